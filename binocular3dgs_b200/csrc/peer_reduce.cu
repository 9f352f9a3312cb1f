// peer_reduce.cu — the data-parallel exchange step as ONE kernel over NVLink peer memory.
//
// training shards by view; every rank renders its own camera and the
// per-Gaussian gradients (92 B/Gaussian at M = 4) are summed over ranks.  The backward kernel
// (K8+K9) already writes its gradients straight into the exchange buffer (Backend.grad_sink);
// here that buffer is SYMMETRIC memory (torch.distributed._symmetric_memory: every rank holds
// the device pointers of all peers' buffers), and the sum is a two-shot all-reduce done in
// place by one kernel: rank r owns the r-th slice of the bucket; for each float4 of its
// slice it loads the N peers' values over NVLink (N independent loads in flight), adds them
// in rank order — every element is reduced by exactly one rank, so all replicas receive
// bit-identical sums — scales, and stores the result into all N buckets.  Traffic per rank:
// (N-1)/N of the bucket in, the same out; no staging buffer, no pack/unpack, no NCCL
// launch/protocol latency.  The two barriers around it (peers' backward stores visible
// before, reduced values landed after) are the symmetric-memory signal-pad barriers issued
// by the host side on the same stream (dp.PeerGradientBucket).
#include <cstdlib>

#include "../../include/b3gs.h"
#include "common.cuh"
#include "kernels.h"

namespace b3 {

struct PeerPtrs {
    float* p[B3GS_MAX_PEERS];
};

template <int N>
__global__ void __launch_bounds__(256) peer_allreduce_kernel(const __grid_constant__ PeerPtrs peers, int rank,
                                                            size_t n4, float scale) {
    // this rank's slice of float4 indices
    const size_t per = (n4 + N - 1) / N;
    const size_t lo = per * rank, hi = lo + per < n4 ? lo + per : n4;
    // kUnroll float4s per thread and iteration: N * kUnroll independent 16-byte NVLink loads in
    // flight per thread before the first add (remote latency is ~2 us)
    constexpr int kUnroll = N <= 2 ? 4 : 2;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t base = lo + t0; base < hi; base += stride * kUnroll) {
        float4 v[kUnroll][N];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const size_t i = base + u * stride;
#pragma unroll
            for (int r = 0; r < N; r++)
                v[u][r] = i < hi ? __ldcg(reinterpret_cast<const float4*>(peers.p[r]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const size_t i = base + u * stride;
            if (i >= hi) break;
            float4 s = v[u][0];
#pragma unroll
            for (int r = 1; r < N; r++) { s.x += v[u][r].x; s.y += v[u][r].y; s.z += v[u][r].z; s.w += v[u][r].w; }
            s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
#pragma unroll
            for (int r = 0; r < N; r++) __stcg(reinterpret_cast<float4*>(peers.p[r]) + i, s);
        }
    }
}

// NVLS variant: `mc` is the MULTICAST address of the same symmetric buffer
// (cuMulticast / NVSwitch).  multimem.ld_reduce returns the sum over all ranks' copies,
// formed inside the switch (one NVLink read instead of N-1), multimem.st broadcasts the
// result to all of them (one write instead of N-1).  The summation order inside the switch is
// unspecified, but every element is still reduced exactly once and then broadcast, so the
// replicas stay bit-identical.
__global__ void __launch_bounds__(256) peer_allreduce_multimem_kernel(float* mc, int world, int rank, size_t n4,
                                                                     float scale) {
    const size_t per = (n4 + world - 1) / world;
    const size_t lo = per * rank, hi = lo + per < n4 ? lo + per : n4;
    constexpr int kUnroll = 4;  // independent switch reductions in flight per thread
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t base = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; base < hi; base += stride * kUnroll) {
        float4 s[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const size_t i = base + u * stride;
            s[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < hi)
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(s[u].x), "=f"(s[u].y), "=f"(s[u].z), "=f"(s[u].w)
                             : "l"(mc + 4 * i)
                             : "memory");
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const size_t i = base + u * stride;
            if (i >= hi) break;
            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + 4 * i),
                         "f"(s[u].x * scale), "f"(s[u].y * scale), "f"(s[u].z * scale), "f"(s[u].w * scale)
                         : "memory");
        }
    }
}

// ---- the same exchange with the two cross-rank barriers INSIDE the kernel ----------------
// Flags live in the symmetric buffer itself, behind the data (64 words): ready[r] / done[r] hold
// the epoch up to which rank r has (a) finished producing its bucket, (b) finished reading and
// writing everybody's.  One launch per step and rank, no host-issued barrier kernels:
//   0. griddepcontrol.wait — with programmatic dependent launch the kernel is already resident
//      while the producer (the backward's last kernel) drains; this is where it waits for it
//   1. block 0 publishes ready[rank] = epoch into every peer's flag area (release, system scope)
//   2. every block waits until ready[r] >= epoch for all r (acquire, system scope)
//   3. the reduction of this rank's slice (plain peer loads/stores, or multimem through the switch)
//   4. the last block to finish publishes done[rank] = epoch to every peer and then waits for
//      done[r] >= epoch from all r: when the kernel ends, every peer's stores into this bucket
//      have landed and every peer has finished reading it (the next step may overwrite it).
// No block waits for a block of the SAME grid that may not be resident (ready flags come from
// block 0, done flags from whichever block finishes last), so any grid size is safe.  Waits are
// bounded: a lost peer becomes a trap (an error on every rank), never a hung GPU.
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_flag(const uint32_t* p, uint32_t epoch) {
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(p) - epoch) < 0) {
        if (clock64() - t0 > (1ll << 33)) __trap();   // ~4 s at 2 GHz
    }
}

constexpr int kFlagWords = 64;   // [0,8) ready, [8,16) done, [16] block counter of this rank

template <int N, bool kMultimem, int kU>
__global__ void __launch_bounds__(256) peer_allreduce_fused_kernel(const __grid_constant__ PeerPtrs peers, float* mc,
                                                                  int world, int rank, size_t n4, size_t flag_off,
                                                                  uint32_t epoch, float scale) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    uint32_t* my_flags = reinterpret_cast<uint32_t*>(peers.p[rank] + flag_off);
    if (blockIdx.x == 0 && threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<uint32_t*>(peers.p[threadIdx.x] + flag_off) + rank, epoch);
    }
    if (threadIdx.x < world) wait_flag(my_flags + threadIdx.x, epoch);
    __syncthreads();

    const size_t per = (n4 + world - 1) / world;
    const size_t lo = per * rank, hi = lo + per < n4 ? lo + per : n4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (kMultimem) {
        constexpr int kUnroll = kU;
        for (size_t base = lo + t0; base < hi; base += stride * kUnroll) {
            float4 s[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; u++) {
                const size_t i = base + u * stride;
                s[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < hi)
                    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(s[u].x), "=f"(s[u].y), "=f"(s[u].z), "=f"(s[u].w)
                                 : "l"(mc + 4 * i)
                                 : "memory");
            }
#pragma unroll
            for (int u = 0; u < kUnroll; u++) {
                const size_t i = base + u * stride;
                if (i >= hi) break;
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + 4 * i),
                             "f"(s[u].x * scale), "f"(s[u].y * scale), "f"(s[u].z * scale), "f"(s[u].w * scale)
                             : "memory");
            }
        }
    } else {
        constexpr int kUnroll = N <= 2 ? kU : (kU > 2 ? kU / 2 : 1);
        for (size_t base = lo + t0; base < hi; base += stride * kUnroll) {
            float4 v[kUnroll][N];
#pragma unroll
            for (int u = 0; u < kUnroll; u++) {
                const size_t i = base + u * stride;
#pragma unroll
                for (int r = 0; r < N; r++)
                    v[u][r] = i < hi ? __ldcg(reinterpret_cast<const float4*>(peers.p[r]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; u++) {
                const size_t i = base + u * stride;
                if (i >= hi) break;
                float4 s = v[u][0];
#pragma unroll
                for (int r = 1; r < N; r++) { s.x += v[u][r].x; s.y += v[u][r].y; s.z += v[u][r].z; s.w += v[u][r].w; }
                s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
#pragma unroll
                for (int r = 0; r < N; r++) __stcg(reinterpret_cast<float4*>(peers.p[r]) + i, s);
            }
        }
    }

    __threadfence_system();   // this thread's stores into the peers' buckets are performed
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
        s_last = atomicAdd(my_flags + 16, 1u) + 1u == gridDim.x;
        if (s_last) my_flags[16] = 0u;   // nobody touches it again before the next call's blocks
    }
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<uint32_t*>(peers.p[threadIdx.x] + flag_off) + 8 + rank, epoch);
        wait_flag(my_flags + 8 + threadIdx.x, epoch);
    }
}

// ---- chunked exchange: the fused kernel over up to five ranges (one chunk of Gaussians) --------
// The flattened index space of the ranges is sliced over the ranks exactly like the whole
// bucket above; same flag protocol (one epoch per chunk call).
template <int N, bool kMultimem>
__global__ void __launch_bounds__(256) peer_allreduce_ranges_kernel(const __grid_constant__ PeerPtrs peers, float* mc,
                                                                   const __grid_constant__ ExchangeRanges rg, int world,
                                                                   int rank, size_t flag_off, uint32_t epoch,
                                                                   float scale) {
    uint32_t* my_flags = reinterpret_cast<uint32_t*>(peers.p[rank] + flag_off);
    if (blockIdx.x == 0 && threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<uint32_t*>(peers.p[threadIdx.x] + flag_off) + rank, epoch);
    }
    if (threadIdx.x < world) wait_flag(my_flags + threadIdx.x, epoch);
    __syncthreads();

    size_t total = 0;
    for (int k = 0; k < rg.n; k++) total += rg.len4[k];
    const size_t per = (total + world - 1) / world;
    const size_t lo = per * rank, hi = lo + per < total ? lo + per : total;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    constexpr int kUnroll = kMultimem ? 4 : (N <= 2 ? 4 : 2);   // independent remote loads in flight per thread
    for (size_t f0 = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; f0 < hi; f0 += stride * kUnroll) {
        size_t idx[kUnroll];
        float4 v[kUnroll][kMultimem ? 1 : N];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            size_t i = f0 + u * stride;
            idx[u] = ~(size_t)0;
            if (i >= hi) continue;
            int k = 0;
            while (i >= rg.len4[k]) { i -= rg.len4[k]; k++; }   // flat index -> (range, offset)
            i += rg.start4[k];
            idx[u] = i;
            if (kMultimem) {
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(v[u][0].x), "=f"(v[u][0].y), "=f"(v[u][0].z), "=f"(v[u][0].w) : "l"(mc + 4 * i) : "memory");
            } else {
#pragma unroll
                for (int r = 0; r < N; r++) v[u][r] = __ldcg(reinterpret_cast<const float4*>(peers.p[r]) + i);
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const size_t i = idx[u];
            if (i == ~(size_t)0) continue;
            float4 s = v[u][0];
            if (!kMultimem) {
#pragma unroll
                for (int r = 1; r < N; r++) { s.x += v[u][r].x; s.y += v[u][r].y; s.z += v[u][r].z; s.w += v[u][r].w; }
            }
            s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
            if (kMultimem) {
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + 4 * i), "f"(s.x),
                             "f"(s.y), "f"(s.z), "f"(s.w) : "memory");
            } else {
#pragma unroll
                for (int r = 0; r < N; r++) __stcg(reinterpret_cast<float4*>(peers.p[r]) + i, s);
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
        s_last = atomicAdd(my_flags + 16, 1u) + 1u == gridDim.x;
        if (s_last) my_flags[16] = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<uint32_t*>(peers.p[threadIdx.x] + flag_off) + 8 + rank, epoch);
        wait_flag(my_flags + 8 + threadIdx.x, epoch);
    }
}

constexpr int kMaxChunks = 16;
struct ExchangePlan {
    int world, rank;
    PeerPtrs peers;
    float* mc;
    size_t n_floats, flag_off;
    unsigned epoch;
    cudaStream_t side;
    cudaEvent_t ev[kMaxChunks + 1];
};
cudaStream_t exchange_stream(ExchangePlan* p) { return p->side; }
cudaEvent_t exchange_event(ExchangePlan* p, int i) { return p->ev[i]; }
float* exchange_base(ExchangePlan* p) { return p->peers.p[p->rank]; }

cudaError_t exchange_chunk(ExchangePlan* p, const ExchangeRanges& r, float scale, cudaStream_t stream) {
    size_t total = 0;
    for (int k = 0; k < r.n; k++) total += r.len4[k];
    if (total == 0) return cudaSuccess;
    size_t blocks = ((total + p->world - 1) / p->world + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks < 1) blocks = 1;
    const unsigned epoch = ++p->epoch;
#define B3_LAUNCH_RANGES(N, MM)                                                                                       \
    peer_allreduce_ranges_kernel<N, MM><<<(unsigned)blocks, 256, 0, stream>>>(p->peers, p->mc, r, p->world, p->rank, \
                                                                             p->flag_off, epoch, scale)
    if (p->mc) {
        B3_LAUNCH_RANGES(1, true);
    } else {
        switch (p->world) {
            case 1: B3_LAUNCH_RANGES(1, false); break;
            case 2: B3_LAUNCH_RANGES(2, false); break;
            case 3: B3_LAUNCH_RANGES(3, false); break;
            case 4: B3_LAUNCH_RANGES(4, false); break;
            case 5: B3_LAUNCH_RANGES(5, false); break;
            case 6: B3_LAUNCH_RANGES(6, false); break;
            case 7: B3_LAUNCH_RANGES(7, false); break;
            default: B3_LAUNCH_RANGES(8, false); break;
        }
    }
#undef B3_LAUNCH_RANGES
    count_launch();
    return cudaGetLastError();
}

}  // namespace b3

using namespace b3;

// The exchange plan: everything the chunked exchange needs, created once per bucket.  `epoch_base`:
// the epoch the flag words currently hold (0 for a fresh buffer); the plan and
// b3gs_peer_allreduce_fused must not be mixed on one buffer without keeping the epochs in step
// (b3gs_exchange_epoch / the epoch argument).
extern "C" int b3gs_exchange_create(int world, int rank, float* const* peer_buffers, float* multicast_buffer,
                                    size_t n_floats, size_t flag_off_floats, void** handle_out) {
    if (world < 1 || world > B3GS_MAX_PEERS || rank < 0 || rank >= world || !peer_buffers || !handle_out ||
        (n_floats & 3) || (flag_off_floats & 3) || flag_off_floats < n_floats)
        return -1;
    ExchangePlan* p = new ExchangePlan();
    p->world = world; p->rank = rank; p->mc = multicast_buffer; p->n_floats = n_floats; p->flag_off = flag_off_floats;
    p->epoch = 0;
    for (int r = 0; r < B3GS_MAX_PEERS; r++) p->peers.p[r] = r < world ? peer_buffers[r] : nullptr;
    if (cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking) != cudaSuccess) { delete p; return -2; }
    for (int i = 0; i <= kMaxChunks; i++)
        if (cudaEventCreateWithFlags(&p->ev[i], cudaEventDisableTiming) != cudaSuccess) { delete p; return -2; }
    *handle_out = p;
    return 0;
}
extern "C" void b3gs_exchange_destroy(void* handle) {
    ExchangePlan* p = static_cast<ExchangePlan*>(handle);
    if (!p) return;
    for (int i = 0; i <= kMaxChunks; i++) cudaEventDestroy(p->ev[i]);
    cudaStreamDestroy(p->side);
    delete p;
}
extern "C" unsigned b3gs_exchange_epoch(void* handle, int set, unsigned value) {
    ExchangePlan* p = static_cast<ExchangePlan*>(handle);
    if (set) p->epoch = value;
    return p->epoch;
}

// n_floats: length of the DATA (a multiple of 4); the buffers must extend kFlagWords 32-bit words
// beyond flag_off_floats (>= n_floats, a multiple of 4), zeroed once when the buffer is created.
// epoch: 1, 2, 3, ... — the same on every rank, one more at every call.
extern "C" int b3gs_peer_allreduce_fused(int world, int rank, float* const* peer_buffers, float* multicast_buffer,
                                         size_t n_floats, size_t flag_off_floats, unsigned epoch, float scale,
                                         void* stream) {
    if (world < 1 || world > B3GS_MAX_PEERS || rank < 0 || rank >= world || !peer_buffers || (n_floats & 3) ||
        (flag_off_floats & 3) || flag_off_floats < n_floats || epoch == 0)
        return -1;
    PeerPtrs pp;
    for (int r = 0; r < B3GS_MAX_PEERS; r++) pp.p[r] = r < world ? peer_buffers[r] : nullptr;
    for (int r = 0; r < world; r++)
        if (!pp.p[r] || (reinterpret_cast<uintptr_t>(pp.p[r]) & 15)) return -1;
    const size_t n4 = n_floats / 4;
    // tuning knobs (read per call so that one process can sweep them): independent 16-byte
    // requests in flight per thread, resident blocks per SM
    const char* eu = getenv("B3GS_AR_UNROLL");
    const char* eb = getenv("B3GS_AR_BPSM");
    // Through the switch (multimem) FEWER requests in flight are faster — 8x B200, 92 MB bucket:
    // 2 per thread x 2 blocks/SM 239 us, 4 x 4 272 us, 8 x 8 277 us (profiles/README.md r02m);
    // plain peer loads want more (2x B200: 4 x 4).
    const int unroll = eu ? atoi(eu) : (multicast_buffer ? 2 : 4);
    const int bpsm = eb && atoi(eb) > 0 ? atoi(eb) : (multicast_buffer ? 2 : 4);
    size_t blocks = ((n4 + world - 1) / world + 255) / 256;
    if (blocks > (size_t)148 * bpsm) blocks = (size_t)148 * bpsm;
    if (blocks < 1) blocks = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)blocks); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0;
    cfg.stream = reinterpret_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // resident while the producer drains
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e;
#define B3_LAUNCH_FUSED_U(N, MM, U) \
    e = cudaLaunchKernelEx(&cfg, peer_allreduce_fused_kernel<N, MM, U>, pp, multicast_buffer, world, rank, n4, flag_off_floats, \
                           (uint32_t)epoch, scale)
#define B3_LAUNCH_FUSED(N, MM)                                   \
    do {                                                         \
        if (unroll >= 8) B3_LAUNCH_FUSED_U(N, MM, 8);            \
        else if (unroll <= 2) B3_LAUNCH_FUSED_U(N, MM, 2);       \
        else B3_LAUNCH_FUSED_U(N, MM, 4);                        \
    } while (0)
    if (multicast_buffer) {
        B3_LAUNCH_FUSED(1, true);
    } else {
        switch (world) {
            case 1: B3_LAUNCH_FUSED(1, false); break;
            case 2: B3_LAUNCH_FUSED(2, false); break;
            case 3: B3_LAUNCH_FUSED(3, false); break;
            case 4: B3_LAUNCH_FUSED(4, false); break;
            case 5: B3_LAUNCH_FUSED(5, false); break;
            case 6: B3_LAUNCH_FUSED(6, false); break;
            case 7: B3_LAUNCH_FUSED(7, false); break;
            default: B3_LAUNCH_FUSED(8, false); break;
        }
    }
#undef B3_LAUNCH_FUSED
#undef B3_LAUNCH_FUSED_U
    count_launch();
    return e == cudaSuccess ? 0 : -2;
}

extern "C" int b3gs_peer_allreduce_multimem(int world, int rank, float* multicast_buffer, size_t n_floats, float scale,
                                            void* stream) {
    if (world < 1 || rank < 0 || rank >= world || !multicast_buffer || (n_floats & 3) ||
        (reinterpret_cast<uintptr_t>(multicast_buffer) & 15))
        return -1;
    const size_t n4 = n_floats / 4;
    if (n4 == 0) return 0;
    size_t blocks = ((n4 + world - 1) / world + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks < 1) blocks = 1;
    peer_allreduce_multimem_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        multicast_buffer, world, rank, n4, scale);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

extern "C" int b3gs_peer_allreduce(int world, int rank, float* const* peer_buffers, size_t n_floats, float scale,
                                   void* stream) {
    if (world < 1 || world > B3GS_MAX_PEERS || rank < 0 || rank >= world || !peer_buffers || (n_floats & 3)) return -1;
    PeerPtrs pp;
    for (int r = 0; r < B3GS_MAX_PEERS; r++) pp.p[r] = r < world ? peer_buffers[r] : nullptr;
    for (int r = 0; r < world; r++)
        if (!pp.p[r] || (reinterpret_cast<uintptr_t>(pp.p[r]) & 15)) return -1;
    const size_t n4 = n_floats / 4;
    if (n4 == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // one slice per rank: size the grid for the slice, capped at 148 SMs x 8 blocks
    size_t blocks = ((n4 + world - 1) / world + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks < 1) blocks = 1;
    switch (world) {
        case 1: peer_allreduce_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(pp, rank, n4, scale); break;
        case 2: peer_allreduce_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(pp, rank, n4, scale); break;
        case 3: peer_allreduce_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(pp, rank, n4, scale); break;
        case 4: peer_allreduce_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(pp, rank, n4, scale); break;
        case 5: peer_allreduce_kernel<5><<<(unsigned)blocks, 256, 0, st>>>(pp, rank, n4, scale); break;
        case 6: peer_allreduce_kernel<6><<<(unsigned)blocks, 256, 0, st>>>(pp, rank, n4, scale); break;
        case 7: peer_allreduce_kernel<7><<<(unsigned)blocks, 256, 0, st>>>(pp, rank, n4, scale); break;
        default: peer_allreduce_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(pp, rank, n4, scale); break;
    }
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
