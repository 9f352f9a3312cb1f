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

}  // namespace b3

using namespace b3;

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
