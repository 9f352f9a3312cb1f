// knn.cu — distCUDA2: mean squared distance to the 3 nearest neighbours of every point
// (reference submodules/simple-knn/simple_knn.cu:185-221, called
// once per training run at scene/gaussian_model.py:134 to initialise the Gaussian scales).
//
// The result is an exact geometric quantity — (d1^2 + d2^2 + d3^2) / 3 over the true three
// nearest neighbours — so any exact search reproduces the reference bit for bit as long as
// the distance expression is evaluated the way its SASS does:
//     d = neighbour - query;  dist = fma(d.z, d.z, fma(d.x, d.x, d.y * d.y));   (the MIDDLE
//     product is rounded alone — the same contraction as common.cuh dot3)
//     out = ((best0 + best1) + best2) / 3.0f   (IEEE division)
// Missing neighbours (P < 4) stay at FLT_MAX as in the reference (the sum overflows to inf).
//
// Reference search: Morton sort, boxes of 1024 consecutive points, one THREAD per query
// walking every box and brute-forcing 1024 points of each box it cannot reject.
// Here: the same Morton order, but a two-level hierarchy (leaves of 32 points = one warp,
// super boxes of 32 leaves) and one WARP per leaf of 32 consecutive queries:
//   * the queries' own leaf is scanned first, which gives every lane a tight bound;
//   * super boxes and leaves are tested 32 at a time (one per lane) against the warp's
//     query box inflated by the worst lane's current 3rd-best distance -> ballots;
//   * an accepted leaf is loaded once (one coalesced 512-byte read), staged in shared
//     memory and broadcast to all lanes with LDS.128; every lane updates its own 3-best.
// No host synchronisation (the reference copies the bounding box to the host twice), no
// device allocation (scratch comes from the caller).
#include <cfloat>

#include "../../include/b3gs.h"
#include "common.cuh"
#include "kernels.h"

namespace b3 {

constexpr int kLeaf = 32, kSuper = 32 * kLeaf;  // points per leaf / per super box
constexpr int kSearchWarps = 4;

struct KnnBox {
    float4 lo, hi;  // xyz used
};

static size_t knn_align(size_t v) { return (v + 255) / 256 * 256; }

// order-preserving float <-> uint map for atomicMin/atomicMax
__device__ __forceinline__ uint32_t f2ord(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// bounds[0..2] = min xyz, bounds[3..5] = max xyz (ordered encoding; initialised by the host
// side to 0xffffffff / 0)
__global__ void __launch_bounds__(256) knn_bounds_kernel(int P, const float* __restrict__ points,
                                                        uint32_t* __restrict__ bounds) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float v = points[3 * (size_t)i + c];
            lo[c] = fminf(lo[c], v);
            hi[c] = fmaxf(hi[c], v);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], d));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], d));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            atomicMin(bounds + c, f2ord(lo[c]));
            atomicMax(bounds + 3 + c, f2ord(hi[c]));
        }
    }
}

__device__ __forceinline__ uint32_t spread10(uint32_t x) {  // 10 bits -> every third bit
    x = (x | (x << 16)) & 0x030000FFu;
    x = (x | (x << 8)) & 0x0300F00Fu;
    x = (x | (x << 4)) & 0x030C30C3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}

// 60-bit Morton code (20 bits per axis) of each point inside the cloud's bounding box, as
// two 30-bit halves.  The reference uses 10 bits per axis (simple_knn.cu:46-73); with SfM
// clouds a few far outliers stretch the box and whole dense clusters collapse into a handful
// of 10-bit cells, inside which the order — and so the leaves — would be arbitrary.
__global__ void __launch_bounds__(256) knn_morton_kernel(int P, const float* __restrict__ points,
                                                        const uint32_t* __restrict__ bounds,
                                                        uint32_t* __restrict__ codes_lo,
                                                        uint32_t* __restrict__ codes_hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    uint32_t lo30 = 0, hi30 = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float lo = ord2f(bounds[c]), hi = ord2f(bounds[3 + c]);
        const float ext = hi - lo;
        const float t = ext > 0.f ? (points[3 * (size_t)i + c] - lo) / ext : 0.f;
        const uint32_t q = (uint32_t)fminf(fmaxf(t * 1048575.f, 0.f), 1048575.f);
        lo30 |= spread10(q & 1023u) << c;
        hi30 |= spread10(q >> 10) << c;
    }
    codes_lo[i] = lo30;
    codes_hi[i] = hi30;
}

__global__ void __launch_bounds__(256) knn_gather_kernel(int P, const uint32_t* __restrict__ src,
                                                        const uint32_t* __restrict__ idx, uint32_t* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) dst[i] = src[idx[i]];
}

// One block = one super box of 1024 Morton-consecutive points = 32 leaves (one per warp).
// Gathers the points into Morton order (float4: xyz + original index bits) and reduces the
// leaf and super boxes.  Positions >= P are padded with +inf points and neutral boxes.
__global__ void __launch_bounds__(kSuper) knn_leaf_kernel(int P, const float* __restrict__ points,
                                                         const uint32_t* __restrict__ ids_by_lo,
                                                         const uint32_t* __restrict__ perm_by_hi,
                                                         float4* __restrict__ sorted_pts,
                                                         KnnBox* __restrict__ leaf_boxes,
                                                         KnnBox* __restrict__ super_boxes) {
    __shared__ KnnBox s_leaf[32];
    const int pos = blockIdx.x * kSuper + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    float4 p = make_float4(INFINITY, INFINITY, INFINITY, 0.f);
    if (pos < P) {
        const uint32_t id = ids_by_lo[perm_by_hi[pos]];  // LSD: low half first, then stable by high half
        p = make_float4(points[3 * (size_t)id], points[3 * (size_t)id + 1], points[3 * (size_t)id + 2],
                        __uint_as_float(id));
        lo[0] = hi[0] = p.x; lo[1] = hi[1] = p.y; lo[2] = hi[2] = p.z;
    }
    sorted_pts[pos] = p;
#pragma unroll
    for (int c = 0; c < 3; c++) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], d));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], d));
        }
    }
    KnnBox b;
    b.lo = make_float4(lo[0], lo[1], lo[2], 0.f);
    b.hi = make_float4(hi[0], hi[1], hi[2], 0.f);
    if (lane == 0) {
        leaf_boxes[blockIdx.x * 32 + warp] = b;
        s_leaf[warp] = b;
    }
    __syncthreads();
    if (warp == 0) {
        b = s_leaf[lane];
        float l[3] = {b.lo.x, b.lo.y, b.lo.z}, h[3] = {b.hi.x, b.hi.y, b.hi.z};
#pragma unroll
        for (int c = 0; c < 3; c++) {
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                l[c] = fminf(l[c], __shfl_xor_sync(0xffffffffu, l[c], d));
                h[c] = fmaxf(h[c], __shfl_xor_sync(0xffffffffu, h[c], d));
            }
        }
        if (lane == 0) {
            KnnBox s;
            s.lo = make_float4(l[0], l[1], l[2], 0.f);
            s.hi = make_float4(h[0], h[1], h[2], 0.f);
            super_boxes[blockIdx.x] = s;
        }
    }
}

// squared distance between two axis-aligned boxes (0 if they overlap); a lower bound of the
// distance between any point of one and any point of the other.  Rounded DOWN-safe: each
// gap is an exact float subtraction result or smaller than the true gap by at most one
// rounding, so the bound is scaled by (1 - 2^-20) before the comparison at the call site.
__device__ __forceinline__ float box_box_dist2(const KnnBox& a, const KnnBox& q) {
    const float gx = fmaxf(fmaxf(a.lo.x - q.hi.x, q.lo.x - a.hi.x), 0.f);
    const float gy = fmaxf(fmaxf(a.lo.y - q.hi.y, q.lo.y - a.hi.y), 0.f);
    const float gz = fmaxf(fmaxf(a.lo.z - q.hi.z, q.lo.z - a.hi.z), 0.f);
    return (gx * gx + gy * gy + gz * gz) * 0.999999f;
}

// simple_knn.cu:132-146 updateKBest<3> with the reference's distance expression
__device__ __forceinline__ void update3(float qx, float qy, float qz, float px, float py, float pz, float& b0, float& b1,
                                        float& b2) {
    const float dx = __fsub_rn(px, qx), dy = __fsub_rn(py, qy), dz = __fsub_rn(pz, qz);
    float dist = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
    if (b0 > dist) { const float t = b0; b0 = dist; dist = t; }
    if (b1 > dist) { const float t = b1; b1 = dist; dist = t; }
    if (b2 > dist) { b2 = dist; }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}

__global__ void __launch_bounds__(kSearchWarps * 32) knn_search_kernel(int P, int n_leaves, int n_super,
                                                                     const float4* __restrict__ sorted_pts,
                                                                     const KnnBox* __restrict__ leaf_boxes,
                                                                     const KnnBox* __restrict__ super_boxes,
                                                                     float* __restrict__ mean_dist2) {
    __shared__ float4 s_pts[kSearchWarps][kLeaf];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int leaf = blockIdx.x * kSearchWarps + warp;
    if (leaf >= n_leaves) return;  // whole warp
    const int pos = leaf * kLeaf + lane;
    const bool valid = pos < P;
    const float4 q = sorted_pts[pos];
    float b0 = FLT_MAX, b1 = FLT_MAX, b2 = FLT_MAX;
    float4* sp = s_pts[warp];

    // own leaf first: all-to-all inside the warp
    sp[lane] = q;
    __syncwarp();
#pragma unroll 8
    for (int j = 0; j < kLeaf; j++) {
        const float4 p = sp[j];
        if (j != lane) update3(q.x, q.y, q.z, p.x, p.y, p.z, b0, b1, b2);
    }
    __syncwarp();
    // then the Morton neighbours on both sides (the reference seeds its bound from the
    // +-3 neighbours in Morton order): without them a nearly empty last leaf would start
    // with an infinite bound and walk the whole cloud
    for (int l = leaf - 1; l <= leaf + 1; l += 2) {
        if (l < 0 || l >= n_leaves) continue;
        sp[lane] = sorted_pts[l * kLeaf + lane];
        __syncwarp();
#pragma unroll 8
        for (int j = 0; j < kLeaf; j++) {
            const float4 p = sp[j];
            update3(q.x, q.y, q.z, p.x, p.y, p.z, b0, b1, b2);
        }
        __syncwarp();
    }
    const KnnBox qbox = leaf_boxes[leaf];
    float bound = warp_max(valid ? b2 : 0.f);

    for (int sb0 = 0; sb0 < n_super; sb0 += 32) {
        const int sb = sb0 + lane;
        bool acc = false;
        if (sb < n_super) acc = box_box_dist2(super_boxes[sb], qbox) <= bound;
        unsigned m = __ballot_sync(0xffffffffu, acc);
        while (m) {
            const int s = sb0 + __ffs(m) - 1;
            m &= m - 1;
            const int lf = s * 32 + lane;
            bool acc2 = false;
            if (lf < n_leaves && (lf < leaf - 1 || lf > leaf + 1)) acc2 = box_box_dist2(leaf_boxes[lf], qbox) <= bound;
            unsigned m2 = __ballot_sync(0xffffffffu, acc2);
            while (m2) {
                const int l = s * 32 + __ffs(m2) - 1;
                m2 &= m2 - 1;
                // exact per-query test (simple_knn.cu:118-130 distBoxPoint) against each lane's
                // own current 3rd-best: scan only if some query can still improve
                const KnnBox lb = leaf_boxes[l];
                const float gx = fmaxf(fmaxf(lb.lo.x - q.x, q.x - lb.hi.x), 0.f);
                const float gy = fmaxf(fmaxf(lb.lo.y - q.y, q.y - lb.hi.y), 0.f);
                const float gz = fmaxf(fmaxf(lb.lo.z - q.z, q.z - lb.hi.z), 0.f);
                const bool want = valid && (gx * gx + gy * gy + gz * gz) * 0.999999f <= b2;
                if (!__any_sync(0xffffffffu, want)) continue;
                sp[lane] = sorted_pts[l * kLeaf + lane];
                __syncwarp();
#pragma unroll 8
                for (int j = 0; j < kLeaf; j++) {
                    const float4 p = sp[j];
                    update3(q.x, q.y, q.z, p.x, p.y, p.z, b0, b1, b2);
                }
                __syncwarp();
                bound = warp_max(valid ? b2 : 0.f);
            }
        }
    }
    if (valid) mean_dist2[__float_as_uint(q.w)] = __fdiv_rn(__fadd_rn(__fadd_rn(b0, b1), b2), 3.0f);
}

}  // namespace b3

using namespace b3;

extern "C" {

size_t b3gs_dist_cuda2_scratch_bytes(int P) {
    if (P <= 0) return 256;
    const size_t n_super = ((size_t)P + kSuper - 1) / kSuper;
    return knn_align(6 * 4) + 3 * knn_align((size_t)P * 4) + 2 * knn_align(sort_keys_iota_scratch_bytes(P)) +
           knn_align(n_super * kSuper * sizeof(float4)) + knn_align(n_super * 32 * sizeof(KnnBox)) +
           knn_align(n_super * sizeof(KnnBox));
}

int b3gs_dist_cuda2(int P, const float* points, float* mean_dist2, void* scratch, size_t scratch_bytes, void* stream) {
    if (P < 0 || (P > 0 && (!points || !mean_dist2 || !scratch))) return -1;
    if (P == 0) return 0;
    if (scratch_bytes < b3gs_dist_cuda2_scratch_bytes(P)) return -1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int n_super = (P + kSuper - 1) / kSuper, n_leaves = (P + kLeaf - 1) / kLeaf;
    char* q = reinterpret_cast<char*>(scratch);
    uint32_t* bounds = reinterpret_cast<uint32_t*>(q); q += knn_align(6 * 4);
    uint32_t* codes_lo = reinterpret_cast<uint32_t*>(q); q += knn_align((size_t)P * 4);
    uint32_t* codes_hi = reinterpret_cast<uint32_t*>(q); q += knn_align((size_t)P * 4);
    uint32_t* hi_by_lo = reinterpret_cast<uint32_t*>(q); q += knn_align((size_t)P * 4);
    char* sort_scratch1 = q;                           q += knn_align(sort_keys_iota_scratch_bytes(P));
    char* sort_scratch2 = q;                           q += knn_align(sort_keys_iota_scratch_bytes(P));
    float4* sorted_pts = reinterpret_cast<float4*>(q); q += knn_align((size_t)n_super * kSuper * sizeof(float4));
    KnnBox* leaf_boxes = reinterpret_cast<KnnBox*>(q); q += knn_align((size_t)n_super * 32 * sizeof(KnnBox));
    KnnBox* super_boxes = reinterpret_cast<KnnBox*>(q);

    if (cudaMemsetAsync(bounds, 0xff, 12, st) != cudaSuccess) return -2;
    if (cudaMemsetAsync(bounds + 3, 0, 12, st) != cudaSuccess) return -2;
    int gb = (P + 255) / 256;
    if (gb > 148 * 8) gb = 148 * 8;
    knn_bounds_kernel<<<gb, 256, 0, st>>>(P, points, bounds);
    knn_morton_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, points, bounds, codes_lo, codes_hi);
    count_launch(2);
    const uint32_t *keys_sorted = nullptr, *ids_by_lo = nullptr, *perm_by_hi = nullptr;
    if (sort_keys_iota_u32(P, codes_lo, 30, sort_scratch1, &keys_sorted, &ids_by_lo, st) != cudaSuccess) return -2;
    knn_gather_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, codes_hi, ids_by_lo, hi_by_lo);
    if (sort_keys_iota_u32(P, hi_by_lo, 30, sort_scratch2, &keys_sorted, &perm_by_hi, st) != cudaSuccess) return -2;
    knn_leaf_kernel<<<n_super, kSuper, 0, st>>>(P, points, ids_by_lo, perm_by_hi, sorted_pts, leaf_boxes, super_boxes);
    count_launch();
    knn_search_kernel<<<(n_leaves + kSearchWarps - 1) / kSearchWarps, kSearchWarps * 32, 0, st>>>(
        P, n_leaves, n_super, sorted_pts, leaf_boxes, super_boxes, mean_dist2);
    count_launch(2);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // extern "C"
