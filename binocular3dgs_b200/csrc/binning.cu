// binning.cu — K2..K5: prefix sum, (tile|depth) key emission, sort, tile ranges.
//
// Reference behaviour: rasterizer_impl.cu:70-111 (duplicateWithKeys), :116-138
// (identifyTileRanges), :35-50 (getHigherMsb), :278 (InclusiveSum), :304-309
// (SortPairs on bits [0, 32+bit)), :311 (memset ranges).
//
// Output contract (bit-exact): point_list[R] ordered by (tile id, float_bits(depth),
// Gaussian index) and ranges[T] = [start,end) of each tile in that list, (0,0) for
// empty tiles.
//
// Round-1 state: the prefix sum, key emission and range detection are hand-written;
// the 64-bit key sort still calls cub::DeviceRadixSort (library code, same call the
// reference makes) — its replacement by the depth-presorted two-pass tile sort
// described in DESIGN.md is the next step on this file.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "kernels.h"

namespace b3 {

// ------------------------------------------------------------------ prefix sum
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;  // 2048

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += n;
    }
    return v;
}

// Block-wide inclusive scan of one value per thread; returns the inclusive value,
// `total` = block sum.
__device__ __forceinline__ uint32_t block_inclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = warp_inclusive_scan(v, lane);
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < (int)(blockDim.x >> 5)) ? s_warp[lane] : 0;
        w = warp_inclusive_scan(w, lane);
        s_warp[lane] = w;
    }
    __syncthreads();
    uint32_t prefix = warp > 0 ? s_warp[warp - 1] : 0;
    total = s_warp[(blockDim.x >> 5) - 1];
    return inc + prefix;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const uint32_t* __restrict__ in,
                                                              uint32_t* __restrict__ sums, int n) {
    __shared__ uint32_t s_warp[32];
    const int base = blockIdx.x * kScanTile;
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        int j = base + i * kScanThreads + threadIdx.x;
        if (j < n) acc += in[j];
    }
    uint32_t total;
    block_inclusive_scan(acc, s_warp, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// Single block: exclusive scan of the tile sums in place.
__global__ void __launch_bounds__(1024) scan_sums_exclusive(uint32_t* __restrict__ sums, int m) {
    __shared__ uint32_t s_warp[32];
    uint32_t carry = 0;
    for (int base = 0; base < m; base += 1024) {
        int j = base + threadIdx.x;
        uint32_t v = j < m ? sums[j] : 0;
        uint32_t total;
        uint32_t inc = block_inclusive_scan(v, s_warp, total);
        if (j < m) sums[j] = carry + inc - v;
        carry += total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply(const uint32_t* __restrict__ in,
                                                          const uint32_t* __restrict__ sums,
                                                          uint32_t* __restrict__ out, int n) {
    __shared__ uint32_t s_warp[32];
    // blocked arrangement: thread t owns items [t*8, t*8+8) of the tile
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        acc += v[i];
    }
    uint32_t total;
    uint32_t inc = block_inclusive_scan(acc, s_warp, total);
    uint32_t run = sums[blockIdx.x] + inc - acc;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        run += v[i];
        if (base + i < n) out[base + i] = run;
    }
}

size_t scan_scratch_elems(int P) { return (size_t)((P + kScanTile - 1) / kScanTile) + 1; }

void launch_inclusive_scan(const uint32_t* in, uint32_t* out, uint32_t* block_sums, int P, cudaStream_t stream) {
    const int tiles = (P + kScanTile - 1) / kScanTile;
    scan_tile_sums<<<tiles, kScanThreads, 0, stream>>>(in, block_sums, P);
    scan_sums_exclusive<<<1, 1024, 0, stream>>>(block_sums, tiles);
    scan_apply<<<tiles, kScanThreads, 0, stream>>>(in, block_sums, out, P);
    count_launch(3);
}

// ------------------------------------------------------------------ key emission
// One (key,value) per (Gaussian, tile) overlap, rows then columns of the rectangle
// (rasterizer_impl.cu:98-108).  key = tile_id << 32 | float_bits(depth).
__global__ void __launch_bounds__(256) emit_keys(int P, const float4* __restrict__ records,
                                                const float* __restrict__ depths, const int* __restrict__ radii,
                                                const uint32_t* __restrict__ offsets, uint64_t* __restrict__ keys,
                                                uint32_t* __restrict__ values, int grid_x, int grid_y) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const int r = radii[idx];
    if (r <= 0) return;
    uint32_t off = idx == 0 ? 0u : offsets[idx - 1];
    const float4 a = records[(size_t)idx * B3_REC_VEC4];
    int x0, y0, x1, y1;
    tile_rect(a.x, a.y, (float)r, grid_x, grid_y, x0, y0, x1, y1);
    const uint64_t dbits = __float_as_uint(depths[idx]);
    for (int y = y0; y < y1; y++) {
        for (int x = x0; x < x1; x++) {
            uint64_t key = (uint64_t)(uint32_t)(y * grid_x + x);
            key = (key << 32) | dbits;
            keys[off] = key;
            values[off] = (uint32_t)idx;
            off++;
        }
    }
}

__global__ void __launch_bounds__(256) tile_ranges(int R, const uint64_t* __restrict__ keys,
                                                  uint2* __restrict__ ranges) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R) return;
    const uint32_t cur = (uint32_t)(keys[idx] >> 32);
    if (idx == 0) {
        ranges[cur].x = 0;
    } else {
        const uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
        if (cur != prev) {
            ranges[prev].y = idx;
            ranges[cur].x = idx;
        }
    }
    if (idx == R - 1) ranges[cur].y = R;
}

// rasterizer_impl.cu:35-50
static uint32_t higher_msb(uint32_t n) {
    uint32_t msb = sizeof(n) * 4;
    uint32_t step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t cub_sort_temp_bytes(int R) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, R);
    return bytes;
}

size_t binning_scratch_bytes(int R) {
    // keys_unsorted u64[R] | keys_sorted u64[R] | values_unsorted u32[R] | cub temp
    size_t r = (size_t)(R > 0 ? R : 0);
    return align_up(r * 8, 256) * 2 + align_up(r * 4, 256) + align_up(cub_sort_temp_bytes(R), 256) + 256;
}

cudaError_t run_binning(const BinningArgs& a, cudaStream_t stream) {
    const int T = a.grid_x * a.grid_y;
    cudaError_t e = cudaMemsetAsync(a.ranges, 0, (size_t)T * sizeof(uint2), stream);
    if (e != cudaSuccess) return e;
    if (a.R <= 0) return cudaSuccess;
    const size_t r = (size_t)a.R;
    char* p = a.scratch;
    uint64_t* keys_unsorted = reinterpret_cast<uint64_t*>(p); p += align_up(r * 8, 256);
    uint64_t* keys_sorted = reinterpret_cast<uint64_t*>(p);   p += align_up(r * 8, 256);
    uint32_t* values_unsorted = reinterpret_cast<uint32_t*>(p); p += align_up(r * 4, 256);
    size_t temp_bytes = cub_sort_temp_bytes(a.R);
    void* temp = p;

    emit_keys<<<(a.P + 255) / 256, 256, 0, stream>>>(a.P, a.records, a.depths, a.radii, a.point_offsets,
                                                    keys_unsorted, values_unsorted, a.grid_x, a.grid_y);
    count_launch();
    const int bit = (int)higher_msb((uint32_t)T);
    e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_unsorted, keys_sorted, values_unsorted,
                                        a.point_list, a.R, 0, 32 + bit, stream);
    if (e != cudaSuccess) return e;
    count_launch(8);  // histogram + onesweep passes (library kernels)
    tile_ranges<<<(a.R + 255) / 256, 256, 0, stream>>>(a.R, keys_sorted, a.ranges);
    count_launch();
    return cudaGetLastError();
}

}  // namespace b3
