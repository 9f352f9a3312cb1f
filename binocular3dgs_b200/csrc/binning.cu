// binning.cu — K2..K5: depth order, (tile, id) instance emission, tile sort, tile ranges.
//
// Reference behaviour: rasterizer_impl.cu:70-111 (duplicateWithKeys), :116-138
// (identifyTileRanges), :35-50 (getHigherMsb), :278 (cub InclusiveSum), :304-309
// (cub SortPairs of 64-bit keys on bits [0, 32+bit)), :311 (memset ranges).
//
// Output contract (bit-exact): point_list[R] ordered by (tile id, float_bits(depth),
// Gaussian index) and ranges[T] = [start,end) of each tile in that list, (0,0) for
// empty tiles.  The reference gets it from one stable LSD radix sort of R 64-bit keys
// (6 onesweep passes, ~152 B/instance).  Any algorithm producing that total order is
// equivalent, so this file uses the structure of the key instead:
//
//   phase 1 (P-sized, runs while the host waits for R):
//     stable LSD sort of the P Gaussians by float_bits(depth)        4 x 8-bit passes
//     (fallback path only) exclusive scan of tiles_touched in that order -> emission offsets
//   phase 2 (R-sized), default — direct tile binning, ONE pass over the instances:
//     count   per (batch of depth-consecutive Gaussians, tile) instance counts     table[nb][T]
//     scan    along the batches per tile, then over the tiles                      -> ranges
//     scatter every batch walks its Gaussians in depth order, band of tile rows by band, with the
//             band's counters in shared memory; the ids are staged in a block-local tile-major
//             buffer and copied out coalesced: 4 B written per instance, nothing else R-sized
//   phase 2, fallback (tile rows wider than kBandTilesMax; B3GS_BINNING=radix):
//     emit (tile, id) instances in depth order, one warp per 32 Gaussians
//     stable LSD sort of the instances by tile id                     ceil(bits(T)/8) passes
//     tile ranges from the sorted tile ids
//
// A stable sort by tile of a sequence already ordered by (depth bits, index) IS the
// order (tile, depth bits, index).  The R-sized work drops from 6 passes over 12-byte
// pairs to 2 passes over 8-byte pairs.  All passes are hand-written: per-block digit
// histograms, a row scan, and a rank-and-scatter kernel that ranks with
// __match_any_sync, reorders through shared memory and writes coalesced runs.  No
// spin-waiting anywhere (three plain kernels per pass), so a bug cannot hang the GPU.
//
// The cub path the first revision used is kept behind B3GS_BINNING=cub for A/B timing.
#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace b3 {

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ------------------------------------------------------------------ block scans
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

// Block-wide inclusive scan of one value per thread (blockDim.x multiple of 32, <= 1024).
__device__ __forceinline__ uint32_t block_inclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    uint32_t inc = warp_inclusive_scan(v, lane);
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < nwarps) ? s_warp[lane] : 0;
        w = warp_inclusive_scan(w, lane);
        s_warp[lane] = w;
    }
    __syncthreads();
    uint32_t prefix = warp > 0 ? s_warp[warp - 1] : 0;
    total = s_warp[nwarps - 1];
    __syncthreads();  // s_warp may be reused by the caller
    return inc + prefix;
}

// ------------------------------------------------------------------ exclusive scan with gather
// out[i] = sum_{j<i} in[idx[j]]  (idx == nullptr: in[j]); total written to *total_out.
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const uint32_t* __restrict__ in,
                                                              const uint32_t* __restrict__ idx,
                                                              uint32_t* __restrict__ sums, int n) {
    __shared__ uint32_t s_warp[32];
    const int base = blockIdx.x * kScanTile;
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        int j = base + i * kScanThreads + threadIdx.x;
        if (j < n) acc += idx ? in[idx[j]] : in[j];
    }
    uint32_t total;
    block_inclusive_scan(acc, s_warp, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_sums_exclusive(uint32_t* __restrict__ sums, int m,
                                                           uint32_t* __restrict__ total_out) {
    __shared__ uint32_t s_warp[32];
    uint32_t carry = 0;
    for (int base = 0; base < m; base += 1024) {
        int j = base + threadIdx.x;
        uint32_t v = j < m ? sums[j] : 0;
        uint32_t total;
        uint32_t inc = block_inclusive_scan(v, s_warp, total);
        if (j < m) sums[j] = carry + inc - v;
        carry += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_exclusive(const uint32_t* __restrict__ in,
                                                                    const uint32_t* __restrict__ idx,
                                                                    const uint32_t* __restrict__ sums,
                                                                    uint32_t* __restrict__ out, int n) {
    __shared__ uint32_t s_warp[32];
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;  // blocked arrangement
    uint32_t v[kScanItems];
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        v[i] = (base + i < n) ? (idx ? in[idx[base + i]] : in[base + i]) : 0;
        acc += v[i];
    }
    uint32_t total;
    uint32_t inc = block_inclusive_scan(acc, s_warp, total);
    uint32_t run = sums[blockIdx.x] + inc - acc;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
}

static size_t scan_tiles(int n) { return (size_t)((n + kScanTile - 1) / kScanTile); }

static void exclusive_scan_gather(const uint32_t* in, const uint32_t* idx, uint32_t* out, uint32_t* sums,
                                  uint32_t* total_out, int n, cudaStream_t stream) {
    const int tiles = (int)scan_tiles(n);
    scan_tile_sums<<<tiles, kScanThreads, 0, stream>>>(in, idx, sums, n);
    scan_sums_exclusive<<<1, 1024, 0, stream>>>(sums, tiles, total_out);
    scan_apply_exclusive<<<tiles, kScanThreads, 0, stream>>>(in, idx, sums, out, n);
    count_launch(3);
}

// ------------------------------------------------------------------ radix pass (8-bit digits)
constexpr int kRadixThreads = 256;
constexpr int kRadixItems = 16;
constexpr int kRadixTile = kRadixThreads * kRadixItems;  // 4096 keys per block (R-sized passes)
// The P-sized depth sort is latency-bound: with 8 keys per thread it has twice the blocks
// (98 at 200k Gaussians instead of 49 on 148 SMs): lego 60.8 -> 48.8 us, fern 69.8 -> 51.3
// (4 keys per thread: 60.1 / 75.4 us — the per-block prefix walk over all blocks takes over).
// The R-sized tile sort keeps 16 (half the histogram table; dtu 0.667 vs 0.726 ms with 8).
constexpr int kDepthItems = 8;
constexpr int kDepthTile = kRadixThreads * kDepthItems;  // 2048
constexpr int kRadixBins = 256;

static int radix_blocks(int n) { return (n + kRadixTile - 1) / kRadixTile; }

// Loads of data another block wrote earlier IN THE SAME (cooperative) kernel must not
// use the non-coherent path; everything else may.
template <bool kCoop, typename T>
__device__ __forceinline__ T ld_u32(const T* p) {
    return kCoop ? __ldcg(p) : __ldg(p);
}

// KeyT is uint32_t for the depth sort and uint16_t for the tile sort (tile ids < 65536:
// 6 instead of 8 bytes per instance and pass).
template <typename KeyT, int kItems = kRadixItems>
struct RadixSmemT {
    uint32_t warp_cnt[kRadixThreads / 32][kRadixBins];  // 8 KB
    uint32_t global_base[kRadixBins];
    uint32_t block_start[kRadixBins];
    uint32_t scan[32];
    KeyT keys[kRadixThreads * kItems];      // 16 or 8 KB at 16 items
    uint32_t vals[kRadixThreads * kItems];  // 16 KB
};
typedef RadixSmemT<uint32_t> RadixSmem;

// hist[d * nb + tile] = number of keys of `tile` with digit d.  s_cnt: 256 words.
template <bool kCoop, typename KeyT, int kItems = kRadixItems>
__device__ __forceinline__ void radix_hist_tile(uint32_t* s_cnt, int tile, const KeyT* __restrict__ keys, int n,
                                                int shift, uint32_t* __restrict__ hist, int nb) {
    constexpr int kTileN = kRadixThreads * kItems;
    const int base = tile * kTileN;
    if (!kCoop && sizeof(KeyT) == 2 && kItems == 16 && base + kTileN <= n) {
        // full tile of 16-bit keys: 16 consecutive keys per thread as two 16-byte loads (a
        // histogram does not care which thread sees which key)
        const uint4* p = reinterpret_cast<const uint4*>(keys + base) + 2 * threadIdx.x;
        const uint4 v0 = __ldg(p), v1 = __ldg(p + 1);
        s_cnt[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int i = 0; i < 8; i++) {
            atomicAdd(&s_cnt[((w[i] & 0xffffu) >> shift) & 0xffu], 1u);
            atomicAdd(&s_cnt[((w[i] >> 16) >> shift) & 0xffu], 1u);
        }
        __syncthreads();
        hist[(size_t)threadIdx.x * nb + tile] = s_cnt[threadIdx.x];
        __syncthreads();
        return;
    }
    // all 16 loads in flight before the first shared atomic (one memory round trip)
    uint32_t k[kItems];
#pragma unroll
    for (int i = 0; i < kItems; i++) {
        const int j = base + i * kRadixThreads + threadIdx.x;
        k[i] = j < n ? (uint32_t)ld_u32<kCoop>(keys + j) : 0u;
    }
    s_cnt[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kItems; i++) {
        const int j = base + i * kRadixThreads + threadIdx.x;
        if (j < n) atomicAdd(&s_cnt[(k[i] >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nb + tile] = s_cnt[threadIdx.x];
    __syncthreads();
}

template <typename KeyT>
__global__ void __launch_bounds__(kRadixThreads) radix_hist(const KeyT* __restrict__ keys, int n, int shift,
                                                           uint32_t* __restrict__ hist, int nb) {
    __shared__ uint32_t s_cnt[kRadixBins];
    radix_hist_tile<false, KeyT>(s_cnt, blockIdx.x, keys, n, shift, hist, nb);
}

// One block per digit: exclusive scan of the digit's row over blocks, row total out.
// Blocked arrangement: each thread owns a contiguous slice of the row, so a row of any
// length costs ONE block-wide scan (the first version looped a block scan per 256
// entries: 36 us for 10k tiles).
__global__ void __launch_bounds__(256) radix_rowscan(uint32_t* __restrict__ hist, int nb,
                                                    uint32_t* __restrict__ digit_totals) {
    // Rows are walked in chunks of 2048 entries: coalesced loads into shared memory (one pad
    // word per 32 so that a thread's 8 consecutive entries are conflict-free), an 8-entry
    // serial scan per thread, one block scan of the 256 partial sums, coalesced stores.  (A
    // thread-owns-a-contiguous-segment version read the row with a stride of nb/256 words:
    // 30 us per pass at 10k radix blocks.)
    constexpr int kPer = 8, kChunk = 256 * kPer;
    __shared__ uint32_t s_tile[kChunk + kChunk / 32];
    __shared__ uint32_t s_warp[32];
    uint32_t* row = hist + (size_t)blockIdx.x * nb;
    uint32_t carry = 0;
    for (int base = 0; base < nb; base += kChunk) {
#pragma unroll
        for (int i = 0; i < kPer; i++) {
            const int l = i * 256 + (int)threadIdx.x, j = base + l;
            s_tile[l + (l >> 5)] = j < nb ? row[j] : 0u;
        }
        __syncthreads();
        uint32_t v[kPer], acc = 0;
#pragma unroll
        for (int i = 0; i < kPer; i++) {
            const int l = (int)threadIdx.x * kPer + i;
            v[i] = s_tile[l + (l >> 5)];
            acc += v[i];
        }
        uint32_t total;
        const uint32_t inc = block_inclusive_scan(acc, s_warp, total);
        uint32_t run = carry + inc - acc;
#pragma unroll
        for (int i = 0; i < kPer; i++) {
            const int l = (int)threadIdx.x * kPer + i;
            s_tile[l + (l >> 5)] = run;
            run += v[i];
        }
        carry += total;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kPer; i++) {
            const int l = i * 256 + (int)threadIdx.x, j = base + l;
            if (j < nb) row[j] = s_tile[l + (l >> 5)];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) digit_totals[blockIdx.x] = carry;
}

// Lanes of the warp holding the same 8-bit digit.  Eight ballots, constant time;
// __match_any_sync iterates over the distinct values and is ~3x slower on high-entropy
// digits (measured: 43 us vs 24 us per pass over 3.1M keys).
__device__ __forceinline__ unsigned match_digit8(uint32_t d) {
    unsigned peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const unsigned vote = __ballot_sync(0xffffffffu, (d >> b) & 1u);
        peers &= ((d >> b) & 1u) ? vote : ~vote;
    }
    return peers;
}

// Stable rank-and-scatter of one 4096-key tile.  kIota: values are the key indices.
// kPreloaded: the caller already put the digit totals in sm.block_start[] and the
// same-digit-earlier-tiles prefix in sm.global_base[] (cooperative in-block-prefix path).
template <bool kIota, bool kCoop, bool kPreloaded, typename KeyT, int kItems = kRadixItems>
__device__ __forceinline__ void radix_scatter_tile(RadixSmemT<KeyT, kItems>& sm, int tile, const KeyT* __restrict__ keys_in,
                                                   const uint32_t* __restrict__ vals_in,
                                                   KeyT* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                   const uint32_t* __restrict__ hist_scanned,
                                                   const uint32_t* __restrict__ digit_totals, int n, int shift,
                                                   int nb) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kTileN = kRadixThreads * kItems;
    const int base = tile * kTileN;
    const unsigned lt_mask = (1u << lane) - 1u;

    // issue every global load of the tile first (one memory round trip), then rank
    uint32_t k[kItems], v[kItems];
    uint32_t rank2[kItems / 2];  // two 16-bit ranks per register (rank < 4096)
#pragma unroll
    for (int r = 0; r < kItems; r++) {
        const int j = base + warp * (32 * kItems) + r * 32 + lane;  // warp-striped: order = (warp, r, lane)
        const bool valid = j < n;
        k[r] = valid ? (uint32_t)ld_u32<kCoop>(keys_in + j) : 0xffffffffu;
        v[r] = valid ? (kIota ? (uint32_t)j : ld_u32<kCoop>(vals_in + j)) : 0u;
    }
#pragma unroll
    for (int w = 0; w < kRadixThreads / 32; w++) sm.warp_cnt[w][tid] = 0;
    {
        // global base of digit `tid` = (keys with a smaller digit) + (same digit, earlier blocks)
        const uint32_t tot = kPreloaded ? sm.block_start[tid] : ld_u32<kCoop>(digit_totals + tid);
        const uint32_t before = kPreloaded ? sm.global_base[tid] : ld_u32<kCoop>(hist_scanned + (size_t)tid * nb + tile);
        uint32_t all;
        const uint32_t inc = block_inclusive_scan(tot, sm.scan, all);
        sm.global_base[tid] = inc - tot + before;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kItems; r++) {
        const int j = base + warp * (32 * kItems) + r * 32 + lane;
        // out-of-range slots sit at the very end of the tile and carry digit 255, so they
        // rank after every real key and are simply not written back.
        const uint32_t d = (j < n) ? ((k[r] >> shift) & 0xffu) : 0xffu;
        const unsigned peers = match_digit8(d);
        const uint32_t cnt = sm.warp_cnt[warp][d];
        __syncwarp();
        const uint32_t rk = cnt + __popc(peers & lt_mask);
        if (r & 1) rank2[r >> 1] |= rk << 16; else rank2[r >> 1] = rk;
        if ((peers & lt_mask) == 0) sm.warp_cnt[warp][d] = cnt + __popc(peers);  // lowest peer updates
        __syncwarp();
    }
    __syncthreads();
    {
        // digit `tid`: exclusive prefix over warps, block total, then prefix over digits
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kRadixThreads / 32; w++) {
            const uint32_t c = sm.warp_cnt[w][tid];
            sm.warp_cnt[w][tid] = run;
            run += c;
        }
        uint32_t all;
        const uint32_t inc = block_inclusive_scan(run, sm.scan, all);
        sm.block_start[tid] = inc - run;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kItems; r++) {
        const int j = base + warp * (32 * kItems) + r * 32 + lane;
        const uint32_t d = (j < n) ? ((k[r] >> shift) & 0xffu) : 0xffu;
        const uint32_t rk = (r & 1) ? (rank2[r >> 1] >> 16) : (rank2[r >> 1] & 0xffffu);
        const uint32_t pos = sm.block_start[d] + sm.warp_cnt[warp][d] + rk;
        sm.keys[pos] = (KeyT)k[r];
        sm.vals[pos] = v[r];
    }
    __syncthreads();
    const int nvalid = min(kTileN, n - base);
    for (int i = tid; i < nvalid; i += kRadixThreads) {
        const uint32_t key = sm.keys[i];
        const uint32_t d = (key >> shift) & 0xffu;
        const uint32_t g = sm.global_base[d] + ((uint32_t)i - sm.block_start[d]);
        keys_out[g] = (KeyT)key;
        vals_out[g] = sm.vals[i];
    }
    __syncthreads();
}

// 3 blocks/SM at 80 registers; 4 (64 registers) and 5 (48) spill and are slower (B200, dtu
// R = 42M: binning 0.667 / 0.681 / 0.727 ms).
template <bool kIota, typename KeyT>
__global__ void __launch_bounds__(kRadixThreads, 3) radix_scatter(const KeyT* __restrict__ keys_in,
                                                              const uint32_t* __restrict__ vals_in,
                                                              KeyT* __restrict__ keys_out,
                                                              uint32_t* __restrict__ vals_out,
                                                              const uint32_t* __restrict__ hist_scanned,
                                                              const uint32_t* __restrict__ digit_totals, int n,
                                                              int shift, int nb) {
    __shared__ RadixSmemT<KeyT> sm;
    radix_scatter_tile<kIota, false, false, KeyT>(sm, blockIdx.x, keys_in, vals_in, keys_out, vals_out, hist_scanned,
                                                  digit_totals, n, shift, nb);
}

// One stable 8-bit pass.  scratch: hist u32[256*nb] + digit_totals u32[256].
template <typename KeyT>
static void radix_pass(const KeyT* keys_in, const uint32_t* vals_in, KeyT* keys_out, uint32_t* vals_out,
                       uint32_t* hist, uint32_t* digit_totals, int n, int shift, bool iota, cudaStream_t stream) {
    const int nb = radix_blocks(n);
    radix_hist<KeyT><<<nb, kRadixThreads, 0, stream>>>(keys_in, n, shift, hist, nb);
    radix_rowscan<<<kRadixBins, 256, 0, stream>>>(hist, nb, digit_totals);
    if (iota)
        radix_scatter<true, KeyT><<<nb, kRadixThreads, 0, stream>>>(keys_in, nullptr, keys_out, vals_out, hist,
                                                                  digit_totals, n, shift, nb);
    else
        radix_scatter<false, KeyT><<<nb, kRadixThreads, 0, stream>>>(keys_in, vals_in, keys_out, vals_out, hist,
                                                                   digit_totals, n, shift, nb);
    count_launch(3);
}

static size_t radix_scratch_elems(int n) { return (size_t)kRadixBins * radix_blocks(n) + kRadixBins; }

// ------------------------------------------------------------------ generic (key, iota) sort
// Stable LSD sort of n 32-bit keys on bits [0, bits) carrying the original index: the
// Morton-order pass of the k-nearest-neighbour initialiser (knn.cu) reuses the rasterizer's
// radix passes.  scratch: keysA keysB idsA idsB (n words each) + histograms.
size_t sort_keys_iota_scratch_bytes(int n) {
    return 4 * align_up((size_t)n * 4, 256) + align_up(radix_scratch_elems(n) * 4, 256);
}

cudaError_t sort_keys_iota_u32(int n, const uint32_t* keys, int bits, char* scratch, const uint32_t** keys_sorted,
                               const uint32_t** ids_sorted, cudaStream_t stream) {
    char* q = scratch;
    uint32_t* kbuf[2]; uint32_t* ibuf[2];
    kbuf[0] = reinterpret_cast<uint32_t*>(q); q += align_up((size_t)n * 4, 256);
    kbuf[1] = reinterpret_cast<uint32_t*>(q); q += align_up((size_t)n * 4, 256);
    ibuf[0] = reinterpret_cast<uint32_t*>(q); q += align_up((size_t)n * 4, 256);
    ibuf[1] = reinterpret_cast<uint32_t*>(q); q += align_up((size_t)n * 4, 256);
    uint32_t* hist = reinterpret_cast<uint32_t*>(q);
    uint32_t* totals = hist + (size_t)kRadixBins * radix_blocks(n);
    const int passes = bits <= 0 ? 1 : (bits + 7) / 8;
    const uint32_t* kin = keys;
    const uint32_t* iin = nullptr;
    for (int p = 0; p < passes; p++) {
        radix_pass<uint32_t>(kin, iin, kbuf[p & 1], ibuf[p & 1], hist, totals, n, 8 * p, p == 0, stream);
        kin = kbuf[p & 1];
        iin = ibuf[p & 1];
    }
    *keys_sorted = kin;
    *ids_sorted = iin;
    return cudaGetLastError();
}

// ------------------------------------------------------------------ phase 1 (P-sized)
// The whole of phase 1 — four radix passes over float_bits(depth) and the gather-scan of
// tiles_touched in the resulting order — as ONE cooperative kernel: the P-sized passes
// are latency-bound (49 tiles at 200k Gaussians), so 15 separate launches cost more in
// launch/drain latency than in work.  Grid-wide barriers replace the launch boundaries;
// a cooperative launch either guarantees co-residency of all blocks or fails, it cannot
// hang.  Falls back to the multi-kernel path when cooperative launch is unavailable.
struct Phase1Args {
    int P;
    const uint32_t* dkeys;
    const uint32_t* tiles_touched;
    const uint32_t* key_bits;  // [0] OR, [1] AND of the visible depth keys
    uint32_t *keysB, *keysC, *valsA, *valsB, *hist, *totals, *sums;
    uint32_t *sorted_ids, *sorted_offsets;
    int need_offsets;
};

// Up to this many radix tiles each block derives its own offsets straight from the
// tile-major histogram (one coalesced column walk) instead of a separate row-scan phase
// and its grid barrier.
constexpr int kInBlockPrefixMaxTiles = 256;

__global__ void __launch_bounds__(kRadixThreads, 2) depth_sort_coop(Phase1Args a) {
    __shared__ RadixSmemT<uint32_t, kDepthItems> sm;
    cg::grid_group grid = cg::this_grid();
    const int n = a.P;
    const int nb = (n + kDepthTile - 1) / kDepthTile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Bytes in which every visible key agrees need no pass.  (Invisible Gaussians carry
    // key 0 and may end up anywhere: they emit no instances.)
    const uint32_t vary = __ldcg(a.key_bits) ^ __ldcg(a.key_bits + 1);
    int num_passes = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) num_passes += ((vary >> (8 * k)) & 0xffu) ? 1 : 0;
    if (__ldcg(a.key_bits + 1) == 0xffffffffu && __ldcg(a.key_bits) == 0u) num_passes = 0;  // nothing visible
    const bool in_block_prefix = nb <= kInBlockPrefixMaxTiles;
    int ip = 0;
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        const int shift = 8 * k;
        if (num_passes == 0 || ((vary >> shift) & 0xffu) == 0) continue;  // grid-uniform
        const bool first = ip == 0, last = ip == num_passes - 1;
        // ping-pong: keys dkeys->B->C->B->C, values iota->A->B->A->..., last pass -> sorted_ids
        const uint32_t* kin = first ? a.dkeys : ((ip & 1) ? a.keysB : a.keysC);
        uint32_t* kout = (ip & 1) ? a.keysC : a.keysB;
        const uint32_t* vin = (ip & 1) ? a.valsA : a.valsB;
        uint32_t* vout = last ? a.sorted_ids : ((ip & 1) ? a.valsB : a.valsA);
        if (in_block_prefix) {
            // tile-major histogram: hist[tile * 256 + digit]
            for (int t = blockIdx.x; t < nb; t += gridDim.x) {
                const int base = t * kDepthTile;
                uint32_t kk[kDepthItems];
#pragma unroll
                for (int i = 0; i < kDepthItems; i++) {
                    const int j = base + i * kRadixThreads + tid;
                    kk[i] = j < n ? __ldcg(kin + j) : 0u;
                }
                sm.global_base[tid] = 0;
                __syncthreads();
#pragma unroll
                for (int i = 0; i < kDepthItems; i++) {
                    const int j = base + i * kRadixThreads + tid;
                    if (j < n) atomicAdd(&sm.global_base[(kk[i] >> shift) & 0xffu], 1u);
                }
                __syncthreads();
                a.hist[(size_t)t * kRadixBins + tid] = sm.global_base[tid];
                __syncthreads();
            }
            grid.sync();
            for (int t = blockIdx.x; t < nb; t += gridDim.x) {
                // digit `tid`: total over all tiles and prefix over the tiles before t
                uint32_t total = 0, before = 0;
#pragma unroll 16   // independent L2 loads: 16 in flight per thread instead of the default 4 (the walk is one latency chain)
                for (int b = 0; b < nb; b++) {
                    const uint32_t v = __ldcg(a.hist + (size_t)b * kRadixBins + tid);
                    total += v;
                    before += b < t ? v : 0u;
                }
                // hand the per-tile scatter its two inputs through shared memory
                sm.block_start[tid] = total;     // digit totals
                sm.global_base[tid] = before;    // same digit, earlier tiles
                __syncthreads();
                if (first)
                    radix_scatter_tile<true, true, true, uint32_t, kDepthItems>(sm, t, kin, nullptr, kout, vout, nullptr, nullptr, n, shift, nb);
                else
                    radix_scatter_tile<false, true, true, uint32_t, kDepthItems>(sm, t, kin, vin, kout, vout, nullptr, nullptr, n, shift, nb);
            }
            grid.sync();
        } else {
            for (int t = blockIdx.x; t < nb; t += gridDim.x)
                radix_hist_tile<true, uint32_t, kDepthItems>(sm.global_base, t, kin, n, shift, a.hist, nb);
            grid.sync();
            // row scan: one warp per digit row
            for (int row = blockIdx.x * (kRadixThreads / 32) + warp; row < kRadixBins;
                 row += gridDim.x * (kRadixThreads / 32)) {
                uint32_t* r = a.hist + (size_t)row * nb;
                uint32_t carry = 0;
                for (int base = 0; base < nb; base += 32) {
                    const int j = base + lane;
                    const uint32_t v = j < nb ? __ldcg(r + j) : 0u;
                    const uint32_t inc = warp_inclusive_scan(v, lane);
                    if (j < nb) r[j] = carry + inc - v;
                    carry += __shfl_sync(0xffffffffu, inc, 31);
                }
                if (lane == 0) a.totals[row] = carry;
            }
            grid.sync();
            for (int t = blockIdx.x; t < nb; t += gridDim.x) {
                if (first)
                    radix_scatter_tile<true, true, false, uint32_t, kDepthItems>(sm, t, kin, nullptr, kout, vout, a.hist, a.totals, n, shift, nb);
                else
                    radix_scatter_tile<false, true, false, uint32_t, kDepthItems>(sm, t, kin, vin, kout, vout, a.hist, a.totals, n, shift, nb);
            }
            grid.sync();
        }
        ip++;
    }
    if (num_passes == 0) {  // all visible keys identical (or nothing visible): identity order
        for (int j = blockIdx.x * kRadixThreads + tid; j < n; j += gridDim.x * kRadixThreads) a.sorted_ids[j] = (uint32_t)j;
        grid.sync();
    }
    if (!a.need_offsets) return;  // grid-uniform: the direct tile binning derives its own offsets
    // exclusive scan of tiles_touched in depth order -> emission offsets
    const int nt = (n + kScanTile - 1) / kScanTile;
    for (int t = blockIdx.x; t < nt; t += gridDim.x) {
        const int base = t * kScanTile;
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < kScanItems; i++) {
            const int j = base + i * kScanThreads + threadIdx.x;
            if (j < n) acc += __ldg(a.tiles_touched + __ldcg(a.sorted_ids + j));
        }
        uint32_t total;
        block_inclusive_scan(acc, sm.scan, total);
        if (threadIdx.x == 0) a.sums[t] = total;
    }
    grid.sync();
    if (blockIdx.x == 0) {
        uint32_t carry = 0;
        for (int base = 0; base < nt; base += kRadixThreads) {
            const int j = base + threadIdx.x;
            const uint32_t v = j < nt ? __ldcg(a.sums + j) : 0u;
            uint32_t total;
            const uint32_t inc = block_inclusive_scan(v, sm.scan, total);
            if (j < nt) a.sums[j] = carry + inc - v;
            carry += total;
        }
    }
    grid.sync();
    for (int t = blockIdx.x; t < nt; t += gridDim.x) {
        const int base = t * kScanTile + threadIdx.x * kScanItems;  // blocked arrangement
        uint32_t v[kScanItems];
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < kScanItems; i++) {
            v[i] = (base + i < n) ? __ldg(a.tiles_touched + __ldcg(a.sorted_ids + base + i)) : 0u;
            acc += v[i];
        }
        uint32_t total;
        const uint32_t inc = block_inclusive_scan(acc, sm.scan, total);
        uint32_t run = __ldcg(a.sums + t) + inc - acc;
#pragma unroll
        for (int i = 0; i < kScanItems; i++) {
            if (base + i < n) a.sorted_offsets[base + i] = run;
            run += v[i];
        }
    }
}

// scratch layout (u32 elements): keysA[P] keysB[P] valsA[P] valsB[P] hist[...] scan_sums[...]
static int depth_blocks(int n) { return (n + kDepthTile - 1) / kDepthTile; }
// histogram scratch of phase 1: the larger of the cooperative (2048-key blocks) and the
// multi-kernel fallback (4096-key blocks) layouts
static size_t depth_scratch_elems(int n) { return (size_t)kRadixBins * depth_blocks(n) + kRadixBins; }

size_t binning_phase1_scratch_bytes(int P) {
    const size_t p = (size_t)(P > 0 ? P : 0);
    return (align_up(p * 4, 256) * 4 + align_up(depth_scratch_elems(P) * 4, 256) +
            align_up((scan_tiles(P) + 1) * 4, 256));
}

// Largest cooperative grid for depth_sort_coop on the current device, 0 if unsupported.
static int coop_grid_limit() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev == cached_dev) return cached;
    cached_dev = dev;
    cached = 0;
    const char* e = getenv("B3GS_PHASE1");
    if (e && !strcmp(e, "multi")) return 0;
    int coop = 0, sms = 0, per_sm = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (!coop || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, depth_sort_coop, kRadixThreads, 0) != cudaSuccess)
        return 0;
    cached = sms * per_sm;
    return cached;
}

cudaError_t run_binning_phase1(const BinningPhase1Args& a, cudaStream_t stream) {
    const size_t p = (size_t)a.P;
    char* q = a.scratch;
    uint32_t* keysB = reinterpret_cast<uint32_t*>(q); q += align_up(p * 4, 256);
    uint32_t* keysC = reinterpret_cast<uint32_t*>(q); q += align_up(p * 4, 256);
    uint32_t* valsA = reinterpret_cast<uint32_t*>(q); q += align_up(p * 4, 256);
    uint32_t* valsB = reinterpret_cast<uint32_t*>(q); q += align_up(p * 4, 256);
    uint32_t* hist = reinterpret_cast<uint32_t*>(q);  q += align_up(depth_scratch_elems(a.P) * 4, 256);
    uint32_t* sums = reinterpret_cast<uint32_t*>(q);
    uint32_t* totals = hist + (size_t)kRadixBins * radix_blocks(a.P);   // multi-kernel fallback layout
    const uint32_t* dkeys = reinterpret_cast<const uint32_t*>(a.depths);
    const int limit = coop_grid_limit();
    if (limit > 0) {
        Phase1Args k;
        k.P = a.P; k.dkeys = dkeys; k.tiles_touched = a.tiles_touched; k.key_bits = a.key_bits;
        k.keysB = keysB; k.keysC = keysC; k.valsA = valsA; k.valsB = valsB;
        k.hist = hist; k.totals = hist + (size_t)kRadixBins * depth_blocks(a.P); k.sums = sums;
        k.sorted_ids = a.sorted_ids; k.sorted_offsets = a.sorted_offsets;
        k.need_offsets = a.need_offsets;
        // at least 32 blocks so that the 256 digit rows find 256 warps
        int grid = depth_blocks(a.P) < 32 ? 32 : depth_blocks(a.P);
        if (grid > limit) grid = limit;
        void* args[] = {&k};
        cudaError_t e = cudaLaunchCooperativeKernel((const void*)depth_sort_coop, dim3(grid), dim3(kRadixThreads), args,
                                                    0, stream);
        if (e == cudaSuccess) {
            count_launch();
            return cudaSuccess;
        }
        (void)cudaGetLastError();  // clear and fall through to the multi-kernel path
    }
    // 4 stable passes over float_bits(depth); depth > 0.2 for every visible Gaussian so
    // unsigned order == float order, and the reference sorts the raw bits anyway.
    radix_pass<uint32_t>(dkeys, nullptr, keysB, valsA, hist, totals, a.P, 0, true, stream);
    radix_pass<uint32_t>(keysB, valsA, keysC, valsB, hist, totals, a.P, 8, false, stream);
    radix_pass<uint32_t>(keysC, valsB, keysB, valsA, hist, totals, a.P, 16, false, stream);
    radix_pass<uint32_t>(keysB, valsA, keysC, a.sorted_ids, hist, totals, a.P, 24, false, stream);
    // emission offsets in depth order
    if (a.need_offsets) exclusive_scan_gather(a.tiles_touched, a.sorted_ids, a.sorted_offsets, sums, nullptr, a.P, stream);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ phase 2 (R-sized)
// One warp per 32 depth-ordered Gaussians.  The warp's instances form one contiguous run
// of the output (their offsets are an exclusive scan), so lane L writes instances
// t0+L of that run, 32 at a time: every lane is busy whatever the rectangle sizes are and
// the stores are fully coalesced.  Gaussians with tiles are first compacted to the low
// lanes; the owner of each instance in a 32-wide window then costs one REDUX.OR (a bit
// per Gaussian starting inside the window) and one POPC per lane.
template <typename KeyT>
__global__ void __launch_bounds__(256) emit_instances(int P, const uint32_t* __restrict__ sorted_ids,
                                                     const uint32_t* __restrict__ sorted_offsets,
                                                     const float4* __restrict__ records, const int* __restrict__ radii,
                                                     KeyT* __restrict__ tile_keys, uint32_t* __restrict__ ids,
                                                     int grid_x, int grid_y) {
    const unsigned full = 0xffffffffu;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t g = 0, off = 0;
    int x0 = 0, y0 = 0, w = 1, cnt = 0;
    if (i < P) {
        g = sorted_ids[i];
        off = sorted_offsets[i];
        const int r = radii[g];
        if (r > 0) {
            const float4 a = records[(size_t)g * B3_REC_VEC4];
            int x1, y1;
            tile_rect(a.x, a.y, (float)r, grid_x, grid_y, x0, y0, x1, y1);
            w = x1 - x0;
            cnt = w * (y1 - y0);
        }
    }
    const unsigned nz = __ballot_sync(full, cnt > 0);
    if (nz == 0) return;
    const int nnz = __popc(nz);
    // compact the Gaussians that own tiles to lanes 0..nnz-1 (offsets strictly increase there)
    const uint32_t warp_begin = __shfl_sync(full, off, __ffs(nz) - 1);
    const uint32_t last = 31 - __clz(nz);
    const uint32_t warp_total = __shfl_sync(full, off + (uint32_t)cnt, last) - warp_begin;
    const unsigned src = __fns(nz, 0, lane + 1);  // lane of the (lane+1)-th set bit, or ~0
    const int s = src < 32 ? (int)src : 0;
    uint32_t crel = __shfl_sync(full, off, s) - warp_begin;
    const uint32_t cxy = __shfl_sync(full, (uint32_t)x0 | ((uint32_t)y0 << 16), s);
    const uint32_t cw = __shfl_sync(full, (uint32_t)w, s);
    const uint32_t cg = __shfl_sync(full, g, s);
    if (lane >= nnz) crel = 0xffffffffu;  // never starts inside a window
    uint32_t below = 0;  // compacted Gaussians starting before the window
    const unsigned le_mask = 0xffffffffu >> (31 - lane);
    for (uint32_t t0 = 0; t0 < warp_total; t0 += 32) {
        const uint32_t d = crel - t0;  // wraps for starts before the window
        const unsigned starts = __reduce_or_sync(full, d < 32u ? (1u << d) : 0u);
        const int owner = (int)below + __popc(starts & le_mask) - 1;
        below += __popc(starts);
        const uint32_t t = t0 + lane;
        const uint32_t orel = __shfl_sync(full, crel, owner & 31);
        const uint32_t oxy = __shfl_sync(full, cxy, owner & 31);
        const uint32_t ow = __shfl_sync(full, cw, owner & 31);
        const uint32_t og = __shfl_sync(full, cg, owner & 31);
        if (t < warp_total) {
            const uint32_t local = t - orel;
            // local / ow for local < 2^22: float quotient, corrected by one step either way
            int ty = (int)(((float)local + 0.5f) * __frcp_rn((float)ow));
            int tx = (int)local - ty * (int)ow;
            if (tx < 0) { ty--; tx += (int)ow; }
            if (tx >= (int)ow) { ty++; tx -= (int)ow; }
            tile_keys[warp_begin + t] = (KeyT)(((oxy >> 16) + ty) * grid_x + (oxy & 0xffffu) + tx);
            ids[warp_begin + t] = og;
        }
    }
}

template <typename KeyT>
__global__ void __launch_bounds__(256) tile_ranges_k(int R, const KeyT* __restrict__ tile_keys,
                                                    uint2* __restrict__ ranges) {
    // 8 consecutive keys per thread plus the predecessor of the first
    constexpr int kPer = 8;
    const int base = (blockIdx.x * blockDim.x + threadIdx.x) * kPer;
    if (base >= R) return;
    uint32_t k[kPer];
    if (base + kPer <= R) {  // 8 keys = one (16-bit) or two (32-bit) 16-byte loads; base is a multiple of 8
        if (sizeof(KeyT) == 2) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(tile_keys + base));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; i++) { k[2 * i] = w[i] & 0xffffu; k[2 * i + 1] = w[i] >> 16; }
        } else {
            const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(tile_keys + base));
            const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(tile_keys + base) + 1);
            k[0] = v0.x; k[1] = v0.y; k[2] = v0.z; k[3] = v0.w; k[4] = v1.x; k[5] = v1.y; k[6] = v1.z; k[7] = v1.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < kPer; i++) k[i] = base + i < R ? (uint32_t)__ldg(tile_keys + base + i) : 0u;
    }
    uint32_t prev = base > 0 ? (uint32_t)__ldg(tile_keys + base - 1) : 0u;
#pragma unroll
    for (int i = 0; i < kPer; i++) {
        const int idx = base + i;
        if (idx >= R) break;
        const uint32_t cur = k[i];
        if (idx == 0) {
            ranges[cur].x = 0;
        } else if (cur != prev) {
            ranges[prev].y = idx;
            ranges[cur].x = idx;
        }
        if (idx == R - 1) ranges[cur].y = R;
        prev = cur;
    }
}

static int tile_passes(int T) {
    int bits = 0;
    while ((1ll << bits) < (long long)T) bits++;
    return bits <= 8 ? 1 : (bits <= 16 ? 2 : (bits <= 24 ? 3 : 4));
}


// ------------------------------------------------------------------ phase 2: direct tile binning
// (large tile grids only, see binning_uses_tile_bins)
// The instance stream in depth order is never materialised.  The P depth-ordered Gaussians
// are cut into nb batches of B; a BLOCK of 8 warps owns a batch, warp w the w-th eighth of it
// (still in depth order).  Final position of instance (g, tile) =
//   tile_start[tile] + #instances of `tile` in earlier batches          (table[b][tile], scanned)
//                    + #instances of `tile` in earlier warps of the block (wcount[b][warp][tile], u8)
//                    + rank among the warp's own Gaussians,
// and the last term is what the warp's private per-tile counter holds when the warp reaches
// g, because a warp walks its Gaussians in depth order and the tiles of one Gaussian are
// distinct (lanes = tiles of the current rectangle: no two lanes touch the same counter).
//   count:   per-warp counters -> wcount (one byte each), summed over the block -> table[b][.]
//   scan:    table[b][t] <- tile_start[t] + sum_{b' < b} table[b'][t]   (three small kernels)
//   scatter: prefix of wcount over the 8 warps on top of table[b][.], then every instance
//            does  pos = counter[warp][tile]++ ; point_list[pos] = id.
// The image is cut into BANDS of whole tile rows (<= kBandTilesMax tiles) and a block makes
// one pass over its batch per band with only that band's counters in shared memory.  Why,
// all measured on B200 at 1M Gaussians / 1600x1200 / 42M instances (profiles/README.md r02c-d):
//  (i)  one warp walking 512 Gaussians against all T counters is a 28 000-instruction dependent
//       chain at 7 warps per SM — 230 us for the count alone; eight short chains per block
//       and several blocks per SM hide most of that latency (156 us, 103 us in the final form);
//  (ii) written in depth order, every 32-byte sector of the output stays half-written for the
//       lifetime of a batch, the open sectors of ~1000 concurrent batches (85 MB) do not fit
//       the L2 and each is evicted and re-fetched (1.8 GB of DRAM traffic for a 169 MB list,
//       1.05 ms).  Band by band, all blocks write the same narrow region at about the same
//       time and most sectors complete while resident (0.38 GB, 0.46 ms);
//  (iii) an instance-parallel variant (lane = Gaussian, per-tile chunk masks for the rank)
//       loses to lane divergence over the rectangle sizes: 1.37 ms.
// With direct 4-byte stores from the walk (~110 G stores/s) depth sort + binning took 0.677 ms
// against 0.777 ms for emit + two radix passes at 42 instances per Gaussian (1M / 1600x1200),
// 0.126 against 0.141 ms at 16 (200k / 800x800); staging the band in shared memory and copying it
// out coalesced (tile_bins_kernel below) brought the 1M case to 0.44 ms.  Bit-identical output
// either way: a stable distribution by tile of a (depth, id)-ordered sequence.
constexpr int kBinChunk = 32;            // batches per scan chunk
constexpr int kBinWarps = 8;             // warps per block = sub-batches per batch
constexpr int kBandTilesMax = 2040;      // widest tile row / largest band the direct path takes (a width is packed in 11 bits)
constexpr int kBandTilesDefault = 1152;  // counters of one band: 8 x 2.25 KB of shared memory
constexpr int kBinPer = 8;               // scatter: tiles per thread of a band's table rows kept in registers (kBandTilesMax / 256)
constexpr int kBandRowsMax = 200;        // rows of a band (a clipped height is packed in 8 bits, 255 = lane idle)
constexpr int kMaxBinBatch = 2040;       // Gaussians per batch: <= 255 per warp (one-byte per-warp counts)
constexpr int kMinBinTiles = 0;          // smaller grids would go through emit + radix (none: the direct path wins or ties at all BASELINE sizes)

struct TileBinArgs {
    int P, B, nb, T, T_pad, grid_x, grid_y, band_rows;
    int band_max;                 // counters per warp (>= tiles of a band)
    int stage_cap;                // scatter: slots of the block-local staging buffer
    uint32_t cap;                 // point_list capacity (instances)
    const uint32_t* sorted_ids;
    const uint2* rects;
    uint32_t* table;              // [nb][T_pad]
    uint2* wcount;                // [nb][T_pad]: the kBinWarps (8) per-warp counts of a tile, one byte each
    uint32_t* point_list;
};

// Batch size: 512 / 768 Gaussians per block (64 / 96 per warp) give 400-1300 blocks at the
// BASELINE sizes; doubled while the table would outweigh the instance list itself (many tiles, few
// instances).  B3GS_BIN_BATCH overrides (tuning).
static int bin_batch_size(int P, int R, int T_pad) {
    static const int forced = [] { const char* e = getenv("B3GS_BIN_BATCH"); return e ? atoi(e) : 0; }();
    if (forced >= 256 && forced <= kMaxBinBatch && forced % 8 == 0) return forced;
    // measured (profiles/README.md r02aa/r02ac): 1M Gaussians 0.354 (768) vs 0.374 (1024) vs 0.41 (640) ms;
    // 200k 0.073 (384: 521 blocks) vs 0.080 (512: 391 blocks, less than one wave of 3 x 148) vs 0.103 (768) ms;
    // 300k 0.089 (512) vs 0.106 (384) ms  ->  the largest size that still fills one wave of the scatter
    int B = 256;
    for (int cand : {768, 512, 384}) {
        if (cand == 768 && P < (1 << 19)) continue;
        if ((P + cand - 1) / cand >= 444) { B = cand; break; }
    }
    const size_t budget = (size_t)(R > 0 ? R : 0) * 8 + ((size_t)32 << 20);
    while (B < 1024 && (size_t)((P + B - 1) / B) * T_pad * 12 > budget) B *= 2;
    return B;
}
static int bin_band_max() {   // B3GS_BIN_BAND: tiles per band (tuning)
    static const int v = [] { const char* e = getenv("B3GS_BIN_BAND"); const int x = e ? atoi(e) : 0;
                              return x >= 64 && x <= kBandTilesMax ? (x + 1) & ~1 : kBandTilesDefault; }();
    return v;
}
static int bin_band_rows(int grid_x, int grid_y) {
    int r = bin_band_max() / grid_x;
    if (r > kBandRowsMax) r = kBandRowsMax;
    if (r > grid_y) r = grid_y;
    return r < 1 ? 1 : r;
}
// shared memory of one block: per-warp band counters (16-bit), the batch's rectangle staging
// (x0|y0<<16, w|h<<16, id), and — scatter — the band's output staging
static size_t bin_smem_bytes(int B, int band_max, bool scatter, int stage_cap) {   // stage_cap = 0: unstaged
    size_t n = (size_t)kBinWarps * band_max * 2 + (size_t)B * 12 + 32 * 4;
    if (scatter) n += (size_t)band_max * 4 + (size_t)(stage_cap ? stage_cap : band_max) * 4;
    return n;
}
// Staging slots of the scatter (one word each: batch-local Gaussian index << 16 | tile of the band):
// whatever three blocks per SM leave (228 KB per SM, 1 KB reserved per block) — at dtu 1.5x the
// mean of a (batch, band); instances beyond it are stored directly.  B3GS_BIN_STAGE overrides.
static int bin_stage_cap(int B, int band_max) {
    static const int forced = [] { const char* e = getenv("B3GS_BIN_STAGE"); const int x = e ? atoi(e) : 0;
                                   return x >= 256 && x <= 32768 ? x : 0; }();
    const long fixed = (long)bin_smem_bytes(B, band_max, true, 0) - (long)band_max * 4;
    long v = forced ? forced : ((228 * 1024) / 3 - 1024 - fixed) / 4;
    if (v > 32768) v = 32768;
    return v > band_max ? (int)v : band_max;   // the wide path keeps band_max 32-bit positions there
}

// count (kScatter = false): per-warp instance counts per tile of each band -> wcount (bytes),
// block totals -> table.
// scatter: the walk assigns every instance its slot in a BLOCK-LOCAL tile-major buffer of the band
// (local_off[tile] + instances of earlier warps + rank in the warp: 16-bit counters) and stages one
// word there (index of the Gaussian in the batch << 16 | tile of the band); the block then copies
// the buffer out: consecutive threads write consecutive slots of a tile's cell, so a warp store
// covers ~6 cells instead of 32 scattered words — the scattered 4-byte global store was what bound
// the first version (profiles/README.md r02n).  The band's rows of the table and of the per-warp
// counts are loaded into registers one band ahead (ncu r02ab: 12 % of the warp samples waited on them).
// kStaged = false: the walk stores straight to global memory (s_base[tile] + slot) — cheaper when a
// (batch, tile) cell holds only ~2 instances and there is nothing to coalesce (fern: 0.093 vs 0.116 ms).
template <bool kScatter, bool kStaged>
__global__ void __launch_bounds__(32 * kBinWarps) tile_bins_kernel(TileBinArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, b = blockIdx.x;
    const int band_max = a.band_max;
    const int kCap = !kScatter ? 0 : (kStaged ? a.stage_cap : band_max);   // unstaged: room for the wide path only
    uint32_t* s_oid = smem;                                  // scatter: [kCap] staged (index in the batch << 16 | tile); first, so
                                                             // that the walk's store address is slot * 4 + a constant
    uint32_t* s_base = s_oid + kCap;                         // scatter: [band_max] global position - local slot
    uint16_t* s_cnt = reinterpret_cast<uint16_t*>(s_base + (kScatter ? band_max : 0));   // [kBinWarps][band_max], 16-byte aligned
    uint32_t* s_xy = reinterpret_cast<uint32_t*>(s_cnt + kBinWarps * band_max);          // B
    uint32_t* s_wh = s_xy + a.B;                             // B
    uint32_t* s_id = s_wh + a.B;                             // B
    uint32_t* s_scan = s_id + a.B;                           // 32 words for the block scan
    uint32_t* row = a.table + (size_t)b * a.T_pad;
    uint2* wrow = a.wcount + (size_t)b * a.T_pad;
    static_assert(kBinWarps == 8, "the per-warp counts of a tile are packed into one 8-byte word");
    uint16_t* cnt = s_cnt + warp * band_max;                 // this warp's counters

    // scatter: this thread's tiles of the band's table / per-warp-count rows, loaded one band ahead
    uint2 c8[kBinPer];
    uint32_t rw[kBinPer];
    auto fetch = [&](int by0) {
        const int by1 = min(a.grid_y, by0 + a.band_rows);
        const int start = by0 * a.grid_x, tiles = (by1 - by0) * a.grid_x;
        const int per = (tiles + 32 * kBinWarps - 1) / (32 * kBinWarps);
        const int k0 = threadIdx.x * per, k1 = min(tiles, k0 + per);
#pragma unroll
        for (int j = 0; j < kBinPer; j++) {
            const bool on = k0 + j < k1;
            c8[j] = on ? __ldg(wrow + start + k0 + j) : make_uint2(0u, 0u);
            rw[j] = on ? __ldg(row + start + k0 + j) : 0u;
        }
    };
    if (kScatter) fetch(0);
    // stage the batch: rectangles (and ids) in depth order
    const int i0 = b * a.B, n = min(a.P, i0 + a.B) - i0;
    for (int k = threadIdx.x; k < n; k += 32 * kBinWarps) {
        const uint32_t g = __ldg(a.sorted_ids + i0 + k);
        const uint2 r = __ldg(a.rects + g);
        s_xy[k] = r.x; s_wh[k] = r.y;
        if (kScatter) s_id[k] = g;
    }
    // The walk's shared-window addresses and loop constants, pinned in registers (left to the compiler they
    // are re-derived from the constant bank / SR_CgaCtaId for every instance: ncu r02ac, 6 % of the samples).
    uint32_t cnt_sa = (uint32_t)__cvta_generic_to_shared(cnt), oid_sa = (uint32_t)__cvta_generic_to_shared(s_oid);
    int gx = a.grid_x;
    uint32_t stage_slots = (uint32_t)kCap;
    asm volatile("" : "+r"(gx), "+r"(stage_slots), "+r"(cnt_sa), "+r"(oid_sa));
    // one instance: rank = this warp's counter of the tile; scatter: stage it (or store it directly)
    auto touch = [&](uint32_t t, uint32_t packed, uint32_t bg) {
        const uint32_t ca = cnt_sa + 2u * t;
        unsigned short slot16;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(slot16) : "r"(ca) : "memory");
        const uint32_t slot = slot16;
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(ca), "h"((unsigned short)(slot + 1u)) : "memory");
        if (kScatter) {
            if (kStaged && slot < stage_slots) {
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(oid_sa + 4u * slot), "r"(packed | t) : "memory");
            } else {
                const uint32_t pos = s_base[t] + slot;
                if (pos < a.cap) a.point_list[pos] = kStaged ? s_id[packed >> 16] : bg;
            }
        }
    };
    const int sub = a.B / kBinWarps;                          // Gaussians per warp
    const int w0 = warp * sub, w1 = min(n, w0 + sub);

    for (int by0 = 0; by0 < a.grid_y; by0 += a.band_rows) {
        const int by1 = min(a.grid_y, by0 + a.band_rows);
        const int band_start = by0 * a.grid_x, band_tiles = (by1 - by0) * a.grid_x;
        uint32_t n_local = 0;
        if (kScatter) {
            // local slots: exclusive scan over the band's tiles of the block's counts (blocked
            // arrangement: a thread owns `per` consecutive tiles), then the warps' prefixes on top
            const int per = (band_tiles + 32 * kBinWarps - 1) / (32 * kBinWarps);
            const int k0 = threadIdx.x * per, k1 = min(band_tiles, k0 + per);
            uint32_t acc = 0;
#pragma unroll
            for (int j = 0; j < kBinPer; j++) {
                // sum of the eight bytes: pairwise adds inside the words (each byte <= 255; zero past k1)
                const uint32_t v = (c8[j].x & 0x00ff00ffu) + ((c8[j].x >> 8) & 0x00ff00ffu) + (c8[j].y & 0x00ff00ffu) + ((c8[j].y >> 8) & 0x00ff00ffu);
                acc += (v & 0xffffu) + (v >> 16);
            }
            const uint32_t inc = block_inclusive_scan(acc, s_scan, n_local);
            uint32_t run = inc - acc;
#pragma unroll
            for (int j = 0; j < kBinPer; j++) {
                const int k = k0 + j;
                if (k < k1) {
                    s_base[k] = rw[j] - run;
#pragma unroll
                    for (int w = 0; w < kBinWarps; w++) {
                        s_cnt[w * band_max + k] = (uint16_t)run;
                        run += ((w < 4 ? c8[j].x : c8[j].y) >> (8 * (w & 3))) & 0xffu;
                    }
                }
            }
            if (by1 < a.grid_y) fetch(by1);   // the next band's rows arrive while this band is walked
        } else {
            // zero the eight counter rows (contiguous, 16-byte aligned: B is a multiple of 8)
            uint4* z = reinterpret_cast<uint4*>(s_cnt);
            for (int k = threadIdx.x; k < band_max; k += 32 * kBinWarps) z[k] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();  // counters initialised (and, first band, the staging complete)

        // More instances in this (batch, band) than 16-bit slots can number (a few hundred Gaussians
        // that each cover most of the image): the warps take turns, in depth order, with 32-bit global
        // positions kept where the ids would have been staged.  Correct, not fast, and rare.
        const bool wide = kScatter && n_local > 65535u;      // block-uniform
        if (wide) {
            uint32_t* pos32 = s_oid;                         // kCap >= band_max words
            for (int k = threadIdx.x; k < band_tiles; k += 32 * kBinWarps) pos32[k] = __ldg(row + band_start + k);
            __syncthreads();
            for (int turn = 0; turn < kBinWarps; turn++) {
                if (warp == turn) {
                    for (int c = w0; c < w1; c++) {          // one Gaussian at a time, lanes over its tiles
                        const uint32_t xy = s_xy[c], w2 = s_wh[c];
                        const int y0 = (int)(xy >> 16), cy0 = max(y0, by0), cy1 = min(y0 + (int)(w2 >> 16), by1);
                        const int bw = (int)(w2 & 0xffffu), x0 = (int)(xy & 0xffffu);
                        if (w2 == 0u || cy1 <= cy0) continue;
                        const uint32_t gid = s_id[c];
                        for (int j = lane; j < bw * (cy1 - cy0); j += 32) {
                            const int yy = j / bw, xx = j - yy * bw;
                            const uint32_t t = (uint32_t)((cy0 - by0 + yy) * a.grid_x + x0 + xx);
                            const uint32_t pos = pos32[t];
                            pos32[t] = pos + 1u;
                            if (pos < a.cap) a.point_list[pos] = gid;
                        }
                        __syncwarp();
                    }
                }
                __syncthreads();
            }
            continue;                                        // next band
        }
        for (int c = w0; c < w1; c += 32) {
            // 32 Gaussians at once: clip to the band, precompute what the serial part needs
            uint32_t t0 = 0u, wh = 0u, g = 0u;
            if (c + lane < w1) {
                const uint32_t xy = s_xy[c + lane], w2 = s_wh[c + lane];
                const int y0 = (int)(xy >> 16), y1 = y0 + (int)(w2 >> 16);
                const int cy0 = max(y0, by0), cy1 = min(y1, by1);
                if (w2 != 0u && cy1 > cy0) {
                    t0 = (uint32_t)((cy0 - by0) * a.grid_x) + (xy & 0xffffu);
                    // width (11 bits) | clipped height << 11 (8 bits) | ceil-reciprocal of the width << 19:
                    // (L * inv) >> 10 == L / width exactly for L <= 32, width <= 32
                    const uint32_t bw = w2 & 0xffffu;
                    wh = bw | ((uint32_t)(cy1 - cy0) << 11) | (bw <= 32u ? (__float2uint_rz(__frcp_ru((float)bw) * 1024.f) + 1u) << 19 : 0u);   // 1024 / bw + 1 without the division routine
                    if (kScatter && !kStaged) g = s_id[c + lane];
                }
            }
            unsigned nz = __ballot_sync(full, wh != 0u);
            while (nz) {
                const int src = __ffs(nz) - 1;
                nz &= nz - 1;
                const uint32_t bt0 = __shfl_sync(full, t0, src), bwh = __shfl_sync(full, wh, src);
                const uint32_t bg = (kScatter && !kStaged) ? __shfl_sync(full, g, src) : 0u;
                const uint32_t packed = (uint32_t)(c + src) << 16;   // staged: the Gaussian's index in the batch
                const int bw = (int)(bwh & 0x7ffu), bh = (int)((bwh >> 11) & 0xffu);
                if (bw <= 32) {
                    // lane L handles column L % bw of rows L / bw, L / bw + rows, ... (rows = 32 / bw whole rows per step)
                    const uint32_t inv = bwh >> 19;
                    const int ry = (int)(((uint32_t)lane * inv) >> 10), rows = (int)((32u * inv) >> 10);
                    const uint32_t tcol = bt0 + (uint32_t)(lane - ry * bw);
                    for (int y = ry < rows ? ry : 255; y < bh; y += rows) touch(tcol + y * gx, packed, bg);
                } else {
                    for (int y = 0; y < bh; y++) {
                        for (int x = lane; x < bw; x += 32) {
                            touch(bt0 + y * gx + x, packed, bg);
                        }
                    }
                }
                __syncwarp();  // the next Gaussian may touch the same counters
            }
        }
        __syncthreads();  // every warp's counts are final / every warp has staged its ids
        if (!kScatter) {
            // per-warp counts (one byte: a warp owns <= 255 Gaussians) and the block total -> global
            for (int k = threadIdx.x; k < band_tiles; k += 32 * kBinWarps) {
                uint32_t sum = 0, lo = 0, hi = 0;
#pragma unroll
                for (int w = 0; w < kBinWarps; w++) {
                    const uint32_t v = s_cnt[w * band_max + k];
                    if (w < 4) lo |= v << (8 * w); else hi |= v << (8 * (w - 4));
                    sum += v;
                }
                wrow[band_start + k] = make_uint2(lo, hi);
                row[band_start + k] = sum;
            }
        } else if (kStaged) {
            // copy the band out: slot i of the local buffer goes to s_base[tile] + i
            const uint32_t staged = min(n_local, (uint32_t)kCap);
            for (uint32_t i = threadIdx.x; i < staged; i += 32 * kBinWarps) {
                const uint32_t v = s_oid[i];
                const uint32_t pos = s_base[v & 0xffffu] + i;
                if (pos < a.cap) a.point_list[pos] = s_id[v >> 16];
            }
        }
        __syncthreads();  // counters and staging free for the next band
    }
}

// chunk_sums[c][t] = sum of table[b][t] over the batches of chunk c; tile_total[t] += that
__global__ void __launch_bounds__(256) bin_chunk_sums(const uint32_t* __restrict__ table, int nb, int T, int T_pad,
                                                     uint32_t* __restrict__ chunk_sums,
                                                     uint32_t* __restrict__ tile_total) {
    const int t = blockIdx.x * 256 + threadIdx.x, c = blockIdx.y;
    if (t >= T) return;
    const int b0 = c * kBinChunk, b1 = min(nb, b0 + kBinChunk);
    uint32_t s = 0;
#pragma unroll 8
    for (int b = b0; b < b1; b++) s += __ldg(table + (size_t)b * T_pad + t);
    chunk_sums[(size_t)c * T_pad + t] = s;
    if (s) atomicAdd(tile_total + t, s);
}

// exclusive scan of the tile totals -> tile_start (in place) and ranges ((0,0) for empty tiles,
// as the reference's memset + identifyTileRanges leave them)
__global__ void __launch_bounds__(1024) bin_tile_scan(uint32_t* __restrict__ tile_total, int T, uint2* __restrict__ ranges,
                                                     uint32_t cap) {
    __shared__ uint32_t s_warp[32];
    constexpr int kPer = 8;   // consecutive tiles per thread: one block scan per 8192 tiles (dtu: 7500)
    uint32_t carry = 0;
    for (int base = 0; base < T; base += 1024 * kPer) {
        const int t0 = base + threadIdx.x * kPer;
        uint32_t v[kPer], sum = 0;
#pragma unroll
        for (int j = 0; j < kPer; j++) {
            v[j] = t0 + j < T ? tile_total[t0 + j] : 0u;
            sum += v[j];
        }
        uint32_t total;
        const uint32_t inc = block_inclusive_scan(sum, s_warp, total);
        uint32_t start = carry + inc - sum;
#pragma unroll
        for (int j = 0; j < kPer; j++) {
            if (t0 + j < T) {
                tile_total[t0 + j] = start;
                // cap: only the no-sync forward can see more instances than the list holds; it drops them
                ranges[t0 + j] = v[j] ? make_uint2(min(start, cap), min(start + v[j], cap)) : make_uint2(0u, 0u);
            }
            start += v[j];
        }
        carry += total;
    }
}

// table[b][t] <- tile_start[t] + (instances of t in earlier batches)
__global__ void __launch_bounds__(256) bin_apply(uint32_t* __restrict__ table, int nb, int T, int T_pad,
                                                const uint32_t* __restrict__ chunk_sums,
                                                const uint32_t* __restrict__ tile_start) {
    const int t = blockIdx.x * 256 + threadIdx.x, c = blockIdx.y;
    if (t >= T) return;
    uint32_t run = tile_start[t];
    for (int k = 0; k < c; k++) run += __ldg(chunk_sums + (size_t)k * T_pad + t);
    const int b0 = c * kBinChunk, b1 = min(nb, b0 + kBinChunk);
    uint32_t v[kBinChunk];
#pragma unroll
    for (int k = 0; k < kBinChunk; k++) v[k] = b0 + k < b1 ? table[(size_t)(b0 + k) * T_pad + t] : 0u;
#pragma unroll
    for (int k = 0; k < kBinChunk; k++) {
        if (b0 + k < b1) table[(size_t)(b0 + k) * T_pad + t] = run;
        run += v[k];
    }
}

static bool use_radix() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B3GS_BINNING");
        v = (e && !strcmp(e, "radix")) ? 1 : 0;
    }
    return v == 1;
}

struct TileBinLayout {
    int B, nb, nchunks, T, T_pad, band_rows;
    size_t table, wcount, chunk_sums, tile_total, total;
    TileBinLayout(int P, int R, int grid_x, int grid_y) {
        T = grid_x * grid_y;
        T_pad = (T + 3) & ~3;
        B = bin_batch_size(P, R, T_pad);
        band_rows = bin_band_rows(grid_x, grid_y);
        nb = P > 0 ? (P + B - 1) / B : 0;
        nchunks = (nb + kBinChunk - 1) / kBinChunk;
        size_t o = 0;
        table = o;      o = align_up(o + (size_t)nb * T_pad * 4, 256);
        wcount = o;     o = align_up(o + (size_t)nb * T_pad * 8, 256);
        chunk_sums = o; o = align_up(o + (size_t)nchunks * T_pad * 4, 256);
        tile_total = o; o = align_up(o + (size_t)T_pad * 4, 256);
        total = o;
    }
};

static cudaError_t tile_bins(const BinningPhase2Args& a, char* q, cudaStream_t stream) {
    const TileBinLayout L(a.P, a.R, a.grid_x, a.grid_y);
    uint32_t* table = reinterpret_cast<uint32_t*>(q + L.table);
    uint32_t* chunk_sums = reinterpret_cast<uint32_t*>(q + L.chunk_sums);
    uint32_t* tile_total = reinterpret_cast<uint32_t*>(q + L.tile_total);
    // stage the output through shared memory when a (batch, tile) cell holds enough instances to coalesce
    static const int forced_staged = [] { const char* e = getenv("B3GS_BIN_STAGED"); return e ? atoi(e) : -1; }();
    const bool staged = forced_staged >= 0 ? forced_staged != 0 : (double)a.R >= 3.5 * (double)L.nb * (double)L.T;
    const int band_max = std::max(bin_band_max(), (a.grid_x + 1) & ~1), stage_cap = bin_stage_cap(L.B, band_max);
    const size_t smem_count = bin_smem_bytes(L.B, band_max, false, 0),
                 smem_scatter = bin_smem_bytes(L.B, band_max, true, staged ? stage_cap : 0);
    static thread_local int attr_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != attr_dev) {  // opt in to more than 48 KB of dynamic shared memory, once per device
        cudaFuncSetAttribute(tile_bins_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(tile_bins_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(tile_bins_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_dev = dev;
    }
    TileBinArgs k;
    k.P = a.P; k.B = L.B; k.nb = L.nb; k.T = L.T; k.T_pad = L.T_pad; k.grid_x = a.grid_x; k.grid_y = a.grid_y;
    k.band_rows = L.band_rows; k.band_max = band_max; k.stage_cap = stage_cap;
    k.cap = (uint32_t)a.R;
    k.sorted_ids = a.sorted_ids; k.rects = a.rects; k.table = table; k.point_list = a.point_list;
    k.wcount = reinterpret_cast<uint2*>(q + L.wcount);
    cudaError_t e = cudaMemsetAsync(tile_total, 0, (size_t)L.T_pad * 4, stream);
    if (e != cudaSuccess) return e;
    tile_bins_kernel<false, false><<<L.nb, 32 * kBinWarps, smem_count, stream>>>(k);
    const dim3 grid((L.T + 255) / 256, L.nchunks);
    bin_chunk_sums<<<grid, 256, 0, stream>>>(table, L.nb, L.T, L.T_pad, chunk_sums, tile_total);
    bin_tile_scan<<<1, 1024, 0, stream>>>(tile_total, L.T, a.ranges, k.cap);
    bin_apply<<<grid, 256, 0, stream>>>(table, L.nb, L.T, L.T_pad, chunk_sums, tile_total);
    if (staged) tile_bins_kernel<true, true><<<L.nb, 32 * kBinWarps, smem_scatter, stream>>>(k);
    else tile_bins_kernel<true, false><<<L.nb, 32 * kBinWarps, smem_scatter, stream>>>(k);
    count_launch(5);
    return cudaGetLastError();
}

// ---- legacy cub path (A/B only)
__global__ void __launch_bounds__(256) emit_keys64(int P, const float4* __restrict__ records,
                                                  const float* __restrict__ depths, const int* __restrict__ radii,
                                                  const uint32_t* __restrict__ sorted_ids,
                                                  const uint32_t* __restrict__ sorted_offsets,
                                                  uint64_t* __restrict__ keys, uint32_t* __restrict__ values, int grid_x,
                                                  int grid_y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const uint32_t idx = sorted_ids[i];
    const int r = radii[idx];
    if (r <= 0) return;
    uint32_t off = sorted_offsets[i];
    const float4 a = records[(size_t)idx * B3_REC_VEC4];
    int x0, y0, x1, y1;
    tile_rect(a.x, a.y, (float)r, grid_x, grid_y, x0, y0, x1, y1);
    const uint64_t dbits = __float_as_uint(depths[idx]);
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
            keys[off] = ((uint64_t)(uint32_t)(y * grid_x + x) << 32) | dbits;
            values[off] = idx;
            off++;
        }
}
__global__ void __launch_bounds__(256) tile_ranges_u64(int R, const uint64_t* __restrict__ keys,
                                                      uint2* __restrict__ ranges) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R) return;
    const uint32_t cur = (uint32_t)(keys[idx] >> 32);
    if (idx == 0) ranges[cur].x = 0;
    else {
        const uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
        if (cur != prev) { ranges[prev].y = idx; ranges[cur].x = idx; }
    }
    if (idx == R - 1) ranges[cur].y = R;
}
static uint32_t higher_msb(uint32_t n) {  // rasterizer_impl.cu:35-50
    uint32_t msb = sizeof(n) * 4, step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}
static size_t cub_sort_temp_bytes(int R) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, R);
    return bytes;
}
static bool use_cub() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B3GS_BINNING");
        v = (e && !strcmp(e, "cub")) ? 1 : 0;
    }
    return v == 1;
}

bool binning_uses_tile_bins(int P, int grid_x, int grid_y) {
    // B3GS_BINNING=bins forces the direct path wherever it is applicable, =radix / =cub disable it
    static const bool forced = [] { const char* e = getenv("B3GS_BINNING"); return e && !strcmp(e, "bins"); }();
    const long long T = (long long)grid_x * grid_y;
    return !use_cub() && !use_radix() && P > 0 && grid_x <= kBandTilesMax && grid_y < 65536 && T < (1ll << 24) &&
           (forced || T >= kMinBinTiles);
}

size_t binning_phase2_scratch_bytes(int P, int R, int grid_x, int grid_y) {
    const size_t r = (size_t)(R > 0 ? R : 0);
    if (binning_uses_tile_bins(P, grid_x, grid_y)) return TileBinLayout(P, R, grid_x, grid_y).total;
    if (use_cub())
        return align_up(r * 8, 256) * 2 + align_up(r * 4, 256) + align_up(cub_sort_temp_bytes(R), 256) + 256;
    // keysA[R] keysB[R] idsB[R] hist
    return align_up(r * 4, 256) * 3 + align_up(radix_scratch_elems(R) * 4, 256);
}

// emit + stable sort by tile id + ranges.  scratch: keysA[R] keysB[R] (KeyT) idsB[R] hist
template <typename KeyT>
static cudaError_t tile_sort(const BinningPhase2Args& a, char* q, cudaStream_t stream) {
    const int T = a.grid_x * a.grid_y;
    const size_t r = (size_t)a.R;
    KeyT* keysA = reinterpret_cast<KeyT*>(q); q += align_up(r * sizeof(KeyT), 256);
    KeyT* keysB = reinterpret_cast<KeyT*>(q); q += align_up(r * sizeof(KeyT), 256);
    uint32_t* idsB = reinterpret_cast<uint32_t*>(q);  q += align_up(r * 4, 256);
    uint32_t* hist = reinterpret_cast<uint32_t*>(q);
    uint32_t* totals = hist + (size_t)kRadixBins * radix_blocks(a.R);
    // ping-pong so that the LAST pass writes the ids into point_list
    const int passes = tile_passes(T);
    uint32_t* ids_cur = (passes & 1) ? idsB : a.point_list;  // where emission writes
    uint32_t* ids_oth = (passes & 1) ? a.point_list : idsB;
    KeyT* keys_cur = keysA;
    KeyT* keys_oth = keysB;
    emit_instances<KeyT><<<(a.P + 255) / 256, 256, 0, stream>>>(a.P, a.sorted_ids, a.sorted_offsets, a.records, a.radii,
                                                               keys_cur, ids_cur, a.grid_x, a.grid_y);
    count_launch();
    for (int p = 0; p < passes; p++) {
        radix_pass<KeyT>(keys_cur, ids_cur, keys_oth, ids_oth, hist, totals, a.R, 8 * p, false, stream);
        KeyT* t = keys_cur; keys_cur = keys_oth; keys_oth = t;
        uint32_t* u = ids_cur; ids_cur = ids_oth; ids_oth = u;
    }
    // ids_cur == a.point_list by construction
    tile_ranges_k<KeyT><<<(a.R + 2047) / 2048, 256, 0, stream>>>(a.R, keys_cur, a.ranges);
    count_launch();
    return cudaGetLastError();
}

cudaError_t run_binning_phase2(const BinningPhase2Args& a, cudaStream_t stream) {
    const int T = a.grid_x * a.grid_y;
    const bool bins = binning_uses_tile_bins(a.P, a.grid_x, a.grid_y);
    cudaError_t e = cudaSuccess;
    if (!bins || a.R <= 0) e = cudaMemsetAsync(a.ranges, 0, (size_t)T * sizeof(uint2), stream);  // the direct binning writes every range itself
    if (e != cudaSuccess) return e;
    if (a.R <= 0) return cudaSuccess;
    const size_t r = (size_t)a.R;
    char* q = a.scratch;
    if (use_cub()) {
        uint64_t* keys_unsorted = reinterpret_cast<uint64_t*>(q); q += align_up(r * 8, 256);
        uint64_t* keys_sorted = reinterpret_cast<uint64_t*>(q);   q += align_up(r * 8, 256);
        uint32_t* values_unsorted = reinterpret_cast<uint32_t*>(q); q += align_up(r * 4, 256);
        size_t temp_bytes = cub_sort_temp_bytes(a.R);
        emit_keys64<<<(a.P + 255) / 256, 256, 0, stream>>>(a.P, a.records, a.depths, a.radii, a.sorted_ids,
                                                          a.sorted_offsets, keys_unsorted, values_unsorted, a.grid_x,
                                                          a.grid_y);
        e = cub::DeviceRadixSort::SortPairs(q, temp_bytes, keys_unsorted, keys_sorted, values_unsorted, a.point_list,
                                            a.R, 0, 32 + (int)higher_msb((uint32_t)T), stream);
        if (e != cudaSuccess) return e;
        tile_ranges_u64<<<(a.R + 255) / 256, 256, 0, stream>>>(a.R, keys_sorted, a.ranges);
        count_launch(10);
        return cudaGetLastError();
    }
    if (bins) return tile_bins(a, q, stream);
    return T <= 65536 ? tile_sort<uint16_t>(a, q, stream) : tile_sort<uint32_t>(a, q, stream);
}

}  // namespace b3
