// kernels.h — host-side launch interface between api.cu and the kernel TUs.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace b3 {

void count_launch(unsigned n = 1);

// ---------------------------------------------------------------- K1 / K10
// raw != 0 (the raw-parameter entry, b3gs_forward_raw / b3gs_backward_raw): `scales`,
// `rotations`, `opacities` hold the reference's RAW parameters and are activated on load
// (exp, normalize, sigmoid), `shs` is f_dc (P,1,3) and `shs_rest` is f_rest (P,M-1,3).
struct PreprocessArgs {
    int P, D, M;
    int raw;
    const float* shs_rest;
    const float* means3D;
    const float* scales;
    float scale_modifier;
    const float* rotations;
    const float* opacities;
    const float* shs;
    const float* cov3D_precomp;
    const float* colors_precomp;
    const float* viewmatrix;
    const float* projmatrix;
    const float* cam_pos;
    int W, H;
    float tan_fovx, tan_fovy, focal_x, focal_y;
    int grid_x, grid_y;
    int prefiltered;
    int* radii;
    float4* records;
    float* depths;
    uint32_t* tiles_touched;
    uint8_t* clamped;
    uint2* rects;            // tile rectangle {x0 | y0 << 16, width | height << 16}; {0, 0} when culled
    uint32_t* num_rendered;  // device counter, zeroed by the caller; += sum(tiles_touched)
};
void launch_preprocess(const PreprocessArgs& a, cudaStream_t stream);
void launch_mark_visible(int P, const float* means3D, const float* viewmatrix, unsigned char* present,
                         cudaStream_t stream);

// ---------------------------------------------------------------- binning (K2-K5)
// Phase 1 needs only P-sized state and is enqueued BEFORE the host reads R, so the GPU
// sorts by depth while the host waits for the count and allocates the R-sized blob.
struct BinningPhase1Args {
    int P;
    const float* depths;            // sort key: float bits
    const uint32_t* tiles_touched;
    const uint32_t* key_bits;       // [0] OR, [1] AND of the visible Gaussians' depth keys
    uint32_t* sorted_ids;           // out: Gaussian ids in (depth bits, id) order, P
    uint32_t* sorted_offsets;       // out: exclusive scan of tiles_touched in that order, P
    int need_offsets;               // 0: skip that scan (the direct tile binning does not use it)
    char* scratch;
};
size_t binning_phase1_scratch_bytes(int P);
cudaError_t run_binning_phase1(const BinningPhase1Args& a, cudaStream_t stream);

struct BinningPhase2Args {
    int P, R;                       // R: instances, or (count_unknown) the capacity of point_list
    int count_unknown;              // != 0: the true count may exceed R; positions are clamped to it
    int grid_x, grid_y;
    const float4* records;
    const float* depths;
    const int* radii;
    const uint32_t* sorted_ids;
    const uint32_t* sorted_offsets;
    const uint2* rects;             // per-Gaussian tile rectangles written by the preprocess
    uint32_t* point_list;           // out: sorted Gaussian ids, R
    uint2* ranges;                  // out: per-tile [start,end), T
    char* scratch;
};
// Which phase-2 algorithm a forward with these sizes uses: true = direct tile binning
// (scratch is a function of P and the tile grid), false = emit + radix passes (scratch is a
// function of R).  Phase 1 must produce sorted_offsets only for the latter.
bool binning_uses_tile_bins(int P, int grid_x, int grid_y);
size_t binning_phase2_scratch_bytes(int P, int R, int grid_x, int grid_y);
cudaError_t run_binning_phase2(const BinningPhase2Args& a, cudaStream_t stream);

// Stable LSD radix sort of n (key, original index) pairs on key bits [0, bits); returns
// pointers (inside scratch) to the sorted keys and the permutation.
size_t sort_keys_iota_scratch_bytes(int n);
cudaError_t sort_keys_iota_u32(int n, const uint32_t* keys, int bits, char* scratch, const uint32_t** keys_sorted,
                               const uint32_t** ids_sorted, cudaStream_t stream);

// ---------------------------------------------------------------- K6 / K7
struct CompositeFwdArgs {
    int W, H, grid_x, grid_y;
    const uint2* ranges;
    const uint32_t* point_list;
    const float4* records;
    const float* background;
    float* out_color;
    float* out_depth;
    float* out_alpha;
    uint32_t* n_contrib;
};
void launch_composite_forward(const CompositeFwdArgs& a, cudaStream_t stream);

struct CompositeBwdArgs {
    int W, H, grid_x, grid_y;
    int P, R;                       // Gaussians and tile instances (kernel-shape heuristic only)
    const uint2* ranges;
    const uint32_t* point_list;
    const float4* records;          // a/b from the forward; colours may be overridden
    const float* colors_override;   // colors_precomp (P,3) or nullptr -> records.c.rgb
    const float* background;
    const float* alphas;            // forward out_alpha
    const uint32_t* n_contrib;
    const float* dL_dpix;           // [3,H,W]
    const float* dL_dpix_depth;     // [H,W]
    const float* dL_dalphas;        // [H,W]
    float* grads;                   // [P, B3_GRAD_STRIDE], zeroed by the caller
};
void launch_composite_backward(const CompositeBwdArgs& a, cudaStream_t stream);
void set_backward_pixels(int n);  // 0 = automatic

// ---------------------------------------------------------------- K8 + K9 fused
struct PreBackwardArgs {
    int P, D, M;
    int first, count;        // set by the launcher: the Gaussian range of this launch
    int raw;                 // as PreprocessArgs::raw; gradients are then w.r.t. the raw parameters
    const float* shs_rest;
    const float* opacities;  // raw mode only (sigmoid'); unused otherwise
    float* dL_dsh_rest;      // raw mode: dL_dsh is d f_dc, this is d f_rest
    const float* means3D;
    const int* radii;
    const float* shs;
    const uint8_t* clamped;
    const float* scales;
    const float* rotations;
    float scale_modifier;
    const float* cov3D_precomp;
    const float* viewmatrix;
    const float* projmatrix;
    const float* campos;
    float focal_x, focal_y, tan_fovx, tan_fovy;
    const float* grads;  // packed accumulator from K7
    float* dL_dmean2D;   // [P,3]
    float* dL_dconic;    // [P,4]
    float* dL_dopacity;  // [P]
    float* dL_dcolor;    // [P,3]
    float* dL_ddepth;    // [P]
    float* dL_dmean3D;   // [P,3]
    float* dL_dcov3D;    // [P,6]
    float* dL_dsh;       // [P,M,3] or nullptr
    float* dL_dscale;    // [P,3]
    float* dL_drot;      // [P,4]
    int accumulate;      // != 0: dL_dmean3D, dL_dsh, dL_dopacity, dL_dscale, dL_drot are added to, not overwritten
    int stage;           // set by the launcher: strided outputs staged through shared memory
};
void launch_preprocess_backward(const PreBackwardArgs& a, cudaStream_t stream, int first = 0, int count = -1);

// ---------------------------------------------------------------- exchange (peer_reduce.cu)
// One chunk of the data-parallel exchange: all-reduce up to six float ranges of the symmetric
// bucket (the five gradient segments of one Gaussian range), barriers inside the kernel.
struct ExchangeRanges {
    int n;
    size_t start4[6], len4[6];      // in float4 units, relative to the bucket base
};
struct ExchangePlan;                // opaque: peers, flags, epoch, side stream, events
cudaError_t exchange_chunk(ExchangePlan* plan, const ExchangeRanges& r, float scale, cudaStream_t stream);
cudaStream_t exchange_stream(ExchangePlan* plan);
cudaEvent_t exchange_event(ExchangePlan* plan, int i);
float* exchange_base(ExchangePlan* plan);

}  // namespace b3
