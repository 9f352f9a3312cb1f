/*
 * b3gs.h — C-ABI of the B200-native differentiable 3D Gaussian Splatting rasterizer.
 *
 * This is the drop-in boundary for the hot path of hanl2010/Binocular3DGS: the
 * differentiable rasterizer the reference reaches through
 *   submodules/diff-gaussian-rasterization/cuda_rasterizer/rasterizer.h:20-90
 *   (CudaRasterizer::Rasterizer::{markVisible, forward, backward})
 * and binds to Python in
 *   submodules/diff-gaussian-rasterization/rasterize_points.cu:35-229, ext.cpp:15-18.
 *
 * Every entry point below takes plain device pointers, sizes and a CUDA stream —
 * no torch types — so the same shared library can be bound from ctypes (what
 * binocular3dgs_b200/_backend.py does), pybind, cgo or JNI.  Argument order and
 * meaning follow the reference interface each function replaces; the only
 * additions are the trailing `stream` (the reference uses the legacy default
 * stream everywhere, rasterizer_impl.cu:148,290,315) and the explicit
 * `num_rendered` out-parameter / error return (the reference returns the count
 * and throws std::runtime_error).
 *
 * All pointers are DEVICE pointers unless stated.  All floating point is FP32.
 * "Optional" pointers follow the reference's null convention
 * (rasterize_points.cu:96-115: empty tensors arrive as nullptr).
 */
#ifndef B3GS_H_INCLUDED
#define B3GS_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define B3GS_API __attribute__((visibility("default")))
#else
#define B3GS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Tile geometry is part of observable behaviour (config.h:15-17): it fixes
 * tiles_touched, the sort keys and the per-tile ranges. */
#define B3GS_TILE_X 16
#define B3GS_TILE_Y 16
#define B3GS_NUM_CHANNELS 3

/* Error codes (negative). 0 = success. */
#define B3GS_OK 0
#define B3GS_ERR_INVALID_ARGUMENT (-1)
#define B3GS_ERR_CUDA (-2)
#define B3GS_ERR_ALLOC (-3)

/*
 * Resize callback, the C form of the reference's
 *   std::function<char*(size_t)>  (rasterizer.h:32-34, rasterize_points.cu:27-33).
 * Must return a device pointer to at least `bytes` bytes, aligned to >= 128 B,
 * that stays valid until the matching b3gs_backward call has been enqueued.
 * `user` is passed back verbatim.  Returning NULL for bytes > 0 aborts the call
 * with B3GS_ERR_ALLOC.
 */
typedef void* (*b3gs_resize_fn)(void* user, size_t bytes);

typedef struct b3gs_buffer {
    b3gs_resize_fn resize;
    void* user;
} b3gs_buffer;

/*
 * Forward rasterization.  Replaces CudaRasterizer::Rasterizer::forward
 * (rasterizer.h:31-58, rasterizer_impl.cu:197-339).
 *
 *   geometry/binning/image : the three opaque state blobs (GeometryState,
 *       BinningState, ImageState in the reference, rasterizer_impl.h:22-73).  Their
 *       layout is private to this library but is a pure function of (P), (R) and
 *       (width,height) respectively, so b3gs_backward can re-derive it.
 *   P, D, M      : #Gaussians, active SH degree (0..3), SH coefficients per Gaussian
 *                  in `shs` (0 when colours are precomputed).
 *   background   : float[3].
 *   means3D      : float[P,3].          shs : float[P,M,3] or NULL.
 *   colors_precomp : float[P,3] or NULL (exactly one of shs / colors_precomp).
 *   opacities    : float[P].            scales : float[P,3] or NULL.
 *   rotations    : float[P,4] (r,x,y,z; NOT normalised here, forward.cu:127) or NULL.
 *   cov3D_precomp: float[P,6] or NULL (exactly one of (scales,rotations) / cov3D).
 *   viewmatrix, projmatrix : float[16], transposed (column-major) convention of
 *                  scene/cameras.py:55-57.   cam_pos : float[3].
 *   out_color    : float[3,H,W]; out_depth, out_alpha : float[H,W]; radii : int[P].
 *                  Outputs need NOT be pre-zeroed (the reference requires zeros,
 *                  rasterize_points.cu:68-71; here every element is written).
 *   prefiltered  : as the reference (auxiliary.h:156-160): a culled point traps.
 *   debug        : synchronise and check after every kernel (auxiliary.h:166-173).
 *   stream       : cudaStream_t to enqueue on.
 *   num_rendered : HOST int, receives R = number of (Gaussian,tile) instances.
 *
 * The call blocks the host once, on `stream`, to read R (the reference blocks the
 * whole device with cudaMemcpy, rasterizer_impl.cu:282).
 */
B3GS_API int b3gs_forward(
    b3gs_buffer geometry, b3gs_buffer binning, b3gs_buffer image,
    int P, int D, int M,
    const float* background,
    int width, int height,
    const float* means3D,
    const float* shs,
    const float* colors_precomp,
    const float* opacities,
    const float* scales,
    float scale_modifier,
    const float* rotations,
    const float* cov3D_precomp,
    const float* viewmatrix,
    const float* projmatrix,
    const float* cam_pos,
    float tan_fovx, float tan_fovy,
    int prefiltered,
    float* out_color,
    float* out_depth,
    float* out_alpha,
    int* radii,
    int debug,
    void* stream,
    int* num_rendered);

/*
 * Forward WITHOUT the host synchronisation (an addition; the reference blocks the whole
 * device at rasterizer_impl.cu:282 and b3gs_forward blocks the calling thread once).
 * Same arguments as b3gs_forward minus `debug`, plus
 *   capacity : how many (Gaussian, tile) instances the binning blob shall hold — the
 *              caller's estimate, e.g. 1.5 x the largest R seen for this (P, width, height);
 *   ticket   : HOST int, receives a ticket for b3gs_count_wait.
 * The call only enqueues work.  R is accumulated on the device and copied to a pinned slot
 * behind the ticket; instances beyond `capacity` are DROPPED (tile ranges are clamped to the
 * list), so the outputs are exact iff R <= capacity.  The caller checks that with
 * b3gs_count_wait — typically when it is about to enqueue the backward, by which time the
 * count has long landed — and, if R > capacity, re-runs b3gs_forward (exact) before using the
 * outputs for anything that matters.  Pass the R returned by b3gs_count_wait to b3gs_backward.
 * b3gs_forward_nosync_supported: 1 when the direct tile binning serves these sizes (its
 * scratch does not depend on R); otherwise use b3gs_forward.
 * b3gs_count_wait blocks until the count behind `ticket` has landed and returns it.  A ticket
 * expires after 256 later b3gs_forward_nosync calls of the process (error, not a hang).
 */
B3GS_API int b3gs_forward_nosync_supported(int P, int width, int height);
B3GS_API int b3gs_forward_nosync(
    b3gs_buffer geometry, b3gs_buffer binning, b3gs_buffer image,
    int P, int D, int M, const float* background, int width, int height, const float* means3D, const float* shs,
    const float* colors_precomp, const float* opacities, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
    const float* cam_pos, float tan_fovx, float tan_fovy, int prefiltered, float* out_color, float* out_depth,
    float* out_alpha, int* radii, void* stream, int capacity, int* ticket);
B3GS_API int b3gs_count_wait(int ticket, int* num_rendered);

/*
 * Backward.  Replaces CudaRasterizer::Rasterizer::backward
 * (rasterizer.h:60-89, rasterizer_impl.cu:343-447).  Argument order follows the
 * reference.  `alphas` is the forward's out_alpha.  Gradient outputs:
 *   dL_dmean2D float[P,3] (x,y used; z written 0), dL_dconic float[P,4] (x,y,-,w),
 *   dL_dopacity float[P], dL_dcolor float[P,3], dL_ddepth float[P],
 *   dL_dmean3D float[P,3], dL_dcov3D float[P,6], dL_dsh float[P,M,3] (or NULL when
 *   M==0), dL_dscale float[P,3], dL_drot float[P,4].
 * Unlike the reference (rasterize_points.cu:158-167) the outputs need NOT be
 * pre-zeroed: every element is written by this call.  dL_dpix_depth and dL_dalphas may
 * be NULL (that image received no gradient: treated as zeros, and not read).  The four
 * intermediates dL_dconic, dL_dcolor, dL_ddepth, dL_dcov3D may be NULL when the caller has no
 * use for them (dL_dcolor is only observable with colors_precomp, dL_dcov3D with
 * cov3D_precomp, the other two never leave rasterize_points.cu:161-162): not written then.
 */
B3GS_API int b3gs_backward(
    int P, int D, int M, int R,
    const float* background,
    int width, int height,
    const float* means3D,
    const float* shs,
    const float* colors_precomp,
    const float* alphas,
    const float* scales,
    float scale_modifier,
    const float* rotations,
    const float* cov3D_precomp,
    const float* viewmatrix,
    const float* projmatrix,
    const float* campos,
    float tan_fovx, float tan_fovy,
    const int* radii,
    char* geom_buffer,
    char* binning_buffer,
    char* image_buffer,
    const float* dL_dpix,
    const float* dL_dpix_depth,
    const float* dL_dalphas,
    float* dL_dmean2D,
    float* dL_dconic,
    float* dL_dopacity,
    float* dL_dcolor,
    float* dL_ddepth,
    float* dL_dmean3D,
    float* dL_dcov3D,
    float* dL_dsh,
    float* dL_dscale,
    float* dL_drot,
    int debug,
    void* stream);

/*
 * b3gs_backward with option flags (an addition; b3gs_backward == flags 0):
 *   B3GS_BWD_ACCUMULATE  the five parameter gradients a trainer sums over the views of one
 *       step — dL_dmean3D, dL_dsh, dL_dopacity, dL_dscale, dL_drot — are ADDED to the values
 *       already in those buffers instead of overwriting them (the reference's binocular step
 *       renders two views before one optimizer step, train.py:100,128; with this flag the
 *       second view's backward accumulates straight into the data-parallel exchange bucket).
 *       The per-view outputs (dL_dmean2D, dL_dconic, dL_dcolor, dL_ddepth, dL_dcov3D) are
 *       written as usual.
 */
#define B3GS_BWD_ACCUMULATE 1u
B3GS_API int b3gs_backward_flags(
    unsigned flags,
    int P, int D, int M, int R, const float* background, int width, int height, const float* means3D,
    const float* shs, const float* colors_precomp, const float* alphas, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
    const float* campos, float tan_fovx, float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer,
    char* image_buffer, const float* dL_dpix, const float* dL_dpix_depth, const float* dL_dalphas, float* dL_dmean2D,
    float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_ddepth, float* dL_dmean3D, float* dL_dcov3D,
    float* dL_dsh, float* dL_dscale, float* dL_drot, int debug, void* stream);

/*
 * ---- The raw-parameter entry (an addition) --------------------
 * The reference evaluates exp / sigmoid / normalize / cat on the raw parameters before every
 * render (scene/gaussian_model.py:95-115, called from gaussian_renderer/__init__.py:54-83)
 * and autograd runs their duals after every backward: ten elementwise kernels and
 * 2 x (11 + 3M) floats per Gaussian of extra traffic each way.  These two entry points take
 * the RAW parameters and fuse the activations into the preprocess and its backward:
 *   xyz float[P,3]; f_dc float[P,1,3]; f_rest float[P,M-1,3] (NULL when M == 1);
 *   opacity_raw float[P]; scaling_raw float[P,3]; rotation_raw float[P,4]
 *   means3D = xyz, shs = cat(f_dc, f_rest), opacities = sigmoid(opacity_raw),
 *   scales = exp(scaling_raw), rotations = rotation_raw / max(|rotation_raw|, 1e-12).
 * Results are bit-identical to b3gs_activate_forward followed by b3gs_forward (the same
 * device functions), and the gradients are those of the raw parameters.
 * b3gs_forward_raw: capacity < 0 -> exact path, *num_rendered_or_ticket receives R;
 *   capacity >= 1 -> no host wait (see b3gs_forward_nosync), it receives the ticket.
 * b3gs_backward_raw: flags as b3gs_backward_flags (B3GS_BWD_ACCUMULATE applies to dL_dxyz,
 *   dL_df_dc, dL_df_rest, dL_dopacity_raw, dL_dscaling_raw, dL_drotation_raw).  The
 *   intermediates (dL_dconic, dL_dcolor, dL_ddepth, dL_dcov3D) are not materialised.
 */
B3GS_API int b3gs_forward_raw(
    b3gs_buffer geometry, b3gs_buffer binning, b3gs_buffer image, int P, int D, int M, const float* background,
    int width, int height, const float* xyz, const float* f_dc, const float* f_rest, const float* opacity_raw,
    const float* scaling_raw, float scale_modifier, const float* rotation_raw, const float* viewmatrix,
    const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy, float* out_color,
    float* out_depth, float* out_alpha, int* radii, void* stream, int capacity, int* num_rendered_or_ticket);
B3GS_API int b3gs_backward_raw(
    unsigned flags, int P, int D, int M, int R, const float* background, int width, int height, const float* xyz,
    const float* f_dc, const float* f_rest, const float* opacity_raw, const float* scaling_raw, float scale_modifier,
    const float* rotation_raw, const float* alphas, const float* viewmatrix, const float* projmatrix,
    const float* campos, float tan_fovx, float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer,
    char* image_buffer, const float* dL_dpix, const float* dL_dpix_depth, const float* dL_dalphas, float* dL_dmean2D,
    float* dL_dxyz, float* dL_df_dc, float* dL_df_rest, float* dL_dopacity_raw, float* dL_dscaling_raw,
    float* dL_drotation_raw, void* stream);

/* Visibility mask.  Replaces Rasterizer::markVisible (rasterizer.h:24-29,
 * rasterizer_impl.cu:54-66,141-153): present[i] = (z_view > 0.2). `present` is
 * bool[P] (1 byte each). */
B3GS_API int b3gs_mark_visible(
    int P,
    const float* means3D,
    const float* viewmatrix,
    const float* projmatrix,
    unsigned char* present,
    void* stream);

/* Sizes of the three opaque blobs.  Pure functions of their arguments.
 * b3gs_binning_bytes(R) is the part b3gs_backward reads (the sorted id list, at offset 0);
 * b3gs_binning_bytes_forward is what b3gs_forward asks the binning callback for: that list
 * plus the forward's scratch, which depends on P and the tile grid for the direct tile
 * binning and on R for the radix fallback (binning.cu). */
B3GS_API size_t b3gs_geometry_bytes(int P);
B3GS_API size_t b3gs_binning_bytes(int R);
B3GS_API size_t b3gs_binning_bytes_forward(int P, int R, int width, int height);
B3GS_API size_t b3gs_image_bytes(int width, int height);

/*
 * Introspection for parity tests (the analogue of slicing the reference's blobs,
 * SURVEY.md §8c): byte offsets of named arrays inside the blobs.  Returns the
 * offset, or (size_t)-1 for an unknown name.
 *   geometry: "depths" f32[P], "tiles_touched" u32[P], "point_offsets" u32[P],
 *             "records" f32[P,12] = {x, y, cull_tau, 0 | conic_x, conic_y, conic_z,
 *             opacity | r, g, b, depth}, "clamped" u8[P] (bit c = channel c clamped),
 *             "rects" u32[P,2] = {x0 | y0 << 16, width | height << 16} in tiles ({0,0}: culled)
 *   binning:  "point_list" u32[R]
 *   image:    "n_contrib" u32[H*W], "ranges" u32[T,2]
 */
B3GS_API size_t b3gs_geometry_offset(int P, const char* name);
B3GS_API size_t b3gs_binning_offset(int R, const char* name);
B3GS_API size_t b3gs_image_offset(int width, int height, const char* name);

/*
 * Optional per-stage device timing, used by bench.py for the roofline figures.  When
 * enabled every stage (preprocess, scan, binning, composite_forward, grad_zero,
 * composite_backward, preprocess_backward) is bracketed by CUDA events on the caller's
 * stream.  b3gs_profile_read() waits for the recorded events, writes the summed
 * device milliseconds and the number of calls per stage into ms_total[n] / calls[n],
 * clears the record and returns the number of (stage, call) samples consumed.
 */
B3GS_API void b3gs_profile_enable(int on);
B3GS_API int b3gs_profile_num_stages(void);
B3GS_API const char* b3gs_profile_stage_name(int i);
B3GS_API int b3gs_profile_read(double* ms_total, unsigned long long* calls, int n);

/*
 * ---- fused photometric loss ---------------------------------
 * loss = (1 - lambda) * mean|img1 - img2| + lambda * (1 - mean SSIM(img1, img2)), the loss
 * the reference evaluates on the rasterizer's output every iteration (train.py:146-147;
 * utils/loss_utils.py:18-21 l1_loss, :36-66 ssim/_ssim with the 11x11 sigma=1.5 window of
 * :26-34, conv2d zero padding 5).  Images are float[C,H,W] (a batch is folded into C).
 *
 * b3gs_photometric_forward writes the three per-pixel partial derivatives of the SSIM
 * map (w.r.t. mu1, E[x^2], E[xy]) needed by the backward, optionally the SSIM map itself
 * (NULL to skip), sums[0] = sum of the SSIM map, sums[1] = sum |img1 - img2| (sums: device
 * double[3], zeroed by the call; [2] is the kernel's block counter) and, if loss_out is
 * not NULL, *loss_out = k_const + k_ssim * sums[0] + k_l1 * sums[1] (device float, written
 * by the last block: the loss value never needs a separate kernel).
 * b3gs_photometric_backward writes dL/dimg1 = g * (k_ssim * dSSIMsum/dimg1 +
 * k_l1 * sign(img1 - img2)) with g = upstream[0], a DEVICE float so the upstream gradient
 * never has to visit the host (for the combined loss: k_ssim = -lambda/(CHW),
 * k_l1 = (1-lambda)/(CHW)).  Gradient flows to img1 only.
 */
B3GS_API int b3gs_photometric_forward(int C, int H, int W, const float* img1, const float* img2, float* dm_dmu1,
                                      float* dm_dsigma1_sq, float* dm_dsigma12, float* ssim_map, double* sums,
                                      float k_const, float k_ssim, float k_l1, float* loss_out, void* stream);
B3GS_API int b3gs_photometric_backward(int C, int H, int W, const float* img1, const float* img2,
                                       const float* dm_dmu1, const float* dm_dsigma1_sq, const float* dm_dsigma12,
                                       const float* upstream, float k_ssim, float k_l1, float* dL_dimg1,
                                       void* stream);

/*
 * ---- binocular-consistency loss -------------------------------
 * What train.py:122-136 evaluates after shift_cam_start on the two renders of a stereo
 * pair:   disparity = k_disp / (depth + 1e-5),  k_disp = focal_x * (-trans_dist);
 *         warped = inverse_warp_images(shifted, disparity), mask = the same warp of ones
 *         (utils/graphics_utils.py:80-125);
 *         loss = l1_loss(warped, gt, mask) + w * SmoothLoss(disparity * mask, gt)
 *         (utils/loss_utils.py:18-21, :68-91; w = 0.05).
 * Images are float[3,H,W], depth float[H,W]; H, W >= 3.
 *
 * b3gs_binocular_forward zeroes and fills sums (device double[4]): [0] = sum over
 * 3*H*W of |warped*mask - gt*mask|, [1] / [2] = sums over (H-2)(W-2) of the x / y
 * smoothness terms, [3] the kernel's block counter; if loss_out is not NULL the last block
 * writes *loss_out = k_l1 * sums[0] + k_sm * (sums[1] + sums[2]) (device float), i.e. the
 * loss for k_l1 = 1/(3HW), k_sm = w/((H-2)(W-2)).
 * b3gs_binocular_backward recomputes from the same inputs (nothing is saved) and writes
 * dL/dshifted (zeroed by the call, accumulated with float REDs) and dL/ddepth;
 * upstream is a DEVICE float[1] holding the upstream gradient g (it never visits the
 * host); k_l1 = 1/(3HW) and k_sm = w/((H-2)(W-2)) scale the two terms.  Gradient flows to `shifted` and `depth` only.
 */
B3GS_API int b3gs_binocular_forward(int H, int W, const float* shifted, const float* depth, const float* gt,
                                    float k_disp, double* sums, float k_l1, float k_sm, float* loss_out,
                                    void* stream);
B3GS_API int b3gs_binocular_backward(int H, int W, const float* shifted, const float* depth, const float* gt,
                                     float k_disp, const float* upstream, float k_l1, float k_sm, float* dL_dshifted,
                                     float* dL_ddepth, void* stream);

/* The constituents as stand-alone operators (drop-in for the reference's Python API).
 * b3gs_warp_*: inverse_warp_images for one image float[C,H,W] and one disparity map
 * float[H,W] (utils/graphics_utils.py:80-125); the backward zeroes dL_dimage and
 * accumulates into it; dL_ddisparity may be NULL.
 * b3gs_smooth_*: SmoothLoss.forward(disparity float[H,W], image float[3,H,W])
 * (utils/loss_utils.py:68-91): sums[2] (device doubles, zeroed by the call) = x / y sums
 * over (H-2)(W-2); the backward writes dL/ddisparity = upstream[0] * k *
 * d(sums[0]+sums[1])/d(disparity) with upstream a DEVICE float[1]. */
B3GS_API int b3gs_warp_forward(int C, int H, int W, const float* image, const float* disparity, float* warped,
                               void* stream);
B3GS_API int b3gs_warp_backward(int C, int H, int W, const float* image, const float* disparity,
                                const float* dL_dwarped, float* dL_dimage, float* dL_ddisparity, void* stream);
B3GS_API int b3gs_smooth_forward(int H, int W, const float* disparity, const float* image, double* sums,
                                 void* stream);
B3GS_API int b3gs_smooth_backward(int H, int W, const float* disparity, const float* image, const float* upstream,
                                  float k, float* dL_ddisparity, void* stream);

/*
 * ---- per-step parameter plumbing ------------------------------
 * The elementwise work scene/gaussian_model.py runs around the rasterizer every
 * iteration, one streaming kernel each.
 *
 * b3gs_activate_forward: raw parameters -> rasterizer inputs (gaussian_model.py:95-115):
 *   shs[P,M,3] = cat(f_dc[P,1,3], f_rest[P,M-1,3]); opacities = sigmoid(opacity_raw[P]);
 *   scales = exp(scaling_raw[P,3]); rotations = normalize(rotation_raw[P,4]) (eps 1e-12).
 * b3gs_activate_backward: the duals, recomputed from the raw parameters; any g_* input
 *   may be NULL (that attribute received no gradient), its output is then left untouched.
 * b3gs_adam_multi: torch.optim.Adam's update (gaussian_model.py:154-163: lr per group,
 *   betas (0.9, 0.999), eps 1e-15, no weight decay / amsgrad) for up to
 *   B3GS_ADAM_MAX_TENSORS tensors in ONE launch.  step_size = lr / (1 - beta1^t) and
 *   inv_bias_correction2_sqrt = 1 / sqrt(1 - beta2^t) are formed by the caller in double,
 *   as torch does.  Updates param, exp_avg, exp_avg_sq in place.
 * b3gs_opacity_decay: opacity_raw <- logit(sigmoid(opacity_raw) * factor)
 *   (gaussian_model.py:307-309).
 * b3gs_densify_stats: for Gaussians with radii > 0 (train.py:170-171,
 *   gaussian_model.py:409-411): xyz_gradient_accum += ||viewspace_grad[:, :2]||,
 *   denom += 1, max_radii2D = max(max_radii2D, radii) (max_radii2D may be NULL).
 *   viewspace_grad is float[P,3] (the rasterizer's dL/dmeans2D).
 */
#define B3GS_ADAM_MAX_TENSORS 8
typedef struct B3gsAdamTensor {
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    size_t n;
    float step_size;
    float inv_bias_correction2_sqrt;
} B3gsAdamTensor;

B3GS_API int b3gs_activate_forward(int P, int M, const float* f_dc, const float* f_rest, const float* opacity_raw,
                                   const float* scaling_raw, const float* rotation_raw, float* shs, float* opacities,
                                   float* scales, float* rotations, void* stream);
B3GS_API int b3gs_activate_backward(int P, int M, const float* opacity_raw, const float* scaling_raw,
                                    const float* rotation_raw, const float* g_shs, const float* g_opacities,
                                    const float* g_scales, const float* g_rotations, float* g_f_dc, float* g_f_rest,
                                    float* g_opacity_raw, float* g_scaling_raw, float* g_rotation_raw, void* stream);
B3GS_API int b3gs_adam_multi(int n_tensors, const B3gsAdamTensor* tensors, double beta1, double beta2, double eps,
                             void* stream);
B3GS_API int b3gs_opacity_decay(int P, float factor, float* opacity_raw, void* stream);
B3GS_API int b3gs_densify_stats(int P, const float* viewspace_grad, const int* radii, float* xyz_gradient_accum,
                                float* denom, float* max_radii2D, void* stream);

/*
 * ---- simple-knn distCUDA2 -------------------------------------
 * mean_dist2[i] = mean of the squared distances from point i to its 3 nearest neighbours
 * (submodules/simple-knn/simple_knn.cu:185-221 SimpleKNN::knn, bound as
 * simple_knn._C.distCUDA2 in spatial.cu:15-25; used at scene/gaussian_model.py:134).
 * points: float[P,3]; mean_dist2: float[P].  Exact search, distances evaluated with the
 * reference's expression, so results are bit-identical to the reference's.  With fewer
 * than 3 other points the missing neighbours count as FLT_MAX (the result overflows to
 * inf), as in the reference.  scratch: device memory of at least
 * b3gs_dist_cuda2_scratch_bytes(P) bytes; no allocation, no host synchronisation.
 */
B3GS_API size_t b3gs_dist_cuda2_scratch_bytes(int P);
B3GS_API int b3gs_dist_cuda2(int P, const float* points, float* mean_dist2, void* scratch, size_t scratch_bytes,
                             void* stream);

/*
 * ---- the data-parallel exchange over NVLink peer memory -------------
 * In-place two-shot all-reduce (sum, then * scale) of one float buffer per rank that lives in
 * symmetric memory: peer_buffers[r] is the device pointer of rank r's buffer as mapped into
 * THIS process (torch.distributed._symmetric_memory buffer_ptrs), all 16-byte aligned and
 * n_floats (a multiple of 4) long.  Rank `rank` reduces the rank-th slice and stores the
 * result into every peer's buffer; all ranks must call it, bracketed by a cross-rank
 * barrier on the same stream before (the peers' producers have finished) and after (the
 * results have landed).  Every element is summed by exactly one rank in rank order, so all
 * replicas end up bit-identical.
 */
#define B3GS_MAX_PEERS 8
B3GS_API int b3gs_peer_allreduce(int world, int rank, float* const* peer_buffers, size_t n_floats, float scale,
                                 void* stream);
/* The same exchange through the NVSwitch (NVLS): multicast_buffer is the multicast address of
 * the symmetric buffer (_SymmetricMemory.multicast_ptr); the sum is formed inside the switch by
 * multimem.ld_reduce and broadcast by multimem.st.  Same calling protocol (barrier before and
 * after); replicas bit-identical, summation order unspecified. */
B3GS_API int b3gs_peer_allreduce_multimem(int world, int rank, float* multicast_buffer, size_t n_floats,
                                          float scale, void* stream);

/* The same exchange with both cross-rank barriers INSIDE the kernel (flag words in the symmetric
 * buffer, system-scope release/acquire) and launched with programmatic stream serialization, so
 * it is resident while the backward's last kernel drains: ONE launch per step, no host-issued
 * barrier.  The buffers must extend 64 32-bit words beyond flag_off_floats (>= n_floats, a multiple
 * of 4), zeroed when created; epoch = 1, 2, 3, ... identical on every rank and incremented per
 * call.  multicast_buffer: NULL for plain peer loads/stores, else the multicast address (NVLS).
 * When the call has completed on `stream`, every rank's bucket holds the reduced values and no
 * peer reads it any more.  Bounded waits: a missing peer traps after ~4 s instead of hanging. */
B3GS_API int b3gs_peer_allreduce_fused(int world, int rank, float* const* peer_buffers, float* multicast_buffer,
                                       size_t n_floats, size_t flag_off_floats, unsigned epoch, float scale,
                                       void* stream);

/*
 * Backward FUSED with the exchange: compute step and collective overlapped tile by tile.
 * b3gs_exchange_create describes the symmetric bucket once (as b3gs_peer_allreduce_fused: peers'
 * pointers, optional multicast address, data length, flag offset) and owns a side stream.
 * b3gs_backward_exchange is b3gs_backward_flags whose five parameter-gradient outputs
 * (dL_dmean3D, dL_dsh, dL_dopacity, dL_dscale, dL_drot) MUST be 16-byte aligned segments of
 * that bucket: the per-Gaussian backward kernel (K8+K9) runs chunk by chunk (8 chunks from 256k
 * Gaussians) on `stream`, and as soon as a chunk is written its ranges of the five segments are
 * all-reduced (sum * scale) over NVLink by the fused kernel on the side stream while the next
 * chunk is being computed; `stream` joins the side stream before the call returns control of the
 * bucket.  With B3GS_BWD_ACCUMULATE (second view of a step) the exchanged values are the sums of
 * both views.  All ranks must make the same sequence of calls.  b3gs_exchange_epoch reads (or,
 * with set != 0, sets) the plan's barrier epoch — needed only when b3gs_peer_allreduce_fused is
 * also used on the same buffer: the two share the flag words and one epoch sequence.
 */
B3GS_API int b3gs_exchange_create(int world, int rank, float* const* peer_buffers, float* multicast_buffer,
                                  size_t n_floats, size_t flag_off_floats, void** handle_out);
B3GS_API void b3gs_exchange_destroy(void* handle);
B3GS_API unsigned b3gs_exchange_epoch(void* handle, int set, unsigned value);
B3GS_API int b3gs_backward_exchange(
    void* exchange, float scale, unsigned flags,
    int P, int D, int M, int R, const float* background, int width, int height, const float* means3D,
    const float* shs, const float* colors_precomp, const float* alphas, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
    const float* campos, float tan_fovx, float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer,
    char* image_buffer, const float* dL_dpix, const float* dL_dpix_depth, const float* dL_dalphas, float* dL_dmean2D,
    float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_ddepth, float* dL_dmean3D, float* dL_dcov3D,
    float* dL_dsh, float* dL_dscale, float* dL_drot, int debug, void* stream);

/* The composite backward exists in three shapes — 1, 2 or 4 pixels per lane (8x4, 8x8, 16x8
 * pixels per warp) — with identical results up to float summation order; n = 0 (default)
 * picks per call from the instances-per-Gaussian ratio, n = 1|2|4 forces one (tests, A/B
 * timing; also the environment variable B3GS_BWD_PIX at load time). */
B3GS_API void b3gs_set_backward_pixels(int n);

/* Last error message of the calling thread ("" if none). */
B3GS_API const char* b3gs_last_error(void);

/* Library version string, and the number of kernels this library launched since
 * load (used by bench.py to report gpu_launches). */
B3GS_API const char* b3gs_version(void);
B3GS_API unsigned long long b3gs_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B3GS_H_INCLUDED */
