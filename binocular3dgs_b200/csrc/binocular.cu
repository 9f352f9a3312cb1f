// binocular.cu — the binocular-consistency loss, forward and
// backward, plus its two constituents as stand-alone operators.
//
// What the reference executes every iteration after shift_cam_start (train.py:122-136):
//     disparity = focal_x * (-trans_dist) / (depth + 1e-5)
//     warped    = inverse_warp_images(shifted_image, disparity)   utils/graphics_utils.py:80-125
//     mask      = inverse_warp_images(ones,          disparity)
//     loss      = l1_loss(warped, gt, mask)                       utils/loss_utils.py:18-21
//               + 0.05 * SmoothLoss(disparity * mask, gt)         utils/loss_utils.py:68-91
// as a Python double loop over batch x channel with advanced-index gathers, a
// device->host->device round trip (`.type(torch.LongTensor).cuda()`), four 3x3 conv2d and
// ~100 small elementwise kernels, then the autograd duals of all of them.
//
// Here: ONE forward kernel (the three partial sums) and ONE backward kernel (dL/dshifted
// by 6 RED.F32 per pixel, merged across neighbouring lanes when taps coincide, and
// dL/ddepth, which is K7's dL_dpix_depth input).  Nothing is saved between them: the
// backward recomputes disparity, taps and the edge weights from the same three inputs
// (28 B/pixel read) instead of reading saved maps.  HBM-bound streaming kernels.
//
// Semantics kept from the reference:
//   * x0 = floor(disparity) carries no gradient; taps c0 = col + x0, c1 = c0 + 1;
//     a pixel is invalid (output 0, no gradient) unless 0 <= c0 and c1 <= W-1;
//   * weights w0 = float(x0 + 1) - disparity, w1 = disparity - float(x0);
//   * the warped mask is w0 + w1 (1 up to rounding) on valid pixels; its derivative
//     w.r.t. the disparity is (-1) + (+1) = 0;
//   * SmoothLoss: central differences 0.5*(f[+1] - f[-1]) on the interior (no padding),
//     image differences summed over the 3 channels, weights exp(-0.33 |.|), two means
//     over (H-2)(W-2); d|z|/dz = sign(z) with sign(0) = 0.
#include "../../include/b3gs.h"
#include "common.cuh"
#include "kernels.h"

namespace b3 {

constexpr int kBX = 32, kBY = 8, kThreads = kBX * kBY;
constexpr float kDepthEps = 1e-5f, kEdgeK = -0.33f;

struct Taps {
    int c0;      // left tap column (valid pixels only)
    float w0, w1, mask;
    bool valid;
};

__device__ __forceinline__ Taps make_taps(float disp, int x, int W) {
    Taps t;
    const float x0f = floorf(disp);
    const float c0f = (float)x + x0f;  // exact whenever it can be in range
    t.valid = (c0f >= 0.f) && (c0f <= (float)(W - 2));  // NaN -> invalid
    t.c0 = t.valid ? (int)c0f : 0;
    t.w0 = __fsub_rn(__fadd_rn(x0f, 1.f), disp);
    t.w1 = __fsub_rn(disp, x0f);
    t.mask = t.valid ? __fadd_rn(t.w0, t.w1) : 0.f;
    return t;
}

// `k / (depth + 1e-5)` with a python scalar on the left is Tensor.__rtruediv__, which
// torch evaluates as reciprocal(depth + 1e-5) * k: two roundings, kept.
__device__ __forceinline__ float disparity_of(float depth, float k_disp) {
    return __fmul_rn(__frcp_rn(__fadd_rn(depth, kDepthEps)), k_disp);
}

// masked disparity (disparity * warped mask) of pixel (x, y); 0 outside the image
__device__ __forceinline__ float masked_disparity(const float* __restrict__ depth, float k_disp, int x, int y, int W,
                                                  int H) {
    if (x < 0 || y < 0 || x >= W || y >= H) return 0.f;
    const float d = disparity_of(depth[(size_t)y * W + x], k_disp);
    return __fmul_rn(d, make_taps(d, x, W).mask);
}

__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

// edge weights of the 3-channel image at interior centre (x, y)
__device__ __forceinline__ void edge_weights(const float* __restrict__ img, size_t plane, int x, int y, int W,
                                             float& wx, float& wy) {
    const size_t o = (size_t)y * W + x;
    float ex = 0.f, ey = 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        const float* p = img + ch * plane + o;
        ex += 0.5f * (p[1] - p[-1]);
        ey += 0.5f * (p[W] - p[-W]);
    }
    wx = expf(kEdgeK * fabsf(ex));
    wy = expf(kEdgeK * fabsf(ey));
}

__device__ __forceinline__ float block_sum(float v, float* s_red, int tid) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((tid & 31) == 0) s_red[tid >> 5] = v;
    __syncthreads();
    float t = (tid < kThreads / 32) ? s_red[tid] : 0.f;
    if (tid < 32) {
#pragma unroll
        for (int d = kThreads / 64; d >= 1; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
    }
    __syncthreads();
    return t;  // valid in thread 0
}

// ------------------------------------------------------------------ smoothness pieces
// Stage the masked disparity of the block's pixels plus a halo into shared memory.
// FROM_DEPTH: compute it from the depth map (fused loss); else read it (SmoothLoss op).
template <int HALO, bool FROM_DEPTH>
__device__ __forceinline__ void stage_disparity(float (*s)[kBX + 2 * HALO], const float* __restrict__ src, float k_disp,
                                                int bx0, int by0, int W, int H, int tid) {
    constexpr int SW = kBX + 2 * HALO, SH = kBY + 2 * HALO;
    for (int i = tid; i < SW * SH; i += kThreads) {
        const int ly = i / SW, lx = i - ly * SW;
        const int x = bx0 + lx - HALO, y = by0 + ly - HALO;
        float v;
        if (FROM_DEPTH) v = masked_disparity(src, k_disp, x, y, W, H);
        else v = (x >= 0 && y >= 0 && x < W && y < H) ? src[(size_t)y * W + x] : 0.f;
        s[ly][lx] = v;
    }
}

// forward: |wx * dx| and |wy * dy| of this thread's pixel if it is an interior centre
template <bool FROM_DEPTH>
__device__ __forceinline__ void smooth_forward_terms(const float* __restrict__ src, float k_disp,
                                                     const float* __restrict__ img, int W, int H, float& tx,
                                                     float& ty) {
    __shared__ float s_d[kBY + 2][kBX + 2];
    const int tid = threadIdx.y * kBX + threadIdx.x;
    const int bx0 = blockIdx.x * kBX, by0 = blockIdx.y * kBY;
    stage_disparity<1, FROM_DEPTH>(s_d, src, k_disp, bx0, by0, W, H, tid);
    __syncthreads();
    const int x = bx0 + threadIdx.x, y = by0 + threadIdx.y;
    tx = ty = 0.f;
    if (x >= 1 && y >= 1 && x < W - 1 && y < H - 1) {
        float wx, wy;
        edge_weights(img, (size_t)H * W, x, y, W, wx, wy);
        const int lx = threadIdx.x + 1, ly = threadIdx.y + 1;
        tx = fabsf(wx * (0.5f * (s_d[ly][lx + 1] - s_d[ly][lx - 1])));
        ty = fabsf(wy * (0.5f * (s_d[ly + 1][lx] - s_d[ly - 1][lx])));
    }
}

// backward: d(sum_x + sum_y)/d(masked disparity of this thread's pixel), unscaled
template <bool FROM_DEPTH>
__device__ __forceinline__ float smooth_backward_term(const float* __restrict__ src, float k_disp,
                                                      const float* __restrict__ img, int W, int H) {
    __shared__ float s_d[kBY + 4][kBX + 4];
    __shared__ float s_sx[kBY + 2][kBX + 2], s_sy[kBY + 2][kBX + 2];
    const int tid = threadIdx.y * kBX + threadIdx.x;
    const int bx0 = blockIdx.x * kBX, by0 = blockIdx.y * kBY;
    stage_disparity<2, FROM_DEPTH>(s_d, src, k_disp, bx0, by0, W, H, tid);
    __syncthreads();
    // per-centre signed weights on the block's pixels plus a halo of 1
    for (int i = tid; i < (kBX + 2) * (kBY + 2); i += kThreads) {
        const int ly = i / (kBX + 2), lx = i - ly * (kBX + 2);
        const int x = bx0 + lx - 1, y = by0 + ly - 1;
        float sx = 0.f, sy = 0.f;
        if (x >= 1 && y >= 1 && x < W - 1 && y < H - 1) {
            float wx, wy;
            edge_weights(img, (size_t)H * W, x, y, W, wx, wy);
            const int dx_ = lx + 1, dy_ = ly + 1;  // same pixel in s_d coordinates
            sx = sgn(wx * (0.5f * (s_d[dy_][dx_ + 1] - s_d[dy_][dx_ - 1]))) * wx;
            sy = sgn(wy * (0.5f * (s_d[dy_ + 1][dx_] - s_d[dy_ - 1][dx_]))) * wy;
        }
        s_sx[ly][lx] = sx;
        s_sy[ly][lx] = sy;
    }
    __syncthreads();
    const int lx = threadIdx.x + 1, ly = threadIdx.y + 1;
    // pixel (x,y) is the +1 neighbour of centre (x-1,y) and the -1 neighbour of (x+1,y)
    return 0.5f * ((s_sx[ly][lx - 1] - s_sx[ly][lx + 1]) + (s_sy[ly - 1][lx] - s_sy[ly + 1][lx]));
}

// ------------------------------------------------------------------ fused loss kernels
__global__ void __launch_bounds__(kThreads) binocular_forward_kernel(int H, int W, const float* __restrict__ shifted,
                                                                    const float* __restrict__ depth,
                                                                    const float* __restrict__ gt, float k_disp,
                                                                    double* __restrict__ sums, float k_l1, float k_sm,
                                                                    float* __restrict__ loss_out) {
    __shared__ float s_red[kThreads / 32];
    const int tid = threadIdx.y * kBX + threadIdx.x;
    float tx, ty;
    smooth_forward_terms<true>(depth, k_disp, gt, W, H, tx, ty);
    const int x = blockIdx.x * kBX + threadIdx.x, y = blockIdx.y * kBY + threadIdx.y;
    float l1 = 0.f;
    if (x < W && y < H) {
        const size_t plane = (size_t)H * W, row = (size_t)y * W;
        const Taps t = make_taps(disparity_of(depth[row + x], k_disp), x, W);
        if (t.valid) {
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                const float* p = shifted + ch * plane + row + t.c0;
                const float warped = __fadd_rn(__fmul_rn(t.w0, p[0]), __fmul_rn(t.w1, p[1]));
                l1 += fabsf(__fsub_rn(__fmul_rn(warped, t.mask), __fmul_rn(gt[ch * plane + row + x], t.mask)));
            }
        }
    }
    const float b0 = block_sum(l1, s_red, tid);
    const float b1 = block_sum(tx, s_red, tid);
    const float b2 = block_sum(ty, s_red, tid);
    if (tid == 0) {
        atomicAdd(sums + 0, (double)b0);
        atomicAdd(sums + 1, (double)b1);
        atomicAdd(sums + 2, (double)b2);
        if (loss_out) {  // the last block to finish forms the loss value
            __threadfence();
            const unsigned long long ticket = atomicAdd(reinterpret_cast<unsigned long long*>(sums + 3), 1ull);
            if (ticket == (unsigned long long)gridDim.x * gridDim.y - 1) {
                __threadfence();
                const volatile double* v = sums;
                *loss_out = (float)((double)k_l1 * v[0] + (double)k_sm * (v[1] + v[2]));
            }
        }
    }
}

// Adds v at column c of `row`; lanes whose neighbour (lane+1) writes the same address fold
// their value into the neighbour first (the right tap of pixel x is the left tap of pixel
// x+1 whenever floor(disparity) is locally constant), halving the RED traffic.
__device__ __forceinline__ void red_pair(float* __restrict__ row, bool valid, int c0, float v0, float v1) {
    // lane L's right tap (c0+1) vs lane L+1's left tap
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int n_c0 = __shfl_down_sync(full, c0, 1);
    const bool n_valid = __shfl_down_sync(full, (int)valid, 1) != 0;
    const bool give = valid && lane < 31 && n_valid && n_c0 == c0 + 1;
    const float from_left = __shfl_up_sync(full, give ? v1 : 0.f, 1);
    if (valid) {
        atomicAdd(row + c0, v0 + (lane > 0 ? from_left : 0.f));
        if (!give) atomicAdd(row + c0 + 1, v1);
    }
}

__global__ void __launch_bounds__(kThreads) binocular_backward_kernel(int H, int W, const float* __restrict__ shifted,
                                                                     const float* __restrict__ depth,
                                                                     const float* __restrict__ gt, float k_disp,
                                                                     const float* __restrict__ upstream, float k_l1,
                                                                     float k_sm,
                                                                     float* __restrict__ dL_dshifted,
                                                                     float* __restrict__ dL_ddepth) {
    const float g_dm = smooth_backward_term<true>(depth, k_disp, gt, W, H);
    const int x = blockIdx.x * kBX + threadIdx.x, y = blockIdx.y * kBY + threadIdx.y;
    const bool in = x < W && y < H;
    const float g_l1 = upstream[0] * k_l1, g_sm = upstream[0] * k_sm;
    const size_t plane = (size_t)H * W, row = (size_t)(in ? y : 0) * W;
    float dep = 0.f, disp = 0.f, g_disp = 0.f;
    Taps t;
    t.valid = false; t.c0 = 0; t.w0 = t.w1 = t.mask = 0.f;
    if (in) {
        dep = depth[row + x];
        disp = disparity_of(dep, k_disp);
        t = make_taps(disp, x, W);
    }
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        float gw = 0.f;
        if (t.valid) {
            const float* p = shifted + ch * plane + row + t.c0;
            const float a = p[0], b = p[1];
            const float warped = __fadd_rn(__fmul_rn(t.w0, a), __fmul_rn(t.w1, b));
            const float diff = __fsub_rn(__fmul_rn(warped, t.mask), __fmul_rn(gt[ch * plane + row + x], t.mask));
            gw = g_l1 * sgn(diff) * t.mask;
            g_disp = fmaf(gw, b - a, g_disp);
        }
        red_pair(dL_dshifted + ch * plane + row, t.valid, t.c0, gw * t.w0, gw * t.w1);
    }
    if (in) {
        g_disp = fmaf(g_sm * g_dm, t.mask, g_disp);
        // d disparity / d depth = -k / (depth + eps)^2 = -disparity / (depth + eps)
        dL_ddepth[row + x] = g_disp * -(disp * __frcp_rn(__fadd_rn(dep, kDepthEps)));
    }
}

// ------------------------------------------------------------------ stand-alone operators
__global__ void __launch_bounds__(256) warp_forward_kernel(int C, int H, int W, const float* __restrict__ image,
                                                          const float* __restrict__ disparity,
                                                          float* __restrict__ warped) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t plane = (size_t)H * W, row = (size_t)y * W;
    const Taps t = make_taps(disparity[row + x], x, W);
    for (int ch = 0; ch < C; ch++) {
        float v = 0.f;
        if (t.valid) {
            const float* p = image + ch * plane + row + t.c0;
            v = __fadd_rn(__fmul_rn(t.w0, p[0]), __fmul_rn(t.w1, p[1]));
        }
        warped[ch * plane + row + x] = v;
    }
}

__global__ void __launch_bounds__(256) warp_backward_kernel(int C, int H, int W, const float* __restrict__ image,
                                                           const float* __restrict__ disparity,
                                                           const float* __restrict__ dL_dwarped,
                                                           float* __restrict__ dL_dimage,
                                                           float* __restrict__ dL_ddisparity) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const bool in = x < W;
    const size_t plane = (size_t)H * W, row = (size_t)y * W;
    Taps t;
    t.valid = false; t.c0 = 0; t.w0 = t.w1 = t.mask = 0.f;
    if (in) t = make_taps(disparity[row + x], x, W);
    float g_disp = 0.f;
    for (int ch = 0; ch < C; ch++) {
        float gw = 0.f;
        if (t.valid) {
            gw = dL_dwarped[ch * plane + row + x];
            const float* p = image + ch * plane + row + t.c0;
            g_disp = fmaf(gw, p[1] - p[0], g_disp);
        }
        red_pair(dL_dimage + ch * plane + row, t.valid, t.c0, gw * t.w0, gw * t.w1);
    }
    if (in && dL_ddisparity) dL_ddisparity[row + x] = g_disp;
}

__global__ void __launch_bounds__(kThreads) smooth_forward_kernel(int H, int W, const float* __restrict__ disparity,
                                                                 const float* __restrict__ image,
                                                                 double* __restrict__ sums) {
    __shared__ float s_red[kThreads / 32];
    const int tid = threadIdx.y * kBX + threadIdx.x;
    float tx, ty;
    smooth_forward_terms<false>(disparity, 0.f, image, W, H, tx, ty);
    const float b1 = block_sum(tx, s_red, tid);
    const float b2 = block_sum(ty, s_red, tid);
    if (tid == 0) {
        atomicAdd(sums + 0, (double)b1);
        atomicAdd(sums + 1, (double)b2);
    }
}

__global__ void __launch_bounds__(kThreads) smooth_backward_kernel(int H, int W, const float* __restrict__ disparity,
                                                                  const float* __restrict__ image,
                                                                  const float* __restrict__ upstream, float k,
                                                                  float* __restrict__ dL_ddisparity) {
    const float g = smooth_backward_term<false>(disparity, 0.f, image, W, H);
    const int x = blockIdx.x * kBX + threadIdx.x, y = blockIdx.y * kBY + threadIdx.y;
    if (x < W && y < H) dL_ddisparity[(size_t)y * W + x] = upstream[0] * k * g;
}

static dim3 tile_grid(int W, int H) { return dim3((W + kBX - 1) / kBX, (H + kBY - 1) / kBY); }

}  // namespace b3

using namespace b3;

extern "C" {

int b3gs_binocular_forward(int H, int W, const float* shifted, const float* depth, const float* gt, float k_disp,
                           double* sums, float k_l1, float k_sm, float* loss_out, void* stream) {
    if (H < 3 || W < 3 || !shifted || !depth || !gt || !sums) return -1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(sums, 0, 4 * sizeof(double), st) != cudaSuccess) return -2;
    binocular_forward_kernel<<<tile_grid(W, H), dim3(kBX, kBY), 0, st>>>(H, W, shifted, depth, gt, k_disp, sums, k_l1,
                                                                        k_sm, loss_out);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int b3gs_binocular_backward(int H, int W, const float* shifted, const float* depth, const float* gt, float k_disp,
                            const float* upstream, float k_l1, float k_sm, float* dL_dshifted, float* dL_ddepth,
                            void* stream) {
    if (H < 3 || W < 3 || !shifted || !depth || !gt || !upstream || !dL_dshifted || !dL_ddepth) return -1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(dL_dshifted, 0, 3 * sizeof(float) * (size_t)H * W, st) != cudaSuccess) return -2;
    binocular_backward_kernel<<<tile_grid(W, H), dim3(kBX, kBY), 0, st>>>(H, W, shifted, depth, gt, k_disp, upstream,
                                                                         k_l1, k_sm,
                                                                         dL_dshifted, dL_ddepth);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int b3gs_warp_forward(int C, int H, int W, const float* image, const float* disparity, float* warped, void* stream) {
    if (C <= 0 || H <= 0 || W <= 0 || !image || !disparity || !warped) return -1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    warp_forward_kernel<<<dim3((W + 255) / 256, H), 256, 0, st>>>(C, H, W, image, disparity, warped);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int b3gs_warp_backward(int C, int H, int W, const float* image, const float* disparity, const float* dL_dwarped,
                       float* dL_dimage, float* dL_ddisparity, void* stream) {
    if (C <= 0 || H <= 0 || W <= 0 || !image || !disparity || !dL_dwarped || !dL_dimage) return -1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(dL_dimage, 0, sizeof(float) * (size_t)C * H * W, st) != cudaSuccess) return -2;
    warp_backward_kernel<<<dim3((W + 255) / 256, H), 256, 0, st>>>(C, H, W, image, disparity, dL_dwarped, dL_dimage,
                                                                  dL_ddisparity);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int b3gs_smooth_forward(int H, int W, const float* disparity, const float* image, double* sums, void* stream) {
    if (H < 3 || W < 3 || !disparity || !image || !sums) return -1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(sums, 0, 2 * sizeof(double), st) != cudaSuccess) return -2;
    smooth_forward_kernel<<<tile_grid(W, H), dim3(kBX, kBY), 0, st>>>(H, W, disparity, image, sums);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int b3gs_smooth_backward(int H, int W, const float* disparity, const float* image, const float* upstream, float k,
                         float* dL_ddisparity, void* stream) {
    if (H < 3 || W < 3 || !disparity || !image || !upstream || !dL_ddisparity) return -1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    smooth_backward_kernel<<<tile_grid(W, H), dim3(kBX, kBY), 0, st>>>(H, W, disparity, image, upstream, k,
                                                                      dL_ddisparity);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // extern "C"
