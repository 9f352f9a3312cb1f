// photometric.cu — fused L1 + SSIM photometric loss, forward and backward.
//
// SURVEY.md §8(f) rank 1, the first "next" row after the rasterizer: the loss the
// reference evaluates on the composite's output every iteration,
//     loss = (1 - lambda) * l1_loss(image, gt) + lambda * (1 - ssim(image, gt))
// (train.py:146-147; utils/loss_utils.py:18-21 l1_loss, :36-66 ssim/_ssim, :26-34 window).
// The reference runs five depthwise 11x11 conv2d calls plus ~15 elementwise kernels in
// the forward and their autograd duals in the backward.  Here:
//
//   photometric_forward_kernel   one pass: per 16x16 tile and channel, loads the 26x26
//       halo of both images into shared memory (zero padding == conv2d padding=5),
//       separable 11-tap Gaussian (horizontal into shared memory, vertical in
//       registers) of x, y, x^2, y^2, xy; the SSIM map; the three partial derivatives
//       dS/dmu1, dS/dsigma1^2, dS/dsigma12 saved for the backward; block-reduced sums of
//       the SSIM map and of |x - y|.
//   photometric_backward_kernel  one pass: separable convolution of the three saved
//       derivative maps (the window is symmetric, so correlation == convolution) and
//       dL/dx = gS * [conv(dS/dmu1) + 2x conv(dS/dsigma1^2) + y conv(dS/dsigma12)]
//               + gL1 * sign(x - y).
//
// HBM-bound by design: forward reads 2 images and writes 3 maps, backward reads 5 and
// writes 1 — 24 B and 24 B per pixel-channel.  Gradient flows to the first image only
// (the render); the ground truth is a constant in the reference's training loop.
#include "../../include/b3gs.h"
#include "common.cuh"
#include "kernels.h"

namespace b3 {

constexpr int kWin = 11, kHalo = 5, kTile = 16, kExt = kTile + 2 * kHalo;  // 26

// gaussian(11, 1.5) normalised, as float32 exactly as utils/loss_utils.py:26-28 builds it
__constant__ float kGauss[kWin] = {0.001028380123898387f,  0.0075987582094967365f, 0.036000773310661316f,
                                   0.10936068743467331f,   0.21300552785396576f,   0.26601171493530273f,
                                   0.21300552785396576f,   0.10936068743467331f,   0.036000773310661316f,
                                   0.0075987582094967365f, 0.001028380123898387f};

// tid: linear thread index of the 16x16 block
__device__ __forceinline__ float block_sum_256(float v, float* s_red, int tid) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    const int lane = tid & 31, warp = tid >> 5;
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = (tid < 8) ? s_red[tid] : 0.f;
    if (warp == 0) {
#pragma unroll
        for (int d = 4; d >= 1; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
    }
    __syncthreads();
    return t;  // valid in thread 0
}

__global__ void __launch_bounds__(256) photometric_forward_kernel(int H, int W, const float* __restrict__ img1,
                                                                 const float* __restrict__ img2,
                                                                 float* __restrict__ dm_dmu1,
                                                                 float* __restrict__ dm_dsigma1,
                                                                 float* __restrict__ dm_dsigma12,
                                                                 float* __restrict__ ssim_map /* may be null */,
                                                                 double* __restrict__ sums /* [0] ssim, [1] l1 */) {
    __shared__ float s1[kExt][kExt + 1], s2[kExt][kExt + 1];
    __shared__ float h[5][kExt][kTile + 1];
    __shared__ float s_red[8];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
    const size_t plane = (size_t)H * W;
    const float* p1 = img1 + c * plane;
    const float* p2 = img2 + c * plane;
    const int tid = threadIdx.y * kTile + threadIdx.x;
    for (int i = tid; i < kExt * kExt; i += 256) {
        const int ly = i / kExt, lx = i - ly * kExt;
        const int gy = y0 + ly - kHalo, gx = x0 + lx - kHalo;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        s1[ly][lx] = in ? p1[(size_t)gy * W + gx] : 0.f;
        s2[ly][lx] = in ? p2[(size_t)gy * W + gx] : 0.f;
    }
    __syncthreads();
    // horizontal pass: 26 rows x 16 columns
    for (int i = tid; i < kExt * kTile; i += 256) {
        const int ly = i / kTile, lx = i - ly * kTile;
        float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; k++) {
            const float w = kGauss[k], u = s1[ly][lx + k], v = s2[ly][lx + k];
            a = fmaf(w, u, a); b = fmaf(w, v, b);
            aa = fmaf(w, u * u, aa); bb = fmaf(w, v * v, bb); ab = fmaf(w, u * v, ab);
        }
        h[0][ly][lx] = a; h[1][ly][lx] = b; h[2][ly][lx] = aa; h[3][ly][lx] = bb; h[4][ly][lx] = ab;
    }
    __syncthreads();
    const int lx = threadIdx.x, ly = threadIdx.y;
    const int gx = x0 + lx, gy = y0 + ly;
    float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; k++) {
        const float w = kGauss[k];
        mu1 = fmaf(w, h[0][ly + k][lx], mu1);
        mu2 = fmaf(w, h[1][ly + k][lx], mu2);
        e11 = fmaf(w, h[2][ly + k][lx], e11);
        e22 = fmaf(w, h[3][ly + k][lx], e22);
        e12 = fmaf(w, h[4][ly + k][lx], e12);
    }
    float ssim_v = 0.f, l1_v = 0.f;
    if (gx < W && gy < H) {
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu1_mu2 = mu1 * mu2;
        const float sigma1_sq = e11 - mu1_sq, sigma2_sq = e22 - mu2_sq, sigma12 = e12 - mu1_mu2;
        const float A = 2.f * mu1_mu2 + C1, B = 2.f * sigma12 + C2;
        const float Cc = mu1_sq + mu2_sq + C1, Dd = sigma1_sq + sigma2_sq + C2;
        const float inv = 1.f / (Cc * Dd);
        ssim_v = A * B * inv;
        const size_t o = c * plane + (size_t)gy * W + gx;
        // dS/dmu1 (through A, B = f(sigma12), Cc, Dd = f(sigma1_sq)); dS/d(E[x^2]); dS/d(E[xy])
        dm_dmu1[o] = 2.f * mu2 * (B - A) * inv - ssim_v * 2.f * mu1 * (1.f / Cc - 1.f / Dd);
        dm_dsigma1[o] = -ssim_v / Dd;
        dm_dsigma12[o] = 2.f * A * inv;
        if (ssim_map) ssim_map[o] = ssim_v;
        l1_v = fabsf(s1[ly + kHalo][lx + kHalo] - s2[ly + kHalo][lx + kHalo]);
    }
    const float bs = block_sum_256(ssim_v, s_red, tid);
    const float bl = block_sum_256(l1_v, s_red, tid);
    if (tid == 0) {
        atomicAdd(sums + 0, (double)bs);
        atomicAdd(sums + 1, (double)bl);
    }
}

__global__ void __launch_bounds__(256) photometric_backward_kernel(int H, int W, const float* __restrict__ img1,
                                                                  const float* __restrict__ img2,
                                                                  const float* __restrict__ dm_dmu1,
                                                                  const float* __restrict__ dm_dsigma1,
                                                                  const float* __restrict__ dm_dsigma12,
                                                                  const float* __restrict__ upstream /* device float[1] */, float k_ssim,
                                                                  float k_l1,
                                                                  float* __restrict__ dL_dimg1) {
    __shared__ float s[3][kExt][kExt + 1];
    __shared__ float h[3][kExt][kTile + 1];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
    const size_t plane = (size_t)H * W;
    const float* m0 = dm_dmu1 + c * plane;
    const float* m1 = dm_dsigma1 + c * plane;
    const float* m2 = dm_dsigma12 + c * plane;
    const int tid = threadIdx.y * kTile + threadIdx.x;
    for (int i = tid; i < kExt * kExt; i += 256) {
        const int ly = i / kExt, lx = i - ly * kExt;
        const int gy = y0 + ly - kHalo, gx = x0 + lx - kHalo;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        const size_t o = (size_t)gy * W + gx;
        s[0][ly][lx] = in ? m0[o] : 0.f;
        s[1][ly][lx] = in ? m1[o] : 0.f;
        s[2][ly][lx] = in ? m2[o] : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < kExt * kTile; i += 256) {
        const int ly = i / kTile, lx = i - ly * kTile;
        float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; k++) {
            const float w = kGauss[k];
            a = fmaf(w, s[0][ly][lx + k], a);
            b = fmaf(w, s[1][ly][lx + k], b);
            d = fmaf(w, s[2][ly][lx + k], d);
        }
        h[0][ly][lx] = a; h[1][ly][lx] = b; h[2][ly][lx] = d;
    }
    __syncthreads();
    const int lx = threadIdx.x, ly = threadIdx.y;
    const int gx = x0 + lx, gy = y0 + ly;
    if (gx >= W || gy >= H) return;
    float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; k++) {
        const float w = kGauss[k];
        a = fmaf(w, h[0][ly + k][lx], a);
        b = fmaf(w, h[1][ly + k][lx], b);
        d = fmaf(w, h[2][ly + k][lx], d);
    }
    const size_t o = c * plane + (size_t)gy * W + gx;
    const float x = img1[o], y = img2[o];
    const float g = upstream[0], gS = g * k_ssim, gL1 = g * k_l1;
    const float diff = x - y;
    const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);  // torch.abs backward: sign(0) = 0
    dL_dimg1[o] = gS * (a + 2.f * x * b + y * d) + gL1 * sgn;
}

}  // namespace b3

using namespace b3;

extern "C" {

int b3gs_photometric_forward(int C, int H, int W, const float* img1, const float* img2, float* dm_dmu1,
                             float* dm_dsigma1_sq, float* dm_dsigma12, float* ssim_map, double* sums, void* stream) {
    if (C <= 0 || H <= 0 || W <= 0 || !img1 || !img2 || !dm_dmu1 || !dm_dsigma1_sq || !dm_dsigma12 || !sums) return -1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(sums, 0, 2 * sizeof(double), st) != cudaSuccess) return -2;
    dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, C), block(kTile, kTile);
    photometric_forward_kernel<<<grid, block, 0, st>>>(H, W, img1, img2, dm_dmu1, dm_dsigma1_sq, dm_dsigma12, ssim_map,
                                                      sums);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int b3gs_photometric_backward(int C, int H, int W, const float* img1, const float* img2, const float* dm_dmu1,
                              const float* dm_dsigma1_sq, const float* dm_dsigma12, const float* upstream,
                              float k_ssim, float k_l1,
                              float* dL_dimg1, void* stream) {
    if (C <= 0 || H <= 0 || W <= 0 || !img1 || !img2 || !dm_dmu1 || !dm_dsigma1_sq || !dm_dsigma12 || !upstream ||
        !dL_dimg1)
        return -1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, C), block(kTile, kTile);
    photometric_backward_kernel<<<grid, block, 0, st>>>(H, W, img1, img2, dm_dmu1, dm_dsigma1_sq, dm_dsigma12, upstream,
                                                       k_ssim, k_l1,
                                                       dL_dimg1);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // extern "C"
