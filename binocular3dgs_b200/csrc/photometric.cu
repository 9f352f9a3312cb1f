// photometric.cu — fused L1 + SSIM photometric loss, forward and backward.
//
// SURVEY.md §8(f) rank 1, the first "next" row after the rasterizer: the loss the
// reference evaluates on the composite's output every iteration,
//     loss = (1 - lambda) * l1_loss(image, gt) + lambda * (1 - ssim(image, gt))
// (train.py:146-147; utils/loss_utils.py:18-21 l1_loss, :36-66 ssim/_ssim, :26-34 window).
// The reference runs five depthwise 11x11 conv2d calls plus ~15 elementwise kernels in
// the forward and their autograd duals in the backward.  Here:
//
//   photometric_forward_kernel   one pass: per 32x32 tile and channel, loads the 42x42
//       halo of both images into shared memory (zero padding == conv2d padding=5),
//       separable 11-tap Gaussian of x, y, x^2, y^2, xy, both directions register-blocked
//       (8 outputs per thread from 18 inputs); the SSIM map; the three partial derivatives
//       dS/dmu1, dS/dsigma1^2, dS/dsigma12 saved for the backward; block-reduced sums of
//       the SSIM map and of |x - y|; the last block forms the loss value.
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

// Tile 32x32 outputs per block and channel, halo 5 -> 42x42 inputs.  Both passes of the
// separable 11-tap filter are register-blocked: a thread produces 8 consecutive outputs
// along the filter direction from 18 inputs held in registers, so each shared-memory
// value is loaded once per 8 outputs instead of once per tap (the first revision issued one
// LDS per FMA and was shared-memory bound at ~7x the HBM time).
constexpr int kWin = 11, kHalo = 5, kTile = 32, kExt = kTile + 2 * kHalo;  // 42
constexpr int kSeg = 8, kSegIn = kSeg + kWin - 1;                          // 8 outputs from 18 inputs
constexpr int kSegs = kTile / kSeg;                                        // 4
constexpr int kThreads = 256;

// gaussian(11, 1.5) normalised, as float32 exactly as utils/loss_utils.py:26-28 builds it
__constant__ float kGauss[kWin] = {0.001028380123898387f,  0.0075987582094967365f, 0.036000773310661316f,
                                   0.10936068743467331f,   0.21300552785396576f,   0.26601171493530273f,
                                   0.21300552785396576f,   0.10936068743467331f,   0.036000773310661316f,
                                   0.0075987582094967365f, 0.001028380123898387f};

// out[i] = sum_k w[k] * in[i + k], i < SEG
template <int SEG>
__device__ __forceinline__ void conv_seg(const float (&in)[SEG + kWin - 1], float (&out)[SEG]) {
#pragma unroll
    for (int i = 0; i < SEG; i++) out[i] = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; k++) {
        const float w = kGauss[k];
#pragma unroll
        for (int i = 0; i < SEG; i++) out[i] = fmaf(w, in[i + k], out[i]);
    }
}
__device__ __forceinline__ void conv8(const float (&in)[kSegIn], float (&out)[kSeg]) { conv_seg<kSeg>(in, out); }

// The vertical pass uses segments of 4 outputs (14 inputs) so that 32 columns x 8 segments
// occupy all 256 threads (8-output segments would leave half the block idle in that phase).
constexpr int kVSeg = 4, kVSegIn = kVSeg + kWin - 1, kVSegs = kTile / kVSeg;
static_assert(kTile * kVSegs == kThreads, "vertical pass: one work item per thread");

// Stage the 42x42 halo of one image plane into shared memory (zero padding == conv2d
// padding=5).  When rows are 16-byte aligned (W % 4 == 0 and an aligned base) the halo is
// read as 42 rows x 12 float4 covering columns [x0-8, x0+40): 2 vector loads per thread
// instead of 7 scalar ones with their index arithmetic.  Column c of the tile's halo sits at
// s[row][c + kPadL] in both cases.
constexpr int kPadL = 3;                         // x0-8 .. x0-6 are loaded but unused
constexpr int kRowW = kExt + kPadL + 3 + 1;      // 42 + 3 left + 3 right (48 loaded) + 1 pad = 49

__device__ __forceinline__ void stage_plane(float (*s)[kRowW], const float* __restrict__ p, int x0, int y0, int H, int W,
                                            bool vec, int tid) {
    if (vec) {
        for (int i = tid; i < kExt * 12; i += kThreads) {
            const int ly = i / 12, q = i - ly * 12;
            const int gy = y0 + ly - kHalo, gx = x0 - 8 + 4 * q;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = *reinterpret_cast<const float4*>(p + (size_t)gy * W + gx);
            float* d = &s[ly][4 * q];
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
    } else {
        for (int i = tid; i < kExt * kExt; i += kThreads) {
            const int ly = i / kExt, lx = i - ly * kExt;
            const int gy = y0 + ly - kHalo, gx = x0 + lx - kHalo;
            const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
            s[ly][lx + kPadL] = in ? p[(size_t)gy * W + gx] : 0.f;
        }
    }
}

__device__ __forceinline__ bool rows_aligned(const float* p, int W, size_t plane) {
    return ((W & 3) == 0) && ((plane & 3) == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// Adds the block's partial sums and, in the last block to finish, turns the totals into the
// loss value: loss = k[0] + k[1] * sums[0] + k[2] * sums[1].  sums[2] is the block ticket.
__device__ __forceinline__ void finish_sums(double* sums, float s0, float s1, unsigned n_blocks, float k0, float k1,
                                            float k2, float* loss_out) {
    atomicAdd(sums + 0, (double)s0);
    atomicAdd(sums + 1, (double)s1);
    if (!loss_out) return;
    __threadfence();
    const unsigned long long ticket = atomicAdd(reinterpret_cast<unsigned long long*>(sums + 2), 1ull);
    if (ticket == n_blocks - 1) {
        __threadfence();
        const double a = *reinterpret_cast<volatile double*>(sums + 0), b = *reinterpret_cast<volatile double*>(sums + 1);
        *loss_out = (float)((double)k0 + (double)k1 * a + (double)k2 * b);
    }
}

__global__ void __launch_bounds__(kThreads) photometric_forward_kernel(
    int H, int W, const float* __restrict__ img1, const float* __restrict__ img2, float* __restrict__ dm_dmu1,
    float* __restrict__ dm_dsigma1, float* __restrict__ dm_dsigma12, float* __restrict__ ssim_map /* may be null */,
    double* __restrict__ sums /* [0] ssim, [1] l1, [2] ticket */, float k0, float k1, float k2,
    float* __restrict__ loss_out /* may be null */) {
    __shared__ float s1[kExt][kRowW], s2[kExt][kRowW];  // 49-word rows: conflict-free segment reads
    __shared__ float h[5][kExt][kTile + 1];
    __shared__ float s_red[2][kThreads / 32];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
    const size_t plane = (size_t)H * W;
    const float* p1 = img1 + c * plane;
    const float* p2 = img2 + c * plane;
    const int tid = threadIdx.x;
    const bool vec = rows_aligned(img1, W, plane) && rows_aligned(img2, W, plane);
    stage_plane(s1, p1, x0, y0, H, W, vec, tid);
    stage_plane(s2, p2, x0, y0, H, W, vec, tid);
    __syncthreads();
    // horizontal pass: 42 rows x 4 segments of 8 outputs
    if (tid < kExt * kSegs) {
        const int r = tid / kSegs, sx = (tid - r * kSegs) * kSeg;
        float a[kSegIn], b[kSegIn], t[kSegIn], o[kSeg];
#pragma unroll
        for (int j = 0; j < kSegIn; j++) { a[j] = s1[r][sx + j + kPadL]; b[j] = s2[r][sx + j + kPadL]; }
        conv8(a, o);
#pragma unroll
        for (int i = 0; i < kSeg; i++) h[0][r][sx + i] = o[i];
        conv8(b, o);
#pragma unroll
        for (int i = 0; i < kSeg; i++) h[1][r][sx + i] = o[i];
#pragma unroll
        for (int j = 0; j < kSegIn; j++) t[j] = a[j] * a[j];
        conv8(t, o);
#pragma unroll
        for (int i = 0; i < kSeg; i++) h[2][r][sx + i] = o[i];
#pragma unroll
        for (int j = 0; j < kSegIn; j++) t[j] = b[j] * b[j];
        conv8(t, o);
#pragma unroll
        for (int i = 0; i < kSeg; i++) h[3][r][sx + i] = o[i];
#pragma unroll
        for (int j = 0; j < kSegIn; j++) t[j] = a[j] * b[j];
        conv8(t, o);
#pragma unroll
        for (int i = 0; i < kSeg; i++) h[4][r][sx + i] = o[i];
    }
    __syncthreads();
    // vertical pass: 32 columns x 8 segments of 4 outputs (all 256 threads), then the SSIM map
    float ssim_sum = 0.f, l1_sum = 0.f;
    {
        const int lx = tid & (kTile - 1), sy = (tid / kTile) * kVSeg;
        float in[kVSegIn], mu1[kVSeg], mu2[kVSeg], e11[kVSeg], e22[kVSeg], e12[kVSeg];
#pragma unroll
        for (int j = 0; j < kVSegIn; j++) in[j] = h[0][sy + j][lx];
        conv_seg<kVSeg>(in, mu1);
#pragma unroll
        for (int j = 0; j < kVSegIn; j++) in[j] = h[1][sy + j][lx];
        conv_seg<kVSeg>(in, mu2);
#pragma unroll
        for (int j = 0; j < kVSegIn; j++) in[j] = h[2][sy + j][lx];
        conv_seg<kVSeg>(in, e11);
#pragma unroll
        for (int j = 0; j < kVSegIn; j++) in[j] = h[3][sy + j][lx];
        conv_seg<kVSeg>(in, e22);
#pragma unroll
        for (int j = 0; j < kVSegIn; j++) in[j] = h[4][sy + j][lx];
        conv_seg<kVSeg>(in, e12);
        const int gx = x0 + lx;
#pragma unroll
        for (int i = 0; i < kVSeg; i++) {
            const int gy = y0 + sy + i;
            if (gx < W && gy < H) {
                const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
                const float mu1_sq = mu1[i] * mu1[i], mu2_sq = mu2[i] * mu2[i], mu1_mu2 = mu1[i] * mu2[i];
                const float sigma1_sq = e11[i] - mu1_sq, sigma2_sq = e22[i] - mu2_sq, sigma12 = e12[i] - mu1_mu2;
                const float A = 2.f * mu1_mu2 + C1, B = 2.f * sigma12 + C2;
                const float Cc = mu1_sq + mu2_sq + C1, Dd = sigma1_sq + sigma2_sq + C2;
                const float inv = 1.f / (Cc * Dd);
                const float ssim_v = A * B * inv;
                const size_t o = c * plane + (size_t)gy * W + gx;
                // dS/dmu1 (through A, B = f(sigma12), Cc, Dd = f(sigma1_sq)); dS/d(E[x^2]); dS/d(E[xy])
                dm_dmu1[o] = 2.f * mu2[i] * (B - A) * inv - ssim_v * 2.f * mu1[i] * (1.f / Cc - 1.f / Dd);
                dm_dsigma1[o] = -ssim_v / Dd;
                dm_dsigma12[o] = 2.f * A * inv;
                if (ssim_map) ssim_map[o] = ssim_v;
                ssim_sum += ssim_v;
                l1_sum += fabsf(s1[sy + i + kHalo][lx + kHalo + kPadL] - s2[sy + i + kHalo][lx + kHalo + kPadL]);
            }
        }
    }
    ssim_sum = warp_sum(ssim_sum);
    l1_sum = warp_sum(l1_sum);
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = ssim_sum; s_red[1][tid >> 5] = l1_sum; }
    __syncthreads();
    if (tid == 0) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < kThreads / 32; w++) { a += s_red[0][w]; b += s_red[1][w]; }
        finish_sums(sums, a, b, gridDim.x * gridDim.y * gridDim.z, k0, k1, k2, loss_out);
    }
}

__global__ void __launch_bounds__(kThreads) photometric_backward_kernel(
    int H, int W, const float* __restrict__ img1, const float* __restrict__ img2, const float* __restrict__ dm_dmu1,
    const float* __restrict__ dm_dsigma1, const float* __restrict__ dm_dsigma12,
    const float* __restrict__ upstream /* device float[1] */, float k_ssim, float k_l1, float* __restrict__ dL_dimg1) {
    __shared__ float s[3][kExt][kRowW];
    __shared__ float h[3][kExt][kTile + 1];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
    const size_t plane = (size_t)H * W;
    const float* m0 = dm_dmu1 + c * plane;
    const float* m1 = dm_dsigma1 + c * plane;
    const float* m2 = dm_dsigma12 + c * plane;
    const int tid = threadIdx.x;
    const bool vec = rows_aligned(dm_dmu1, W, plane) && rows_aligned(dm_dsigma1, W, plane) &&
                     rows_aligned(dm_dsigma12, W, plane);
    stage_plane(s[0], m0, x0, y0, H, W, vec, tid);
    stage_plane(s[1], m1, x0, y0, H, W, vec, tid);
    stage_plane(s[2], m2, x0, y0, H, W, vec, tid);
    __syncthreads();
    if (tid < kExt * kSegs) {
        const int r = tid / kSegs, sx = (tid - r * kSegs) * kSeg;
        float in[kSegIn], o[kSeg];
#pragma unroll
        for (int q = 0; q < 3; q++) {
#pragma unroll
            for (int j = 0; j < kSegIn; j++) in[j] = s[q][r][sx + j + kPadL];
            conv8(in, o);
#pragma unroll
            for (int i = 0; i < kSeg; i++) h[q][r][sx + i] = o[i];
        }
    }
    __syncthreads();
    const int lx = tid & (kTile - 1), sy = (tid / kTile) * kVSeg;
    float in[kVSegIn], a[kVSeg], b[kVSeg], d[kVSeg];
#pragma unroll
    for (int j = 0; j < kVSegIn; j++) in[j] = h[0][sy + j][lx];
    conv_seg<kVSeg>(in, a);
#pragma unroll
    for (int j = 0; j < kVSegIn; j++) in[j] = h[1][sy + j][lx];
    conv_seg<kVSeg>(in, b);
#pragma unroll
    for (int j = 0; j < kVSegIn; j++) in[j] = h[2][sy + j][lx];
    conv_seg<kVSeg>(in, d);
    const int gx = x0 + lx;
    if (gx >= W) return;
    const float g = upstream[0], gS = g * k_ssim, gL1 = g * k_l1;
#pragma unroll
    for (int i = 0; i < kVSeg; i++) {
        const int gy = y0 + sy + i;
        if (gy >= H) break;
        const size_t o = c * plane + (size_t)gy * W + gx;
        const float x = img1[o], y = img2[o];
        const float diff = x - y;
        const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);  // torch.abs backward: sign(0) = 0
        dL_dimg1[o] = gS * (a[i] + 2.f * x * b[i] + y * d[i]) + gL1 * sgn;
    }
}

}  // namespace b3

using namespace b3;

extern "C" {

int b3gs_photometric_forward(int C, int H, int W, const float* img1, const float* img2, float* dm_dmu1,
                             float* dm_dsigma1_sq, float* dm_dsigma12, float* ssim_map, double* sums, float k_const,
                             float k_ssim, float k_l1, float* loss_out, void* stream) {
    if (C <= 0 || H <= 0 || W <= 0 || !img1 || !img2 || !dm_dmu1 || !dm_dsigma1_sq || !dm_dsigma12 || !sums) return -1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(sums, 0, 3 * sizeof(double), st) != cudaSuccess) return -2;
    dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, C), block(kThreads);
    photometric_forward_kernel<<<grid, block, 0, st>>>(H, W, img1, img2, dm_dmu1, dm_dsigma1_sq, dm_dsigma12, ssim_map,
                                                      sums, k_const, k_ssim, k_l1, loss_out);
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
    dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, C), block(kThreads);
    photometric_backward_kernel<<<grid, block, 0, st>>>(H, W, img1, img2, dm_dmu1, dm_dsigma1_sq, dm_dsigma12, upstream,
                                                       k_ssim, k_l1,
                                                       dL_dimg1);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // extern "C"
