// parameters.cu — per-step parameter plumbing.
//
// What the reference runs around the rasterizer every iteration, each as a string of small
// elementwise torch kernels over P Gaussians (scene/gaussian_model.py):
//   :95-115  get_scaling = exp, get_rotation = normalize, get_opacity = sigmoid,
//            get_features = cat(f_dc, f_rest)                  (per render, + autograd duals)
//   :154-163 torch.optim.Adam(lr=0, eps=1e-15) over six parameter groups   (train.py:192)
//   :307-309 opacity_decay: opacity <- logit(sigmoid(opacity) * factor)    (train.py:163-165)
//   :409-411 add_densification_stats + max_radii2D update (train.py:170-171), boolean-mask
//            indexing (nonzero -> host sync) three times
//
// Here each is ONE streaming kernel (HBM-bound; bytes per Gaussian stated at each):
//   activate_forward_kernel / activate_backward_kernel   raw parameters -> rasterizer inputs
//   adam_multi_kernel        all parameter groups in one launch (multi-tensor table)
//   opacity_decay_kernel     in place
//   densify_stats_kernel     no index lists, no host sync
#include "../../include/b3gs.h"
#include "common.cuh"
#include "kernels.h"

namespace b3 {

// ------------------------------------------------------------------ activations
// forward: read (4 + 3M + 3 + 4) floats, write the same count.  One thread per Gaussian
// for the small attributes; the SH copy is a flat coalesced loop.
__global__ void __launch_bounds__(256) activate_forward_kernel(int P, int M, const float* __restrict__ f_dc,
                                                              const float* __restrict__ f_rest,
                                                              const float* __restrict__ opacity_raw,
                                                              const float* __restrict__ scaling_raw,
                                                              const float* __restrict__ rotation_raw,
                                                              float* __restrict__ shs, float* __restrict__ opacities,
                                                              float* __restrict__ scales,
                                                              float* __restrict__ rotations) {
    // 32-bit indices: the host side rejects P * 3M >= 2^31
    const unsigned stride = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    // torch.cat((f_dc, f_rest), dim=1): (P,1,3) ++ (P,M-1,3) -> (P,M,3)
    const unsigned row = 3u * M, n_sh = (unsigned)P * row;
    for (unsigned i = tid; i < n_sh; i += stride) {
        const unsigned g = i / row, k = i - g * row;
        shs[i] = k < 3 ? f_dc[g * 3 + k] : f_rest[g * (row - 3) + (k - 3)];
    }
    for (unsigned i = tid; i < 3u * P; i += stride) scales[i] = act_scale(scaling_raw[i]);
    for (unsigned g = tid; g < (unsigned)P; g += stride) {
        opacities[g] = act_opacity(opacity_raw[g]);
        reinterpret_cast<float4*>(rotations)[g] = act_rotation(reinterpret_cast<const float4*>(rotation_raw)[g]);
    }
}

// backward: recomputes the activations from the raw parameters (nothing saved).
__global__ void __launch_bounds__(256) activate_backward_kernel(
    int P, int M, const float* __restrict__ opacity_raw, const float* __restrict__ scaling_raw,
    const float* __restrict__ rotation_raw, const float* __restrict__ g_shs, const float* __restrict__ g_opacities,
    const float* __restrict__ g_scales, const float* __restrict__ g_rotations, float* __restrict__ g_f_dc,
    float* __restrict__ g_f_rest, float* __restrict__ g_opacity_raw, float* __restrict__ g_scaling_raw,
    float* __restrict__ g_rotation_raw) {
    const unsigned stride = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned row = 3u * M, n_sh = (unsigned)P * row;
    if (g_shs) {
        for (unsigned i = tid; i < n_sh; i += stride) {
            const unsigned g = i / row, k = i - g * row;
            const float v = g_shs[i];
            if (k < 3) g_f_dc[g * 3 + k] = v;
            else g_f_rest[g * (row - 3) + (k - 3)] = v;
        }
    }
    if (g_scales)
        for (unsigned i = tid; i < 3u * P; i += stride) g_scaling_raw[i] = g_scales[i] * act_scale(scaling_raw[i]);
    for (unsigned g = tid; g < (unsigned)P; g += stride) {
        if (g_opacities) g_opacity_raw[g] = act_opacity_grad(opacity_raw[g], g_opacities[g]);
        if (g_rotations)
            reinterpret_cast<float4*>(g_rotation_raw)[g] = act_rotation_grad(
                reinterpret_cast<const float4*>(rotation_raw)[g], reinterpret_cast<const float4*>(g_rotations)[g]);
    }
}

// ------------------------------------------------------------------ multi-tensor Adam
// torch.optim.Adam (no weight decay, no amsgrad, not maximize), the update of
// torch/optim/adam.py::_single_tensor_adam:
//     m <- lerp(m, g, 1-beta1);  v <- v*beta2 + (1-beta2) g g
//     p <- p - (lr / (1-beta1^t)) * m / (sqrt(v)/sqrt(1-beta2^t) + eps)
// step_size and 1/sqrt(bias_correction2) are formed on the host in double like torch does.
// 28 B read + 12 B written per element.
struct AdamTable {
    B3gsAdamTensor t[B3GS_ADAM_MAX_TENSORS];
    unsigned block_start[B3GS_ADAM_MAX_TENSORS + 1];
    int n_tensors;
    float w1, beta2, w2, eps;  // w1 = 1-beta1, w2 = 1-beta2, differences taken in double on the host
};

constexpr int kAdamThreads = 256, kAdamPerThread = 4, kAdamPerBlock = kAdamThreads * kAdamPerThread;

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, float w1, float beta2, float w2,
                                            float step_size, float inv_bc2_sqrt, float eps) {
    m = fmaf(w1, g - m, m);
    v = fmaf(w2 * g, g, v * beta2);
    const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
    p = fmaf(-step_size, __fdiv_rn(m, denom), p);
}

__global__ void __launch_bounds__(kAdamThreads) adam_multi_kernel(const __grid_constant__ AdamTable tab) {
    int k = 0;
#pragma unroll
    for (int i = 1; i < B3GS_ADAM_MAX_TENSORS; i++)
        if (i < tab.n_tensors && blockIdx.x >= tab.block_start[i]) k = i;
    const B3gsAdamTensor& t = tab.t[k];
    const size_t base = (size_t)(blockIdx.x - tab.block_start[k]) * kAdamPerBlock + (size_t)threadIdx.x * kAdamPerThread;
    if (base >= t.n) return;
    const float w1 = tab.w1, w2 = tab.w2;
    const bool vec = (base + kAdamPerThread <= t.n) &&
                     (((uintptr_t)t.param | (uintptr_t)t.grad | (uintptr_t)t.exp_avg | (uintptr_t)t.exp_avg_sq) & 15) == 0;
    if (vec) {
        float4 p = *reinterpret_cast<float4*>(t.param + base);
        const float4 g = *reinterpret_cast<const float4*>(t.grad + base);
        float4 m = *reinterpret_cast<float4*>(t.exp_avg + base);
        float4 v = *reinterpret_cast<float4*>(t.exp_avg_sq + base);
        adam_update(p.x, g.x, m.x, v.x, w1, tab.beta2, w2, t.step_size, t.inv_bias_correction2_sqrt, tab.eps);
        adam_update(p.y, g.y, m.y, v.y, w1, tab.beta2, w2, t.step_size, t.inv_bias_correction2_sqrt, tab.eps);
        adam_update(p.z, g.z, m.z, v.z, w1, tab.beta2, w2, t.step_size, t.inv_bias_correction2_sqrt, tab.eps);
        adam_update(p.w, g.w, m.w, v.w, w1, tab.beta2, w2, t.step_size, t.inv_bias_correction2_sqrt, tab.eps);
        *reinterpret_cast<float4*>(t.param + base) = p;
        *reinterpret_cast<float4*>(t.exp_avg + base) = m;
        *reinterpret_cast<float4*>(t.exp_avg_sq + base) = v;
    } else {
        for (size_t i = base; i < base + kAdamPerThread && i < t.n; i++) {
            float p = t.param[i], m = t.exp_avg[i], v = t.exp_avg_sq[i];
            adam_update(p, t.grad[i], m, v, w1, tab.beta2, w2, t.step_size, t.inv_bias_correction2_sqrt, tab.eps);
            t.param[i] = p; t.exp_avg[i] = m; t.exp_avg_sq[i] = v;
        }
    }
}

// ------------------------------------------------------------------ opacity decay
// scene/gaussian_model.py:307-309: opacity_raw <- inverse_sigmoid(sigmoid(opacity_raw) * factor),
// inverse_sigmoid(x) = log(x / (1 - x)) (utils/general_utils.py:18-19).  8 B/Gaussian.
__global__ void __launch_bounds__(256) opacity_decay_kernel(int P, float factor, float* __restrict__ opacity_raw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float y = __fmul_rn(__fdiv_rn(1.f, 1.f + expf(-opacity_raw[i])), factor);
    opacity_raw[i] = logf(__fdiv_rn(y, __fsub_rn(1.f, y)));
}

// ------------------------------------------------------------------ densification statistics
// train.py:170-171 + scene/gaussian_model.py:409-411 with visibility_filter = radii > 0
// (gaussian_renderer/__init__.py:101):
//     max_radii2D[vis] = max(max_radii2D[vis], radii[vis])
//     xyz_gradient_accum[vis] += || viewspace_grad[vis, :2] ||;   denom[vis] += 1
// 28 B read + 12 B written per visible Gaussian, 4 B per culled one.
__global__ void __launch_bounds__(256) densify_stats_kernel(int P, const float* __restrict__ viewspace_grad,
                                                           const int* __restrict__ radii,
                                                           float* __restrict__ xyz_gradient_accum,
                                                           float* __restrict__ denom, float* __restrict__ max_radii2D) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int r = radii[i];
    if (r <= 0) return;
    const float gx = viewspace_grad[3 * (size_t)i], gy = viewspace_grad[3 * (size_t)i + 1];
    xyz_gradient_accum[i] += sqrtf(gx * gx + gy * gy);
    denom[i] += 1.f;
    if (max_radii2D) max_radii2D[i] = fmaxf(max_radii2D[i], (float)r);
}

static int grid_for(size_t n, int threads, int cap) {
    size_t b = (n + threads - 1) / threads;
    return (int)(b < 1 ? 1 : (b > (size_t)cap ? (size_t)cap : b));
}

}  // namespace b3

using namespace b3;

extern "C" {

int b3gs_activate_forward(int P, int M, const float* f_dc, const float* f_rest, const float* opacity_raw,
                          const float* scaling_raw, const float* rotation_raw, float* shs, float* opacities,
                          float* scales, float* rotations, void* stream) {
    if (P < 0 || M < 1 || (M > 1 && !f_rest) || (size_t)P * 3 * M >= (1ull << 31)) return -1;
    if (P == 0) return 0;
    if (!f_dc || !opacity_raw || !scaling_raw || !rotation_raw || !shs || !opacities || !scales || !rotations) return -1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // 148 SMs x 8 resident blocks of 256 threads; grid-stride beyond that
    activate_forward_kernel<<<grid_for((size_t)P * 3 * M, 256, 148 * 8), 256, 0, st>>>(
        P, M, f_dc, f_rest, opacity_raw, scaling_raw, rotation_raw, shs, opacities, scales, rotations);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int b3gs_activate_backward(int P, int M, const float* opacity_raw, const float* scaling_raw, const float* rotation_raw,
                           const float* g_shs, const float* g_opacities, const float* g_scales,
                           const float* g_rotations, float* g_f_dc, float* g_f_rest, float* g_opacity_raw,
                           float* g_scaling_raw, float* g_rotation_raw, void* stream) {
    if (P < 0 || M < 1 || (size_t)P * 3 * M >= (1ull << 31)) return -1;
    if (P == 0) return 0;
    if ((g_shs && (!g_f_dc || (M > 1 && !g_f_rest))) || (g_opacities && (!g_opacity_raw || !opacity_raw)) ||
        (g_scales && (!g_scaling_raw || !scaling_raw)) || (g_rotations && (!g_rotation_raw || !rotation_raw)))
        return -1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    activate_backward_kernel<<<grid_for((size_t)P * 3 * M, 256, 148 * 8), 256, 0, st>>>(
        P, M, opacity_raw, scaling_raw, rotation_raw, g_shs, g_opacities, g_scales, g_rotations, g_f_dc, g_f_rest,
        g_opacity_raw, g_scaling_raw, g_rotation_raw);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int b3gs_adam_multi(int n_tensors, const B3gsAdamTensor* tensors, double beta1, double beta2, double eps,
                    void* stream) {
    if (n_tensors < 0 || n_tensors > B3GS_ADAM_MAX_TENSORS || (n_tensors > 0 && !tensors)) return -1;
    AdamTable tab;
    tab.n_tensors = 0;
    tab.w1 = (float)(1.0 - beta1); tab.beta2 = (float)beta2; tab.w2 = (float)(1.0 - beta2); tab.eps = (float)eps;
    unsigned blocks = 0;
    for (int i = 0; i < n_tensors; i++) {
        const B3gsAdamTensor& t = tensors[i];
        if (t.n == 0) continue;
        if (!t.param || !t.grad || !t.exp_avg || !t.exp_avg_sq) return -1;
        tab.t[tab.n_tensors] = t;
        tab.block_start[tab.n_tensors] = blocks;
        blocks += (unsigned)((t.n + kAdamPerBlock - 1) / kAdamPerBlock);
        tab.n_tensors++;
    }
    for (int i = tab.n_tensors; i <= B3GS_ADAM_MAX_TENSORS; i++) tab.block_start[i] = blocks;
    if (blocks == 0) return 0;
    adam_multi_kernel<<<blocks, kAdamThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(tab);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int b3gs_opacity_decay(int P, float factor, float* opacity_raw, void* stream) {
    if (P < 0 || (P > 0 && !opacity_raw)) return -1;
    if (P == 0) return 0;
    opacity_decay_kernel<<<(P + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(P, factor, opacity_raw);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int b3gs_densify_stats(int P, const float* viewspace_grad, const int* radii, float* xyz_gradient_accum, float* denom,
                       float* max_radii2D, void* stream) {
    if (P < 0 || (P > 0 && (!viewspace_grad || !radii || !xyz_gradient_accum || !denom))) return -1;
    if (P == 0) return 0;
    densify_stats_kernel<<<(P + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        P, viewspace_grad, radii, xyz_gradient_accum, denom, max_radii2D);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // extern "C"
