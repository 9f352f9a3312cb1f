// common.cuh — shared types and the reference-exact arithmetic primitives.
//
// The reference (diff-gaussian-rasterization, built by nvcc with the default
// -fmad=true) does not compute what its C++ source literally says: ptxas contracts
// multiplies and adds into FMAs in a specific pattern.  Tile keys embed
// float_bits(depth) and the tile rectangle comes from ceil(3*sqrt(lambda)), so a
// 1-ulp difference in the preprocess flips sort keys and radii.  The pattern below
// was decoded from the reference's sm_100a SASS (profiles/ref_forward_sm100a.sass)
// and is written with explicit round-to-nearest intrinsics so that neither nvcc
// nor a future toolkit can re-associate it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define B3_TILE_X 16
#define B3_TILE_Y 16
#define B3_TILE_PIXELS 256

// Per-Gaussian record written by the preprocess and gathered by both composite
// kernels: three 16-byte vectors, 48 B, so one Gaussian is three LDG.128.
//   a = { x, y, tau, 0 }         pixel-space mean and the conservative culling threshold:
//                                alpha can reach 1/255 only where q(d) <= tau
//   b = { conic.x, conic.y, conic.z, opacity }
//   c = { r, g, b, depth }
#define B3_REC_VEC4 3
#define B3_GRAD_STRIDE 12  // packed per-Gaussian gradient accumulator, floats

// Layout of the packed gradient accumulator (one RED instruction per warp-survivor).
#define B3_G_MEAN2D_X 0
#define B3_G_MEAN2D_Y 1
#define B3_G_CONIC_X 2
#define B3_G_CONIC_Y 3
#define B3_G_CONIC_W 4
#define B3_G_OPACITY 5
#define B3_G_COLOR_R 6
#define B3_G_COLOR_G 7
#define B3_G_COLOR_B 8
#define B3_G_DEPTH 9

namespace b3 {

// a0*b0 + a1*b1 + a2*b2 as the reference's SASS evaluates it:
// the MIDDLE product is rounded alone, then the first and the last are fused in.
__device__ __forceinline__ float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
    return __fmaf_rn(a2, b2, __fmaf_rn(a0, b0, __fmul_rn(a1, b1)));
}

// matrix[i]*x + matrix[i+4]*y + matrix[i+8]*z + matrix[i+12]  (auxiliary.h:58-77)
__device__ __forceinline__ float xform(float m0, float m4, float m8, float m12, float x, float y, float z) {
    return __fadd_rn(dot3(m0, x, m4, y, m8, z), m12);
}

// ((v + 1.0) * S - 1.0) * 0.5 in DOUBLE, rounded once to float (auxiliary.h:41-44;
// SASS: F2F.F64.F32, DADD, DFMA, DMUL, F2F.F32.F64).
__device__ __forceinline__ float ndc2pix(float v, int S) {
    double d = __dadd_rn((double)v, 1.0);
    d = __fma_rn(d, (double)S, -1.0);
    d = __dmul_rn(d, 0.5);
    return __double2float_rn(d);
}

// Tile rectangle (auxiliary.h:46-56): float arithmetic, truncation toward zero,
// clamp to [0, grid].  `rf` is the radius already converted int -> float.
__device__ __forceinline__ void tile_rect(float px, float py, float rf, int grid_x, int grid_y,
                                          int& x0, int& y0, int& x1, int& y1) {
    x0 = min(grid_x, max(0, __float2int_rz(__fmul_rn(__fsub_rn(px, rf), 0.0625f))));
    y0 = min(grid_y, max(0, __float2int_rz(__fmul_rn(__fsub_rn(py, rf), 0.0625f))));
    x1 = min(grid_x, max(0, __float2int_rz(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(px, rf), 16.0f), -1.0f), 0.0625f))));
    y1 = min(grid_y, max(0, __float2int_rz(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(py, rf), 16.0f), -1.0f), 0.0625f))));
}

// Gaussian falloff exponent at pixel offset (dx,dy) (forward.cu:339, backward.cu:525).
// SASS: s = fma(dx, dx*cx, dy*(dy*cz)); power = fma(s, -0.5, -(dy*(dx*cy))).
__device__ __forceinline__ float gauss_power(float dx, float dy, float cx, float cy, float cz) {
    float s = __fmaf_rn(dx, __fmul_rn(dx, cx), __fmul_rn(dy, __fmul_rn(dy, cz)));
    return __fmaf_rn(s, -0.5f, -__fmul_rn(dy, __fmul_rn(dx, cy)));
}

// ---- the reference's parameter activations (scene/gaussian_model.py:27-40, :95-115) ------
// One definition for the stand-alone activation kernels (parameters.cu) and for the raw-
// parameter entry that fuses them into the preprocess and its backward: explicit rounding, so
// the fused and the unfused path are bit-identical whatever the surrounding code is.
__device__ __forceinline__ float act_scale(float raw) { return expf(raw); }                       // torch.exp
__device__ __forceinline__ float act_opacity(float raw) {                                        // torch.sigmoid
    return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-raw)));
}
// torch.nn.functional.normalize(q, dim=1) = q / max(||q||_2, 1e-12) as torch 2.x evaluates it on
// CUDA for a (P,4) tensor: the reduction keeps two strided accumulators, i.e.
// (x0^2 + x2^2) + (x1^2 + x3^2) with every square rounded first, then sqrt, then one IEEE
// division per component.  Found by tools/activation_probe.py (0 mismatching rows of 2M against
// torch 2.11; every other order of the four squares differs in ~15 % of the rows) — a last-bit
// difference in a rotation is enough to flip an alpha < 1/255 test somewhere in an image.
__device__ __forceinline__ float quat_norm(const float4& q) {
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(q.x, q.x), __fmul_rn(q.z, q.z)),
                                __fadd_rn(__fmul_rn(q.y, q.y), __fmul_rn(q.w, q.w))));
}
__device__ __forceinline__ float4 act_rotation(const float4& q) {
    const float n = fmaxf(quat_norm(q), 1e-12f);
    return make_float4(__fdiv_rn(q.x, n), __fdiv_rn(q.y, n), __fdiv_rn(q.z, n), __fdiv_rn(q.w, n));
}
// duals: gradient w.r.t. the raw parameter from the gradient u w.r.t. the activated one
__device__ __forceinline__ float act_opacity_grad(float raw, float u) {
    const float s = act_opacity(raw);
    return u * ((1.f - s) * s);
}
__device__ __forceinline__ float4 act_rotation_grad(const float4& q, const float4& u) {
    const float norm = quat_norm(q);
    if (norm > 1e-12f) {
        // y = q/|q|:  dq = (u - y (y.u)) / |q|
        const float inv = __fdiv_rn(1.f, norm);
        const float yx = q.x * inv, yy = q.y * inv, yz = q.z * inv, yw = q.w * inv;
        const float d = yx * u.x + yy * u.y + yz * u.z + yw * u.w;
        return make_float4((u.x - yx * d) * inv, (u.y - yy * d) * inv, (u.z - yz * d) * inv, (u.w - yw * d) * inv);
    }
    // clamped branch: y = q / 1e-12, the clamp passes no gradient to the norm
    return make_float4(u.x * 1e12f, u.y * 1e12f, u.z * 1e12f, u.w * 1e12f);
}

// ---- TMA 1-D bulk copy (cp.async.bulk, SASS: UBLKCP) + mbarrier, sm_90+/sm_100a -------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// one arrival + the number of bytes the bulk copies will deliver
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy; dst, src and bytes must be multiples of 16
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy (SASS: UBLKCP as well); dst, src and bytes multiples of 16.  The
// generic-proxy writes to `src` must be made visible to the async proxy first.
__device__ __forceinline__ void bulk_copy_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// wait until the bulk stores of this thread have finished READING shared memory
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a lost completion becomes a trap (an error), never a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); spin++) {
        if (spin > (1u << 24)) __trap();
    }
}

}  // namespace b3
