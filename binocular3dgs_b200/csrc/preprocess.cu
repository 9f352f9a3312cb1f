// preprocess.cu — K1: per-Gaussian forward preprocess, and mark_visible (K10).
//
// Reference behaviour: forward.cu:156-256 (preprocessCUDA) with its helpers
// computeCov3D :118-152, computeCov2D :74-113, computeColorFromSH :20-71 and
// auxiliary.h:41-164 (ndc2Pix, getRect, transformPoint*, in_frustum);
// rasterizer_impl.cu:54-66 (checkFrustum).
//
// Design: one thread per Gaussian; the 12-byte-stride arrays (means, scales) are
// staged through shared memory with perfectly coalesced 4-byte loads, rotations are
// one aligned 16-byte load.  All per-Gaussian results that the composite kernels
// consume are packed into ONE 48-byte record (three STG.128), so the hot kernels
// gather a Gaussian with three LDG.128 instead of chasing five arrays.
// The arithmetic is the reference's, operation for operation (see common.cuh).
#include <cstdio>
#include "common.cuh"
#include "kernels.h"

namespace b3 {

__device__ const float kSH_C0 = 0.28209479177387814f;
__device__ const float kSH_C1 = 0.4886025119029199f;

struct PreParams {
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
    // outputs
    int* radii;
    float4* records;
    float* depths;
    uint32_t* tiles_touched;
    uint8_t* clamped;
    uint2* rects;
    uint32_t* num_rendered;
};

// SH -> RGB, forward.cu:20-71.  dir components are IEEE divisions by the IEEE
// square root of dot3(dx,dx,dy,dy,dz,dz).  Each basis coefficient is formed as a
// scalar first and then fused into the running sum, in coefficient order.
// `sh0` holds coefficient 0; coefficient k >= 1 is sh[3k + c] (for the concatenated (P,M,3)
// tensor both are the same pointer; for the raw-parameter entry sh0 = f_dc and sh = f_rest
// shifted down by one coefficient).
__device__ __forceinline__ void sh_to_rgb(int deg, const float* __restrict__ sh0, const float* __restrict__ sh,
                                          float x, float y, float z, float& r, float& g, float& b) {
    float c0 = __fmul_rn(sh0[0], 0.28209479177387814f);
    float c1 = __fmul_rn(sh0[1], 0.28209479177387814f);
    float c2 = __fmul_rn(sh0[2], 0.28209479177387814f);
#define B3_SH_ACC(coef, k)                       \
    {                                            \
        float cf_ = (coef);                      \
        c0 = __fmaf_rn(cf_, sh[3 * (k) + 0], c0); \
        c1 = __fmaf_rn(cf_, sh[3 * (k) + 1], c1); \
        c2 = __fmaf_rn(cf_, sh[3 * (k) + 2], c2); \
    }
    if (deg > 0) {
        B3_SH_ACC(-__fmul_rn(y, 0.4886025119029199f), 1);
        B3_SH_ACC(__fmul_rn(z, 0.4886025119029199f), 2);
        B3_SH_ACC(-__fmul_rn(x, 0.4886025119029199f), 3);
        if (deg > 1) {
            float xx = __fmul_rn(x, x), yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
            float xy = __fmul_rn(x, y), yz = __fmul_rn(y, z), xz = __fmul_rn(x, z);
            float zz2 = __fadd_rn(zz, zz);
            B3_SH_ACC(__fmul_rn(xy, 1.0925484305920792f), 4);
            B3_SH_ACC(__fmul_rn(yz, -1.0925484305920792f), 5);
            B3_SH_ACC(__fmul_rn(__fsub_rn(__fsub_rn(zz2, xx), yy), 0.31539156525252005f), 6);
            B3_SH_ACC(__fmul_rn(xz, -1.0925484305920792f), 7);
            float xx_yy = __fsub_rn(xx, yy);
            B3_SH_ACC(__fmul_rn(xx_yy, 0.5462742152960396f), 8);
            if (deg > 2) {
                float t4 = __fsub_rn(__fmaf_rn(zz, 4.0f, -xx), yy);  // 4zz - xx - yy
                B3_SH_ACC(__fmul_rn(__fmul_rn(y, -0.5900435899266435f), __fmaf_rn(xx, 3.0f, -yy)), 9);
                B3_SH_ACC(__fmul_rn(__fmul_rn(xy, 2.890611442640554f), z), 10);
                B3_SH_ACC(__fmul_rn(__fmul_rn(y, -0.4570457994644658f), t4), 11);
                B3_SH_ACC(__fmul_rn(__fmul_rn(z, 0.3731763325901154f),
                                    __fmaf_rn(yy, -3.0f, __fmaf_rn(xx, -3.0f, zz2))), 12);
                B3_SH_ACC(__fmul_rn(t4, __fmul_rn(x, -0.4570457994644658f)), 13);
                B3_SH_ACC(__fmul_rn(xx_yy, __fmul_rn(z, 1.445305721320277f)), 14);
                B3_SH_ACC(__fmul_rn(__fmul_rn(x, -0.5900435899266435f), __fmaf_rn(yy, -3.0f, xx)), 15);
            }
        }
    }
#undef B3_SH_ACC
    r = c0; g = c1; b = c2;
}

// Conservative cut-off for culling: this Gaussian can reach alpha >= 1/255 at pixel
// offset d only if  q(d) = 0.5*(cx dx^2 + cz dy^2) + cy dx dy <= tau,  tau = ln(255*o).
// The composite kernels minimise q exactly over a warp's pixel rectangle and skip the
// entry when the minimum exceeds the value returned here.  This is NOT in the
// reference; it must never exclude a pixel the exact per-pixel test
// (power <= 0 && o*exp(power) >= 1/255) would accept, hence tau is inflated (1 % + 0.05,
// which dominates the float error of evaluating q for conics that are comfortably
// positive definite) and culling is disabled (+huge) for anything unusual.
//   returns  < 0     : can never contribute (o*G < 1/255 for every G <= 1)
//            >= 3e38 : never cull
__device__ __forceinline__ float cull_threshold(float px, float py, float cx, float cy, float cz, float o) {
    const float kNever = 3.0e38f;
    if (!(o >= 0.0f)) return kNever;          // NaN / negative opacity
    if (o < 0.0039f) return -1.0f;            // 0.0039 < 1/255: o*G < 1/255 whenever G <= 1
    const float det = cx * cz - cy * cy;
    if (!(cx > 0.0f) || !(cz > 0.0f) || !(det > 1e-3f * cx * cz)) return kNever;  // not safely PSD
    if (!(cx < 1.0e4f) || !(cz < 1.0e4f)) return kNever;
    if (!(fabsf(px) < 1.0e5f) || !(fabsf(py) < 1.0e5f)) return kNever;
    return 1.01f * logf(255.0f * o) + 0.05f;
}

// 6 blocks/SM (40 registers, no spills): 888 resident blocks hold the 782 blocks of a 200k-
// Gaussian scene in ONE wave; at 5 (48 registers, 740 slots) the last 42 blocks formed a
// second wave of this latency-bound kernel (B200, lego: 19.8 -> 16.6 us).
__global__ void __launch_bounds__(256, 6) preprocess_kernel(PreParams p) {
    __shared__ float s_mean[256 * 3];
    __shared__ float s_scale[256 * 3];
    const int base = blockIdx.x * 256;
    const int n = min(256, p.P - base);
    // coalesced staging of the 12-byte-stride inputs
    for (int i = threadIdx.x; i < n * 3; i += 256) {
        s_mean[i] = p.means3D[(size_t)base * 3 + i];
        if (p.scales) s_scale[i] = p.raw ? act_scale(p.scales[(size_t)base * 3 + i]) : p.scales[(size_t)base * 3 + i];
    }
    __syncthreads();
    const int t = threadIdx.x;
    const bool in_range = t < n;
    const int idx = base + (in_range ? t : 0);

    const float x = s_mean[3 * t], y = s_mean[3 * t + 1], z = s_mean[3 * t + 2];
    const float* __restrict__ V = p.viewmatrix;
    const float* __restrict__ Pm = p.projmatrix;

    int radius_out = 0;
    uint32_t tiles = 0;
    float depth = 0.0f;
    float4 ra = make_float4(0.f, 0.f, -1.f, 0.f), rb = make_float4(0.f, 0.f, 0.f, 0.f),
           rc = make_float4(0.f, 0.f, 0.f, 0.f);
    uint8_t clamp_bits = 0;
    uint2 rect = make_uint2(0u, 0u);

    // in_frustum (auxiliary.h:139-164): only the view-space z is live.
    const float pz = xform(V[2], V[6], V[10], V[14], x, y, z);
    bool alive = in_range && !(pz <= 0.2f);
    if (in_range && !alive && p.prefiltered) {
        printf("Point is filtered although prefiltered is set. This shouldn't happen!");
        __trap();
    }
    if (alive) {
        // clip-space projection (forward.cu:197-200)
        const float hx = xform(Pm[0], Pm[4], Pm[8], Pm[12], x, y, z);
        const float hy = xform(Pm[1], Pm[5], Pm[9], Pm[13], x, y, z);
        const float hw = xform(Pm[3], Pm[7], Pm[11], Pm[15], x, y, z);
        const float p_w = __frcp_rn(__fadd_rn(hw, 0.0000001f));
        const float projx = __fmul_rn(hx, p_w), projy = __fmul_rn(hy, p_w);

        // 3D covariance (forward.cu:118-152) or the precomputed one
        float c0, c1, c2, c3, c4, c5;
        if (p.cov3D_precomp) {
            const float* c = p.cov3D_precomp + (size_t)idx * 6;
            c0 = c[0]; c1 = c[1]; c2 = c[2]; c3 = c[3]; c4 = c[4]; c5 = c[5];
        } else {
            float4 q = reinterpret_cast<const float4*>(p.rotations)[idx];
            if (p.raw) q = act_rotation(q);
            const float qr = q.x, qx = q.y, qy = q.z, qz = q.w;
            const float sx = __fmul_rn(s_scale[3 * t], p.scale_modifier);
            const float sy = __fmul_rn(s_scale[3 * t + 1], p.scale_modifier);
            const float sz = __fmul_rn(s_scale[3 * t + 2], p.scale_modifier);
            const float xz = __fmul_rn(qx, qz), rx = __fmul_rn(qr, qx), rz = __fmul_rn(qr, qz);
            const float yy = __fmul_rn(qy, qy), zz = __fmul_rn(qz, qz);
            const float xz_p_ry = __fmaf_rn(qr, qy, xz), xz_m_ry = __fmaf_rn(-qr, qy, xz);
            const float yz_m_rx = __fmaf_rn(qy, qz, -rx), yz_p_rx = __fmaf_rn(qy, qz, rx);
            const float xy_m_rz = __fmaf_rn(qx, qy, -rz), xy_p_rz = __fmaf_rn(qx, qy, rz);
            const float xx_p_yy = __fmaf_rn(qx, qx, yy);
            const float yy_p_zz = __fadd_rn(yy, zz);
            const float xx_p_zz = __fmaf_rn(qx, qx, zz);
            const float R00 = __fsub_rn(1.0f, __fadd_rn(yy_p_zz, yy_p_zz));
            const float R11 = __fsub_rn(1.0f, __fadd_rn(xx_p_zz, xx_p_zz));
            const float R22 = __fsub_rn(1.0f, __fadd_rn(xx_p_yy, xx_p_yy));
            // M = S * R (glm column-major): M[i][j] = s_j * R[i][j]
            const float M00 = __fmul_rn(sx, R00), M01 = __fmul_rn(sy, __fadd_rn(xy_m_rz, xy_m_rz)),
                        M02 = __fmul_rn(sz, __fadd_rn(xz_p_ry, xz_p_ry));
            const float M10 = __fmul_rn(sx, __fadd_rn(xy_p_rz, xy_p_rz)), M11 = __fmul_rn(sy, R11),
                        M12 = __fmul_rn(sz, __fadd_rn(yz_m_rx, yz_m_rx));
            const float M20 = __fmul_rn(sx, __fadd_rn(xz_m_ry, xz_m_ry)),
                        M21 = __fmul_rn(sy, __fadd_rn(yz_p_rx, yz_p_rx)), M22 = __fmul_rn(sz, R22);
            c0 = dot3(M00, M00, M01, M01, M02, M02);
            c1 = dot3(M00, M10, M01, M11, M02, M12);
            c2 = dot3(M00, M20, M01, M21, M02, M22);
            c3 = dot3(M10, M10, M11, M11, M12, M12);
            c4 = dot3(M10, M20, M11, M21, M12, M22);
            c5 = dot3(M20, M20, M21, M21, M22, M22);
        }

        // EWA 2D covariance (forward.cu:74-113)
        const float tx = xform(V[0], V[4], V[8], V[12], x, y, z);
        const float ty = xform(V[1], V[5], V[9], V[13], x, y, z);
        const float tz = pz;
        const float limx = __fmul_rn(1.3f, p.tan_fovx), limy = __fmul_rn(1.3f, p.tan_fovy);
        const float kx = fminf(fmaxf(__fdiv_rn(tx, tz), -limx), limx);
        const float ky = fminf(fmaxf(__fdiv_rn(ty, tz), -limy), limy);
        const float tz2 = __fmul_rn(tz, tz);
        const float J00 = __fdiv_rn(p.focal_x, tz);
        const float J02 = __fdiv_rn(__fmul_rn(__fmul_rn(tz, -kx), p.focal_x), tz2);
        const float J11 = __fdiv_rn(p.focal_y, tz);
        const float J12 = __fdiv_rn(__fmul_rn(__fmul_rn(tz, -ky), p.focal_y), tz2);
        const float T00 = __fmaf_rn(V[2], J02, __fmul_rn(V[0], J00));
        const float T01 = __fmaf_rn(V[6], J02, __fmul_rn(V[4], J00));
        const float T02 = __fmaf_rn(V[10], J02, __fmul_rn(V[8], J00));
        const float T10 = __fmaf_rn(V[2], J12, __fmul_rn(V[1], J11));
        const float T11 = __fmaf_rn(V[6], J12, __fmul_rn(V[5], J11));
        const float T12 = __fmaf_rn(V[10], J12, __fmul_rn(V[9], J11));
        const float A00 = dot3(T00, c0, T01, c1, T02, c2);
        const float A10 = dot3(T00, c1, T01, c3, T02, c4);
        const float A20 = dot3(T00, c2, T01, c4, T02, c5);
        const float A01 = dot3(T10, c0, T11, c1, T12, c2);
        const float A11 = dot3(T10, c1, T11, c3, T12, c4);
        const float A21 = dot3(T10, c2, T11, c4, T12, c5);
        const float ca = __fadd_rn(dot3(T00, A00, T01, A10, T02, A20), 0.3f);
        const float cb = dot3(T00, A01, T01, A11, T02, A21);
        const float cc = __fadd_rn(dot3(T10, A01, T11, A11, T12, A21), 0.3f);

        // invert (forward.cu:219-223)
        const float det = __fmaf_rn(ca, cc, -__fmul_rn(cb, cb));
        if (det != 0.0f) {
            const float det_inv = __frcp_rn(det);
            const float conx = __fmul_rn(cc, det_inv);
            const float cony = __fmul_rn(cb, -det_inv);
            const float conz = __fmul_rn(ca, det_inv);

            // radius and tile rectangle (forward.cu:229-237)
            const float mid = __fmul_rn(__fadd_rn(ca, cc), 0.5f);
            const float sq = __fsqrt_rn(fmaxf(__fmaf_rn(mid, mid, -det), 0.1f));
            const float lam = fmaxf(__fadd_rn(mid, sq), __fsub_rn(mid, sq));
            const int ri = __float2int_ru(__fmul_rn(__fsqrt_rn(lam), 3.0f));
            const float rf = (float)ri;
            const float pix_x = ndc2pix(projx, p.W), pix_y = ndc2pix(projy, p.H);
            int x0, y0, x1, y1;
            tile_rect(pix_x, pix_y, rf, p.grid_x, p.grid_y, x0, y0, x1, y1);
            const uint32_t area = (uint32_t)(x1 - x0) * (uint32_t)(y1 - y0);
            if (area != 0) {
                float cr, cg, cbl;
                if (p.colors_precomp == nullptr) {
                    const float* camp = p.cam_pos;
                    const float dx = __fsub_rn(x, camp[0]), dy = __fsub_rn(y, camp[1]), dz = __fsub_rn(z, camp[2]);
                    const float len = __fsqrt_rn(dot3(dx, dx, dy, dy, dz, dz));
                    float sr, sg, sb;
                    const float* sh0 = p.raw ? p.shs + (size_t)idx * 3 : p.shs + (size_t)idx * p.M * 3;
                    const float* shk = p.raw ? p.shs_rest + (size_t)idx * (p.M - 1) * 3 - 3 : sh0;
                    sh_to_rgb(p.D, sh0, shk, __fdiv_rn(dx, len), __fdiv_rn(dy, len), __fdiv_rn(dz, len), sr, sg, sb);
                    sr = __fadd_rn(sr, 0.5f); sg = __fadd_rn(sg, 0.5f); sb = __fadd_rn(sb, 0.5f);
                    if (sr < 0.0f) { clamp_bits |= 1; sr = 0.0f; }
                    if (sg < 0.0f) { clamp_bits |= 2; sg = 0.0f; }
                    if (sb < 0.0f) { clamp_bits |= 4; sb = 0.0f; }
                    cr = sr; cg = sg; cbl = sb;
                } else {
                    const float* c = p.colors_precomp + (size_t)idx * 3;
                    cr = c[0]; cg = c[1]; cbl = c[2];
                }
                const float op = p.raw ? act_opacity(p.opacities[idx]) : p.opacities[idx];
                const float tau = cull_threshold(pix_x, pix_y, conx, cony, conz, op);
                depth = pz;
                radius_out = ri;
                tiles = area;
                rect = make_uint2((uint32_t)x0 | ((uint32_t)y0 << 16), (uint32_t)(x1 - x0) | ((uint32_t)(y1 - y0) << 16));
                ra = make_float4(pix_x, pix_y, tau, 0.0f);
                rb = make_float4(conx, cony, conz, op);
                rc = make_float4(cr, cg, cbl, pz);
            }
        }
    }
    if (in_range) {
        p.radii[idx] = radius_out;
        p.tiles_touched[idx] = tiles;
        p.depths[idx] = depth;
        p.clamped[idx] = clamp_bits;
        p.rects[idx] = rect;
        float4* rec = p.records + (size_t)idx * B3_REC_VEC4;
        rec[0] = ra; rec[1] = rb; rec[2] = rc;
    }
    // R = sum(tiles_touched): warp shuffle reduce, one shared atomic per warp, one global
    // atomic per block — replaces the reference's full prefix sum + read of its last element.
    // Also the OR and AND of the visible depth keys: bytes in which all keys agree need
    // no radix pass (counter words: [0] R, [1] OR, [2] AND).
    __shared__ uint32_t s_total, s_or, s_and;
    if (threadIdx.x == 0) { s_total = 0; s_or = 0; s_and = 0xffffffffu; }
    __syncthreads();
    const uint32_t dbits = __float_as_uint(depth);
    const uint32_t wsum = __reduce_add_sync(0xffffffffu, tiles);
    const uint32_t wor = __reduce_or_sync(0xffffffffu, tiles ? dbits : 0u);
    const uint32_t wand = __reduce_and_sync(0xffffffffu, tiles ? dbits : 0xffffffffu);
    if ((threadIdx.x & 31) == 0 && wsum) {
        atomicAdd(&s_total, wsum);
        atomicOr(&s_or, wor);
        atomicAnd(&s_and, wand);
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_total) {
        atomicAdd(p.num_rendered, s_total);
        atomicOr(p.num_rendered + 1, s_or);
        atomicAnd(p.num_rendered + 2, s_and);
    }
}

__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ V,
                                    unsigned char* __restrict__ present) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float x = means3D[3 * (size_t)idx], y = means3D[3 * (size_t)idx + 1], z = means3D[3 * (size_t)idx + 2];
    const float pz = xform(V[2], V[6], V[10], V[14], x, y, z);
    present[idx] = !(pz <= 0.2f);
}

void launch_preprocess(const PreprocessArgs& a, cudaStream_t stream) {
    PreParams p;
    p.P = a.P; p.D = a.D; p.M = a.M;
    p.raw = a.raw; p.shs_rest = a.shs_rest;
    p.means3D = a.means3D; p.scales = a.scales; p.scale_modifier = a.scale_modifier;
    p.rotations = a.rotations; p.opacities = a.opacities; p.shs = a.shs;
    p.cov3D_precomp = a.cov3D_precomp; p.colors_precomp = a.colors_precomp;
    p.viewmatrix = a.viewmatrix; p.projmatrix = a.projmatrix; p.cam_pos = a.cam_pos;
    p.W = a.W; p.H = a.H; p.tan_fovx = a.tan_fovx; p.tan_fovy = a.tan_fovy;
    p.focal_x = a.focal_x; p.focal_y = a.focal_y;
    p.grid_x = a.grid_x; p.grid_y = a.grid_y; p.prefiltered = a.prefiltered;
    p.radii = a.radii; p.records = a.records; p.depths = a.depths;
    p.tiles_touched = a.tiles_touched; p.clamped = a.clamped; p.rects = a.rects; p.num_rendered = a.num_rendered;
    preprocess_kernel<<<(a.P + 255) / 256, 256, 0, stream>>>(p);
    count_launch();
}

void launch_mark_visible(int P, const float* means3D, const float* viewmatrix, unsigned char* present,
                         cudaStream_t stream) {
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, viewmatrix, present);
    count_launch();
}

}  // namespace b3
