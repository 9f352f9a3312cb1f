// preprocess_backward.cu — K8 + K9 fused: per-Gaussian backward of the preprocess.
//
// Reference behaviour: backward.cu:144-274 (computeCov2DCUDA), :347-412
// (preprocessCUDA), :20-139 (SH backward), :278-341 (cov3D backward),
// auxiliary.h:107-117 (dnormvdv).  The reference runs two kernels that re-read the
// same inputs and accumulate into pre-zeroed outputs; here one streaming pass reads
// the packed accumulator written by the composite backward and WRITES every
// output element (zeros for Gaussians with radius <= 0), so no memset of the ten
// gradient tensors is needed (rasterize_points.cu:158-167 zero-fills them).
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace b3 {

// column-major 3x3 with glm semantics: m[c][r]
struct M3 {
    float m[3][3];
};
__device__ __forceinline__ M3 m3_mul(const M3& A, const M3& B) {
    M3 R;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) R.m[i][j] = A.m[0][j] * B.m[i][0] + A.m[1][j] * B.m[i][1] + A.m[2][j] * B.m[i][2];
    return R;
}
__device__ __forceinline__ M3 m3_t(const M3& A) {
    M3 R;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) R.m[i][j] = A.m[j][i];
    return R;
}

__device__ __forceinline__ void cov3d_from_scale_rot(float sx, float sy, float sz, float4 q, M3& R, M3& M,
                                                     float (&c)[6]) {
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    R.m[0][0] = 1.f - 2.f * (y * y + z * z); R.m[0][1] = 2.f * (x * y - r * z); R.m[0][2] = 2.f * (x * z + r * y);
    R.m[1][0] = 2.f * (x * y + r * z); R.m[1][1] = 1.f - 2.f * (x * x + z * z); R.m[1][2] = 2.f * (y * z - r * x);
    R.m[2][0] = 2.f * (x * z - r * y); R.m[2][1] = 2.f * (y * z + r * x); R.m[2][2] = 1.f - 2.f * (x * x + y * y);
    const float s[3] = {sx, sy, sz};
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) M.m[i][j] = s[j] * R.m[i][j];
    // Sigma = M^T M : Sigma[i][j] = sum_k M[j][k] * M[i][k]
    c[0] = M.m[0][0] * M.m[0][0] + M.m[0][1] * M.m[0][1] + M.m[0][2] * M.m[0][2];
    c[1] = M.m[0][0] * M.m[1][0] + M.m[0][1] * M.m[1][1] + M.m[0][2] * M.m[1][2];
    c[2] = M.m[0][0] * M.m[2][0] + M.m[0][1] * M.m[2][1] + M.m[0][2] * M.m[2][2];
    c[3] = M.m[1][0] * M.m[1][0] + M.m[1][1] * M.m[1][1] + M.m[1][2] * M.m[1][2];
    c[4] = M.m[1][0] * M.m[2][0] + M.m[1][1] * M.m[2][1] + M.m[1][2] * M.m[2][2];
    c[5] = M.m[2][0] * M.m[2][0] + M.m[2][1] * M.m[2][1] + M.m[2][2] * M.m[2][2];
}

template <int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks) preprocess_backward_kernel(PreBackwardArgs p) {
    const int idx = p.first + blockIdx.x * blockDim.x + threadIdx.x;
    // Optional (p.stage, B3GS_PREBWD_STAGE=1; OFF by default): outputs with a 12-byte (or 12 M-byte) stride per
    // Gaussian — dL_dsh, dL_dmean3D, dL_dscale, dL_dmean2D — staged per block in shared memory
    // ([component][thread], padded) and written out with consecutive threads on consecutive floats, so that a
    // warp store covers 4 full sectors instead of 12-32 partial ones (ncu r02t: 15.7 sectors per store request).
    // Measured on B200: it LOSES — 1M Gaussians 82.4 vs 75.3 us, 200k 25.1 vs 25.0, 300k 32.0 vs 31.3 — the L2
    // already merges the partial sectors (DRAM writes = the algorithmic bytes either way) and the barrier plus
    // the extra shared-memory round trip cost more than the shorter store queue saves.
    extern __shared__ float s_out[];
    const bool staged = p.stage != 0;
    constexpr int kPitch = kThreads + 1;
    const int c_mean3d = 3 * p.M, c_scale = c_mean3d + 3, c_mean2d = c_scale + 3;
    const bool live = idx < p.first + p.count;
    if (!live && !staged) return;
    const size_t i = (size_t)idx;
    if (live) {

    float dmean[3] = {0.f, 0.f, 0.f};
    float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float dscale[3] = {0.f, 0.f, 0.f};
    float4 drot = make_float4(0.f, 0.f, 0.f, 0.f);
    float g[B3_GRAD_STRIDE];
#pragma unroll
    for (int k = 0; k < B3_GRAD_STRIDE; k++) g[k] = 0.f;

    const bool visible = p.radii[idx] > 0;
    const bool acc = p.accumulate != 0;
    const int ncoef = (p.D + 1) * (p.D + 1);
    // coefficient 0 goes to dsh0[0..2], coefficient k >= 1 to dshk[3k..3k+2]: one tensor (P,M,3), or —
    // raw-parameter entry — d f_dc (P,1,3) and d f_rest (P,M-1,3) shifted down by one coefficient
    float* dsh0 = p.dL_dsh ? (p.raw ? p.dL_dsh + i * 3 : p.dL_dsh + i * p.M * 3) : nullptr;
    float* dshk = p.dL_dsh ? (p.raw ? p.dL_dsh_rest + i * (p.M - 1) * 3 - 3 : dsh0) : nullptr;
    float4 q_raw = make_float4(0.f, 0.f, 0.f, 0.f);
    float scale_act[3] = {1.f, 1.f, 1.f};   // d scale / d raw scale

    if (visible) {
        const float4* gp = reinterpret_cast<const float4*>(p.grads + i * B3_GRAD_STRIDE);
        const float4 g0 = gp[0], g1 = gp[1], g2 = gp[2];
        g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w;
        g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
        g[8] = g2.x; g[9] = g2.y;

        const float mx = p.means3D[3 * i], my = p.means3D[3 * i + 1], mz = p.means3D[3 * i + 2];
        const float* __restrict__ V = p.viewmatrix;
        const float* __restrict__ Pm = p.projmatrix;

        // ---- 3D covariance (recomputed; the forward does not store it)
        float c3d[6];
        M3 R, M;
        float sx = 0.f, sy = 0.f, sz = 0.f;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.cov3D_precomp) {
#pragma unroll
            for (int k = 0; k < 6; k++) c3d[k] = p.cov3D_precomp[6 * i + k];
        } else {
            float s0 = p.scales[3 * i], s1 = p.scales[3 * i + 1], s2 = p.scales[3 * i + 2];
            q = reinterpret_cast<const float4*>(p.rotations)[idx];
            if (p.raw) {
                s0 = act_scale(s0); s1 = act_scale(s1); s2 = act_scale(s2);
                scale_act[0] = s0; scale_act[1] = s1; scale_act[2] = s2;
                q_raw = q;
                q = act_rotation(q);
            }
            sx = p.scale_modifier * s0;
            sy = p.scale_modifier * s1;
            sz = p.scale_modifier * s2;
            cov3d_from_scale_rot(sx, sy, sz, q, R, M, c3d);
        }

        // ---- K8: conic -> cov2D -> cov3D, and mean through J (backward.cu:144-274)
        {
            const float dconx = g[B3_G_CONIC_X], dcony = g[B3_G_CONIC_Y], dconz = g[B3_G_CONIC_W];
            float tx = V[0] * mx + V[4] * my + V[8] * mz + V[12];
            float ty = V[1] * mx + V[5] * my + V[9] * mz + V[13];
            const float tz = V[2] * mx + V[6] * my + V[10] * mz + V[14];
            const float limx = 1.3f * p.tan_fovx, limy = 1.3f * p.tan_fovy;
            const float txtz = tx / tz, tytz = ty / tz;
            tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
            ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
            const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
            const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
            const float hx = p.focal_x, hy = p.focal_y;

            M3 J, W, Vrk;
            J.m[0][0] = hx / tz; J.m[0][1] = 0.f; J.m[0][2] = -(hx * tx) / (tz * tz);
            J.m[1][0] = 0.f; J.m[1][1] = hy / tz; J.m[1][2] = -(hy * ty) / (tz * tz);
            J.m[2][0] = 0.f; J.m[2][1] = 0.f; J.m[2][2] = 0.f;
            W.m[0][0] = V[0]; W.m[0][1] = V[4]; W.m[0][2] = V[8];
            W.m[1][0] = V[1]; W.m[1][1] = V[5]; W.m[1][2] = V[9];
            W.m[2][0] = V[2]; W.m[2][1] = V[6]; W.m[2][2] = V[10];
            Vrk.m[0][0] = c3d[0]; Vrk.m[0][1] = c3d[1]; Vrk.m[0][2] = c3d[2];
            Vrk.m[1][0] = c3d[1]; Vrk.m[1][1] = c3d[3]; Vrk.m[1][2] = c3d[4];
            Vrk.m[2][0] = c3d[2]; Vrk.m[2][1] = c3d[4]; Vrk.m[2][2] = c3d[5];
            const M3 T = m3_mul(W, J);
            const M3 cov2D = m3_mul(m3_mul(m3_t(T), m3_t(Vrk)), T);
            const float a = cov2D.m[0][0] + 0.3f, b = cov2D.m[0][1], c = cov2D.m[1][1] + 0.3f;
            const float denom = a * c - b * b;
            float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
            const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
            if (denom2inv != 0.f) {
                dL_da = denom2inv * (-c * c * dconx + 2 * b * c * dcony + (denom - a * c) * dconz);
                dL_dc = denom2inv * (-a * a * dconz + 2 * a * b * dcony + (denom - a * c) * dconx);
                dL_db = denom2inv * 2 * (b * c * dconx - (denom + 2 * b * b) * dcony + a * b * dconz);
                dcov[0] = T.m[0][0] * T.m[0][0] * dL_da + T.m[0][0] * T.m[1][0] * dL_db + T.m[1][0] * T.m[1][0] * dL_dc;
                dcov[3] = T.m[0][1] * T.m[0][1] * dL_da + T.m[0][1] * T.m[1][1] * dL_db + T.m[1][1] * T.m[1][1] * dL_dc;
                dcov[5] = T.m[0][2] * T.m[0][2] * dL_da + T.m[0][2] * T.m[1][2] * dL_db + T.m[1][2] * T.m[1][2] * dL_dc;
                dcov[1] = 2 * T.m[0][0] * T.m[0][1] * dL_da + (T.m[0][0] * T.m[1][1] + T.m[0][1] * T.m[1][0]) * dL_db +
                          2 * T.m[1][0] * T.m[1][1] * dL_dc;
                dcov[2] = 2 * T.m[0][0] * T.m[0][2] * dL_da + (T.m[0][0] * T.m[1][2] + T.m[0][2] * T.m[1][0]) * dL_db +
                          2 * T.m[1][0] * T.m[1][2] * dL_dc;
                dcov[4] = 2 * T.m[0][2] * T.m[0][1] * dL_da + (T.m[0][1] * T.m[1][2] + T.m[0][2] * T.m[1][1]) * dL_db +
                          2 * T.m[1][1] * T.m[1][2] * dL_dc;
            }
            // dL/dT (upper 2x3)
            float dT0[3], dT1[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float tv0 = T.m[0][0] * Vrk.m[k][0] + T.m[0][1] * Vrk.m[k][1] + T.m[0][2] * Vrk.m[k][2];
                const float tv1 = T.m[1][0] * Vrk.m[k][0] + T.m[1][1] * Vrk.m[k][1] + T.m[1][2] * Vrk.m[k][2];
                dT0[k] = 2 * tv0 * dL_da + tv1 * dL_db;
                dT1[k] = 2 * tv1 * dL_dc + tv0 * dL_db;
            }
            const float dJ00 = W.m[0][0] * dT0[0] + W.m[0][1] * dT0[1] + W.m[0][2] * dT0[2];
            const float dJ02 = W.m[2][0] * dT0[0] + W.m[2][1] * dT0[1] + W.m[2][2] * dT0[2];
            const float dJ11 = W.m[1][0] * dT1[0] + W.m[1][1] * dT1[1] + W.m[1][2] * dT1[2];
            const float dJ12 = W.m[2][0] * dT1[0] + W.m[2][1] * dT1[1] + W.m[2][2] * dT1[2];
            const float itz = 1.f / tz, itz2 = itz * itz, itz3 = itz2 * itz;
            const float dtx = x_grad_mul * -hx * itz2 * dJ02;
            const float dty = y_grad_mul * -hy * itz2 * dJ12;
            const float dtz = -hx * itz2 * dJ00 - hy * itz2 * dJ11 + (2 * hx * tx) * itz3 * dJ02 + (2 * hy * ty) * itz3 * dJ12;
            dmean[0] = V[0] * dtx + V[1] * dty + V[2] * dtz;
            dmean[1] = V[4] * dtx + V[5] * dty + V[6] * dtz;
            dmean[2] = V[8] * dtx + V[9] * dty + V[10] * dtz;
        }

        // ---- K9: projection and depth terms (backward.cu:372-403)
        {
            const float m_hom_w = Pm[3] * mx + Pm[7] * my + Pm[11] * mz + Pm[15];
            const float m_w = 1.0f / (m_hom_w + 0.0000001f);
            const float mul1 = (Pm[0] * mx + Pm[4] * my + Pm[8] * mz + Pm[12]) * m_w * m_w;
            const float mul2 = (Pm[1] * mx + Pm[5] * my + Pm[9] * mz + Pm[13]) * m_w * m_w;
            const float d2x = g[B3_G_MEAN2D_X], d2y = g[B3_G_MEAN2D_Y];
            dmean[0] += (Pm[0] * m_w - Pm[3] * mul1) * d2x + (Pm[1] * m_w - Pm[3] * mul2) * d2y;
            dmean[1] += (Pm[4] * m_w - Pm[7] * mul1) * d2x + (Pm[5] * m_w - Pm[7] * mul2) * d2y;
            dmean[2] += (Pm[8] * m_w - Pm[11] * mul1) * d2x + (Pm[9] * m_w - Pm[11] * mul2) * d2y;
            const float mul3 = V[2] * mx + V[6] * my + V[10] * mz + V[14];
            const float dd = g[B3_G_DEPTH];
            dmean[0] += (V[2] - V[3] * mul3) * dd;
            dmean[1] += (V[6] - V[7] * mul3) * dd;
            dmean[2] += (V[10] - V[11] * mul3) * dd;
        }

        // ---- SH backward (backward.cu:20-139)
        if (p.shs) {
            const float* sh = p.raw ? p.shs_rest + i * (p.M - 1) * 3 - 3 : p.shs + i * p.M * 3;  // only k >= 1 is read
            const float dox = mx - p.campos[0], doy = my - p.campos[1], doz = mz - p.campos[2];
            const float len = sqrtf(dox * dox + doy * doy + doz * doz);
            const float x = dox / len, y = doy / len, z = doz / len;
            const uint8_t cl = p.clamped[idx];
            float dRGB[3] = {g[B3_G_COLOR_R], g[B3_G_COLOR_G], g[B3_G_COLOR_B]};
            if (cl & 1) dRGB[0] = 0.f;
            if (cl & 2) dRGB[1] = 0.f;
            if (cl & 4) dRGB[2] = 0.f;
            float dRGBdx[3] = {0.f, 0.f, 0.f}, dRGBdy[3] = {0.f, 0.f, 0.f}, dRGBdz[3] = {0.f, 0.f, 0.f};
#define SHV(k, ch) sh[3 * (k) + (ch)]
#define DSH(k, coef)                                             \
    {                                                            \
        const float cf_ = (coef);                                \
        float* d_ = ((k) == 0 ? dsh0 : dshk) + 3 * (k);          \
        if (staged) {                                            \
            s_out[(3 * (k) + 0) * kPitch + threadIdx.x] = cf_ * dRGB[0]; \
            s_out[(3 * (k) + 1) * kPitch + threadIdx.x] = cf_ * dRGB[1]; \
            s_out[(3 * (k) + 2) * kPitch + threadIdx.x] = cf_ * dRGB[2]; \
        } else if (acc) {                                        \
            d_[0] += cf_ * dRGB[0];                              \
            d_[1] += cf_ * dRGB[1];                              \
            d_[2] += cf_ * dRGB[2];                              \
        } else {                                                 \
            d_[0] = cf_ * dRGB[0];                               \
            d_[1] = cf_ * dRGB[1];                               \
            d_[2] = cf_ * dRGB[2];                               \
        }                                                        \
    }
            const float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f;
            const float C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                 -1.0925484305920792f, 0.5462742152960396f};
            const float C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                 0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                 -0.5900435899266435f};
            DSH(0, C0);
            if (p.D > 0) {
                DSH(1, -C1 * y);
                DSH(2, C1 * z);
                DSH(3, -C1 * x);
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    dRGBdx[ch] = -C1 * SHV(3, ch);
                    dRGBdy[ch] = -C1 * SHV(1, ch);
                    dRGBdz[ch] = C1 * SHV(2, ch);
                }
                if (p.D > 1) {
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    DSH(4, C2[0] * xy);
                    DSH(5, C2[1] * yz);
                    DSH(6, C2[2] * (2.f * zz - xx - yy));
                    DSH(7, C2[3] * xz);
                    DSH(8, C2[4] * (xx - yy));
#pragma unroll
                    for (int ch = 0; ch < 3; ch++) {
                        dRGBdx[ch] += C2[0] * y * SHV(4, ch) + C2[2] * 2.f * -x * SHV(6, ch) + C2[3] * z * SHV(7, ch) +
                                      C2[4] * 2.f * x * SHV(8, ch);
                        dRGBdy[ch] += C2[0] * x * SHV(4, ch) + C2[1] * z * SHV(5, ch) + C2[2] * 2.f * -y * SHV(6, ch) +
                                      C2[4] * 2.f * -y * SHV(8, ch);
                        dRGBdz[ch] += C2[1] * y * SHV(5, ch) + C2[2] * 2.f * 2.f * z * SHV(6, ch) + C2[3] * x * SHV(7, ch);
                    }
                    if (p.D > 2) {
                        DSH(9, C3[0] * y * (3.f * xx - yy));
                        DSH(10, C3[1] * xy * z);
                        DSH(11, C3[2] * y * (4.f * zz - xx - yy));
                        DSH(12, C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
                        DSH(13, C3[4] * x * (4.f * zz - xx - yy));
                        DSH(14, C3[5] * z * (xx - yy));
                        DSH(15, C3[6] * x * (xx - 3.f * yy));
#pragma unroll
                        for (int ch = 0; ch < 3; ch++) {
                            dRGBdx[ch] += C3[0] * SHV(9, ch) * 3.f * 2.f * xy + C3[1] * SHV(10, ch) * yz +
                                          C3[2] * SHV(11, ch) * -2.f * xy + C3[3] * SHV(12, ch) * -3.f * 2.f * xz +
                                          C3[4] * SHV(13, ch) * (-3.f * xx + 4.f * zz - yy) +
                                          C3[5] * SHV(14, ch) * 2.f * xz + C3[6] * SHV(15, ch) * 3.f * (xx - yy);
                            dRGBdy[ch] += C3[0] * SHV(9, ch) * 3.f * (xx - yy) + C3[1] * SHV(10, ch) * xz +
                                          C3[2] * SHV(11, ch) * (-3.f * yy + 4.f * zz - xx) +
                                          C3[3] * SHV(12, ch) * -3.f * 2.f * yz + C3[4] * SHV(13, ch) * -2.f * xy +
                                          C3[5] * SHV(14, ch) * -2.f * yz + C3[6] * SHV(15, ch) * -3.f * 2.f * xy;
                            dRGBdz[ch] += C3[1] * SHV(10, ch) * xy + C3[2] * SHV(11, ch) * 4.f * 2.f * yz +
                                          C3[3] * SHV(12, ch) * 3.f * (2.f * zz - xx - yy) +
                                          C3[4] * SHV(13, ch) * 4.f * 2.f * xz + C3[5] * SHV(14, ch) * (xx - yy);
                        }
                    }
                }
            }
#undef SHV
#undef DSH
            if (staged) {
                for (int c = 3 * ncoef; c < 3 * p.M; c++) s_out[c * kPitch + threadIdx.x] = 0.f;
            } else {
                for (int k = ncoef; k < p.M && !acc; k++) {
                    dshk[3 * k] = 0.f; dshk[3 * k + 1] = 0.f; dshk[3 * k + 2] = 0.f;
                }
            }
            const float ddx = dRGBdx[0] * dRGB[0] + dRGBdx[1] * dRGB[1] + dRGBdx[2] * dRGB[2];
            const float ddy = dRGBdy[0] * dRGB[0] + dRGBdy[1] * dRGB[1] + dRGBdy[2] * dRGB[2];
            const float ddz = dRGBdz[0] * dRGB[0] + dRGBdz[1] * dRGB[1] + dRGBdz[2] * dRGB[2];
            // dnormvdv (auxiliary.h:107-117)
            const float sum2 = dox * dox + doy * doy + doz * doz;
            const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
            dmean[0] += ((+sum2 - dox * dox) * ddx - doy * dox * ddy - doz * dox * ddz) * invsum32;
            dmean[1] += (-dox * doy * ddx + (sum2 - doy * doy) * ddy - doz * doy * ddz) * invsum32;
            dmean[2] += (-dox * doz * ddx - doy * doz * ddy + (sum2 - doz * doz) * ddz) * invsum32;
        }

        // ---- cov3D -> scale, rotation (backward.cu:278-341)
        if (p.scales) {
            M3 dS;
            dS.m[0][0] = dcov[0]; dS.m[0][1] = 0.5f * dcov[1]; dS.m[0][2] = 0.5f * dcov[2];
            dS.m[1][0] = 0.5f * dcov[1]; dS.m[1][1] = dcov[3]; dS.m[1][2] = 0.5f * dcov[4];
            dS.m[2][0] = 0.5f * dcov[2]; dS.m[2][1] = 0.5f * dcov[4]; dS.m[2][2] = dcov[5];
            M3 dM = m3_mul(M, dS);
#pragma unroll
            for (int a2 = 0; a2 < 3; a2++)
#pragma unroll
                for (int b2 = 0; b2 < 3; b2++) dM.m[a2][b2] *= 2.0f;
            const M3 Rt = m3_t(R);
            M3 dMt = m3_t(dM);
            dscale[0] = Rt.m[0][0] * dMt.m[0][0] + Rt.m[0][1] * dMt.m[0][1] + Rt.m[0][2] * dMt.m[0][2];
            dscale[1] = Rt.m[1][0] * dMt.m[1][0] + Rt.m[1][1] * dMt.m[1][1] + Rt.m[1][2] * dMt.m[1][2];
            dscale[2] = Rt.m[2][0] * dMt.m[2][0] + Rt.m[2][1] * dMt.m[2][1] + Rt.m[2][2] * dMt.m[2][2];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                dMt.m[0][k] *= sx; dMt.m[1][k] *= sy; dMt.m[2][k] *= sz;
            }
            const float r = q.x, x = q.y, y = q.z, z = q.w;
            drot.x = 2 * z * (dMt.m[0][1] - dMt.m[1][0]) + 2 * y * (dMt.m[2][0] - dMt.m[0][2]) +
                     2 * x * (dMt.m[1][2] - dMt.m[2][1]);
            drot.y = 2 * y * (dMt.m[1][0] + dMt.m[0][1]) + 2 * z * (dMt.m[2][0] + dMt.m[0][2]) +
                     2 * r * (dMt.m[1][2] - dMt.m[2][1]) - 4 * x * (dMt.m[2][2] + dMt.m[1][1]);
            drot.z = 2 * x * (dMt.m[1][0] + dMt.m[0][1]) + 2 * r * (dMt.m[2][0] - dMt.m[0][2]) +
                     2 * z * (dMt.m[1][2] + dMt.m[2][1]) - 4 * y * (dMt.m[2][2] + dMt.m[0][0]);
            drot.w = 2 * r * (dMt.m[0][1] - dMt.m[1][0]) + 2 * x * (dMt.m[2][0] + dMt.m[0][2]) +
                     2 * y * (dMt.m[1][2] + dMt.m[2][1]) - 4 * z * (dMt.m[1][1] + dMt.m[0][0]);
            // the reference scales by the modifier implicitly through s = mod*scale:
            // dL/dscale as written is w.r.t. the UNSCALED parameter only through s; it
            // stores dot(Rt, dMt) without the modifier (backward.cu:322-325) — kept.
        }
    } else if (dsh0 && staged) {
        for (int c = 0; c < 3 * p.M; c++) s_out[c * kPitch + threadIdx.x] = 0.f;
    } else if (dsh0 && !acc) {
        dsh0[0] = 0.f; dsh0[1] = 0.f; dsh0[2] = 0.f;
        for (int k = 3; k < p.M * 3; k++) dshk[k] = 0.f;
    }

    // ---- write every output element
    if (staged) {
        s_out[(c_mean2d + 0) * kPitch + threadIdx.x] = g[B3_G_MEAN2D_X];
        s_out[(c_mean2d + 1) * kPitch + threadIdx.x] = g[B3_G_MEAN2D_Y];
        s_out[(c_mean2d + 2) * kPitch + threadIdx.x] = 0.f;
    } else {
        p.dL_dmean2D[3 * i] = g[B3_G_MEAN2D_X];
        p.dL_dmean2D[3 * i + 1] = g[B3_G_MEAN2D_Y];
        p.dL_dmean2D[3 * i + 2] = 0.f;
    }
    // intermediates a caller may not want (NULL): the reference materialises all of them
    if (p.dL_dconic)
        reinterpret_cast<float4*>(p.dL_dconic)[idx] = make_float4(g[B3_G_CONIC_X], g[B3_G_CONIC_Y], 0.f, g[B3_G_CONIC_W]);
    if (p.dL_dcolor) {
        p.dL_dcolor[3 * i] = g[B3_G_COLOR_R];
        p.dL_dcolor[3 * i + 1] = g[B3_G_COLOR_G];
        p.dL_dcolor[3 * i + 2] = g[B3_G_COLOR_B];
    }
    if (p.dL_ddepth) p.dL_ddepth[idx] = g[B3_G_DEPTH];
    if (p.dL_dcov3D) {
#pragma unroll
        for (int k = 0; k < 6; k++) p.dL_dcov3D[6 * i + k] = dcov[k];
    }
    if (p.raw && visible) {
        // chain rule through the activations: the outputs are gradients of the RAW parameters
        g[B3_G_OPACITY] = act_opacity_grad(p.opacities[idx], g[B3_G_OPACITY]);
        dscale[0] *= scale_act[0]; dscale[1] *= scale_act[1]; dscale[2] *= scale_act[2];
        drot = act_rotation_grad(q_raw, drot);
    }
    // The five parameter gradients: overwritten, or — B3GS_BWD_ACCUMULATE, the second view of a
    // step writing into the same exchange bucket — added to what the earlier view left there
    // (dL_dsh was accumulated where it was formed, above).
    if (staged) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            s_out[(c_mean3d + k) * kPitch + threadIdx.x] = dmean[k];
            s_out[(c_scale + k) * kPitch + threadIdx.x] = dscale[k];
        }
        if (acc) {
            p.dL_dopacity[idx] += g[B3_G_OPACITY];
            float4* r4 = reinterpret_cast<float4*>(p.dL_drot) + idx;
            const float4 o = *r4;
            *r4 = make_float4(o.x + drot.x, o.y + drot.y, o.z + drot.z, o.w + drot.w);
        } else {
            p.dL_dopacity[idx] = g[B3_G_OPACITY];
            reinterpret_cast<float4*>(p.dL_drot)[idx] = drot;
        }
    } else if (acc) {
        p.dL_dopacity[idx] += g[B3_G_OPACITY];
        p.dL_dmean3D[3 * i] += dmean[0];
        p.dL_dmean3D[3 * i + 1] += dmean[1];
        p.dL_dmean3D[3 * i + 2] += dmean[2];
        p.dL_dscale[3 * i] += dscale[0];
        p.dL_dscale[3 * i + 1] += dscale[1];
        p.dL_dscale[3 * i + 2] += dscale[2];
        float4* r4 = reinterpret_cast<float4*>(p.dL_drot) + idx;
        const float4 o = *r4;
        *r4 = make_float4(o.x + drot.x, o.y + drot.y, o.z + drot.z, o.w + drot.w);
    } else {
        p.dL_dopacity[idx] = g[B3_G_OPACITY];
        p.dL_dmean3D[3 * i] = dmean[0];
        p.dL_dmean3D[3 * i + 1] = dmean[1];
        p.dL_dmean3D[3 * i + 2] = dmean[2];
        p.dL_dscale[3 * i] = dscale[0];
        p.dL_dscale[3 * i + 1] = dscale[1];
        p.dL_dscale[3 * i + 2] = dscale[2];
        reinterpret_cast<float4*>(p.dL_drot)[idx] = drot;
    }
    }  // live
    if (!staged) return;
    __syncthreads();
    // copy-out: element e of a region = (Gaussian e / ncomp of the block, component e % ncomp)
    const int i0 = p.first + blockIdx.x * kThreads;
    const int nvalid = min(kThreads, p.first + p.count - i0);
    auto flush = [&](float* dst, int comp0, int ncomp, bool add) {
        const int total = nvalid * ncomp;
        int t = (int)threadIdx.x / ncomp, c = (int)threadIdx.x - t * ncomp;
        const int dt = kThreads / ncomp, dc = kThreads - dt * ncomp;
        for (int e = threadIdx.x; e < total; e += kThreads) {
            const float v = s_out[(comp0 + c) * kPitch + t];
            if (add) dst[e] += v; else dst[e] = v;
            t += dt; c += dc;
            if (c >= ncomp) { c -= ncomp; t++; }
        }
    };
    const bool add = p.accumulate != 0;
    if (p.dL_dsh) {
        if (p.raw) {
            flush(p.dL_dsh + (size_t)i0 * 3, 0, 3, add);
            if (p.M > 1) flush(p.dL_dsh_rest + (size_t)i0 * (p.M - 1) * 3, 3, 3 * (p.M - 1), add);
        } else {
            flush(p.dL_dsh + (size_t)i0 * p.M * 3, 0, 3 * p.M, add);
        }
    }
    flush(p.dL_dmean3D + (size_t)i0 * 3, c_mean3d, 3, add);
    flush(p.dL_dscale + (size_t)i0 * 3, c_scale, 3, add);
    flush(p.dL_dmean2D + (size_t)i0 * 3, c_mean2d, 3, false);
}

// Gaussians [first, first + count) (count < 0: all P)
void launch_preprocess_backward(const PreBackwardArgs& a0, cudaStream_t stream, int first, int count) {
    PreBackwardArgs a = a0;
    a.first = first;
    a.count = count < 0 ? a.P - first : count;
    if (a.count <= 0) return;
    // threads per block x minimum blocks per SM (register cap); B3GS_PREBWD_SHAPE = 0..5 overrides.
    // Measured on B200, 1M / 200k Gaussians: 256x3 (80 regs, 184 B spilled) 87.0 / 27.2 us, 128x6 80.5 / 25.8,
    // 128x5 84.5 / 27.1, 128x4 (no spills) 90.8 / 27.1, 64x10 (96 regs, 40 B spilled) 78.7 / 25.2; the same shape in
    // the current form of the kernel (16 B spilled): 75.3 / 25.0 us.  More 64-thread blocks per SM beat fewer spills:
    // 64x8 (125 regs, none) 88.4 / 27.1, 64x9 76.5 / 25.5, 64x10 75.4 / 25.1, 64x12 (80 regs, 108 B) 72.8 / 24.8 =
    // 3.39 TB/s, 52 % of the measured HBM peak (default), 64x14 (72 regs, 208 B) 73.0 / 25.6, 64x16 (64 regs) 81.0 / 30.3.
    static const int shape = [] { const char* e = getenv("B3GS_PREBWD_SHAPE"); return e ? atoi(e) : 5; }();
    // staged outputs (B3GS_PREBWD_STAGE=1; measured slower, see the kernel): (3 M + 9) columns of threads + 1 floats
    static const int stage = [] { const char* e = getenv("B3GS_PREBWD_STAGE"); return e ? atoi(e) : 0; }();
    const int threads = shape >= 4 && shape <= 5 ? 64 : (shape >= 1 && shape <= 3 ? 128 : 256);
    size_t smem = (size_t)(3 * a.M + 9) * (threads + 1) * sizeof(float);
    a.stage = stage && smem <= 48 * 1024;
    if (!a.stage) smem = 0;
    switch (shape) {
        case 1: preprocess_backward_kernel<128, 6><<<(a.count + 127) / 128, 128, smem, stream>>>(a); break;   // 85 regs
        case 2: preprocess_backward_kernel<128, 5><<<(a.count + 127) / 128, 128, smem, stream>>>(a); break;   // 102 regs
        case 3: preprocess_backward_kernel<128, 4><<<(a.count + 127) / 128, 128, smem, stream>>>(a); break;   // 128 regs
        case 4: preprocess_backward_kernel<64, 10><<<(a.count + 63) / 64, 64, smem, stream>>>(a); break;      // 96 regs
        case 5: preprocess_backward_kernel<64, 12><<<(a.count + 63) / 64, 64, smem, stream>>>(a); break;      // 80 regs
        default: preprocess_backward_kernel<256, 3><<<(a.count + 255) / 256, 256, smem, stream>>>(a); break;  // 80 regs
    }
    count_launch();
}

}  // namespace b3
