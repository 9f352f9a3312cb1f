/*
 * oracle.c — CPU restatement of the reference rasterizer.  TEST INFRASTRUCTURE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (binocular3dgs_b200/) never does.
 *
 * What it restates (reference = /root/reference/submodules/diff-gaussian-rasterization):
 *   orc_preprocess           forward.cu:156-256 (preprocessCUDA) + :20-71 (SH), :74-113
 *                            (cov2D), :118-152 (cov3D); auxiliary.h:41-77,139-164
 *   orc_binning              rasterizer_impl.cu:70-111 (duplicateWithKeys), :304-309
 *                            (stable sort by tile|depth key), :116-138 (identifyTileRanges)
 *   orc_render_forward       forward.cu:261-381 (renderCUDA)
 *   orc_render_backward      backward.cu:415-601 (renderCUDA)
 *   orc_preprocess_backward  backward.cu:144-274 (computeCov2DCUDA), :347-412
 *                            (preprocessCUDA), :20-139 (SH), :278-341 (cov3D)
 *
 * Parity pinning: the reference ships no golden vectors or tests, and
 * it cannot run in the build container (CUDA-only, no GPU).  The oracle is pinned
 * against outputs of the reference's own kernels (oracle/_ref/libdgr_ref.so) generated
 * on the GPU box by tests/golden/make_golden.py and committed under tests/golden/;
 * tests/test_oracle_golden.py checks the oracle against them on CPU.
 *
 * Arithmetic: the reference is compiled by nvcc with -fmad=true, so what its GPU
 * executes is NOT the source-level expression tree but the FMA-contracted one.  The
 * forward preprocess below follows the contraction pattern decoded from the reference's
 * sm_100a SASS (profiles/ref_forward_sm100a.sass), written with explicit fmaf(); this
 * file must be compiled with -ffp-contract=off.  Every operation in the preprocess is
 * IEEE (add, mul, fma, div, sqrt, double ops), so depths/radii/tile counts/keys are
 * bit-exact on the CPU.  The composite uses libm expf where the GPU uses
 * ex2.approx-based expf (<= 2 ulp), so images agree to ~1e-6, not bitwise.
 * Gradient sums are accumulated in double (the reference's float atomics make its own
 * gradients run-to-run non-deterministic in the last bits).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16

/* ------------------------------------------------------------------ helpers */
static inline float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
    /* a0*b0 + a1*b1 + a2*b2 as contracted by ptxas: middle product rounded alone */
    return fmaf(a2, b2, fmaf(a0, b0, a1 * b1));
}
static inline float xform(const float* m, int i, float x, float y, float z) {
    return dot3(m[i], x, m[i + 4], y, m[i + 8], z) + m[i + 12];
}
static inline float ndc2pix(float v, int S) {
    double d = (double)v + 1.0;
    d = fma(d, (double)S, -1.0);
    d = d * 0.5;
    return (float)d;
}
/* CUDA float->int conversions saturate and map NaN to 0 */
static inline int f2i_sat(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (-2147483647 - 1);
    return (int)f; /* truncation toward zero */
}
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static void tile_rect(float px, float py, float rf, int gx, int gy, int* x0, int* y0, int* x1, int* y1) {
    *x0 = imin(gx, imax(0, f2i_sat((px - rf) * 0.0625f)));
    *y0 = imin(gy, imax(0, f2i_sat((py - rf) * 0.0625f)));
    *x1 = imin(gx, imax(0, f2i_sat((((px + rf) + 16.0f) + -1.0f) * 0.0625f)));
    *y1 = imin(gy, imax(0, f2i_sat((((py + rf) + 16.0f) + -1.0f) * 0.0625f)));
}
static inline float gauss_power(float dx, float dy, float cx, float cy, float cz) {
    float s = fmaf(dx, dx * cx, dy * (dy * cz));
    return fmaf(s, -0.5f, -(dy * (dx * cy)));
}

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

/* ------------------------------------------------------------------ K1 */
static void sh_to_rgb(int deg, const float* sh, float x, float y, float z, float* out) {
    float c[3];
    for (int ch = 0; ch < 3; ch++) c[ch] = sh[ch] * SH_C0;
#define ACC(coef, k)                                                      \
    do {                                                                  \
        float cf_ = (coef);                                               \
        for (int ch = 0; ch < 3; ch++) c[ch] = fmaf(cf_, sh[3 * (k) + ch], c[ch]); \
    } while (0)
    if (deg > 0) {
        ACC(-(y * SH_C1), 1);
        ACC(z * SH_C1, 2);
        ACC(-(x * SH_C1), 3);
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            float zz2 = zz + zz;
            ACC(xy * SH_C2[0], 4);
            ACC(yz * SH_C2[1], 5);
            ACC(((zz2 - xx) - yy) * SH_C2[2], 6);
            ACC(xz * SH_C2[3], 7);
            float xx_yy = xx - yy;
            ACC(xx_yy * SH_C2[4], 8);
            if (deg > 2) {
                float t4 = fmaf(zz, 4.0f, -xx) - yy;
                ACC((y * SH_C3[0]) * fmaf(xx, 3.0f, -yy), 9);
                ACC((xy * SH_C3[1]) * z, 10);
                ACC((y * SH_C3[2]) * t4, 11);
                ACC((z * SH_C3[3]) * fmaf(yy, -3.0f, fmaf(xx, -3.0f, zz2)), 12);
                ACC(t4 * (x * SH_C3[4]), 13);
                ACC(xx_yy * (z * SH_C3[5]), 14);
                ACC((x * SH_C3[6]) * fmaf(yy, -3.0f, xx), 15);
            }
        }
    }
#undef ACC
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2];
}

static void cov3d_fwd(const float* scale, float mod, const float* rot, float* c) {
    const float r = rot[0], x = rot[1], y = rot[2], z = rot[3];
    const float sx = scale[0] * mod, sy = scale[1] * mod, sz = scale[2] * mod;
    const float xz = x * z, rx = r * x, rz = r * z, yy = y * y, zz = z * z;
    const float xz_p_ry = fmaf(r, y, xz), xz_m_ry = fmaf(-r, y, xz);
    const float yz_m_rx = fmaf(y, z, -rx), yz_p_rx = fmaf(y, z, rx);
    const float xy_m_rz = fmaf(x, y, -rz), xy_p_rz = fmaf(x, y, rz);
    const float xx_p_yy = fmaf(x, x, yy), yy_p_zz = yy + zz, xx_p_zz = fmaf(x, x, zz);
    const float R00 = 1.0f - (yy_p_zz + yy_p_zz), R11 = 1.0f - (xx_p_zz + xx_p_zz), R22 = 1.0f - (xx_p_yy + xx_p_yy);
    const float M00 = sx * R00, M01 = sy * (xy_m_rz + xy_m_rz), M02 = sz * (xz_p_ry + xz_p_ry);
    const float M10 = sx * (xy_p_rz + xy_p_rz), M11 = sy * R11, M12 = sz * (yz_m_rx + yz_m_rx);
    const float M20 = sx * (xz_m_ry + xz_m_ry), M21 = sy * (yz_p_rx + yz_p_rx), M22 = sz * R22;
    c[0] = dot3(M00, M00, M01, M01, M02, M02);
    c[1] = dot3(M00, M10, M01, M11, M02, M12);
    c[2] = dot3(M00, M20, M01, M21, M02, M22);
    c[3] = dot3(M10, M10, M11, M11, M12, M12);
    c[4] = dot3(M10, M20, M11, M21, M12, M22);
    c[5] = dot3(M20, M20, M21, M21, M22, M22);
}

void orc_preprocess(int P, int D, int M, const float* means3D, const float* scales, float scale_modifier,
                    const float* rotations, const float* opacities, const float* shs, const float* cov3D_precomp,
                    const float* colors_precomp, const float* V, const float* Pm, const float* cam_pos, int W, int H,
                    float tan_fovx, float tan_fovy, int* radii, float* means2D, float* depths, float* cov3Ds,
                    float* rgb, float* conic_opacity, uint32_t* tiles_touched, uint8_t* clamped) {
    const float focal_y = H / (2.0f * tan_fovy), focal_x = W / (2.0f * tan_fovx);
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; idx++) {
        radii[idx] = 0;
        tiles_touched[idx] = 0;
        depths[idx] = 0.f;
        means2D[2 * idx] = means2D[2 * idx + 1] = 0.f;
        for (int k = 0; k < 4; k++) conic_opacity[4 * idx + k] = 0.f;
        for (int k = 0; k < 3; k++) { rgb[3 * idx + k] = 0.f; clamped[3 * idx + k] = 0; }
        for (int k = 0; k < 6; k++) cov3Ds[6 * idx + k] = 0.f;
        const float x = means3D[3 * idx], y = means3D[3 * idx + 1], z = means3D[3 * idx + 2];
        const float pz = xform(V, 2, x, y, z);
        if (pz <= 0.2f) continue;
        const float hx = xform(Pm, 0, x, y, z), hy = xform(Pm, 1, x, y, z), hw = xform(Pm, 3, x, y, z);
        const float p_w = 1.0f / (hw + 0.0000001f);
        const float projx = hx * p_w, projy = hy * p_w;
        float c[6];
        if (cov3D_precomp) {
            for (int k = 0; k < 6; k++) c[k] = cov3D_precomp[6 * idx + k];
        } else {
            cov3d_fwd(scales + 3 * idx, scale_modifier, rotations + 4 * idx, c);
            for (int k = 0; k < 6; k++) cov3Ds[6 * idx + k] = c[k];
        }
        const float tx = xform(V, 0, x, y, z), ty = xform(V, 1, x, y, z), tz = pz;
        const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
        const float kx = fminf(fmaxf(tx / tz, -limx), limx), ky = fminf(fmaxf(ty / tz, -limy), limy);
        const float tz2 = tz * tz;
        const float J00 = focal_x / tz, J02 = ((tz * -kx) * focal_x) / tz2;
        const float J11 = focal_y / tz, J12 = ((tz * -ky) * focal_y) / tz2;
        const float T00 = fmaf(V[2], J02, V[0] * J00), T01 = fmaf(V[6], J02, V[4] * J00), T02 = fmaf(V[10], J02, V[8] * J00);
        const float T10 = fmaf(V[2], J12, V[1] * J11), T11 = fmaf(V[6], J12, V[5] * J11), T12 = fmaf(V[10], J12, V[9] * J11);
        const float A00 = dot3(T00, c[0], T01, c[1], T02, c[2]);
        const float A10 = dot3(T00, c[1], T01, c[3], T02, c[4]);
        const float A20 = dot3(T00, c[2], T01, c[4], T02, c[5]);
        const float A01 = dot3(T10, c[0], T11, c[1], T12, c[2]);
        const float A11 = dot3(T10, c[1], T11, c[3], T12, c[4]);
        const float A21 = dot3(T10, c[2], T11, c[4], T12, c[5]);
        const float ca = dot3(T00, A00, T01, A10, T02, A20) + 0.3f;
        const float cb = dot3(T00, A01, T01, A11, T02, A21);
        const float cc = dot3(T10, A01, T11, A11, T12, A21) + 0.3f;
        const float det = fmaf(ca, cc, -(cb * cb));
        if (det == 0.0f) continue;
        const float det_inv = 1.0f / det;
        const float conx = cc * det_inv, cony = cb * -det_inv, conz = ca * det_inv;
        const float mid = (ca + cc) * 0.5f;
        const float sq = sqrtf(fmaxf(fmaf(mid, mid, -det), 0.1f));
        const float lam = fmaxf(mid + sq, mid - sq);
        const float r3 = sqrtf(lam) * 3.0f;
        int ri;
        if (r3 != r3) ri = 0; else { float cf = ceilf(r3); ri = f2i_sat(cf); }
        const float rf = (float)ri;
        const float pix_x = ndc2pix(projx, W), pix_y = ndc2pix(projy, H);
        int x0, y0, x1, y1;
        tile_rect(pix_x, pix_y, rf, gx, gy, &x0, &y0, &x1, &y1);
        const uint32_t area = (uint32_t)(x1 - x0) * (uint32_t)(y1 - y0);
        if (area == 0) continue;
        if (!colors_precomp) {
            const float dx = x - cam_pos[0], dy = y - cam_pos[1], dz = z - cam_pos[2];
            const float len = sqrtf(dot3(dx, dx, dy, dy, dz, dz));
            float res[3];
            sh_to_rgb(D, shs + (size_t)idx * M * 3, dx / len, dy / len, dz / len, res);
            for (int ch = 0; ch < 3; ch++) {
                float v = res[ch] + 0.5f;
                clamped[3 * idx + ch] = v < 0.0f;
                rgb[3 * idx + ch] = v < 0.0f ? 0.0f : v;
            }
        } else {
            for (int ch = 0; ch < 3; ch++) rgb[3 * idx + ch] = colors_precomp[3 * idx + ch];
        }
        depths[idx] = pz;
        radii[idx] = ri;
        means2D[2 * idx] = pix_x;
        means2D[2 * idx + 1] = pix_y;
        conic_opacity[4 * idx] = conx; conic_opacity[4 * idx + 1] = cony; conic_opacity[4 * idx + 2] = conz;
        conic_opacity[4 * idx + 3] = opacities[idx];
        tiles_touched[idx] = area;
    }
}

void orc_mark_visible(int P, const float* means3D, const float* V, uint8_t* present) {
    for (int i = 0; i < P; i++)
        present[i] = !(xform(V, 2, means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]) <= 0.2f);
}

/* ------------------------------------------------------------------ K2..K5 */
typedef struct { uint64_t key; uint32_t val; uint32_t seq; } kv_t;
static int kv_cmp(const void* a, const void* b) {
    const kv_t* p = (const kv_t*)a; const kv_t* q = (const kv_t*)b;
    if (p->key != q->key) return p->key < q->key ? -1 : 1;
    return p->seq < q->seq ? -1 : (p->seq > q->seq);  /* stable: emission order */
}

/* Returns R, or -1 on allocation failure.  point_list must hold sum(tiles_touched)
 * entries; ranges holds 2*T u32 (start,end), zero for empty tiles. */
long long orc_binning(int P, const float* means2D, const float* depths, const int* radii, int W, int H,
                      uint32_t* point_list, uint32_t* ranges, uint64_t* keys_out /* may be NULL */) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    size_t R = 0;
    for (int i = 0; i < P; i++) {
        if (radii[i] > 0) {
            int x0, y0, x1, y1;
            tile_rect(means2D[2 * i], means2D[2 * i + 1], (float)radii[i], gx, gy, &x0, &y0, &x1, &y1);
            R += (size_t)(x1 - x0) * (size_t)(y1 - y0);
        }
    }
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)gx * gy);
    if (R == 0) return 0;
    kv_t* kv = (kv_t*)malloc(sizeof(kv_t) * R);
    if (!kv) return -1;
    size_t off = 0;
    for (int i = 0; i < P; i++) {
        if (radii[i] <= 0) continue;
        int x0, y0, x1, y1;
        tile_rect(means2D[2 * i], means2D[2 * i + 1], (float)radii[i], gx, gy, &x0, &y0, &x1, &y1);
        uint32_t dbits;
        memcpy(&dbits, &depths[i], 4);
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < x1; x++) {
                kv[off].key = ((uint64_t)(uint32_t)(y * gx + x) << 32) | dbits;
                kv[off].val = (uint32_t)i;
                kv[off].seq = (uint32_t)off;
                off++;
            }
    }
    qsort(kv, R, sizeof(kv_t), kv_cmp);
    for (size_t i = 0; i < R; i++) {
        point_list[i] = kv[i].val;
        if (keys_out) keys_out[i] = kv[i].key;
        uint32_t cur = (uint32_t)(kv[i].key >> 32);
        if (i == 0) ranges[2 * cur] = 0;
        else {
            uint32_t prev = (uint32_t)(kv[i - 1].key >> 32);
            if (cur != prev) { ranges[2 * prev + 1] = (uint32_t)i; ranges[2 * cur] = (uint32_t)i; }
        }
        if (i == R - 1) ranges[2 * cur + 1] = (uint32_t)R;
    }
    free(kv);
    return (long long)R;
}

/* rasterizer_impl.cu:35-50 */
uint32_t orc_higher_msb(uint32_t n) {
    uint32_t msb = sizeof(n) * 4, step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

/* ------------------------------------------------------------------ K6 */
void orc_render_forward(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* means2D,
                        const float* features, const float* depths, const float* conic_opacity, const float* bg,
                        float* out_color, float* out_depth, float* out_alpha, uint32_t* n_contrib) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const size_t plane = (size_t)W * H;
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < gx * gy; tile++) {
        const int tx = tile % gx, ty = tile / gx;
        const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        for (int ly = 0; ly < TILE; ly++)
            for (int lx = 0; lx < TILE; lx++) {
                const int px = tx * TILE + lx, py = ty * TILE + ly;
                if (px >= W || py >= H) continue;
                const float pxf = (float)px, pyf = (float)py;
                float T = 1.0f, C[3] = {0, 0, 0}, weight = 0.f, Dd = 0.f;
                uint32_t contributor = 0, last = 0;
                for (uint32_t i = r0; i < r1; i++) {
                    contributor++;
                    const uint32_t g = point_list[i];
                    const float dx = means2D[2 * g] - pxf, dy = means2D[2 * g + 1] - pyf;
                    const float* co = conic_opacity + 4 * (size_t)g;
                    const float power = gauss_power(dx, dy, co[0], co[1], co[2]);
                    if (power > 0.0f) continue;
                    const float alpha = fminf(0.99f, co[3] * expf(power));
                    if (alpha < 1.0f / 255.0f) continue;
                    const float test_T = T * (1.0f - alpha);
                    if (test_T < 0.0001f) break; /* done */
                    for (int ch = 0; ch < 3; ch++) C[ch] = fmaf(T, alpha * features[3 * (size_t)g + ch], C[ch]);
                    weight = fmaf(T, alpha, weight);
                    Dd = fmaf(T, alpha * depths[g], Dd);
                    T = test_T;
                    last = contributor;
                }
                const size_t pix = (size_t)py * W + px;
                n_contrib[pix] = last;
                for (int ch = 0; ch < 3; ch++) out_color[ch * plane + pix] = fmaf(bg[ch], T, C[ch]);
                out_alpha[pix] = weight;
                out_depth[pix] = Dd;
            }
    }
}

/* ------------------------------------------------------------------ K7 */
/* Gradients are accumulated in double (acc arrays), then rounded to float. */
void orc_render_backward(int P, int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* bg,
                         const float* means2D, const float* conic_opacity, const float* colors, const float* depths,
                         const float* alphas, const uint32_t* n_contrib, const float* dL_dpix,
                         const float* dL_dpix_depth, const float* dL_dalphas, float* dL_dmean2D /*P*3*/,
                         float* dL_dconic /*P*4*/, float* dL_dopacity /*P*/, float* dL_dcolors /*P*3*/,
                         float* dL_ddepths /*P*/) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const size_t plane = (size_t)W * H;
    double* acc = (double*)calloc((size_t)P * 10, sizeof(double));
    if (!acc) return;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < gx * gy; tile++) {
        const int tx = tile % gx, ty = tile / gx;
        const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        for (int ly = 0; ly < TILE; ly++)
            for (int lx = 0; lx < TILE; lx++) {
                const int px = tx * TILE + lx, py = ty * TILE + ly;
                if (px >= W || py >= H) continue;
                const size_t pix = (size_t)py * W + px;
                const float pxf = (float)px, pyf = (float)py;
                const float T_final = 1.0f - alphas[pix];
                float T = T_final;
                const uint32_t last_contributor = n_contrib[pix];
                uint32_t contributor = r1 - r0;
                float accum_rec[3] = {0, 0, 0}, accum_depth_rec = 0.f, accum_alpha_rec = 0.f;
                float last_alpha = 0.f, last_color[3] = {0, 0, 0}, last_depth = 0.f;
                const float dLdp[3] = {dL_dpix[pix], dL_dpix[plane + pix], dL_dpix[2 * plane + pix]};
                const float dLdD = dL_dpix_depth[pix], dLdA = dL_dalphas[pix];
                for (uint32_t k = r1; k > r0; k--) {
                    contributor--;
                    if (contributor >= last_contributor) continue;
                    const uint32_t g = point_list[k - 1];
                    const float dx = means2D[2 * g] - pxf, dy = means2D[2 * g + 1] - pyf;
                    const float* co = conic_opacity + 4 * (size_t)g;
                    const float power = gauss_power(dx, dy, co[0], co[1], co[2]);
                    if (power > 0.0f) continue;
                    const float G = expf(power);
                    const float alpha = fminf(0.99f, co[3] * G);
                    if (alpha < 1.0f / 255.0f) continue;
                    T = T / (1.f - alpha);
                    const float w = alpha * T;
                    float dL_dopa = 0.0f;
                    double* a = acc + (size_t)g * 10;
                    for (int ch = 0; ch < 3; ch++) {
                        const float c = colors[3 * (size_t)g + ch];
                        accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                        last_color[ch] = c;
                        dL_dopa += (c - accum_rec[ch]) * dLdp[ch];
#pragma omp atomic
                        a[6 + ch] += (double)(w * dLdp[ch]);
                    }
                    const float c_d = depths[g];
                    accum_depth_rec = last_alpha * last_depth + (1.f - last_alpha) * accum_depth_rec;
                    last_depth = c_d;
                    dL_dopa += (c_d - accum_depth_rec) * dLdD;
#pragma omp atomic
                    a[9] += (double)(w * dLdD);
                    accum_alpha_rec = last_alpha + (1.f - last_alpha) * accum_alpha_rec;
                    dL_dopa += (1 - accum_alpha_rec) * dLdA;
                    dL_dopa *= T;
                    last_alpha = alpha;
                    float bg_dot = 0;
                    for (int i = 0; i < 3; i++) bg_dot += bg[i] * dLdp[i];
                    dL_dopa += (-T_final / (1.f - alpha)) * bg_dot;
                    const float dL_dG = co[3] * dL_dopa;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * co[0] - gdy * co[1];
                    const float dG_ddely = -gdy * co[2] - gdx * co[1];
#pragma omp atomic
                    a[0] += (double)(dL_dG * dG_ddelx * ddelx_dx);
#pragma omp atomic
                    a[1] += (double)(dL_dG * dG_ddely * ddely_dy);
#pragma omp atomic
                    a[2] += (double)(-0.5f * gdx * dx * dL_dG);
#pragma omp atomic
                    a[3] += (double)(-0.5f * gdx * dy * dL_dG);
#pragma omp atomic
                    a[4] += (double)(-0.5f * gdy * dy * dL_dG);
#pragma omp atomic
                    a[5] += (double)(G * dL_dopa);
                }
            }
    }
    for (int i = 0; i < P; i++) {
        const double* a = acc + (size_t)i * 10;
        dL_dmean2D[3 * i] = (float)a[0]; dL_dmean2D[3 * i + 1] = (float)a[1]; dL_dmean2D[3 * i + 2] = 0.f;
        dL_dconic[4 * i] = (float)a[2]; dL_dconic[4 * i + 1] = (float)a[3]; dL_dconic[4 * i + 2] = 0.f;
        dL_dconic[4 * i + 3] = (float)a[4];
        dL_dopacity[i] = (float)a[5];
        dL_dcolors[3 * i] = (float)a[6]; dL_dcolors[3 * i + 1] = (float)a[7]; dL_dcolors[3 * i + 2] = (float)a[8];
        dL_ddepths[i] = (float)a[9];
    }
    free(acc);
}

/* ------------------------------------------------------------------ K8 + K9 */
/* glm-style column-major 3x3: m[c][r] */
typedef struct { float m[3][3]; } M3;
static M3 m3_mul(const M3* A, const M3* B) {
    M3 R;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            R.m[i][j] = A->m[0][j] * B->m[i][0] + A->m[1][j] * B->m[i][1] + A->m[2][j] * B->m[i][2];
    return R;
}
static M3 m3_t(const M3* A) {
    M3 R;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R.m[i][j] = A->m[j][i];
    return R;
}

void orc_preprocess_backward(int P, int D, int M, const float* means3D, const int* radii, const float* shs,
                             const uint8_t* clamped /*3P*/, const float* scales, const float* rotations,
                             float scale_modifier, const float* cov3Ds, const float* V, const float* Pm,
                             float focal_x, float focal_y, float tan_fovx, float tan_fovy, const float* campos,
                             const float* dL_dmean2D /*3P*/, const float* dL_dconic /*4P*/, const float* dL_dcolor,
                             const float* dL_ddepth, float* dL_dmeans /*3P*/, float* dL_dcov /*6P*/,
                             float* dL_dsh /*P*M*3*/, float* dL_dscale /*3P*/, float* dL_drot /*4P*/) {
    memset(dL_dmeans, 0, sizeof(float) * 3 * (size_t)P);
    memset(dL_dcov, 0, sizeof(float) * 6 * (size_t)P);
    if (dL_dsh) memset(dL_dsh, 0, sizeof(float) * 3 * (size_t)P * M);
    memset(dL_dscale, 0, sizeof(float) * 3 * (size_t)P);
    memset(dL_drot, 0, sizeof(float) * 4 * (size_t)P);
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; idx++) {
        if (!(radii[idx] > 0)) continue;
        const float* cov3D = cov3Ds + 6 * (size_t)idx;
        const float mx = means3D[3 * idx], my = means3D[3 * idx + 1], mz = means3D[3 * idx + 2];
        /* ---- computeCov2DCUDA */
        const float dconx = dL_dconic[4 * idx], dcony = dL_dconic[4 * idx + 1], dconz = dL_dconic[4 * idx + 3];
        float tx = V[0] * mx + V[4] * my + V[8] * mz + V[12];
        float ty = V[1] * mx + V[5] * my + V[9] * mz + V[13];
        const float tz = V[2] * mx + V[6] * my + V[10] * mz + V[14];
        const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
        const float txtz = tx / tz, tytz = ty / tz;
        tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
        ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
        const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
        const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
        M3 J = {{{focal_x / tz, 0.f, -(focal_x * tx) / (tz * tz)}, {0.f, focal_y / tz, -(focal_y * ty) / (tz * tz)}, {0, 0, 0}}};
        M3 Wm = {{{V[0], V[4], V[8]}, {V[1], V[5], V[9]}, {V[2], V[6], V[10]}}};
        M3 Vrk = {{{cov3D[0], cov3D[1], cov3D[2]}, {cov3D[1], cov3D[3], cov3D[4]}, {cov3D[2], cov3D[4], cov3D[5]}}};
        M3 T = m3_mul(&Wm, &J);
        M3 Tt = m3_t(&T), Vt = m3_t(&Vrk);
        M3 tmp = m3_mul(&Tt, &Vt);
        M3 cov2D = m3_mul(&tmp, &T);
        const float a = cov2D.m[0][0] + 0.3f, b = cov2D.m[0][1], c = cov2D.m[1][1] + 0.3f;
        const float denom = a * c - b * b;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        float* dcv = dL_dcov + 6 * (size_t)idx;
        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dconx + 2 * b * c * dcony + (denom - a * c) * dconz);
            dL_dc = denom2inv * (-a * a * dconz + 2 * a * b * dcony + (denom - a * c) * dconx);
            dL_db = denom2inv * 2 * (b * c * dconx - (denom + 2 * b * b) * dcony + a * b * dconz);
            dcv[0] = (T.m[0][0] * T.m[0][0] * dL_da + T.m[0][0] * T.m[1][0] * dL_db + T.m[1][0] * T.m[1][0] * dL_dc);
            dcv[3] = (T.m[0][1] * T.m[0][1] * dL_da + T.m[0][1] * T.m[1][1] * dL_db + T.m[1][1] * T.m[1][1] * dL_dc);
            dcv[5] = (T.m[0][2] * T.m[0][2] * dL_da + T.m[0][2] * T.m[1][2] * dL_db + T.m[1][2] * T.m[1][2] * dL_dc);
            dcv[1] = 2 * T.m[0][0] * T.m[0][1] * dL_da + (T.m[0][0] * T.m[1][1] + T.m[0][1] * T.m[1][0]) * dL_db + 2 * T.m[1][0] * T.m[1][1] * dL_dc;
            dcv[2] = 2 * T.m[0][0] * T.m[0][2] * dL_da + (T.m[0][0] * T.m[1][2] + T.m[0][2] * T.m[1][0]) * dL_db + 2 * T.m[1][0] * T.m[1][2] * dL_dc;
            dcv[4] = 2 * T.m[0][2] * T.m[0][1] * dL_da + (T.m[0][1] * T.m[1][2] + T.m[0][2] * T.m[1][1]) * dL_db + 2 * T.m[1][1] * T.m[1][2] * dL_dc;
        }
        float dT0[3], dT1[3];
        for (int k = 0; k < 3; k++) {
            const float tv0 = T.m[0][0] * Vrk.m[k][0] + T.m[0][1] * Vrk.m[k][1] + T.m[0][2] * Vrk.m[k][2];
            const float tv1 = T.m[1][0] * Vrk.m[k][0] + T.m[1][1] * Vrk.m[k][1] + T.m[1][2] * Vrk.m[k][2];
            dT0[k] = 2 * tv0 * dL_da + tv1 * dL_db;
            dT1[k] = 2 * tv1 * dL_dc + tv0 * dL_db;
        }
        const float dJ00 = Wm.m[0][0] * dT0[0] + Wm.m[0][1] * dT0[1] + Wm.m[0][2] * dT0[2];
        const float dJ02 = Wm.m[2][0] * dT0[0] + Wm.m[2][1] * dT0[1] + Wm.m[2][2] * dT0[2];
        const float dJ11 = Wm.m[1][0] * dT1[0] + Wm.m[1][1] * dT1[1] + Wm.m[1][2] * dT1[2];
        const float dJ12 = Wm.m[2][0] * dT1[0] + Wm.m[2][1] * dT1[1] + Wm.m[2][2] * dT1[2];
        const float itz = 1.f / tz, itz2 = itz * itz, itz3 = itz2 * itz;
        const float dtx = x_grad_mul * -focal_x * itz2 * dJ02;
        const float dty = y_grad_mul * -focal_y * itz2 * dJ12;
        const float dtz = -focal_x * itz2 * dJ00 - focal_y * itz2 * dJ11 + (2 * focal_x * tx) * itz3 * dJ02 + (2 * focal_y * ty) * itz3 * dJ12;
        float dmean[3];
        dmean[0] = V[0] * dtx + V[1] * dty + V[2] * dtz;
        dmean[1] = V[4] * dtx + V[5] * dty + V[6] * dtz;
        dmean[2] = V[8] * dtx + V[9] * dty + V[10] * dtz;
        /* ---- preprocessCUDA (backward) */
        {
            const float m_hom_w = Pm[3] * mx + Pm[7] * my + Pm[11] * mz + Pm[15];
            const float m_w = 1.0f / (m_hom_w + 0.0000001f);
            const float mul1 = (Pm[0] * mx + Pm[4] * my + Pm[8] * mz + Pm[12]) * m_w * m_w;
            const float mul2 = (Pm[1] * mx + Pm[5] * my + Pm[9] * mz + Pm[13]) * m_w * m_w;
            const float d2x = dL_dmean2D[3 * idx], d2y = dL_dmean2D[3 * idx + 1];
            dmean[0] += (Pm[0] * m_w - Pm[3] * mul1) * d2x + (Pm[1] * m_w - Pm[3] * mul2) * d2y;
            dmean[1] += (Pm[4] * m_w - Pm[7] * mul1) * d2x + (Pm[5] * m_w - Pm[7] * mul2) * d2y;
            dmean[2] += (Pm[8] * m_w - Pm[11] * mul1) * d2x + (Pm[9] * m_w - Pm[11] * mul2) * d2y;
            const float mul3 = V[2] * mx + V[6] * my + V[10] * mz + V[14];
            const float dd = dL_ddepth[idx];
            dmean[0] += (V[2] - V[3] * mul3) * dd;
            dmean[1] += (V[6] - V[7] * mul3) * dd;
            dmean[2] += (V[10] - V[11] * mul3) * dd;
        }
        if (shs) {
            const float* sh = shs + (size_t)idx * M * 3;
            float* dsh = dL_dsh + (size_t)idx * M * 3;
            const float dox = mx - campos[0], doy = my - campos[1], doz = mz - campos[2];
            const float len = sqrtf(dox * dox + doy * doy + doz * doz);
            const float x = dox / len, y = doy / len, z = doz / len;
            float dRGB[3];
            for (int ch = 0; ch < 3; ch++) dRGB[ch] = dL_dcolor[3 * idx + ch] * (clamped[3 * idx + ch] ? 0.f : 1.f);
            float dx_[3] = {0, 0, 0}, dy_[3] = {0, 0, 0}, dz_[3] = {0, 0, 0};
#define SHV(k, ch) sh[3 * (k) + (ch)]
#define DSH(k, coef) do { float cf_ = (coef); for (int ch = 0; ch < 3; ch++) dsh[3 * (k) + ch] = cf_ * dRGB[ch]; } while (0)
            DSH(0, SH_C0);
            if (D > 0) {
                DSH(1, -SH_C1 * y); DSH(2, SH_C1 * z); DSH(3, -SH_C1 * x);
                for (int ch = 0; ch < 3; ch++) { dx_[ch] = -SH_C1 * SHV(3, ch); dy_[ch] = -SH_C1 * SHV(1, ch); dz_[ch] = SH_C1 * SHV(2, ch); }
                if (D > 1) {
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    DSH(4, SH_C2[0] * xy); DSH(5, SH_C2[1] * yz); DSH(6, SH_C2[2] * (2.f * zz - xx - yy));
                    DSH(7, SH_C2[3] * xz); DSH(8, SH_C2[4] * (xx - yy));
                    for (int ch = 0; ch < 3; ch++) {
                        dx_[ch] += SH_C2[0] * y * SHV(4, ch) + SH_C2[2] * 2.f * -x * SHV(6, ch) + SH_C2[3] * z * SHV(7, ch) + SH_C2[4] * 2.f * x * SHV(8, ch);
                        dy_[ch] += SH_C2[0] * x * SHV(4, ch) + SH_C2[1] * z * SHV(5, ch) + SH_C2[2] * 2.f * -y * SHV(6, ch) + SH_C2[4] * 2.f * -y * SHV(8, ch);
                        dz_[ch] += SH_C2[1] * y * SHV(5, ch) + SH_C2[2] * 2.f * 2.f * z * SHV(6, ch) + SH_C2[3] * x * SHV(7, ch);
                    }
                    if (D > 2) {
                        DSH(9, SH_C3[0] * y * (3.f * xx - yy)); DSH(10, SH_C3[1] * xy * z);
                        DSH(11, SH_C3[2] * y * (4.f * zz - xx - yy)); DSH(12, SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
                        DSH(13, SH_C3[4] * x * (4.f * zz - xx - yy)); DSH(14, SH_C3[5] * z * (xx - yy));
                        DSH(15, SH_C3[6] * x * (xx - 3.f * yy));
                        for (int ch = 0; ch < 3; ch++) {
                            dx_[ch] += SH_C3[0] * SHV(9, ch) * 3.f * 2.f * xy + SH_C3[1] * SHV(10, ch) * yz + SH_C3[2] * SHV(11, ch) * -2.f * xy +
                                       SH_C3[3] * SHV(12, ch) * -3.f * 2.f * xz + SH_C3[4] * SHV(13, ch) * (-3.f * xx + 4.f * zz - yy) +
                                       SH_C3[5] * SHV(14, ch) * 2.f * xz + SH_C3[6] * SHV(15, ch) * 3.f * (xx - yy);
                            dy_[ch] += SH_C3[0] * SHV(9, ch) * 3.f * (xx - yy) + SH_C3[1] * SHV(10, ch) * xz +
                                       SH_C3[2] * SHV(11, ch) * (-3.f * yy + 4.f * zz - xx) + SH_C3[3] * SHV(12, ch) * -3.f * 2.f * yz +
                                       SH_C3[4] * SHV(13, ch) * -2.f * xy + SH_C3[5] * SHV(14, ch) * -2.f * yz + SH_C3[6] * SHV(15, ch) * -3.f * 2.f * xy;
                            dz_[ch] += SH_C3[1] * SHV(10, ch) * xy + SH_C3[2] * SHV(11, ch) * 4.f * 2.f * yz +
                                       SH_C3[3] * SHV(12, ch) * 3.f * (2.f * zz - xx - yy) + SH_C3[4] * SHV(13, ch) * 4.f * 2.f * xz +
                                       SH_C3[5] * SHV(14, ch) * (xx - yy);
                        }
                    }
                }
            }
#undef SHV
#undef DSH
            const float ddx = dx_[0] * dRGB[0] + dx_[1] * dRGB[1] + dx_[2] * dRGB[2];
            const float ddy = dy_[0] * dRGB[0] + dy_[1] * dRGB[1] + dy_[2] * dRGB[2];
            const float ddz = dz_[0] * dRGB[0] + dz_[1] * dRGB[1] + dz_[2] * dRGB[2];
            const float sum2 = dox * dox + doy * doy + doz * doz;
            const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
            dmean[0] += ((+sum2 - dox * dox) * ddx - doy * dox * ddy - doz * dox * ddz) * invsum32;
            dmean[1] += (-dox * doy * ddx + (sum2 - doy * doy) * ddy - doz * doy * ddz) * invsum32;
            dmean[2] += (-dox * doz * ddx - doy * doz * ddy + (sum2 - doz * doz) * ddz) * invsum32;
        }
        for (int k = 0; k < 3; k++) dL_dmeans[3 * idx + k] = dmean[k];
        if (scales) {
            const float* rot = rotations + 4 * (size_t)idx;
            const float r = rot[0], x = rot[1], y = rot[2], z = rot[3];
            M3 Rm = {{{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                      {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                      {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}}};
            const float s[3] = {scale_modifier * scales[3 * idx], scale_modifier * scales[3 * idx + 1], scale_modifier * scales[3 * idx + 2]};
            M3 Mm;
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Mm.m[i][j] = s[j] * Rm.m[i][j];
            M3 dS = {{{dcv[0], 0.5f * dcv[1], 0.5f * dcv[2]}, {0.5f * dcv[1], dcv[3], 0.5f * dcv[4]}, {0.5f * dcv[2], 0.5f * dcv[4], dcv[5]}}};
            M3 dM = m3_mul(&Mm, &dS);
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) dM.m[i][j] *= 2.0f;
            M3 Rt = m3_t(&Rm), dMt = m3_t(&dM);
            for (int k = 0; k < 3; k++)
                dL_dscale[3 * idx + k] = Rt.m[k][0] * dMt.m[k][0] + Rt.m[k][1] * dMt.m[k][1] + Rt.m[k][2] * dMt.m[k][2];
            for (int k = 0; k < 3; k++) { dMt.m[0][k] *= s[0]; dMt.m[1][k] *= s[1]; dMt.m[2][k] *= s[2]; }
            float* dq = dL_drot + 4 * (size_t)idx;
            dq[0] = 2 * z * (dMt.m[0][1] - dMt.m[1][0]) + 2 * y * (dMt.m[2][0] - dMt.m[0][2]) + 2 * x * (dMt.m[1][2] - dMt.m[2][1]);
            dq[1] = 2 * y * (dMt.m[1][0] + dMt.m[0][1]) + 2 * z * (dMt.m[2][0] + dMt.m[0][2]) + 2 * r * (dMt.m[1][2] - dMt.m[2][1]) - 4 * x * (dMt.m[2][2] + dMt.m[1][1]);
            dq[2] = 2 * x * (dMt.m[1][0] + dMt.m[0][1]) + 2 * r * (dMt.m[2][0] - dMt.m[0][2]) + 2 * z * (dMt.m[1][2] + dMt.m[2][1]) - 4 * y * (dMt.m[2][2] + dMt.m[0][0]);
            dq[3] = 2 * r * (dMt.m[0][1] - dMt.m[1][0]) + 2 * x * (dMt.m[2][0] + dMt.m[0][2]) + 2 * y * (dMt.m[1][2] + dMt.m[2][1]) - 4 * z * (dMt.m[1][1] + dMt.m[0][0]);
        }
    }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
