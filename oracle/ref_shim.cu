/*
 * ref_shim.cu — TEST INFRASTRUCTURE, not product code.
 *
 * A C-ABI veneer over the UNMODIFIED reference rasterizer
 * (/root/reference/submodules/diff-gaussian-rasterization/cuda_rasterizer/
 *  {forward,backward,rasterizer_impl}.cu), which oracle/Makefile compiles where the
 * sources lie and links with this file into oracle/_ref/libdgr_ref.so.  No
 * reference source is copied into this repository.  The veneer exists so the
 * reference's own CUDA kernels can run on the GPU box as "Oracle A": the parity
 * tests compare against it bit for bit, and bench.py --impl reference times it.
 *
 * Entry points mirror include/b3gs.h one for one (same argument order, which is
 * the reference's own, rasterizer.h:24-89) with the prefix dgr_ref_, so the same
 * Python driver code can bind either library.  The reference launches everything
 * on the legacy default stream and ignores `stream`.
 */
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>

#include "rasterizer.h"
#include "rasterizer_impl.h"

typedef void* (*dgr_resize_fn)(void* user, size_t bytes);
struct dgr_buffer { dgr_resize_fn resize; void* user; };

static thread_local std::string g_err;

static std::function<char*(size_t)> wrap(dgr_buffer b) {
    return [b](size_t n) { return reinterpret_cast<char*>(b.resize(b.user, n)); };
}

extern "C" {

int dgr_ref_forward(
    dgr_buffer geometry, dgr_buffer binning, dgr_buffer image,
    int P, int D, int M, const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp,
    const float* opacities, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
    const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy,
    int prefiltered, float* out_color, float* out_depth, float* out_alpha, int* radii,
    int debug, void* /*stream*/, int* num_rendered)
{
    try {
        int r = 0;
        if (P != 0) {
            r = CudaRasterizer::Rasterizer::forward(
                wrap(geometry), wrap(binning), wrap(image), P, D, M, background, width, height,
                means3D, shs, colors_precomp, opacities, scales, scale_modifier, rotations,
                cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy,
                prefiltered != 0, out_color, out_depth, out_alpha, radii, debug != 0);
        }
        if (num_rendered) *num_rendered = r;
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -2;
    }
}

int dgr_ref_backward(
    int P, int D, int M, int R, const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp, const float* alphas,
    const float* scales, float scale_modifier, const float* rotations,
    const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
    const float* campos, float tan_fovx, float tan_fovy, const int* radii,
    char* geom_buffer, char* binning_buffer, char* image_buffer,
    const float* dL_dpix, const float* dL_dpix_depth, const float* dL_dalphas,
    float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
    float* dL_ddepth, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
    float* dL_dscale, float* dL_drot, int debug, void* /*stream*/)
{
    try {
        if (P != 0) {
            CudaRasterizer::Rasterizer::backward(
                P, D, M, R, background, width, height, means3D, shs, colors_precomp, alphas,
                scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos,
                tan_fovx, tan_fovy, radii, geom_buffer, binning_buffer, image_buffer,
                dL_dpix, dL_dpix_depth, dL_dalphas, dL_dmean2D, dL_dconic, dL_dopacity,
                dL_dcolor, dL_ddepth, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot,
                debug != 0);
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -2;
    }
}

int dgr_ref_mark_visible(int P, const float* means3D, const float* viewmatrix,
                         const float* projmatrix, unsigned char* present, void* /*stream*/)
{
    if (P != 0) {
        CudaRasterizer::Rasterizer::markVisible(
            P, const_cast<float*>(means3D), const_cast<float*>(viewmatrix),
            const_cast<float*>(projmatrix), reinterpret_cast<bool*>(present));
    }
    return 0;
}

size_t dgr_ref_geometry_bytes(int P) {
    return CudaRasterizer::required<CudaRasterizer::GeometryState>(P);
}
size_t dgr_ref_binning_bytes(int R) {
    return CudaRasterizer::required<CudaRasterizer::BinningState>(R);
}
size_t dgr_ref_image_bytes(int width, int height) {
    return CudaRasterizer::required<CudaRasterizer::ImageState>((size_t)width * height);
}

/* Offsets of the reference's internal arrays inside its blobs, derived by running
 * its own fromChunk on a base the caller's allocation would have (any 128-B
 * aligned base gives the same offsets). */
size_t dgr_ref_geometry_offset(int P, const char* name) {
    char* base = reinterpret_cast<char*>(uintptr_t(1) << 20);
    char* p = base;
    auto g = CudaRasterizer::GeometryState::fromChunk(p, P);
    auto off = [&](const void* q) { return size_t(reinterpret_cast<const char*>(q) - base); };
    if (!strcmp(name, "depths")) return off(g.depths);
    if (!strcmp(name, "clamped")) return off(g.clamped);
    if (!strcmp(name, "means2D")) return off(g.means2D);
    if (!strcmp(name, "cov3D")) return off(g.cov3D);
    if (!strcmp(name, "conic_opacity")) return off(g.conic_opacity);
    if (!strcmp(name, "rgb")) return off(g.rgb);
    if (!strcmp(name, "tiles_touched")) return off(g.tiles_touched);
    if (!strcmp(name, "point_offsets")) return off(g.point_offsets);
    return (size_t)-1;
}
size_t dgr_ref_binning_offset(int R, const char* name) {
    char* base = reinterpret_cast<char*>(uintptr_t(1) << 20);
    char* p = base;
    auto b = CudaRasterizer::BinningState::fromChunk(p, R);
    auto off = [&](const void* q) { return size_t(reinterpret_cast<const char*>(q) - base); };
    if (!strcmp(name, "point_list")) return off(b.point_list);
    if (!strcmp(name, "point_list_keys")) return off(b.point_list_keys);
    if (!strcmp(name, "point_list_unsorted")) return off(b.point_list_unsorted);
    if (!strcmp(name, "point_list_keys_unsorted")) return off(b.point_list_keys_unsorted);
    return (size_t)-1;
}
size_t dgr_ref_image_offset(int width, int height, const char* name) {
    char* base = reinterpret_cast<char*>(uintptr_t(1) << 20);
    char* p = base;
    auto im = CudaRasterizer::ImageState::fromChunk(p, (size_t)width * height);
    auto off = [&](const void* q) { return size_t(reinterpret_cast<const char*>(q) - base); };
    if (!strcmp(name, "n_contrib")) return off(im.n_contrib);
    if (!strcmp(name, "ranges")) return off(im.ranges);
    return (size_t)-1;
}

const char* dgr_ref_last_error(void) { return g_err.c_str(); }
const char* dgr_ref_version(void) { return "diff-gaussian-rasterization@8829d14 (reference kernels, unmodified)"; }
unsigned long long dgr_ref_launch_count(void) { return 0; }

}  // extern "C"
