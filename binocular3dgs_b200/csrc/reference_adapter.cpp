// reference_adapter.cpp — the reference's own C++ rasterizer interface, implemented on libb3gs.
//
// The reference's torch glue (submodules/diff-gaussian-rasterization/rasterize_points.cu:91,
// :171, :221) calls three static methods declared in cuda_rasterizer/rasterizer.h:20-90 and
// defined in cuda_rasterizer/rasterizer_impl.cu.  This translation unit DEFINES those three
// methods by forwarding to the C-ABI of include/b3gs.h, so a maintainer of the reference
// swaps the CUDA library without touching a line of their sources:
//
//     sources = ["rasterize_points.cu", "ext.cpp",                      # theirs, unmodified
//                "<this repo>/binocular3dgs_b200/csrc/reference_adapter.cpp"]
//     libraries = ["b3gs"]                                               # instead of cuda_rasterizer/*.cu
//
// It is compiled against THE REFERENCE'S header (include path = their checkout), which is
// the point: if their interface and this adapter ever disagree, the build fails.
// baseline/build_adapter.sh does exactly that build; tests/test_gpu_stock_reference.py runs
// the reference's Python package on top of the result and compares it with the stock build.
//
// Streams: the reference enqueues everything on the legacy default stream
// (rasterizer_impl.cu:148,290,315) and its interface has no stream parameter, so the adapter
// passes the legacy default stream too.  Errors: the reference throws std::runtime_error
// (rasterizer_impl.cu:243-246, auxiliary.h:166-173); so does the adapter, with
// b3gs_last_error() as the message.
#include <stdexcept>
#include <string>

#include "cuda_rasterizer/rasterizer.h"   // the reference's declaration of CudaRasterizer::Rasterizer

#include "../../include/b3gs.h"

namespace {

// std::function<char*(size_t)> -> b3gs_buffer {fn, user}
void* call_resize(void* user, size_t bytes) {
    return (*static_cast<std::function<char*(size_t)>*>(user))(bytes);
}
b3gs_buffer as_buffer(std::function<char*(size_t)>& f) { return b3gs_buffer{call_resize, &f}; }

void check(int rc, const char* what) {
    if (rc != B3GS_OK) throw std::runtime_error(std::string(what) + ": " + b3gs_last_error());
}

}  // namespace

void CudaRasterizer::Rasterizer::markVisible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present) {
    check(b3gs_mark_visible(P, means3D, viewmatrix, projmatrix, reinterpret_cast<unsigned char*>(present), nullptr),
          "Rasterizer::markVisible");
}

int CudaRasterizer::Rasterizer::forward(std::function<char*(size_t)> geometryBuffer,
                                        std::function<char*(size_t)> binningBuffer,
                                        std::function<char*(size_t)> imageBuffer, const int P, int D, int M,
                                        const float* background, const int width, int height, const float* means3D,
                                        const float* shs, const float* colors_precomp, const float* opacities,
                                        const float* scales, const float scale_modifier, const float* rotations,
                                        const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                                        const float* cam_pos, const float tan_fovx, float tan_fovy,
                                        const bool prefiltered, float* out_color, float* out_depth, float* out_alpha,
                                        int* radii, bool debug) {
    int rendered = 0;
    check(b3gs_forward(as_buffer(geometryBuffer), as_buffer(binningBuffer), as_buffer(imageBuffer), P, D, M, background,
                       width, height, means3D, shs, colors_precomp, opacities, scales, scale_modifier, rotations,
                       cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy, prefiltered ? 1 : 0,
                       out_color, out_depth, out_alpha, radii, debug ? 1 : 0, nullptr, &rendered),
          "Rasterizer::forward");
    return rendered;
}

void CudaRasterizer::Rasterizer::backward(const int P, int D, int M, int R, const float* background, const int width,
                                          int height, const float* means3D, const float* shs,
                                          const float* colors_precomp, const float* alphas, const float* scales,
                                          const float scale_modifier, const float* rotations,
                                          const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                                          const float* campos, const float tan_fovx, float tan_fovy, const int* radii,
                                          char* geom_buffer, char* binning_buffer, char* image_buffer,
                                          const float* dL_dpix, const float* dL_dpix_depth, const float* dL_dalphas,
                                          float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                                          float* dL_ddepth, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                                          float* dL_dscale, float* dL_drot, bool debug) {
    check(b3gs_backward(P, D, M, R, background, width, height, means3D, shs, colors_precomp, alphas, scales,
                        scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy,
                        radii, geom_buffer, binning_buffer, image_buffer, dL_dpix, dL_dpix_depth, dL_dalphas,
                        dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_ddepth, dL_dmean3D, dL_dcov3D, dL_dsh,
                        dL_dscale, dL_drot, debug ? 1 : 0, nullptr),
          "Rasterizer::backward");
}
