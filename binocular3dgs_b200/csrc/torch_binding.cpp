// torch_binding.cpp — the compiled host side above the C-ABI: a pybind11 module with the
// reference's `_C` functions, same names, argument order and return tuples
// (submodules/diff-gaussian-rasterization/ext.cpp:15-18, rasterize_points.h:18-70).
//
// It does what rasterize_points.cu:35-229 does — allocate the outputs and the three opaque
// blobs as torch tensors (the blobs through resize callbacks, rasterize_points.cu:27-33),
// unwrap data pointers, call the library — but the library it calls is libb3gs.so through
// include/b3gs.h, on torch's CURRENT stream.  There is no kernel code in this file.
//
// binocular3dgs_b200/_backend.py implements the same surface with ctypes and stays the
// reference implementation of the host side (and the only one the parity tests use to drive
// the reference's kernels); this module exists because the host side is on the critical
// path: after the forward's one host synchronisation the GPU has ~250 us of work queued, and
// everything the host does before it enqueues the backward in excess of that is GPU idle
// time (tools/host_overhead.py).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <stdexcept>
#include <string>
#include <tuple>

#include "../../include/b3gs.h"

namespace {

void* resize_cb(void* user, size_t n) {  // resizeFunctional, rasterize_points.cu:27-33
    auto* t = static_cast<torch::Tensor*>(user);
    t->resize_({(long long)n});
    return t->data_ptr();
}

// reference null convention (rasterize_points.cu:96-115): empty tensor -> nullptr
const float* fptr(const torch::Tensor& t, const char* name, torch::Tensor& keep) {
    if (!t.defined() || t.numel() == 0) return nullptr;
    TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
    TORCH_CHECK(t.scalar_type() == torch::kFloat32, name, " must be float32");
    keep = t.contiguous();
    if (reinterpret_cast<uintptr_t>(keep.data_ptr()) % 16 != 0) keep = keep.clone();  // the kernels load float4s
    return keep.data_ptr<float>();
}

[[noreturn]] void fail(const char* what, int rc) {
    throw std::runtime_error(std::string("b3gs.") + what + " failed (" + std::to_string(rc) + "): " + b3gs_last_error());
}

// capacity < 0: b3gs_forward (exact, one host wait); >= 0: b3gs_forward_nosync, first element = ticket
std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
forward_common(const int capacity, const torch::Tensor& background, const torch::Tensor& means3D,
               const torch::Tensor& colors, const torch::Tensor& opacity, const torch::Tensor& scales,
               const torch::Tensor& rotations, const float scale_modifier, const torch::Tensor& cov3D_precomp,
               const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix, const float tan_fovx,
               const float tan_fovy, const int image_height, const int image_width, const torch::Tensor& sh,
               const int degree, const torch::Tensor& campos, const bool prefiltered, const bool debug) {
    if (means3D.ndimension() != 2 || means3D.size(1) != 3) AT_ERROR("means3D must have dimensions (num_points, 3)");
    TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor (no CPU path exists)");
    const int P = (int)means3D.size(0), H = image_height, W = image_width;
    const c10::cuda::CUDAGuard guard(means3D.device());
    auto fopt = means3D.options().dtype(torch::kFloat32);
    auto bopt = means3D.options().dtype(torch::kByte);
    torch::Tensor out_color = torch::empty({3, H, W}, fopt), out_depth = torch::empty({1, H, W}, fopt),
                  out_alpha = torch::empty({1, H, W}, fopt),
                  radii = torch::empty({P}, means3D.options().dtype(torch::kInt32));
    torch::Tensor geom = torch::empty({0}, bopt), binning = torch::empty({0}, bopt), img = torch::empty({0}, bopt);
    const int M = (sh.dim() >= 2 && sh.size(0) != 0) ? (int)sh.size(1) : 0;
    torch::Tensor k[11];
    int rendered = 0;
    const float *p_bg = fptr(background, "background", k[0]), *p_m3 = fptr(means3D, "means3D", k[1]),
                *p_sh = fptr(sh, "sh", k[2]), *p_col = fptr(colors, "colors_precomp", k[3]),
                *p_op = fptr(opacity, "opacities", k[4]), *p_sc = fptr(scales, "scales", k[5]),
                *p_rot = fptr(rotations, "rotations", k[6]), *p_cov = fptr(cov3D_precomp, "cov3D_precomp", k[7]),
                *p_vm = fptr(viewmatrix, "viewmatrix", k[8]), *p_pm = fptr(projmatrix, "projmatrix", k[9]),
                *p_cam = fptr(campos, "campos", k[10]);
    const b3gs_buffer bg_{resize_cb, &geom}, bb_{resize_cb, &binning}, bi_{resize_cb, &img};
    void* stream = at::cuda::getCurrentCUDAStream().stream();
    const int rc =
        capacity < 0
            ? b3gs_forward(bg_, bb_, bi_, P, degree, M, p_bg, W, H, p_m3, p_sh, p_col, p_op, p_sc, scale_modifier, p_rot,
                           p_cov, p_vm, p_pm, p_cam, tan_fovx, tan_fovy, prefiltered ? 1 : 0, out_color.data_ptr<float>(),
                           out_depth.data_ptr<float>(), out_alpha.data_ptr<float>(), P ? radii.data_ptr<int>() : nullptr,
                           debug ? 1 : 0, stream, &rendered)
            : b3gs_forward_nosync(bg_, bb_, bi_, P, degree, M, p_bg, W, H, p_m3, p_sh, p_col, p_op, p_sc, scale_modifier,
                                  p_rot, p_cov, p_vm, p_pm, p_cam, tan_fovx, tan_fovy, prefiltered ? 1 : 0,
                                  out_color.data_ptr<float>(), out_depth.data_ptr<float>(), out_alpha.data_ptr<float>(),
                                  P ? radii.data_ptr<int>() : nullptr, stream, capacity, &rendered);
    if (rc != 0) fail("rasterize_gaussians", rc);
    return std::make_tuple(rendered, out_color, out_depth, out_alpha, radii, geom, binning, img);
}

std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
rasterize_gaussians(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& colors,
                    const torch::Tensor& opacity, const torch::Tensor& scales, const torch::Tensor& rotations,
                    const float scale_modifier, const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
                    const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy, const int image_height,
                    const int image_width, const torch::Tensor& sh, const int degree, const torch::Tensor& campos,
                    const bool prefiltered, const bool debug) {
    return forward_common(-1, background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                          viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos,
                          prefiltered, debug);
}

// first element of the result: the ticket for count_wait (include/b3gs.h: b3gs_forward_nosync)
std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
rasterize_gaussians_nosync(const int capacity, const torch::Tensor& background, const torch::Tensor& means3D,
                           const torch::Tensor& colors, const torch::Tensor& opacity, const torch::Tensor& scales,
                           const torch::Tensor& rotations, const float scale_modifier,
                           const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
                           const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
                           const int image_height, const int image_width, const torch::Tensor& sh, const int degree,
                           const torch::Tensor& campos, const bool prefiltered, const bool debug) {
    TORCH_CHECK(capacity >= 1, "capacity must be >= 1");
    return forward_common(capacity, background, means3D, colors, opacity, scales, rotations, scale_modifier,
                          cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh,
                          degree, campos, prefiltered, debug);
}

int count_wait(const int ticket) {
    int r = 0;
    const int rc = b3gs_count_wait(ticket, &r);
    if (rc != 0) fail("count_wait", rc);
    return r;
}

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
           torch::Tensor>
rasterize_gaussians_backward(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii,
                             const torch::Tensor& colors, const torch::Tensor& scales, const torch::Tensor& rotations,
                             const float scale_modifier, const torch::Tensor& cov3D_precomp,
                             const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix, const float tan_fovx,
                             const float tan_fovy, const torch::Tensor& dL_dout_color,
                             const c10::optional<torch::Tensor>& dL_dout_depth,
                             const c10::optional<torch::Tensor>& dL_dout_alpha, const torch::Tensor& sh,
                             const int degree, const torch::Tensor& campos, const torch::Tensor& geomBuffer, const int R,
                             const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer,
                             const torch::Tensor& alphas, const bool debug, const bool skip_unobservable) {
    const int P = (int)means3D.size(0);
    const int H = (int)dL_dout_color.size(1), W = (int)dL_dout_color.size(2);
    const int M = (sh.dim() >= 2 && sh.size(0) != 0) ? (int)sh.size(1) : 0;
    const c10::cuda::CUDAGuard guard(means3D.device());
    auto fopt = means3D.options().dtype(torch::kFloat32);
    // the reference zero-fills all ten (rasterize_points.cu:158-167); the library writes every element
    auto alloc = [&](at::IntArrayRef s) { return P == 0 ? torch::zeros(s, fopt) : torch::empty(s, fopt); };
    // dL_dconic / dL_ddepth never leave rasterize_points.cu:161-162: not materialised.  With
    // skip_unobservable (the autograd surface) neither are dL_dcolors on the SH path and dL_dcov3D
    // on the scale + rotation path, whose gradients autograd discards; they come back undefined (None).
    const bool want_colors = !(skip_unobservable && colors.numel() == 0);
    const bool want_cov3D = !(skip_unobservable && cov3D_precomp.numel() == 0);
    torch::Tensor dL_dmeans3D = alloc({P, 3}), dL_dmeans2D = alloc({P, 3}), dL_dopacity = alloc({P, 1}),
                  dL_dsh = alloc({P, M, 3}), dL_dscales = alloc({P, 3}), dL_drotations = alloc({P, 4});
    torch::Tensor dL_dcolors, dL_dcov3D;
    if (want_colors) dL_dcolors = alloc({P, 3});
    if (want_cov3D) dL_dcov3D = alloc({P, 6});
    if (P != 0) {
        torch::Tensor k[14];
        torch::Tensor none;
        torch::Tensor rad = radii.contiguous();
        const int rc = b3gs_backward(
            P, degree, M, R, fptr(background, "background", k[0]), W, H, fptr(means3D, "means3D", k[1]),
            fptr(sh, "sh", k[2]), fptr(colors, "colors_precomp", k[3]), fptr(alphas, "alphas", k[4]),
            fptr(scales, "scales", k[5]), scale_modifier, fptr(rotations, "rotations", k[6]),
            fptr(cov3D_precomp, "cov3D_precomp", k[7]), fptr(viewmatrix, "viewmatrix", k[8]),
            fptr(projmatrix, "projmatrix", k[9]), fptr(campos, "campos", k[10]), tan_fovx, tan_fovy, rad.data_ptr<int>(),
            reinterpret_cast<char*>(geomBuffer.data_ptr()),
            binningBuffer.numel() ? reinterpret_cast<char*>(binningBuffer.data_ptr()) : nullptr,
            reinterpret_cast<char*>(imageBuffer.data_ptr()), fptr(dL_dout_color, "dL_dout_color", k[11]),
            fptr(dL_dout_depth.has_value() ? *dL_dout_depth : none, "dL_dout_depth", k[12]),
            fptr(dL_dout_alpha.has_value() ? *dL_dout_alpha : none, "dL_dout_alpha", k[13]),
            dL_dmeans2D.data_ptr<float>(), nullptr, dL_dopacity.data_ptr<float>(),
            want_colors ? dL_dcolors.data_ptr<float>() : nullptr, nullptr, dL_dmeans3D.data_ptr<float>(),
            want_cov3D ? dL_dcov3D.data_ptr<float>() : nullptr, M ? dL_dsh.data_ptr<float>() : nullptr, dL_dscales.data_ptr<float>(),
            dL_drotations.data_ptr<float>(), debug ? 1 : 0, at::cuda::getCurrentCUDAStream().stream());
        if (rc != 0) fail("rasterize_gaussians_backward", rc);
    }
    return std::make_tuple(dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales,
                           dL_drotations);
}

torch::Tensor mark_visible(const torch::Tensor& means3D, const torch::Tensor& viewmatrix,
                           const torch::Tensor& projmatrix) {
    const int P = (int)means3D.size(0);
    torch::Tensor present = torch::zeros({P}, means3D.options().dtype(torch::kBool));
    if (P != 0) {
        const c10::cuda::CUDAGuard guard(means3D.device());
        torch::Tensor k[3];
        const int rc = b3gs_mark_visible(P, fptr(means3D, "means3D", k[0]), fptr(viewmatrix, "viewmatrix", k[1]),
                                         fptr(projmatrix, "projmatrix", k[2]),
                                         reinterpret_cast<unsigned char*>(present.data_ptr<bool>()),
                                         at::cuda::getCurrentCUDAStream().stream());
        if (rc != 0) fail("mark_visible", rc);
    }
    return present;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("rasterize_gaussians", &rasterize_gaussians);
    m.def("rasterize_gaussians_backward", &rasterize_gaussians_backward, py::arg("background"), py::arg("means3D"),
          py::arg("radii"), py::arg("colors"), py::arg("scales"), py::arg("rotations"), py::arg("scale_modifier"),
          py::arg("cov3D_precomp"), py::arg("viewmatrix"), py::arg("projmatrix"), py::arg("tan_fovx"),
          py::arg("tan_fovy"), py::arg("dL_dout_color"), py::arg("dL_dout_depth"), py::arg("dL_dout_alpha"),
          py::arg("sh"), py::arg("degree"), py::arg("campos"), py::arg("geomBuffer"), py::arg("R"),
          py::arg("binningBuffer"), py::arg("imageBuffer"), py::arg("alphas"), py::arg("debug"),
          py::arg("skip_unobservable") = false);
    m.def("mark_visible", &mark_visible);
    m.def("rasterize_gaussians_nosync", &rasterize_gaussians_nosync);
    m.def("count_wait", &count_wait, py::call_guard<py::gil_scoped_release>());
    m.def("launch_count", []() { return (unsigned long long)b3gs_launch_count(); });
    m.def("version", []() { return std::string(b3gs_version()); });
}
