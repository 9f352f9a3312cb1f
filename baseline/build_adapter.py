#!/usr/bin/env python
"""baseline/build_adapter.py — the reference's OWN torch glue, unmodified, on top of libb3gs.

Builds the extension a maintainer of the reference would build to swap the CUDA library
(INTEGRATION.md §3):

    sources   rasterize_points.cu + ext.cpp        the reference's, compiled where they lie
              binocular3dgs_b200/csrc/reference_adapter.cpp   defines CudaRasterizer::Rasterizer
                                                   ::forward/backward/markVisible on the C-ABI
    headers   the reference's rasterize_points.h, cuda_rasterizer/rasterizer.h, config.h
    links     binocular3dgs_b200/libb3gs.so        instead of cuda_rasterizer/{forward,backward,rasterizer_impl}.cu

Output (git-ignored, travels to the GPU box): baseline/_ref/adapter/diff_gaussian_rasterization/
with the reference's own __init__.py next to the built _C.so.  No reference file is modified
or added to the repository's history.  No-op when /root/reference is absent (GPU box).
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
DGR = os.path.join(REF, "submodules", "diff-gaussian-rasterization")
OUT = os.path.join(ROOT, "baseline", "_ref", "adapter", "diff_gaussian_rasterization")


def build(verbose=False):
    if not os.path.isdir(DGR):
        print("build_adapter.py: %s not present (GPU box?) — using the prebuilt baseline/_ref/adapter" % DGR)
        return None
    from torch.utils import cpp_extension
    pkg = os.path.join(ROOT, "binocular3dgs_b200")
    adapter = os.path.join(pkg, "csrc", "reference_adapter.cpp")
    srcs = [os.path.join(DGR, "rasterize_points.cu"), os.path.join(DGR, "ext.cpp"), adapter]
    out_so = os.path.join(OUT, "_C.so")
    deps = srcs + [os.path.join(ROOT, "include", "b3gs.h")]
    if os.path.exists(out_so) and all(os.path.getmtime(out_so) >= os.path.getmtime(d) for d in deps):
        return out_so
    bdir = os.path.join(ROOT, "baseline", "_ref", "adapter", "build")
    os.makedirs(bdir, exist_ok=True)
    os.makedirs(OUT, exist_ok=True)
    built = os.path.join(bdir, "_C.so")
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    try:
        cpp_extension.load(
            name="_C", sources=srcs, build_directory=bdir, verbose=verbose, is_python_module=False, with_cuda=True,
            extra_include_paths=[DGR, os.path.join(ROOT, "include")],
            extra_cflags=["-O2", "-std=c++17"], extra_cuda_cflags=["-O2", "-std=c++17"],
            # absolute run path: baseline/_ref/adapter/... -> binocular3dgs_b200/ ($$: ninja, quotes: sh)
            extra_ldflags=["-L" + pkg, "-lb3gs", "-Wl,-rpath,'$$ORIGIN/../../../../binocular3dgs_b200'"])
    except OSError:
        if not os.path.exists(built):     # load() dlopens from the build directory, where $ORIGIN does not resolve
            raise
    shutil.copy2(built, out_so)
    shutil.copy2(os.path.join(DGR, "diff_gaussian_rasterization", "__init__.py"), os.path.join(OUT, "__init__.py"))
    return out_so


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
