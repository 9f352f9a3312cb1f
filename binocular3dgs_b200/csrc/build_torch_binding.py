"""Build binocular3dgs_b200/_b3gs_torch.so — the compiled host side (csrc/torch_binding.cpp)
above the C-ABI — in-tree, so it travels to the GPU box with the snapshot.
Run by __graft_entry__.build() after libb3gs.so exists.  Needs torch headers, g++, ninja;
no nvcc (the file contains no kernel code).  ~1 minute."""
import os
import shutil
import sys


def build(verbose=False):
    from torch.utils import cpp_extension
    here = os.path.dirname(os.path.abspath(__file__))
    pkg = os.path.dirname(here)
    root = os.path.dirname(pkg)
    out = os.path.join(pkg, "_b3gs_torch.so")
    src = os.path.join(here, "torch_binding.cpp")
    deps = [src, os.path.join(root, "include", "b3gs.h"), os.path.join(pkg, "libb3gs.so")]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps[:2]):
        return out
    bdir = os.path.join(here, "build", "torch_binding")
    os.makedirs(bdir, exist_ok=True)
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    built = os.path.join(bdir, "_b3gs_torch.so")
    try:
        cpp_extension.load(
            name="_b3gs_torch", sources=[src], build_directory=bdir, verbose=verbose, is_python_module=False,
            extra_include_paths=[os.path.join(root, "include"), os.path.join(cuda_home, "include")],
            extra_cflags=["-O2", "-std=c++17"], with_cuda=True,
            extra_ldflags=["-L" + pkg, "-lb3gs", "-Wl,-rpath,'$$ORIGIN'"])   # $$: ninja, quotes: sh
    except OSError:
        # load() also tries to dlopen the result from the build directory, where the
        # $ORIGIN rpath cannot find libb3gs.so; the link itself has succeeded by then
        if not os.path.exists(built):
            raise
    shutil.copy2(built, out)
    return out


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
