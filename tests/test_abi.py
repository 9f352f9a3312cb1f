"""The C-ABI library loads on a CPU-only box and exports every symbol include/b3gs.h
declares; argument validation works without a GPU (no compute calls)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b3gs.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"B3GS_API\s+[\w\s\*]+?\b(b3gs_\w+)\s*\(", src)))


def test_header_declares_the_reference_entry_points():
    syms = declared_symbols()
    for s in ("b3gs_forward", "b3gs_backward", "b3gs_mark_visible", "b3gs_last_error"):
        assert s in syms
    assert len(syms) >= 12


def test_library_exports_every_declared_symbol():
    from binocular3dgs_b200 import _backend
    lib = ctypes.CDLL(_backend.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/b3gs.h but not exported by libb3gs.so"


def test_layout_queries_are_pure_functions():
    from binocular3dgs_b200 import _backend
    nat = _backend.native()
    assert nat._geometry_bytes(1000) == nat._geometry_bytes(1000)
    assert nat._geometry_bytes(2000) > nat._geometry_bytes(1000)
    assert nat._image_bytes(800, 800) >= 800 * 800 * 4 + 2500 * 8
    assert nat._binning_offset(12345, b"point_list") == 0
    assert nat._geometry_offset(1000, b"no_such_array") == ctypes.c_size_t(-1).value
    offs = [nat._geometry_offset(1000, n) for n in (b"records", b"depths", b"tiles_touched", b"point_offsets", b"clamped")]
    assert offs == sorted(offs) and len(set(offs)) == len(offs)
    assert all(o % 256 == 0 for o in offs)


def test_argument_validation_without_gpu():
    from binocular3dgs_b200 import _backend
    nat = _backend.native()
    g, b, i = nat._buffers()
    r = ctypes.c_int(0)
    fake = 0x1000
    rc = nat._forward(g, b, i, -1, 1, 4, fake, 16, 16, None, None, None, None, None, 1.0, None, None, None, None,
                      None, 0.5, 0.5, 0, fake, fake, fake, None, 0, None, ctypes.byref(r))
    assert rc == -1 and b"bad sizes" in nat._last_error()
    rc = nat._forward(g, b, i, 4, 7, 4, fake, 16, 16, None, None, None, None, None, 1.0, None, None, None, None,
                      None, 0.5, 0.5, 0, fake, fake, fake, None, 0, None, ctypes.byref(r))
    assert rc == -1
    rc = nat._mark_visible(-5, None, None, None, None, None)
    assert rc == -1


def test_missing_library_is_a_loud_import_error(tmp_path):
    from binocular3dgs_b200._backend import Backend
    with pytest.raises(ImportError, match="no CPU fallback"):
        Backend(str(tmp_path / "nope.so"), "b3gs_", False, "b3gs")


def test_cpu_tensor_is_rejected_not_silently_computed():
    import binocular3dgs_b200 as b3
    s = b3.GaussianRasterizationSettings(16, 16, 0.5, 0.5, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0,
                                         torch.zeros(3), False, False)
    r = b3.GaussianRasterizer(s)
    with pytest.raises(RuntimeError, match="CUDA"):
        r(means3D=torch.zeros(4, 3), means2D=torch.zeros(4, 3), opacities=torch.ones(4, 1),
          colors_precomp=torch.ones(4, 3), scales=torch.ones(4, 3), rotations=torch.ones(4, 4))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package (nor the drop-in module
    shims) may import, load or execute it — the product path has no CPU / reference fallback."""
    import glob
    offenders = []
    files = (glob.glob(os.path.join(ROOT, "binocular3dgs_b200", "**", "*.py"), recursive=True)
             + glob.glob(os.path.join(ROOT, "binocular3dgs_b200", "csrc", "*.c*"))
             + glob.glob(os.path.join(ROOT, "diff_gaussian_rasterization", "*.py"))
             + glob.glob(os.path.join(ROOT, "simple_knn", "*.py")))
    assert len(files) > 15
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle|#include\s+\".*oracle)|liboracle|libdgr_ref|libknn_ref|_ref/")
    for f in files:
        for i, line in enumerate(open(f, errors="replace"), 1):
            if pat.search(line) and "oracle/refbackend.py" not in line and not line.lstrip().startswith(("#", "//", "*", '"'))\
                    and "``oracle" not in line:
                offenders.append((os.path.relpath(f, ROOT), i, line.strip()))
    assert not offenders, offenders
