"""TEST INFRASTRUCTURE — binds oracle/_ref/libdgr_ref.so (the reference's own CUDA kernels).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` may import this module.
It reuses the product's ctypes host code (``binocular3dgs_b200._backend.Backend``) with
the symbol prefix ``dgr_ref_`` so both libraries are driven by identical Python; the
reference needs zero-filled outputs (rasterize_points.cu:68-71,158-167), ours does not.
"""
import os

from binocular3dgs_b200._backend import Backend

REF_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libdgr_ref.so")
_reference = None


def available() -> bool:
    return os.path.exists(REF_LIB_PATH)


def reference() -> Backend:
    global _reference
    if _reference is None:
        _reference = Backend(REF_LIB_PATH, "dgr_ref_", needs_zeroed_outputs=True, name="dgr_ref")
    return _reference
