"""distCUDA2 behind the name the reference imports:
``from simple_knn._C import distCUDA2`` (scene/gaussian_model.py:20, used at :134).

``distCUDA2(points)`` takes a float32 CUDA tensor (P,3) and returns the (P,) mean squared
distance of every point to its three nearest neighbours, bit-identical to the reference's
kernel (submodules/simple-knn/simple_knn.cu) through the C-ABI ``b3gs_dist_cuda2``.
CUDA only; there is no CPU path.
"""
from __future__ import annotations

import ctypes

import torch

from . import _backend

_lib = None


def _fns():
    global _lib
    if _lib is None:
        lib = _backend.native().lib
        lib.b3gs_dist_cuda2_scratch_bytes.argtypes = [ctypes.c_int]
        lib.b3gs_dist_cuda2_scratch_bytes.restype = ctypes.c_size_t
        lib.b3gs_dist_cuda2.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_size_t, ctypes.c_void_p]
        lib.b3gs_dist_cuda2.restype = ctypes.c_int
        _lib = lib
    return _lib


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    """spatial.cu:15-25: ``points`` (P,3) float32 CUDA -> ``means`` (P,) float32."""
    if points.dim() != 2 or points.shape[1] != 3:
        raise RuntimeError("points must be (P,3)")
    if not points.is_cuda:
        raise RuntimeError("points must be a CUDA tensor (no CPU path exists)")
    if points.dtype != torch.float32:
        raise RuntimeError("points must be float32")
    pts = points.contiguous()
    P = int(pts.shape[0])
    dev = pts.device
    with torch.cuda.device(dev):
        out = torch.empty(P, dtype=torch.float32, device=dev)
        if P == 0:
            return out
        nbytes = _fns().b3gs_dist_cuda2_scratch_bytes(P)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        rc = _fns().b3gs_dist_cuda2(P, pts.data_ptr(), out.data_ptr(), scratch.data_ptr(), nbytes,
                                    torch.cuda.current_stream(dev).cuda_stream)
    if rc != 0:
        raise RuntimeError(f"b3gs_dist_cuda2 failed ({rc})")
    return out
