"""binocular3dgs_b200 — B200-native differentiable 3D Gaussian Splatting rasterizer.

Drop-in for the ``diff_gaussian_rasterization`` package that
hanl2010/Binocular3DGS's ``gaussian_renderer/__init__.py:14`` imports: the same four
public names over hand-written sm_100a CUDA kernels reached through the C-ABI of
``include/b3gs.h``.  Importing this package loads ``libb3gs.so``; there is no CPU or
PyTorch fallback, a missing library is an ImportError.
"""
from . import _backend
from .rasterizer import GaussianRasterizationSettings, cpu_deep_copy_tuple, make_surface

_C = _backend.preferred()
_surface = make_surface(_C)
_RasterizeGaussians = _surface._RasterizeGaussians
rasterize_gaussians = _surface.rasterize_gaussians
GaussianRasterizer = _surface.GaussianRasterizer
async_policy = _surface.async_policy      # rasterizer.AsyncCountPolicy: when the forward may skip its host wait

__all__ = [
    "GaussianRasterizationSettings",
    "GaussianRasterizer",
    "rasterize_gaussians",
    "_RasterizeGaussians",
    "_C",
    "make_surface",
]
__version__ = "0.1.0"
