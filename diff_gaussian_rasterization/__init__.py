"""Drop-in module name.  ``gaussian_renderer/__init__.py:14`` of the reference does
``from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer``;
with this repository on ``sys.path`` that import resolves here and the reference's
render adapter runs unmodified on the sm_100a kernels."""
from binocular3dgs_b200 import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    _C,
    _RasterizeGaussians,
    cpu_deep_copy_tuple,
    rasterize_gaussians,
)
