"""The ``diff_gaussian_rasterization`` operator surface, bound to a rasterizer backend.

Mirrors, name for name and argument for argument, the Python API the reference's
render adapter imports (``gaussian_renderer/__init__.py:14,36-51,85-93``) from
``submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py``:

* ``GaussianRasterizationSettings``  (…/__init__.py:160-172) — NamedTuple, 12 fields
* ``GaussianRasterizer``             (:174-223) — nn.Module, ``forward`` / ``markVisible``
* ``rasterize_gaussians``            (:21-42)
* ``_RasterizeGaussians``            (:44-158) — autograd.Function, 9 inputs, 4 outputs

``make_surface(backend)`` builds these four objects on top of any object exposing the
reference's ``_C`` functions, so the test-suite can instantiate the identical surface
over the reference's own kernels for A/B parity.  The package-level names are the
surface bound to our sm_100a library.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import NamedTuple

import torch
import torch.nn as nn


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def cpu_deep_copy_tuple(input_tuple):
    """Host snapshot of an argument tuple (…/__init__.py:17-19), for debug dumps."""
    return tuple(item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple)


def make_surface(_C) -> SimpleNamespace:
    """Build the operator surface over a ``_C``-like backend."""

    class _RasterizeGaussians(torch.autograd.Function):
        @staticmethod
        def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                    raster_settings):
            s = raster_settings
            # argument order of the C++ entry point (rasterize_points.h:18-38)
            args = (
                s.bg, means3D, colors_precomp, opacities, scales, rotations, s.scale_modifier, cov3Ds_precomp,
                s.viewmatrix, s.projmatrix, s.tanfovx, s.tanfovy, s.image_height, s.image_width, sh, s.sh_degree,
                s.campos, s.prefiltered, s.debug,
            )
            if s.debug:
                # replayable input snapshot on failure (…/__init__.py:83-90)
                cpu_args = cpu_deep_copy_tuple(args)
                try:
                    out = _C.rasterize_gaussians(*args)
                except Exception as ex:
                    torch.save(cpu_args, "snapshot_fw.dump")
                    print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                    raise ex
            else:
                out = _C.rasterize_gaussians(*args)
            num_rendered, color, depth, alpha, radii, geomBuffer, binningBuffer, imgBuffer = out

            ctx.raster_settings = s
            ctx.num_rendered = num_rendered
            ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh,
                                  geomBuffer, binningBuffer, imgBuffer, alpha)
            ctx.mark_non_differentiable(radii)
            if getattr(_C, "accepts_null_grads", False):
                # outputs that received no gradient arrive as None instead of materialised
                # zero images (the C-ABI takes NULL for dL/ddepth and dL/dalpha)
                ctx.set_materialize_grads(False)
            return color, radii, depth, alpha

        @staticmethod
        def backward(ctx, grad_color, grad_radii, grad_depth, grad_alpha):
            s = ctx.raster_settings
            (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer,
             imgBuffer, alpha) = ctx.saved_tensors
            if grad_color is None:      # only reachable with set_materialize_grads(False)
                grad_color = torch.zeros((3, s.image_height, s.image_width), dtype=alpha.dtype, device=alpha.device)
                if grad_depth is None and grad_alpha is None:
                    grad_depth = torch.zeros_like(alpha)
            # argument order of the C++ entry point (rasterize_points.h:40-65)
            args = (
                s.bg, means3D, radii, colors_precomp, scales, rotations, s.scale_modifier, cov3Ds_precomp,
                s.viewmatrix, s.projmatrix, s.tanfovx, s.tanfovy, grad_color, grad_depth, grad_alpha, sh,
                s.sh_degree, s.campos, geomBuffer, ctx.num_rendered, binningBuffer, imgBuffer, alpha, s.debug,
            )
            # a gradient sink (dp.GradientBucket) hands its views to ONE backward per step when
            # autograd is the caller: autograd sums the results of several backwards itself
            tracks = hasattr(_C, "in_autograd")
            if tracks:
                _C.in_autograd = True
            try:
                if s.debug:
                    cpu_args = cpu_deep_copy_tuple(args)
                    try:
                        grads = _C.rasterize_gaussians_backward(*args)
                    except Exception as ex:
                        torch.save(cpu_args, "snapshot_bw.dump")
                        print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                        raise ex
                else:
                    grads = _C.rasterize_gaussians_backward(*args)
            finally:
                if tracks:
                    _C.in_autograd = False
            (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh,
             grad_scales, grad_rotations) = grads
            # one gradient per forward input (…/__init__.py:146-158)
            return (grad_means3D, grad_means2D, grad_sh, grad_colors_precomp, grad_opacities, grad_scales,
                    grad_rotations, grad_cov3Ds_precomp, None)

    def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                            raster_settings):
        return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                         cov3Ds_precomp, raster_settings)

    class GaussianRasterizer(nn.Module):
        def __init__(self, raster_settings):
            super().__init__()
            self.raster_settings = raster_settings

        def markVisible(self, positions):
            # boolean mask of points passing the near-plane test (…/__init__.py:179-188)
            with torch.no_grad():
                s = self.raster_settings
                return _C.mark_visible(positions, s.viewmatrix, s.projmatrix)

        def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                    cov3D_precomp=None):
            # same validation and messages as the reference (…/__init__.py:194-198)
            if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
                raise Exception('Please provide excatly one of either SHs or precomputed colors!')
            if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                    (scales is not None or rotations is not None) and cov3D_precomp is not None):
                raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
            # absent optionals travel as empty tensors == null pointers (:200-210)
            empty = torch.Tensor([])
            shs = empty if shs is None else shs
            colors_precomp = empty if colors_precomp is None else colors_precomp
            scales = empty if scales is None else scales
            rotations = empty if rotations is None else rotations
            cov3D_precomp = empty if cov3D_precomp is None else cov3D_precomp
            return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                       cov3D_precomp, self.raster_settings)

    return SimpleNamespace(
        _C=_C,
        _RasterizeGaussians=_RasterizeGaussians,
        rasterize_gaussians=rasterize_gaussians,
        GaussianRasterizer=GaussianRasterizer,
        GaussianRasterizationSettings=GaussianRasterizationSettings,
    )
