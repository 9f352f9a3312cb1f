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

import os
import warnings
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


class AsyncCountPolicy:
    """When may the forward skip its one host wait (include/b3gs.h: b3gs_forward_nosync)?

    The reference reads ``num_rendered`` back inside every forward (rasterizer_impl.cu:282)
    because the binning buffer is sized from it.  Here the first forward of a given
    (device, P, width, height) takes the exact path and records R; later TRAINING forwards
    size the buffer from ``margin x`` the largest R seen (plus a floor), do not wait, and hand
    autograd a ticket instead of the integer.  The backward resolves the ticket (by then the
    count has long landed) and, should R have exceeded the buffer, re-renders exactly into the
    same output tensors before computing gradients, and warns: the loss of that one step was
    formed from a truncated image.  ``B3GS_ASYNC=0`` (or ``policy.enabled = False``) keeps every
    forward exact; forwards without gradient (evaluation) and ``debug=True`` always are.
    """

    def __init__(self):
        self.enabled = os.environ.get("B3GS_ASYNC", "1") != "0"
        self.margin, self.floor = 1.5, 1 << 16
        self.hwm = {}
        self.overflows = 0
        self.forced_capacity = None        # tests: provoke an overflow

    def capacity(self, key):
        if not self.enabled:
            return None
        if self.forced_capacity is not None:
            return int(self.forced_capacity)
        h = self.hwm.get(key)
        return None if h is None else min(int(h * self.margin) + self.floor, 2 ** 31 - 1)

    def observe(self, key, r):
        if r > self.hwm.get(key, -1):
            self.hwm[key] = int(r)


class _PendingCount:
    """``ctx.num_rendered`` of a no-sync forward: a ticket, and what a re-render needs."""
    __slots__ = ("ticket", "capacity", "key", "args", "outputs")

    def __init__(self, ticket, capacity, key, args, outputs):
        self.ticket, self.capacity, self.key, self.args, self.outputs = ticket, capacity, key, args, outputs


def make_surface(_C) -> SimpleNamespace:
    """Build the operator surface over a ``_C``-like backend."""
    policy = AsyncCountPolicy() if hasattr(_C, "rasterize_gaussians_nosync") else None

    def resolve(pending):
        """-> (R, replacement blobs or None).  Waits for the count of a no-sync forward."""
        r = _C.count_wait(pending.ticket)
        policy.observe(pending.key, r)
        if r <= pending.capacity:
            return r, None
        policy.overflows += 1
        warnings.warn("binocular3dgs_b200: %d tile instances exceeded the binning buffer sized for %d by the "
                      "no-sync forward; re-rendering exactly (this step's loss saw a truncated image). "
                      "Set B3GS_ASYNC=0 to make every forward exact." % (r, pending.capacity), RuntimeWarning)
        R, color, depth, alpha, radii, geom, binning, img = _C.rasterize_gaussians(*pending.args)
        for old, new in zip(pending.outputs, (color, depth, alpha, radii)):
            old.copy_(new)                  # aliases of the storage the caller (and autograd) holds
        return R, (geom, binning, img)

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
            pending = None
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
                capacity = key = None
                if policy is not None and means3D.is_cuda and any(ctx.needs_input_grad):
                    P = int(means3D.shape[0])
                    key = (means3D.device.index, P, int(s.image_width), int(s.image_height))
                    if P > 0 and _C.nosync_supported(P, s.image_width, s.image_height):
                        capacity = policy.capacity(key)
                if capacity is not None:
                    out = _C.rasterize_gaussians_nosync(capacity, *args)
                    # detached aliases: the returned tensors themselves will point at this node
                    # (grad_fn), and holding them here would be a reference cycle that only the
                    # garbage collector breaks — a render's worth of device memory leaked per step
                    pending = _PendingCount(out[0], capacity, key, args, tuple(t.detach() for t in out[1:5]))
                else:
                    out = _C.rasterize_gaussians(*args)
                    if key is not None:
                        policy.observe(key, out[0])
            num_rendered, color, depth, alpha, radii, geomBuffer, binningBuffer, imgBuffer = out
            if pending is not None:
                num_rendered = pending

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
            num_rendered = ctx.num_rendered
            if isinstance(num_rendered, _PendingCount):
                num_rendered, fresh = resolve(num_rendered)
                if fresh is not None:
                    geomBuffer, binningBuffer, imgBuffer = fresh
            if grad_color is None:      # only reachable with set_materialize_grads(False)
                grad_color = torch.zeros((3, s.image_height, s.image_width), dtype=alpha.dtype, device=alpha.device)
                if grad_depth is None and grad_alpha is None:
                    grad_depth = torch.zeros_like(alpha)
            # argument order of the C++ entry point (rasterize_points.h:40-65)
            args = (
                s.bg, means3D, radii, colors_precomp, scales, rotations, s.scale_modifier, cov3Ds_precomp,
                s.viewmatrix, s.projmatrix, s.tanfovx, s.tanfovy, grad_color, grad_depth, grad_alpha, sh,
                s.sh_degree, s.campos, geomBuffer, num_rendered, binningBuffer, imgBuffer, alpha, s.debug,
            )
            # a gradient sink (dp.GradientBucket) hands its views to ONE backward per step when
            # autograd is the caller: autograd sums the results of several backwards itself
            tracks = hasattr(_C, "in_autograd")
            if tracks:
                _C.in_autograd = True
            # autograd discards the gradients of inputs that were empty placeholders: no need to
            # materialise dL_dcolors (SH path) / dL_dcov3D (scale + rotation path)
            kw = {"skip_unobservable": True} if getattr(_C, "supports_skip_unobservable", False) else {}
            try:
                if s.debug:
                    cpu_args = cpu_deep_copy_tuple(args)
                    try:
                        grads = _C.rasterize_gaussians_backward(*args, **kw)
                    except Exception as ex:
                        torch.save(cpu_args, "snapshot_bw.dump")
                        print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                        raise ex
                else:
                    grads = _C.rasterize_gaussians_backward(*args, **kw)
            finally:
                if tracks:
                    _C.in_autograd = False
            (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh,
             grad_scales, grad_rotations) = grads
            # one gradient per forward input (…/__init__.py:146-158)
            return (grad_means3D, grad_means2D, grad_sh, grad_colors_precomp, grad_opacities, grad_scales,
                    grad_rotations, grad_cov3Ds_precomp, None)

    class _RasterizeGaussiansRaw(torch.autograd.Function):
        """The raw-parameter entry (include/b3gs.h: b3gs_forward_raw / b3gs_backward_raw): what
        render() + GaussianModel compute as  rasterize(xyz, cat(f_dc, f_rest), sigmoid(opacity),
        exp(scaling), normalize(rotation))  (gaussian_renderer/__init__.py:54-93,
        scene/gaussian_model.py:95-115) in ONE operator whose gradients are those of the raw
        parameters."""

        @staticmethod
        def forward(ctx, xyz, means2D, f_dc, f_rest, opacity, scaling, rotation, raster_settings):
            s = raster_settings
            args = (s.bg, xyz, f_dc, f_rest, opacity, scaling, rotation, s.scale_modifier, s.viewmatrix, s.projmatrix,
                    s.tanfovx, s.tanfovy, s.image_height, s.image_width, s.sh_degree, s.campos)
            capacity = key = pending = None
            if policy is not None and any(ctx.needs_input_grad) and not s.debug:
                P = int(xyz.shape[0])
                key = (xyz.device.index, P, int(s.image_width), int(s.image_height))
                if P > 0 and _C.nosync_supported(P, s.image_width, s.image_height):
                    capacity = policy.capacity(key)
            out = _C.rasterize_raw(*args, capacity=capacity)
            num_rendered, color, depth, alpha, radii, geomBuffer, binningBuffer, imgBuffer = out
            if capacity is not None:
                num_rendered = _PendingCount(out[0], capacity, key, args, tuple(t.detach() for t in out[1:5]))
            elif key is not None:
                policy.observe(key, out[0])
            ctx.raster_settings, ctx.num_rendered = s, num_rendered
            ctx.save_for_backward(xyz, f_dc, f_rest, opacity, scaling, rotation, radii, geomBuffer, binningBuffer,
                                  imgBuffer, alpha)
            ctx.mark_non_differentiable(radii)
            ctx.set_materialize_grads(False)
            return color, radii, depth, alpha

        @staticmethod
        def backward(ctx, grad_color, grad_radii, grad_depth, grad_alpha):
            s = ctx.raster_settings
            xyz, f_dc, f_rest, opacity, scaling, rotation, radii, geomBuffer, binningBuffer, imgBuffer, alpha = \
                ctx.saved_tensors
            num_rendered = ctx.num_rendered
            if isinstance(num_rendered, _PendingCount):
                pending = num_rendered
                num_rendered = _C.count_wait(pending.ticket)
                policy.observe(pending.key, num_rendered)
                if num_rendered > pending.capacity:      # as resolve(), through the raw entry
                    policy.overflows += 1
                    warnings.warn("binocular3dgs_b200: %d tile instances exceeded the binning buffer sized for %d by "
                                  "the no-sync forward; re-rendering exactly (this step's loss saw a truncated image)."
                                  % (num_rendered, pending.capacity), RuntimeWarning)
                    num_rendered, c2, d2, a2, r2, geomBuffer, binningBuffer, imgBuffer = _C.rasterize_raw(*pending.args)
                    for old, new in zip(pending.outputs, (c2, d2, a2, r2)):
                        old.copy_(new)
            if grad_color is None:
                grad_color = torch.zeros((3, s.image_height, s.image_width), dtype=alpha.dtype, device=alpha.device)
                if grad_depth is None and grad_alpha is None:
                    grad_depth = torch.zeros_like(alpha)
            g2d, g_xyz, g_dc, g_rest, g_op, g_sc, g_rot = _C.rasterize_raw_backward(
                s.bg, xyz, f_dc, f_rest, opacity, scaling, rotation, s.scale_modifier, s.viewmatrix, s.projmatrix,
                s.tanfovx, s.tanfovy, grad_color, grad_depth, grad_alpha, s.sh_degree, s.campos, radii, geomBuffer,
                num_rendered, binningBuffer, imgBuffer, alpha)
            return g_xyz, g2d, g_dc, g_rest, g_op, g_sc, g_rot, None

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

        def forward_raw(self, xyz, means2D, f_dc, f_rest, opacity, scaling, rotation):
            """Render straight from the reference's RAW parameters (GaussianModel._xyz,
            ._features_dc, ._features_rest, ._opacity, ._scaling, ._rotation): the activations of
            scene/gaussian_model.py:95-115 and their gradients are fused into the operator.
            An addition to the reference surface; returns (color, radii, depth, alpha)."""
            if not hasattr(_C, "rasterize_raw"):
                raise RuntimeError("this backend has no raw-parameter entry")
            return _RasterizeGaussiansRaw.apply(xyz, means2D, f_dc, f_rest, opacity, scaling, rotation,
                                                self.raster_settings)

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
        async_policy=policy,
        _RasterizeGaussians=_RasterizeGaussians,
        rasterize_gaussians=rasterize_gaussians,
        GaussianRasterizer=GaussianRasterizer,
        GaussianRasterizationSettings=GaussianRasterizationSettings,
    )
