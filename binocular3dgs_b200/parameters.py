"""Per-step parameter plumbing behind the reference's own names.

* ``activate`` — what ``GaussianModel.get_features / get_opacity / get_scaling /
  get_rotation`` (scene/gaussian_model.py:95-115) compute for one render, with autograd,
  as one forward and one backward kernel instead of ~7 + ~10 elementwise launches.
* ``FusedAdam`` — a ``torch.optim.Adam`` subclass (same constructor, same ``param_groups``
  and per-parameter ``state`` layout: ``step``, ``exp_avg``, ``exp_avg_sq``), so the
  reference's optimizer surgery during densification
  (``cat_tensors_to_optimizer`` / ``_prune_optimizer`` / ``replace_tensor_to_optimizer``,
  gaussian_model.py:255-343) keeps working; ``step()`` is ONE kernel for all groups.
* ``opacity_decay`` — gaussian_model.py:307-309 in place.
* ``add_densification_stats`` — train.py:170-171 + gaussian_model.py:409-411 without
  boolean-mask indexing (no ``nonzero``, no host sync).

CUDA float32 only; no CPU path.
"""
from __future__ import annotations

import ctypes
import math

import torch

from . import _backend

_V, _I, _F, _D = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double
ADAM_MAX_TENSORS = 8


class _AdamTensor(ctypes.Structure):
    """B3gsAdamTensor (include/b3gs.h)."""
    _fields_ = [("param", _V), ("grad", _V), ("exp_avg", _V), ("exp_avg_sq", _V), ("n", ctypes.c_size_t),
                ("step_size", _F), ("inv_bias_correction2_sqrt", _F)]


_lib = None


def _fns():
    global _lib
    if _lib is None:
        lib = _backend.native().lib
        for name, args in (("b3gs_activate_forward", [_I, _I] + [_V] * 10),
                           ("b3gs_activate_backward", [_I, _I] + [_V] * 13),
                           ("b3gs_adam_multi", [_I, ctypes.POINTER(_AdamTensor), _D, _D, _D, _V]),
                           ("b3gs_opacity_decay", [_I, _F, _V, _V]),
                           ("b3gs_densify_stats", [_I, _V, _V, _V, _V, _V, _V])):
            fn = getattr(lib, name)
            fn.argtypes, fn.restype = args, _I
        _lib = lib
    return _lib


def _call(fn, *args):
    rc = fn(*args)
    if rc != 0:
        raise RuntimeError(f"{fn.__name__} failed ({rc})")


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _f32_cuda(t: torch.Tensor, name: str, shape=None) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (no CPU path exists)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise RuntimeError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
    t = t.contiguous()
    return t if t.data_ptr() % 16 == 0 else t.clone()      # the kernels load float4s


# ------------------------------------------------------------------ activations
class _Activate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f_dc, f_rest, opacity_raw, scaling_raw, rotation_raw):
        P, M = int(f_dc.shape[0]), 1 + int(f_rest.shape[1])
        dev = f_dc.device
        with torch.cuda.device(dev):
            shs = torch.empty((P, M, 3), dtype=torch.float32, device=dev)
            opacities = torch.empty((P, 1), dtype=torch.float32, device=dev)
            scales = torch.empty((P, 3), dtype=torch.float32, device=dev)
            rotations = torch.empty((P, 4), dtype=torch.float32, device=dev)
            _call(_fns().b3gs_activate_forward, P, M, f_dc.data_ptr(), f_rest.data_ptr() if M > 1 else None,
                  opacity_raw.data_ptr(), scaling_raw.data_ptr(), rotation_raw.data_ptr(), shs.data_ptr(),
                  opacities.data_ptr(), scales.data_ptr(), rotations.data_ptr(), _stream(dev))
        ctx.save_for_backward(opacity_raw, scaling_raw, rotation_raw)
        ctx.meta = (P, M)
        return shs, opacities, scales, rotations

    @staticmethod
    def backward(ctx, g_shs, g_opacities, g_scales, g_rotations):
        opacity_raw, scaling_raw, rotation_raw = ctx.saved_tensors
        P, M = ctx.meta
        dev = opacity_raw.device
        need = ctx.needs_input_grad

        def grad_in(g, wanted):
            if g is None or not wanted:
                return None
            g = g.contiguous()
            return g if g.data_ptr() % 16 == 0 else g.clone()

        g_shs = grad_in(g_shs, need[0] or need[1])
        g_opacities, g_scales, g_rotations = grad_in(g_opacities, need[2]), grad_in(g_scales, need[3]), grad_in(g_rotations, need[4])
        with torch.cuda.device(dev):
            new = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
            o_dc = new(P, 1, 3) if g_shs is not None else None
            o_rest = new(P, M - 1, 3) if g_shs is not None else None
            o_op = new(P, 1) if g_opacities is not None else None
            o_sc = new(P, 3) if g_scales is not None else None
            o_rot = new(P, 4) if g_rotations is not None else None
            p = lambda t: None if (t is None or t.numel() == 0) else t.data_ptr()
            _call(_fns().b3gs_activate_backward, P, M, opacity_raw.data_ptr(), scaling_raw.data_ptr(),
                  rotation_raw.data_ptr(), p(g_shs), p(g_opacities), p(g_scales), p(g_rotations), p(o_dc), p(o_rest),
                  p(o_op), p(o_sc), p(o_rot), _stream(dev))
        return o_dc, o_rest, o_op, o_sc, o_rot


def activate(features_dc, features_rest, opacity, scaling, rotation):
    """(get_features, get_opacity, get_scaling, get_rotation) of scene/gaussian_model.py:95-115
    from the raw parameters ``_features_dc (P,1,3)``, ``_features_rest (P,M-1,3)``,
    ``_opacity (P,1)``, ``_scaling (P,3)``, ``_rotation (P,4)``."""
    P = int(features_dc.shape[0])
    if features_dc.dim() != 3 or tuple(features_dc.shape[1:]) != (1, 3):
        raise RuntimeError("features_dc must be (P,1,3)")
    if features_rest.dim() != 3 or features_rest.shape[0] != P or features_rest.shape[2] != 3:
        raise RuntimeError("features_rest must be (P,M-1,3)")
    return _Activate.apply(_f32_cuda(features_dc, "features_dc"), _f32_cuda(features_rest, "features_rest"),
                           _f32_cuda(opacity, "opacity", (P, 1)), _f32_cuda(scaling, "scaling", (P, 3)),
                           _f32_cuda(rotation, "rotation", (P, 4)))


# ------------------------------------------------------------------ Adam
class FusedAdam(torch.optim.Adam):
    """``torch.optim.Adam(params, lr=0.0, eps=1e-15)`` as the reference builds it
    (gaussian_model.py:154-163), stepping every group in one kernel launch.
    Options the reference does not use (weight decay, amsgrad, maximize, capturable,
    sparse gradients) are rejected instead of being silently ignored."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, **kw):
        for k in ("weight_decay", "amsgrad", "maximize", "capturable", "differentiable", "fused", "foreach"):
            if kw.get(k):
                raise NotImplementedError(f"FusedAdam does not implement {k}")
            kw.pop(k, None)
        if kw:
            raise TypeError(f"unexpected arguments {sorted(kw)}")
        super().__init__(params, lr=lr, betas=betas, eps=eps)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        by_key = {}
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            if group.get("weight_decay", 0) or group.get("amsgrad") or group.get("maximize"):
                raise NotImplementedError("FusedAdam: weight_decay / amsgrad / maximize are not implemented")
            lr = group["lr"]
            if isinstance(lr, torch.Tensor):
                lr = float(lr)
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError("FusedAdam needs float32 CUDA parameters (no CPU path exists)")
                if not p.is_contiguous():
                    raise RuntimeError("FusedAdam needs contiguous parameters")
                state = self.state[p]
                if len(state) == 0:      # torch/optim/adam.py::_init_group
                    state["step"] = torch.tensor(0.0, dtype=torch.float32)
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["step"] += 1
                t = float(state["step"])
                grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                entry = _AdamTensor(p.data_ptr(), grad.data_ptr(), state["exp_avg"].data_ptr(),
                                    state["exp_avg_sq"].data_ptr(), p.numel(), lr / (1.0 - beta1 ** t),
                                    1.0 / math.sqrt(1.0 - beta2 ** t))
                by_key.setdefault((p.device, beta1, beta2, group["eps"]), []).append((entry, grad))
        for (dev, beta1, beta2, eps), entries in by_key.items():
            with torch.cuda.device(dev):
                for i in range(0, len(entries), ADAM_MAX_TENSORS):
                    chunk = entries[i:i + ADAM_MAX_TENSORS]
                    table = (_AdamTensor * len(chunk))(*[e for e, _ in chunk])
                    _call(_fns().b3gs_adam_multi, len(chunk), table, beta1, beta2, eps, _stream(dev))
        return loss


# ------------------------------------------------------------------ opacity decay, densify stats
@torch.no_grad()
def opacity_decay(opacity: torch.Tensor, factor: float = 0.99) -> torch.Tensor:
    """In place on the raw opacity parameter (P,1): ``inverse_sigmoid(sigmoid(x) * factor)``
    (scene/gaussian_model.py:307-309).  Returns the same tensor."""
    t = opacity.data
    if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
        raise RuntimeError("opacity must be a contiguous float32 CUDA tensor (no CPU path exists)")
    with torch.cuda.device(t.device):
        _call(_fns().b3gs_opacity_decay, t.numel(), float(factor), t.data_ptr(), _stream(t.device))
    return opacity


@torch.no_grad()
def add_densification_stats(viewspace_grad, radii, xyz_gradient_accum, denom, max_radii2D=None):
    """train.py:170-171 and scene/gaussian_model.py:409-411 for ``visibility_filter = radii > 0``:
    ``max_radii2D[v] = max(max_radii2D[v], radii[v])``;
    ``xyz_gradient_accum[v] += norm(viewspace_grad[v, :2])``; ``denom[v] += 1``.  In place.
    ``viewspace_grad`` is the (P,3) gradient of ``screenspace_points``; ``radii`` the int32
    (P,) output of the rasterizer."""
    P = int(radii.shape[0])
    g = _f32_cuda(viewspace_grad, "viewspace_grad", (P, 3))
    if radii.dtype != torch.int32 or not radii.is_cuda:
        raise RuntimeError("radii must be an int32 CUDA tensor")
    for t, name in ((xyz_gradient_accum, "xyz_gradient_accum"), (denom, "denom"), (max_radii2D, "max_radii2D")):
        if t is None:
            continue
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != P:
            raise RuntimeError(f"{name} must be a contiguous float32 CUDA tensor with P elements")
    with torch.cuda.device(g.device):
        _call(_fns().b3gs_densify_stats, P, g.data_ptr(), radii.contiguous().data_ptr(), xyz_gradient_accum.data_ptr(),
              denom.data_ptr(), None if max_radii2D is None else max_radii2D.data_ptr(), _stream(g.device))
