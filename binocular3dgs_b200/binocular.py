"""Binocular-consistency loss behind the reference's own names.

``inverse_warp_images`` keeps the signature of ``utils/graphics_utils.py:80`` and
``SmoothLoss`` the one of ``utils/loss_utils.py:68-91`` so ``train.py:128-136`` runs
unchanged on them; ``binocular_consistency_loss`` is those nine statements evaluated by
ONE forward and ONE backward kernel through the C-ABI (``b3gs_binocular_forward/backward``).

Gradient flows to the warped image and the disparity / depth; the ground-truth image and the
constant ``mask`` are constants in the reference's training loop (a tensor there that
requires grad is rejected loudly).  CUDA float32 only: there is no CPU path.
"""
from __future__ import annotations

import ctypes

import torch

from . import _backend

_V, _I, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
_lib = None


def _fns():
    global _lib
    if _lib is None:
        lib = _backend.native().lib
        for name, args in (("b3gs_binocular_forward", [_I, _I, _V, _V, _V, _F, _V, _F, _F, _V, _V]),
                           ("b3gs_binocular_backward", [_I, _I, _V, _V, _V, _F, _V, _F, _F, _V, _V, _V]),
                           ("b3gs_warp_forward", [_I, _I, _I, _V, _V, _V, _V]),
                           ("b3gs_warp_backward", [_I, _I, _I, _V, _V, _V, _V, _V, _V]),
                           ("b3gs_smooth_forward", [_I, _I, _V, _V, _V, _V]),
                           ("b3gs_smooth_backward", [_I, _I, _V, _V, _V, _F, _V, _V])):
            fn = getattr(lib, name)
            fn.argtypes, fn.restype = args, _I
        _lib = lib
    return _lib


def _check(t: torch.Tensor, name: str, shape_tail=None):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (no CPU path exists)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32")
    if shape_tail is not None and tuple(t.shape[-len(shape_tail):]) != tuple(shape_tail):
        raise RuntimeError(f"{name} has shape {tuple(t.shape)}, expected (..., {', '.join(map(str, shape_tail))})")
    return t.contiguous()


def _call(fn, *args):
    rc = fn(*args)
    if rc != 0:
        raise RuntimeError(f"{fn.__name__} failed ({rc})")


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


# ------------------------------------------------------------------ inverse_warp_images
class _InverseWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, disparity):
        B, C, H, W = image.shape
        dev = image.device
        out = torch.empty_like(image)
        with torch.cuda.device(dev):
            for b in range(B):
                _call(_fns().b3gs_warp_forward, C, H, W, image[b].data_ptr(), disparity[b].data_ptr(),
                      out[b].data_ptr(), _stream(dev))
        ctx.save_for_backward(image, disparity)
        return out

    @staticmethod
    def backward(ctx, g):
        image, disparity = ctx.saved_tensors
        B, C, H, W = image.shape
        dev = image.device
        g = g.contiguous()
        g_img = torch.empty_like(image)
        g_disp = torch.empty_like(disparity)
        with torch.cuda.device(dev):
            for b in range(B):
                _call(_fns().b3gs_warp_backward, C, H, W, image[b].data_ptr(), disparity[b].data_ptr(),
                      g[b].data_ptr(), g_img[b].data_ptr(), g_disp[b].data_ptr(), _stream(dev))
        return g_img, g_disp


def inverse_warp_images(image, disparity, row_indices=None, column_indices=None):
    """utils/graphics_utils.py:80-125.  image (B,C,H,W), disparity (B,1,H,W) -> (B,C,H,W).
    ``row_indices`` / ``column_indices`` are the reference's precomputed index grids
    (train.py:53-54); the kernel derives them from the thread index, so they are accepted
    and only shape-checked."""
    if image.dim() != 4 or disparity.dim() != 4 or disparity.shape[1] != 1:
        raise RuntimeError("image must be (B,C,H,W) and disparity (B,1,H,W)")
    B, C, H, W = image.shape
    image = _check(image, "image")
    disparity = _check(disparity, "disparity", (H, W))
    if disparity.shape[0] != B:
        raise RuntimeError("image and disparity batch sizes differ")
    for idx, name in ((row_indices, "row_indices"), (column_indices, "column_indices")):
        if idx is not None and tuple(idx.shape) != (H, W):
            raise RuntimeError(f"{name} must be ({H},{W})")
    return _InverseWarp.apply(image, disparity)


# ------------------------------------------------------------------ SmoothLoss
class _Smooth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disparity, image):
        H, W = disparity.shape[-2:]
        dev = disparity.device
        with torch.cuda.device(dev):
            sums = torch.empty(2, dtype=torch.float64, device=dev)
            _call(_fns().b3gs_smooth_forward, H, W, disparity.data_ptr(), image.data_ptr(), sums.data_ptr(),
                  _stream(dev))
        ctx.save_for_backward(disparity, image)
        return (sums.sum() / float((H - 2) * (W - 2))).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        disparity, image = ctx.saved_tensors
        H, W = disparity.shape[-2:]
        dev = disparity.device
        with torch.cuda.device(dev):
            up = g.reshape(1).to(torch.float32).contiguous()
            out = torch.empty_like(disparity)
            _call(_fns().b3gs_smooth_backward, H, W, disparity.data_ptr(), image.data_ptr(), up.data_ptr(),
                  1.0 / float((H - 2) * (W - 2)), out.data_ptr(), _stream(dev))
        return out, None


class SmoothLoss(torch.nn.Module):
    """utils/loss_utils.py:68-91: edge-aware first-order smoothness of a disparity map.
    The reference builds four fixed 3x3 Conv2d layers; the kernel hard-codes the same
    central-difference stencils, so this module has no parameters."""

    def forward(self, disparity, image):
        if image.dim() != 4 or image.shape[0] != 1 or image.shape[1] != 3:
            raise RuntimeError("image must be (1,3,H,W)")
        H, W = image.shape[-2:]
        if H < 3 or W < 3:
            raise RuntimeError("SmoothLoss needs H, W >= 3 (3x3 stencils without padding)")
        if disparity.numel() != H * W:
            raise RuntimeError("disparity must hold one (H,W) map")
        if image.requires_grad:
            raise NotImplementedError("gradient w.r.t. the image (ground truth) is not implemented")
        return _Smooth.apply(_check(disparity, "disparity", (H, W)), _check(image, "image"))


# ------------------------------------------------------------------ the fused loss
class _Binocular(torch.autograd.Function):
    @staticmethod
    def forward(ctx, shifted, depth, gt, k_disp: float, smooth_weight: float):
        H, W = depth.shape[-2:]
        dev = depth.device
        with torch.cuda.device(dev):
            sums = torch.empty(4, dtype=torch.float64, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            k_l1, k_sm = 1.0 / float(3 * H * W), smooth_weight / float((H - 2) * (W - 2))
            _call(_fns().b3gs_binocular_forward, H, W, shifted.data_ptr(), depth.data_ptr(), gt.data_ptr(), k_disp,
                  sums.data_ptr(), k_l1, k_sm, loss.data_ptr(), _stream(dev))
        ctx.save_for_backward(shifted, depth, gt)
        ctx.meta = (H, W, k_disp, k_l1, k_sm)
        return loss

    @staticmethod
    def backward(ctx, g):
        shifted, depth, gt = ctx.saved_tensors
        H, W, k_disp, k_l1, k_sm = ctx.meta
        dev = depth.device
        with torch.cuda.device(dev):
            up = g.reshape(1).to(torch.float32).contiguous()
            g_shifted = torch.empty_like(shifted)
            g_depth = torch.empty_like(depth)
            _call(_fns().b3gs_binocular_backward, H, W, shifted.data_ptr(), depth.data_ptr(), gt.data_ptr(), k_disp,
                  up.data_ptr(), k_l1, k_sm, g_shifted.data_ptr(), g_depth.data_ptr(), _stream(dev))
        return g_shifted, g_depth, None, None, None


def binocular_consistency_loss(shifted_image, depth, gt_image, focal_x: float, trans_dist: float,
                               smooth_weight: float = 0.05):
    """train.py:128-136 in one call: the second render ``shifted_image`` (3,H,W) of the
    camera translated by ``trans_dist`` along its x axis is warped back onto the first
    view with the disparity ``focal_x * (-trans_dist) / (depth + 1e-5)`` of the first
    render's ``depth`` (1,H,W) and compared with ``gt_image`` (3,H,W):
    ``l1_loss(warped, gt, mask) + smooth_weight * SmoothLoss(disparity * mask, gt)``."""
    if shifted_image.dim() != 3 or shifted_image.shape[0] != 3:
        raise RuntimeError("shifted_image must be (3,H,W)")
    H, W = shifted_image.shape[-2:]
    if H < 3 or W < 3:
        raise RuntimeError("the smoothness term needs H, W >= 3")
    if depth.numel() != H * W:
        raise RuntimeError("depth must be (1,H,W)")
    if gt_image.requires_grad:
        raise NotImplementedError("gradient w.r.t. the ground-truth image is not implemented")
    k_disp = float(focal_x) * (-float(trans_dist))          # python double, as the reference forms it
    return _Binocular.apply(_check(shifted_image, "shifted_image"), _check(depth, "depth", (H, W)),
                            _check(gt_image, "gt_image", (3, H, W)), k_disp, float(smooth_weight))
