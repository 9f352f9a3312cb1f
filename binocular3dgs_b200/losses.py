"""Fused photometric loss behind the reference's loss API.

``ssim`` and ``l1_loss`` keep the names, argument meaning and return values of
``utils/loss_utils.py:18-21`` and ``:36-66``; ``photometric_loss`` is the combination the
trainer actually uses every iteration (``train.py:146-147``) evaluated by ONE forward and
ONE backward kernel through the C-ABI (``b3gs_photometric_forward/backward``) instead of
five depthwise conv2d calls plus elementwise kernels and their autograd duals.

Gradient flows to the first image (the render) only — the ground truth is a constant in
the reference's training loop; a ground truth that requires grad is rejected loudly.
"""
from __future__ import annotations

import ctypes

import torch

from . import _backend

_V = ctypes.c_void_p
_lib = None


def _fns():
    global _lib
    if _lib is None:
        lib = _backend.native().lib
        lib.b3gs_photometric_forward.argtypes = [ctypes.c_int] * 3 + [_V] * 7 + [ctypes.c_float] * 3 + [_V] * 2
        lib.b3gs_photometric_forward.restype = ctypes.c_int
        lib.b3gs_photometric_backward.argtypes = [ctypes.c_int] * 3 + [_V] * 6 + [ctypes.c_float] * 2 + [_V] * 2
        lib.b3gs_photometric_backward.restype = ctypes.c_int
        _lib = lib
    return _lib


def _as_chw(img: torch.Tensor, name: str):
    if img.dim() not in (3, 4):
        raise RuntimeError(f"{name} must be (C,H,W) or (N,C,H,W)")
    if not img.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (no CPU path exists)")
    if img.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32")
    H, W = int(img.shape[-2]), int(img.shape[-1])
    return img.contiguous().view(-1, H, W), H, W


class _Photometric(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img1, img2, w_ssim: float, w_l1: float, const: float):
        if img2.requires_grad:
            raise NotImplementedError("gradient w.r.t. the second image (ground truth) is not implemented")
        if img1.shape != img2.shape:
            raise RuntimeError("image shapes differ")
        a, H, W = _as_chw(img1, "img1")
        b, _, _ = _as_chw(img2, "img2")
        C = int(a.shape[0])
        dev = a.device
        with torch.cuda.device(dev):
            maps = torch.empty((3, C, H, W), dtype=torch.float32, device=dev)
            sums = torch.empty(3, dtype=torch.float64, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            n = float(C * H * W)
            # loss = const + w_ssim * mean(SSIM) + w_l1 * mean|x-y|, formed by the kernel's last block
            rc = _fns().b3gs_photometric_forward(C, H, W, a.data_ptr(), b.data_ptr(), maps[0].data_ptr(),
                                                 maps[1].data_ptr(), maps[2].data_ptr(), None, sums.data_ptr(),
                                                 const, w_ssim / n, w_l1 / n, loss.data_ptr(),
                                                 torch.cuda.current_stream(dev).cuda_stream)
        if rc != 0:
            raise RuntimeError(f"b3gs_photometric_forward failed ({rc})")
        ctx.save_for_backward(a, b, maps)
        ctx.meta = (C, H, W, w_ssim / n, w_l1 / n, img1.shape)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        a, b, maps = ctx.saved_tensors
        C, H, W, ks, kl, shape = ctx.meta
        dev = a.device
        with torch.cuda.device(dev):
            up = grad_out.reshape(1).to(torch.float32).contiguous()
            grad = torch.empty((C, H, W), dtype=torch.float32, device=dev)
            rc = _fns().b3gs_photometric_backward(C, H, W, a.data_ptr(), b.data_ptr(), maps[0].data_ptr(),
                                                  maps[1].data_ptr(), maps[2].data_ptr(), up.data_ptr(), ks, kl,
                                                  grad.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        if rc != 0:
            raise RuntimeError(f"b3gs_photometric_backward failed ({rc})")
        return grad.view(shape), None, None, None, None


def ssim(img1, img2, window_size=11, size_average=True):
    """Mean SSIM, utils/loss_utils.py:36-66 (11x11 Gaussian window, sigma 1.5, zero padding)."""
    if window_size != 11 or not size_average:
        raise NotImplementedError("only window_size=11, size_average=True (the reference's call sites)")
    return _Photometric.apply(img1, img2, 1.0, 0.0, 0.0)


def l1_loss(network_output, gt, mask=None):
    """utils/loss_utils.py:18-21, verbatim semantics (two elementwise torch kernels)."""
    if mask is not None:
        return torch.abs((network_output * mask - gt * mask)).mean()
    return torch.abs((network_output - gt)).mean()


def photometric_loss(image, gt, lambda_dssim: float = 0.2):
    """(1 - lambda) * l1_loss(image, gt) + lambda * (1 - ssim(image, gt))  (train.py:146-147)."""
    return _Photometric.apply(image, gt, -float(lambda_dssim), 1.0 - float(lambda_dssim), float(lambda_dssim))
