"""Timing of the fused photometric loss against the same loss
written the way the reference writes it (utils/loss_utils.py: five depthwise conv2d +
elementwise ops, autograd backward), both on the GPU, CUDA events, L2 flushed between
iterations.  Used by bench.py ("next_rows") and runnable on its own."""
import os
import sys
from math import exp

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _torch_style_loss(img, gt, lam=0.2):
    """What train.py:146-147 + utils/loss_utils.py:36-66 execute (restated, same ops)."""
    C = img.size(-3)
    g = torch.Tensor([exp(-(x - 11 // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
    g = (g / g.sum()).unsqueeze(1)
    w = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0).expand(C, 1, 11, 11).contiguous().to(img.device)
    mu1, mu2 = F.conv2d(img, w, padding=5, groups=C), F.conv2d(gt, w, padding=5, groups=C)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img * img, w, padding=5, groups=C) - mu1_sq
    s2 = F.conv2d(gt * gt, w, padding=5, groups=C) - mu2_sq
    s12 = F.conv2d(img * gt, w, padding=5, groups=C) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim = (((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))).mean()
    return (1.0 - lam) * torch.abs(img - gt).mean() + lam * (1.0 - ssim)


def measure(dev, C=3, H=800, W=800, iters=30, warmup=5):
    from binocular3dgs_b200 import losses
    g = torch.Generator().manual_seed(0)
    gt = torch.rand(C, H, W, generator=g).to(dev)
    img = (gt + 0.05 * torch.randn(C, H, W, generator=g).to(dev)).clamp(0, 1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def run(fn):
        ts = []
        for i in range(warmup + iters):
            a = img.clone().requires_grad_(True)
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            v = fn(a, gt)
            v.backward()
            e1.record()
            torch.cuda.synchronize()
            if i >= warmup:
                ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2], float(v.detach()), a.grad

    # the two kernels alone, through the C-ABI on preallocated buffers
    fns = losses._fns()
    maps = torch.empty((3, C, H, W), device=dev)
    sums = torch.empty(3, dtype=torch.float64, device=dev)
    loss_out = torch.empty((), device=dev)
    up = torch.ones(1, device=dev)
    grad = torch.empty((C, H, W), device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    kt = []
    for i in range(warmup + iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fns.b3gs_photometric_forward(C, H, W, img.data_ptr(), gt.data_ptr(), maps[0].data_ptr(), maps[1].data_ptr(),
                                     maps[2].data_ptr(), None, sums.data_ptr(), 0.2, -0.2 / (C * H * W),
                                     0.8 / (C * H * W), loss_out.data_ptr(), st)
        fns.b3gs_photometric_backward(C, H, W, img.data_ptr(), gt.data_ptr(), maps[0].data_ptr(), maps[1].data_ptr(),
                                      maps[2].data_ptr(), up.data_ptr(), -0.2 / (C * H * W), 0.8 / (C * H * W),
                                      grad.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        if i >= warmup:
            kt.append(e0.elapsed_time(e1))
    t_kernels = sorted(kt)[len(kt) // 2]
    t_fused, v_fused, g_fused = run(lambda a, b: losses.photometric_loss(a, b, 0.2))
    t_torch, v_torch, g_torch = run(_torch_style_loss)
    n = C * H * W
    alg_bytes = n * (4 * 2 + 4 * 3) + n * (4 * 5 + 4)     # fwd: 2 reads + 3 writes; bwd: 5 reads + 1 write
    return {"what": "0.8*L1 + 0.2*(1-SSIM) forward+backward, %dx%dx%d" % (C, H, W), "fused_ms": round(t_fused, 4),
            "torch_reference_style_ms": round(t_torch, 4), "speedup": round(t_torch / t_fused, 2),
            "kernels_only_ms": round(t_kernels, 4), "alg_bytes": alg_bytes,
            "kernels_GBps": round(alg_bytes / t_kernels / 1e6, 1),
            "value_abs_diff": abs(v_fused - v_torch),
            "grad_rel_diff": float((g_fused - g_torch).abs().max() / g_torch.abs().max())}


def _torch_style_warp(image, disparity, r_ind, c_ind):
    """What utils/graphics_utils.py:80-125 executes (restated, same ops, including the
    LongTensor round trip through the host)."""
    x0 = torch.floor(disparity).type(torch.LongTensor).to(image.device)
    x1 = x0 + 1
    W = image.size(3)
    batches = []
    for b in range(image.size(0)):
        chans = []
        for ch in range(image.size(1)):
            c0 = c_ind + x0[b, 0]
            bad0 = (c0 < 0) | (c0 >= W)
            c0[c0 >= W] = W - 1
            c0[c0 < 0] = 0
            c1 = c_ind + x1[b, 0]
            bad1 = (c1 < 0) | (c1 >= W)
            c1[c1 >= W] = W - 1
            c1[c1 < 0] = 0
            v = ((x1[b, 0] - disparity[b, 0]) * image[b, ch, r_ind, c0] +
                 (disparity[b, 0] - x0[b, 0]) * image[b, ch, r_ind, c1]).unsqueeze(0).unsqueeze(0)
            v[0, 0, bad0] = 0.0
            v[0, 0, bad1] = 0.0
            chans.append(v)
        batches.append(torch.cat(chans, 1))
    return torch.cat(batches, 0)


class _TorchStyleSmooth(torch.nn.Module):
    """utils/loss_utils.py:68-91 restated: four fixed 3x3 convolutions."""

    def __init__(self, dev):
        super().__init__()
        kx = torch.tensor([[0, 0, 0], [-0.5, 0, 0.5], [0, 0, 0]], device=dev)
        self.wx3, self.wy3 = kx.expand(1, 3, 3, 3).contiguous(), kx.t().expand(1, 3, 3, 3).contiguous()
        self.wx1, self.wy1 = kx.view(1, 1, 3, 3).contiguous(), kx.t().reshape(1, 1, 3, 3).contiguous()

    def forward(self, disparity, image):
        ex = torch.exp(F.conv2d(image, self.wx3).abs() * -0.33)
        ey = torch.exp(F.conv2d(image, self.wy3).abs() * -0.33)
        return (ex * F.conv2d(disparity, self.wx1)).abs().mean() + (ey * F.conv2d(disparity, self.wy1)).abs().mean()


def measure_binocular(dev, H=756, W=1008, iters=30, warmup=5, focal_x=815.0, trans_dist=0.27):
    """train.py:128-136 forward + backward: fused kernels vs the reference's op sequence."""
    from binocular3dgs_b200 import binocular
    g = torch.Generator().manual_seed(1)
    k = torch.ones(1, 1, 9, 9) / 81.0
    gt = F.conv2d(torch.rand(3, 1, H, W, generator=g), k, padding=4).squeeze(1).clamp(0, 1).to(dev)
    shifted = (gt + 0.05 * torch.randn(3, H, W, generator=g).to(dev)).clamp(0, 1)
    depth = (2.0 + 6.0 * F.conv2d(torch.rand(1, 1, H, W, generator=g), k, padding=4).squeeze(1)).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = torch.arange(0, H).view(-1, 1).repeat(1, W).to(dev)
    cols = torch.arange(0, W).repeat(H, 1).to(dev)
    ones = torch.ones((1, H, W), dtype=torch.float32, device=dev)
    smooth = _TorchStyleSmooth(dev)

    def torch_style(a, d):
        disparity = focal_x * (-trans_dist) / (d + 1e-5)
        warped = _torch_style_warp(a.unsqueeze(0), disparity.unsqueeze(0), rows, cols)
        mask = _torch_style_warp(ones.unsqueeze(0), disparity.unsqueeze(0), rows, cols)
        l1 = torch.abs(warped * mask - gt.unsqueeze(0) * mask).mean()
        return l1 + 0.05 * smooth(disparity * mask, gt.unsqueeze(0))

    def fused(a, d):
        return binocular.binocular_consistency_loss(a, d, gt, focal_x, trans_dist)

    def run(fn):
        ts = []
        for i in range(warmup + iters):
            a = shifted.clone().requires_grad_(True)
            d = depth.clone().requires_grad_(True)
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            v = fn(a, d)
            v.backward()
            e1.record()
            torch.cuda.synchronize()
            if i >= warmup:
                ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2], float(v.detach()), a.grad, d.grad

    fns = binocular._fns()
    sums = torch.empty(4, dtype=torch.float64, device=dev)
    loss_out = torch.empty((), device=dev)
    up = torch.ones(1, device=dev)
    g_s, g_d = torch.empty_like(shifted), torch.empty_like(depth)
    st = torch.cuda.current_stream(dev).cuda_stream
    k_disp = focal_x * (-trans_dist)
    kt = []
    for i in range(warmup + iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fns.b3gs_binocular_forward(H, W, shifted.data_ptr(), depth.data_ptr(), gt.data_ptr(), k_disp, sums.data_ptr(),
                                   1.0 / (3 * H * W), 0.05 / ((H - 2) * (W - 2)), loss_out.data_ptr(), st)
        fns.b3gs_binocular_backward(H, W, shifted.data_ptr(), depth.data_ptr(), gt.data_ptr(), k_disp,
                                    up.data_ptr(), 1.0 / (3 * H * W), 0.05 / ((H - 2) * (W - 2)), g_s.data_ptr(),
                                    g_d.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        if i >= warmup:
            kt.append(e0.elapsed_time(e1))
    t_kernels = sorted(kt)[len(kt) // 2]
    t_fused, v_fused, gs_f, gd_f = run(fused)
    t_torch, v_torch, gs_t, gd_t = run(torch_style)
    n = H * W
    alg_bytes = n * 28 + n * (28 + 12 + 24 + 4)   # fwd: 7 reads; bwd: 7 reads, memset 3, 6 RED (read-modify-write = 3 planes r+w), 1 write
    return {"what": "binocular consistency loss (warp + masked L1 + 0.05*smooth) forward+backward, 3x%dx%d" % (H, W),
            "fused_ms": round(t_fused, 4), "torch_reference_style_ms": round(t_torch, 4),
            "speedup": round(t_torch / t_fused, 2), "kernels_only_ms": round(t_kernels, 4), "alg_bytes": alg_bytes,
            "kernels_GBps": round(alg_bytes / t_kernels / 1e6, 1), "value_abs_diff": abs(v_fused - v_torch),
            "grad_shifted_rel_diff": float((gs_f - gs_t).abs().max() / gs_t.abs().max()),
            "grad_depth_rel_diff": float((gd_f - gd_t).abs().max() / gd_t.abs().max())}


if __name__ == "__main__":
    import json
    print(json.dumps(measure(torch.device("cuda:0"))))
    print(json.dumps(measure_binocular(torch.device("cuda:0"))))
