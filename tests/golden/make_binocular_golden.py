"""Golden vectors for the binocular-consistency loss FROM THE REFERENCE'S OWN functions:
utils/graphics_utils.py inverse_warp_images, utils/loss_utils.py l1_loss and SmoothLoss,
combined exactly as train.py:128-136 does.

The reference hard-codes `.cuda()`; it is otherwise pure PyTorch, so this script stubs
Tensor.cuda()/Module.cuda() to the identity and runs it on CPU in the build container:

    python tests/golden/make_binocular_golden.py   # needs /root/reference; writes tests/golden/bino_*.npz

Inputs are regenerated from the stored seed; stored are the warped image, the warped mask,
SmoothLoss, the combined loss and its autograd gradients w.r.t. the shifted image and the
depth map, plus stand-alone inverse_warp_images / SmoothLoss gradients for the drop-in
operators.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
# trans_dist < 0 shifts content right (disparity > 0), > 0 left; near_depth makes the
# disparity exceed the image width for some pixels (both taps out of range).
CASES = {
    "bino_small": dict(H=37, W=53, seed=11, focal_x=60.0, trans_dist=0.31, depth_lo=2.0, depth_hi=6.0),
    "bino_negative_shift": dict(H=48, W=64, seed=12, focal_x=75.0, trans_dist=-0.22, depth_lo=1.5, depth_hi=5.0),
    "bino_out_of_range": dict(H=20, W=33, seed=13, focal_x=90.0, trans_dist=0.4, depth_lo=0.0, depth_hi=3.0),
}


def make_inputs(H, W, seed, focal_x, trans_dist, depth_lo, depth_hi):
    """shifted image (3,H,W), depth (1,H,W) with zeros where depth_lo == 0 (empty pixels
    render depth 0 in the rasterizer), gt (3,H,W)."""
    g = torch.Generator().manual_seed(seed)
    k = torch.ones(1, 1, 5, 5) / 25.0
    base = torch.rand(3, H, W, generator=g)
    gt = torch.nn.functional.conv2d(base.unsqueeze(1), k, padding=2).squeeze(1).clamp(0, 1)
    shifted = (gt + 0.1 * torch.randn(3, H, W, generator=g)).clamp(0, 1)
    d = torch.rand(1, H, W, generator=g)
    d = torch.nn.functional.conv2d(d.unsqueeze(1), k, padding=2).squeeze(1) * 25.0 / 16.0
    depth = depth_lo + (depth_hi - depth_lo) * d.clamp(0, 1)
    if depth_lo == 0.0:
        depth = torch.where(torch.rand(1, H, W, generator=g) < 0.15, torch.zeros_like(depth), depth)
    return shifted.contiguous(), depth.contiguous(), gt.contiguous()


def reference_functions():
    torch.Tensor.cuda = lambda self, *a, **k: self          # the reference hard-codes .cuda()
    torch.nn.Module.cuda = lambda self, *a, **k: self
    sys.path.insert(0, "/root/reference")
    for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
        del sys.modules[k]
    from utils.graphics_utils import inverse_warp_images
    from utils.loss_utils import SmoothLoss, l1_loss
    return inverse_warp_images, SmoothLoss, l1_loss


def reference_loss(fns, shifted_image, depth, gt_image, focal_x, trans_dist):
    """train.py:52-55 and :128-136, same statements."""
    inverse_warp_images, SmoothLoss, l1_loss = fns
    image_height, image_width = depth.shape[-2:]
    row_indices = torch.arange(0, image_height).view(-1, 1).repeat(1, image_width)
    column_indices = torch.arange(0, image_width).repeat(image_height, 1)
    mask = torch.ones((1, image_height, image_width), dtype=torch.float32)
    smooth_loss = SmoothLoss()
    disparity = focal_x * (-trans_dist) / (depth + 1e-5)
    warped_image = inverse_warp_images(shifted_image.unsqueeze(0), disparity.unsqueeze(0), row_indices, column_indices)
    shift_mask = inverse_warp_images(mask.unsqueeze(0), disparity.unsqueeze(0), row_indices, column_indices)
    sm = smooth_loss.forward(disparity=disparity * shift_mask, image=gt_image.unsqueeze(0))
    l1 = l1_loss(warped_image, gt_image.unsqueeze(0), mask=shift_mask)
    return l1 + 0.05 * sm, l1, sm, warped_image, shift_mask, disparity


def main():
    fns = reference_functions()
    inverse_warp_images, SmoothLoss, _ = fns
    for name, c in CASES.items():
        shifted, depth, gt = make_inputs(**c)
        a, d = shifted.clone().requires_grad_(True), depth.clone().requires_grad_(True)
        loss, l1, sm, warped, shift_mask, disparity = reference_loss(fns, a, d, gt, c["focal_x"], c["trans_dist"])
        loss.backward()
        out = dict(loss=np.float64(loss.item()), l1=np.float64(l1.item()), smooth=np.float64(sm.item()),
                   warped=warped.detach().numpy()[0], shift_mask=shift_mask.detach().numpy()[0, 0],
                   grad_shifted=a.grad.numpy(), grad_depth=d.grad.numpy())
        # stand-alone operators with a dense upstream gradient (drop-in surface)
        H, W = depth.shape[-2:]
        rows = torch.arange(0, H).view(-1, 1).repeat(1, W)
        cols = torch.arange(0, W).repeat(H, 1)
        g = torch.Generator().manual_seed(c["seed"] + 100)
        up = torch.randn(1, 3, H, W, generator=g)
        a2 = shifted.clone().requires_grad_(True)
        disp2 = disparity.detach().clone().requires_grad_(True)
        w2 = inverse_warp_images(a2.unsqueeze(0), disp2.unsqueeze(0), rows, cols)
        w2.backward(up)
        out.update(warp_up=up.numpy()[0], warp_grad_image=a2.grad.numpy(), warp_grad_disparity=disp2.grad.numpy())
        disp3 = (disparity.detach() * shift_mask.detach()[0]).clone().requires_grad_(True)   # (1,H,W)
        s3 = SmoothLoss().forward(disparity=disp3.unsqueeze(0), image=gt.unsqueeze(0))
        s3.backward()
        out.update(smooth_grad_disparity=disp3.grad.numpy())
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: float(v) for k, v in out.items() if np.ndim(v) == 0},
              "valid frac", float((shift_mask != 0).float().mean()))


if __name__ == "__main__":
    main()
