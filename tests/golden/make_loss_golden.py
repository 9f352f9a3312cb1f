"""Golden vectors for the photometric loss FROM THE REFERENCE'S OWN utils/loss_utils.py.

The reference's losses are pure PyTorch, so they run on CPU in the build container:

    python tests/golden/make_loss_golden.py        # needs /root/reference; writes tests/golden/loss_*.npz

Inputs are regenerated from the stored seed; stored are the reference's ssim(), l1_loss(),
the combined training loss (train.py:146-147, lambda_dssim = 0.2) and their autograd
gradients w.r.t. the first image.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"loss_small": dict(C=3, H=37, W=53, seed=5), "loss_tile_edges": dict(C=3, H=64, W=48, seed=6),
         "loss_one_channel": dict(C=1, H=20, W=90, seed=7)}


def make_images(C, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(C, H, W, generator=g)
    k = torch.ones(1, 1, 5, 5) / 25.0
    smooth = torch.nn.functional.conv2d(base.unsqueeze(1), k, padding=2).squeeze(1)   # structure, not white noise
    gt = smooth.clamp(0, 1)
    img = (smooth + 0.1 * torch.randn(C, H, W, generator=g)).clamp(0, 1)
    return img.contiguous(), gt.contiguous()


def main():
    sys.path.insert(0, "/root/reference")
    from utils.loss_utils import l1_loss, ssim  # the reference's own functions
    for name, c in CASES.items():
        img, gt = make_images(**c)
        out = {}
        for key, fn in (("ssim", lambda a: ssim(a, gt)), ("l1", lambda a: l1_loss(a, gt)),
                        ("train", lambda a: 0.8 * l1_loss(a, gt) + 0.2 * (1.0 - ssim(a, gt)))):
            a = img.clone().requires_grad_(True)
            v = fn(a)
            v.backward()
            out[key] = np.float64(v.item())
            out["grad_" + key] = a.grad.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: float(v) for k, v in out.items() if not k.startswith("grad")})


if __name__ == "__main__":
    main()
