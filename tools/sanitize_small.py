"""Every kernel family once at small, ragged sizes — the workload for
`compute-sanitizer --tool memcheck|racecheck|initcheck python tools/sanitize_small.py`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from binocular3dgs_b200 import _backend, binocular, losses, parameters  # noqa: E402
from binocular3dgs_b200.simple_knn import distCUDA2  # noqa: E402
from workloads import make_camera, make_pixel_grads, make_scene  # noqa: E402

dev = torch.device("cuda:0")
nat = _backend.native()
for (P, W, H, kind) in ((3000, 97, 61, "cube"), (1500, 64, 48, "shell"), (5, 16, 16, "cube")):
    scene, cam = make_scene(P, seed=P, kind=kind).to(dev), make_camera(W, H).to(dev)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    grads = tuple(g.to(dev) for g in make_pixel_grads(W, H, 2))
    for pix in (1, 2, 4):
        nat.lib.b3gs_set_backward_pixels(pix)
        util.surface_forward_backward(nat, scene, cam, bg, grads)
nat.lib.b3gs_set_backward_pixels(0)
util.surface_forward_backward(_backend.preferred(), scene, cam, bg, grads)
nat.mark_visible(scene.means3D, cam.world_view_transform, cam.full_proj_transform)
g = torch.Generator().manual_seed(0)
for (C, H, W) in ((3, 37, 53), (3, 64, 48), (1, 33, 100)):
    a = torch.rand(C, H, W, generator=g).to(dev).requires_grad_(True)
    b = torch.rand(C, H, W, generator=g).to(dev)
    losses.photometric_loss(a, b, 0.2).backward()
    losses.ssim(a, b).backward()
for (H, W) in ((37, 53), (20, 33), (3, 3)):
    s = torch.rand(3, H, W, generator=g).to(dev).requires_grad_(True)
    d = (torch.rand(1, H, W, generator=g) * 5).to(dev).requires_grad_(True)
    gt = torch.rand(3, H, W, generator=g).to(dev)
    binocular.binocular_consistency_loss(s, d, gt, 60.0, 0.3).backward()
    disp = (torch.rand(1, 1, H, W, generator=g) * 40 - 20).to(dev).requires_grad_(True)
    w = binocular.inverse_warp_images(s.detach().unsqueeze(0).requires_grad_(True), disp)
    w.sum().backward()
    binocular.SmoothLoss()(disp.detach().requires_grad_(True), gt.unsqueeze(0)).backward()
for (P, M) in ((257, 4), (64, 1), (130, 16)):
    raw = dict(f_dc=torch.randn(P, 1, 3), f_rest=torch.randn(P, M - 1, 3), opacity=torch.randn(P, 1),
               scaling=torch.randn(P, 3), rotation=torch.randn(P, 4))
    raw = {k: torch.nn.Parameter(v.to(dev)) for k, v in raw.items()}
    acts = parameters.activate(raw["f_dc"], raw["f_rest"], raw["opacity"], raw["scaling"], raw["rotation"])
    sum(a.sum() for a in acts).backward()
    opt = parameters.FusedAdam([{"params": [v], "lr": 1e-3} for v in raw.values()], lr=0.0, eps=1e-15)
    opt.step()
    parameters.opacity_decay(raw["opacity"], 0.99)
    parameters.add_densification_stats(torch.randn(P, 3, device=dev), torch.randint(0, 5, (P,), device=dev, dtype=torch.int32),
                                       torch.zeros(P, 1, device=dev), torch.zeros(P, 1, device=dev), torch.zeros(P, device=dev))
for P in (1, 3, 33, 1025, 5000):
    distCUDA2(torch.rand(P, 3, generator=g).to(dev))
torch.cuda.synchronize()
print("sanitize_small: done")
