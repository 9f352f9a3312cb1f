"""Host-side cost of one operator call (Python + ctypes + allocator), measured on a problem
so small that the GPU is never the bottleneck.  usage: python tools/host_overhead.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from binocular3dgs_b200 import _backend, losses, parameters  # noqa: E402
from binocular3dgs_b200.rasterizer import make_surface  # noqa: E402
from workloads import make_camera, make_pixel_grads, make_scene  # noqa: E402

dev = torch.device("cuda:0")
scene, cam = make_scene(2000, seed=1).to(dev), make_camera(64, 64).to(dev)
bg = torch.zeros(3, device=dev)
grads = tuple(g.to(dev) for g in make_pixel_grads(64, 64))
nat = _backend.native()
S = make_surface(_backend.native() if os.environ.get("B3GS_HOST") == "ctypes" else _backend.preferred())
print("host side:", type(S._C).__name__)
e = torch.empty(0)


def wall(fn, n=300):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / n * 1e6


out = util.raw_forward(nat, scene, cam, bg)
print("raw forward   us/call:", round(wall(lambda: util.raw_forward(nat, scene, cam, bg)), 1))
print("raw backward  us/call:", round(wall(lambda: nat.rasterize_gaussians_backward(
    bg, scene.means3D, out["radii"], e, scene.scales, scene.rotations, 1.0, e, cam.world_view_transform,
    cam.full_proj_transform, cam.tanfovx, cam.tanfovy, grads[0], grads[1], grads[2], scene.shs, scene.sh_degree,
    cam.camera_center, out["geom"], out["R"], out["binning"], out["img"], out["alpha"], False)), 1))
leaves = [t.detach().clone().requires_grad_(True) for t in scene.tensors()]
m3, sc, ro, op, sh = leaves
settings = util.settings_for(cam, bg, scene.sh_degree)
gt = torch.rand(3, 64, 64, device=dev)


def step_l1():
    m2 = torch.zeros_like(m3, requires_grad=True)
    color, radii, depth, alpha = S.GaussianRasterizer(settings)(means3D=m3, means2D=m2, opacities=op, shs=sh, scales=sc,
                                                                rotations=ro)
    loss = (color - gt).abs().mean()
    loss.backward()


def step_fused():
    m2 = torch.zeros_like(m3, requires_grad=True)
    color, radii, depth, alpha = S.GaussianRasterizer(settings)(means3D=m3, means2D=m2, opacities=op, shs=sh, scales=sc,
                                                                rotations=ro)
    losses.photometric_loss(color, gt).backward()


print("surface fwd + L1 + autograd bwd  us/step:", round(wall(step_l1), 1))
print("surface fwd + fused photometric + bwd us/step:", round(wall(step_fused), 1))
raw = {k: torch.nn.Parameter(v) for k, v in dict(a=torch.randn(2000, 3, device=dev), b=torch.randn(2000, 4, device=dev)).items()}
for v in raw.values():
    v.grad = torch.randn_like(v)
opt = parameters.FusedAdam([{"params": [v], "lr": 1e-3} for v in raw.values()], lr=0.0, eps=1e-15)
print("FusedAdam.step us/call:", round(wall(opt.step), 1))
opt2 = torch.optim.Adam([{"params": [v], "lr": 1e-3} for v in raw.values()], lr=0.0, eps=1e-15)
print("torch Adam.step us/call:", round(wall(opt2.step), 1))

if os.environ.get("B3GS_CPROFILE"):
    import cProfile
    import pstats
    pr = cProfile.Profile()
    for _ in range(50):
        step_l1()
    torch.cuda.synchronize()
    pr.enable()
    for _ in range(300):
        step_l1()
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(45)
