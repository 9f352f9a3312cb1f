"""Run N forward+backward steps of one backend on one config (for ncu launch lists).
usage: python tools/gpu_step.py <native|ref> <config[:kind]> <steps>"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from binocular3dgs_b200 import _backend  # noqa: E402
from workloads import CONFIGS, make_camera, make_pixel_grads, make_scene  # noqa: E402
import util  # noqa: E402

which, name, steps = sys.argv[1], sys.argv[2], int(sys.argv[3])
kind = "cube"
if ":" in name:
    name, kind = name.split(":")
if which == "ref":
    from oracle import refbackend
    back = refbackend.reference()
else:
    back = _backend.native()
cfg = CONFIGS[name]
dev = torch.device("cuda:0")
scene = make_scene(cfg["P"], kind=kind).to(dev)
cam = make_camera(cfg["width"], cfg["height"], cfg["fovx"]).to(dev)
bg = torch.zeros(3, device=dev)
grads = tuple(g.to(dev) for g in make_pixel_grads(cfg["width"], cfg["height"]))
e = torch.empty(0)
for _ in range(steps):
    out = util.raw_forward(back, scene, cam, bg)
    back.rasterize_gaussians_backward(
        bg, scene.means3D, out["radii"], e, scene.scales, scene.rotations, 1.0, e, cam.world_view_transform,
        cam.full_proj_transform, cam.tanfovx, cam.tanfovy, grads[0], grads[1], grads[2], scene.shs, scene.sh_degree,
        cam.camera_center, out["geom"], out["R"], out["binning"], out["img"], out["alpha"], False)
torch.cuda.synchronize()
print("done", which, name, steps, "R", out["R"])
