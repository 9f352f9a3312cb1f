"""Generate the golden fixtures in this directory FROM THE REFERENCE'S OWN CUDA KERNELS.

Run on a GPU box (the reference rasterizer is CUDA-only and cannot run in the build
container):

    gpurun -- 'python tests/golden/make_golden.py'      # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/

Each fixture records a seeded synthetic case (regenerated from the stored parameters by
``workloads``, so inputs are not stored) and the outputs of
oracle/_ref/libdgr_ref.so — the unmodified reference kernels compiled from
/root/reference — for that case: radii, depth bits, tiles_touched, the sorted
point_list, per-tile ranges, n_contrib, the three images and all gradients.  The CPU
oracle is checked against these in tests/test_oracle_golden.py (CPU) and the CUDA
library in tests/test_gpu_parity.py (GPU).  Gradients from the reference are summed with
float atomics, so two reference runs are stored-as-one plus their observed spread.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from workloads import make_camera, make_pixel_grads, make_scene  # noqa: E402
from oracle import refbackend  # noqa: E402
import util  # noqa: E402

CASES = {
    # name: scene kwargs, camera kwargs, extras
    "cube_sh1": dict(scene=dict(P=3000, seed=11, kind="cube", sh_degree=1), cam=dict(width=160, height=128),
                     bg=[0.0, 0.0, 0.0]),
    "shell_sh3_ragged": dict(scene=dict(P=2500, seed=12, kind="shell", sh_degree=3),
                             cam=dict(width=150, height=100, azimuth=1.1, elevation=-0.4), bg=[0.2, 0.5, 0.7]),
    "cube_sh0of3_big": dict(scene=dict(P=1500, seed=13, kind="cube", sh_degree=0, max_sh_degree=3,
                                       scale_lo=0.02, scale_hi=0.4),
                            cam=dict(width=96, height=80, distance=2.0), bg=[1.0, 1.0, 1.0]),
    "cube_inside": dict(scene=dict(P=3000, seed=14, kind="cube", sh_degree=2),
                        cam=dict(width=128, height=128, distance=0.5, fovx=1.2), bg=[0.0, 0.3, 0.0]),
    "cube_scale_mod": dict(scene=dict(P=2000, seed=15, kind="cube", sh_degree=1), cam=dict(width=112, height=64),
                           bg=[0.0, 0.0, 0.0], scale_modifier=1.7),
}


def run_case(name, spec, dev, back):
    scene = make_scene(**spec["scene"]).to(dev)
    cam = make_camera(**spec["cam"]).to(dev)
    W, H, P = cam.image_width, cam.image_height, scene.P
    bg = torch.tensor(spec["bg"], device=dev)
    sm = spec.get("scale_modifier", 1.0)
    grads = tuple(g.to(dev) for g in make_pixel_grads(W, H, seed=spec["scene"]["seed"] + 100))
    fwd = util.raw_forward(back, scene, cam, bg, scale_modifier=sm)
    ins = util.internals(back, fwd, P, W, H)
    g1 = util.surface_forward_backward(back, scene, cam, bg, grads, scale_modifier=sm)
    g2 = util.surface_forward_backward(back, scene, cam, bg, grads, scale_modifier=sm)
    out = dict(
        R=np.int64(fwd["R"]), radii=fwd["radii"].cpu().numpy(), depth_bits=ins["depths"].cpu().numpy(),
        tiles_touched=ins["tiles_touched"].cpu().numpy(), point_list=ins["point_list"].cpu().numpy(),
        ranges=ins["ranges"].cpu().numpy(), n_contrib=ins["n_contrib"].cpu().numpy(),
        means2D=ins["means2D"].cpu().numpy(), conic_opacity=ins["conic_opacity"].cpu().numpy(),
        rgb=ins["rgb"].cpu().numpy(), color=fwd["color"].cpu().numpy(), depth=fwd["depth"].cpu().numpy(),
        alpha=fwd["alpha"].cpu().numpy(),
    )
    spread = {}
    for k in g1:
        if k.startswith("g_"):
            out[k] = g1[k].cpu().numpy()
            spread[k] = util.rel_err(g2[k], g1[k])
    out["grad_spread"] = np.array([spread[k] for k in sorted(spread)], np.float64)
    return out


def main():
    dev = torch.device("cuda:0")
    back = refbackend.reference()
    outdir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    for name, spec in CASES.items():
        out = run_case(name, spec, dev, back)
        path = os.path.join(outdir, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "R", int(out["R"]), "visible", int((out["radii"] > 0).sum()), "bytes", os.path.getsize(path),
              "grad spread max", float(out["grad_spread"].max()))


if __name__ == "__main__":
    main()
