"""Forward-only timing (CUDA events) through both host sides, against the sum of the kernel
times reported by the per-stage timers: how much of the forward is not kernel time."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from binocular3dgs_b200 import _backend  # noqa: E402
from workloads import CONFIGS, make_camera, make_scene  # noqa: E402

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "lego"
cfg = CONFIGS[name]
scene = make_scene(cfg["P"]).to(dev)
cam = make_camera(cfg["width"], cfg["height"], cfg["fovx"]).to(dev)
bg = torch.zeros(3, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(back, iters=40):
    ts = []
    for i in range(iters + 10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        util.raw_forward(back, scene, cam, bg)
        b.record()
        torch.cuda.synchronize()
        if i >= 10:
            ts.append(a.elapsed_time(b))
    return float(np.median(ts))


nat, comp = _backend.native(), _backend.preferred()
print("forward, ctypes host   : %.1f us" % (timeit(nat) * 1e3))
print("forward, compiled host : %.1f us" % (timeit(comp) * 1e3))
nat.profile_enable(True)
nat.profile_read()
for _ in range(20):
    flush.zero_()
    util.raw_forward(nat, scene, cam, bg)
torch.cuda.synchronize()
prof = nat.profile_read()
nat.profile_enable(False)
print({k: round(v[0] / max(v[1], 1) * 1e3, 1) for k, v in prof.items() if v[1]})
