"""Golden vectors for distCUDA2 FROM THE REFERENCE'S OWN KERNEL.

simple-knn is CUDA-only, so this script runs on the GPU box (it is how tests/golden/knn_*.npz
were made): it drives oracle/_ref/libknn_ref.so — submodules/simple-knn/simple_knn.cu compiled
unmodified by oracle/Makefile — over seeded point clouds and stores the outputs.

    gpurun -- python tests/golden/make_knn_golden.py gpurun_out/      # then copy knn_*.npz here

Inputs are regenerated from the stored seed by make_points().
"""
import ctypes
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CASES = {
    "knn_uniform": dict(P=5000, kind="uniform", seed=41),
    "knn_clustered": dict(P=4099, kind="clustered", seed=42),     # SfM-like: dense blobs + outliers
    "knn_duplicates": dict(P=1500, kind="duplicates", seed=43),   # repeated points -> zero distances
    "knn_planar": dict(P=2048, kind="planar", seed=44),           # degenerate z extent (Morton axis collapses)
    "knn_tiny": dict(P=3, kind="uniform", seed=45),               # fewer than 3 neighbours -> FLT_MAX / 3
    "knn_leaf_tail": dict(P=33, kind="uniform", seed=46),         # last leaf holds a single point
}


def make_points(P, kind, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "uniform":
        pts = torch.rand(P, 3, generator=g) * 2.6 - 1.3
    elif kind == "clustered":
        centres = torch.randn(12, 3, generator=g) * 2.0
        which = torch.randint(0, 12, (P,), generator=g)
        pts = centres[which] + torch.randn(P, 3, generator=g) * (0.02 + 0.2 * torch.rand(P, 1, generator=g))
        pts[: P // 50] = torch.randn(P // 50, 3, generator=g) * 30.0
    elif kind == "duplicates":
        base = torch.rand(P // 3, 3, generator=g)
        pts = base[torch.randint(0, P // 3, (P,), generator=g)]
    elif kind == "planar":
        pts = torch.rand(P, 3, generator=g)
        pts[:, 2] = 0.25
    else:
        raise ValueError(kind)
    return pts.float().contiguous()


def reference_dist_cuda2(points_cuda):
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libknn_ref.so"))
    lib.knn_ref_dist_cuda2.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.knn_ref_dist_cuda2.restype = ctypes.c_int
    out = torch.empty(points_cuda.shape[0], dtype=torch.float32, device=points_cuda.device)
    torch.cuda.synchronize()
    rc = lib.knn_ref_dist_cuda2(points_cuda.shape[0], points_cuda.data_ptr(), out.data_ptr())
    assert rc == 0, rc
    return out


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    for name, c in CASES.items():
        pts = make_points(**c).cuda()
        d = reference_dist_cuda2(pts).cpu().numpy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), mean_dist2=d)
        print(name, d[:4], "inf:", int(np.isinf(d).sum()))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else HERE)
