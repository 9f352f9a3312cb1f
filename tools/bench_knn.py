"""Timing of distCUDA2: this library vs the reference's own kernel
(oracle/_ref/libknn_ref.so, unmodified simple_knn.cu), same GPU, same points; wall time with
synchronisation on both sides (the reference allocates, synchronises and copies to the host
internally, so CUDA events on a stream would not cover it)."""
import importlib.util
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("mkg", os.path.join(ROOT, "tests", "golden", "make_knn_golden.py"))
mkg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mkg)


def measure(P=1_000_000, kind="clustered", iters=5, which="both"):
    """which: "native" (this library only; nothing under oracle/ is touched), "reference" (the
    reference's kernel only) or "both" (A/B with a bit-identity check)."""
    from simple_knn._C import distCUDA2
    pts = mkg.make_points(P, kind, 1).cuda()
    have_ref = which != "native" and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libknn_ref.so"))

    def wall(fn):
        best = 1e9
        for _ in range(iters):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = fn(pts)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return best * 1e3, out
    res = {"what": "distCUDA2, %d points (%s)" % (P, kind)}
    if which != "reference":
        distCUDA2(pts)
        t_new, a = wall(distCUDA2)
        res["native_ms"] = round(t_new, 3)
    if have_ref:
        mkg.reference_dist_cuda2(pts)
        t_ref, b = wall(mkg.reference_dist_cuda2)
        res["reference_kernel_ms"] = round(t_ref, 3)
        if which == "both":
            res.update(speedup=round(t_ref / t_new, 2), bit_identical=bool(torch.equal(a, b)))
    return res


if __name__ == "__main__":
    for P, kind in ((100_000, "clustered"), (1_000_000, "clustered"), (1_000_000, "uniform")):
        print(json.dumps(measure(P, kind)))
