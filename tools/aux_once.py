"""Launch each auxiliary kernel family a few times (for ncu captures of the §8(f) rows)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_knn  # noqa: E402
import bench_loss  # noqa: E402
import bench_params  # noqa: E402

dev = torch.device("cuda:0")
print(bench_loss.measure(dev, 3, 756, 1008, iters=2, warmup=1))
print(bench_loss.measure_binocular(dev, iters=2, warmup=1))
print(bench_params.measure(dev, 300_000, 4, iters=2, warmup=1))
print(bench_knn.measure(1_000_000, "clustered", 1))
