"""Run a few native training iterations (tools/bench_iteration.py) for ncu launch lists.
usage: python tools/iter_once.py <config> <iters>"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_iteration  # noqa: E402

name, iters = sys.argv[1], int(sys.argv[2])
print(bench_iteration.measure(torch.device("cuda:0"), name, iters, 3, which=("native",)))
