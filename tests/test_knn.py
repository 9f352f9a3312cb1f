"""distCUDA2.

CPU: the oracle (oracle/knn_oracle.py) against golden vectors produced by the reference's
own kernel on a B200 (tests/golden/make_knn_golden.py), to 1 ulp (the oracle emulates the
float32 FMA through float64), and against a float64 brute force.
GPU (-m gpu): b3gs_dist_cuda2 through the drop-in ``simple_knn._C.distCUDA2`` against the
same golden vectors BIT-EXACT, against the reference kernel live (oracle/_ref/libknn_ref.so)
bit-exact at 100k / 1M points, and against the oracle."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import knn_oracle as ko

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_knn_golden", os.path.join(HERE, "golden", "make_knn_golden.py"))
mkg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mkg)
NAMES = sorted(mkg.CASES)


def ulp_diff(a, b):
    """max distance in float32 ulps; identical infinities count as 0."""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    return int(np.where(same_inf, 0, np.abs(ia - ib)).max()) if a.size else 0


def golden(name):
    path = os.path.join(HERE, "golden", name + ".npz")
    if not os.path.exists(path):
        pytest.skip("golden vector not generated yet (needs the GPU box)")
    return np.load(path)["mean_dist2"]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_golden(name):
    want = golden(name)
    got = ko.dist_cuda2(mkg.make_points(**mkg.CASES[name]).numpy())
    assert got.shape == want.shape
    assert ulp_diff(got, want) <= 1


def test_oracle_vs_float64_brute_force():
    pts = mkg.make_points(700, "clustered", 3).numpy()
    d2 = ((pts[:, None, :].astype(np.float64) - pts[None, :, :].astype(np.float64)) ** 2).sum(-1)
    np.fill_diagonal(d2, np.inf)
    want = np.sort(d2, axis=1)[:, :3].mean(1)
    got = ko.dist_cuda2(pts).astype(np.float64)
    assert np.abs(got - want).max() <= 1e-6 * want.max()
    # the k-d tree path (P > 2048) agrees with the brute-force path
    pts = mkg.make_points(3000, "uniform", 4).numpy()
    d2 = ((pts[:, None, :].astype(np.float64) - pts[None, :, :].astype(np.float64)) ** 2).sum(-1)
    np.fill_diagonal(d2, np.inf)
    want = np.sort(d2, axis=1)[:, :3].mean(1)
    assert np.abs(ko.dist_cuda2(pts).astype(np.float64) - want).max() <= 1e-6 * want.max()


def test_oracle_edge_cases():
    assert np.isinf(ko.dist_cuda2(np.zeros((1, 3), np.float32))).all()
    # P = 2: two of the three "neighbours" are FLT_MAX -> the sum overflows; P = 3: one is,
    # FLT_MAX absorbs the two real distances and the mean is FLT_MAX / 3
    assert np.isinf(ko.dist_cuda2(np.random.default_rng(0).random((2, 3)).astype(np.float32))).all()
    assert (ko.dist_cuda2(np.random.default_rng(0).random((3, 3)).astype(np.float32)) == ko.FLT_MAX / np.float32(3)).all()
    four = np.array([[0, 0, 0], [1, 0, 0], [0, 2, 0], [0, 0, 3]], np.float32)
    assert np.allclose(ko.dist_cuda2(four), [(1 + 4 + 9) / 3, (1 + 5 + 10) / 3, (4 + 5 + 13) / 3, (9 + 10 + 13) / 3])
    assert (ko.dist_cuda2(np.ones((5, 3), np.float32)) == 0).all()


def test_api_surface_and_loud_failures():
    from simple_knn._C import distCUDA2
    with pytest.raises(RuntimeError, match="CUDA"):
        distCUDA2(torch.zeros(8, 3))
    with pytest.raises(RuntimeError):
        distCUDA2(torch.zeros(8, 4))


# ------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_kernel_matches_reference_golden_bit_exact(name):
    from simple_knn._C import distCUDA2
    want = golden(name)
    got = distCUDA2(mkg.make_points(**mkg.CASES[name]).cuda()).cpu().numpy()
    assert ulp_diff(got, want) == 0
    assert (np.isinf(got) == np.isinf(want)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("P,kind", [(100_000, "clustered"), (1_000_000, "uniform"), (200_001, "planar")])
def test_kernel_matches_reference_kernel_live_bit_exact(P, kind):
    if not os.path.exists(os.path.join(mkg.ROOT, "oracle", "_ref", "libknn_ref.so")):
        pytest.skip("oracle/_ref/libknn_ref.so not built")
    from simple_knn._C import distCUDA2
    pts = mkg.make_points(P, kind, 7).cuda()
    want = mkg.reference_dist_cuda2(pts)
    got = distCUDA2(pts)
    torch.cuda.synchronize()
    assert torch.equal(got, want)


@pytest.mark.gpu
def test_kernel_vs_oracle_and_edge_cases():
    from simple_knn._C import distCUDA2
    pts = mkg.make_points(20_000, "clustered", 9)
    assert ulp_diff(distCUDA2(pts.cuda()).cpu().numpy(), ko.dist_cuda2(pts.numpy())) <= 1
    assert distCUDA2(torch.zeros(0, 3).cuda()).shape == (0,)
    assert torch.isinf(distCUDA2(torch.rand(2, 3).cuda())).all()
    assert (distCUDA2(torch.ones(40, 3).cuda()) == 0).all()
    # non-contiguous input is legal (the reference calls .contiguous(), spatial.cu:23)
    wide = torch.rand(500, 6).cuda()
    assert torch.equal(distCUDA2(wide[:, ::2]), distCUDA2(wide[:, ::2].contiguous()))
