"""TEST INFRASTRUCTURE — CPU restatement of simple-knn's distCUDA2
(submodules/simple-knn/simple_knn.cu:132-183 updateKBest / boxMeanDist, :185-221 knn).

The reference's Morton boxes only prune the search; what it returns for every point is the
mean of the squared distances to its 3 exact nearest neighbours, each evaluated in float32 as
    d = neighbour - query;  dist = fma(d.z, d.z, fma(d.x, d.x, d.y * d.y))
(the contraction nvcc emits for `d.x*d.x + d.y*d.y + d.z*d.z` — the middle product is
rounded alone — read from the SASS of the reference compiled for sm_100a) and combined as ((b0 + b1) + b2) / 3.0f, with FLT_MAX for
missing neighbours when P < 4.  Here: candidates from a float64 k-d tree (scipy), the
float32 expression emulated through float64 (products of two float32 are exact in float64;
the fused add is rounded once to float64 and once to float32 — double rounding can differ
from a true FMA by 1 ulp in rare cases, hence the 1-ulp allowance in the tests).
Pinned against outputs of the reference's own kernel run on a B200
(tests/golden/make_knn_golden.py -> tests/golden/knn_*.npz).
"""
import numpy as np

FLT_MAX = np.float32(3.4028234663852886e38)


def _fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def _dist32(q, p):
    """q (..., 3), p (..., 3) float32 -> float32 squared distance, reference expression."""
    d = (p - q).astype(np.float32)
    t = (d[..., 1] * d[..., 1]).astype(np.float32)
    return _fma32(d[..., 2], d[..., 2], _fma32(d[..., 0], d[..., 0], t))


def dist_cuda2(points, k_candidates=12):
    points = np.ascontiguousarray(points, dtype=np.float32)
    P = points.shape[0]
    best = np.full((P, 3), FLT_MAX, dtype=np.float32)
    if P > 1:
        if P <= 2048:
            idx = np.broadcast_to(np.arange(P)[None, :], (P, P))
        else:
            from scipy.spatial import cKDTree
            _, idx = cKDTree(points.astype(np.float64)).query(points.astype(np.float64), k=min(P, k_candidates))
        d = _dist32(points[:, None, :], points[idx])
        # exclude the query itself (exactly one occurrence: duplicates of it DO count)
        self_col = np.argmax(idx == np.arange(P)[:, None], axis=1)
        has_self = (idx == np.arange(P)[:, None]).any(axis=1)
        d = d.copy()
        d[np.arange(P)[has_self], self_col[has_self]] = np.inf
        d.sort(axis=1)
        k = min(3, d.shape[1])
        take = d[:, :k]
        best[:, :k] = np.where(np.isfinite(take), take, FLT_MAX)
    with np.errstate(over="ignore"):
        s = ((best[:, 0] + best[:, 1]).astype(np.float32) + best[:, 2]).astype(np.float32)
        return (s / np.float32(3.0)).astype(np.float32)
