"""Independent checks of pieces of the CPU oracle: SH against the
textbook real-SH polynomials (the reference's Python twin utils/sh_utils.py:57-112),
cov3D against R S S^T R^T (utils/general_utils.py:64-110), a hand-computed pinhole
projection, tile-rect / key packing / getHigherMsb known answers, the sort contract as
a hypothesis property, and the analytic backward against central finite differences."""
import math

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from workloads import make_camera, make_scene
from oracle import cpu_oracle as orc

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]


def eval_sh64(deg, sh, d):
    """float64 restatement of utils/sh_utils.py:57-112 for one direction; sh is (M,3)."""
    x, y, z = d
    r = C0 * sh[0]
    if deg > 0:
        r = r - C1 * y * sh[1] + C1 * z * sh[2] - C1 * x * sh[3]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        r = r + C2[0] * xy * sh[4] + C2[1] * yz * sh[5] + C2[2] * (2 * zz - xx - yy) * sh[6] + C2[3] * xz * sh[7] + C2[4] * (xx - yy) * sh[8]
    if deg > 2:
        r = (r + C3[0] * y * (3 * xx - yy) * sh[9] + C3[1] * xy * z * sh[10] + C3[2] * y * (4 * zz - xx - yy) * sh[11]
             + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[12] + C3[4] * x * (4 * zz - xx - yy) * sh[13]
             + C3[5] * z * (xx - yy) * sh[14] + C3[6] * x * (xx - 3 * yy) * sh[15])
    return r


def _pre(scene, cam, **kw):
    a = [t.numpy() for t in scene.tensors()]
    return orc.preprocess(a[0], a[1], a[2], a[3], a[4], cam.world_view_transform.numpy(),
                          cam.full_proj_transform.numpy(), cam.camera_center.numpy(), cam.image_width,
                          cam.image_height, cam.tanfovx, cam.tanfovy, scene.sh_degree, **kw)


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_sh_matches_eval_sh(deg):
    scene = make_scene(400, seed=deg, sh_degree=deg, max_sh_degree=3)
    cam = make_camera(160, 120)
    pre = _pre(scene, cam)
    vis = pre["radii"] > 0
    m, sh, campos = scene.means3D.numpy().astype(np.float64), scene.shs.numpy().astype(np.float64), cam.camera_center.numpy().astype(np.float64)
    for i in np.nonzero(vis)[0][:200]:
        d = m[i] - campos
        d /= np.linalg.norm(d)
        want = np.maximum(eval_sh64(deg, sh[i], d) + 0.5, 0.0)
        assert np.abs(pre["rgb"][i] - want).max() < 2e-6
        assert ((eval_sh64(deg, sh[i], d) + 0.5 < -1e-6) <= pre["clamped"][i].astype(bool)).all()


def test_cov3d_matches_rotation_scaling_product():
    scene = make_scene(300, seed=5)
    cam = make_camera(128, 128)
    for mod in (1.0, 1.7):
        pre = _pre(scene, cam, scale_modifier=mod)
        q, s = scene.rotations.numpy().astype(np.float64), scene.scales.numpy().astype(np.float64) * mod
        for i in np.nonzero(pre["radii"] > 0)[0][:100]:
            r, x, y, z = q[i]
            R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)],
                          [2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)],
                          [2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]])
            L = R @ np.diag(s[i])
            S = L @ L.T
            want = np.array([S[0, 0], S[0, 1], S[0, 2], S[1, 1], S[1, 2], S[2, 2]])
            assert np.abs(pre["cov3D"][i] - want).max() < 1e-6 * max(1.0, np.abs(want).max())


def test_pinhole_projection_by_hand():
    # identity camera at the origin looking down +z: pixel = ((x/z * f) + W/2) - 0.5
    W, H, fov = 64, 48, 0.9
    V = np.eye(4, dtype=np.float32)
    from workloads import _projection
    fovy = 2 * math.atan(H / (2 * (W / (2 * math.tan(fov / 2)))))
    Pm = (V @ _projection(0.01, 100.0, fov, fovy).numpy().T).astype(np.float32)  # transposed convention
    means = np.array([[0.0, 0.0, 2.0], [0.3, -0.2, 3.0], [0.0, 0.0, 0.1]], np.float32)
    pre = orc.preprocess(means, np.full((3, 3), 0.05, np.float32), np.tile([1, 0, 0, 0], (3, 1)).astype(np.float32),
                         np.ones((3, 1), np.float32), None, V, Pm, np.zeros(3, np.float32), W, H,
                         math.tan(fov / 2), math.tan(fovy / 2), 0, colors_precomp=np.ones((3, 3), np.float32))
    f = W / (2 * math.tan(fov / 2))
    assert pre["radii"][2] == 0 and pre["tiles_touched"][2] == 0          # behind the 0.2 near plane
    assert abs(pre["means2D"][0, 0] - (W / 2 - 0.5)) < 1e-4 and abs(pre["means2D"][0, 1] - (H / 2 - 0.5)) < 1e-4
    assert abs(pre["means2D"][1, 0] - (0.3 / 3.0 * f + W / 2 - 0.5)) < 1e-3
    assert abs(pre["means2D"][1, 1] - (-0.2 / 3.0 * f + H / 2 - 0.5)) < 1e-3
    assert pre["depths"][0] == 2.0 and pre["depths"][1] == 3.0
    # isotropic sigma=0.05 at z=2: cov2D = (f*0.05/2)^2 + 0.3 on the diagonal
    sig2 = (f * 0.05 / 2.0) ** 2 + 0.3
    assert abs(1.0 / pre["conic_opacity"][0, 0] - sig2) < 1e-3 * sig2
    assert pre["radii"][0] == math.ceil(3 * math.sqrt(sig2))


def test_higher_msb_known_answers():
    # rasterizer_impl.cu:35-50 on the four BASELINE tile counts -> 42/44/44/45 sort bits
    assert [orc.higher_msb(n) for n in (625, 2500, 3024, 7500)] == [10, 12, 12, 13]
    assert orc.higher_msb(1) == 1 and orc.higher_msb(255) == 8 and orc.higher_msb(256) == 9


def test_tile_rect_and_key_packing_known_answers():
    # one Gaussian at pixel (40.2, 17.9) radius 10 on a 5x4 tile grid: x tiles 1..3, y tiles 0..1
    means2D = np.array([[40.2, 17.9]], np.float32)
    out = orc.binning(means2D, np.array([1.5], np.float32), np.array([10], np.int32), 80, 64, want_keys=True)
    assert out["R"] == 6
    dbits = int(np.array([1.5], np.float32).view(np.uint32)[0])
    want_tiles = [0 * 5 + 1, 0 * 5 + 2, 0 * 5 + 3, 1 * 5 + 1, 1 * 5 + 2, 1 * 5 + 3]
    assert [int(k >> np.uint64(32)) for k in out["keys"]] == want_tiles
    assert all(int(k & np.uint64(0xFFFFFFFF)) == dbits for k in out["keys"])
    r = out["ranges"]
    assert (r[1] == [0, 1]).all() and (r[3] == [2, 3]).all() and (r[8] == [5, 6]).all() and (r[0] == [0, 0]).all()
    # clamping: far off-screen to the left with a huge radius still clamps to the grid
    out = orc.binning(np.array([[-500.0, 10.0]], np.float32), np.array([1.0], np.float32), np.array([600], np.int32), 80, 64)
    assert out["R"] == 5 * 4


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 300), st.integers(0, 2 ** 31 - 1), st.sampled_from([(64, 48), (80, 80), (150, 33)]))
def test_sort_contract_property(P, seed, wh):
    """Output is a permutation ordered by (tile, depth bits, Gaussian index) and ranges
    partition it (rasterizer_impl.cu:304-309 stable sort + :116-138)."""
    W, H = wh
    rng = np.random.default_rng(seed)
    means2D = (rng.random((P, 2)) * [W * 1.4, H * 1.4] - [W * 0.2, H * 0.2]).astype(np.float32)
    depths = rng.choice(np.array([0.3, 0.5, 1.0, 2.0, 7.5], np.float32), P)  # many ties
    radii = rng.integers(0, 40, P).astype(np.int32)
    out = orc.binning(means2D, depths, radii, W, H, want_keys=True)
    R, pl, keys, ranges = out["R"], out["point_list"], out["keys"], out["ranges"].astype(np.int64)
    gx = (W + 15) // 16
    assert (radii[pl] > 0).all()
    trip = np.stack([(keys >> np.uint64(32)).astype(np.int64), (keys & np.uint64(0xFFFFFFFF)).astype(np.int64),
                     pl.astype(np.int64)], 1)
    assert (np.lexsort((trip[:, 2], trip[:, 1], trip[:, 0])) == np.arange(R)).all()
    assert (trip[:, 1] == depths.view(np.uint32)[pl]).all()
    lens = ranges[:, 1] - ranges[:, 0]
    assert lens.sum() == R and (lens >= 0).all()
    nz = np.nonzero(lens)[0]
    assert (ranges[nz[1:], 0] == ranges[nz[:-1], 1]).all() if len(nz) > 1 else True
    for t in nz[:20]:
        assert (trip[ranges[t, 0]:ranges[t, 1], 0] == t).all()
    counts = np.bincount(pl, minlength=P)
    x0 = np.clip(np.trunc((means2D[:, 0] - radii) / 16), 0, gx)
    assert (counts[radii == 0] == 0).all() and counts.sum() == R and x0.min() >= 0


def _loss_and_grads(scene_arrays, cam, wts, W, H, deg):
    m, s, q, o, sh = scene_arrays
    cam_args = (cam.world_view_transform.numpy(), cam.full_proj_transform.numpy(), cam.camera_center.numpy())
    bg = np.array([0.3, 0.1, 0.2], np.float32)
    f = orc.rasterize_forward(m, s, q, o, sh, *cam_args, bg, W, H, cam.tanfovx, cam.tanfovy, deg)
    loss = float((f["color"].astype(np.float64) * wts[0]).sum() + (f["depth"].astype(np.float64) * wts[1]).sum()
                 + (f["alpha"].astype(np.float64) * wts[2]).sum())
    return loss, f, cam_args, bg


def test_backward_matches_finite_differences():
    """Analytic backward (oracle restatement of backward.cu) vs central differences of
    the oracle forward on a tiny scene, loss = <w, color> + <w, depth> + <w, alpha>."""
    W, H, deg = 32, 32, 2
    scene = make_scene(24, seed=21, sh_degree=deg, scale_lo=0.08, scale_hi=0.25)
    cam = make_camera(W, H, distance=3.0)
    arrs = [t.numpy().astype(np.float32).copy() for t in scene.tensors()]
    rng = np.random.default_rng(0)
    wts = [rng.standard_normal((3, H, W)), rng.standard_normal((1, H, W)) * 0.3, rng.standard_normal((1, H, W))]
    _, f, cam_args, bg = _loss_and_grads(arrs, cam, wts, W, H, deg)
    b = orc.rasterize_backward(f, arrs[0], arrs[1], arrs[2], arrs[4], *cam_args, bg, W, H, cam.tanfovx, cam.tanfovy,
                               deg, *(w.astype(np.float32) for w in wts))
    names = ["dL_dmeans3D", "dL_dscales", "dL_drotations", "dL_dopacity", "dL_dsh"]
    # The forward is piecewise smooth (alpha >= 1/255 cut-off, T < 1e-4 stop, tile
    # rectangles), and the reference's backward — which the oracle restates — ignores the
    # jump terms.  A central difference that straddles a jump is wrong by jump/eps, so
    # each entry is probed at three step sizes and judged on the best one.
    base_eps = [4e-3, 2e-3, 4e-3, 4e-3, 1e-2]
    errs = []
    for ai, (nm, e0) in enumerate(zip(names, base_eps)):
        g = b[nm].reshape(arrs[ai].shape)
        scale = np.abs(g).max()
        for idx in np.argsort(-np.abs(g).reshape(-1))[:6]:       # the 6 largest entries of each tensor
            ix = np.unravel_index(idx, g.shape)
            keep = arrs[ai][ix]
            best = np.inf
            for e in (e0, e0 / 4, e0 / 16):
                arrs[ai][ix] = keep + e
                lp = _loss_and_grads(arrs, cam, wts, W, H, deg)[0]
                arrs[ai][ix] = keep - e
                lm = _loss_and_grads(arrs, cam, wts, W, H, deg)[0]
                arrs[ai][ix] = keep
                best = min(best, abs((lp - lm) / (2 * e) - g[ix]) / scale)
            errs.append(best)
    errs = np.array(errs)
    assert len(errs) == 30
    assert (errs <= 0.01).sum() >= 27, errs
    assert errs.max() <= 0.2, errs
