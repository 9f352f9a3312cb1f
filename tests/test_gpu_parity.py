"""Parity tests proper (run on the B200 with -m gpu).  Everything goes through the
C-ABI (ctypes -> libb3gs.so).  Three checkers, in decreasing strictness:

* golden fixtures produced by the reference's own CUDA kernels (tests/golden/*.npz);
* the reference's own kernels run live (oracle/_ref/libdgr_ref.so) when present;
* the CPU oracle (oracle/liboracle.so).

Bars: bit-exact for radii, depth bits, tiles_touched, R, the sorted point_list,
ranges, n_contrib; images <= 1e-5 max-abs (they are in fact bit-identical to the
reference, asserted where the reference is available); gradients within
max(6 x the reference's measured run-to-run spread, 2e-5) of tensor scale.
"""
import numpy as np
import pytest
import torch

import cases
import util
from workloads import CONFIGS, Scene, make_camera, make_pixel_grads, make_scene

pytestmark = pytest.mark.gpu
GRAD_KEYS = ("g_means3D", "g_means2D", "g_scales", "g_rotations", "g_opacities", "g_shs")


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def nat():
    from binocular3dgs_b200 import _backend
    return _backend.native()


@pytest.fixture(scope="module")
def ref():
    from oracle import refbackend
    if not refbackend.available():
        pytest.skip("oracle/_ref/libdgr_ref.so not present")
    return refbackend.reference()


def _bits(t):
    return t.contiguous().view(torch.int32)


# ------------------------------------------------------------------ golden fixtures
@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_native_vs_golden(name, nat, dev):
    g = cases.load_golden(name)
    scene, cam, bg, grads, sm = cases.make_case(name)
    scene, cam = scene.to(dev), cam.to(dev)
    bg_t = torch.from_numpy(bg).to(dev)
    W, H, P = cam.image_width, cam.image_height, scene.P
    fwd = util.raw_forward(nat, scene, cam, bg_t, scale_modifier=sm)
    ins = util.internals(nat, fwd, P, W, H)
    vis = g["radii"] > 0
    assert fwd["R"] == int(g["R"])
    assert (fwd["radii"].cpu().numpy() == g["radii"]).all()
    assert (ins["depths"].cpu().numpy()[vis] == g["depth_bits"][vis]).all()
    assert (ins["tiles_touched"].cpu().numpy() == np.where(vis, g["tiles_touched"], 0)).all()
    assert (ins["point_list"].cpu().numpy() == g["point_list"]).all()
    assert (ins["ranges"].cpu().numpy() == g["ranges"]).all()
    assert (ins["n_contrib"].cpu().numpy() == g["n_contrib"]).all()
    for k in ("means2D", "conic_opacity", "rgb"):
        assert (_bits(ins[k]).cpu().numpy()[vis] == g[k].view(np.int32)[vis]).all(), k
    for k in ("color", "depth", "alpha"):
        # same GPU arithmetic as the reference -> bit-identical, far inside the 1e-5 bar
        assert (_bits(fwd[k]).cpu().numpy() == g[k].view(np.int32)).all(), k
    out = util.surface_forward_backward(nat, scene, cam, bg_t, tuple(t.to(dev) for t in grads), scale_modifier=sm)
    spread = dict(zip(sorted(GRAD_KEYS), g["grad_spread"]))
    for k in GRAD_KEYS:
        tol = max(6 * spread[k], 2e-5)
        err = util.rel_err(out[k].cpu(), torch.from_numpy(g[k]))
        assert err <= tol, (k, err, tol)


# ------------------------------------------------------------------ live reference
@pytest.mark.parametrize("cfg,kind", [("plumbing", "cube"), ("lego", "cube"), ("lego", "shell"), ("fern", "cube"), ("dtu", "cube")])
def test_native_vs_reference_kernels(cfg, kind, nat, ref, dev):
    c = CONFIGS[cfg]
    W, H, P = c["width"], c["height"], c["P"]
    scene = make_scene(P, seed=7, kind=kind).to(dev)
    cam = make_camera(W, H, c["fovx"], azimuth=0.7).to(dev)
    bg = torch.tensor([0.3, 0.6, 0.1], device=dev)
    fn, fr = util.raw_forward(nat, scene, cam, bg), util.raw_forward(ref, scene, cam, bg)
    inn, inr = util.internals(nat, fn, P, W, H), util.internals(ref, fr, P, W, H)
    vis = fr["radii"] > 0
    assert fn["R"] == fr["R"] and (fn["radii"] == fr["radii"]).all()
    assert (inn["depths"][vis] == inr["depths"][vis]).all()
    assert (inn["tiles_touched"][vis] == inr["tiles_touched"][vis]).all()
    assert (inn["point_list"] == inr["point_list"]).all()
    assert (inn["ranges"] == inr["ranges"]).all()
    assert (inn["n_contrib"] == inr["n_contrib"]).all()
    for k in ("color", "depth", "alpha"):
        assert util.max_abs(fn[k], fr[k]) <= 1e-5
        assert (_bits(fn[k]) == _bits(fr[k])).all(), k
    grads = tuple(t.to(dev) for t in make_pixel_grads(W, H, 8))
    gn = util.surface_forward_backward(nat, scene, cam, bg, grads)
    gr = util.surface_forward_backward(ref, scene, cam, bg, grads)
    gr2 = util.surface_forward_backward(ref, scene, cam, bg, grads)
    def outside(g):
        return [(k, util.rel_err(g[k], gr[k]), max(6 * util.rel_err(gr2[k], gr[k]), 2e-5)) for k in GRAD_KEYS
                if util.rel_err(g[k], gr[k]) > max(6 * util.rel_err(gr2[k], gr[k]), 2e-5)]

    bad = outside(gn)
    if bad:   # float atomics on both sides: every run is one draw; a real error is outside the bar twice
        bad = outside(util.surface_forward_backward(nat, scene, cam, bg, grads))
    assert not bad, bad


# ------------------------------------------------------------------ CPU oracle
def _oracle(scene_c, cam_c, bg, grads_c, **kw):
    from oracle import cpu_oracle as orc
    a = [t.numpy() for t in scene_c.tensors()]
    cam_args = (cam_c.world_view_transform.numpy(), cam_c.full_proj_transform.numpy(), cam_c.camera_center.numpy())
    of = orc.rasterize_forward(a[0], a[1], a[2], a[3], a[4], *cam_args, bg, cam_c.image_width, cam_c.image_height,
                               cam_c.tanfovx, cam_c.tanfovy, scene_c.sh_degree, **kw)
    return orc, a, cam_args, of


@pytest.mark.parametrize("P,W,H,kind,deg", [(20000, 400, 400, "cube", 1), (30000, 333, 217, "shell", 3),
                                             (5000, 1008, 756, "cube", 2)])
def test_native_vs_cpu_oracle(P, W, H, kind, deg, nat, dev):
    scene_c, cam_c = make_scene(P, seed=31, kind=kind, sh_degree=deg), make_camera(W, H, elevation=-0.3)
    bg = np.array([0.5, 0.5, 0.5], np.float32)
    grads_c = make_pixel_grads(W, H, 32)
    orc, a, cam_args, of = _oracle(scene_c, cam_c, bg, grads_c)
    scene, cam, bg_t = scene_c.to(dev), cam_c.to(dev), torch.from_numpy(bg).to(dev)
    fwd = util.raw_forward(nat, scene, cam, bg_t)
    ins = util.internals(nat, fwd, P, W, H)
    assert fwd["R"] == of["R"]
    assert (fwd["radii"].cpu().numpy() == of["radii"]).all()
    assert (ins["depths"].cpu().numpy() == of["depths"].view(np.int32)).all()
    assert (ins["point_list"].cpu().numpy() == of["point_list"].view(np.int32)).all()
    assert (ins["ranges"].cpu().numpy() == of["ranges"].reshape(-1).view(np.int32)).all()
    for k in ("color", "alpha"):
        assert np.abs(fwd[k].cpu().numpy() - of[k]).max() <= 1e-5
    assert np.abs(fwd["depth"].cpu().numpy() - of["depth"]).max() <= 1e-5 * max(1.0, float(of["depth"].max()))
    # backward: oracle fed with the GPU forward (alpha, n_contrib) isolates K7..K9
    of2 = dict(of)
    of2["alpha"], of2["n_contrib"] = fwd["alpha"].cpu().numpy(), ins["n_contrib"].cpu().numpy().view(np.uint32)
    ob = orc.rasterize_backward(of2, a[0], a[1], a[2], a[4], *cam_args, bg, W, H, cam_c.tanfovx, cam_c.tanfovy, deg,
                                *(t.numpy() for t in grads_c))
    out = util.surface_forward_backward(nat, scene, cam, bg_t, tuple(t.to(dev) for t in grads_c))
    pairs = dict(g_means3D="dL_dmeans3D", g_scales="dL_dscales", g_rotations="dL_drotations",
                 g_opacities="dL_dopacity", g_shs="dL_dsh", g_means2D="dL_dmean2D")
    for k, ok in pairs.items():
        assert util.rel_err(out[k].cpu(), torch.from_numpy(ob[ok]).reshape(out[k].shape)) <= 1e-4, k


# ------------------------------------------------------------------ edge cases
def _settings(nat, cam, bg, deg=1, **kw):
    return util.settings_for(cam, bg, deg, **kw)


def test_empty_input_returns_zero_images(nat, dev):
    # rasterize_points.cu:83: P == 0 skips everything; outputs stay zero (not background)
    cam = make_camera(64, 48).to(dev)
    scene = Scene(*(torch.zeros(s, device=dev) for s in ((0, 3), (0, 3), (0, 4), (0, 1), (0, 4, 3))), 1)
    out = util.raw_forward(nat, scene, cam, torch.ones(3, device=dev))
    assert out["R"] == 0 and out["radii"].numel() == 0
    assert float(out["color"].abs().max()) == 0 and float(out["alpha"].abs().max()) == 0
    S = util.make_surface(nat)
    m = torch.zeros(0, 3, device=dev, requires_grad=True)
    c, r, d, a = S.GaussianRasterizer(util.settings_for(cam, torch.ones(3, device=dev), 1))(
        means3D=m, means2D=torch.zeros(0, 3, device=dev), opacities=torch.zeros(0, 1, device=dev),
        shs=torch.zeros(0, 4, 3, device=dev), scales=torch.zeros(0, 3, device=dev), rotations=torch.zeros(0, 4, device=dev))
    c.sum().backward()
    assert m.grad.shape == (0, 3)


def test_all_culled_renders_background(nat, dev):
    # R == 0 (rasterizer_impl.cu:314): every Gaussian behind the camera
    cam = make_camera(80, 64).to(dev)
    scene = make_scene(500, seed=1).to(dev)
    scene.means3D[:] = scene.means3D * 0.01 + cam.camera_center - 3.0 * (-cam.camera_center / cam.camera_center.norm())
    bg = torch.tensor([0.25, 0.5, 0.75], device=dev)
    out = util.raw_forward(nat, scene, cam, bg)
    assert out["R"] == 0 and int(out["radii"].max()) == 0
    assert torch.equal(out["color"], bg[:, None, None].expand(3, 64, 80))
    assert float(out["alpha"].abs().max()) == 0 and float(out["depth"].abs().max()) == 0
    g = util.surface_forward_backward(nat, scene, cam, bg, tuple(t.to(dev) for t in make_pixel_grads(80, 64)))
    for k in GRAD_KEYS:
        assert float(g[k].abs().max()) == 0


def test_huge_gaussians_cover_every_tile(nat, dev):
    from oracle import cpu_oracle as orc
    W, H = 200, 120
    scene_c = make_scene(64, seed=2, scale_lo=1.0, scale_hi=3.0)
    cam_c = make_camera(W, H)
    bg = np.zeros(3, np.float32)
    _, a, cam_args, of = _oracle(scene_c, cam_c, bg, None)
    out = util.raw_forward(nat, scene_c.to(dev), cam_c.to(dev), torch.zeros(3, device=dev))
    T = ((W + 15) // 16) * ((H + 15) // 16)
    assert out["R"] == of["R"] and of["tiles_touched"].max() == T
    ins = util.internals(nat, out, 64, W, H)
    assert (ins["point_list"].cpu().numpy() == of["point_list"].view(np.int32)).all()
    assert np.abs(out["color"].cpu().numpy() - of["color"]).max() <= 1e-5


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_sh_degrees_with_16_coefficients(deg, nat, dev):
    W, H = 96, 96
    scene_c, cam_c = make_scene(3000, seed=40 + deg, sh_degree=deg, max_sh_degree=3), make_camera(W, H)
    bg = np.zeros(3, np.float32)
    _, a, cam_args, of = _oracle(scene_c, cam_c, bg, None)
    out = util.surface_forward_backward(nat, scene_c.to(dev), cam_c.to(dev), torch.zeros(3, device=dev),
                                        tuple(t.to(dev) for t in make_pixel_grads(W, H)))
    assert np.abs(out["color"].cpu().numpy() - of["color"]).max() <= 1e-5
    n = (deg + 1) ** 2
    if n < 16:
        assert float(out["g_shs"][:, n:, :].abs().max()) == 0  # untouched coefficients get exact zeros
    assert float(out["g_shs"][:, :n, :].abs().max()) > 0


def test_precomputed_colors_and_covariance(nat, dev):
    from oracle import cpu_oracle as orc
    W, H, P = 128, 96, 4000
    scene_c, cam_c = make_scene(P, seed=50), make_camera(W, H)
    bg = np.array([0.1, 0.1, 0.1], np.float32)
    _, a, cam_args, of = _oracle(scene_c, cam_c, bg, None)
    colors, cov = torch.from_numpy(of["rgb"]).to(dev), torch.from_numpy(of["cov3D"]).to(dev)
    # cov3D of culled Gaussians is zero in the oracle; give them a valid matrix
    cov[cov.abs().sum(1) == 0] = torch.tensor([1e-4, 0, 0, 1e-4, 0, 1e-4], device=dev)
    scene, cam = scene_c.to(dev), cam_c.to(dev)
    base = util.raw_forward(nat, scene, cam, torch.from_numpy(bg).to(dev))
    pre = util.raw_forward(nat, scene, cam, torch.from_numpy(bg).to(dev), colors_precomp=colors, cov3D_precomp=cov)
    assert pre["R"] == base["R"]
    for k in ("color", "depth", "alpha"):
        assert (_bits(pre[k]) == _bits(base[k])).all(), k
    # gradients flow to the precomputed inputs (…/__init__.py:146-158 return order)
    S = util.make_surface(nat)
    colors.requires_grad_(True); cov.requires_grad_(True)
    m3 = scene.means3D.clone().requires_grad_(True)
    op = scene.opacities.clone().requires_grad_(True)
    c, r, d, al = S.GaussianRasterizer(util.settings_for(cam, torch.from_numpy(bg).to(dev), 1))(
        means3D=m3, means2D=torch.zeros_like(m3, requires_grad=True), opacities=op, colors_precomp=colors,
        cov3D_precomp=cov)
    (c.sum() + d.sum() + al.sum()).backward()
    assert colors.grad.shape == (P, 3) and cov.grad.shape == (P, 6)
    assert float(colors.grad.abs().max()) > 0 and float(cov.grad.abs().max()) > 0 and float(m3.grad.abs().max()) > 0


def test_non_contiguous_inputs_and_side_stream(nat, dev):
    W, H, P = 112, 80, 3000
    scene, cam = make_scene(P, seed=60).to(dev), make_camera(W, H).to(dev)
    bg = torch.zeros(3, device=dev)
    base = util.raw_forward(nat, scene, cam, bg)
    wide = torch.zeros(P, 6, device=dev)
    wide[:, ::2] = scene.means3D
    nc = Scene(wide[:, ::2], scene.scales.t().contiguous().t(), scene.rotations, scene.opacities, scene.shs, 1)
    assert not nc.means3D.is_contiguous()
    s = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s):
        out = util.raw_forward(nat, nc, cam, bg)
    s.synchronize()
    assert out["R"] == base["R"]
    for k in ("color", "depth", "alpha"):
        assert (_bits(out[k]) == _bits(base[k])).all()


def test_two_forwards_in_flight_before_backward(nat, dev):
    # train.py:100,128,149: binocular pair = two forwards, then one backward through both
    W, H, P = 128, 96, 5000
    scene = make_scene(P, seed=70).to(dev)
    cam1, cam2 = make_camera(W, H).to(dev), make_camera(W, H, shift_x=0.2).to(dev)
    bg = torch.zeros(3, device=dev)
    grads = tuple(t.to(dev) for t in make_pixel_grads(W, H))
    S = util.make_surface(nat)
    leaves = [t.clone().requires_grad_(True) for t in scene.tensors()]

    def render(cam):
        m, s, q, o, sh = leaves
        return S.GaussianRasterizer(util.settings_for(cam, bg, 1))(
            means3D=m, means2D=torch.zeros_like(m, requires_grad=True), opacities=o, shs=sh, scales=s, rotations=q)

    c1, _, d1, a1 = render(cam1)
    c2, _, d2, a2 = render(cam2)
    torch.autograd.backward([c1, d1, a1, c2], [grads[0], grads[1], grads[2], grads[0]])
    both = [l.grad.clone() for l in leaves]
    g1 = util.surface_forward_backward(nat, scene, cam1, bg, grads)
    g2 = util.surface_forward_backward(nat, scene, cam2, bg, (grads[0], torch.zeros_like(grads[1]), torch.zeros_like(grads[2])))
    for b, k in zip(both, ("g_means3D", "g_scales", "g_rotations", "g_opacities", "g_shs")):
        assert util.rel_err(b, g1[k] + g2[k]) <= 2e-5, k


def test_debug_mode_and_mark_visible(nat, dev, tmp_path, monkeypatch):
    W, H, P = 64, 64, 2000
    scene, cam = make_scene(P, seed=80).to(dev), make_camera(W, H).to(dev)
    bg = torch.zeros(3, device=dev)
    a = util.raw_forward(nat, scene, cam, bg, debug=False)
    b = util.raw_forward(nat, scene, cam, bg, debug=True)
    assert (_bits(a["color"]) == _bits(b["color"])).all()
    S = util.make_surface(nat)
    vis = S.GaussianRasterizer(util.settings_for(cam, bg, 1)).markVisible(scene.means3D)
    z = (torch.cat([scene.means3D, torch.ones(P, 1, device=dev)], 1) @ cam.world_view_transform)[:, 2]
    assert vis.dtype == torch.bool and bool((vis == (z > 0.2)).all())
    # debug snapshot on failure (…/__init__.py:83-90): a bad SH tensor makes the C call fail
    monkeypatch.chdir(tmp_path)
    bad = util.settings_for(cam, bg, 3, debug=True)
    with pytest.raises(RuntimeError):
        S.GaussianRasterizer(bad)(means3D=scene.means3D, means2D=torch.zeros_like(scene.means3D),
                                   opacities=scene.opacities, shs=scene.shs, scales=scene.scales,
                                   rotations=scene.rotations)   # M=4 < (3+1)^2
    assert (tmp_path / "snapshot_fw.dump").exists()


# ------------------------------------------------------------------ full-size properties
@pytest.mark.parametrize("cfg", ["lego", "dtu"])
def test_full_size_structure_and_determinism(cfg, nat, dev):
    """At BASELINE sizes the oracle is too slow for a per-element check; use
    size-independent properties: the list is sorted by (tile, depth bits, index), ranges
    partition it, sum(tiles_touched) == R, two forwards are bit-identical, and the
    backward is linear in the upstream gradients."""
    c = CONFIGS[cfg]
    W, H, P = c["width"], c["height"], c["P"]
    scene, cam = make_scene(P, seed=90).to(dev), make_camera(W, H, c["fovx"]).to(dev)
    bg = torch.zeros(3, device=dev)
    f1, f2 = util.raw_forward(nat, scene, cam, bg), util.raw_forward(nat, scene, cam, bg)
    i1 = util.internals(nat, f1, P, W, H)
    R = f1["R"]
    assert R == int(i1["tiles_touched"].long().sum()) and R > P
    for k in ("color", "depth", "alpha"):
        assert (_bits(f1[k]) == _bits(f2[k])).all()
    pl = i1["point_list"].long()
    rng = i1["ranges"].view(-1, 2).long()
    lens = rng[:, 1] - rng[:, 0]
    assert int(lens.sum()) == R
    tile_of = torch.repeat_interleave(torch.arange(rng.shape[0], device=dev), lens)
    depth_bits = i1["depths"].long()[pl]
    key = (tile_of << 32) | depth_bits
    assert bool((key[1:] >= key[:-1]).all())
    ties = key[1:] == key[:-1]
    assert bool((pl[1:][ties] > pl[:-1][ties]).all())
    assert bool((f1["radii"][pl] > 0).all())
    grads = tuple(t.to(dev) for t in make_pixel_grads(W, H))
    g1 = util.surface_forward_backward(nat, scene, cam, bg, grads)
    g2 = util.surface_forward_backward(nat, scene, cam, bg, tuple(2 * t for t in grads))
    for k in GRAD_KEYS:
        assert util.rel_err(g2[k], 2 * g1[k]) <= 2e-5, k
        assert bool(torch.isfinite(g1[k]).all())


def test_reference_render_adapter_end_to_end(nat, dev):
    """The reference's render() contract on top of our module, with a duck-typed model
    (the adapter source itself is only present in the build container)."""
    import binocular3dgs_b200 as b3
    W, H, P = 96, 64, 3000
    scene, cam = make_scene(P, seed=99).to(dev), make_camera(W, H).to(dev)
    settings = b3.GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0,
                                                cam.world_view_transform, cam.full_proj_transform, 1,
                                                cam.camera_center, False, False)
    xyz = scene.means3D.clone().requires_grad_(True)
    screenspace = torch.zeros_like(xyz, requires_grad=True) + 0
    screenspace.retain_grad()
    img, radii, depth, alpha = b3.GaussianRasterizer(settings)(
        means3D=xyz, means2D=screenspace, shs=scene.shs, colors_precomp=None, opacities=scene.opacities,
        scales=scene.scales, rotations=scene.rotations, cov3D_precomp=None)
    assert img.shape == (3, H, W) and depth.shape == (1, H, W) and alpha.shape == (1, H, W)
    assert radii.dtype == torch.int32 and radii.shape == (P,)
    (img.mean() + 0.1 * depth.mean()).backward()
    assert screenspace.grad.shape == (P, 3) and float(screenspace.grad[:, 2].abs().max()) == 0
    assert float(screenspace.grad[radii > 0][:, :2].abs().max()) > 0


@pytest.mark.gpu
def test_unused_outputs_pass_null_gradients_and_match_explicit_zeros():
    """Only the colour image feeds the loss: depth/alpha gradients travel as NULL through the
    C-ABI (no materialised zero images) and give the same result as explicit zeros."""
    from binocular3dgs_b200 import _backend
    from binocular3dgs_b200.rasterizer import make_surface
    from workloads import make_camera, make_pixel_grads, make_scene
    dev = torch.device("cuda:0")
    scene, cam = make_scene(3000, seed=12).to(dev), make_camera(120, 90).to(dev)
    bg = torch.zeros(3, device=dev)
    gc = make_pixel_grads(120, 90, 13)[0].to(dev)
    S = make_surface(_backend.native())

    def run(only_color):
        leaves = [t.detach().clone().requires_grad_(True) for t in scene.tensors()]
        m3, sc, ro, op, sh = leaves
        m2 = torch.zeros_like(m3, requires_grad=True)
        color, radii, depth, alpha = S.GaussianRasterizer(util.settings_for(cam, bg, scene.sh_degree))(
            means3D=m3, means2D=m2, opacities=op, shs=sh, scales=sc, rotations=ro)
        if only_color:
            color.backward(gc)
        else:
            torch.autograd.backward([color, depth, alpha], [gc, torch.zeros_like(depth), torch.zeros_like(alpha)])
        return [t.grad for t in leaves] + [m2.grad]

    for a, b in zip(run(True), run(False)):
        assert util.rel_err(a, b) <= 2e-5


@pytest.mark.gpu
def test_backward_kernel_shapes_agree():
    """The backward exists with 1, 2 and 4 pixels per lane (chosen per call from the
    instances-per-Gaussian ratio); all three must give the same gradients up to float
    summation order, on a ragged image (tile rows/columns cut by the border)."""
    from binocular3dgs_b200 import _backend
    nat = _backend.native()
    dev = torch.device("cuda:0")
    scene, cam = make_scene(20000, seed=21).to(dev), make_camera(333, 190).to(dev)
    bg = torch.tensor([0.2, 0.1, 0.3], device=dev)
    grads = tuple(g.to(dev) for g in make_pixel_grads(333, 190, 22))
    outs = {}
    try:
        for n in (1, 2, 4):
            nat.lib.b3gs_set_backward_pixels(n)
            outs[n] = util.surface_forward_backward(nat, scene, cam, bg, grads)
    finally:
        nat.lib.b3gs_set_backward_pixels(0)
    for n in (2, 4):
        for k in GRAD_KEYS:
            assert util.rel_err(outs[n][k], outs[1][k]) <= 5e-5, (n, k)


@pytest.mark.gpu
def test_compiled_host_side_matches_ctypes_host_side():
    """csrc/torch_binding.cpp and _backend.py are two host sides over the same C-ABI calls:
    identical images, radii and blobs' contents; gradients equal up to the order of the
    float REDs; NULL depth/alpha gradients accepted by both."""
    from binocular3dgs_b200 import _backend
    from binocular3dgs_b200.rasterizer import make_surface
    comp = _backend.preferred()
    if not isinstance(comp, _backend.CompiledBackend):
        pytest.skip("compiled host side not built")
    nat = _backend.native()
    dev = torch.device("cuda:0")
    scene, cam = make_scene(5000, seed=31).to(dev), make_camera(160, 120).to(dev)
    bg = torch.tensor([0.3, 0.2, 0.1], device=dev)
    grads = tuple(g.to(dev) for g in make_pixel_grads(160, 120, 32))
    a, b = util.raw_forward(comp, scene, cam, bg), util.raw_forward(nat, scene, cam, bg)
    assert a["R"] == b["R"]
    for k in ("color", "depth", "alpha", "radii"):
        assert torch.equal(a[k], b[k]), k
    ga = util.surface_forward_backward(comp, scene, cam, bg, grads)
    gb = util.surface_forward_backward(nat, scene, cam, bg, grads)
    for k in GRAD_KEYS:
        assert util.rel_err(ga[k], gb[k]) <= 2e-5, k
    before = nat.launch_count()
    assert comp.mark_visible(scene.means3D, cam.world_view_transform, cam.full_proj_transform).dtype == torch.bool
    assert nat.launch_count() == before + 1          # one library instance behind both
    with pytest.raises(RuntimeError, match="num_points, 3"):
        comp.rasterize_gaussians(bg, scene.means3D.view(-1), *([torch.empty(0)] * 4), 1.0, torch.empty(0),
                                 cam.world_view_transform, cam.full_proj_transform, 0.5, 0.5, 8, 8, torch.empty(0), 0,
                                 cam.camera_center, False, False)


@pytest.mark.gpu
def test_misaligned_views_and_odd_bucket_are_safe():
    """Contiguous views at a 4-byte offset (float4 loads in the kernels) and the DP gradient
    sink with an odd P (segment starts) must not fault and must give the same numbers."""
    from binocular3dgs_b200 import _backend, dp
    nat = _backend.native()
    dev = torch.device("cuda:0")
    P = 3001
    scene, cam = make_scene(P, seed=41).to(dev), make_camera(96, 64).to(dev)
    bg = torch.zeros(3, device=dev)
    grads = tuple(g.to(dev) for g in make_pixel_grads(96, 64, 42))
    ref = util.surface_forward_backward(nat, scene, cam, bg, grads)
    flat = torch.zeros(4 * P + 1, device=dev)
    flat[1:] = scene.rotations.reshape(-1)
    odd = Scene(scene.means3D, scene.scales, flat[1:].view(P, 4), scene.opacities, scene.shs, scene.sh_degree)
    assert odd.rotations.data_ptr() % 16 != 0 and odd.rotations.is_contiguous()
    for back in (nat, _backend.preferred()):
        got = util.surface_forward_backward(back, odd, cam, bg, grads)
        assert torch.equal(got["color"], ref["color"])
        for k in GRAD_KEYS:
            assert util.rel_err(got[k], ref[k]) <= 2e-5, k
    bucket = dp.GradientBucket(P, 4, dev)
    nat.grad_sink = bucket.views()
    try:
        got = util.surface_forward_backward(nat, scene, cam, bg, grads)
    finally:
        nat.grad_sink = None
    torch.cuda.synchronize()
    assert util.rel_err(bucket.views()["rotations"], ref["g_rotations"]) <= 2e-5
    assert util.rel_err(bucket.views()["shs"], ref["g_shs"]) <= 2e-5


# ------------------------------------------------------------------ the forward without its host wait
def _run_surface(S, scene, cam, bg, grads, cam2=None):
    leaves = [t.detach().clone().requires_grad_(True) for t in scene.tensors()]
    m3, sc, ro, op, sh = leaves
    outs = []
    for c in (cam,) if cam2 is None else (cam, cam2):      # both forwards are in flight before the backward
        m2 = torch.zeros_like(m3, requires_grad=True)
        outs.append(S.GaussianRasterizer(util.settings_for(c, bg, scene.sh_degree))(
            means3D=m3, means2D=m2, opacities=op, shs=sh, scales=sc, rotations=ro))
    gc, gd, ga = grads
    color, radii, depth, alpha = outs[0]
    loss = (color * gc).sum() + (depth * gd).sum() + (alpha * ga).sum()
    if cam2 is not None:
        loss = loss + (outs[1][0] * gc).sum()
    loss.backward()
    res = dict(color=color.detach(), depth=depth.detach(), alpha=alpha.detach(), radii=radii)
    res.update({"g%d" % i: t.grad for i, t in enumerate(leaves)})
    return res


@pytest.mark.gpu
@pytest.mark.parametrize("host", ["compiled", "ctypes"])
def test_nosync_forward_equals_exact_forward_and_recovers_from_overflow(host):
    """b3gs_forward_nosync (include/b3gs.h): after the first (exact) forward of a size, training
    forwards do not wait for R.  Same images bit for bit, same gradients up to the order of the
    float REDs; an undersized buffer is detected when the backward resolves the ticket, the
    forward is redone exactly into the same output tensors, and a warning says so."""
    from binocular3dgs_b200 import _backend
    from binocular3dgs_b200.rasterizer import make_surface
    back = _backend.preferred() if host == "compiled" else _backend.native()
    if host == "compiled" and not isinstance(back, _backend.CompiledBackend):
        pytest.skip("compiled host side not built")
    dev = torch.device("cuda:0")
    W, H, P = 333, 190, 20000
    scene, cam = make_scene(P, seed=51).to(dev), make_camera(W, H).to(dev)
    cam2 = make_camera(W, H, shift_x=0.2).to(dev)
    bg = torch.tensor([0.2, 0.1, 0.3], device=dev)
    grads = tuple(g.to(dev) for g in make_pixel_grads(W, H, 52))
    S = make_surface(back)
    pol = S.async_policy
    assert pol is not None and pol.enabled
    exact = _run_surface(S, scene, cam, bg, grads)                       # first call of this size: exact path
    key = (0, P, W, H)
    R = pol.hwm[key]
    assert R > 0 and pol.capacity(key) >= R
    launched = back.launch_count()
    lazy = _run_surface(S, scene, cam, bg, grads)                        # second call: no host wait
    assert back.launch_count() > launched and pol.overflows == 0
    for k in ("color", "depth", "alpha", "radii"):
        assert torch.equal(lazy[k], exact[k]), k
    for i in range(5):
        assert util.rel_err(lazy["g%d" % i], exact["g%d" % i]) <= 2e-5, i
    # the binocular pair: two tickets outstanding
    pair_exact = None
    try:
        pol.enabled = False
        pair_exact = _run_surface(S, scene, cam, bg, grads, cam2)
    finally:
        pol.enabled = True
    pair_lazy = _run_surface(S, scene, cam, bg, grads, cam2)
    assert torch.equal(pair_lazy["color"], pair_exact["color"])
    for i in range(5):
        assert util.rel_err(pair_lazy["g%d" % i], pair_exact["g%d" % i]) <= 2e-5, i
    # overflow: a buffer half the size it needs
    pol.forced_capacity = max(1, R // 2)
    try:
        with pytest.warns(RuntimeWarning, match="re-rendering exactly"):
            over = _run_surface(S, scene, cam, bg, grads)
    finally:
        pol.forced_capacity = None
    assert pol.overflows == 1
    for k in ("color", "depth", "alpha", "radii"):
        assert torch.equal(over[k], exact[k]), k                          # re-rendered into the same tensors
    for i in range(5):
        assert util.rel_err(over["g%d" % i], exact["g%d" % i]) <= 2e-5, i
    # without gradients (evaluation) the forward stays exact: no ticket, nothing to resolve
    with torch.no_grad():
        c = S.GaussianRasterizer(util.settings_for(cam, bg, scene.sh_degree))(
            means3D=scene.means3D, means2D=torch.zeros_like(scene.means3D), opacities=scene.opacities,
            shs=scene.shs, scales=scene.scales, rotations=scene.rotations)[0]
    assert torch.equal(c, exact["color"])


@pytest.mark.gpu
def test_binning_with_screen_filling_gaussians(nat, ref, dev):
    """A few thousand Gaussians that each cover most of the image: more instances per (batch of
    depth-ordered Gaussians, band of tile rows) than the direct tile binning's 16-bit slots can
    number — its wide path (binning.cu) — and lists of thousands of entries in every tile."""
    W, H, P = 512, 384, 3000
    scene = make_scene(P, seed=61, scale_lo=0.8, scale_hi=2.0)
    scene.opacities[:] = 0.02 + 0.05 * scene.opacities          # faint: no early termination, long walks
    scene = scene.to(dev)
    cam = make_camera(W, H, azimuth=0.5).to(dev)
    bg = torch.tensor([0.1, 0.1, 0.1], device=dev)
    fn, fr = util.raw_forward(nat, scene, cam, bg), util.raw_forward(ref, scene, cam, bg)
    inn, inr = util.internals(nat, fn, P, W, H), util.internals(ref, fr, P, W, H)
    T = ((W + 15) // 16) * ((H + 15) // 16)
    assert fn["R"] == fr["R"] and fn["R"] > 512 * 65535 // 64 and fn["R"] > 0.5 * int((fr["radii"] > 0).sum()) * T
    assert (inn["point_list"] == inr["point_list"]).all()
    assert (inn["ranges"] == inr["ranges"]).all()
    assert (inn["n_contrib"] == inr["n_contrib"]).all()
    for k in ("color", "depth", "alpha"):
        assert (_bits(fn[k]) == _bits(fr[k])).all(), k


@pytest.mark.gpu
@pytest.mark.parametrize("W,H", [(20000, 48), (48, 20000), (32640, 16), (1250 * 16, 400)])
def test_binning_on_extreme_tile_grids(W, H, nat, ref, dev):
    """Tile rows wider than the default band of the direct tile binning (one row per band, up to the
    2040-tile limit of the packed widths), 1250 rows of 3 tiles (bands capped at 200 rows), and a
    grid whose bands hold several wide rows: lists, ranges and images stay those of the reference."""
    P = 6000
    scene = make_scene(P, seed=67, scale_lo=0.02, scale_hi=0.3).to(dev)
    cam = make_camera(W, H, azimuth=0.3).to(dev)
    bg = torch.tensor([0.2, 0.1, 0.4], device=dev)
    fn, fr = util.raw_forward(nat, scene, cam, bg), util.raw_forward(ref, scene, cam, bg)
    inn, inr = util.internals(nat, fn, P, W, H), util.internals(ref, fr, P, W, H)
    assert fn["R"] == fr["R"] and fn["R"] > 0
    assert (inn["point_list"] == inr["point_list"]).all()
    assert (inn["ranges"] == inr["ranges"]).all()
    assert (inn["n_contrib"] == inr["n_contrib"]).all()
    for k in ("color", "depth", "alpha"):
        assert (_bits(fn[k]) == _bits(fr[k])).all(), k


@pytest.mark.gpu
def test_forward_and_backward_capture_into_a_cuda_graph(nat, dev):
    """With the no-sync forward nothing on the path blocks the host, so forward + backward can be
    captured into ONE CUDA graph (12 kernel nodes, the cooperative depth sort included) and replayed
    with new camera matrices written into the captured input tensors: same images bit for bit, same
    gradients up to the order of the float REDs."""
    W, H, P = 400, 300, 30000
    scene = make_scene(P, seed=71).to(dev)
    cams = [make_camera(W, H, azimuth=a).to(dev) for a in (0.3, 1.4)]
    bg = torch.tensor([0.2, 0.3, 0.1], device=dev)
    gc, gd, ga = (t.to(dev) for t in make_pixel_grads(W, H, 72))
    e = torch.empty(0)
    view, proj, center = (t.clone() for t in (cams[0].world_view_transform, cams[0].full_proj_transform,
                                              cams[0].camera_center))

    def forward(nosync_capacity=None):
        args = (bg, scene.means3D, e, scene.opacities, scene.scales, scene.rotations, 1.0, e, view, proj,
                cams[0].tanfovx, cams[0].tanfovy, H, W, scene.shs, scene.sh_degree, center, False, False)
        return (nat.rasterize_gaussians(*args) if nosync_capacity is None
                else nat.rasterize_gaussians_nosync(nosync_capacity, *args))

    def backward(out, R):
        return nat.rasterize_gaussians_backward(bg, scene.means3D, out[4], e, scene.scales, scene.rotations, 1.0, e, view,
                                                proj, cams[0].tanfovx, cams[0].tanfovy, gc, gd, ga, scene.shs,
                                                scene.sh_degree, center, out[5], R, out[6], out[7], out[3], False)

    def set_camera(c):
        view.copy_(c.world_view_transform); proj.copy_(c.full_proj_transform); center.copy_(c.camera_center)

    eager = []
    for c in cams:                                   # exact, eager: what every replay must reproduce
        set_camera(c)
        out = forward()
        eager.append((out, [g.clone() for g in backward(out, out[0])]))
    capacity = 2 * max(o[0][0] for o in eager)
    side = torch.cuda.Stream(device=dev)             # warm-up on a side stream, as torch's capture recipe asks
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        out = forward(capacity)
        backward(out, eager[-1][0][0])
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        g_out = forward(capacity)
        g_grads = backward(g_out, capacity // 2)      # R only steers the backward's kernel shape
    for c, (out, grads) in zip(cams, eager):
        set_camera(c)
        graph.replay()
        torch.cuda.synchronize()
        for k in (1, 2, 3, 4):                       # colour, depth, alpha, radii
            assert torch.equal(g_out[k], out[k]), k
        for a, b in zip(g_grads, grads):
            assert util.rel_err(a, b) <= 2e-5
