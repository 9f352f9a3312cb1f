"""GPU-box diagnostic: parity of libb3gs vs the reference kernels + first timings.
Run:  python tools/gpu_probe.py [config ...]     (writes gpurun_out/probe_<cfg>.json)
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from binocular3dgs_b200 import _backend  # noqa: E402
from workloads import CONFIGS, make_camera, make_pixel_grads, make_scene  # noqa: E402
from oracle import cpu_oracle as orc  # noqa: E402
from oracle import refbackend  # noqa: E402
import util  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, warm=5, iters=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def probe(name, kind="cube", seed=0):
    cfg = CONFIGS[name]
    W, H, P = cfg["width"], cfg["height"], cfg["P"]
    scene = make_scene(P, seed=seed, kind=kind).to(dev)
    cam = make_camera(W, H, cfg["fovx"]).to(dev)
    bg = torch.tensor([0.0, 0.0, 0.0], device=dev)
    grads = tuple(g.to(dev) for g in make_pixel_grads(W, H, seed + 1))
    nat, ref = _backend.native(), refbackend.reference()
    res = {"config": name, "kind": kind, "P": P, "W": W, "H": H}

    fn = util.raw_forward(nat, scene, cam, bg)
    fr = util.raw_forward(ref, scene, cam, bg)
    torch.cuda.synchronize()
    res["R_native"], res["R_ref"] = fn["R"], fr["R"]
    inn, inr = util.internals(nat, fn, P, W, H), util.internals(ref, fr, P, W, H)
    vis = fr["radii"] > 0
    res["visible"] = int(vis.sum())
    res["radii_mismatch"] = int((fn["radii"] != fr["radii"]).sum())
    for k in ("depths", "tiles_touched"):
        res[k + "_mismatch"] = int((inn[k][vis] != inr[k][vis]).sum())
    for k in ("means2D", "conic_opacity", "rgb"):
        a, b = inn[k][vis].view(torch.int32), inr[k][vis].view(torch.int32)
        res[k + "_bit_mismatch"] = int((a != b).sum())
        res[k + "_maxabs"] = util.max_abs(inn[k][vis], inr[k][vis])
    if fn["R"] == fr["R"]:
        res["point_list_mismatch"] = int((inn["point_list"] != inr["point_list"]).sum())
    res["ranges_mismatch"] = int((inn["ranges"] != inr["ranges"]).sum())
    res["n_contrib_mismatch"] = int((inn["n_contrib"] != inr["n_contrib"]).sum())
    for k in ("color", "depth", "alpha"):
        res[k + "_maxabs"] = util.max_abs(fn[k], fr[k])
        res[k + "_bit_mismatch"] = int((fn[k].view(torch.int32) != fr[k].view(torch.int32)).sum())
    res["depth_max"] = float(fr["depth"].max())
    nc = inr["n_contrib"].double()
    rng = inr["ranges"].view(-1, 2).long()
    res["tile_len_mean"] = float((rng[:, 1] - rng[:, 0]).double().mean())
    res["tile_len_max"] = int((rng[:, 1] - rng[:, 0]).max())
    res["n_contrib_mean"] = float(nc.mean())

    # backward parity through the autograd surface
    gn = util.surface_forward_backward(nat, scene, cam, bg, grads)
    gr = util.surface_forward_backward(ref, scene, cam, bg, grads)
    gr2 = util.surface_forward_backward(ref, scene, cam, bg, grads)
    for k in gn:
        if k.startswith("g_"):
            res[k + "_rel"] = util.rel_err(gn[k], gr[k])
            res[k + "_refspread"] = util.rel_err(gr2[k], gr[k])
            res[k + "_finite"] = bool(torch.isfinite(gn[k]).all())

    # CPU oracle vs reference (pinning): preprocess bitwise, binning exact, image tolerance
    if P <= 300_000:
        sc = scene.to("cpu")
        cc = cam.to("cpu")
        t0 = time.time()
        of = orc.rasterize_forward(sc.means3D.numpy(), sc.scales.numpy(), sc.rotations.numpy(), sc.opacities.numpy(),
                                   sc.shs.numpy(), cc.world_view_transform.numpy(), cc.full_proj_transform.numpy(),
                                   cc.camera_center.numpy(), np.zeros(3, np.float32), W, H, cc.tanfovx, cc.tanfovy,
                                   sc.sh_degree)
        res["oracle_fwd_s"] = time.time() - t0
        res["oracle_threads"] = orc.num_threads()
        v = vis.cpu().numpy()
        res["oracle_R"] = of["R"]
        res["oracle_radii_mismatch"] = int((of["radii"] != fr["radii"].cpu().numpy()).sum())
        res["oracle_depth_bit_mismatch"] = int((of["depths"].view(np.int32)[v] != inr["depths"].cpu().numpy()[v]).sum())
        res["oracle_tiles_mismatch"] = int((of["tiles_touched"].view(np.int32)[v] != inr["tiles_touched"].cpu().numpy()[v]).sum())
        for k in ("means2D", "conic_opacity", "rgb"):
            res["oracle_" + k + "_bit_mismatch"] = int((of[k].view(np.int32)[v] != inr[k].cpu().numpy().view(np.int32)[v]).sum())
        if of["R"] == fr["R"]:
            res["oracle_point_list_mismatch"] = int((of["point_list"].view(np.int32) != inr["point_list"].cpu().numpy()).sum())
            res["oracle_ranges_mismatch"] = int((of["ranges"].reshape(-1).view(np.int32) != inr["ranges"].cpu().numpy()).sum())
        res["oracle_n_contrib_mismatch"] = int((of["n_contrib"].view(np.int32) != inr["n_contrib"].cpu().numpy()).sum())
        for k in ("color", "depth", "alpha"):
            res["oracle_" + k + "_maxabs"] = float(np.abs(of[k].astype(np.float64) - fr[k].cpu().numpy()).max())
        gc, gd, ga = (g.cpu().numpy() for g in grads)
        t0 = time.time()
        ob = orc.rasterize_backward(of, sc.means3D.numpy(), sc.scales.numpy(), sc.rotations.numpy(), sc.shs.numpy(),
                                    cc.world_view_transform.numpy(), cc.full_proj_transform.numpy(),
                                    cc.camera_center.numpy(), np.zeros(3, np.float32), W, H, cc.tanfovx, cc.tanfovy,
                                    sc.sh_degree, gc, gd, ga)
        res["oracle_bwd_s"] = time.time() - t0
        pairs = dict(g_means3D="dL_dmeans3D", g_scales="dL_dscales", g_rotations="dL_drotations",
                     g_opacities="dL_dopacity", g_shs="dL_dsh", g_means2D="dL_dmean2D")
        for k, ok in pairs.items():
            o = torch.from_numpy(ob[ok]).to(dev)
            res["oracle_" + k + "_rel_vs_ref"] = util.rel_err(gr[k], o)
            res["oracle_" + k + "_rel_vs_native"] = util.rel_err(gn[k], o)

    # timings (device time, CUDA events)
    S_n, S_r = util.make_surface(nat), util.make_surface(ref)

    def step(back, S):
        def f():
            out = util.raw_forward(back, scene, cam, bg)
            e = torch.empty(0)
            back.rasterize_gaussians_backward(
                bg, scene.means3D, out["radii"], e, scene.scales, scene.rotations, 1.0, e, cam.world_view_transform,
                cam.full_proj_transform, cam.tanfovx, cam.tanfovy, grads[0], grads[1], grads[2], scene.shs,
                scene.sh_degree, cam.camera_center, out["geom"], out["R"], out["binning"], out["img"], out["alpha"],
                False)
        return f

    def fwd_only(back):
        return lambda: util.raw_forward(back, scene, cam, bg)

    res["t_fwd_native_ms"], _ = timeit(fwd_only(nat))
    res["t_fwd_ref_ms"], _ = timeit(fwd_only(ref))
    res["t_step_native_ms"], res["t_step_native_best_ms"] = timeit(step(nat, S_n))
    res["t_step_ref_ms"], res["t_step_ref_best_ms"] = timeit(step(ref, S_r))
    # wall-clock per step (host overhead included)
    for nm, back in (("native", nat), ("ref", ref)):
        f = step(back, None)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(20):
            f()
        torch.cuda.synchronize()
        res["wall_step_%s_ms" % nm] = (time.time() - t0) / 20 * 1e3
    res["speedup_step"] = res["t_step_ref_ms"] / res["t_step_native_ms"]
    return res


if __name__ == "__main__":
    names = sys.argv[1:] or ["plumbing", "lego"]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    print(torch.cuda.get_device_name(0), torch.__version__)
    for nm in names:
        kind = "cube"
        if ":" in nm:
            nm, kind = nm.split(":")
        r = probe(nm, kind)
        with open(os.path.join(ROOT, "gpurun_out", "probe_%s_%s.json" % (nm, kind)), "w") as f:
            json.dump(r, f, indent=1)
        print(json.dumps(r, indent=1))
