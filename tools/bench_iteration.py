"""One full training iteration of the reference's loop (train.py:83-194) on synthetic data,
timed two ways on the same GPU:

  reference_style  the reference's own rasterizer kernels (oracle/_ref, unmodified) driven
                   through the same autograd surface, and everything around them written
                   the way the reference writes it: torch activations per render
                   (gaussian_model.py:95-115), five-conv SSIM + L1 (loss_utils.py), the
                   Python-loop inverse warp + SmoothLoss (graphics_utils.py:80-125,
                   loss_utils.py:68-91), boolean-mask densification statistics
                   (gaussian_model.py:409-411) and torch.optim.Adam.
  native           this library for every one of those rows: rasterizer (§8 a), fused
                   photometric loss (f1), fused binocular loss (f2), fused activations,
                   densification statistics and one-launch Adam (f3); the activations are
                   evaluated once per iteration for both views of the pair.

An iteration = render the training view + render the shifted view (binocular pair,
train.py:122-127) + losses + backward + densification statistics + optimizer step,
opacity decay, i.e. 2 "views" of BASELINE.json's metric.  CUDA events around the whole loop,
host-side Python included (it is part of what a user waits for).
Used by bench.py ("next_rows") and runnable on its own:  python tools/bench_iteration.py fern
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

LRS = dict(xyz=0.00016, f_dc=0.0025, f_rest=0.0025 / 20.0, opacity=0.05, scaling=0.005, rotation=0.001)
ORDER = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


def _raw_parameters(scene, dev):
    """Invert the activations of a synthetic scene into the reference's raw parameters."""
    op = scene.opacities.clamp(1e-4, 1 - 1e-4)
    return {"xyz": scene.means3D.clone(), "f_dc": scene.shs[:, :1].clone(), "f_rest": scene.shs[:, 1:].clone(),
            "opacity": torch.log(op / (1 - op)), "scaling": torch.log(scene.scales), "rotation": scene.rotations.clone()}


def measure(dev, config="fern", iters=20, warmup=5, which=("native", "reference_style")):
    import bench_loss
    from binocular3dgs_b200 import _backend, binocular, losses, parameters
    from binocular3dgs_b200.rasterizer import GaussianRasterizationSettings, make_surface
    from workloads import CONFIGS, make_camera, make_scene

    cfg = CONFIGS[config]
    W, H, P = cfg["width"], cfg["height"], cfg["P"]
    scene = make_scene(P, seed=0).to(dev)
    cam = make_camera(W, H, cfg["fovx"]).to(dev)
    trans_dist = 0.23
    cam_shift = make_camera(W, H, cfg["fovx"], shift_x=trans_dist).to(dev)
    focal_x = W / (2.0 * cam.tanfovx)
    bg = torch.zeros(3, device=dev)
    g = torch.Generator().manual_seed(2)
    gt = torch.rand(3, H, W, generator=g).to(dev)
    rows = torch.arange(0, H).view(-1, 1).repeat(1, W).to(dev)
    cols = torch.arange(0, W).repeat(H, 1).to(dev)
    ones = torch.ones((1, H, W), dtype=torch.float32, device=dev)

    out = {"what": "train.py iteration on %s: 2 renders (binocular pair) + photometric + binocular loss + backward + "
                   "densification statistics + Adam; P=%d, %dx%d" % (config, P, W, H), "views_per_iteration": 2}

    def run(style):
        if style in ("native", "native_raw"):
            surface = make_surface(_backend.preferred())
            adam_cls = parameters.FusedAdam
        else:
            from oracle import refbackend
            surface = make_surface(refbackend.reference())
            adam_cls = torch.optim.Adam
            smooth = bench_loss._TorchStyleSmooth(dev)
        params = {k: torch.nn.Parameter(v.contiguous()) for k, v in _raw_parameters(scene, dev).items()}
        opt = adam_cls([{"params": [params[k]], "lr": LRS[k], "name": k} for k in ORDER], lr=0.0, eps=1e-15)
        accum, denom, max_radii = torch.zeros(P, 1, device=dev), torch.zeros(P, 1, device=dev), torch.zeros(P, device=dev)

        def render(camera, acts=None):
            screenspace = torch.zeros_like(params["xyz"], requires_grad=True)
            if style == "native_raw":
                # the raw-parameter entry: activations and their gradients inside the operator
                rast = surface.GaussianRasterizer(GaussianRasterizationSettings(
                    image_height=camera.image_height, image_width=camera.image_width, tanfovx=camera.tanfovx,
                    tanfovy=camera.tanfovy, bg=bg, scale_modifier=1.0, viewmatrix=camera.world_view_transform,
                    projmatrix=camera.full_proj_transform, sh_degree=1, campos=camera.camera_center,
                    prefiltered=False, debug=False))
                image, radii, depth, alpha = rast.forward_raw(params["xyz"], screenspace, params["f_dc"],
                                                              params["f_rest"], params["opacity"], params["scaling"],
                                                              params["rotation"])
                return image, radii, depth, screenspace
            if style == "native":
                # both views of the pair see the same parameters: activate once per iteration
                shs, opacity, scales, rotations = acts
            else:
                shs = torch.cat((params["f_dc"], params["f_rest"]), dim=1)
                opacity, scales = torch.sigmoid(params["opacity"]), torch.exp(params["scaling"])
                rotations = torch.nn.functional.normalize(params["rotation"])
            rast = surface.GaussianRasterizer(GaussianRasterizationSettings(
                image_height=camera.image_height, image_width=camera.image_width, tanfovx=camera.tanfovx,
                tanfovy=camera.tanfovy, bg=bg, scale_modifier=1.0, viewmatrix=camera.world_view_transform,
                projmatrix=camera.full_proj_transform, sh_degree=1, campos=camera.camera_center, prefiltered=False,
                debug=False))
            image, radii, depth, alpha = rast(means3D=params["xyz"], means2D=screenspace, opacities=opacity, shs=shs,
                                              scales=scales, rotations=rotations)
            return image, radii, depth, screenspace

        def iteration():
            acts = None
            if style == "native":
                acts = parameters.activate(params["f_dc"], params["f_rest"], params["opacity"], params["scaling"],
                                           params["rotation"])
            image, radii, depth, screenspace = render(cam, acts)
            shifted_image = render(cam_shift, acts)[0]
            if style != "reference_style":
                disparity_loss = binocular.binocular_consistency_loss(shifted_image, depth, gt, focal_x, trans_dist)
                loss = losses.photometric_loss(image, gt, 0.2)
            else:
                disparity = focal_x * (-trans_dist) / (depth + 1e-5)
                warped = bench_loss._torch_style_warp(shifted_image.unsqueeze(0), disparity.unsqueeze(0), rows, cols)
                mask = bench_loss._torch_style_warp(ones.unsqueeze(0), disparity.unsqueeze(0), rows, cols)
                disparity_loss = (torch.abs(warped * mask - gt.unsqueeze(0) * mask).mean() +
                                  0.05 * smooth(disparity * mask, gt.unsqueeze(0)))
                loss = bench_loss._torch_style_loss(image, gt, 0.2)
            total = loss + disparity_loss
            total.backward()
            with torch.no_grad():
                # train.py:163-165 (opacity_decay defaults to True, factor 0.995)
                if style != "reference_style":
                    parameters.opacity_decay(params["opacity"], 0.995)
                else:
                    o_ = torch.sigmoid(params["opacity"]) * 0.995
                    params["opacity"].data = torch.log(o_ / (1 - o_))
                if style != "reference_style":
                    parameters.add_densification_stats(screenspace.grad, radii, accum, denom, max_radii)
                else:
                    vis = radii > 0
                    max_radii[vis] = torch.max(max_radii[vis], radii[vis])
                    accum[vis] += torch.norm(screenspace.grad[vis, :2], dim=-1, keepdim=True)
                    denom[vis] += 1
                opt.step()
                opt.zero_grad(set_to_none=True)
            return total

        for _ in range(warmup):
            last = iteration()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            last = iteration()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters, float(last.detach())

    for style in which:
        try:
            ms, loss = run(style)
            out[style] = {"ms_per_iteration": round(ms, 4), "views_per_s": round(2000.0 / ms, 1), "last_loss": loss}
        except Exception as ex:   # e.g. oracle/_ref not built
            out[style] = {"error": repr(ex)}
    if all("ms_per_iteration" in out.get(s, {}) for s in ("native", "reference_style")):
        out["speedup"] = round(out["reference_style"]["ms_per_iteration"] / out["native"]["ms_per_iteration"], 2)
    return out


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "fern"
    print(json.dumps(measure(torch.device("cuda:0"), name)))
