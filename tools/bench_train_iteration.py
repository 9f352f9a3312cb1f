"""One training iteration of the reference's loop (train.py:83-194) built from the REFERENCE'S OWN
FILES — gaussian_renderer.render, scene.gaussian_model.GaussianModel (activations, Adam setup,
add_densification_stats, opacity_decay), utils.loss_utils (l1_loss, ssim, SmoothLoss),
utils.graphics_utils.inverse_warp_images — loaded unmodified from baseline/_ref/reference_tree
(or /root/reference), on a rasterizer module of the caller's choice:

  style "reference"  the stock diff_gaussian_rasterization (baseline/_ref): the reference as it is
  style "dropin"     the same files with THIS repository's drop-in as `diff_gaussian_rasterization`:
                     what a user gets by switching the rasterizer and nothing else

Only the loop body itself is restated here (train.py is a script, not a function): the statements
of train.py:84-194 that run every iteration on the LLFF binocular configuration (binocular
consistency on, opacity decay on, no densify step inside the timed window).  Synthetic scene and
cameras as everywhere else (workloads.py); 2 "views" of BASELINE.json's metric per iteration.
"""
import json
import os
import random
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))


def measure(dev, style="reference", config="fern", iters=20, warmup=5):
    import reference_loader
    from workloads import CONFIGS, make_camera, make_scene
    if style == "reference":
        dgr = reference_loader.stock()
    elif style == "dropin":
        import binocular3dgs_b200 as dgr
    else:
        raise ValueError(style)
    if dgr is None:
        return {"error": "baseline/_ref/diff_gaussian_rasterization is not built"}
    loaded = reference_loader.render_adapter(dgr, "gaussian_renderer_for_" + style)
    if loaded is None:
        return {"error": "the reference tree is not available (baseline/_ref/reference_tree)"}
    gr, GaussianModel = loaded
    loss_utils = reference_loader.reference_module("utils.loss_utils")
    graphics_utils = reference_loader.reference_module("utils.graphics_utils")
    l1_loss, ssim, SmoothLoss = loss_utils.l1_loss, loss_utils.ssim, loss_utils.SmoothLoss
    inverse_warp_images = graphics_utils.inverse_warp_images

    cfg = CONFIGS[config]
    W, H, P = cfg["width"], cfg["height"], cfg["P"]
    scene = make_scene(P, seed=0)
    # the reference's GaussianModel holding the synthetic scene as raw parameters, its own optimizer
    gaussians = GaussianModel(scene.sh_degree)
    gaussians.active_sh_degree = scene.sh_degree
    op = scene.opacities.clamp(1e-4, 1 - 1e-4)
    par = lambda t: torch.nn.Parameter(t.to(dev).contiguous().requires_grad_(True))
    gaussians._xyz, gaussians._features_dc, gaussians._features_rest = par(scene.means3D), par(scene.shs[:, :1]), par(scene.shs[:, 1:])
    gaussians._opacity, gaussians._scaling, gaussians._rotation = par(torch.log(op / (1 - op))), par(torch.log(scene.scales)), par(scene.rotations)
    gaussians.max_radii2D = torch.zeros(P, device=dev)
    gaussians.spatial_lr_scale = 1.0
    opt = types.SimpleNamespace(percent_dense=0.01, position_lr_init=0.00016, position_lr_final=0.0000016,
                                position_lr_delay_mult=0.01, position_lr_max_steps=30_000, feature_lr=0.0025,
                                opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001, lambda_dssim=0.2)   # arguments/__init__.py:75-84
    gaussians.training_setup(opt)
    pipe = types.SimpleNamespace(convert_SHs_python=False, compute_cov3D_python=False, debug=False)
    background = torch.tensor([0, 0, 0], dtype=torch.float32, device=dev)
    cam = make_camera(W, H, cfg["fovx"]).to(dev)
    g = torch.Generator().manual_seed(2)
    gt_image = torch.rand(3, H, W, generator=g).to(dev)
    focal_x = W / (2.0 * cam.tanfovx)
    # train.py:54-60
    row_indices = torch.arange(0, H).view(-1, 1).repeat(1, W).cuda()
    column_indices = torch.arange(0, W).repeat(H, 1).cuda()
    mask = torch.ones((1, H, W), dtype=torch.float32).cuda()
    smooth_loss = SmoothLoss()
    random.seed(0)
    torch.manual_seed(0)
    shifted = {}

    def iteration(it):
        gaussians.update_learning_rate(it)                                    # train.py:85
        render_pkg = gr.render(cam, gaussians, pipe, background)              # :100
        image, viewspace_point_tensor = render_pkg["render"], render_pkg["viewspace_points"]
        visibility_filter, radii, depth = render_pkg["visibility_filter"], render_pkg["radii"], render_pkg["rendered_depth"]
        trans_dist = torch.rand(1) * 0.4                                       # :122-126 (cam_trans_dist 0.4)
        trans_dist = (trans_dist * random.choice([-1.0, 1.0])).item()
        key = round(trans_dist, 2)                                             # Scene.getShiftedCamera (scene/__init__.py:96-115)
        if key not in shifted:
            shifted[key] = make_camera(W, H, cfg["fovx"], shift_x=key).to(dev)
        render_pkg = gr.render(shifted[key], gaussians, pipe, background)      # :127
        shifted_image = render_pkg["render"]
        disparity = focal_x * (-key) / (depth + 1e-5)                          # :130-136
        warped_image = inverse_warp_images(shifted_image.unsqueeze(0), disparity.unsqueeze(0), row_indices, column_indices)
        shift_mask = inverse_warp_images(mask.unsqueeze(0), disparity.unsqueeze(0), row_indices, column_indices)
        disparity_loss = (l1_loss(warped_image, gt_image.unsqueeze(0), mask=shift_mask) +
                          0.05 * smooth_loss.forward(disparity=disparity * shift_mask, image=gt_image.unsqueeze(0)))
        Ll1 = l1_loss(image, gt_image)                                         # :146-149
        loss = (1.0 - opt.lambda_dssim) * Ll1 + opt.lambda_dssim * (1.0 - ssim(image, gt_image))
        total_loss = loss + disparity_loss
        total_loss.backward()
        with torch.no_grad():
            gaussians.opacity_decay(factor=0.995)                              # :163-165
            gaussians.max_radii2D[visibility_filter] = torch.max(gaussians.max_radii2D[visibility_filter],
                                                                 radii[visibility_filter])           # :170
            gaussians.add_densification_stats(viewspace_point_tensor, visibility_filter)             # :171
            gaussians.optimizer.step()                                         # :192-193
            gaussians.optimizer.zero_grad(set_to_none=True)
        return total_loss

    for it in range(1, warmup + 1):
        last = iteration(it)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(warmup + 1, warmup + iters + 1):
        last = iteration(it)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return {"what": "train.py:84-194 on %s from the reference's own files (render, GaussianModel, loss_utils, "
                    "graphics_utils, torch.optim.Adam); P=%d, %dx%d; rasterizer: %s" % (config, P, W, H, style),
            "ms_per_iteration": round(ms, 4), "views_per_s": round(2000.0 / ms, 1), "views_per_iteration": 2,
            "last_loss": float(last.detach())}


if __name__ == "__main__":
    print(json.dumps(measure(torch.device("cuda:0"), sys.argv[1] if len(sys.argv) > 1 else "reference")))
