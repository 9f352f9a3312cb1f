"""Shared helpers for the parity tests: drive a backend (ours or the reference veneer)
through the operator surface and collect outputs plus sliced internals."""
from __future__ import annotations

import numpy as np
import torch

from binocular3dgs_b200.rasterizer import GaussianRasterizationSettings, make_surface
from workloads import Camera, Scene


def settings_for(cam: Camera, bg: torch.Tensor, sh_degree: int, scale_modifier=1.0, debug=False):
    return GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
        bg=bg, scale_modifier=scale_modifier, viewmatrix=cam.world_view_transform,
        projmatrix=cam.full_proj_transform, sh_degree=sh_degree, campos=cam.camera_center, prefiltered=False,
        debug=debug)


def raw_forward(backend, scene: Scene, cam: Camera, bg, scale_modifier=1.0, colors_precomp=None, cov3D_precomp=None,
                debug=False):
    """Call the _C-level forward; return outputs and the blobs."""
    e = torch.empty(0)
    out = backend.rasterize_gaussians(
        bg, scene.means3D, e if colors_precomp is None else colors_precomp, scene.opacities,
        e if cov3D_precomp is not None else scene.scales, e if cov3D_precomp is not None else scene.rotations,
        scale_modifier, e if cov3D_precomp is None else cov3D_precomp, cam.world_view_transform,
        cam.full_proj_transform, cam.tanfovx, cam.tanfovy, cam.image_height, cam.image_width,
        e if colors_precomp is not None else scene.shs, scene.sh_degree, cam.camera_center, False, debug)
    keys = ("R", "color", "depth", "alpha", "radii", "geom", "binning", "img")
    return dict(zip(keys, out))


def internals(backend, fwd, P, W, H):
    """Slice bit-exact-comparable internals out of the blobs."""
    R = fwd["R"]
    T = ((W + 15) // 16) * ((H + 15) // 16)
    out = {}
    out["depths"] = backend.blob_view(fwd["geom"], "geometry", "depths", torch.int32, P, P).clone()
    out["tiles_touched"] = backend.blob_view(fwd["geom"], "geometry", "tiles_touched", torch.int32, P, P).clone()
    out["point_list"] = backend.blob_view(fwd["binning"], "binning", "point_list", torch.int32, R, R).clone() \
        if R > 0 else torch.empty(0, dtype=torch.int32)
    out["ranges"] = backend.blob_view(fwd["img"], "image", "ranges", torch.int32, 2 * T, W, H).clone()
    out["n_contrib"] = backend.blob_view(fwd["img"], "image", "n_contrib", torch.int32, W * H, W, H).clone()
    if backend.prefix == "b3gs_":
        rec = backend.blob_view(fwd["geom"], "geometry", "records", torch.float32, 12 * P, P).view(P, 12)
        out["means2D"] = rec[:, 0:2].clone()
        out["conic_opacity"] = rec[:, 4:8].clone()
        out["rgb"] = rec[:, 8:11].clone()
    else:
        out["means2D"] = backend.blob_view(fwd["geom"], "geometry", "means2D", torch.float32, 2 * P, P).view(P, 2).clone()
        out["conic_opacity"] = backend.blob_view(fwd["geom"], "geometry", "conic_opacity", torch.float32, 4 * P, P).view(P, 4).clone()
        out["rgb"] = backend.blob_view(fwd["geom"], "geometry", "rgb", torch.float32, 3 * P, P).view(P, 3).clone()
    return out


def surface_forward_backward(backend, scene: Scene, cam: Camera, bg, grads, scale_modifier=1.0):
    """Forward + backward through the autograd surface; returns outputs and leaf grads."""
    S = make_surface(backend)
    leaves = [t.detach().clone().requires_grad_(True) for t in scene.tensors()]
    means3D, scales, rotations, opacities, shs = leaves
    means2D = torch.zeros_like(means3D, requires_grad=True)
    rast = S.GaussianRasterizer(settings_for(cam, bg, scene.sh_degree, scale_modifier))
    color, radii, depth, alpha = rast(means3D=means3D, means2D=means2D, opacities=opacities, shs=shs, scales=scales,
                                      rotations=rotations)
    gc, gd, ga = grads
    torch.autograd.backward([color, depth, alpha], [gc, gd, ga])
    return dict(color=color.detach(), depth=depth.detach(), alpha=alpha.detach(), radii=radii,
                g_means3D=means3D.grad, g_means2D=means2D.grad, g_scales=scales.grad, g_rotations=rotations.grad,
                g_opacities=opacities.grad, g_shs=shs.grad)


def max_abs(a, b):
    return float((a.double() - b.double()).abs().max()) if a.numel() else 0.0


def rel_err(a, b):
    """max |a-b| scaled by the tensor's max magnitude (gradient comparison)."""
    denom = float(b.double().abs().max())
    return max_abs(a, b) / denom if denom > 0 else max_abs(a, b)
