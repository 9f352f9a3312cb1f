"""Seeded synthetic scenes and cameras for parity tests and ``bench.py``.

There are no datasets in this environment, so workloads are generated: everything is
drawn on the CPU from ``torch.Generator().manual_seed(seed)`` (bit-reproducible across
machines) and then moved to the device.  Cameras are built exactly like the reference's
``Camera`` (``scene/cameras.py:49-58``, ``utils/graphics_utils.py:38-71``): transposed
world-to-view matrix, transposed projection, their product, and the camera centre from
the inverse view matrix.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

SH_C0 = 0.28209479177387814


@dataclass
class Camera:
    image_width: int
    image_height: int
    FoVx: float
    FoVy: float
    world_view_transform: torch.Tensor  # (4,4) transposed W2C
    full_proj_transform: torch.Tensor  # (4,4)
    camera_center: torch.Tensor  # (3,)
    R: np.ndarray
    T: np.ndarray

    @property
    def tanfovx(self) -> float:
        return math.tan(self.FoVx * 0.5)

    @property
    def tanfovy(self) -> float:
        return math.tan(self.FoVy * 0.5)

    def to(self, device):
        return Camera(self.image_width, self.image_height, self.FoVx, self.FoVy,
                      self.world_view_transform.to(device), self.full_proj_transform.to(device),
                      self.camera_center.to(device), self.R, self.T)


def _world2view2(R, t, translate=np.array([0.0, 0.0, 0.0]), scale=1.0):
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.transpose()
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    C2W = np.linalg.inv(Rt)
    C2W[:3, 3] = (C2W[:3, 3] + translate) * scale
    return np.float32(np.linalg.inv(C2W))


def _projection(znear, zfar, fovX, fovY):
    tx, ty = math.tan(fovX / 2), math.tan(fovY / 2)
    top, right = ty * znear, tx * znear
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (2 * right)
    P[1, 1] = 2.0 * znear / (2 * top)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def make_camera(width: int, height: int, fovx: float = 0.6911, distance: float = 4.0, azimuth: float = 0.3,
                elevation: float = 0.2, shift_x: float = 0.0, znear: float = 0.01, zfar: float = 100.0) -> Camera:
    """Pinhole camera on a sphere of radius ``distance`` looking at the origin.

    ``shift_x`` translates the camera along its own x axis the way
    ``Scene.getShiftedCamera`` does for the binocular pair (scene/__init__.py:96-115).
    """
    c = distance * np.array([math.cos(elevation) * math.sin(azimuth), -math.sin(elevation),
                             -math.cos(elevation) * math.cos(azimuth)])
    fwd = -c / np.linalg.norm(c)
    down0 = np.array([0.0, 1.0, 0.0])  # image y points down (COLMAP convention)
    right = np.cross(down0, fwd)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R_w2c = np.stack([right, down, fwd], axis=0)  # rows
    t = -R_w2c @ c
    R = R_w2c.transpose()  # Camera.R is the transpose of the W2C rotation
    fovy = 2 * math.atan(height / (2 * (width / (2 * math.tan(fovx / 2)))))
    trans = np.array([0.0, 0.0, 0.0])
    if shift_x != 0.0:
        w2c = _world2view2(R, t)
        point_world = np.linalg.inv(w2c.astype(np.float64)) @ np.array([shift_x, 0.0, 0.0, 1.0])
        trans = point_world[:3] - c
    wvt = torch.tensor(_world2view2(R, t, trans)).transpose(0, 1).contiguous()
    proj = _projection(znear, zfar, fovx, fovy).transpose(0, 1)
    full = (wvt.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0).contiguous()
    center = wvt.inverse()[3, :3].contiguous()
    return Camera(width, height, fovx, fovy, wvt, full, center, R, t)


@dataclass
class Scene:
    means3D: torch.Tensor  # (P,3)
    scales: torch.Tensor  # (P,3) activated (positive)
    rotations: torch.Tensor  # (P,4) unit quaternions (r,x,y,z)
    opacities: torch.Tensor  # (P,1) in (0,1)
    shs: torch.Tensor  # (P,M,3)
    sh_degree: int

    @property
    def P(self) -> int:
        return int(self.means3D.shape[0])

    def to(self, device):
        return Scene(*(t.to(device) for t in (self.means3D, self.scales, self.rotations, self.opacities, self.shs)),
                     self.sh_degree)

    def tensors(self):
        return self.means3D, self.scales, self.rotations, self.opacities, self.shs


def make_scene(P: int, seed: int = 0, kind: str = "cube", sh_degree: int = 1, max_sh_degree: int | None = None,
               scale_lo: float = 0.005, scale_hi: float = 0.05) -> Scene:
    """``cube``: uniform in [-1.3,1.3]^3 (the reference's own random init,
    scene/dataset_readers.py:269-275).  ``shell``: three noisy concentric spheres with
    surface-aligned flat Gaussians (trained-scene-like early termination)."""
    g = torch.Generator().manual_seed(seed)
    M = ((sh_degree if max_sh_degree is None else max_sh_degree) + 1) ** 2
    U = lambda *s: torch.rand(*s, generator=g)
    N = lambda *s: torch.randn(*s, generator=g)
    lo, hi = math.log(scale_lo), math.log(scale_hi)
    if kind == "cube":
        means = (U(P, 3) * 2 - 1) * 1.3
        scales = torch.exp(U(P, 3) * (hi - lo) + lo)
        rot = torch.nn.functional.normalize(N(P, 4), dim=1)
    elif kind == "shell":
        d = torch.nn.functional.normalize(N(P, 3), dim=1)
        radius = torch.tensor([0.6, 0.9, 1.2])[torch.randint(0, 3, (P,), generator=g)] + 0.02 * N(P)
        means = d * radius[:, None]
        tang = torch.exp(U(P, 2) * (hi - lo) + lo)
        scales = torch.cat([tang, tang.min(dim=1, keepdim=True).values * 0.25], dim=1)
        # rotate local z onto the sphere normal d: q = (1 + z.d, z x d) normalised
        z = torch.tensor([0.0, 0.0, 1.0]).expand(P, 3)
        rot = torch.cat([(1 + (z * d).sum(1, keepdim=True)), torch.cross(z, d, dim=1)], dim=1)
        rot = torch.nn.functional.normalize(rot + 1e-8, dim=1)
    else:
        raise ValueError(kind)
    opac = torch.sigmoid(N(P, 1) * 2.0)
    shs = torch.zeros(P, M, 3)
    shs[:, 0, :] = (U(P, 3) - 0.5) / SH_C0
    if M > 1:
        shs[:, 1:, :] = N(P, M - 1, 3) * 0.05
    return Scene(means.contiguous(), scales.contiguous(), rot.contiguous(), opac.contiguous(), shs.contiguous(),
                 sh_degree)


def make_pixel_grads(width: int, height: int, seed: int = 1):
    """Upstream gradients dL/dcolor (3,H,W), dL/ddepth (1,H,W), dL/dalpha (1,H,W)."""
    g = torch.Generator().manual_seed(seed)
    n = width * height
    return (torch.randn(3, height, width, generator=g) / (3 * n),
            torch.randn(1, height, width, generator=g) / n,
            torch.randn(1, height, width, generator=g) / n)


# BASELINE.json configs (SURVEY.md §8 sizes)
CONFIGS = {
    "plumbing": dict(P=10_000, width=400, height=400, fovx=0.6911),
    "lego": dict(P=200_000, width=800, height=800, fovx=0.6911),
    "fern": dict(P=300_000, width=1008, height=756, fovx=2 * math.atan(1008 / (2 * 815.0))),
    "dtu": dict(P=1_000_000, width=1600, height=1200, fovx=2 * math.atan(800 / 2892.0)),
}
