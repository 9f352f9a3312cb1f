"""TEST INFRASTRUCTURE — numpy front end of oracle/liboracle.so (the CPU restatement).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline /
``--impl reference`` legs) may import this module.  See oracle.c for what each function
restates (reference file:line) and for the parity-pinning statement.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "cpu"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.orc_binning.restype = ctypes.c_longlong
        _lib.orc_higher_msb.restype = ctypes.c_uint32
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int):
    lib().orc_set_num_threads(ctypes.c_int(n))


def higher_msb(n: int) -> int:
    return int(lib().orc_higher_msb(ctypes.c_uint32(n)))


def preprocess(means3D, scales, rotations, opacities, shs, viewmatrix, projmatrix, campos, W, H, tanfovx, tanfovy,
               sh_degree, scale_modifier=1.0, cov3D_precomp=None, colors_precomp=None):
    means3D = _f(means3D)
    P = means3D.shape[0]
    scales, rotations, opacities, shs = _f(scales), _f(rotations), _f(opacities), _f(shs)
    cov3D_precomp, colors_precomp = _f(cov3D_precomp), _f(colors_precomp)
    V, Pm, cam = _f(viewmatrix).reshape(-1), _f(projmatrix).reshape(-1), _f(campos)
    M = 0 if shs is None else shs.shape[1]
    out = dict(
        radii=np.zeros(P, np.int32), means2D=np.zeros((P, 2), np.float32), depths=np.zeros(P, np.float32),
        cov3D=np.zeros((P, 6), np.float32), rgb=np.zeros((P, 3), np.float32),
        conic_opacity=np.zeros((P, 4), np.float32), tiles_touched=np.zeros(P, np.uint32),
        clamped=np.zeros((P, 3), np.uint8),
    )
    lib().orc_preprocess(
        ctypes.c_int(P), ctypes.c_int(sh_degree), ctypes.c_int(M), _p(means3D), _p(scales),
        ctypes.c_float(scale_modifier), _p(rotations), _p(opacities), _p(shs), _p(cov3D_precomp), _p(colors_precomp),
        _p(V), _p(Pm), _p(cam), ctypes.c_int(W), ctypes.c_int(H), ctypes.c_float(tanfovx), ctypes.c_float(tanfovy),
        _p(out["radii"]), _p(out["means2D"]), _p(out["depths"]), _p(out["cov3D"]), _p(out["rgb"]),
        _p(out["conic_opacity"]), _p(out["tiles_touched"]), _p(out["clamped"]),
    )
    return out


def binning(means2D, depths, radii, W, H, want_keys=False):
    P = radii.shape[0]
    gx, gy = (W + 15) // 16, (H + 15) // 16
    means2D, depths = _f(means2D), _f(depths)
    radii = np.ascontiguousarray(radii, np.int32)
    # upper bound on R: recomputed inside; allocate from the rect areas
    # (cheap python-side estimate = call once with a generous buffer)
    r = np.maximum(radii, 0).astype(np.float32)
    x0 = np.clip(np.trunc((means2D[:, 0] - r) * np.float32(0.0625)), 0, gx)
    x1 = np.clip(np.trunc((means2D[:, 0] + r + np.float32(15.0)) * np.float32(0.0625)) + 1, 0, gx)
    y0 = np.clip(np.trunc((means2D[:, 1] - r) * np.float32(0.0625)), 0, gy)
    y1 = np.clip(np.trunc((means2D[:, 1] + r + np.float32(15.0)) * np.float32(0.0625)) + 1, 0, gy)
    cap = int(np.sum(np.where(radii > 0, (x1 - x0 + 1) * (y1 - y0 + 1), 0))) + 16
    point_list = np.zeros(cap, np.uint32)
    ranges = np.zeros((gx * gy, 2), np.uint32)
    keys = np.zeros(cap, np.uint64) if want_keys else None
    R = lib().orc_binning(ctypes.c_int(P), _p(means2D), _p(depths), _p(radii), ctypes.c_int(W), ctypes.c_int(H),
                          _p(point_list), _p(ranges), _p(keys))
    if R < 0:
        raise MemoryError("orc_binning")
    out = dict(R=int(R), point_list=point_list[:R], ranges=ranges)
    if want_keys:
        out["keys"] = keys[:R]
    return out


def render_forward(W, H, ranges, point_list, means2D, colors, depths, conic_opacity, bg):
    out_color = np.zeros((3, H, W), np.float32)
    out_depth = np.zeros((1, H, W), np.float32)
    out_alpha = np.zeros((1, H, W), np.float32)
    n_contrib = np.zeros(H * W, np.uint32)
    ranges = np.ascontiguousarray(ranges, np.uint32)
    point_list = np.ascontiguousarray(point_list, np.uint32)
    a = [_f(means2D), _f(colors), _f(depths), _f(conic_opacity), _f(bg)]
    lib().orc_render_forward(ctypes.c_int(W), ctypes.c_int(H), _p(ranges), _p(point_list), *[_p(t) for t in a],
                             _p(out_color), _p(out_depth), _p(out_alpha), _p(n_contrib))
    return dict(color=out_color, depth=out_depth, alpha=out_alpha, n_contrib=n_contrib)


def render_backward(P, W, H, ranges, point_list, bg, means2D, conic_opacity, colors, depths, alphas, n_contrib,
                    dL_dpix, dL_dpix_depth, dL_dalphas):
    out = dict(dL_dmean2D=np.zeros((P, 3), np.float32), dL_dconic=np.zeros((P, 4), np.float32),
               dL_dopacity=np.zeros((P, 1), np.float32), dL_dcolors=np.zeros((P, 3), np.float32),
               dL_ddepths=np.zeros((P, 1), np.float32))
    ranges = np.ascontiguousarray(ranges, np.uint32)
    point_list = np.ascontiguousarray(point_list, np.uint32)
    n_contrib = np.ascontiguousarray(n_contrib, np.uint32)
    a = [_f(bg), _f(means2D), _f(conic_opacity), _f(colors), _f(depths), _f(alphas)]
    b = [_f(dL_dpix), _f(dL_dpix_depth), _f(dL_dalphas)]
    lib().orc_render_backward(ctypes.c_int(P), ctypes.c_int(W), ctypes.c_int(H), _p(ranges), _p(point_list),
                              *[_p(t) for t in a], _p(n_contrib), *[_p(t) for t in b], _p(out["dL_dmean2D"]),
                              _p(out["dL_dconic"]), _p(out["dL_dopacity"]), _p(out["dL_dcolors"]),
                              _p(out["dL_ddepths"]))
    return out


def preprocess_backward(means3D, radii, shs, clamped, scales, rotations, scale_modifier, cov3D, viewmatrix,
                        projmatrix, W, H, tanfovx, tanfovy, campos, sh_degree, dL_dmean2D, dL_dconic, dL_dcolor,
                        dL_ddepth):
    means3D = _f(means3D)
    P = means3D.shape[0]
    shs, scales, rotations = _f(shs), _f(scales), _f(rotations)
    M = 0 if shs is None else shs.shape[1]
    out = dict(dL_dmeans3D=np.zeros((P, 3), np.float32), dL_dcov3D=np.zeros((P, 6), np.float32),
               dL_dsh=np.zeros((P, M, 3), np.float32), dL_dscales=np.zeros((P, 3), np.float32),
               dL_drotations=np.zeros((P, 4), np.float32))
    focal_y = np.float32(H) / (np.float32(2.0) * np.float32(tanfovy))
    focal_x = np.float32(W) / (np.float32(2.0) * np.float32(tanfovx))
    radii = np.ascontiguousarray(radii, np.int32)
    clamped = np.ascontiguousarray(clamped, np.uint8)
    a = [_f(cov3D), _f(viewmatrix).reshape(-1), _f(projmatrix).reshape(-1)]
    g = [_f(dL_dmean2D), _f(dL_dconic), _f(dL_dcolor), _f(dL_ddepth)]
    lib().orc_preprocess_backward(
        ctypes.c_int(P), ctypes.c_int(sh_degree), ctypes.c_int(M), _p(means3D), _p(radii), _p(shs), _p(clamped),
        _p(scales), _p(rotations), ctypes.c_float(scale_modifier), *[_p(t) for t in a], ctypes.c_float(focal_x),
        ctypes.c_float(focal_y), ctypes.c_float(tanfovx), ctypes.c_float(tanfovy), _p(_f(campos)),
        *[_p(t) for t in g], _p(out["dL_dmeans3D"]), _p(out["dL_dcov3D"]), _p(out["dL_dsh"] if M else None),
        _p(out["dL_dscales"]), _p(out["dL_drotations"]),
    )
    return out


def rasterize_forward(means3D, scales, rotations, opacities, shs, viewmatrix, projmatrix, campos, bg, W, H, tanfovx,
                      tanfovy, sh_degree, scale_modifier=1.0, cov3D_precomp=None, colors_precomp=None):
    """Whole forward (K1..K6).  Returns a dict with the images and every intermediate."""
    pre = preprocess(means3D, scales, rotations, opacities, shs, viewmatrix, projmatrix, campos, W, H, tanfovx,
                     tanfovy, sh_degree, scale_modifier, cov3D_precomp, colors_precomp)
    b = binning(pre["means2D"], pre["depths"], pre["radii"], W, H)
    img = render_forward(W, H, b["ranges"], b["point_list"], pre["means2D"], pre["rgb"], pre["depths"],
                         pre["conic_opacity"], bg)
    return {**pre, **b, **img}


def rasterize_backward(fwd, means3D, scales, rotations, shs, viewmatrix, projmatrix, campos, bg, W, H, tanfovx,
                       tanfovy, sh_degree, dL_dcolor, dL_ddepth, dL_dalpha, scale_modifier=1.0, cov3D_precomp=None,
                       colors_precomp=None):
    """Whole backward (K7..K9) from a forward dict."""
    P = np.asarray(means3D).shape[0]
    rb = render_backward(P, W, H, fwd["ranges"], fwd["point_list"], bg, fwd["means2D"], fwd["conic_opacity"],
                         fwd["rgb"], fwd["depths"], fwd["alpha"], fwd["n_contrib"], dL_dcolor, dL_ddepth, dL_dalpha)
    cov = fwd["cov3D"] if cov3D_precomp is None else cov3D_precomp
    pb = preprocess_backward(means3D, fwd["radii"], shs if colors_precomp is None else None, fwd["clamped"],
                             scales if cov3D_precomp is None else None, rotations if cov3D_precomp is None else None,
                             scale_modifier, cov, viewmatrix, projmatrix, W, H, tanfovx, tanfovy, campos, sh_degree,
                             rb["dL_dmean2D"], rb["dL_dconic"], rb["dL_dcolors"], rb["dL_ddepths"])
    return {**rb, **pb}
