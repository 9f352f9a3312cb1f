"""TEST INFRASTRUCTURE — numpy (float64) restatement of the per-step parameter plumbing
(SURVEY.md §8(f) rank 3).

Follows scene/gaussian_model.py:95-115 (activations: torch.exp, torch.sigmoid,
torch.nn.functional.normalize with eps 1e-12, torch.cat of the SH features), :154-163 +
torch/optim/adam.py (Adam, betas (0.9, 0.999), eps 1e-15, lr per group, no weight decay /
amsgrad), :307-309 + utils/general_utils.py:18-19 (opacity decay through inverse_sigmoid),
:409-411 + train.py:170-171 (densification statistics on radii > 0).

The arithmetic lives in torch (a dependency of the reference, pinned there to 2.1.1; 2.11
here): pinned by tests/golden/make_parameters_golden.py, which runs the reference's
statements with torch on CPU -> tests/golden/params_*.npz.
"""
import numpy as np


def activate(f_dc, f_rest, opacity_raw, scaling_raw, rotation_raw):
    f_dc, f_rest, o, s, q = (a.astype(np.float64) for a in (f_dc, f_rest, opacity_raw, scaling_raw, rotation_raw))
    n = np.maximum(np.sqrt((q * q).sum(1, keepdims=True)), 1e-12)
    return np.concatenate([f_dc, f_rest], axis=1), 1.0 / (1.0 + np.exp(-o)), np.exp(s), q / n


def activate_grad(opacity_raw, scaling_raw, rotation_raw, g_shs, g_opacities, g_scales, g_rotations):
    o, s, q, g_shs, g_o, g_s, g_q = (a.astype(np.float64) for a in (opacity_raw, scaling_raw, rotation_raw, g_shs,
                                                                    g_opacities, g_scales, g_rotations))
    sig = 1.0 / (1.0 + np.exp(-o))
    norm = np.sqrt((q * q).sum(1, keepdims=True))
    y = q / np.maximum(norm, 1e-12)
    g_rot = np.where(norm > 1e-12, (g_q - y * (y * g_q).sum(1, keepdims=True)) / np.maximum(norm, 1e-12), g_q * 1e12)
    return g_shs[:, :1], g_shs[:, 1:], g_o * sig * (1 - sig), g_s * np.exp(s), g_rot


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-15):
    """One torch.optim.Adam update (step counts from 1).  Returns new (p, m, v)."""
    p, g, m, v = (a.astype(np.float64) for a in (p, g, m, v))
    m = m + (g - m) * (1 - beta1)
    v = v * beta2 + (1 - beta2) * g * g
    step_size = lr / (1 - beta1 ** step)
    denom = np.sqrt(v) / np.sqrt(1 - beta2 ** step) + eps
    return p - step_size * m / denom, m, v


def opacity_decay(opacity_raw, factor=0.99):
    y = factor / (1.0 + np.exp(-opacity_raw.astype(np.float64)))
    return np.log(y / (1 - y))


def densify_stats(viewspace_grad, radii, accum, denom, max_radii2D):
    vis = radii > 0
    g = viewspace_grad.astype(np.float64)
    accum, denom, max_radii2D = (a.astype(np.float64).copy() for a in (accum, denom, max_radii2D))
    accum[vis] += np.sqrt(g[vis, 0] ** 2 + g[vis, 1] ** 2)
    denom[vis] += 1
    max_radii2D[vis] = np.maximum(max_radii2D[vis], radii[vis])
    return accum, denom, max_radii2D
