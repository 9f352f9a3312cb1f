"""TEST INFRASTRUCTURE — float64 numpy restatement of the reference's photometric loss.

Follows utils/loss_utils.py: gaussian/create_window :26-34 (the 1-D window is built in
float32 exactly as the reference does, the 2-D window is its float32 outer product),
_ssim :46-66 (five depthwise correlations with zero padding 5, C1 = 0.01^2, C2 = 0.03^2,
mean of the map), l1_loss :18-21, and the combination at train.py:146-147.  The
analytic gradient w.r.t. the first image is derived from the same formulas.  Pinned
against the reference's own functions by tests/golden/make_loss_golden.py (run on CPU in
the build container) -> tests/golden/loss_*.npz.
"""
import numpy as np


# The float32 values utils/loss_utils.py:26-28 gaussian(11, 1.5) produces (torch.Tensor of
# the python doubles, divided by its float32 sum); tests/test_loss.py re-derives them with
# torch.
_WINDOW = np.array([0.001028380123898387, 0.0075987582094967365, 0.036000773310661316, 0.10936068743467331,
                    0.21300552785396576, 0.26601171493530273, 0.21300552785396576, 0.10936068743467331,
                    0.036000773310661316, 0.0075987582094967365, 0.001028380123898387], dtype=np.float32)


def window_1d():
    return _WINDOW.copy()


def window_2d():
    g = window_1d()
    return np.outer(g, g).astype(np.float32).astype(np.float64)  # float32 products, like .mm().float()


def _corr(img, w2):
    """Depthwise 11x11 correlation, zero padding 5.  img: (C,H,W) float64."""
    C, H, W = img.shape
    p = np.zeros((C, H + 10, W + 10))
    p[:, 5:5 + H, 5:5 + W] = img
    out = np.zeros_like(img)
    for dy in range(11):
        for dx in range(11):
            out += w2[dy, dx] * p[:, dy:dy + H, dx:dx + W]
    return out


def ssim_terms(x, y):
    w2 = window_2d()
    x, y = x.astype(np.float64), y.astype(np.float64)
    mu1, mu2 = _corr(x, w2), _corr(y, w2)
    e11, e22, e12 = _corr(x * x, w2), _corr(y * y, w2), _corr(x * y, w2)
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    s1, s2, s12 = e11 - mu1 * mu1, e22 - mu2 * mu2, e12 - mu1 * mu2
    A, B = 2 * mu1 * mu2 + C1, 2 * s12 + C2
    Cc, Dd = mu1 * mu1 + mu2 * mu2 + C1, s1 + s2 + C2
    S = A * B / (Cc * Dd)
    return S, (mu1, mu2, A, B, Cc, Dd, w2)


def ssim(x, y):
    return float(ssim_terms(x, y)[0].mean())


def ssim_grad(x, y):
    """d mean(SSIM) / dx."""
    S, (mu1, mu2, A, B, Cc, Dd, w2) = ssim_terms(x, y)
    x, y = x.astype(np.float64), y.astype(np.float64)
    inv = 1.0 / (Cc * Dd)
    d_mu1 = 2 * mu2 * (B - A) * inv - S * 2 * mu1 * (1 / Cc - 1 / Dd)
    d_e11 = -S / Dd
    d_e12 = 2 * A * inv
    g = _corr(d_mu1, w2) + 2 * x * _corr(d_e11, w2) + y * _corr(d_e12, w2)  # symmetric window
    return g / x.size


def l1(x, y):
    return float(np.abs(x.astype(np.float64) - y.astype(np.float64)).mean())


def photometric(x, y, lam=0.2):
    return (1 - lam) * l1(x, y) + lam * (1 - ssim(x, y))


def photometric_grad(x, y, lam=0.2):
    return (1 - lam) * np.sign(x.astype(np.float64) - y.astype(np.float64)) / x.size - lam * ssim_grad(x, y)
