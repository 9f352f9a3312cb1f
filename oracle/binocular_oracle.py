"""TEST INFRASTRUCTURE — float64 numpy restatement of the reference's binocular-consistency
loss.

Follows utils/graphics_utils.py:80-125 (inverse_warp_images: horizontal bilinear gather at
column + disparity, both taps clamped, pixels whose either tap leaves the image set to 0),
utils/loss_utils.py:18-21 (masked l1_loss), :68-91 (SmoothLoss: central differences
[-0.5, 0, 0.5] without padding, image edges summed over channels, exp(-0.33 |.|)) and the
combination at train.py:128-136:

    disparity = focal_x * (-trans_dist) / (depth + 1e-5)
    warped    = inverse_warp_images(shifted_image, disparity)
    mask      = inverse_warp_images(ones,          disparity)
    loss      = l1_loss(warped, gt, mask) + 0.05 * SmoothLoss(disparity * mask, gt)

The analytic gradients (to the shifted image and to the depth) are derived from the same
formulas; floor() and the validity mask carry no gradient, and d mask / d disparity is
(-1) + (+1) = 0 exactly as in the reference's autograd graph.  Pinned against the
reference's own functions by tests/golden/make_binocular_golden.py (run on CPU in the build
container with Tensor.cuda() stubbed to the identity) -> tests/golden/bino_*.npz.
"""
import numpy as np

EPS = 1e-5
EDGE_K = 0.33


def _taps(disparity):
    """disparity (H,W) -> c0, c1 (clamped int columns), w0, w1, valid."""
    H, W = disparity.shape
    cols = np.arange(W)[None, :]
    x0 = np.floor(disparity).astype(np.int64)
    x1 = x0 + 1
    c0, c1 = cols + x0, cols + x1
    valid = (c0 >= 0) & (c0 < W) & (c1 >= 0) & (c1 < W)
    return np.clip(c0, 0, W - 1), np.clip(c1, 0, W - 1), x1 - disparity, disparity - x0, valid


def inverse_warp(image, disparity):
    """image (C,H,W), disparity (H,W) -> warped (C,H,W).  graphics_utils.py:80-125."""
    image, disparity = image.astype(np.float64), disparity.astype(np.float64)
    c0, c1, w0, w1, valid = _taps(disparity)
    rows = np.arange(disparity.shape[0])[:, None]
    out = w0[None] * image[:, rows, c0] + w1[None] * image[:, rows, c1]
    return out * valid[None]


def inverse_warp_grad(image, disparity, g_out):
    """Gradients of sum(g_out * warped) w.r.t. image and disparity."""
    image, disparity, g_out = (a.astype(np.float64) for a in (image, disparity, g_out))
    C, H, W = image.shape
    c0, c1, w0, w1, valid = _taps(disparity)
    rows = np.broadcast_to(np.arange(H)[:, None], (H, W))
    g = g_out * valid[None]
    g_img = np.zeros_like(image)
    for ch in range(C):
        np.add.at(g_img[ch], (rows, c0), g[ch] * w0)
        np.add.at(g_img[ch], (rows, c1), g[ch] * w1)
    g_disp = (g * (image[:, rows, c1] - image[:, rows, c0])).sum(0)
    return g_img, g_disp


def _edges(image):
    """exp(-0.33 |sum_ch central difference|) along x and y, (H-2, W-2) each."""
    image = image.astype(np.float64)
    ex = (0.5 * (image[:, 1:-1, 2:] - image[:, 1:-1, :-2])).sum(0)
    ey = (0.5 * (image[:, 2:, 1:-1] - image[:, :-2, 1:-1])).sum(0)
    return np.exp(-EDGE_K * np.abs(ex)), np.exp(-EDGE_K * np.abs(ey))


def smooth_loss(disparity, image):
    """SmoothLoss.forward(disparity (H,W), image (3,H,W)).  loss_utils.py:68-91."""
    d = disparity.astype(np.float64)
    wx, wy = _edges(image)
    dx = 0.5 * (d[1:-1, 2:] - d[1:-1, :-2])
    dy = 0.5 * (d[2:, 1:-1] - d[:-2, 1:-1])
    return float(np.abs(wx * dx).mean() + np.abs(wy * dy).mean())


def smooth_loss_grad(disparity, image):
    """d smooth_loss / d disparity, (H,W)."""
    d = disparity.astype(np.float64)
    H, W = d.shape
    wx, wy = _edges(image)
    n = float((H - 2) * (W - 2))
    sx = np.sign(wx * 0.5 * (d[1:-1, 2:] - d[1:-1, :-2])) * wx * 0.5 / n
    sy = np.sign(wy * 0.5 * (d[2:, 1:-1] - d[:-2, 1:-1])) * wy * 0.5 / n
    g = np.zeros_like(d)
    g[1:-1, 2:] += sx
    g[1:-1, :-2] -= sx
    g[2:, 1:-1] += sy
    g[:-2, 1:-1] -= sy
    return g


def disparity_of(depth, focal_x, trans_dist):
    return focal_x * (-trans_dist) / (depth.astype(np.float64) + EPS)


def binocular_terms(shifted, depth, gt, focal_x, trans_dist):
    """depth (H,W) -> (masked L1 term, smoothness term)."""
    disp = disparity_of(depth, focal_x, trans_dist)
    warped = inverse_warp(shifted, disp)
    mask = inverse_warp(np.ones((1,) + disp.shape), disp)[0]
    l1 = float(np.abs(warped * mask[None] - gt.astype(np.float64) * mask[None]).mean())
    return l1, smooth_loss(disp * mask, gt)


def binocular_loss(shifted, depth, gt, focal_x, trans_dist, smooth_weight=0.05):
    l1, sm = binocular_terms(shifted, depth, gt, focal_x, trans_dist)
    return l1 + smooth_weight * sm


def binocular_loss_grad(shifted, depth, gt, focal_x, trans_dist, smooth_weight=0.05):
    """(d loss / d shifted (3,H,W), d loss / d depth (H,W))."""
    depth64, gt64 = depth.astype(np.float64), gt.astype(np.float64)
    disp = disparity_of(depth, focal_x, trans_dist)
    warped = inverse_warp(shifted, disp)
    mask = inverse_warp(np.ones((1,) + disp.shape), disp)[0]
    g_warped = np.sign(warped * mask[None] - gt64 * mask[None]) * mask[None] / warped.size
    g_img, g_disp = inverse_warp_grad(shifted, disp, g_warped)
    g_disp = g_disp + smooth_weight * smooth_loss_grad(disp * mask, gt) * mask
    return g_img, g_disp * (-disp / (depth64 + EPS))
