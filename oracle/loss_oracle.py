"""TEST INFRASTRUCTURE ONLY — numpy float64 restatement of the reference's photometric loss and its gradient.

Follows gaussian_splatting/utils/loss_utils.py: l1_loss (17-18), gaussian (23-25), create_window (27-31),
_ssim (41-63: zero-padded grouped conv2d with the 11x11 window, C1 = 0.01^2, C2 = 0.03^2, mean), combined as
gs_localization/gs/7scenes_gs_full_dslam.py:165-166.  Pinned against tests/golden/ref_loss.npz, produced by
importing the reference module itself (tests/golden/make_loss_golden.py).  Never imported by the product path."""
import numpy as np


def window_1d(size=11, sigma=1.5):
    g = np.exp(-((np.arange(size) - size // 2) ** 2) / (2.0 * sigma ** 2))
    return g / g.sum()


def _filter(img, w):
    """zero-padded separable correlation of a [C,H,W] array with the outer product w w^T"""
    r = len(w) // 2
    C, H, W = img.shape
    p = np.zeros((C, H + 2 * r, W + 2 * r))
    p[:, r:r + H, r:r + W] = img
    h = sum(w[k] * p[:, :, k:k + W] for k in range(len(w)))
    return sum(w[k] * h[:, k:k + H, :] for k in range(len(w)))


def l1_ssim_loss_grad(img, gt, lam=0.2):
    """returns (loss, l1, ssim, dloss/dimg) in float64"""
    x, y = np.asarray(img, np.float64), np.asarray(gt, np.float64)
    w = window_1d()
    n = x.size
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    mu1, mu2 = _filter(x, w), _filter(y, w)
    e11, e22, e12 = _filter(x * x, w), _filter(y * y, w), _filter(x * y, w)
    s1, s2, s12 = e11 - mu1 ** 2, e22 - mu2 ** 2, e12 - mu1 * mu2
    A1, A2, B1, B2 = 2 * mu1 * mu2 + C1, 2 * s12 + C2, mu1 ** 2 + mu2 ** 2 + C1, s1 + s2 + C2
    ssim_map = A1 * A2 / (B1 * B2)
    l1, ssim = np.abs(x - y).mean(), ssim_map.mean()
    loss = (1 - lam) * l1 + lam * (1 - ssim)
    # chain rule through the five filtered moments (the window is symmetric, so the adjoint filter is the filter)
    g = -lam / n
    d_mu1 = 2 * mu2 * (A2 - A1) / (B1 * B2) - ssim_map * (2 * mu1 / B1 - 2 * mu1 / B2)
    d_e11 = -ssim_map / B2
    d_e12 = 2 * A1 / (B1 * B2)
    grad = _filter(g * d_mu1, w) + 2 * x * _filter(g * d_e11, w) + y * _filter(g * d_e12, w)
    grad += (1 - lam) / n * np.sign(x - y)
    return loss, l1, ssim, grad


def tracking_loss_grad(image, depth, opacity, gt_image, gt_depth, grad_mask, exposure=(0.0, 0.0), opacity_threshold=0.5,
                       depth_weight=0.02):
    """get_loss_tracking of gs_localization/pipelines/tools/descent_utils.py:85-123 and its gradient, float64.
    image, gt_image [3,H,W]; depth, opacity [1,H,W]; gt_depth [H,W] or None (monocular); grad_mask [1,H,W] or None.
    Returns (loss, dL/dimage, dL/ddepth, dL/d(exposure_a, exposure_b)).  Pinned against tests/golden/ref_tracking.npz."""
    I, G = np.asarray(image, np.float64), np.asarray(gt_image, np.float64)
    H, W = I.shape[1:]
    gm = np.ones((1, H, W)) if grad_mask is None else np.asarray(grad_mask, np.float64).reshape(1, H, W)
    om = (np.asarray(opacity, np.float32).reshape(1, H, W) > np.float32(opacity_threshold)).astype(np.float64)
    a, b = float(exposure[0]), float(exposure[1])
    ea = np.exp(a)
    d = (ea * I + b) * gm - G * gm
    loss = (om * np.abs(d)).mean()
    sg = np.sign(d) * om * gm / d.size
    dI = sg * ea
    da, db = (sg * ea * I).sum(), sg.sum()
    dD = np.zeros((1, H, W))
    if gt_depth is not None:
        gd = np.asarray(gt_depth, np.float64).reshape(1, H, W)
        D = np.asarray(depth, np.float64).reshape(1, H, W)
        dm = (np.asarray(gt_depth, np.float32).reshape(1, H, W) > np.float32(0.01)) * om * gm
        dd = D * dm - gd * dm
        loss += depth_weight * np.abs(dd).mean()
        dD = depth_weight * np.sign(dd) * dm / dd.size
    return loss, dI, dD, np.array([da, db])


def depth_loss_grad(depth, pseudo, gt, inv_numerator=1000.0, w_pearson=0.01, w_l1=0.05):
    """Depth terms of the map-training loss (gs_localization/gs/7scenes_gs_full_dslam.py:168-184), float64:
    w_p * min(1 - r(-m, d), 1 - r(k/(m+200), d)) + w_l1 * mean|d*mask - gt*mask|, mask = gt > 0.
    r is torchmetrics' pearson_corrcoef in the reference — a third-party dependency that is neither vendored nor
    version-pinned there and is absent from this image: PARITY UNPINNED against it; restated from the definition of
    the Pearson coefficient and checked against np.corrcoef and torch autograd in tests/test_losses.py."""
    d = np.asarray(depth, np.float64).ravel()
    n = d.size
    loss, grad = 0.0, np.zeros(n)
    if pseudo is not None:
        m32 = np.asarray(pseudo, np.float32).ravel()
        cands = [-m32.astype(np.float64), (np.float32(inv_numerator) / (m32 + np.float32(200.0))).astype(np.float64)]
        dc = d - d.mean()
        Sdd = (dc * dc).sum()
        best = None
        for x in cands:
            xc = x - x.mean()
            r = (xc * dc).sum() / np.sqrt((xc * xc).sum() * Sdd)
            if best is None or (1 - r) < best[0]:
                best = (1 - r, r, xc)
        l, r, xc = best
        loss += w_pearson * l
        grad += -w_pearson * (xc / np.sqrt((xc * xc).sum() * Sdd) - r * dc / Sdd)
    if gt is not None:
        g = np.asarray(gt, np.float64).ravel()
        mask = np.asarray(gt, np.float32).ravel() > 0
        diff = (d - g) * mask
        loss += w_l1 * np.abs(diff).mean()
        grad += w_l1 * np.sign(diff) / n
    return loss, grad.reshape(np.shape(depth))
