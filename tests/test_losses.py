"""Photometric loss: oracle pinned on the reference's own outputs; CUDA parity through the C ABI."""
import os

import numpy as np
import pytest

from oracle import loss_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_loss.npz")
CASES = ["a", "b", "c"]


@pytest.mark.parametrize("case", CASES)
def test_loss_oracle_matches_reference_golden(case):
    g = np.load(GOLD)
    loss, l1, ssim, grad = loss_oracle.l1_ssim_loss_grad(g[f"{case}_img"], g[f"{case}_gt"], 0.2)
    assert abs(l1 - g[f"{case}_l1"]) < 1e-6
    assert abs(ssim - g[f"{case}_ssim"]) < 1e-5
    assert abs(loss - g[f"{case}_loss"]) < 1e-5
    ref = g[f"{case}_grad"].astype(np.float64)
    assert np.abs(grad - ref).max() <= 1e-4 * np.abs(ref).max()


def test_loss_oracle_gradient_is_the_derivative():
    rng = np.random.default_rng(5)
    y = rng.random((2, 14, 17))
    x = np.clip(y + 0.1 * rng.standard_normal(y.shape), 0.01, 0.99)
    _, _, _, grad = loss_oracle.l1_ssim_loss_grad(x, y, 0.2)
    for _ in range(6):
        i = tuple(rng.integers(0, s) for s in x.shape)
        e = np.zeros_like(x); e[i] = 1e-6
        num = (loss_oracle.l1_ssim_loss_grad(x + e, y, 0.2)[0] - loss_oracle.l1_ssim_loss_grad(x - e, y, 0.2)[0]) / 2e-6
        assert abs(num - grad[i]) < 1e-6 * max(1.0, abs(grad[i]) * 1e3)


# ---------------------------------------------------------------------------------------------- GPU
def _cuda_loss(img, gt, lam):
    import torch
    from gs_localization_b200 import _lib
    lib = _lib.load()
    x = torch.from_numpy(np.ascontiguousarray(img, np.float32)).cuda()
    y = torch.from_numpy(np.ascontiguousarray(gt, np.float32)).cuda()
    C, H, W = x.shape
    loss = torch.zeros(1, device="cuda")
    grad = torch.full_like(x, float("nan"))
    scratch = torch.empty(3 * x.numel() + 2, device="cuda")
    _lib.check(lib.gsr_l1_ssim_loss_grad(x.data_ptr(), y.data_ptr(), C, H, W, lam, loss.data_ptr(), grad.data_ptr(),
                                         scratch.data_ptr(), torch.cuda.current_stream().cuda_stream), "loss")
    torch.cuda.synchronize()
    return loss.item(), grad.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_loss_matches_reference_golden(case):
    g = np.load(GOLD)
    loss, grad = _cuda_loss(g[f"{case}_img"], g[f"{case}_gt"], 0.2)
    assert abs(loss - g[f"{case}_loss"]) < 1e-5               # tolerance: 1e-5 absolute on the scalar
    ref = g[f"{case}_grad"]
    assert np.abs(grad - ref).max() <= 1e-4 * np.abs(ref).max()  # 1e-4 of the largest gradient entry


@pytest.mark.gpu
@pytest.mark.parametrize("shape,lam", [((3, 480, 640), 0.2), ((3, 1, 1), 0.2), ((3, 17, 16), 0.0), ((1, 33, 47), 1.0)])
def test_cuda_loss_matches_oracle(shape, lam):
    rng = np.random.default_rng(11)
    y = rng.random(shape).astype(np.float32)
    x = np.clip(y + 0.1 * rng.standard_normal(shape), 0, 1).astype(np.float32)
    if x[0].size > 1:
        x[..., 0, 0] = y[..., 0, 0]                            # an exact tie: sign(0) = 0 like torch
    want_loss, _, _, want = loss_oracle.l1_ssim_loss_grad(x, y, lam)
    loss, grad = _cuda_loss(x, y, lam)
    assert abs(loss - want_loss) < 2e-5
    assert np.abs(grad - want).max() <= 1e-4 * np.abs(want).max() + 1e-9


@pytest.mark.gpu
def test_loss_autograd_wrapper_feeds_rasterizer():
    import torch
    from gs_localization_b200.losses import l1_ssim_loss
    rng = np.random.default_rng(3)
    y = torch.from_numpy(rng.random((3, 40, 56)).astype(np.float32)).cuda()
    x = (y + 0.1 * torch.randn_like(y)).clamp(0, 1).requires_grad_(True)
    loss = l1_ssim_loss(x * 1.0, y, 0.2)
    (2.0 * loss).backward()
    want_loss, _, _, want = loss_oracle.l1_ssim_loss_grad(x.detach().cpu().numpy(), y.cpu().numpy(), 0.2)
    assert abs(loss.item() - want_loss) < 2e-5
    assert np.abs(x.grad.cpu().numpy() - 2.0 * want).max() <= 1e-4 * np.abs(2.0 * want).max()


# ------------------------------------------------------------------- tracking loss (descent_utils.py:85-123)
TGOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_tracking.npz")
TCASES = ["rgbd", "mono", "rgbd_default"]


def _tracking_inputs(g, case):
    mono = bool(g[f"{case}_mono"])
    return dict(image=g[f"{case}_image"], depth=g[f"{case}_depth"], opacity=g[f"{case}_opacity"], gt_image=g[f"{case}_gt"],
                gt_depth=None if mono else g[f"{case}_gt_depth"], grad_mask=g[f"{case}_grad_mask"], exposure=g[f"{case}_exposure"],
                opacity_threshold=0.5, depth_weight=float(g[f"{case}_depth_weight"]))


def _check_tracking(got, g, case):
    loss, dI, dD, dE = got
    assert abs(loss - g[f"{case}_loss"]) < 1e-6
    assert np.abs(dI - g[f"{case}_dimage"]).max() <= 1e-5 * np.abs(g[f"{case}_dimage"]).max()
    assert np.abs(np.reshape(dD, g[f"{case}_ddepth"].shape) - g[f"{case}_ddepth"]).max() <= 1e-5 * max(np.abs(g[f"{case}_ddepth"]).max(), 1e-9)
    assert np.abs(dE - g[f"{case}_dexposure"]).max() <= 1e-4 * np.abs(g[f"{case}_dexposure"]).max() + 1e-7


@pytest.mark.parametrize("case", TCASES)
def test_tracking_oracle_matches_reference_golden(case):
    g = np.load(TGOLD)
    _check_tracking(loss_oracle.tracking_loss_grad(**_tracking_inputs(g, case)), g, case)


def _cuda_tracking(image, depth, opacity, gt_image, gt_depth, grad_mask, exposure, opacity_threshold, depth_weight):
    import torch
    from gs_localization_b200 import _lib
    lib = _lib.load()
    dev = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()
    I, D, O, G, GD, GM, E = map(dev, (image, depth, opacity, gt_image, gt_depth, grad_mask, exposure))
    H, W = I.shape[1:]
    loss, dI, dD, dE = torch.zeros(1, device="cuda"), torch.full_like(I, float("nan")), torch.full_like(D, float("nan")), torch.zeros(2, device="cuda")
    p = lambda t: 0 if t is None else t.data_ptr()
    _lib.check(lib.gsr_tracking_loss_grad(p(I), p(D), p(O), p(G), p(GD), p(GM), p(E), H, W, opacity_threshold, depth_weight,
                                          p(loss), p(dI), p(dD), p(dE), torch.cuda.current_stream().cuda_stream), "tracking")
    torch.cuda.synchronize()
    return loss.item(), dI.cpu().numpy(), dD.cpu().numpy(), dE.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("case", TCASES)
def test_cuda_tracking_loss_matches_reference_golden(case):
    g = np.load(TGOLD)
    _check_tracking(_cuda_tracking(**_tracking_inputs(g, case)), g, case)


@pytest.mark.gpu
def test_cuda_tracking_loss_full_size_and_defaults():
    rng = np.random.default_rng(2)
    H, W = 480, 640
    gt = rng.random((3, H, W)).astype(np.float32)
    img = np.clip(gt + 0.1 * rng.standard_normal(gt.shape), 0, 1).astype(np.float32)
    gd = (rng.random((H, W)) * 4).astype(np.float32)
    dep = (gd[None] + 0.1 * rng.standard_normal((1, H, W))).astype(np.float32)
    opa = rng.random((1, H, W)).astype(np.float32)
    for kw in (dict(gt_depth=gd, grad_mask=(rng.random((1, H, W)) > 0.5), exposure=np.float32([0.1, 0.05])),
               dict(gt_depth=None, grad_mask=None, exposure=None)):
        want = loss_oracle.tracking_loss_grad(img, dep, opa, gt, kw["gt_depth"], kw["grad_mask"],
                                              (0.0, 0.0) if kw["exposure"] is None else kw["exposure"], 0.5, 0.02)
        got = _cuda_tracking(img, dep, opa, gt, kw["gt_depth"], kw["grad_mask"], kw["exposure"], 0.5, 0.02)
        assert abs(got[0] - want[0]) < 1e-5
        assert np.abs(got[1] - want[1]).max() <= 1e-5 * np.abs(want[1]).max()
        assert np.abs(got[2] - want[2]).max() <= 1e-5 * max(np.abs(want[2]).max(), 1e-9)
        if kw["exposure"] is not None:
            assert np.abs(got[3] - want[3]).max() <= 1e-3 * np.abs(want[3]).max()   # float32 sums of 9.2e5 signed terms


def test_grad_mask_matches_reference_golden():
    import torch
    from gs_localization_b200 import tracking
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_grad_mask.npz"))
    img = torch.from_numpy(g["img"])
    gv, gh = tracking.image_gradient(img)
    assert np.abs(gv.numpy() - g["grad_v"]).max() < 1e-6 and np.abs(gh.numpy() - g["grad_h"]).max() < 1e-6
    assert np.array_equal(tracking.compute_grad_mask(img, 1.1, "tum").numpy(), g["tum"])
    assert np.array_equal(tracking.compute_grad_mask(img, 4, "replica").numpy(), g["replica"])


# ----------------------------------------------------------- depth terms of the training loss (7scenes_gs_full_dslam.py:168-184)
def _depth_case(seed, n=(48, 64)):
    rng = np.random.default_rng(seed)
    gt = (1.0 + 3.0 * rng.random(n)).astype(np.float32)
    gt[rng.random(n) < 0.15] = 0.0
    true = np.where(gt > 0, gt, 2.0)
    pseudo = (800.0 / true + 20.0 * rng.standard_normal(n)).astype(np.float32)      # MiDaS-like inverse depth
    depth = (true + 0.2 * rng.standard_normal(n)).astype(np.float32)
    return depth, pseudo, gt


def test_depth_loss_oracle_vs_corrcoef_and_autograd():
    import torch
    depth, pseudo, gt = _depth_case(0)
    loss, grad = loss_oracle.depth_loss_grad(depth, pseudo, gt)
    d64 = depth.astype(np.float64).ravel()
    r1 = np.corrcoef(-pseudo.astype(np.float64).ravel(), d64)[0, 1]
    r2 = np.corrcoef((np.float32(1000) / (pseudo + np.float32(200))).astype(np.float64).ravel(), d64)[0, 1]
    l1 = np.abs((d64 - gt.ravel()) * (gt.ravel() > 0)).mean()
    assert abs(loss - (0.01 * min(1 - r1, 1 - r2) + 0.05 * l1)) < 1e-12
    # autograd of the same expression written with torch.corrcoef
    d = torch.from_numpy(d64).requires_grad_(True)
    m = torch.from_numpy(pseudo.astype(np.float64).ravel())
    g = torch.from_numpy(gt.astype(np.float64).ravel())
    cands = [1 - torch.corrcoef(torch.stack([-m, d]))[0, 1],
             1 - torch.corrcoef(torch.stack([torch.from_numpy((np.float32(1000) / (pseudo + np.float32(200))).astype(np.float64).ravel()), d]))[0, 1]]
    mask = (g > 0).double()
    L = 0.01 * min(cands) + 0.05 * (d * mask - g * mask).abs().mean()
    L.backward()
    assert abs(L.item() - loss) < 1e-12
    assert np.abs(d.grad.numpy() - grad.ravel()).max() <= 1e-9 * np.abs(grad).max()


@pytest.mark.gpu
@pytest.mark.parametrize("use_pseudo,use_gt,n", [(True, True, (480, 640)), (True, False, (33, 7)), (False, True, (5, 5)), (True, True, (840, 1297))])
def test_cuda_depth_loss_matches_oracle(use_pseudo, use_gt, n):
    import torch
    from gs_localization_b200 import losses
    depth, pseudo, gt = _depth_case(3, n)
    want_loss, want = loss_oracle.depth_loss_grad(depth, pseudo if use_pseudo else None, gt if use_gt else None)
    d = torch.from_numpy(depth).cuda().requires_grad_(True)
    L = losses.depth_loss(d, torch.from_numpy(pseudo).cuda() if use_pseudo else None, torch.from_numpy(gt).cuda() if use_gt else None)
    L.backward()
    assert abs(L.item() - want_loss) < 1e-6
    assert np.abs(d.grad.cpu().numpy() - want).max() <= 1e-4 * np.abs(want).max()
