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
