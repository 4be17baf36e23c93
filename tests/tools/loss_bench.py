"""Times the fused L1+SSIM loss/gradient (gsr_l1_ssim_loss_grad) against the framework formulation the reference
uses (grouped conv2d + element-wise ops + autograd; loss_utils.py:41-63 restated in torch) at 640x480x3."""
import json
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from gs_localization_b200.losses import l1_ssim_loss  # noqa: E402


def torch_loss(img, gt, lam, window):
    C = img.shape[0]
    x, y = img[None], gt[None]
    mu1, mu2 = F.conv2d(x, window, padding=5, groups=C), F.conv2d(y, window, padding=5, groups=C)
    s1 = F.conv2d(x * x, window, padding=5, groups=C) - mu1 * mu1
    s2 = F.conv2d(y * y, window, padding=5, groups=C) - mu2 * mu2
    s12 = F.conv2d(x * y, window, padding=5, groups=C) - mu1 * mu2
    m = ((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 * mu1 + mu2 * mu2 + 1e-4) * (s1 + s2 + 9e-4))
    return (1 - lam) * (img - gt).abs().mean() + lam * (1 - m.mean())


def timed(fn, n=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    H, W, C = 480, 640, 3
    torch.manual_seed(0)
    gt = torch.rand(C, H, W, device="cuda")
    img = (gt + 0.1 * torch.randn_like(gt)).clamp(0, 1).requires_grad_(True)
    g1 = torch.exp(-((torch.arange(11.0) - 5) ** 2) / (2 * 1.5 ** 2))
    g1 = g1 / g1.sum()
    window = (g1[:, None] @ g1[None, :])[None, None].expand(C, 1, 11, 11).contiguous().cuda()

    def ours():
        img.grad = None
        l1_ssim_loss(img, gt, 0.2).backward()

    def ref():
        img.grad = None
        torch_loss(img, gt, 0.2, window).backward()

    ours(); go = img.grad.clone(); lo = l1_ssim_loss(img, gt, 0.2).item()
    ref(); gr = img.grad.clone(); lr = torch_loss(img, gt, 0.2, window).item()
    t_ours, t_ref = timed(ours), timed(ref)
    px = C * H * W
    print(json.dumps({"tool": "loss_bench", "image": [C, H, W], "fused_ms": t_ours, "framework_ms": t_ref, "speedup": t_ref / t_ours,
                      "loss_abs_diff": abs(lo - lr), "grad_rel_diff": ((go - gr).abs().max() / gr.abs().max()).item(),
                      # algorithmic bytes: read x,y twice, write+read 3 maps, write grad  = (2*2 + 3*2 + 1) * 4 B / element
                      "algorithmic_bytes": 11 * 4 * px}))


if __name__ == "__main__":
    main()
