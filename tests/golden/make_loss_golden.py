"""Generates tests/golden/ref_loss.npz by importing the reference's own loss functions
(/root/reference/gaussian_splatting/utils/loss_utils.py:17-64) on CPU in the build container.
Run here only; the GPU box reads the committed fixture."""
import importlib.util
import os

import numpy as np
import torch

REF = "/root/reference/gaussian_splatting/utils/loss_utils.py"
spec = importlib.util.spec_from_file_location("ref_loss_utils", REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

out = {}
for name, (C, H, W), seed in [("a", (3, 37, 53), 0), ("b", (3, 64, 80), 1), ("c", (1, 20, 9), 2)]:
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(C, H, W, generator=g)
    # smooth-ish image correlated with gt so SSIM is in its usual range
    img = (gt + 0.15 * torch.randn(C, H, W, generator=g)).clamp(0, 1).requires_grad_(True)
    lam = 0.2
    l1 = ref.l1_loss(img, gt)
    ss = ref.ssim(img, gt)
    loss = (1.0 - lam) * l1 + lam * (1.0 - ss)
    loss.backward()
    out[f"{name}_img"] = img.detach().numpy()
    out[f"{name}_gt"] = gt.numpy()
    out[f"{name}_l1"] = np.float32(l1.item())
    out[f"{name}_ssim"] = np.float32(ss.item())
    out[f"{name}_loss"] = np.float32(loss.item())
    out[f"{name}_grad"] = img.grad.numpy()
np.savez_compressed(os.path.join(os.path.dirname(__file__), "ref_loss.npz"), **out)
print({k: v.shape if hasattr(v, "shape") and v.shape else float(v) for k, v in out.items()})
