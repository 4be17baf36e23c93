"""Generates tests/golden/ref_tracking.npz by importing the reference's own tracking loss
(/root/reference/gs_localization/pipelines/tools/descent_utils.py:85-123) on CPU in the build container.
The reference moves the query image with `.cuda()`; the stub below returns the CPU tensor instead."""
import importlib.util
import os

import numpy as np
import torch

spec = importlib.util.spec_from_file_location("ref_descent", "/root/reference/gs_localization/pipelines/tools/descent_utils.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


class _Img:
    def __init__(self, t):
        self.t = t

    def cuda(self):
        return self.t


class _View:
    pass


out = {}
for name, (H, W), seed, mono, alpha in [("rgbd", (30, 44), 0, False, 0.9), ("mono", (17, 23), 1, True, None), ("rgbd_default", (8, 8), 2, False, None)]:
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(3, H, W, generator=g)
    image = (gt + 0.1 * torch.randn(3, H, W, generator=g)).clamp(0, 1).requires_grad_(True)
    gt_depth = torch.rand(H, W, generator=g) * 3
    gt_depth[torch.rand(H, W, generator=g) < 0.2] = 0.0           # invalid depth pixels
    depth = (gt_depth[None] + 0.2 * torch.randn(1, H, W, generator=g)).requires_grad_(True)
    opacity = torch.rand(1, H, W, generator=g)
    v = _View()
    v.original_image = _Img(gt)
    v.depth = gt_depth.numpy()
    v.grad_mask = torch.rand(1, H, W, generator=g) > 0.4
    v.exposure_a = torch.tensor([0.03], requires_grad=True)
    v.exposure_b = torch.tensor([-0.02], requires_grad=True)
    cfg = {"Training": {"monocular": mono, "opacity_threshold": 0.5}}
    if alpha is not None:
        cfg["Training"]["alpha"] = alpha
    loss = ref.get_loss_tracking(cfg, image, depth, opacity, v)
    loss.backward()
    out.update({f"{name}_image": image.detach().numpy(), f"{name}_gt": gt.numpy(), f"{name}_depth": depth.detach().numpy(),
                f"{name}_gt_depth": gt_depth.numpy(), f"{name}_opacity": opacity.numpy(),
                f"{name}_grad_mask": v.grad_mask.numpy(), f"{name}_exposure": np.float32([0.03, -0.02]),
                f"{name}_mono": np.bool_(mono), f"{name}_depth_weight": np.float32(1 - (alpha if alpha is not None else 0.98)),
                f"{name}_loss": np.float32(loss.item()), f"{name}_dimage": image.grad.numpy(),
                f"{name}_ddepth": (depth.grad if depth.grad is not None else torch.zeros_like(depth)).numpy(),
                f"{name}_dexposure": np.float32([v.exposure_a.grad.item(), v.exposure_b.grad.item()])})
np.savez_compressed(os.path.join(os.path.dirname(__file__), "ref_tracking.npz"), **out)
print({k: float(v) for k, v in out.items() if k.endswith("_loss")})
