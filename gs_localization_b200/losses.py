"""Fused map-training loss: (1 - lambda) * L1 + lambda * (1 - SSIM), forward and gradient in two CUDA passes.

Mirrors `l1_loss` and `ssim` of gaussian_splatting/utils/loss_utils.py:17-64 as LoGS combines them
(gs_localization/gs/7scenes_gs_full_dslam.py:165-166).  No framework fallback: CUDA tensors only."""
from __future__ import annotations

import torch

from . import _lib


class _L1SSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, gt, lambda_dssim):
        if not image.is_cuda:
            raise RuntimeError("gs_localization_b200: tensors must be CUDA tensors (no CPU fallback exists)")
        lib = _lib.load()
        x = image.detach().contiguous().float()
        y = gt.detach().contiguous().float()
        C, H, W = x.shape[-3:]
        loss = torch.zeros(1, dtype=torch.float32, device=x.device)
        grad = torch.empty_like(x)
        scratch = torch.empty(3 * x.numel() + 2, dtype=torch.float32, device=x.device)
        _lib.check(lib.gsr_l1_ssim_loss_grad(x.data_ptr(), y.data_ptr(), int(C), int(H), int(W), float(lambda_dssim),
                                             loss.data_ptr(), grad.data_ptr(), scratch.data_ptr(),
                                             torch.cuda.current_stream(x.device).cuda_stream), "gsr_l1_ssim_loss_grad")
        ctx.save_for_backward(grad)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


def l1_ssim_loss(image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float = 0.2) -> torch.Tensor:
    """(1 - lambda_dssim) * l1_loss(image, gt) + lambda_dssim * (1 - ssim(image, gt)); image, gt: [C,H,W]."""
    return _L1SSIM.apply(image, gt, lambda_dssim)
