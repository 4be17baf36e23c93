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


class _DepthLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, pseudo_depth, gt_depth, inv_numerator, pearson_weight, l1_weight):
        if not depth.is_cuda:
            raise RuntimeError("gs_localization_b200: tensors must be CUDA tensors (no CPU fallback exists)")
        lib = _lib.load()
        f = lambda t: None if t is None else t.detach().to(torch.float32).contiguous()
        d, m, g = f(depth), f(pseudo_depth), f(gt_depth)
        loss = torch.zeros(1, dtype=torch.float32, device=d.device)
        grad = torch.empty_like(d)
        scratch = torch.empty(9, dtype=torch.float64, device=d.device)
        p = lambda t: 0 if t is None else t.data_ptr()
        _lib.check(lib.gsr_depth_loss_grad(p(d), p(m), p(g), d.numel(), float(inv_numerator), float(pearson_weight), float(l1_weight),
                                           p(loss), p(grad), p(scratch), torch.cuda.current_stream(d.device).cuda_stream),
                   "gsr_depth_loss_grad")
        ctx.save_for_backward(grad)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None, None


def depth_loss(depth: torch.Tensor, pseudo_depth: torch.Tensor | None, gt_depth: torch.Tensor | None = None,
               inv_numerator: float = 1000.0, pearson_weight: float = 0.01, l1_weight: float = 0.05) -> torch.Tensor:
    """pearson_weight * min(1 - pearson(-m, d), 1 - pearson(inv_numerator / (m + 200), d)) + l1_weight * l1(d*mask, gt*mask),
    mask = gt > 0 (7scenes_gs_full_dslam.py:168-184); any shape, reduced over all elements."""
    return _DepthLoss.apply(depth, pseudo_depth, gt_depth, inv_numerator, pearson_weight, l1_weight)
