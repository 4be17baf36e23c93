"""Host-side mirror of LoGS' tracking loss and query-image gradient mask (same names and argument meaning as
gs_localization/pipelines/tools/descent_utils.py and tools/camera_utils.py:164-192).

`get_loss_tracking` runs the fused CUDA kernel (gsr_tracking_loss_grad: forward and gradient in one pass over the
pixels) behind a torch.autograd.Function, so the reference's loop body works unchanged; the graph-captured loop in
`localization.GraphRefiner` calls the same kernel directly.  The gradient mask is computed once per query (not on
the per-iteration path) with a handful of framework ops."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib


def image_gradient(image: torch.Tensor):
    """Scharr gradients with reflect padding, normalised by 1/32 (the sum of the absolute kernel weights) (descent_utils.py:33-49); image [C,H,W]."""
    c = image.shape[0]
    k_h = torch.tensor([[3, 0, -3], [10, 0, -10], [3, 0, -3]], dtype=torch.float32, device=image.device)
    k_v = k_h.t().contiguous()
    p_img = F.pad(image, (1, 1, 1, 1), mode="reflect")[None]
    norm = 1.0 / 32.0
    grad_v = norm * F.conv2d(p_img, k_v.view(1, 1, 3, 3).repeat(c, 1, 1, 1), groups=c)
    grad_h = norm * F.conv2d(p_img, k_h.view(1, 1, 3, 3).repeat(c, 1, 1, 1), groups=c)
    return grad_v[0], grad_h[0]


def image_gradient_mask(image: torch.Tensor, eps: float = 0.01):
    """True where the whole 3x3 neighbourhood has |value| > eps (descent_utils.py:52-66)."""
    c = image.shape[0]
    ones = torch.ones((c, 1, 3, 3), dtype=torch.float32, device=image.device)
    p_img = (F.pad(image, (1, 1, 1, 1), mode="reflect")[None].abs() > eps).float()
    full = F.conv2d(p_img, ones, groups=c)[0] == 9.0
    return full, full.clone()


def compute_grad_mask(original_image: torch.Tensor, edge_threshold: float, dataset_type: str = "tum") -> torch.Tensor:
    """Camera.compute_grad_mask (tools/camera_utils.py:164-192): pixels whose Scharr gradient magnitude of the gray
    image exceeds edge_threshold x the median; for "replica" the median is taken per cell of a 32x32 grid."""
    gray = original_image.mean(dim=0, keepdim=True)
    gv, gh = image_gradient(gray)
    mv, mh = image_gradient_mask(gray)
    intensity = torch.sqrt((gv * mv) ** 2 + (gh * mh) ** 2)
    if dataset_type == "replica":
        _, h, w = original_image.shape
        bh, bw = int(h / 32), int(w / 32)
        out = intensity.clone()
        for r in range(32):
            for c in range(32):
                block = out[:, r * bh:(r + 1) * bh, c * bw:(c + 1) * bw]
                th = block.median() * edge_threshold
                hit = block > th
                block[hit] = 1
                block[~hit] = 0
        return out
    return intensity > intensity.median() * edge_threshold


class _TrackingLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, depth, exposure_a, exposure_b, opacity, gt_image, gt_depth, grad_mask, opacity_threshold, depth_weight):
        if not image.is_cuda:
            raise RuntimeError("gs_localization_b200: tensors must be CUDA tensors (no CPU fallback exists)")
        lib = _lib.load()
        f = lambda t: None if t is None else t.detach().to(torch.float32).contiguous()
        I, D, O, G, GD, GM = f(image), f(depth), f(opacity), f(gt_image), f(gt_depth), f(grad_mask)
        E = torch.cat([exposure_a.detach().reshape(1), exposure_b.detach().reshape(1)]).float()
        H, W = I.shape[-2:]
        loss, dE = torch.zeros(1, device=I.device), torch.zeros(2, device=I.device)
        dI, dD = torch.empty_like(I), torch.empty_like(D)
        p = lambda t: 0 if t is None else t.data_ptr()
        _lib.check(lib.gsr_tracking_loss_grad(p(I), p(D), p(O), p(G), p(GD), p(GM), p(E), int(H), int(W), float(opacity_threshold),
                                              float(depth_weight), p(loss), p(dI), p(dD), p(dE),
                                              torch.cuda.current_stream(I.device).cuda_stream), "gsr_tracking_loss_grad")
        ctx.save_for_backward(dI, dD, dE)
        ctx.shapes = (exposure_a.shape, exposure_b.shape)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        dI, dD, dE = ctx.saved_tensors
        return (dI * g, dD * g, (dE[0] * g).reshape(ctx.shapes[0]), (dE[1] * g).reshape(ctx.shapes[1]), None, None, None, None, None, None)


def get_loss_tracking(config, image, depth, opacity, viewpoint, initialization=False):
    """descent_utils.py:85-123 — `viewpoint` provides original_image, depth (numpy or tensor), grad_mask,
    exposure_a, exposure_b; `config["Training"]` provides monocular, opacity_threshold and (optionally) alpha."""
    tr = config["Training"]
    gt_depth = None
    if not tr["monocular"]:
        gt_depth = viewpoint.depth
        if not torch.is_tensor(gt_depth):
            gt_depth = torch.from_numpy(gt_depth)
        gt_depth = gt_depth.to(dtype=torch.float32, device=image.device)
    alpha = tr["alpha"] if "alpha" in tr else 0.98
    return _TrackingLoss.apply(image, depth, viewpoint.exposure_a, viewpoint.exposure_b, opacity,
                               viewpoint.original_image.to(image.device), gt_depth, viewpoint.grad_mask,
                               tr["opacity_threshold"], 1.0 - alpha)
