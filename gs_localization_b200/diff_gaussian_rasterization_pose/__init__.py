"""`diff_gaussian_rasterization_pose` surface used by LoGS pose refinement.

The reference imports this package (gs_localization/pipelines/tools/__init__.py:15-18) but
does not vendor or pin it; only its call sites are known: settings gain `projmatrix_raw`
(:67), forward gains `theta=` / `rho=` (:126-127, :139-140) and returns
(image, radii, depth, opacity, n_touched) (:130).  `theta` / `rho` are the camera's
`cam_rot_delta` / `cam_trans_delta` Adam parameters; they are zero when the rasterizer is
called (tools/pose_utils.py:120-121) and only receive gradients:

    dL/d[rho, theta]  of  L(exp(tau) @ T_w2c)  at tau = 0   (tools/pose_utils.py:90-122)

computed by the fused SE(3) chain rule in the preprocess-backward kernel (csrc/backward_pre.cu)
as the full analytic derivative of the in-tree forward (projection, EWA covariance, rendered
depth and SH view direction).  `n_touched[i]` counts the pixels where Gaussian i was blended
while transmittance was still above 0.5 (unpinned upstream).
"""
from typing import NamedTuple

import torch
import torch.nn as nn

from ..diff_gaussian_rasterization import _C


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    projmatrix_raw: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, theta, rho,
                        raster_settings):
    return _RasterizeGaussiansPose.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                         cov3Ds_precomp, theta, rho, raster_settings)


class _RasterizeGaussiansPose(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, theta, rho,
                raster_settings):
        rs = raster_settings
        (num_rendered, color, depth, alpha, radii, geomBuffer, binningBuffer, imgBuffer, n_touched) = _C._forward_impl(
            rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier, cov3Ds_precomp, rs.viewmatrix,
            rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, sh, rs.sh_degree, rs.campos,
            rs.prefiltered, rs.debug, want_n_touched=True)
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
                              binningBuffer, imgBuffer, alpha)
        ctx.mark_non_differentiable(radii, n_touched)
        return color, radii, depth, alpha, n_touched

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_alpha, grad_n_touched):
        rs = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer, imgBuffer,
         alpha) = ctx.saved_tensors
        H, W, dev = rs.image_height, rs.image_width, means3D.device
        z = lambda c: torch.zeros(c, H, W, dtype=torch.float32, device=dev)
        grad_color = z(3) if grad_color is None else grad_color
        grad_depth = z(1) if grad_depth is None else grad_depth
        grad_alpha = z(1) if grad_alpha is None else grad_alpha
        # inputs: means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, theta, rho
        n = ctx.needs_input_grad
        needs = dict(means3D=n[0], means2D=n[1], sh=n[2], colors=n[3], opacity=n[4], scales=n[5], rotations=n[6],
                     cov3D=n[7])
        want_pose = bool(n[8] or n[9])
        grads, dL_dtau = _C._backward_impl(
            rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier, cov3Ds_precomp, rs.viewmatrix,
            rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_color, grad_depth, grad_alpha, sh, rs.sh_degree, rs.campos,
            geomBuffer, ctx.num_rendered, binningBuffer, imgBuffer, alpha, rs.debug,
            projmatrix_raw=rs.projmatrix_raw, want_pose=want_pose, needs=needs)
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh, grad_scales,
         grad_rotations) = grads
        grad_rho = dL_dtau[:3] if want_pose and n[9] else None
        grad_theta = dL_dtau[3:] if want_pose and n[8] else None
        return (grad_means3D, grad_means2D, grad_sh, grad_colors_precomp, grad_opacities, grad_scales, grad_rotations,
                grad_cov3Ds_precomp, grad_theta, grad_rho, None)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            return _C.mark_visible(positions, rs.viewmatrix, rs.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, theta=None, rho=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        e = lambda: torch.Tensor([])
        shs = e() if shs is None else shs
        colors_precomp = e() if colors_precomp is None else colors_precomp
        scales = e() if scales is None else scales
        rotations = e() if rotations is None else rotations
        cov3D_precomp = e() if cov3D_precomp is None else cov3D_precomp
        dev = means3D.device
        if theta is None:
            theta = torch.zeros(3, dtype=torch.float32, device=dev)
        if rho is None:
            rho = torch.zeros(3, dtype=torch.float32, device=dev)
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   theta, rho, rs)
