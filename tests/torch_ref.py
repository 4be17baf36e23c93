"""Dense, differentiable float64 PyTorch restatement of the rasterizer math (test infrastructure).

It evaluates every (pixel, Gaussian) pair densely, so it is only usable for a few hundred
Gaussians at thumbnail resolution — but torch.autograd then yields gradients that are
independent of both the CUDA kernels and the C++ oracle.  The discrete decisions of the
reference (tile rectangle membership, `power > 0`, `alpha < 1/255`, `T < 1e-4` termination,
the ±1.3·tanfov clamp gate) enter as detached masks, exactly as the reference's analytic
backward treats them.

Reference math: forward.cu:20-379, backward.cu:144-176 (clamp gate) of
gaussian_splatting/submodules/diff-gaussian-rasterization/cuda_rasterizer/.
Pose convention: T_w2c <- exp(tau) T_w2c, tau = [rho, theta]
(gs_localization/pipelines/tools/pose_utils.py:90-122).
"""
from __future__ import annotations

import math

import torch

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435]


def skew_t(v):
    z = torch.zeros((), dtype=v.dtype)
    return torch.stack([torch.stack([z, -v[2], v[1]]), torch.stack([v[2], z, -v[0]]), torch.stack([-v[1], v[0], z])])


def se3_exp_t(tau):
    """Differentiable SE3 exp, small-angle branch (tools/pose_utils.py:54-102), valid at tau≈0."""
    rho, theta = tau[:3], tau[3:]
    Wm = skew_t(theta)
    W2 = Wm @ Wm
    I = torch.eye(3, dtype=tau.dtype)
    Rm = I + Wm + 0.5 * W2
    V = I + 0.5 * Wm + (1.0 / 6.0) * W2
    T = torch.eye(4, dtype=tau.dtype)
    T = T.clone()
    T[:3, :3] = Rm
    T[:3, 3] = V @ rho
    return T


def eval_sh(deg, sh, dirs):
    """sh [P,M,3], dirs [P,3] normalised -> [P,3] (forward.cu:30-62)."""
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    res = SH_C0 * sh[:, 0]
    if deg > 0:
        res = res - SH_C1 * y * sh[:, 1] + SH_C1 * z * sh[:, 2] - SH_C1 * x * sh[:, 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            res = (res + SH_C2[0] * xy * sh[:, 4] + SH_C2[1] * yz * sh[:, 5] + SH_C2[2] * (2 * zz - xx - yy) * sh[:, 6]
                   + SH_C2[3] * xz * sh[:, 7] + SH_C2[4] * (xx - yy) * sh[:, 8])
            if deg > 2:
                res = (res + SH_C3[0] * y * (3 * xx - yy) * sh[:, 9] + SH_C3[1] * xy * z * sh[:, 10]
                       + SH_C3[2] * y * (4 * zz - xx - yy) * sh[:, 11]
                       + SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12]
                       + SH_C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + SH_C3[5] * z * (xx - yy) * sh[:, 14]
                       + SH_C3[6] * x * (xx - 3 * yy) * sh[:, 15])
    return res


def render(means3D, shs, opacities, scales, rotations, w2c, raw_proj_t, W, H, tanfovx, tanfovy, bg, deg,
           tau=None, scale_modifier=1.0, depth_feeds_geometry=False, colors_precomp=None, cov3D_precomp=None):
    """Returns (color[3,H,W], depth[1,H,W], alpha[1,H,W], aux dict).  All float64.

    w2c: [4,4] world-to-camera (untransposed); raw_proj_t: projmatrix_raw in the transposed
    storage LoGS uses (tools/camera_utils.py:135-137).  tau (6,) optional pose delta.
    depth_feeds_geometry=False reproduces the reference's parameter gradients (the rendered
    depth's dependence on the splat depth itself is dropped, backward.cu:540-543);
    True gives the complete derivative, used for the pose gradient.
    """
    dt = torch.float64
    P = means3D.shape[0]
    T_w2c = w2c.to(dt)
    if tau is not None:
        T_w2c = se3_exp_t(tau) @ T_w2c
    view = T_w2c.t()                      # transposed storage
    full = view @ raw_proj_t.to(dt)
    campos = torch.linalg.inv(view)[3, :3]
    ones = torch.ones(P, 1, dtype=dt)
    ph = torch.cat([means3D, ones], 1)
    p_view = ph @ view                    # [P,4]
    p_hom = ph @ full
    p_w = 1.0 / (p_hom[:, 3] + 1e-7)
    ndc = p_hom[:, :2] * p_w[:, None]
    pix = torch.stack([((ndc[:, 0] + 1.0) * W - 1.0) * 0.5, ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5], 1)
    tz = p_view[:, 2]
    # cov3D (unnormalised quaternion, forward.cu:127)
    if cov3D_precomp is None:
        r, x, y, z = rotations[:, 0], rotations[:, 1], rotations[:, 2], rotations[:, 3]
        Rstd = torch.stack([
            torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)], 1),
            torch.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)], 1),
            torch.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1)], 1)  # [P,3,3]
        S = scale_modifier * scales
        RS = Rstd * S[:, None, :]
        Sigma = RS @ RS.transpose(1, 2)
    else:
        c = cov3D_precomp
        Sigma = torch.stack([torch.stack([c[:, 0], c[:, 1], c[:, 2]], 1), torch.stack([c[:, 1], c[:, 3], c[:, 4]], 1),
                             torch.stack([c[:, 2], c[:, 4], c[:, 5]], 1)], 1)
    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    txtz, tytz = p_view[:, 0] / tz, p_view[:, 1] / tz
    inx = (txtz >= -limx) & (txtz <= limx)
    iny = (tytz >= -limy) & (tytz <= limy)
    tx = torch.where(inx, p_view[:, 0], (txtz.clamp(-limx, limx) * tz).detach())
    ty = torch.where(iny, p_view[:, 1], (tytz.clamp(-limy, limy) * tz).detach())
    zero = torch.zeros_like(tz)
    J = torch.stack([torch.stack([fx / tz, zero, -(fx * tx) / (tz * tz)], 1),
                     torch.stack([zero, fy / tz, -(fy * ty) / (tz * tz)], 1)], 1)  # [P,2,3]
    Rc = T_w2c[:3, :3]
    JW = J @ Rc                           # [P,2,3]
    cov2 = JW @ Sigma @ JW.transpose(1, 2)
    a = cov2[:, 0, 0] + 0.3
    b = cov2[:, 0, 1]
    c2 = cov2[:, 1, 1] + 0.3
    det = a * c2 - b * b
    conic = torch.stack([c2 / det, -b / det, a / det], 1)
    mid = 0.5 * (a + c2)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam)).detach()
    visible = (tz > 0.2) & (det != 0)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    pixd = pix.detach()

    def lo(p, g):
        return torch.clamp(torch.trunc((p - radius) / 16.0), 0, g)

    def hi(p, g):
        return torch.clamp(torch.trunc((p + radius + 15.0) / 16.0), 0, g)

    rminx, rmaxx, rminy, rmaxy = lo(pixd[:, 0], gx), hi(pixd[:, 0], gx), lo(pixd[:, 1], gy), hi(pixd[:, 1], gy)
    visible = visible & ((rmaxx - rminx) * (rmaxy - rminy) > 0)
    # colour
    if colors_precomp is None:
        d = means3D - campos[None]
        d = d / d.norm(dim=1, keepdim=True)
        rgb_raw = eval_sh(deg, shs, d) + 0.5
        rgb = torch.clamp(rgb_raw, min=0.0)   # zero gradient where clamped
    else:
        rgb = colors_precomp
    depth_val = tz if depth_feeds_geometry else tz.detach()
    # depth order: float32 depth bits, ties by index (stable), as the radix sort does
    order = torch.argsort(tz.detach().float(), stable=True)
    order = order[visible[order]]
    ys, xs = torch.meshgrid(torch.arange(H, dtype=dt), torch.arange(W, dtype=dt), indexing="ij")
    pxf, pyf = xs.reshape(-1), ys.reshape(-1)           # [N]
    tile_x, tile_y = torch.floor(pxf / 16), torch.floor(pyf / 16)
    o = order
    dx = pix[o, 0][None, :] - pxf[:, None]              # [N,K]
    dy = pix[o, 1][None, :] - pyf[:, None]
    power = -0.5 * (conic[o, 0][None] * dx * dx + conic[o, 2][None] * dy * dy) - conic[o, 1][None] * dx * dy
    alpha = torch.clamp(opacities[o, 0][None] * torch.exp(power), max=0.99)
    in_tile = ((tile_x[:, None] >= rminx[o][None]) & (tile_x[:, None] < rmaxx[o][None])
               & (tile_y[:, None] >= rminy[o][None]) & (tile_y[:, None] < rmaxy[o][None]))
    mask = in_tile & (power.detach() <= 0) & (alpha.detach() >= 1.0 / 255.0)
    am = alpha * mask
    one_m = 1.0 - am
    Tincl = torch.cumprod(one_m, 1)
    Texcl = torch.cat([torch.ones_like(Tincl[:, :1]), Tincl[:, :-1]], 1)
    # termination: the first contributing splat with T*(1-alpha) < 1e-4 and all after it are dropped
    stop = (mask & (Tincl.detach() < 1e-4)).to(dt)
    alive = (torch.cumsum(stop, 1) == 0)
    am = am * alive
    one_m = 1.0 - am
    Tincl = torch.cumprod(one_m, 1)
    Texcl = torch.cat([torch.ones_like(Tincl[:, :1]), Tincl[:, :-1]], 1)
    w = am * Texcl                                     # [N,K]
    Tfinal = Tincl[:, -1] if Tincl.shape[1] else torch.ones_like(pxf)
    color = w @ rgb[o] + Tfinal[:, None] * bg[None].to(dt)
    depth = w @ depth_val[o]
    out_alpha = 1.0 - Tfinal
    contrib = (mask & alive)
    idxs = torch.arange(1, contrib.shape[1] + 1)[None].expand_as(contrib)
    n_contrib = torch.where(contrib, idxs, torch.zeros_like(idxs)).max(dim=1).values if contrib.shape[1] else torch.zeros(W * H, dtype=torch.long)
    aux = dict(radius=radius, visible=visible, pix=pix, conic=conic, order=order, n_contrib=n_contrib.reshape(H, W),
               rgb=rgb, depth=tz)
    return color.t().reshape(3, H, W), depth.reshape(1, H, W), out_alpha.reshape(1, H, W), aux
