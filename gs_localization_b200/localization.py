"""Pose refinement on the B200 rasterizer: the loop body of LoGS' `gradient_decent`
(gs_localization/pipelines/7scenes_localize_full_dslam.py:29-93) without its dataset I/O.

A `PoseCamera` carries the world-to-camera pose and the two zero-initialised Adam parameters
`cam_rot_delta` / `cam_trans_delta` (tools/camera_utils.py:38-110); each iteration renders through
`diff_gaussian_rasterization_pose`, takes an L1 (+ optional depth L1) tracking loss
(tools/descent_utils.py:85-123 with all-ones masks), steps Adam and folds the deltas into the pose
with the left-multiplicative SE(3) update of tools/pose_utils.py:105-122.
"""
from __future__ import annotations

import math

import torch

from . import synthetic as syn
from .diff_gaussian_rasterization_pose import GaussianRasterizationSettings, GaussianRasterizer


def so3_exp(theta: torch.Tensor) -> torch.Tensor:
    """tools/pose_utils.py:54-70 on device tensors."""
    z = torch.zeros((), dtype=theta.dtype, device=theta.device)
    Wm = torch.stack([torch.stack([z, -theta[2], theta[1]]), torch.stack([theta[2], z, -theta[0]]),
                      torch.stack([-theta[1], theta[0], z])])
    W2 = Wm @ Wm
    a = torch.linalg.norm(theta)
    I = torch.eye(3, dtype=theta.dtype, device=theta.device)
    if float(a) < 1e-5:
        return I + Wm + 0.5 * W2, Wm, W2, a
    return I + (torch.sin(a) / a) * Wm + ((1 - torch.cos(a)) / a**2) * W2, Wm, W2, a


def se3_exp(tau: torch.Tensor) -> torch.Tensor:
    """tools/pose_utils.py:73-102; tau = [rho, theta]."""
    rho, theta = tau[:3], tau[3:]
    Rm, Wm, W2, a = so3_exp(theta)
    I = torch.eye(3, dtype=tau.dtype, device=tau.device)
    if float(a) < 1e-5:
        V = I + 0.5 * Wm + (1.0 / 6.0) * W2
    else:
        V = I + Wm * ((1.0 - torch.cos(a)) / a**2) + W2 * ((a - torch.sin(a)) / a**3)
    T = torch.eye(4, dtype=tau.dtype, device=tau.device)
    T[:3, :3] = Rm
    T[:3, 3] = V @ rho
    return T


class PoseCamera:
    """Mirror of the MonoGS-derived Camera used by LoGS (tools/camera_utils.py): pose + pose deltas."""

    def __init__(self, cam: syn.Camera, device):
        self.device = torch.device(device)
        self.W, self.H = cam.W, cam.H
        self.tanfovx, self.tanfovy = cam.tanfovx, cam.tanfovy
        self.w2c = cam.w2c.to(torch.float32).to(self.device)        # R | T
        self.projection_matrix = syn.projection_matrix2(cam.znear, cam.zfar, cam.cx, cam.cy, cam.fx, cam.fy, cam.W,
                                                        cam.H).t().contiguous().float().to(self.device)
        self.cam_rot_delta = torch.nn.Parameter(torch.zeros(3, device=self.device))
        self.cam_trans_delta = torch.nn.Parameter(torch.zeros(3, device=self.device))
        # tools/camera_utils.py:66-91: the query image, its depth, gradient mask and exposure pair
        self.exposure_a = torch.nn.Parameter(torch.tensor([0.0], device=self.device))
        self.exposure_b = torch.nn.Parameter(torch.tensor([0.0], device=self.device))
        self.original_image, self.depth, self.grad_mask = None, None, None

    def compute_grad_mask(self, config):
        """tools/camera_utils.py:164-192."""
        from . import tracking
        self.grad_mask = tracking.compute_grad_mask(self.original_image, config["Training"]["edge_threshold"],
                                                    config["Dataset"]["type"])

    # tools/camera_utils.py:144-158
    @property
    def world_view_transform(self):
        return self.w2c.t().contiguous()

    @property
    def full_proj_transform(self):
        return self.world_view_transform @ self.projection_matrix

    @property
    def camera_center(self):
        Rm, t = self.w2c[:3, :3], self.w2c[:3, 3]
        return -(Rm.t() @ t)

    def update_pose(self, converged_threshold: float = 1e-4) -> bool:
        """tools/pose_utils.py:105-122."""
        with torch.no_grad():
            tau = torch.cat([self.cam_trans_delta, self.cam_rot_delta])
            self.w2c = se3_exp(tau) @ self.w2c
            converged = bool(tau.norm() < converged_threshold)
            self.cam_rot_delta.zero_()
            self.cam_trans_delta.zero_()
        return converged


def render_pose(gmap: syn.GaussianMap, cam: PoseCamera, bg: torch.Tensor):
    """tools/__init__.py:24-153 reduced to its rasterizer call."""
    rs = GaussianRasterizationSettings(
        image_height=cam.H, image_width=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg, scale_modifier=1.0,
        viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, projmatrix_raw=cam.projection_matrix,
        sh_degree=gmap.sh_degree, campos=cam.camera_center, prefiltered=False, debug=False)
    # means2D is only a gradient sink in the reference API (never read by the kernels): no need to allocate one here
    return GaussianRasterizer(rs)(means3D=gmap.means3D, means2D=gmap.means3D, opacities=gmap.opacities, shs=gmap.shs,
                                  scales=gmap.scales, rotations=gmap.rotations, theta=cam.cam_rot_delta,
                                  rho=cam.cam_trans_delta)


def refine_pose(gmap: syn.GaussianMap, cam: PoseCamera, target: torch.Tensor, iters: int = 50, lr: float = 1e-3,
                target_depth: torch.Tensor | None = None, depth_weight: float = 0.01, converge: bool = False):
    """`gradient_decent` (7scenes_localize_full_dslam.py:29-93): Adam on (rot, trans) deltas, fixed iteration count
    unless `converge` re-enables the reference's early break.  The map carries no gradient (pose-only refinement)."""
    bg = torch.zeros(3, device=cam.device)
    opt = torch.optim.Adam([{"params": [cam.cam_rot_delta], "lr": lr}, {"params": [cam.cam_trans_delta], "lr": lr}])
    loss = None
    for _ in range(iters):
        image, radii, depth, opacity, n_touched = render_pose(gmap, cam, bg)
        loss = (image - target).abs().mean()
        if target_depth is not None:
            loss = loss + depth_weight * (depth - target_depth).abs().mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        if cam.update_pose() and converge:
            break
    return cam.w2c, loss


def gradient_decent(gmap: syn.GaussianMap, viewpoint: PoseCamera, config, iters: int = 50, lr: float = 1e-3,
                    converged_threshold: float | None = 1e-4):
    """The reference loop as written (7scenes_localize_full_dslam.py:29-93): Adam over rotation, translation and the two
    exposure scalars, the full tracking loss (fused kernel behind autograd), update_pose, early break on convergence."""
    from . import tracking
    bg = torch.zeros(3, device=viewpoint.device)
    opt = torch.optim.Adam([{"params": [viewpoint.cam_rot_delta], "lr": lr}, {"params": [viewpoint.cam_trans_delta], "lr": lr},
                            {"params": [viewpoint.exposure_a], "lr": lr}, {"params": [viewpoint.exposure_b], "lr": lr}])
    loss = None
    for _ in range(iters):
        image, radii, depth, opacity, n_touched = render_pose(gmap, viewpoint, bg)
        opt.zero_grad()
        loss = tracking.get_loss_tracking(config, image, depth, opacity, viewpoint)
        loss.backward()
        with torch.no_grad():
            opt.step()
            converged = viewpoint.update_pose(converged_threshold if converged_threshold is not None else -1.0)
        if converged:
            break
    return viewpoint.w2c, loss


class TrackingLoss:
    """Options of LoGS' full tracking loss (tools/descent_utils.py:85-123) for the fused loops: opacity threshold,
    RGB-D weight alpha (depth term weighs 1 - alpha; `monocular` drops it), and Adam on the exposure pair
    (7scenes_localize_full_dslam.py:48-61).  The per-query gradient mask comes from tracking.compute_grad_mask."""

    def __init__(self, opacity_threshold: float = 0.5, alpha: float = 0.98, monocular: bool = False, exposure_lr: float = 1e-3,
                 optimise_exposure: bool = True):
        self.opacity_threshold, self.alpha, self.monocular = float(opacity_threshold), float(alpha), bool(monocular)
        self.exposure_lr, self.optimise_exposure = float(exposure_lr), bool(optimise_exposure)

    @staticmethod
    def from_config(config) -> "TrackingLoss":
        tr = config["Training"]
        return TrackingLoss(tr["opacity_threshold"], tr["alpha"] if "alpha" in tr else 0.98, tr["monocular"])


def refine_pose_fused(gmap: syn.GaussianMap, cam: PoseCamera, target: torch.Tensor, iters: int = 50, lr: float = 1e-3,
                      lr_rot: float | None = None, target_depth: torch.Tensor | None = None, depth_weight: float = 0.01,
                      converge_threshold: float | None = None, tracking: "TrackingLoss | None" = None,
                      grad_mask: torch.Tensor | None = None, exposure: torch.Tensor | None = None):
    """Same loop as `refine_pose`, without framework ops on the per-iteration path: the rasterizer's C ABI
    forward, a fused L1 loss+gradient kernel, the pose-only backward (SE(3) chain rule fused, no per-Gaussian
    gradients written) and one single-thread kernel doing Adam + SE3_exp + the new view constants.  The only
    host wait per iteration is the rasterizer's num_rendered poll; everything else is queued ahead."""
    import ctypes as C

    from . import _lib
    from .diff_gaussian_rasterization import _C

    lib = _lib.load()
    dev = cam.device
    H, W = cam.H, cam.W
    f32 = dict(dtype=torch.float32, device=dev)
    w2c = cam.w2c.contiguous().clone()
    view = cam.world_view_transform.contiguous().clone()
    proj = cam.full_proj_transform.contiguous().clone()
    campos = cam.camera_center.contiguous().clone()
    raw = cam.projection_matrix.contiguous()
    adam_m, adam_v, step, tau_norm = torch.zeros(6, **f32), torch.zeros(6, **f32), torch.zeros(1, **f32), torch.zeros(1, **f32)
    loss = torch.zeros(1, **f32)
    dL_dpix = torch.empty(3, H, W, **f32)
    dL_ddepth = torch.zeros(1, H, W, **f32)
    zeros_a = torch.zeros(1, H, W, **f32)
    bg = torch.zeros(3, **f32)
    e = torch.Tensor([])
    needs = dict(means3D=False, means2D=False, sh=False, colors=False, opacity=False, scales=False, rotations=False, cov3D=False)
    stream = torch.cuda.current_stream(dev).cuda_stream
    p = lambda t: t.data_ptr()
    lr_rot = lr if lr_rot is None else lr_rot
    target = target.contiguous()
    if tracking is not None:
        gm = None if grad_mask is None else grad_mask.to(**f32).contiguous()
        exposure = torch.zeros(2, **f32) if exposure is None else exposure
        dE, exp_m, exp_v, exp_step = torch.zeros(2, **f32), torch.zeros(2, **f32), torch.zeros(2, **f32), torch.zeros(1, **f32)
        gt_d = None if (tracking.monocular or target_depth is None) else target_depth.contiguous()
    for it in range(iters):
        fwd = _C._forward_impl(bg, gmap.means3D, e, gmap.opacities, gmap.scales, gmap.rotations, 1.0, e, view.view(4, 4),
                               proj.view(4, 4), cam.tanfovx, cam.tanfovy, H, W, gmap.shs, gmap.sh_degree, campos, False, False)
        R, color, depth, alpha, radii, geom, binning, img, _ = fwd
        loss.zero_()
        gD = zeros_a
        if tracking is not None:
            _lib.check(lib.gsr_tracking_loss_grad(p(color), p(depth), p(alpha), p(target), 0 if gt_d is None else p(gt_d),
                                                  0 if gm is None else p(gm), p(exposure), H, W, tracking.opacity_threshold,
                                                  1.0 - tracking.alpha, p(loss), p(dL_dpix), p(dL_ddepth), p(dE), stream),
                       "gsr_tracking_loss_grad")
            gD = dL_ddepth
        else:
            _lib.check(lib.gsr_l1_loss_grad(p(color), p(target), p(dL_dpix), 3 * H * W, 1.0, p(loss), stream), "gsr_l1_loss_grad")
        if tracking is None and target_depth is not None:
            _lib.check(lib.gsr_l1_loss_grad(p(depth), p(target_depth), p(dL_ddepth), H * W, depth_weight, p(loss), stream),
                       "gsr_l1_loss_grad")
            gD = dL_ddepth
        _, dL_dtau = _C._backward_impl(bg, gmap.means3D, radii, e, gmap.scales, gmap.rotations, 1.0, e, view.view(4, 4),
                                       proj.view(4, 4), cam.tanfovx, cam.tanfovy, dL_dpix, gD, zeros_a, gmap.shs,
                                       gmap.sh_degree, campos, geom, R, binning, img, alpha, False, projmatrix_raw=raw,
                                       want_pose=True, needs=needs)
        _lib.check(lib.gsr_pose_adam_step(p(dL_dtau), p(adam_m), p(adam_v), p(step), float(lr), float(lr_rot), p(w2c), p(raw),
                                          p(view), p(proj), p(campos), p(tau_norm), stream), "gsr_pose_adam_step")
        if tracking is not None and tracking.optimise_exposure:
            _lib.check(lib.gsr_exposure_adam_step(p(exposure), p(dE), p(exp_m), p(exp_v), p(exp_step), tracking.exposure_lr, stream),
                       "gsr_exposure_adam_step")
        elif tracking is not None:
            dE.zero_()
        if converge_threshold is not None and float(tau_norm) < converge_threshold:   # the reference's early break (host sync)
            break
    cam.w2c = w2c.view(4, 4)
    return cam.w2c, loss


class GraphRefiner:
    """One pose-refinement iteration captured as a CUDA graph and replayed `iters` times per query.

    All scratch (geometry / image / binning buffers, images, gradients, Adam and pose state) is allocated once
    for a (map, image size) pair; an iteration is the sync-free forward (gsr_rasterize_forward_async), the fused
    L1 loss+gradient kernel, the pose-only backward and the Adam+SE3 kernel — no host round trip, no allocation,
    so the launch cost of an iteration is one graph replay.  The binning capacity is taken from a first eager
    forward (x1.5); if a later pose overflows it the query is re-run through the eager fused loop."""

    def __init__(self, gmap: syn.GaussianMap, cam: PoseCamera, lr: float = 1e-3, lr_rot: float | None = None,
                 depth_weight: float | None = None, tracking: "TrackingLoss | None" = None):
        import ctypes as C

        from . import _lib
        from .diff_gaussian_rasterization import _C

        self.lib, self._lib, self.C = _lib.load(), _lib, C
        self.gmap, self.dev = gmap, cam.device
        self.H, self.W, self.tanfovx, self.tanfovy = cam.H, cam.W, cam.tanfovx, cam.tanfovy
        self.lr, self.lr_rot = float(lr), float(lr if lr_rot is None else lr_rot)
        self.depth_weight, self.tracking = depth_weight, tracking
        dev, H, W = self.dev, self.H, self.W
        P = int(gmap.means3D.shape[0])
        self.P, self.M = P, int(gmap.shs.shape[1])
        f32 = dict(dtype=torch.float32, device=dev)
        byte = dict(dtype=torch.uint8, device=dev)
        # capacity from one eager forward at the starting pose
        e = torch.Tensor([])
        bg = torch.zeros(3, **f32)
        fwd = _C._forward_impl(bg, gmap.means3D, e, gmap.opacities, gmap.scales, gmap.rotations, 1.0, e,
                               cam.world_view_transform, cam.full_proj_transform, cam.tanfovx, cam.tanfovy, H, W, gmap.shs,
                               gmap.sh_degree, cam.camera_center, False, False)
        cnt = (C.c_uint * 3)()
        _lib.check(self.lib.gsr_read_counters(fwd[5].data_ptr(), P, cnt, torch.cuda.current_stream(dev).cuda_stream), "gsr_read_counters")
        self.capacity = int(cnt[0] * 2.0) + 65536     # headroom for the other views of the run: growing means re-capturing
        self.global_sort = int(cnt[2] * 1.5 > 4096)
        lib = self.lib
        self.geom = torch.empty(lib.gsr_geometry_bytes(P), **byte)
        self.img = torch.empty(lib.gsr_image_bytes(W, H), **byte)
        self.binning = torch.empty(lib.gsr_binning_bytes(self.capacity, W, H), **byte)
        self.color, self.depth, self.alpha = torch.empty(3, H, W, **f32), torch.empty(1, H, W, **f32), torch.empty(1, H, W, **f32)
        self.radii = torch.empty(P, dtype=torch.int32, device=dev)
        self.dL_dpix, self.dL_ddepth, self.zeros1 = torch.empty(3, H, W, **f32), torch.zeros(1, H, W, **f32), torch.zeros(1, H, W, **f32)
        self.dL_dtau, self.loss = torch.zeros(6, **f32), torch.zeros(1, **f32)
        self.adam_m, self.adam_v, self.step, self.tau_norm = (torch.zeros(6, **f32), torch.zeros(6, **f32), torch.zeros(1, **f32),
                                                              torch.zeros(1, **f32))
        self.w2c, self.view, self.proj, self.campos = (torch.zeros(16, **f32), torch.zeros(16, **f32), torch.zeros(16, **f32),
                                                       torch.zeros(3, **f32))
        self.raw = cam.projection_matrix.contiguous().clone()
        self.bg = bg
        self.target = torch.zeros(3, H, W, **f32)
        self.target_depth = torch.zeros(1, H, W, **f32)
        self.grad_mask = torch.ones(1, H, W, **f32)
        self.exposure, self.dL_dexposure = torch.zeros(2, **f32), torch.zeros(2, **f32)
        self.exp_m, self.exp_v, self.exp_step = torch.zeros(2, **f32), torch.zeros(2, **f32), torch.zeros(1, **f32)
        # the map is read-only for the whole run: pack the cull pass's inputs once (16 bytes per Gaussian)
        self.cull_records = torch.empty(P, 4, **f32)
        _lib.check(lib.gsr_build_cull_records(P, gmap.means3D.data_ptr(), gmap.scales.data_ptr(), gmap.rotations.data_ptr(),
                                              self.cull_records.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "gsr_build_cull_records")
        self.graph = None

    def _iteration(self):
        lib, chk, g = self.lib, self._lib.check, self.gmap
        p = lambda t: t.data_ptr()
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        H, W = self.H, self.W
        chk(lib.gsr_rasterize_forward_async(
            p(self.geom), p(self.binning), self.capacity, self.global_sort, p(self.img), self.P, g.sh_degree, self.M,
            p(self.bg), W, H, p(g.means3D), p(g.shs), None, p(g.opacities), p(g.scales), 1.0, p(g.rotations), None,
            p(self.view), p(self.proj), p(self.campos), self.tanfovx, self.tanfovy,
            p(self.color), p(self.depth), p(self.alpha), p(self.radii), None, p(self.cull_records), stream), "gsr_rasterize_forward_async")
        self.loss.zero_()
        gD = self.zeros1
        tr = self.tracking
        if tr is not None:
            chk(lib.gsr_tracking_loss_grad(p(self.color), p(self.depth), p(self.alpha), p(self.target),
                                           0 if tr.monocular else p(self.target_depth), p(self.grad_mask), p(self.exposure), H, W,
                                           tr.opacity_threshold, 1.0 - tr.alpha, p(self.loss), p(self.dL_dpix), p(self.dL_ddepth),
                                           p(self.dL_dexposure), stream), "gsr_tracking_loss_grad")
            gD = self.dL_ddepth
        else:
            chk(lib.gsr_l1_loss_grad(p(self.color), p(self.target), p(self.dL_dpix), 3 * H * W, 1.0, p(self.loss), stream), "gsr_l1_loss_grad")
        if tr is None and self.depth_weight is not None:
            chk(lib.gsr_l1_loss_grad(p(self.depth), p(self.target_depth), p(self.dL_ddepth), H * W, float(self.depth_weight),
                                     p(self.loss), stream), "gsr_l1_loss_grad")
            gD = self.dL_ddepth
        chk(lib.gsr_rasterize_backward(
            self.P, g.sh_degree, self.M, self.capacity, p(self.bg), W, H, p(g.means3D), p(g.shs), None, p(self.alpha),
            p(g.scales), 1.0, p(g.rotations), None, p(self.view), p(self.proj), p(self.raw), p(self.campos),
            self.tanfovx, self.tanfovy, p(self.radii), p(self.geom), p(self.binning), p(self.img),
            p(self.dL_dpix), p(gD), p(self.zeros1), None, None, None, None, None, None, None, None, None,
            p(self.dL_dtau), 0, stream), "gsr_rasterize_backward")
        chk(lib.gsr_pose_adam_step(p(self.dL_dtau), p(self.adam_m), p(self.adam_v), p(self.step), self.lr, self.lr_rot, p(self.w2c),
                                   p(self.raw), p(self.view), p(self.proj), p(self.campos), p(self.tau_norm), stream),
            "gsr_pose_adam_step")
        if tr is not None:
            if tr.optimise_exposure:
                chk(lib.gsr_exposure_adam_step(p(self.exposure), p(self.dL_dexposure), p(self.exp_m), p(self.exp_v), p(self.exp_step),
                                               tr.exposure_lr, stream), "gsr_exposure_adam_step")
            else:
                self.dL_dexposure.zero_()

    def _load_query(self, cam: PoseCamera, target, target_depth, grad_mask=None):
        self.w2c.copy_(cam.w2c.reshape(-1))
        self.view.copy_(cam.world_view_transform.reshape(-1))
        self.proj.copy_(cam.full_proj_transform.reshape(-1))
        self.campos.copy_(cam.camera_center)
        self.target.copy_(target)
        if target_depth is not None:
            self.target_depth.copy_(target_depth)
        if grad_mask is not None:
            self.grad_mask.copy_(grad_mask.reshape(1, self.H, self.W))
        else:
            self.grad_mask.fill_(1.0)
        for t in (self.adam_m, self.adam_v, self.step, self.exposure, self.dL_dexposure, self.exp_m, self.exp_v, self.exp_step):
            t.zero_()
        # the overflow flag is sticky across the replays of a query (any iteration that exceeded the binning capacity
        # invalidates the trajectory); it is cleared here and read once in collect()
        self._lib.check(self.lib.gsr_clear_overflow(self.geom.data_ptr(), self.P, torch.cuda.current_stream(self.dev).cuda_stream),
                        "gsr_clear_overflow")

    def _counters(self):
        cnt = (self.C.c_uint * 3)()
        self._lib.check(self.lib.gsr_read_counters(self.geom.data_ptr(), self.P, cnt,
                                                   torch.cuda.current_stream(self.dev).cuda_stream), "gsr_read_counters")
        return int(cnt[0]), int(cnt[1]), int(cnt[2])

    def _ensure_capacity(self):
        """One un-captured sync-free forward at the query's starting pose; grow the binning buffer (and drop the
        captured graph, whose nodes hold the old pointers) if this view needs more than 80 % of the capacity."""
        p = lambda t: t.data_ptr()
        g = self.gmap
        self._lib.check(self.lib.gsr_rasterize_forward_async(
            p(self.geom), p(self.binning), self.capacity, self.global_sort, p(self.img), self.P, g.sh_degree, self.M,
            p(self.bg), self.W, self.H, p(g.means3D), p(g.shs), None, p(g.opacities), p(g.scales), 1.0, p(g.rotations), None,
            p(self.view), p(self.proj), p(self.campos), self.tanfovx, self.tanfovy,
            p(self.color), p(self.depth), p(self.alpha), p(self.radii), None, p(self.cull_records),
            torch.cuda.current_stream(self.dev).cuda_stream), "gsr_rasterize_forward_async")
        R, overflow, longest = self._counters()
        need_global = int(longest * 1.5 > 4096)
        if R * 1.25 > self.capacity or need_global > self.global_sort:
            # twice the view's instances: every growth drops the captured graph, so grow rarely
            self._resize(max(self.capacity, int(R * 2.0) + 65536), max(self.global_sort, need_global))
            return True
        return False

    def _resize(self, capacity: int, global_sort: int):
        """New binning buffer of `capacity` instances (and sort path); the captured graph holds the old pointers and is dropped."""
        if capacity == self.capacity and global_sort == self.global_sort:
            return
        self.capacity, self.global_sort = int(capacity), int(global_sort)
        self.binning = torch.empty(self.lib.gsr_binning_bytes(self.capacity, self.W, self.H), dtype=torch.uint8, device=self.dev)
        self.graph = None

    def refine(self, cam: PoseCamera, target: torch.Tensor, iters: int = 50, target_depth: torch.Tensor | None = None,
               grad_mask: torch.Tensor | None = None):
        """Refine one query; returns (w2c [4,4], loss [1]) like refine_pose_fused."""
        self.submit(cam, target, iters, target_depth, grad_mask)
        return self.collect()

    def collect(self):
        """Wait for the query submitted last and return (w2c, loss)."""
        cam, target, iters, target_depth, grad_mask = self._pending
        if self._counters()[1]:   # some iteration overflowed the binning capacity: redo this query eagerly
            self.exposure.zero_()
            return refine_pose_fused(self.gmap, cam, target, iters=iters, lr=self.lr, lr_rot=self.lr_rot,
                                     target_depth=target_depth if (self.depth_weight is not None or self.tracking is not None) else None,
                                     depth_weight=self.depth_weight or 0.01, tracking=self.tracking, grad_mask=grad_mask,
                                     exposure=self.exposure)
        cam.w2c = self.w2c.view(4, 4).clone()
        return cam.w2c, self.loss.clone()

    def submit(self, cam: PoseCamera, target: torch.Tensor, iters: int = 50, target_depth: torch.Tensor | None = None,
               grad_mask: torch.Tensor | None = None):
        """Queue one query on the current stream without waiting for it (several refiners on different streams keep
        the GPU busy: the latency-bound binning kernels of one query overlap the blend kernels of another)."""
        self._pending = (cam, target, iters, target_depth, grad_mask)
        self._load_query(cam, target, target_depth, grad_mask)
        self._ensure_capacity()
        if self.graph is None:
            self._iteration()                    # warm-up outside capture (lazy CUDA state, side streams)
            torch.cuda.synchronize(self.dev)
            self._load_query(cam, target, target_depth, grad_mask)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._iteration()
            self._load_query(cam, target, target_depth, grad_mask)   # capture does not execute, but keep the state explicit
        for _ in range(iters):
            self.graph.replay()


class BatchedGraphRefiner:
    """B independent queries against one map per CUDA-graph launch (queries are independent, the map
    is read-only).  One graph holds the refinement iteration of B `GraphRefiner`s as B parallel branches; replaying it
    `iters` times refines B queries at once.  The latency-bound kernels of a small frame (C1: fourteen launches of a
    few microseconds each per iteration) run side by side instead of one after the other, and the host pays one graph
    launch per iteration for the whole batch; a large frame's blend kernels, which fill the GPU on their own, overlap
    with the other branches' binning kernels.  Every branch runs exactly the kernels of the unbatched loop on its own
    scratch, so the refined poses are those of `GraphRefiner`."""

    def __init__(self, gmap: syn.GaussianMap, cam: PoseCamera, batch: int = 4, **kw):
        self.dev = cam.device
        self.refiners = [GraphRefiner(gmap, cam, **kw) for _ in range(int(batch))]
        self.streams = [torch.cuda.Stream(self.dev) for _ in range(int(batch))]
        self.graph = None
        self._captured_with = None

    @property
    def batch(self) -> int:
        return len(self.refiners)

    def _capture(self):
        main = torch.cuda.current_stream(self.dev)
        for r in self.refiners:                 # warm-up outside capture (lazy CUDA state, side streams)
            r._iteration()
        torch.cuda.synchronize(self.dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            main_c = torch.cuda.current_stream(self.dev)
            for r, s_ in zip(self.refiners, self.streams):
                s_.wait_stream(main_c)
                with torch.cuda.stream(s_):
                    r._iteration()
            for s_ in self.streams:
                main_c.wait_stream(s_)
        self._captured_with = [(r.capacity, r.global_sort) for r in self.refiners]
        del main

    def refine_batch(self, cams, targets, iters: int = 50, target_depths=None, grad_masks=None):
        """Refine len(cams) <= batch queries; returns [(w2c, loss), ...].  A partial batch repeats its last query in the idle
        branches (their results are discarded)."""
        self.submit_batch(cams, targets, iters, target_depths, grad_masks)
        return self.collect_batch()

    def submit_batch(self, cams, targets, iters: int = 50, target_depths=None, grad_masks=None):
        """Queue the refinement of len(cams) <= batch queries on the current stream without waiting for it
        (`collect_batch` does); the only host wait in here is the capacity check of each query's starting view."""
        n = len(cams)
        assert 1 <= n <= self.batch
        td = list(target_depths) if target_depths is not None else [None] * n
        gm = list(grad_masks) if grad_masks is not None else [None] * n
        for i, r in enumerate(self.refiners):
            j = min(i, n - 1)
            r._pending = (cams[j], targets[j], iters, td[j], gm[j])
            r._load_query(cams[j], targets[j], td[j], gm[j])
            r._ensure_capacity()                # may grow this branch's binning buffer
        # one branch grew: bring every branch to the same size now, so that a similar view landing in another branch later
        # does not cost another capture of the whole batch graph
        cap, gs = max(r.capacity for r in self.refiners), max(r.global_sort for r in self.refiners)
        for r in self.refiners:
            r._resize(cap, gs)
        if self.graph is None or self._captured_with != [(r.capacity, r.global_sort) for r in self.refiners]:
            self._capture()
            for i, r in enumerate(self.refiners):
                j = min(i, n - 1)
                r._load_query(cams[j], targets[j], td[j], gm[j])
        for _ in range(iters):
            self.graph.replay()
        self._submitted = n

    def collect_batch(self):
        """Wait for the batch submitted last; [(w2c, loss), ...] in the order of its queries."""
        out = []
        for i in range(self._submitted):
            r = self.refiners[i]
            r.graph = None                      # collect()'s eager fallback must not reuse a stale single-query graph
            out.append(r.collect())
        self._submitted = 0
        return out


class PipelinedBatchRefiner:
    """A stream of queries against one map through `depth` BatchedGraphRefiners that take turns, each on its own CUDA
    stream: while one set's graph replays run, the host loads the next batch into the other set (pose / target copies,
    the capacity check of each starting view, which is a forward plus a host read) and queues its replays, and collects
    the first set only when it needs it again.  The GPU never waits for the host between batches; every query still runs
    exactly the kernels of `GraphRefiner`, so the refined poses are unchanged (queries are independent: loop of
    gs_localization/pipelines/7scenes_localize_full_dslam.py:352-365)."""

    def __init__(self, gmap: syn.GaussianMap, cam: PoseCamera, batch: int = 4, depth: int = 2, **kw):
        self.dev = cam.device
        self.sets = [BatchedGraphRefiner(gmap, cam, batch=batch, **kw) for _ in range(int(depth))]
        self.streams = [torch.cuda.Stream(self.dev) for _ in range(int(depth))]

    @property
    def batch(self) -> int:
        return self.sets[0].batch

    def refine_all(self, cams, targets, iters: int = 50, target_depths=None, grad_masks=None):
        """Refine all queries; returns [(w2c, loss), ...] in query order."""
        n, B, D = len(cams), self.batch, len(self.sets)
        td = list(target_depths) if target_depths is not None else [None] * n
        gm = list(grad_masks) if grad_masks is not None else [None] * n
        main = torch.cuda.current_stream(self.dev)
        for s_ in self.streams:
            s_.wait_stream(main)                  # the targets were produced on the caller's stream
        results = [None] * n
        pending = [None] * D                      # per set: (first, last) of the batch in flight

        def collect(k):
            if pending[k] is None:
                return
            a, b = pending[k]
            with torch.cuda.stream(self.streams[k]):
                results[a:b] = self.sets[k].collect_batch()
            pending[k] = None

        for j, a in enumerate(range(0, n, B)):
            k, b = j % D, min(n, a + B)
            collect(k)
            with torch.cuda.stream(self.streams[k]):
                self.sets[k].submit_batch(cams[a:b], targets[a:b], iters, td[a:b], gm[a:b])
            pending[k] = (a, b)
        for k in range(D):
            collect(k)
        for s_ in self.streams:
            main.wait_stream(s_)
        return results
