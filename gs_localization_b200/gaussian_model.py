"""Map side of LoGS on the B200 library: the GaussianModel of gaussian_splatting/scene/gaussian_model.py with the
same method names and argument meaning, re-laid-out for a training iteration without framework autograd.

Differences in layout (not in results):
  * SH features live in ONE [P,M,3] tensor; `_features_dc` / `_features_rest` are views of it.  The reference
    concatenates the two on every render (gaussian_model.py:108-111) and autograd splits the gradient again.
  * The activated copies the rasterizer reads (sigmoid opacity, exp scaling, normalised rotation) are persistent
    buffers refreshed by the fused optimiser kernel (gsr_map_adam_step), which also applies the chain rule through
    the activations, Adam on all six groups and the densification statistics.
  * Adam state is held as plain tensors next to the parameters (torch.optim.Adam's state dict in the reference),
    so densification edits parameters and state with one gather each.

`training_step` = rasterize forward -> fused L1+SSIM loss/gradient (+ fused depth terms) -> rasterize backward
-> fused optimiser kernel: 7scenes_gs_full_dslam.py:159-242 without its dataset and logging."""
from __future__ import annotations

import ctypes as C
import math
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib
from . import io as gio
from .diff_gaussian_rasterization import _C as _rast
from .simple_knn._C import distCUDA2

SH_C0 = 0.28209479177387814
GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


def inverse_sigmoid(x):
    return torch.log(x / (1 - x))


def RGB2SH(rgb):
    return (rgb - 0.5) / SH_C0


def get_expon_lr_func(lr_init, lr_final, lr_delay_steps=0, lr_delay_mult=1.0, max_steps=1000000):
    """utils/general_utils.py:29-62: log-linear interpolation with an optional sine warm-up."""
    def helper(step):
        if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
            return 0.0
        delay = 1.0
        if lr_delay_steps > 0:
            delay = lr_delay_mult + (1 - lr_delay_mult) * math.sin(0.5 * math.pi * min(max(step / lr_delay_steps, 0.0), 1.0))
        t = min(max(step / max_steps, 0.0), 1.0)
        return delay * math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t)
    return helper


def build_rotation(r):
    """utils/general_utils.py:95-118 (wxyz quaternion, normalised here) -> [N,3,3]."""
    q = r / torch.sqrt((r * r).sum(dim=1))[:, None]
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=1).reshape(-1, 3, 3)


def default_training_args(**over):
    """arguments/__init__.py:72-92 (OptimizationParams)."""
    d = dict(iterations=30_000, position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01,
             position_lr_max_steps=30_000, feature_lr=0.0025, opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001,
             percent_dense=0.01, lambda_dssim=0.2, densification_interval=100, opacity_reset_interval=3000,
             densify_from_iter=500, densify_until_iter=15_000, densify_grad_threshold=0.0002)
    d.update(over)
    return SimpleNamespace(**d)


class GaussianModel:
    def __init__(self, sh_degree: int, device="cuda"):
        self.active_sh_degree = 0
        self.max_sh_degree = sh_degree
        self.device = torch.device(device)
        e = torch.empty(0, device=self.device)
        self._xyz = self._features = self._scaling = self._rotation = self._opacity = e
        self.max_radii2D = self.xyz_gradient_accum = self.denom = e
        self.percent_dense = 0
        self.spatial_lr_scale = 0
        self._state = None          # {"m": [5 tensors], "v": [5 tensors]} in the order xyz, features, opacity, scaling, rotation
        self._lrs = None
        self._skip = set()          # groups whose parameter was replaced since the last gradient (reference: .grad is None)
        self._steps = {g: 0 for g in GROUPS}   # torch.optim.Adam keeps one step counter per parameter

    # ------------------------------------------------------------------ reference-shaped accessors
    @property
    def _features_dc(self):
        return self._features[:, :1]

    @property
    def _features_rest(self):
        return self._features[:, 1:]

    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_features(self):
        return self._features

    @property
    def get_opacity(self):
        return self._opacity_act

    @property
    def get_scaling(self):
        return self._scaling_act

    @property
    def get_rotation(self):
        return self._rotation_act

    def oneupSHdegree(self):
        if self.active_sh_degree < self.max_sh_degree:
            self.active_sh_degree += 1

    def _refresh_activations(self):
        self._opacity_act = torch.sigmoid(self._opacity)
        self._scaling_act = torch.exp(self._scaling)
        self._rotation_act = torch.nn.functional.normalize(self._rotation)

    def _params(self):
        return [self._xyz, self._features, self._opacity, self._scaling, self._rotation]

    def _set_params(self, tensors):
        self._xyz, self._features, self._opacity, self._scaling, self._rotation = [t.contiguous() for t in tensors]
        self._refresh_activations()

    # ------------------------------------------------------------------ construction
    def create_from_pcd(self, points, colors, spatial_lr_scale: float):
        """gaussian_model.py:124-148; `points`, `colors` are [N,3] arrays (BasicPointCloud.points / .colors)."""
        self.spatial_lr_scale = spatial_lr_scale
        dev = self.device
        xyz = torch.as_tensor(np.asarray(points)).float().to(dev)
        M = (self.max_sh_degree + 1) ** 2
        features = torch.zeros((xyz.shape[0], M, 3), device=dev)
        features[:, 0, :] = RGB2SH(torch.as_tensor(np.asarray(colors)).float().to(dev))
        dist2 = torch.clamp_min(distCUDA2(xyz), 0.0000001)
        scales = torch.log(torch.sqrt(dist2))[..., None].repeat(1, 3)
        rots = torch.zeros((xyz.shape[0], 4), device=dev)
        rots[:, 0] = 1
        opacities = inverse_sigmoid(0.1 * torch.ones((xyz.shape[0], 1), dtype=torch.float, device=dev))
        self._set_params([xyz, features, opacities, scales, rots])
        self.max_radii2D = torch.zeros(xyz.shape[0], device=dev)

    def from_raw(self, raw: gio.RawGaussians):
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(self.device)
        self._set_params([t(raw.xyz), torch.cat([t(raw.features_dc), t(raw.features_rest)], dim=1), t(raw.opacity), t(raw.scaling),
                          t(raw.rotation)])
        self.max_radii2D = torch.zeros(self._xyz.shape[0], device=self.device)
        self.active_sh_degree = self.max_sh_degree

    def to_raw(self) -> gio.RawGaussians:
        n = lambda x: x.detach().cpu().numpy()
        return gio.RawGaussians(n(self._xyz), n(self._features_dc), n(self._features_rest), n(self._opacity), n(self._scaling),
                                n(self._rotation))

    def save_ply(self, path):
        gio.save_ply(path, self.to_raw())

    def load_ply(self, path):
        self.from_raw(gio.load_ply_raw(path, self.max_sh_degree))

    # ------------------------------------------------------------------ optimiser
    def training_setup(self, training_args):
        """gaussian_model.py:150-173: six Adam groups (eps 1e-15), exponential position schedule."""
        self.percent_dense = training_args.percent_dense
        N = self._xyz.shape[0]
        self.xyz_gradient_accum = torch.zeros((N, 1), device=self.device)
        self.denom = torch.zeros((N, 1), device=self.device)
        self._lrs = {"xyz": training_args.position_lr_init * self.spatial_lr_scale, "f_dc": training_args.feature_lr,
                     "f_rest": training_args.feature_lr / 20.0, "opacity": training_args.opacity_lr,
                     "scaling": training_args.scaling_lr, "rotation": training_args.rotation_lr}
        self._state = {"m": [torch.zeros_like(p) for p in self._params()], "v": [torch.zeros_like(p) for p in self._params()]}
        self._steps = {g: 0 for g in GROUPS}
        self.xyz_scheduler_args = get_expon_lr_func(lr_init=training_args.position_lr_init * self.spatial_lr_scale,
                                                    lr_final=training_args.position_lr_final * self.spatial_lr_scale,
                                                    lr_delay_mult=training_args.position_lr_delay_mult,
                                                    max_steps=training_args.position_lr_max_steps)

    def update_learning_rate(self, iteration):
        self._lrs["xyz"] = self.xyz_scheduler_args(iteration)
        return self._lrs["xyz"]

    def optimizer_step(self, grads, dL_dmeans2D=None, radii=None, stats=False, adam=True):
        """`gaussians.optimizer.step()` (+ the statistics of 7scenes_gs_full_dslam.py:229-230 when `stats`).
        grads = (dL_dxyz, dL_dfeatures, dL_dopacity_act, dL_dscaling_act, dL_drotation_act)."""
        lib = _lib.load()
        P, M = int(self._xyz.shape[0]), int(self._features.shape[1])
        if adam:
            for g in GROUPS:
                if g not in self._skip:
                    self._steps[g] += 1
        steps = (C.c_int * 6)(*[max(self._steps[g], 1) for g in GROUPS])
        lrs = (C.c_float * 6)(*[(-1.0 if g in self._skip else float(self._lrs[g])) for g in GROUPS])
        tab = lambda ts: (C.c_void_p * 5)(*[None if t is None else t.data_ptr() for t in ts])
        p = lambda t: None if t is None else t.data_ptr()
        _lib.check(lib.gsr_map_adam_step(
            P, M, int(stats), int(adam), steps, lrs, 0.9, 0.999, 1e-15, tab(self._params()),
            tab(grads if adam else [None] * 5), tab(self._state["m"]), tab(self._state["v"]), p(self._opacity_act), p(self._scaling_act),
            p(self._rotation_act), p(dL_dmeans2D), p(radii), p(self.max_radii2D), p(self.xyz_gradient_accum), p(self.denom),
            torch.cuda.current_stream(self.device).cuda_stream), "gsr_map_adam_step")
        if adam:
            self._skip.clear()

    # ------------------------------------------------------------------ densification (gaussian_model.py:258-407)
    def reset_opacity(self):
        new = inverse_sigmoid(torch.min(self.get_opacity, torch.ones_like(self.get_opacity) * 0.01))
        self._opacity = new.contiguous()
        self._state["m"][2] = torch.zeros_like(new)
        self._state["v"][2] = torch.zeros_like(new)
        self._opacity_act = torch.sigmoid(self._opacity)
        self._skip.add("opacity")

    def _gather(self, src, fresh):
        """Rebuild parameters and Adam state from rows `src` of the current ones; rows flagged `fresh` start with zero state."""
        keep = (~fresh).to(torch.float32)
        self._set_params([p.index_select(0, src) for p in self._params()])
        for k in ("m", "v"):
            self._state[k] = [(s.index_select(0, src) * keep.view(-1, *([1] * (s.dim() - 1)))).contiguous() for s in self._state[k]]

    def prune_points(self, mask):
        valid = (~mask).nonzero(as_tuple=True)[0]
        self._gather(valid, torch.zeros(valid.shape[0], dtype=torch.bool, device=self.device))
        self.xyz_gradient_accum = self.xyz_gradient_accum[valid]
        self.denom = self.denom[valid]
        self.max_radii2D = self.max_radii2D[valid]
        self._skip.update(GROUPS)

    def densify_and_prune(self, max_grad, min_opacity, extent, max_screen_size):
        """clone small high-gradient Gaussians, split large ones in two (children sampled inside the parent, scale
        / 1.6), drop the split parents, then prune by opacity / size — same order, same surviving row order and the
        same torch.normal draw as the reference (gaussian_model.py:352-402)."""
        N = self._xyz.shape[0]
        grads = self.xyz_gradient_accum / self.denom
        grads[grads.isnan()] = 0.0
        gnorm = torch.norm(grads, dim=-1)
        big = torch.max(self.get_scaling, dim=1).values > self.percent_dense * extent
        clone = (gnorm >= max_grad) & ~big
        # the split test sees the clones too (their padded gradient is zero)
        n_clone = int(clone.sum())
        split = (grads.squeeze(-1) >= max_grad) & big
        if max_grad <= 0:
            raise ValueError("densify_grad_threshold must be positive")
        ci, si = clone.nonzero(as_tuple=True)[0], split.nonzero(as_tuple=True)[0]
        n_split = int(si.shape[0])
        stds = self.get_scaling[si].repeat(2, 1)
        samples = torch.normal(mean=torch.zeros((stds.size(0), 3), device=self.device), std=stds)
        rots = build_rotation(self._rotation[si]).repeat(2, 1, 1)
        child_xyz = torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + self._xyz[si].repeat(2, 1)
        child_scaling = torch.log(self.get_scaling[si].repeat(2, 1) / (0.8 * 2))
        keep = (~split).nonzero(as_tuple=True)[0]
        src = torch.cat([keep, ci, si, si])
        fresh = torch.cat([torch.zeros(keep.shape[0], dtype=torch.bool, device=self.device),
                           torch.ones(n_clone + 2 * n_split, dtype=torch.bool, device=self.device)])
        self._gather(src, fresh)
        first_child = keep.shape[0] + n_clone
        self._xyz[first_child:] = child_xyz
        self._scaling[first_child:] = child_scaling
        self._refresh_activations()
        total = src.shape[0]
        self.xyz_gradient_accum = torch.zeros((total, 1), device=self.device)
        self.denom = torch.zeros((total, 1), device=self.device)
        self.max_radii2D = torch.zeros(total, device=self.device)

        prune_mask = (self.get_opacity < min_opacity).squeeze(-1)
        if max_screen_size:
            big_vs = self.max_radii2D > max_screen_size        # all zeros here: densification_postfix just reset it (reference quirk)
            big_ws = self.get_scaling.max(dim=1).values > 0.1 * extent
            prune_mask = prune_mask | big_vs | big_ws
        self.prune_points(prune_mask)
        return dict(before=N, cloned=n_clone, split=n_split, pruned=int(prune_mask.sum()), after=int(self._xyz.shape[0]))

    def add_densification_stats(self, viewspace_grad, update_filter):
        """gaussian_model.py:405-407 as a stand-alone call (the fused path does this inside optimizer_step)."""
        self.xyz_gradient_accum[update_filter] += torch.norm(viewspace_grad[update_filter, :2], dim=-1, keepdim=True)
        self.denom[update_filter] += 1

    # ------------------------------------------------------------------ one training iteration, no autograd
    def compute_gradients(self, cam, gt_image, bg, opt, iteration: int, pseudo_depth=None, gt_depth=None, after_forward=None):
        """Render + loss + backward of one view (7scenes_gs_full_dslam.py:128-190) without touching the parameters.
        Returns (loss[1], grads, dL_dmeans2D, outputs); grads = (dL_dxyz, dL_dfeatures, dL_dopacity_act, dL_dscaling_act,
        dL_drotation_act), dense over all Gaussians and exactly zero outside `outputs["radii"] > 0`."""
        lib = _lib.load()
        dev = self.device
        self.update_learning_rate(iteration)
        if iteration % 1000 == 0:
            self.oneupSHdegree()
        view, proj, _, campos = cam.matrices(dev)
        e = torch.Tensor([])
        fwd = _rast._forward_impl(bg, self._xyz, e, self._opacity_act, self._scaling_act, self._rotation_act, 1.0, e, view, proj,
                                  cam.tanfovx, cam.tanfovy, cam.H, cam.W, self._features, self.active_sh_degree, campos, False, False)
        R, color, depth, alpha, radii, geom, binning, img, _ = fwd
        if after_forward is not None:
            after_forward(radii)        # e.g. the data-parallel exchange starts counting its rows while the backward runs
        stream = torch.cuda.current_stream(dev).cuda_stream
        p = lambda t: None if t is None else t.data_ptr()
        loss = torch.zeros(1, device=dev)
        dL_dcolor = torch.empty_like(color)
        scratch = torch.empty(3 * color.numel() + 2, device=dev)
        _lib.check(lib.gsr_l1_ssim_loss_grad(p(color), p(gt_image), 3, cam.H, cam.W, float(opt.lambda_dssim), p(loss), p(dL_dcolor),
                                             p(scratch), stream), "gsr_l1_ssim_loss_grad")
        dL_ddepth = torch.zeros_like(depth)
        if pseudo_depth is not None or gt_depth is not None:
            dscratch = torch.empty(9, dtype=torch.float64, device=dev)
            _lib.check(lib.gsr_depth_loss_grad(p(depth), p(pseudo_depth), p(gt_depth), depth.numel(), 1000.0, 0.01, 0.05, p(loss),
                                               p(dL_ddepth), p(dscratch), stream), "gsr_depth_loss_grad")
        needs = dict(means3D=True, means2D=True, sh=True, colors=False, opacity=True, scales=True, rotations=True, cov3D=False)
        grads, _ = _rast._backward_impl(bg, self._xyz, radii, e, self._scaling_act, self._rotation_act, 1.0, e, view, proj, cam.tanfovx,
                                        cam.tanfovy, dL_dcolor, dL_ddepth, torch.zeros_like(alpha), self._features, self.active_sh_degree,
                                        campos, geom, R, binning, img, alpha, False, needs=needs)
        dL_dmeans2D, _, dL_dopacity, dL_dmeans3D, _, dL_dsh, dL_dscales, dL_drotations = grads
        g = (dL_dmeans3D, dL_dsh, dL_dopacity, dL_dscales, dL_drotations)
        return loss, g, dL_dmeans2D, dict(render=color, depth=depth, alpha=alpha, radii=radii)

    def apply_gradients(self, g, dL_dmeans2D, radii, opt, iteration: int, extent: float = 1.0, stats_done: bool = False):
        """Statistics, densification / opacity reset on their schedule, and the optimiser step
        (7scenes_gs_full_dslam.py:225-242).  Under data parallelism `g` is the sum over ranks and `stats_done=True` says
        that the densification statistics of ALL views of the step have already been accumulated by the exchange
        (parallel.DataParallelTrainer): every replica then holds the same statistics and takes the same densification
        decisions.  Passing only the local dL_dmeans2D / radii in a data-parallel run would let the replicas diverge at
        the first densification."""
        info = None
        with_stats = not stats_done
        if iteration < opt.densify_until_iter:
            densify = iteration > opt.densify_from_iter and iteration % opt.densification_interval == 0
            reset = iteration % opt.opacity_reset_interval == 0
            if densify or reset:
                if with_stats:
                    self.optimizer_step(g, dL_dmeans2D, radii, stats=True, adam=False)
                if densify:
                    info = self.densify_and_prune(opt.densify_grad_threshold, 0.005, extent,
                                                  20 if iteration > opt.opacity_reset_interval else None)
                if reset:
                    self.reset_opacity()
                if not densify:          # only the opacity group lost its gradient
                    self.optimizer_step(g, adam=True)
                else:
                    self._skip.clear()   # every parameter was replaced: optimizer.step() finds no gradients at all
            elif with_stats:
                self.optimizer_step(g, dL_dmeans2D, radii, stats=True, adam=True)
            else:
                self.optimizer_step(g, adam=True)
        else:
            self.optimizer_step(g, adam=True)
        return info

    def training_step(self, cam, gt_image, bg, opt, iteration: int, pseudo_depth=None, gt_depth=None, extent: float = 1.0):
        """7scenes_gs_full_dslam.py:128-242 for one view: returns (loss tensor[1], dict of render outputs)."""
        loss, g, dL_dmeans2D, out = self.compute_gradients(cam, gt_image, bg, opt, iteration, pseudo_depth, gt_depth)
        out["densify"] = self.apply_gradients(g, dL_dmeans2D, out["radii"], opt, iteration, extent)
        return loss, out
