"""Call-site conformance of the drop-in packages, run in a fresh interpreter (tests/test_gpu_callsite.py).

`gs_localization_b200/dropin` goes FIRST on sys.path; the reference's own import lines then resolve to this repo:

    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer     (gaussian_renderer/__init__.py:14)
    from diff_gaussian_rasterization_pose import GaussianRasterizationSettings, GaussianRasterizer (pipelines/tools/__init__.py:15-18)
    from simple_knn._C import distCUDA2                                                            (scene/gaussian_model.py:20)

and the two `render` functions are replayed with the keyword calls, tensor shapes and autograd set-up of
gaussian_splatting/gaussian_renderer/__init__.py:36-95 and gs_localization/pipelines/tools/__init__.py:58-141
(`screenspace_points = zeros_like(xyz, requires_grad=True) + 0` with retain_grad, `[P,1]` opacities from a sigmoid,
`shs = pc.get_features` as a cat of f_dc / f_rest, exp / normalize activations, masked branch by boolean indexing).
The same replay then runs on the UNMODIFIED reference package (oracle/_ref, imported under an alias so both can live
in one process) and the results are compared.  Prints one JSON line; exit status 0 = conformant."""
import importlib.util
import json
import math
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "gs_localization_b200", "dropin"))
sys.path.insert(1, ROOT)

import torch  # noqa: E402

import diff_gaussian_rasterization as dgr  # noqa: E402  (the drop-in)
import diff_gaussian_rasterization_pose as dgr_pose  # noqa: E402
from simple_knn._C import distCUDA2  # noqa: E402

from gs_localization_b200 import synthetic as syn  # noqa: E402

assert os.path.join("gs_localization_b200", "dropin") in dgr.__file__, dgr.__file__
assert os.path.join("gs_localization_b200", "dropin") in dgr_pose.__file__, dgr_pose.__file__


def load_reference_alias():
    d = os.path.join(ROOT, "oracle", "_ref", "diff_gaussian_rasterization")
    if not os.path.exists(os.path.join(d, "_C.so")):
        return None
    spec = importlib.util.spec_from_file_location("ref_dgr", os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_dgr"] = mod
    spec.loader.exec_module(mod)
    return mod


class Model:
    """Stand-in for GaussianModel's property surface (scene/gaussian_model.py:84-111): raw parameters + activations."""

    def __init__(self, m: syn.GaussianMap, dev):
        P = m.means3D.shape[0]
        par = lambda t: torch.nn.Parameter(t.to(dev).contiguous().requires_grad_(True))
        self._xyz = par(m.means3D)
        self._features_dc = par(m.shs[:, :1, :])
        self._features_rest = par(m.shs[:, 1:, :])
        self._opacity = par(torch.logit(m.opacities.clamp(1e-4, 1 - 1e-4)))
        self._scaling = par(torch.log(m.scales))
        self._rotation = par(m.rotations * 1.7)          # unnormalised storage, normalised by the getter
        self.active_sh_degree = self.max_sh_degree = m.sh_degree
        assert self._opacity.shape == (P, 1)

    get_xyz = property(lambda s: s._xyz)
    get_opacity = property(lambda s: torch.sigmoid(s._opacity))
    get_scaling = property(lambda s: torch.exp(s._scaling))
    get_rotation = property(lambda s: torch.nn.functional.normalize(s._rotation))
    get_features = property(lambda s: torch.cat((s._features_dc, s._features_rest), dim=1))

    def params(self):
        return [self._xyz, self._features_dc, self._features_rest, self._opacity, self._scaling, self._rotation]


def viewpoint(cam: syn.Camera, dev):
    view, full, raw, campos = cam.matrices(dev)
    vp = types.SimpleNamespace(FoVx=2 * math.atan(cam.tanfovx), FoVy=2 * math.atan(cam.tanfovy), image_height=cam.H, image_width=cam.W,
                               world_view_transform=view, full_proj_transform=full, projection_matrix=raw, camera_center=campos)
    vp.cam_rot_delta = torch.nn.Parameter(torch.zeros(3, device=dev))
    vp.cam_trans_delta = torch.nn.Parameter(torch.zeros(3, device=dev))
    return vp


def render_plain(pkg, viewpoint_camera, pc, bg_color, scaling_modifier=1.0):
    """keyword call of gaussian_renderer/__init__.py:36-95"""
    screenspace_points = torch.zeros_like(pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True, device="cuda") + 0
    screenspace_points.retain_grad()
    raster_settings = pkg.GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5), tanfovy=math.tan(viewpoint_camera.FoVy * 0.5), bg=bg_color,
        scale_modifier=scaling_modifier, viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform, sh_degree=pc.active_sh_degree, campos=viewpoint_camera.camera_center,
        prefiltered=False, debug=False)
    rasterizer = pkg.GaussianRasterizer(raster_settings=raster_settings)
    rendered_image, radii, depth, alpha = rasterizer(
        means3D=pc.get_xyz, means2D=screenspace_points, shs=pc.get_features, colors_precomp=None, opacities=pc.get_opacity,
        scales=pc.get_scaling, rotations=pc.get_rotation, cov3D_precomp=None)
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii,
            "depth": depth, "alpha": alpha}


def render_pose(pkg, viewpoint_camera, pc, bg_color, scaling_modifier=1.0, mask=None):
    """keyword calls of pipelines/tools/__init__.py:58-141 (both branches)"""
    screenspace_points = torch.zeros_like(pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True, device="cuda") + 0
    screenspace_points.retain_grad()
    raster_settings = pkg.GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5), tanfovy=math.tan(viewpoint_camera.FoVy * 0.5), bg=bg_color,
        scale_modifier=scaling_modifier, viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform, projmatrix_raw=viewpoint_camera.projection_matrix,
        sh_degree=pc.active_sh_degree, campos=viewpoint_camera.camera_center, prefiltered=False, debug=False)
    rasterizer = pkg.GaussianRasterizer(raster_settings=raster_settings)
    means3D, means2D, opacity = pc.get_xyz, screenspace_points, pc.get_opacity
    scales, rotations, shs = pc.get_scaling, pc.get_rotation, pc.get_features
    if mask is not None:
        out = rasterizer(means3D=means3D[mask], means2D=means2D[mask], shs=shs[mask], colors_precomp=None, opacities=opacity[mask],
                         scales=scales[mask], rotations=rotations[mask], cov3D_precomp=None, theta=viewpoint_camera.cam_rot_delta,
                         rho=viewpoint_camera.cam_trans_delta)
    else:
        out = rasterizer(means3D=means3D, means2D=means2D, shs=shs, colors_precomp=None, opacities=opacity, scales=scales,
                         rotations=rotations, cov3D_precomp=None, theta=viewpoint_camera.cam_rot_delta,
                         rho=viewpoint_camera.cam_trans_delta)
    rendered_image, radii, depth, opacity_img, n_touched = out     # the upstream rasterizer returns five values on both branches
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii,
            "depth": depth, "opacity": opacity_img, "n_touched": n_touched}


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main():
    dev = torch.device("cuda:0")
    cfg = dict(P=40_000, W=320, H=240, deg=3, f=262.5, box=1.0, sigma0=0.04)
    gmap = syn.make_map(cfg["P"], cfg["deg"], cfg["sigma0"], cfg["box"], seed=0)
    cam = syn.make_camera(cfg, 3)
    bg = torch.tensor([0.0, 0.0, 0.0], device=dev)
    g = torch.Generator().manual_seed(0)
    w_img = torch.rand(3, cam.H, cam.W, generator=g).to(dev)
    w_dep = (torch.rand(1, cam.H, cam.W, generator=g) * 0.1).to(dev)
    report = {}

    def run(render_fn, pkg, **kw):
        pc, vp = Model(gmap, dev), viewpoint(cam, dev)
        pkt = render_fn(pkg, vp, pc, bg, **kw)
        loss = (pkt["render"] * w_img).sum() + (pkt["depth"] * w_dep).sum()
        loss.backward()
        torch.cuda.synchronize()
        grads = [p.grad.detach().clone() for p in pc.params()]
        return pkt, grads, pkt["viewspace_points"].grad.detach().clone(), vp

    # ---- plain package: ours vs the reference build, same call
    pkt, grads, vsg, _ = run(render_plain, dgr)
    assert pkt["render"].shape == (3, cam.H, cam.W) and pkt["depth"].shape == (1, cam.H, cam.W) and pkt["alpha"].shape == (1, cam.H, cam.W)
    assert pkt["radii"].dtype == torch.int32 and pkt["visibility_filter"].dtype == torch.bool and int(pkt["visibility_filter"].sum()) > 0
    assert vsg.shape == (cfg["P"], 3) and float(vsg[~pkt["visibility_filter"]].abs().sum()) == 0.0
    # densification statistic of the training loop (gaussian_model.py:405-407): norm of viewspace grad[:, :2] of the visible rows
    stat = torch.norm(vsg[pkt["visibility_filter"], :2], dim=-1, keepdim=True)
    assert bool(torch.isfinite(stat).all()) and float(stat.sum()) > 0
    ref = load_reference_alias()
    report["reference_build"] = ref is not None
    if ref is not None:
        rpkt, rgrads, rvsg, _ = run(render_plain, ref)
        assert torch.equal(pkt["radii"], rpkt["radii"])
        assert float((pkt["alpha"] - rpkt["alpha"]).abs().max()) == 0.0
        assert float((pkt["render"] - rpkt["render"]).abs().max()) <= 1e-4
        assert float((pkt["depth"] - rpkt["depth"]).abs().max()) <= 1e-4 * max(1.0, float(rpkt["depth"].abs().max()))
        report["viewspace_grad_rel"] = rel(vsg, rvsg)
        assert report["viewspace_grad_rel"] <= 1e-3
        report["param_grad_rel"] = [rel(a, b) for a, b in zip(grads, rgrads)]
        assert max(report["param_grad_rel"]) <= 1e-3, report["param_grad_rel"]

    # ---- pose package: same images / parameter gradients as the plain call, five return values, pose gradients flow
    ppkt, pgrads, pvsg, vp = run(render_pose, dgr_pose)
    assert torch.equal(ppkt["render"], pkt["render"]) and torch.equal(ppkt["radii"], pkt["radii"])
    assert ppkt["n_touched"].shape == (cfg["P"],) and int(ppkt["n_touched"][~ppkt["visibility_filter"]].sum()) == 0
    assert max(rel(a, b) for a, b in zip(pgrads, grads)) <= 1e-5 and rel(pvsg, vsg) <= 1e-5
    assert vp.cam_rot_delta.grad is not None and vp.cam_trans_delta.grad is not None
    assert float(vp.cam_rot_delta.grad.abs().sum()) > 0 and float(vp.cam_trans_delta.grad.abs().sum()) > 0
    # masked branch (:117-128): a boolean mask over the Gaussians; must equal rendering the sub-map
    mask = torch.zeros(cfg["P"], dtype=torch.bool, device=dev)
    mask[::2] = True
    mpkt, mgrads, mvsg, mvp = run(render_pose, dgr_pose, mask=mask)
    sub = syn.GaussianMap(gmap.means3D[::2], gmap.shs[::2], gmap.opacities[::2], gmap.scales[::2], gmap.rotations[::2], gmap.sh_degree)
    pc_sub, vp_sub = Model(sub, dev), viewpoint(cam, dev)
    spkt = render_pose(dgr_pose, vp_sub, pc_sub, bg)
    assert torch.equal(mpkt["render"], spkt["render"]) and mpkt["radii"].shape == (int(mask.sum()),)
    assert float(mgrads[0][~mask].abs().sum()) == 0.0 and float(mvsg[~mask].abs().sum()) == 0.0   # unselected rows get no gradient

    # ---- simple_knn
    pts = gmap.means3D[:20_000].to(dev)
    d2 = distCUDA2(pts)
    assert d2.shape == (20_000,) and d2.dtype == torch.float32 and bool((d2 > 0).all())
    sk = os.path.join(ROOT, "oracle", "_ref", "simple_knn")
    if os.path.exists(os.path.join(sk, "_C.so")):
        spec = importlib.util.spec_from_file_location("ref_simple_knn._C", os.path.join(sk, "_C.so"))
        try:
            rk = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(rk)
            report["knn_bit_exact"] = bool(torch.equal(d2, rk.distCUDA2(pts)))
            assert report["knn_bit_exact"]
        except ImportError as ex:           # built under another module name: covered by tests/test_knn.py
            report["knn_bit_exact"] = f"reference module not importable here: {ex}"
    report["ok"] = True
    print(json.dumps(report))


if __name__ == "__main__":
    main()
