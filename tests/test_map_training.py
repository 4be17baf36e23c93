"""Map-training side: densification against the sequential oracle (CPU), the fused optimiser
kernel against torch.optim.Adam + autograd through the activations, and a whole training iteration against the
framework formulation of the same iteration (GPU)."""
import numpy as np
import pytest
import torch

from gs_localization_b200 import gaussian_model as gm
from gs_localization_b200 import io as gio
from gs_localization_b200 import synthetic as syn
from oracle import densify_oracle


def _raw(P, deg, seed, spread=1.0):
    g = syn.make_map(P, deg, 0.05, spread, seed=seed)
    raw = gio.deactivate(g)
    rng = np.random.default_rng(seed)
    return raw._replace(rotation=(raw.rotation * (0.5 + rng.random((P, 1)))).astype(np.float32))   # unnormalised, like a trained map


def test_densify_and_prune_matches_sequential_oracle():
    P, deg = 4000, 2
    raw = _raw(P, deg, 1)
    model = gm.GaussianModel(deg, device="cpu")
    model.from_raw(raw)
    args = gm.default_training_args()
    model.spatial_lr_scale = 1.0
    model.training_setup(args)
    g = torch.Generator().manual_seed(0)
    for k in ("m", "v"):
        model._state[k] = [torch.rand(p.shape, generator=g) for p in model._params()]
    model.xyz_gradient_accum = torch.rand(P, 1, generator=g) * 0.002
    model.denom = torch.randint(0, 4, (P, 1), generator=g).float()          # zeros give NaN -> 0
    model.max_radii2D = torch.rand(P, generator=g) * 40
    extent = 5.0
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    names = densify_oracle.NAMES
    split5 = lambda ts: {"xyz": ts[0], "f_dc": ts[1][:, :1], "f_rest": ts[1][:, 1:], "opacity": ts[2], "scaling": ts[3], "rotation": ts[4]}
    S = {"p": {k: v.clone() for k, v in split5(model._params()).items()}, "m": {k: v.clone() for k, v in split5(model._state["m"]).items()},
         "v": {k: v.clone() for k, v in split5(model._state["v"]).items()}, "accum": model.xyz_gradient_accum.clone(),
         "denom": model.denom.clone(), "max_radii2D": model.max_radii2D.clone()}
    torch.manual_seed(123)
    densify_oracle.densify_and_prune(S, args.densify_grad_threshold, 0.005, extent, 20, args.percent_dense)
    torch.manual_seed(123)
    info = model.densify_and_prune(args.densify_grad_threshold, 0.005, extent, 20)
    assert info["cloned"] > 50 and info["split"] > 50 and info["pruned"] > 0, info
    assert info["after"] == S["p"]["xyz"].shape[0] == info["before"] + info["cloned"] + info["split"] - info["pruned"]
    got = {"p": split5(model._params()), "m": split5(model._state["m"]), "v": split5(model._state["v"])}
    for w in ("p", "m", "v"):
        for k in names:
            assert torch.equal(got[w][k], S[w][k]), (w, k)
    assert torch.equal(model.xyz_gradient_accum, S["accum"]) and torch.equal(model.denom, S["denom"]) and torch.equal(model.max_radii2D, S["max_radii2D"])
    assert torch.equal(model.get_opacity, torch.sigmoid(model._opacity)) and model._skip == set(gm.GROUPS)


def test_expon_lr_and_reset_opacity_cpu():
    f = gm.get_expon_lr_func(1.6e-4, 1.6e-6, lr_delay_mult=0.01, max_steps=30000)
    assert abs(f(0) - 1.6e-4) < 1e-12 and abs(f(30000) - 1.6e-6) < 1e-12 and abs(f(15000) - 1.6e-5) < 1e-11 and f(-1) == 0.0
    g = gm.get_expon_lr_func(1.0, 0.01, lr_delay_steps=100, lr_delay_mult=0.1, max_steps=1000)
    assert abs(g(0) - 0.1) < 1e-12 and abs(g(100) - 10 ** (-0.2)) < 1e-12
    model = gm.GaussianModel(1, device="cpu")
    model.from_raw(_raw(100, 1, 2))
    model.spatial_lr_scale = 1.0
    model.training_setup(gm.default_training_args())
    model._state["m"][2] += 1.0
    model.reset_opacity()
    assert float(model.get_opacity.max()) <= 0.01 + 1e-7 and float(model._state["m"][2].abs().max()) == 0.0 and model._skip == {"opacity"}


# --------------------------------------------------------------------------------------------------- GPU
def _torch_reference_model(raw, dev, lrs):
    t = lambda a: torch.nn.Parameter(torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(dev))
    P = dict(xyz=t(raw.xyz), f_dc=t(raw.features_dc), f_rest=t(raw.features_rest), opacity=t(raw.opacity), scaling=t(raw.scaling),
             rotation=t(raw.rotation))
    opt = torch.optim.Adam([{"params": [P[k]], "lr": lrs[k], "name": k} for k in gm.GROUPS], lr=0.0, eps=1e-15)
    return P, opt


@pytest.mark.gpu
@pytest.mark.parametrize("deg", [0, 2, 3])
def test_fused_optimizer_matches_torch_adam(deg):
    dev = "cuda:0"
    Pn = 5003
    raw = _raw(Pn, deg, 4)
    model = gm.GaussianModel(deg, device=dev)
    model.from_raw(raw)
    model.spatial_lr_scale = 2.0
    args = gm.default_training_args()
    model.training_setup(args)
    P, opt = _torch_reference_model(raw, dev, dict(model._lrs))
    g = torch.Generator(device="cpu").manual_seed(7)
    for it in range(4):
        G = [torch.randn(p.shape, generator=g).to(dev) * 1e-3 for p in model._params()]
        radii = (torch.rand(Pn, generator=g) * 30 - 10).clamp(min=0).int().to(dev)
        g2d = torch.randn(Pn, 3, generator=g).to(dev)
        want_accum = model.xyz_gradient_accum.clone()
        vis = radii > 0
        want_accum[vis] += g2d[vis, :2].norm(dim=-1, keepdim=True)
        want_denom = model.denom + vis[:, None].float()
        want_radii = torch.where(vis, torch.max(model.max_radii2D, radii.float()), model.max_radii2D)
        if it == 2:                                   # reset_opacity between backward and step: that group is skipped
            model.reset_opacity()
            with torch.no_grad():
                new = gm.inverse_sigmoid(torch.min(torch.sigmoid(P["opacity"]), torch.ones_like(P["opacity"]) * 0.01))
            st = opt.state[P["opacity"]]
            st["exp_avg"], st["exp_avg_sq"] = torch.zeros_like(new), torch.zeros_like(new)
            del opt.state[P["opacity"]]
            P["opacity"] = torch.nn.Parameter(new)
            opt.param_groups[3]["params"][0] = P["opacity"]
            opt.state[P["opacity"]] = st
        feats = torch.cat([P["f_dc"], P["f_rest"]], dim=1)
        loss = (P["xyz"] * G[0]).sum() + (feats * G[1]).sum() + (torch.sigmoid(P["opacity"]) * G[2]).sum() + \
            (torch.exp(P["scaling"]) * G[3]).sum() + (torch.nn.functional.normalize(P["rotation"]) * G[4]).sum()
        opt.zero_grad(set_to_none=True)
        if it == 2:
            grads = torch.autograd.grad(loss, [P[k] for k in gm.GROUPS if k != "opacity"])
            for k, gr in zip([k for k in gm.GROUPS if k != "opacity"], grads):
                P[k].grad = gr
        else:
            loss.backward()
        opt.step()
        model.optimizer_step(G, g2d, radii, stats=True, adam=True)
        ref = [P["xyz"], torch.cat([P["f_dc"], P["f_rest"]], dim=1), P["opacity"], P["scaling"], P["rotation"]]
        for name, a, b in zip(("xyz", "features", "opacity", "scaling", "rotation"), model._params(), ref):
            assert (a - b.detach()).abs().max().item() <= 2e-6 * max(1.0, b.abs().max().item()), (it, name)
        assert torch.allclose(model.get_opacity, torch.sigmoid(model._opacity), atol=1e-6)
        assert torch.allclose(model.get_scaling, torch.exp(model._scaling), rtol=1e-5)
        assert torch.allclose(model.get_rotation, torch.nn.functional.normalize(model._rotation), atol=1e-6)
        assert torch.allclose(model.xyz_gradient_accum, want_accum, rtol=1e-6) and torch.equal(model.denom, want_denom)
        assert torch.equal(model.max_radii2D, want_radii)
    assert model._steps["opacity"] == 3 and model._steps["xyz"] == 4


@pytest.mark.gpu
def test_training_step_matches_framework_iteration():
    """Three iterations of GaussianModel.training_step vs the same iterations written the reference's way: autograd
    through activations + drop-in rasterizer + torch l1/ssim + Pearson/L1 depth terms + torch.optim.Adam."""
    import torch.nn.functional as F
    from gs_localization_b200.diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    dev = "cuda:0"
    cfg = dict(P=30_000, W=200, H=152, deg=2, f=150.0, box=1.0, sigma0=0.05)
    raw = _raw(cfg["P"], cfg["deg"], 0)
    model = gm.GaussianModel(cfg["deg"], device=dev)
    model.from_raw(raw)
    model.spatial_lr_scale = 1.0
    args = gm.default_training_args()
    model.training_setup(args)
    P, opt = _torch_reference_model(raw, dev, dict(model._lrs))
    sched = gm.get_expon_lr_func(args.position_lr_init, args.position_lr_final, lr_delay_mult=args.position_lr_delay_mult,
                                 max_steps=args.position_lr_max_steps)
    g1 = torch.exp(-((torch.arange(11.0) - 5) ** 2) / (2 * 1.5 ** 2))
    g1 = g1 / g1.sum()
    window = (g1[:, None] @ g1[None, :])[None, None].expand(3, 1, 11, 11).contiguous().to(dev)

    def ssim(a, b):
        a, b = a[None], b[None]
        mu1, mu2 = F.conv2d(a, window, padding=5, groups=3), F.conv2d(b, window, padding=5, groups=3)
        s1 = F.conv2d(a * a, window, padding=5, groups=3) - mu1 * mu1
        s2 = F.conv2d(b * b, window, padding=5, groups=3) - mu2 * mu2
        s12 = F.conv2d(a * b, window, padding=5, groups=3) - mu1 * mu2
        return (((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 * mu1 + mu2 * mu2 + 1e-4) * (s1 + s2 + 9e-4))).mean()

    def pearson(x, y):
        xc, yc = x - x.mean(), y - y.mean()
        return (xc * yc).sum() / torch.sqrt((xc * xc).sum() * (yc * yc).sum())

    bg = torch.zeros(3, device=dev)
    gen = torch.Generator().manual_seed(5)
    for it in range(1, 4):
        cam = syn.make_camera(cfg, it)
        view, proj, _, campos = cam.matrices(dev)
        gt = torch.rand(3, cam.H, cam.W, generator=gen).to(dev)
        gt_depth = (torch.rand(1, cam.H, cam.W, generator=gen) * 3).to(dev)
        gt_depth[:, :10] = 0
        pseudo = (500.0 / (gt_depth + 1) + torch.randn(1, cam.H, cam.W, generator=gen).to(dev))
        # framework iteration
        opt.param_groups[0]["lr"] = sched(it)
        rs = GaussianRasterizationSettings(cam.H, cam.W, cam.tanfovx, cam.tanfovy, bg, 1.0, view, proj, model.active_sh_degree, campos, False, False)
        shs = torch.cat([P["f_dc"], P["f_rest"]], dim=1)
        m2 = torch.zeros_like(P["xyz"], requires_grad=True)
        color, radii, depth, alpha = GaussianRasterizer(rs)(means3D=P["xyz"], means2D=m2, opacities=torch.sigmoid(P["opacity"]), shs=shs,
                                                            scales=torch.exp(P["scaling"]), rotations=F.normalize(P["rotation"]))
        loss = 0.8 * (color - gt).abs().mean() + 0.2 * (1 - ssim(color, gt))
        d, m_ = depth.reshape(-1), pseudo.reshape(-1)
        loss = loss + 0.01 * min(1 - pearson(-m_, d), 1 - pearson(1000 / (m_ + 200.0), d))
        mask = (gt_depth > 0).float()
        loss = loss + 0.05 * (depth * mask - gt_depth * mask).abs().mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        # fused iteration
        loss_f, out = model.training_step(cam, gt, bg, args, it, pseudo_depth=pseudo, gt_depth=gt_depth)
        assert abs(float(loss_f) - float(loss)) <= 2e-5 * max(1.0, abs(float(loss))), (it, float(loss_f), float(loss))
        assert torch.equal(out["radii"], radii)
        ref = [P["xyz"], torch.cat([P["f_dc"], P["f_rest"]], dim=1), P["opacity"], P["scaling"], P["rotation"]]
        for name, a, b in zip(("xyz", "features", "opacity", "scaling", "rotation"), model._params(), ref):
            # Adam's first steps move every touched entry by ~lr regardless of gradient size, so compare against lr
            lr = {"xyz": sched(it), "features": args.feature_lr, "opacity": args.opacity_lr, "scaling": args.scaling_lr, "rotation": args.rotation_lr}[name]
            bad = ((a - b.detach()).abs() > 0.02 * lr).float().mean().item()
            assert bad < 2e-3, (it, name, bad)
        vis = radii > 0
        assert torch.equal(model.denom.squeeze(-1), vis.float() * 1 + (model.denom.squeeze(-1) - vis.float()))
    assert float(model.denom.max()) >= 1.0 and float(model.xyz_gradient_accum.max()) > 0


@pytest.mark.gpu
def test_training_loop_with_densification_and_opacity_reset():
    """A short schedule on the GPU that crosses a densification and an opacity reset: shapes stay consistent, the
    optimiser skips exactly the groups torch would skip, the loss keeps falling on a fixed view."""
    dev = "cuda:0"
    cfg = dict(P=20_000, W=160, H=120, deg=1, f=120.0, box=1.0, sigma0=0.05)
    rng = np.random.default_rng(0)
    pts = (rng.random((cfg["P"], 3)) - 0.5) * [6, 4, 6]
    cols = rng.random((cfg["P"], 3))
    model = gm.GaussianModel(cfg["deg"], device=dev)
    model.create_from_pcd(pts, cols, spatial_lr_scale=1.0)
    assert model.get_scaling.shape == (cfg["P"], 3) and float(model.get_opacity.mean()) == pytest.approx(0.1, abs=1e-6)
    args = gm.default_training_args(densify_from_iter=3, densification_interval=4, opacity_reset_interval=6, densify_until_iter=12,
                                    densify_grad_threshold=1e-7)
    model.training_setup(args)
    bg = torch.zeros(3, device=dev)
    cam = syn.make_camera(cfg, 0)
    gt = torch.rand(3, cam.H, cam.W, generator=torch.Generator().manual_seed(1)).to(dev)
    sizes, losses = [], []
    for it in range(1, 15):
        loss, out = model.training_step(cam, gt, bg, args, it, extent=3.0)
        n = model.get_xyz.shape[0]
        sizes.append(n)
        losses.append(float(loss))
        assert out["radii"].shape[0] <= n or out["densify"] is not None
        for t in model._params() + model._state["m"] + model._state["v"]:
            assert t.shape[0] == n
        assert model.xyz_gradient_accum.shape == (n, 1) and model.denom.shape == (n, 1) and model.max_radii2D.shape == (n,)
        assert model.get_opacity.shape == (n, 1) and model.get_rotation.shape == (n, 4)
        if it in (4, 8):                               # densification iterations: no optimiser step at all
            assert out["densify"] is not None and out["densify"]["after"] == n
            assert float(model.xyz_gradient_accum.abs().sum()) == 0.0
        if it == 6:                                    # opacity reset: that group skips one step
            assert float(model.get_opacity.max()) <= 0.01 + 1e-6 or model._steps["opacity"] < model._steps["xyz"]
    assert len(set(sizes)) > 1                          # the map changed size
    assert model._steps["opacity"] < model._steps["xyz"]
    assert np.isfinite(losses).all() and losses[4] < losses[0]    # falling until the first opacity reset (iteration 6)
