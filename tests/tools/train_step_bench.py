"""One map-training iteration (render -> L1+SSIM loss -> backward -> Adam on all six groups + densification
statistics) at a BASELINE config, two ways on the same B200:

  fused      GaussianModel.training_step: C-ABI rasterizer + fused loss kernels + gsr_map_adam_step, no autograd
  framework  the reference's formulation of the same iteration: autograd through the activations and torch.cat of
             the SH features, conv2d SSIM, torch.optim.Adam, run on (a) this repo's drop-in rasterizer and
             (b) the reference rasterizer build (oracle/_ref) when present

  python tools/train_step_bench.py [C4] [iters]"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gs_localization_b200 import gaussian_model as gm, io as gio, synthetic as syn  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = "cuda:0"
cfg = syn.CONFIGS[name]
gmap = syn.make_map(cfg["P"], cfg["deg"], cfg["sigma0"], cfg["box"], seed=0)
raw = gio.deactivate(gmap)
args = gm.default_training_args()
bg = torch.zeros(3, device=dev)
cams = [syn.make_camera(cfg, i) for i in range(8)]
gen = torch.Generator().manual_seed(0)
gts = [torch.rand(3, cams[0].H, cams[0].W, generator=gen).to(dev) for _ in range(8)]

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


def timed(step):
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for i in range(iters):
        step(3 + i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def fused():
    model = gm.GaussianModel(cfg["deg"], device=dev)
    model.from_raw(raw)
    model.spatial_lr_scale = 1.0
    model.training_setup(args)
    return timed(lambda i: model.training_step(cams[i % 8], gts[i % 8], bg, args, 1 + i))


def framework(pkg):
    t = lambda a: torch.nn.Parameter(torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(dev))
    P = dict(xyz=t(raw.xyz), f_dc=t(raw.features_dc), f_rest=t(raw.features_rest), opacity=t(raw.opacity), scaling=t(raw.scaling),
             rotation=t(raw.rotation))
    lrs = dict(xyz=args.position_lr_init, f_dc=args.feature_lr, f_rest=args.feature_lr / 20, opacity=args.opacity_lr, scaling=args.scaling_lr,
               rotation=args.rotation_lr)
    opt = torch.optim.Adam([{"params": [P[k]], "lr": lrs[k]} for k in gm.GROUPS], lr=0.0, eps=1e-15)
    accum, denom, maxr = torch.zeros(cfg["P"], 1, device=dev), torch.zeros(cfg["P"], 1, device=dev), torch.zeros(cfg["P"], device=dev)

    def step(i):
        cam = cams[i % 8]
        view, proj, _, campos = cam.matrices(dev)
        rs = pkg.GaussianRasterizationSettings(cam.H, cam.W, cam.tanfovx, cam.tanfovy, bg, 1.0, view, proj, cfg["deg"], campos, False, False)
        m2 = torch.zeros_like(P["xyz"], requires_grad=True)
        color, radii, depth, alpha = pkg.GaussianRasterizer(rs)(
            means3D=P["xyz"], means2D=m2, opacities=torch.sigmoid(P["opacity"]), shs=torch.cat([P["f_dc"], P["f_rest"]], dim=1),
            scales=torch.exp(P["scaling"]), rotations=F.normalize(P["rotation"]))
        gt = gts[i % 8]
        loss = 0.8 * (color - gt).abs().mean() + 0.2 * (1 - ssim(color, gt))
        loss.backward()
        with torch.no_grad():
            vis = radii > 0
            maxr[vis] = torch.max(maxr[vis], radii[vis])
            accum[vis] += torch.norm(m2.grad[vis, :2], dim=-1, keepdim=True)
            denom[vis] += 1
            opt.step()
            opt.zero_grad(set_to_none=True)
    return timed(step)


row = {"tool": "train_step_bench", "config": name, "P": cfg["P"], "image": [cams[0].W, cams[0].H], "iters": iters}
row["fused_ms"] = fused()
torch.cuda.empty_cache()
import gs_localization_b200.diff_gaussian_rasterization as ours  # noqa: E402
row["framework_on_b200_rasterizer_ms"] = framework(ours)
torch.cuda.empty_cache()
ref_dir = os.path.join(ROOT, "oracle", "_ref")
if os.path.exists(os.path.join(ref_dir, "diff_gaussian_rasterization", "_C.so")):
    sys.path.insert(0, ref_dir)
    for k in [k for k in sys.modules if k.startswith("diff_gaussian_rasterization")]:
        del sys.modules[k]
    import diff_gaussian_rasterization as refpkg  # noqa: E402
    row["reference_ms"] = framework(refpkg)
    row["speedup_vs_reference"] = row["reference_ms"] / row["fused_ms"]
# optimiser kernel alone against its HBM floor: 7 passes over (11 + 3M) floats per Gaussian
model = gm.GaussianModel(cfg["deg"], device=dev)
model.from_raw(raw)
model.spatial_lr_scale = 1.0
model.training_setup(args)
G = [torch.randn_like(p) * 1e-3 for p in model._params()]
t = timed(lambda i: model.optimizer_step(G, adam=True))
M = (cfg["deg"] + 1) ** 2
bytes_ = 7 * 4 * (11 + 3 * M) * cfg["P"] + 4 * 8 * cfg["P"]
row["adam_kernel_ms"] = t
row["adam_GBs"] = bytes_ / t / 1e6
print(json.dumps(row))
