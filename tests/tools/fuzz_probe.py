"""For a fuzz seed: gradient error of ours and of the reference build against the float64 CPU oracle."""
import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import util
from test_gpu_fuzz import _case
from gs_localization_b200.diff_gaussian_rasterization import _C as ours
from oracle.oracle import Oracle
DEV = "cuda:0"
ref = util.load_reference()
for seed in [int(a) for a in sys.argv[1:]]:
    m, cam, smod = _case(seed)
    bg = torch.tensor([0.2, 0.0, 0.4])
    args = list(util.c_args(m, cam, bg, DEV)); args[6] = smod
    g = torch.Generator().manual_seed(seed)
    gC = (torch.rand(3, cam.H, cam.W, generator=g) - 0.5); gD = (torch.rand(1, cam.H, cam.W, generator=g) - 0.5) * 0.1; gA = (torch.rand(1, cam.H, cam.W, generator=g) - 0.5) * 0.1
    (bgt, means3D, col, opac, scales, rots, sm, cov, view, proj, tfx, tfy, Hh, Ww, sh, deg, campos, pf, dbg) = args
    outs = {}
    for name, mod in (("ours", ours), ("ref", ref._C)):
        res = []
        for rep in range(2):
            R, color, depth, alpha, radii, geom, binning, img = mod.rasterize_gaussians(*args)
            res.append(mod.rasterize_gaussians_backward(bgt, means3D, radii, col, scales, rots, sm, cov, view, proj, tfx, tfy, gC.to(DEV), gD.to(DEV), gA.to(DEV), sh, deg, campos, geom, R, binning, img, alpha, False))
        outs[name] = res
    v, p, raw, c = cam.matrices()
    o = Oracle("f64")
    o.forward(bg, m.means3D, None, m.opacities, m.scales, m.rotations, smod, None, v, p, cam.tanfovx, cam.tanfovy, cam.H, cam.W, m.shs, m.sh_degree, c)
    want = o.backward(gC.numpy(), gD.numpy(), gA.numpy())
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"]
    for k in (0, 2, 3, 5, 6, 7):
        w = want[names[k]]
        eo = util.rel_err(outs["ours"][0][k].cpu().numpy().reshape(w.shape), w)
        er = util.rel_err(outs["ref"][0][k].cpu().numpy().reshape(w.shape), w)
        rr = util.rel_err(outs["ref"][0][k].cpu().numpy(), outs["ref"][1][k].cpu().numpy())
        oo = util.rel_err(outs["ours"][0][k].cpu().numpy(), outs["ours"][1][k].cpu().numpy())
        print(seed, names[k], "ours-vs-f64 %.2e  ref-vs-f64 %.2e  ref-vs-ref %.2e  ours-vs-ours %.2e  |want| %.3e" % (eo, er, rr, oo, np.linalg.norm(w)))
