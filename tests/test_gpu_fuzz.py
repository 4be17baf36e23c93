"""Randomised bit-exactness sweep against the reference build: unnormalised quaternions, extreme scales, scale
modifiers, cameras inside the cloud, odd image sizes.  Guards the conservative screen-radius bound of the preprocess
cull (a wrongly culled Gaussian shows up as a radius / list mismatch) and the tile-list machinery."""
import numpy as np
import pytest
import torch

import util
from gs_localization_b200 import synthetic as syn
from gs_localization_b200.diff_gaussian_rasterization import _C as ours

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _case(seed):
    rng = np.random.default_rng(seed)
    P = int(rng.integers(500, 8000))
    W, H = int(rng.integers(17, 400)), int(rng.integers(17, 300))
    deg = int(rng.integers(0, 4))
    f = float(rng.uniform(20.0, 600.0))
    cfg = dict(P=P, W=W, H=H, deg=deg, f=f, box=float(rng.uniform(0.3, 3.0)), sigma0=float(10 ** rng.uniform(-2.5, -0.3)))
    m = syn.make_map(P, deg, cfg["sigma0"], cfg["box"], seed=seed)
    g = torch.Generator().manual_seed(seed)
    # unnormalised quaternions and a heavy tail of scales (the reference accepts both)
    rot = m.rotations * torch.exp(torch.randn(P, 1, generator=g) * 0.6)
    scales = m.scales * torch.exp(torch.randn(P, 3, generator=g) * 1.0)
    m = m._replace(rotations=rot, scales=scales)
    cam = syn.make_camera(cfg, int(rng.integers(0, 1000)))
    return m, cam, float(rng.choice([0.5, 1.0, 1.0, 2.5]))


@pytest.mark.parametrize("seed", range(24))
def test_random_scene_bit_exact_vs_reference(seed):
    if not util.reference_available():
        pytest.skip("oracle/_ref not built (reference sources absent at build time)")
    ref = util.load_reference()
    m, cam, scale_modifier = _case(seed)
    bg = torch.tensor([0.2, 0.0, 0.4])
    args = list(util.c_args(m, cam, bg, DEV))
    args[6] = scale_modifier
    R, color, depth, alpha, radii, geom, binning, img = ours.rasterize_gaussians(*args)
    rR, rcolor, rdepth, ralpha, rradii, rgeom, rbin, rimg = ref._C.rasterize_gaussians(*args)
    torch.cuda.synchronize()
    P = m.means3D.shape[0]
    assert R == rR and torch.equal(radii, rradii), (seed, R, rR, int((radii != rradii).sum()))
    st = ours.export_state(P, R, cam.W, cam.H, geom, binning, img)
    rs = util.ref_unpack_state(P, rR, cam.W, cam.H, rgeom, rbin, rimg)
    for k in ("keys", "list", "ranges", "n_contrib"):
        assert torch.equal(st[k], rs[k]), (seed, k)
    assert float((alpha - ralpha).abs().max()) == 0.0
    assert float((color - rcolor).abs().max()) <= 1e-4 and float((depth - rdepth).abs().max()) <= 1e-4 * max(1.0, float(rdepth.abs().max()))
    # backward against the reference on the same upstream gradient
    g = torch.Generator().manual_seed(seed)
    gC = (torch.rand(3, cam.H, cam.W, generator=g) - 0.5).to(DEV)
    gD = (torch.rand(1, cam.H, cam.W, generator=g) - 0.5).to(DEV) * 0.1
    gA = (torch.rand(1, cam.H, cam.W, generator=g) - 0.5).to(DEV) * 0.1
    (bgt, means3D, col, opac, scales, rots, smod, cov, view, proj, tfx, tfy, Hh, Ww, sh, deg, campos, pf, dbg) = args
    mine = ours.rasterize_gaussians_backward(bgt, means3D, radii, col, scales, rots, smod, cov, view, proj, tfx, tfy, gC, gD, gA, sh, deg,
                                             campos, geom, R, binning, img, alpha, False)
    theirs = ref._C.rasterize_gaussians_backward(bgt, means3D, rradii, col, scales, rots, smod, cov, view, proj, tfx, tfy, gC, gD, gA, sh,
                                                 deg, campos, rgeom, rR, rbin, rimg, ralpha, False)
    # Heavy-tailed scales make some of these gradients ill-conditioned in float32: the reference's own run-to-run
    # spread (float atomics) reaches tens of percent on the worst seeds, where no float32 answer is meaningful.  A
    # tensor that goes through the covariance chain is compared only where the reference agrees with itself to 5e-5; there this implementation must agree
    # with it to the north-star 1e-3 (x2 for the two atomics-ordered sums being compared).
    theirs2 = ref._C.rasterize_gaussians_backward(bgt, means3D, rradii, col, scales, rots, smod, cov, view, proj, tfx, tfy, gC, gD, gA, sh,
                                                  deg, campos, rgeom, rR, rbin, rimg, ralpha, False)
    compared = 0
    for k, (a, b, b2) in enumerate(zip(mine, theirs, theirs2)):
        if k in (1, 4) or b.numel() == 0:       # colours / cov3D are not inputs on this path
            continue
        bn, b2n = b.double().cpu().numpy(), b2.double().cpu().numpy()
        if k in (3, 6, 7) and util.rel_err(b2n, bn) > 5e-5:     # means3D / scales / rotations: only where well conditioned
            continue
        compared += 1
        assert util.rel_err(a.double().cpu().numpy(), bn) <= 2e-3, (seed, k)
    assert compared >= 3, (seed, compared)      # means2D, opacity and SH are always well conditioned
