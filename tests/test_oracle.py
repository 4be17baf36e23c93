"""CPU tests of the oracle (oracle/gsr_oracle.cpp), the checker every GPU parity test relies on.

Pinning: the reference ships no golden vectors, so tests/golden/ref_*.npz hold OUTPUTS OF THE
UNMODIFIED REFERENCE CUDA rasterizer (oracle/_ref, generated on a B200 by
tests/golden/make_golden.py).  The float32 oracle must reproduce them: bit-exactly for radii,
tile counts, sorted keys, point list and tile ranges; within tolerance for images and gradients
(CPU exp vs MUFU.EX2, float-atomic ordering in the reference).  The float64 oracle is pinned
independently against torch.autograd on a dense differentiable restatement (tests/torch_ref.py).
"""
import os

import numpy as np
import pytest
import torch

import torch_ref
import util
from golden.make_golden import BG, GOLDEN_SCENES, loss_weights
from oracle.oracle import Oracle, get_higher_msb

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run(m, cam, bg, prec="f32", **kw):
    view, proj, raw, campos = cam.matrices()
    o = Oracle(prec)
    o.forward(bg, m.means3D, None, m.opacities, m.scales, m.rotations, 1.0, None, view, proj, cam.tanfovx, cam.tanfovy,
              cam.H, cam.W, m.shs, m.sh_degree, campos, **kw)
    return o


@pytest.mark.parametrize("name", list(GOLDEN_SCENES))
def test_oracle_matches_reference_golden(name):
    ref = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    m, cam = util.scene(**GOLDEN_SCENES[name])
    o = run(m, cam, torch.tensor(BG))
    g, b = o.geometry(), o.binning()
    vis = g["radii"] > 0
    # integer / index work: bit-exact against the reference build
    assert o.R == int(ref["num_rendered"])
    assert np.array_equal(g["radii"], ref["radii"])
    assert np.array_equal(g["tiles_touched"], ref["tiles_touched"].view(np.uint32))
    assert np.array_equal(b["keys"], ref["keys"].view(np.uint64))
    assert np.array_equal(b["list"], ref["point_list"].view(np.uint32))
    assert np.array_equal(b["ranges"], ref["ranges"].view(np.uint32))
    for k_o, k_r in (("depths", "vis_depths"), ("means2D", "vis_means2D"), ("conic_opacity", "vis_conic_opacity"),
                     ("cov3D", "vis_cov3D")):
        assert np.array_equal(g[k_o][vis].view(np.uint32), ref[k_r].view(np.uint32)), k_o
    np.testing.assert_allclose(g["rgb"][vis], ref["vis_rgb"], atol=2e-6)
    # n_contrib / images: a (pixel, splat) pair exactly on a threshold may flip between CPU exp and MUFU.EX2
    assert (b["n_contrib"] != ref["n_contrib"].view(np.uint32)).mean() <= 2e-3
    c, d, a = o.images()
    for got, want in ((c, ref["color"]), (d, ref["depth"]), (a, ref["alpha"])):
        err = np.abs(got - want) / max(1.0, float(np.abs(want).max()))
        assert (err > 1e-4).mean() <= 1e-3 and err.max() <= 2 / 255
    # gradients (reference accumulates with float atomics): 1e-3 relative
    wc, wd, wa = loss_weights(cam.H, cam.W)
    gr = o.backward(wc, wd, wa)
    for k in ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"):
        assert util.rel_err(gr[k][vis], ref[k]) <= 1e-3, k


def test_oracle_f64_matches_torch_autograd():
    """Independent pin of the float64 oracle: dense differentiable torch model, autograd gradients,
    including the pose gradient dL/d(rho, theta) derived two different ways."""
    cfg = dict(P=400, W=48, H=32, deg=3, f=40.0, box=1.0, sigma0=0.25)
    m, cam = util.scene(**{k: cfg[k] for k in ("P", "W", "H", "deg", "f", "sigma0")}, seed=3, query=1)
    H, W = cam.H, cam.W
    rng = np.random.default_rng(1)
    wc = rng.standard_normal((3, H, W)).astype(np.float32)
    wd = (0.3 * rng.standard_normal((1, H, W))).astype(np.float32)
    bg = torch.tensor([0.1, 0.2, 0.3])
    o = run(m, cam, bg, "f64")
    c, d, a = o.images()
    g = o.backward(wc, wd, None)
    view, proj, raw, campos = cam.matrices()
    leaf = lambda t: t.double().clone().requires_grad_(True)
    means, shs, opa, sc, rot = map(leaf, (m.means3D, m.shs, m.opacities, m.scales, m.rotations))
    tau = torch.zeros(6, dtype=torch.float64, requires_grad=True)
    w2c = view.double().t()
    for feeds in (False, True):
        col, dep, alp, aux = torch_ref.render(means, shs, opa, sc, rot, w2c, raw.double(), W, H, cam.tanfovx, cam.tanfovy,
                                              bg.double(), 3, tau=tau, depth_feeds_geometry=feeds)
        L = (col * torch.from_numpy(wc).double()).sum() + (dep * torch.from_numpy(wd).double()).sum()
        grads = torch.autograd.grad(L, [means, shs, opa, sc, rot, tau])
        if not feeds:
            assert np.abs(col.detach().numpy() - c).max() < 1e-5
            assert np.abs(dep.detach().numpy() - d).max() < 1e-5
            assert np.abs(alp.detach().numpy() - a).max() < 1e-5
            for name, gt in zip(("dL_dmeans3D", "dL_dsh", "dL_dopacity", "dL_dscales", "dL_drotations"), grads[:5]):
                assert util.rel_err(g[name], gt.numpy()) < 1e-5, name
        else:
            assert util.rel_err(g["dL_dtau"], grads[5].numpy()) < 1e-5


def test_oracle_invariants():
    m, cam = util.scene(P=5000, W=112, H=80, deg=1, f=90.0, sigma0=0.1)
    o = run(m, cam, torch.zeros(3), count_touched=True)
    g, b = o.geometry(), o.binning()
    assert o.R == int(g["tiles_touched"].sum()) == int(g["point_offsets"][-1])
    keys = b["keys"]
    assert np.all(np.diff(keys.astype(np.uint64)) >= 0) if o.R > 1 else True          # sortedness
    assert sorted(b["keys_unsorted"].tolist()) == keys.tolist()                       # permutation
    # stability: equal keys keep Gaussian order
    same = keys[1:] == keys[:-1]
    assert np.all(b["list"][1:][same] > b["list"][:-1][same])
    # ranges partition the list by tile
    T = b["ranges"].shape[0]
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    for t in range(T):
        s, e = b["ranges"][t]
        assert np.all(tiles[s:e] == t)
    assert int((b["ranges"][:, 1].astype(np.int64) - b["ranges"][:, 0]).sum()) == o.R
    # n_contrib never exceeds the tile's list length; n_touched only for visible Gaussians
    H, W = cam.H, cam.W
    lens = (b["ranges"][:, 1].astype(np.int64) - b["ranges"][:, 0]).reshape((H + 15) // 16, (W + 15) // 16)
    assert np.all(b["n_contrib"] <= np.kron(lens, np.ones((16, 16), np.int64))[:H, :W])
    assert np.all(b["n_touched"][g["radii"] == 0] == 0) and b["n_touched"].sum() > 0


def test_oracle_f32_f64_agree():
    m, cam = util.scene(P=3000, W=96, H=64, deg=2, f=80.0, sigma0=0.12)
    c32, d32, a32 = run(m, cam, torch.zeros(3), "f32").images()
    c64, d64, a64 = run(m, cam, torch.zeros(3), "f64").images()
    err = np.abs(c32 - c64)
    assert (err > 1e-4).mean() < 1e-3 and err.max() < 2 / 255


def test_get_higher_msb():
    # rasterizer_impl.cu:35-50: number of bits needed for the tile index part of the sort key
    for n, want in ((1, 1), (2, 2), (80, 7), (1200, 11), (2304, 12), (4346, 13), (8160, 13), (65535, 16), (65536, 17)):
        assert get_higher_msb(n) == want, n


def test_empty_scene():
    m, cam = util.scene(P=16, W=32, H=16)
    view, proj, raw, campos = cam.matrices()
    o = Oracle("f32")
    R = o.forward(torch.tensor([0.5, 0.25, 0.125]), np.zeros((0, 3), np.float32), None, np.zeros((0, 1), np.float32),
                  np.zeros((0, 3), np.float32), np.zeros((0, 4), np.float32), 1.0, None, view, proj, cam.tanfovx, cam.tanfovy,
                  cam.H, cam.W, np.zeros((0, 16, 3), np.float32), 3, campos)
    assert R == 0
    c, d, a = o.images()
    assert np.allclose(c[0], 0.5) and np.allclose(c[2], 0.125) and not a.any() and not d.any()
