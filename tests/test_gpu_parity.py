"""GPU parity tests: the CUDA path (through the C ABI) against
  (A) the unmodified reference CUDA rasterizer built for sm_100a (oracle/_ref), bit-exact for
      radii / tiles / sort keys / point list / tile ranges / n_contrib;
  (B) the CPU oracle (oracle/gsr_oracle.cpp): bit-exact for the integer/binning work, images
      <= 1e-4 max abs (north_star), gradients <= 1e-3 relative vs the float64 oracle.
"""
import numpy as np
import pytest
import torch

import util
from gs_localization_b200.diff_gaussian_rasterization import _C as ours
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

IMG_TOL = 1e-4        # north_star: images agree to <= 1e-4 max abs
GRAD_REL_TOL = 1e-3   # north_star: parameter and pose gradients agree to <= 1e-3 relative

SCENES = {
    "tiny": dict(P=500, W=48, H=32, deg=3, f=40.0, sigma0=0.2),
    "C1": dict(P=10_000, W=160, H=120, deg=0, f=131.25, sigma0=0.05),
    "ragged": dict(P=3000, W=75, H=53, deg=2, f=70.0, sigma0=0.1),      # W,H not multiples of 16
    "deg1": dict(P=4000, W=128, H=96, deg=1, f=100.0, sigma0=0.08),
    "mid": dict(P=60_000, W=320, H=240, deg=3, f=262.5, sigma0=0.04),
    # tile lists longer than the shared-memory tile sort handles (> 4096): exercises the global onesweep path
    "dense": dict(P=200_000, W=64, H=48, deg=1, f=50.0, sigma0=0.1),
    # more than 10240 cells in the coverage grid: scan_tiles works in global memory instead of shared memory
    "wide": dict(P=4000, W=2064, H=1552, deg=0, f=1500.0, sigma0=0.05),
}


def _images_close(color, depth, alpha, oc, od, oa):
    """CUDA vs CPU oracle.  The CPU's exp differs from MUFU.EX2 by an ulp or two, so a
    (pixel, splat) pair sitting exactly on the alpha >= 1/255 threshold can flip; such a flip
    moves that pixel by at most ~1/255.  Hence: all but a handful of pixels within the
    north_star 1e-4, and no pixel off by more than one threshold flip.  (Against the
    reference build on the same GPU the comparison is exact, see *_vs_reference_*.)"""
    dscale = max(1.0, float(np.abs(od).max()))
    for got, want, scale in ((color, oc, 1.0), (alpha, oa, 1.0), (depth, od, dscale)):
        err = np.abs(got.cpu().numpy() - want) / scale
        assert (err > IMG_TOL).mean() <= 1e-3, (err > IMG_TOL).mean()
        assert err.max() <= 2.0 / 255.0, err.max()


def run_ours(m, cam, bg, **kw):
    args = util.c_args(m, cam, bg, DEV, **kw)
    out = ours.rasterize_gaussians(*args)
    torch.cuda.synchronize()
    return args, out


def run_oracle(m, cam, bg, prec="f32", colors_precomp=None, cov3D_precomp=None, count_touched=False):
    view, proj, raw, campos = cam.matrices()
    o = Oracle(prec)
    o.forward(bg, m.means3D, colors_precomp, m.opacities, None if cov3D_precomp is not None else m.scales,
              None if cov3D_precomp is not None else m.rotations, 1.0, cov3D_precomp, view, proj, cam.tanfovx,
              cam.tanfovy, cam.H, cam.W, None if colors_precomp is not None else m.shs, m.sh_degree, campos,
              count_touched=count_touched)
    return o


@pytest.mark.parametrize("name", list(SCENES))
def test_forward_vs_oracle(name):
    m, cam = util.scene(**SCENES[name])
    bg = torch.tensor([0.1, 0.3, 0.2])
    args, (R, color, depth, alpha, radii, geom, binning, img) = run_ours(m, cam, bg)
    o = run_oracle(m, cam, bg)
    P = m.means3D.shape[0]
    assert R == o.R
    st = ours.export_state(P, R, cam.W, cam.H, geom, binning, img)
    g, b = o.geometry(), o.binning()
    if name == "dense":
        lens = b["ranges"][:, 1].astype(np.int64) - b["ranges"][:, 0]
        assert lens.max() > 4096, lens.max()      # really on the global-sort path
    # integer / index work: bit-exact
    assert np.array_equal(radii.cpu().numpy(), g["radii"])
    assert np.array_equal(st["tiles_touched"].cpu().numpy().view(np.uint32), g["tiles_touched"])
    vis = g["radii"] > 0
    assert np.array_equal(st["depths"].cpu().numpy()[vis].view(np.uint32), g["depths"][vis].view(np.uint32))
    assert np.array_equal(st["means2D"].cpu().numpy()[vis].view(np.uint32), g["means2D"][vis].view(np.uint32))
    assert np.array_equal(st["conic_opacity"].cpu().numpy()[vis].view(np.uint32), g["conic_opacity"][vis].view(np.uint32))
    assert np.array_equal(st["cov3D"].cpu().numpy()[vis].view(np.uint32), g["cov3D"][vis].view(np.uint32))
    assert np.array_equal(st["keys"].cpu().numpy().view(np.uint64), b["keys"])
    assert np.array_equal(st["list"].cpu().numpy().view(np.uint32), b["list"])
    assert np.array_equal(st["ranges"].cpu().numpy().view(np.uint32), b["ranges"])
    np.testing.assert_allclose(st["rgb"].cpu().numpy()[vis], g["rgb"][vis], atol=2e-6)
    # floating point: images within the north_star tolerance; n_contrib may flip where CPU expf
    # and MUFU.EX2 differ by an ulp at a threshold (bit-exactness is checked against the reference build)
    oc, od, oa = o.images()
    _images_close(color, depth, alpha, oc, od, oa)
    mism = (st["n_contrib"].cpu().numpy().view(np.uint32) != b["n_contrib"]).mean()
    assert mism <= 2e-3, mism


@pytest.mark.parametrize("name", ["tiny", "C1", "ragged", "mid", "dense", "wide"])
def test_forward_vs_reference_bit_exact(name):
    if not util.reference_available():
        pytest.skip("oracle/_ref not built (reference sources absent at build time)")
    ref = util.load_reference()
    m, cam = util.scene(**SCENES[name])
    bg = torch.tensor([0.1, 0.3, 0.2])
    args, (R, color, depth, alpha, radii, geom, binning, img) = run_ours(m, cam, bg)
    rR, rcolor, rdepth, ralpha, rradii, rgeom, rbin, rimg = ref._C.rasterize_gaussians(*args)
    torch.cuda.synchronize()
    P = m.means3D.shape[0]
    assert R == rR
    st = ours.export_state(P, R, cam.W, cam.H, geom, binning, img)
    rs = util.ref_unpack_state(P, rR, cam.W, cam.H, rgeom, rbin, rimg)
    assert torch.equal(radii, rradii)
    vis = rradii > 0
    assert torch.equal(st["tiles_touched"], rs["tiles_touched"])
    for k in ("depths", "means2D", "conic_opacity", "cov3D"):
        assert torch.equal(st[k][vis].view(torch.int32), rs[k][vis].view(torch.int32)), k
    for k in ("keys", "list", "ranges", "n_contrib"):
        assert torch.equal(st[k], rs[k]), k
    # with identical lists and identical per-pair arithmetic the images are bit-identical too
    assert (color - rcolor).abs().max().item() <= IMG_TOL
    assert (alpha - ralpha).abs().max().item() == 0.0
    assert (depth - rdepth).abs().max().item() <= IMG_TOL
    assert torch.equal(st["clamped"][vis], rs["clamped"][vis].to(torch.uint8))


@pytest.mark.parametrize("kind", ["pairs", "runs", "plane"])
def test_equal_depths_are_ordered_like_the_reference(kind):
    """The tile sort runs on 32-bit words (quantised depth | index in the bucket) and resolves runs of equal quantised
    depth afterwards (binning.cu, tile_sort_small).  Exact depth ties are where that can go wrong: the reference's stable
    radix sort orders them by Gaussian id.  pairs: every Gaussian twice (runs of 2, ranked in place); runs: groups of
    3-12 coincident Gaussians; plane: a fronto-parallel plane seen by an unrotated camera — every instance of a tile has
    the SAME depth, hundreds per tile (the long-run fall-back to the 64-bit network)."""
    if not util.reference_available():
        pytest.skip("oracle/_ref not built (reference sources absent at build time)")
    ref = util.load_reference()
    from gs_localization_b200 import synthetic as syn
    g = torch.Generator().manual_seed(11)
    if kind == "plane":
        P, W, H, f = 6000, 128, 96, 90.0
        m = syn.make_map(P, 1, 0.03, 1.0, seed=5)
        means = torch.stack([(torch.rand(P, generator=g) - 0.5) * 3.0, (torch.rand(P, generator=g) - 0.5) * 2.2,
                             torch.full((P,), 2.0)], 1)
        m = m._replace(means3D=means.float().contiguous())
        w2c = torch.eye(4, dtype=torch.float64)
        w2c[2, 3] = 0.25
        cam = syn.Camera(w2c, W, H, f, f, W / 2.0, H / 2.0)
    else:
        m, cam = util.scene(P=24_000, W=256, H=192, deg=1, f=200.0, sigma0=0.05)
        P = m.means3D.shape[0]
        means = m.means3D.clone()
        if kind == "pairs":
            means[P // 2:] = means[: P // 2]
        else:
            src = torch.randint(0, P // 16, (P,), generator=g)      # ~16 Gaussians per distinct position
            keep = torch.rand(P, generator=g) < 0.5
            means = torch.where(keep[:, None], means, means[src])
        m = m._replace(means3D=means.contiguous())
    bg = torch.tensor([0.1, 0.3, 0.2])
    args, (R, color, depth, alpha, radii, geom, binning, img) = run_ours(m, cam, bg)
    rR, rcolor, rdepth, ralpha, rradii, rgeom, rbin, rimg = ref._C.rasterize_gaussians(*args)
    torch.cuda.synchronize()
    assert R == rR and R > 0
    st = ours.export_state(P, R, cam.W, cam.H, geom, binning, img)
    rs = util.ref_unpack_state(P, rR, cam.W, cam.H, rgeom, rbin, rimg)
    keys = rs["keys"].cpu().numpy().view(np.uint64)
    ties = int((keys[1:] == keys[:-1]).sum())
    assert ties > (R // 4 if kind != "runs" else R // 8), (ties, R)       # the scene really is full of equal keys
    for k in ("keys", "list", "ranges", "n_contrib"):
        assert torch.equal(st[k], rs[k]), k
    assert (alpha - ralpha).abs().max().item() == 0.0


def test_packed_expf_is_expf_bit_for_bit():
    """The forward's alpha decides n_contrib, and it comes from expf_pair (render.cu): CUDA's accurate expf with the four
    roundings that have a packed form issued once for two values.  Every bit must be expf()'s: swept on the device over the
    exponent's whole useful range (the blend only evaluates power <= 0), around zero, and at the special values."""
    import ctypes as C
    from gs_localization_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)
    n = 1 << 23
    x = torch.empty(n)
    x[: n // 2] = torch.rand(n // 2, generator=g) * -110.0                      # the blend's domain
    x[n // 2: 3 * n // 4] = (torch.rand(n // 4, generator=g) - 0.5) * 200.0     # both signs, overflow and underflow ends
    x[3 * n // 4:] = torch.randn(n // 4, generator=g) * torch.logspace(-30, 1, n // 4)
    x[:12] = torch.tensor([0.0, -0.0, float("inf"), float("-inf"), float("nan"), -87.33655, -88.0, -103.97, -104.0, 88.72, 88.73, -1e-45])
    xd = x.to(DEV)
    ref, fast = torch.empty_like(xd), torch.empty_like(xd)
    p = lambda t: C.c_void_p(t.data_ptr())
    lib.gsr_selftest_expf.restype = C.c_int
    rc = lib.gsr_selftest_expf(p(xd), p(ref), p(fast), C.c_int(n), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert rc == 0
    bad = ref.view(torch.int32) != fast.view(torch.int32)
    assert not bool(bad.any()), (int(bad.sum()), x[bad.cpu()][:8], ref[bad][:8], fast[bad][:8])
    assert torch.equal(ref[:2].cpu(), torch.ones(2)) and float(ref[3]) == 0.0


def _grads_ours(m, cam, bg, wc, wd, wa):
    args, (R, color, depth, alpha, radii, geom, binning, img) = run_ours(m, cam, bg)
    (bgt, means3D, colors, opac, scales, rots, smod, cov, view, proj, tfx, tfy, H, W, sh, deg, campos, pf, dbg) = args
    t = lambda a: torch.from_numpy(a).to(DEV)
    res = ours.rasterize_gaussians_backward(bgt, means3D, radii, colors, scales, rots, smod, cov, view, proj, tfx, tfy,
                                            t(wc), t(wd), t(wa), sh, deg, campos, geom, R, binning, img, alpha, False)
    torch.cuda.synchronize()
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"]
    return dict(zip(names, [r.cpu().numpy() for r in res])), args, (R, radii, geom, binning, img, alpha)


@pytest.mark.parametrize("name", ["tiny", "C1", "ragged", "mid", "dense"])
def test_backward_vs_oracle_f64(name):
    m, cam = util.scene(**SCENES[name])
    bg = torch.tensor([0.1, 0.3, 0.2])
    rng = np.random.default_rng(0)
    H, W = cam.H, cam.W
    wc = rng.standard_normal((3, H, W)).astype(np.float32)
    wd = (0.3 * rng.standard_normal((1, H, W))).astype(np.float32)
    wa = (0.2 * rng.standard_normal((1, H, W))).astype(np.float32)
    got, _, _ = _grads_ours(m, cam, bg, wc, wd, wa)
    o = run_oracle(m, cam, bg, "f64")
    want = o.backward(wc, wd, wa)
    for k in got:
        if k == "dL_dcov3D":
            continue   # intermediate of the scale/rotation path; checked through scales/rotations
        e = util.rel_err(got[k], want[k])
        assert e <= GRAD_REL_TOL, (k, e)


def test_backward_vs_reference():
    if not util.reference_available():
        pytest.skip("oracle/_ref not built")
    ref = util.load_reference()
    m, cam = util.scene(**SCENES["mid"])
    bg = torch.tensor([0.1, 0.3, 0.2])
    rng = np.random.default_rng(0)
    H, W = cam.H, cam.W
    wc = rng.standard_normal((3, H, W)).astype(np.float32)
    wd = (0.3 * rng.standard_normal((1, H, W))).astype(np.float32)
    wa = (0.2 * rng.standard_normal((1, H, W))).astype(np.float32)
    got, args, _ = _grads_ours(m, cam, bg, wc, wd, wa)
    (bgt, means3D, colors, opac, scales, rots, smod, cov, view, proj, tfx, tfy, H, W, sh, deg, campos, pf, dbg) = args
    rR, rcolor, rdepth, ralpha, rradii, rgeom, rbin, rimg = ref._C.rasterize_gaussians(*args)
    t = lambda a: torch.from_numpy(a).to(DEV)
    rres = ref._C.rasterize_gaussians_backward(bgt, means3D, rradii, colors, scales, rots, smod, cov, view, proj, tfx, tfy,
                                               t(wc), t(wd), t(wa), sh, deg, campos, rgeom, rR, rbin, rimg, ralpha, False)
    torch.cuda.synchronize()
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"]
    for k, r in zip(names, rres):
        e = util.rel_err(got[k], r.cpu().numpy())
        assert e <= GRAD_REL_TOL, (k, e)


def test_precomputed_colors_and_cov():
    m, cam = util.scene(**SCENES["deg1"])
    bg = torch.tensor([0.0, 0.0, 0.0])
    o0 = run_oracle(m, cam, bg)
    g0 = o0.geometry()
    colors = torch.rand(m.means3D.shape[0], 3, generator=torch.Generator().manual_seed(1))
    cov = torch.from_numpy(g0["cov3D"].copy())
    # culled Gaussians have no cov3D in the oracle state: give them something harmless
    cov[torch.from_numpy(g0["radii"] <= 0)] = torch.tensor([1e-4, 0, 0, 1e-4, 0, 1e-4])
    args, (R, color, depth, alpha, radii, geom, binning, img) = run_ours(m, cam, bg, colors_precomp=colors, cov3D_precomp=cov)
    o = run_oracle(m, cam, bg, colors_precomp=colors, cov3D_precomp=cov)
    assert R == o.R
    oc, od, oa = o.images()
    _images_close(color, depth, alpha, oc, od, oa)
    assert np.array_equal(radii.cpu().numpy(), o.geometry()["radii"])


def test_empty_and_all_culled():
    bg = torch.tensor([0.2, 0.4, 0.6])
    m, cam = util.scene(P=64, W=40, H=24)
    # (a) zero Gaussians: reference returns zero images (rasterize_points.cu:83)
    empty = m._replace(means3D=m.means3D[:0], shs=m.shs[:0], opacities=m.opacities[:0], scales=m.scales[:0], rotations=m.rotations[:0])
    args = util.c_args(empty, cam, bg, DEV)
    R, color, depth, alpha, radii, *_ = ours.rasterize_gaussians(*args)
    assert R == 0 and color.abs().max().item() == 0 and radii.numel() == 0
    # (b) everything behind the camera: background only
    view = cam.matrices()[0]
    behind = m._replace(means3D=m.means3D * 0 + (torch.linalg.inv(view.double().t())[:3, :3] @ torch.tensor([0.0, 0.0, -5.0], dtype=torch.float64)).float() + torch.linalg.inv(view.double().t())[:3, 3].float())
    args = util.c_args(behind, cam, bg, DEV)
    R, color, depth, alpha, radii, geom, binning, img = ours.rasterize_gaussians(*args)
    torch.cuda.synchronize()
    assert R == 0 and int(radii.max()) == 0
    assert torch.allclose(color, bg.to(DEV)[:, None, None].expand_as(color))
    assert alpha.abs().max().item() == 0 and depth.abs().max().item() == 0
    # and its backward is all zeros
    (bgt, means3D, colors, opac, scales, rots, smod, cov, view_, proj, tfx, tfy, H, W, sh, deg, campos, pf, dbg) = args
    res = ours.rasterize_gaussians_backward(bgt, means3D, radii, colors, scales, rots, smod, cov, view_, proj, tfx, tfy,
                                            torch.ones_like(color), torch.ones_like(depth), torch.ones_like(alpha), sh, deg,
                                            campos, geom, R, binning, img, alpha, False)
    assert all(float(r.abs().max()) == 0 for r in res if r.numel())


def test_mark_visible():
    m, cam = util.scene(**SCENES["C1"])
    view, proj, raw, campos = cam.matrices(DEV)
    got = ours.mark_visible(m.means3D.to(DEV), view, proj)
    assert got.dtype == torch.bool and got.shape == (m.means3D.shape[0],)
    o = run_oracle(m, cam, torch.zeros(3))
    # every Gaussian the oracle kept is marked
    assert got.cpu().numpy()[o.geometry()["radii"] > 0].all()
    if util.reference_available():
        # exact: the reference's own checkFrustum kernel (rasterizer_impl.cu:54-66) on the same inputs, on the ragged
        # 1M-point headline map too (P not a multiple of the block size)
        ref = util.load_reference()
        assert torch.equal(got, ref._C.mark_visible(m.means3D.to(DEV), view, proj))
        from gs_localization_b200 import synthetic as syn
        cfg = syn.CONFIGS["headline"]
        big = syn.make_map(cfg["P"] + 77, 0, cfg["sigma0"], cfg["box"], seed=3).means3D.to(DEV)
        for q in range(3):
            v, p_, _, _ = syn.make_camera(cfg, q).matrices(DEV)
            assert torch.equal(ours.mark_visible(big, v, p_), ref._C.mark_visible(big, v, p_)), q
    else:
        pv = (m.means3D.double() @ view.cpu().double()[:3, :3] + view.cpu().double()[3, :3]).numpy()
        assert ((pv[:, 2] > 0.2) == got.cpu().numpy()).mean() > 0.999


def test_sort_pairs_standalone():
    from gs_localization_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    for n, end_bit in ((1, 40), (1000, 43), (4096, 45), (4097, 33), (300_000, 43), (1_000_003, 45)):
        keys = torch.randint(0, 2 ** 62, (n,), generator=g, dtype=torch.int64) & ((1 << end_bit) - 1)
        keys[: n // 3] = keys[n // 3: 2 * (n // 3)][: n // 3]     # plenty of duplicates: stability matters
        vals = torch.arange(n, dtype=torch.int32)
        kd, vd = keys.to(DEV), vals.to(DEV)
        ko, vo, kt, vt = torch.empty_like(kd), torch.empty_like(vd), torch.empty_like(kd), torch.empty_like(vd)
        temp = torch.empty(lib.gsr_sort_temp_bytes(n), dtype=torch.uint8, device=DEV)
        p = lambda t: C.c_void_p(t.data_ptr())
        rc = lib.gsr_sort_pairs(p(kd), p(ko), p(vd), p(vo), p(kt), p(vt), n, end_bit, p(temp),
                                C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, lib.gsr_last_error()
        torch.cuda.synchronize()
        order = np.argsort(keys.numpy(), kind="stable")
        assert np.array_equal(ko.cpu().numpy(), keys.numpy()[order])
        assert np.array_equal(vo.cpu().numpy(), vals.numpy()[order])


def test_sort_pairs_at_c5_scale():
    """gsr_sort_pairs at the instance count of config C5 (n = 3e8, 45 key bits: six 8-bit passes, 73 K tiles per pass):
    32-bit cursor / offset arithmetic, the look-back status arrays and the temp sizing at full size.  Checked on the
    device against torch.sort(stable=True) (an independent implementation) — keys with the duplicate structure of
    tile|depth keys (few distinct high words, many equal depths), so stability is exercised."""
    from gs_localization_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    n, end_bit = 300_000_000, 45
    g = torch.Generator(device=DEV).manual_seed(1)
    tile = torch.randint(0, 8160, (n,), generator=g, dtype=torch.int64, device=DEV)
    depth = torch.randint(0, 1 << 20, (n,), generator=g, dtype=torch.int64, device=DEV) << 9     # 2^20 distinct depth words
    keys = (tile << 32) | depth
    del tile, depth
    vals = torch.arange(n, dtype=torch.int32, device=DEV)
    ko, vo, kt, vt = torch.empty_like(keys), torch.empty_like(vals), torch.empty_like(keys), torch.empty_like(vals)
    temp = torch.empty(lib.gsr_sort_temp_bytes(n), dtype=torch.uint8, device=DEV)
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.gsr_sort_pairs(p(keys), p(ko), p(vals), p(vo), p(kt), p(vt), n, end_bit, p(temp),
                            C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.gsr_last_error()
    torch.cuda.synchronize()
    del kt, vt, temp
    want_k, order = torch.sort(keys, stable=True)
    assert torch.equal(ko, want_k)
    del want_k, ko
    assert torch.equal(vo.to(torch.int64), order)


def test_speculative_launch_overflow_and_reuse():
    """The forward launches its tail speculatively into a binning buffer sized from the previous call with
    the same (P, W, H).  Drive it through: no history -> history hit -> a view with far more instances
    (overflow, tail re-run) -> back to a small view; every result must still match the oracle bit-exactly."""
    sc = dict(P=20_000, W=160, H=128, deg=1, f=120.0, sigma0=0.06)
    m, cam_a = util.scene(**sc, query=0)
    bg = torch.tensor([0.0, 0.1, 0.2])
    # a second map with the same P but much larger splats -> many more instances for the same shapes
    big = m._replace(scales=m.scales * 4.0)
    seq = [(m, cam_a), (m, cam_a), (big, cam_a), (m, cam_a), (big, cam_a)]
    Rs = []
    for mm, cam in seq:
        args, (R, color, depth, alpha, radii, geom, binning, img) = run_ours(mm, cam, bg)
        o = run_oracle(mm, cam, bg)
        assert R == o.R
        st = ours.export_state(mm.means3D.shape[0], R, cam.W, cam.H, geom, binning, img)
        b = o.binning()
        assert np.array_equal(st["keys"].cpu().numpy().view(np.uint64), b["keys"])
        assert np.array_equal(st["list"].cpu().numpy().view(np.uint32), b["list"])
        assert np.array_equal(st["ranges"].cpu().numpy().view(np.uint32), b["ranges"])
        oc, od, oa = o.images()
        _images_close(color, depth, alpha, oc, od, oa)
        # and the backward still finds the sorted list in the (possibly over-allocated) binning buffer
        got, _, _ = _grads_from(args, (R, color, depth, alpha, radii, geom, binning, img))
        want = o.backward(np.ones((3, cam.H, cam.W), np.float32), None, None)
        assert util.rel_err(got["dL_dmeans3D"], want["dL_dmeans3D"]) <= GRAD_REL_TOL
        Rs.append(R)
    assert Rs[2] > 1.3 * Rs[1]      # the third call really overflowed the 25 % headroom


def _grads_from(args, fwd):
    (R, color, depth, alpha, radii, geom, binning, img) = fwd
    (bgt, means3D, colors, opac, scales, rots, smod, cov, view, proj, tfx, tfy, H, W, sh, deg, campos, pf, dbg) = args
    res = ours.rasterize_gaussians_backward(bgt, means3D, radii, colors, scales, rots, smod, cov, view, proj, tfx, tfy,
                                            torch.ones_like(color), torch.zeros_like(depth), torch.zeros_like(alpha), sh, deg,
                                            campos, geom, R, binning, img, alpha, False)
    torch.cuda.synchronize()
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"]
    return dict(zip(names, [r.cpu().numpy() for r in res])), args, fwd


def test_backward_with_masked_upstream_gradients():
    """LoGS' tracking loss masks most pixels (tools/descent_utils.py:85-123): pixels with all-zero upstream
    gradients are skipped by the backward; the result must equal the unmasked computation of the same loss."""
    m, cam = util.scene(**SCENES["mid"])
    bg = torch.tensor([0.1, 0.3, 0.2])
    rng = np.random.default_rng(3)
    H, W = cam.H, cam.W
    mask = (rng.random((1, H, W)) < 0.15).astype(np.float32)
    wc = rng.standard_normal((3, H, W)).astype(np.float32) * mask
    wd = (0.3 * rng.standard_normal((1, H, W))).astype(np.float32) * mask
    wa = np.zeros((1, H, W), np.float32)
    got, _, _ = _grads_ours(m, cam, bg, wc, wd, wa)
    want = run_oracle(m, cam, bg, "f64").backward(wc, wd, wa)
    for k in ("dL_dmeans2D", "dL_dopacity", "dL_dmeans3D", "dL_dsh", "dL_dscales", "dL_drotations"):
        assert util.rel_err(got[k], want[k]) <= GRAD_REL_TOL, k


def test_unused_outputs_need_no_zero_gradients():
    """A loss on the colour image alone reaches the library with NULL depth / alpha upstream gradients (the compiled
    autograd node does not materialise them); the result must equal the same loss with explicit zero gradients."""
    from gs_localization_b200.diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    m, cam = util.scene(**SCENES["deg1"])
    view, proj, _, campos = cam.matrices(DEV)
    bg = torch.tensor([0.2, 0.1, 0.3], device=DEV)
    d = m.to(DEV)
    w = torch.from_numpy(np.random.default_rng(5).standard_normal((3, cam.H, cam.W)).astype(np.float32)).to(DEV)
    rs = GaussianRasterizationSettings(cam.H, cam.W, cam.tanfovx, cam.tanfovy, bg, 1.0, view, proj, m.sh_degree, campos, False, False)

    def grads(explicit_zeros):
        params = [t.clone().requires_grad_(True) for t in (d.means3D, d.shs, d.opacities, d.scales, d.rotations)]
        m2 = torch.zeros_like(params[0], requires_grad=True)
        color, radii, depth, alpha = GaussianRasterizer(rs)(means3D=params[0], means2D=m2, opacities=params[2], shs=params[1],
                                                             scales=params[3], rotations=params[4])
        loss = (color * w).sum()
        if explicit_zeros:
            loss = loss + (depth * 0.0).sum() + (alpha * 0.0).sum()
        loss.backward()
        return [p.grad.cpu().numpy() for p in params] + [m2.grad.cpu().numpy()]

    for a, b in zip(grads(False), grads(True)):
        assert np.isfinite(a).all() and np.abs(a).max() > 0
        assert util.rel_err(a, b) <= 1e-4


def test_async_forward_long_lists_match_default_forward():
    """gsr_rasterize_forward_async (caller-owned buffers, no host wait) on a frame with tile lists beyond the shared-memory
    sort, with global_sort=1 (the long-list path: Gaussian-level depth sort, ordered emission, stable tile passes) and
    with global_sort=0 (long buckets fall to the in-place network): identical images, lists and n_contrib as the
    default forward."""
    import ctypes as C
    from gs_localization_b200 import _lib
    lib = _lib.load()
    m, cam = util.scene(**SCENES["dense"])
    bg = torch.tensor([0.1, 0.3, 0.2])
    args, (R, color, depth, alpha, radii, geom, binning, img) = run_ours(m, cam, bg)
    P, W, H = m.means3D.shape[0], cam.W, cam.H
    st = ours.export_state(P, R, W, H, geom, binning, img)
    (bgt, means3D, colors, opac, scales, rots, smod, cov, view, proj, tfx, tfy, Hh, Ww, sh, deg, campos, pf, dbg) = args
    byte = dict(dtype=torch.uint8, device=DEV)
    cap = R + 1000
    g2 = torch.empty(lib.gsr_geometry_bytes(P), **byte)
    i2 = torch.empty(lib.gsr_image_bytes(W, H), **byte)
    b2 = torch.empty(lib.gsr_binning_bytes(cap, W, H), **byte)
    c2, d2, a2 = torch.empty_like(color), torch.empty_like(depth), torch.empty_like(alpha)
    r2 = torch.empty_like(radii)
    p = lambda t: C.c_void_p(t.data_ptr())
    # packed static map for the cull pass (gsr_build_cull_records): with and without, the results must be identical
    rec = torch.empty(P, 4, device=DEV)
    assert lib.gsr_build_cull_records(P, p(means3D), p(scales), p(rots), p(rec), C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    for global_sort, records in ((1, None), (0, None), (1, rec)):   # 0: long buckets fall to the in-place network, same result
        rc = lib.gsr_rasterize_forward_async(p(g2), p(b2), cap, global_sort, p(i2), P, deg, int(sh.shape[1]), p(bgt), W, H, p(means3D), p(sh),
                                             None, p(opac), p(scales), 1.0, p(rots), None, p(view), p(proj), p(campos), tfx, tfy,
                                             p(c2), p(d2), p(a2), p(r2), None, None if records is None else p(records),
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, lib.gsr_last_error()
        torch.cuda.synchronize()
        cnt = (C.c_uint * 3)()
        assert lib.gsr_read_counters(p(g2), P, cnt, C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
        assert cnt[0] == R and cnt[2] > 4096
        st2 = ours.export_state(P, R, W, H, g2, b2, i2)
        for k in ("keys", "list", "ranges", "n_contrib"):
            assert torch.equal(st[k], st2[k]), (global_sort, records is not None, k)
        assert torch.equal(color, c2) and torch.equal(alpha, a2) and torch.equal(depth, d2) and torch.equal(radii, r2)
