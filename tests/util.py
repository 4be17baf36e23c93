"""Shared helpers for the parity tests: scenes, and one runner per implementation."""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")

from gs_localization_b200 import synthetic as syn  # noqa: E402


def scene(P=2000, W=96, H=64, deg=3, f=80.0, sigma0=0.12, box=1.0, seed=0, query=0):
    cfg = dict(P=P, W=W, H=H, deg=deg, f=f, box=box, sigma0=sigma0)
    m = syn.make_map(P, deg, sigma0, box, seed=seed)
    cam = syn.make_camera(cfg, query)
    return m, cam


def reference_available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "diff_gaussian_rasterization", "_C.so"))


def load_reference():
    """The UNMODIFIED reference package built by oracle/build_ref.sh, imported under its own name
    from oracle/_ref (our drop-in lives under gs_localization_b200.*, so there is no clash)."""
    if "diff_gaussian_rasterization" in sys.modules:
        mod = sys.modules["diff_gaussian_rasterization"]
        if os.path.dirname(os.path.dirname(mod.__file__)) == REF_DIR:
            return mod
        raise RuntimeError("another diff_gaussian_rasterization is already imported")
    sys.path.insert(0, REF_DIR)
    try:
        return importlib.import_module("diff_gaussian_rasterization")
    finally:
        sys.path.remove(REF_DIR)


def c_args(m, cam, bg, dev, colors_precomp=None, cov3D_precomp=None):
    """Argument tuple of _C.rasterize_gaussians (reference __init__.py:60-80)."""
    view, proj, raw, campos = cam.matrices(dev)
    e = torch.Tensor([])
    t = lambda x: x.to(dev)
    return (t(bg), t(m.means3D), e if colors_precomp is None else t(colors_precomp), t(m.opacities),
            e if cov3D_precomp is not None else t(m.scales), e if cov3D_precomp is not None else t(m.rotations), 1.0,
            e if cov3D_precomp is None else t(cov3D_precomp), view, proj, cam.tanfovx, cam.tanfovy, cam.H, cam.W,
            e if colors_precomp is not None else t(m.shs), m.sh_degree, campos, False, False)


def ref_unpack_state(P, R, W, H, geom, binning, img):
    """Carve the REFERENCE's three byte buffers (GeometryState/BinningState/ImageState::fromChunk,
    rasterizer_impl.cu:155-193): 128-byte aligned slabs in declaration order."""
    def carve(buf, off, count, dtype, itemsize):
        base = buf.data_ptr()
        a = (base + off + 127) // 128 * 128 - base
        n = count * itemsize
        return buf[a:a + n].view(dtype), a + n
    out = {}
    off = 0
    out["depths"], off = carve(geom, off, P, torch.float32, 4)
    cl, off = carve(geom, off, 3 * P, torch.uint8, 1)
    out["clamped"] = cl.view(P, 3)
    out["internal_radii"], off = carve(geom, off, P, torch.int32, 4)
    m2, off = carve(geom, off, 2 * P, torch.float32, 4)
    out["means2D"] = m2.view(P, 2)
    c3, off = carve(geom, off, 6 * P, torch.float32, 4)
    out["cov3D"] = c3.view(P, 6)
    co, off = carve(geom, off, 4 * P, torch.float32, 4)
    out["conic_opacity"] = co.view(P, 4)
    rgb, off = carve(geom, off, 3 * P, torch.float32, 4)
    out["rgb"] = rgb.view(P, 3)
    out["tiles_touched"], off = carve(geom, off, P, torch.int32, 4)
    off = 0
    out["list"], off = carve(binning, off, R, torch.int32, 4)
    out["list_unsorted"], off = carve(binning, off, R, torch.int32, 4)
    out["keys"], off = carve(binning, off, R, torch.int64, 8)
    out["keys_unsorted"], off = carve(binning, off, R, torch.int64, 8)
    off = 0
    N = W * H
    nc, off = carve(img, off, N, torch.int32, 4)
    out["n_contrib"] = nc.view(H, W)
    rg, off = carve(img, off, 2 * N, torch.int32, 4)
    T = ((W + 15) // 16) * ((H + 15) // 16)
    out["ranges"] = rg.view(N, 2)[:T]
    return out


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = np.linalg.norm(a - b)
    n = max(np.linalg.norm(b), 1e-30)
    return d / n
