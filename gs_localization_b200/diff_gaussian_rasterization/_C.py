"""`_C` surface of the reference extension, implemented over the C ABI.

Mirrors the three pybind entry points of the reference
(gaussian_splatting/submodules/diff-gaussian-rasterization/ext.cpp:15-19,
rasterize_points.h:18-70) — same names, argument order, return tuples and error
behaviour — but the work is done by libgsr_b200.so (include/gsr_b200.h).  The torch
side only owns memory and the current stream.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib

NUM_CHANNELS = 3

# Two equivalent bindings over the same C ABI: the compiled one (binding/torch_binding.cpp -> _gsr_torch.so, the
# counterpart of the reference's rasterize_points.cu glue; ~3x less host time per call) and the ctypes one below.
# GSR_BINDING=ctypes|compiled forces one; by default the compiled module is used when it has been built.
import os as _os

_B = None
if _os.environ.get("GSR_BINDING", "compiled") != "ctypes":
    try:
        _lib.load()                      # resolves libgsr_b200.so first (and fails loudly if it is missing)
        from .. import _gsr_torch as _B  # noqa: F401
    except ImportError:
        if _os.environ.get("GSR_BINDING") == "compiled":
            raise
        _B = None
_NEED_BITS = dict(means3D=0, means2D=1, sh=2, colors=3, opacity=4, scales=5, rotations=6, cov3D=7)


def _ptr(t: torch.Tensor):
    """Device pointer of a contiguous float/int tensor; None for the reference's empty placeholders."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _cf(t: torch.Tensor, dev, dtype=torch.float32):
    """.contiguous().data<float>() of the reference (rasterize_points.cu:96-115)."""
    if t is None or t.numel() == 0:
        return None
    if t.dtype is dtype and t.is_contiguous() and t.device == dev:   # the hot case: nothing to do
        return t
    if t.device != dev:
        t = t.to(dev)
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


class _DeviceGuard:
    """torch.cuda.device(dev) only when dev is not already current (the context manager is not free)."""

    def __init__(self, dev):
        self.dev = dev
        self.ctx = None

    def __enter__(self):
        if torch.cuda.current_device() != (self.dev.index if self.dev.index is not None else torch.cuda.current_device()):
            self.ctx = torch.cuda.device(self.dev)
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


# Scratch allocation callbacks (gsr_alloc_fn).  Created once: building a ctypes callback per call is slow.
# The library calls them synchronously from inside gsr_rasterize_forward on the calling thread.
import threading

_tls = threading.local()


def _make_alloc(name):
    def alloc(nbytes, _user):
        t = torch.empty(int(nbytes), dtype=torch.uint8, device=_tls.dev)
        _tls.bufs[name] = t
        return t.data_ptr()
    return _lib.ALLOC_FN(alloc)


_ALLOC_GEOM, _ALLOC_BINNING, _ALLOC_IMG = _make_alloc("geom"), _make_alloc("binning"), _make_alloc("img")


_EMPTY = torch.Tensor([])


def _t(x):
    return _EMPTY if x is None else x


def _require_cuda(means3D):
    if not means3D.is_cuda:
        raise RuntimeError("gs_localization_b200: tensors must be CUDA tensors (no CPU fallback exists)")


def _forward_impl(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                  projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos, prefiltered, debug,
                  want_n_touched=False):
    if means3D.ndimension() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")  # rasterize_points.cu:57-59
    _require_cuda(means3D)
    if _B is not None:
        try:
            return _B.forward(background, means3D, _t(colors), opacity, _t(scales), _t(rotations), float(scale_modifier),
                              _t(cov3D_precomp), viewmatrix, projmatrix, float(tan_fovx), float(tan_fovy), int(image_height),
                              int(image_width), _t(sh), int(degree), campos, bool(prefiltered), bool(debug), bool(want_n_touched))
        except RuntimeError as ex:
            raise _lib.GsrError(str(ex)) from None
    lib = _lib.load()
    dev = means3D.device
    P, H, W = int(means3D.size(0)), int(image_height), int(image_width)
    f32 = dict(dtype=torch.float32, device=dev)
    byte = dict(dtype=torch.uint8, device=dev)
    n_touched = torch.zeros(P if want_n_touched else 0, dtype=torch.int32, device=dev)
    if P == 0:  # rasterize_points.cu:83: nothing is launched, outputs stay zero
        z = lambda *s: torch.zeros(*s, **f32)
        e = lambda: torch.empty(0, **byte)
        return (0, z(NUM_CHANNELS, H, W), z(1, H, W), z(1, H, W), torch.zeros(0, dtype=torch.int32, device=dev),
                e(), e(), e(), n_touched)

    out_color = torch.empty(NUM_CHANNELS, H, W, **f32)
    out_depth = torch.empty(1, H, W, **f32)
    out_alpha = torch.empty(1, H, W, **f32)
    radii = torch.empty(P, dtype=torch.int32, device=dev)
    M = int(sh.size(1)) if sh is not None and sh.numel() != 0 else 0

    keep = [_cf(background, dev), _cf(means3D, dev), _cf(sh, dev), _cf(colors, dev), _cf(opacity, dev), _cf(scales, dev),
            _cf(rotations, dev), _cf(cov3D_precomp, dev), _cf(viewmatrix, dev), _cf(projmatrix, dev), _cf(campos, dev)]
    bg, m3, shc, col, opa, sc, rot, cov, vm, pm, cp = keep
    bufs = {}
    _tls.dev, _tls.bufs = dev, bufs
    with _DeviceGuard(dev):
        R = lib.gsr_rasterize_forward(
            _ALLOC_GEOM, _ALLOC_BINNING, _ALLOC_IMG, None, P, int(degree), M,
            _ptr(bg), W, H, _ptr(m3), _ptr(shc), _ptr(col), _ptr(opa),
            _ptr(sc), float(scale_modifier), _ptr(rot), _ptr(cov),
            _ptr(vm), _ptr(pm), _ptr(cp), float(tan_fovx), float(tan_fovy), int(bool(prefiltered)),
            _ptr(out_color), _ptr(out_depth), _ptr(out_alpha), _ptr(radii), _ptr(n_touched) if want_n_touched else None,
            int(bool(debug)), _stream(dev))
    _lib.check(R, "gsr_rasterize_forward")
    e = torch.empty(0, **byte)
    return (int(R), out_color, out_depth, out_alpha, radii, bufs.get("geom", e), bufs.get("binning", e),
            bufs.get("img", e), n_touched)


def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                        projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos, prefiltered, debug):
    """RasterizeGaussiansCUDA (rasterize_points.cu:35-119).
    Returns (num_rendered, color, depth, alpha, radii, geomBuffer, binningBuffer, imgBuffer)."""
    return _forward_impl(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                         projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos, prefiltered,
                         debug)[:8]


def _backward_impl(background, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                   projmatrix, tan_fovx, tan_fovy, dL_dout_color, dL_dout_depth, dL_dout_alpha, sh, degree, campos,
                   geomBuffer, R, binningBuffer, imageBuffer, alpha, debug, projmatrix_raw=None, want_pose=False,
                   needs=None):
    _require_cuda(means3D)
    if _B is not None:
        mask = 0xff
        if needs:
            mask = 0
            for k, bit in _NEED_BITS.items():
                if needs.get(k, True):
                    mask |= 1 << bit
        try:
            res = _B.backward(background, means3D, radii, _t(colors), _t(scales), _t(rotations), float(scale_modifier),
                              _t(cov3D_precomp), viewmatrix, projmatrix, float(tan_fovx), float(tan_fovy), dL_dout_color,
                              dL_dout_depth, dL_dout_alpha, _t(sh), int(degree), campos, geomBuffer, int(R), binningBuffer,
                              imageBuffer, alpha, bool(debug), _t(projmatrix_raw), bool(want_pose), mask)
        except RuntimeError as ex:
            raise _lib.GsrError(str(ex)) from None
        return res[:8], res[8]
    lib = _lib.load()
    dev = means3D.device
    P = int(means3D.size(0))
    H, W = int(dL_dout_color.size(1)), int(dL_dout_color.size(2))
    M = int(sh.size(1)) if sh is not None and sh.numel() != 0 else 0
    f32 = dict(dtype=torch.float32, device=dev)
    needs = needs or {}
    want = lambda k: needs.get(k, True)
    # Every row is written by the kernel (zeros for culled Gaussians): torch.empty, not the
    # reference's nine torch::zeros (rasterize_points.cu:158-166).  P == 0 launches nothing.
    mk = torch.empty if P > 0 else torch.zeros
    dL_dmeans3D = mk(P, 3, **f32) if want("means3D") else None
    dL_dmeans2D = mk(P, 3, **f32) if want("means2D") else None
    dL_dcolors = mk(P, NUM_CHANNELS, **f32) if want("colors") else None
    dL_dopacity = mk(P, 1, **f32) if want("opacity") else None
    dL_dcov3D = mk(P, 6, **f32) if want("cov3D") else None
    dL_dsh = mk(P, M, 3, **f32) if want("sh") else None
    dL_dscales = mk(P, 3, **f32) if want("scales") else None
    dL_drotations = mk(P, 4, **f32) if want("rotations") else None
    dL_dtau = torch.zeros(6, **f32) if want_pose else None

    keep = [_cf(background, dev), _cf(means3D, dev), _cf(sh, dev), _cf(colors, dev), _cf(alpha, dev), _cf(scales, dev),
            _cf(rotations, dev), _cf(cov3D_precomp, dev), _cf(viewmatrix, dev), _cf(projmatrix, dev),
            _cf(projmatrix_raw, dev), _cf(campos, dev), _cf(radii, dev, torch.int32), _cf(dL_dout_color, dev),
            _cf(dL_dout_depth, dev), _cf(dL_dout_alpha, dev)]
    bg, m3, shc, col, alp, sc, rot, cov, vm, pm, praw, cp, rad, gC, gD, gA = keep
    if P > 0:
        with _DeviceGuard(dev):
            rc = lib.gsr_rasterize_backward(
                P, int(degree), M, int(R), _ptr(bg), W, H, _ptr(m3), _ptr(shc), _ptr(col), _ptr(alp),
                _ptr(sc), float(scale_modifier), _ptr(rot), _ptr(cov), _ptr(vm), _ptr(pm), _ptr(praw), _ptr(cp),
                float(tan_fovx), float(tan_fovy), _ptr(rad),
                _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer),
                _ptr(gC), _ptr(gD), _ptr(gA),
                _ptr(dL_dmeans2D), None, _ptr(dL_dopacity), _ptr(dL_dcolors),
                _ptr(dL_dmeans3D), _ptr(dL_dcov3D), _ptr(dL_dsh), _ptr(dL_dscales), _ptr(dL_drotations),
                _ptr(dL_dtau), int(bool(debug)), _stream(dev))
        _lib.check(rc, "gsr_rasterize_backward")
    return (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations), dL_dtau


def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp,
                                 viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, dL_dout_depth, dL_dout_alpha,
                                 sh, degree, campos, geomBuffer, R, binningBuffer, imageBuffer, alpha, debug):
    """RasterizeGaussiansBackwardCUDA (rasterize_points.cu:121-206).  Returns, in the reference's order
    (:205): (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)."""
    return _backward_impl(background, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                          projmatrix, tan_fovx, tan_fovy, dL_dout_color, dL_dout_depth, dL_dout_alpha, sh, degree, campos,
                          geomBuffer, R, binningBuffer, imageBuffer, alpha, debug)[0]


def mark_visible(means3D, viewmatrix, projmatrix):
    """markVisible (rasterize_points.cu:208-227): bool[P], True where view-space z > 0.2."""
    _require_cuda(means3D)
    if _B is not None:
        return _B.mark_visible(means3D, viewmatrix, projmatrix)
    lib = _lib.load()
    dev = means3D.device
    P = int(means3D.size(0))
    present = torch.zeros(P, dtype=torch.bool, device=dev)
    if P != 0:
        m3, vm, pm = _cf(means3D, dev), _cf(viewmatrix, dev), _cf(projmatrix, dev)
        with _DeviceGuard(dev):
            _lib.check(lib.gsr_mark_visible(P, _ptr(m3), _ptr(vm), _ptr(pm), present.data_ptr(), _stream(dev)),
                       "gsr_mark_visible")
    return present


def export_state(P, R, W, H, geomBuffer, binningBuffer, imgBuffer):
    """Test helper: the private intermediates, re-emitted in the reference's GeometryState /
    BinningState / ImageState formats (rasterizer_impl.h:29-64)."""
    lib = _lib.load()
    dev = geomBuffer.device if geomBuffer.numel() else imgBuffer.device
    T = ((W + 15) // 16) * ((H + 15) // 16)
    z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
    out = dict(
        depths=z(P, torch.float32), means2D=z((P, 2), torch.float32), cov3D=z((P, 6), torch.float32),
        conic_opacity=z((P, 4), torch.float32), rgb=z((P, 3), torch.float32), clamped=z((P, 3), torch.uint8),
        tiles_touched=z(P, torch.int32), keys=z(R, torch.int64), list=z(R, torch.int32),
        ranges=z((T, 2), torch.int32), n_contrib=z((H, W), torch.int32))
    order = ["depths", "means2D", "cov3D", "conic_opacity", "rgb", "clamped", "tiles_touched", "keys", "list", "ranges",
             "n_contrib"]
    with torch.cuda.device(dev):
        rc = lib.gsr_export_state(P, int(R), W, H, _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imgBuffer),
                                  *[_ptr(out[k]) for k in order], _stream(dev))
    _lib.check(rc, "gsr_export_state")
    torch.cuda.synchronize(dev)
    # the library keeps offsets per visible Gaussian only; the reference's per-Gaussian array is their prefix sum
    out["point_offsets"] = torch.cumsum(out["tiles_touched"].to(torch.int64), 0).to(torch.int32)
    return out
