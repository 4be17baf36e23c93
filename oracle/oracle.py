"""TEST INFRASTRUCTURE — ctypes front-end of the CPU oracle (oracle/gsr_oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may
import this module.  The product package gs_localization_b200 never does.

`Oracle("f32")` restates the reference rasterizer bit-faithfully in float32 (reference
files cited function by function in gsr_oracle.cpp); `Oracle("f64")` evaluates the same
splat lists in double and is the ground truth for gradient tolerances.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgsr_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (building the checker is not using it)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "gsr_oracle.cpp")
    ):
        subprocess.check_call(["make", "-C", _HERE, "libgsr_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_expf.restype = C.c_float
        _lib.orc_expf.argtypes = [C.c_float]
        _lib.orc_get_higher_msb.restype = C.c_uint
        _lib.orc_get_higher_msb.argtypes = [C.c_uint]
        for pfx in ("orc32", "orc64"):
            getattr(_lib, pfx + "_create").restype = C.c_void_p
            getattr(_lib, pfx + "_destroy").argtypes = [C.c_void_p]
            f = getattr(_lib, pfx + "_forward")
            f.restype = C.c_longlong
            f.argtypes = (
                [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
                + [C.c_void_p] * 5
                + [C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int]
            )
            getattr(_lib, pfx + "_backward").argtypes = [C.c_void_p] * 4
            getattr(_lib, pfx + "_get_images").argtypes = [C.c_void_p] * 4
            getattr(_lib, pfx + "_get_geometry").argtypes = [C.c_void_p] * 10
            getattr(_lib, pfx + "_get_binning").argtypes = [C.c_void_p] * 8
            getattr(_lib, pfx + "_get_counters").argtypes = [C.c_void_p] * 3
            getattr(_lib, pfx + "_get_grads").argtypes = [C.c_void_p] * 11
    return _lib


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


def get_higher_msb(n: int) -> int:
    return lib().orc_get_higher_msb(int(n))


def _f32(a):
    if a is None:
        return None
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a if a.size else None


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One forward (+ optional backward) of the reference algorithm on the CPU."""

    def __init__(self, precision: str = "f32"):
        assert precision in ("f32", "f64")
        self.pfx = "orc32" if precision == "f32" else "orc64"
        self.real = np.float32 if precision == "f32" else np.float64
        self._l = lib()
        self._h = C.c_void_p(getattr(self._l, self.pfx + "_create")())
        self.P = self.M = self.W = self.H = 0
        self.R = 0

    def __del__(self):
        try:
            getattr(self._l, self.pfx + "_destroy")(self._h)
        except Exception:
            pass

    def _fn(self, name):
        return getattr(self._l, f"{self.pfx}_{name}")

    # argument names follow _C.rasterize_gaussians (rasterize_points.h:18-38)
    def forward(self, bg, means3D, colors_precomp, opacities, scales, rotations, scale_modifier, cov3D_precomp,
                viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos,
                count_touched=False):
        means3D = _f32(means3D)
        P = 0 if means3D is None else means3D.shape[0]
        sh = _f32(sh)
        M = 0 if sh is None else sh.shape[1]
        keep = [_f32(bg), means3D, sh, _f32(colors_precomp), _f32(opacities), _f32(scales), _f32(rotations),
                _f32(cov3D_precomp), _f32(viewmatrix), _f32(projmatrix), _f32(campos)]
        self._keep = keep
        self.P, self.M, self.W, self.H = P, M, int(image_width), int(image_height)
        self.R = int(self._fn("forward")(
            self._h, P, int(degree), M, _p(keep[0]), self.W, self.H, _p(keep[1]), _p(keep[2]), _p(keep[3]),
            _p(keep[4]), _p(keep[5]), float(scale_modifier), _p(keep[6]), _p(keep[7]), _p(keep[8]), _p(keep[9]),
            _p(keep[10]), float(tan_fovx), float(tan_fovy), int(bool(count_touched))))
        self._count_touched = bool(count_touched)
        return self.R

    def images(self):
        N = self.W * self.H
        color = np.empty((3, self.H, self.W), self.real)
        depth = np.empty((1, self.H, self.W), self.real)
        alpha = np.empty((1, self.H, self.W), self.real)
        assert color.size == 3 * N
        self._fn("get_images")(self._h, _p(color), _p(depth), _p(alpha))
        return color, depth, alpha

    def geometry(self):
        P = self.P
        out = dict(
            radii=np.empty(P, np.int32), tiles_touched=np.empty(P, np.uint32), point_offsets=np.empty(P, np.uint32),
            depths=np.empty(P, self.real), means2D=np.empty((P, 2), self.real), cov3D=np.empty((P, 6), self.real),
            conic_opacity=np.empty((P, 4), self.real), rgb=np.empty((P, 3), self.real),
            clamped=np.empty((P, 3), np.uint8))
        self._fn("get_geometry")(self._h, *[_p(out[k]) for k in (
            "radii", "tiles_touched", "point_offsets", "depths", "means2D", "cov3D", "conic_opacity", "rgb", "clamped")])
        return out

    def binning(self):
        R = self.R
        T = ((self.W + 15) // 16) * ((self.H + 15) // 16)
        out = dict(
            keys_unsorted=np.empty(R, np.uint64), list_unsorted=np.empty(R, np.uint32), keys=np.empty(R, np.uint64),
            list=np.empty(R, np.uint32), ranges=np.empty((T, 2), np.uint32),
            n_contrib=np.empty((self.H, self.W), np.uint32),
            n_touched=np.empty(self.P if self._count_touched else 0, np.int32))
        self._fn("get_binning")(self._h, *[_p(out[k]) for k in (
            "keys_unsorted", "list_unsorted", "keys", "list", "ranges", "n_contrib", "n_touched")])
        return out

    def counters(self):
        a, b = C.c_longlong(0), C.c_longlong(0)
        self._fn("get_counters")(self._h, C.byref(a), C.byref(b))
        return dict(pairs_evaluated=a.value, pairs_contributing=b.value)

    def backward(self, dL_dcolor, dL_ddepth, dL_dalpha):
        g = [_f32(dL_dcolor), _f32(dL_ddepth), _f32(dL_dalpha)]
        N = self.W * self.H
        if g[1] is None:
            g[1] = np.zeros(N, np.float32)
        if g[2] is None:
            g[2] = np.zeros(N, np.float32)
        assert g[0].size == 3 * N and g[1].size == N and g[2].size == N
        self._fn("backward")(self._h, _p(g[0]), _p(g[1]), _p(g[2]))
        P, M = self.P, self.M
        out = dict(
            dL_dmeans2D=np.empty((P, 3), self.real), dL_dcolors=np.empty((P, 3), self.real),
            dL_dopacity=np.empty((P, 1), self.real), dL_dmeans3D=np.empty((P, 3), self.real),
            dL_dcov3D=np.empty((P, 6), self.real), dL_dsh=np.empty((P, M, 3), self.real),
            dL_dscales=np.empty((P, 3), self.real), dL_drotations=np.empty((P, 4), self.real),
            dL_dconic=np.empty((P, 3), self.real), dL_dtau=np.empty(6, self.real))
        self._fn("get_grads")(self._h, *[_p(out[k]) for k in (
            "dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales",
            "dL_drotations", "dL_dconic", "dL_dtau")])
        return out
