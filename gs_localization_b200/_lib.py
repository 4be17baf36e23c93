"""ctypes binding of libgsr_b200.so (C ABI declared in include/gsr_b200.h)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgsr_b200.so")

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_size_t, C.c_void_p)

_P = C.c_void_p
_lib = None

# name -> (restype, argtypes); one entry per symbol declared in include/gsr_b200.h
SIGNATURES = {
    "gsr_abi_version": (C.c_int, []),
    "gsr_last_error": (C.c_char_p, []),
    "gsr_launch_count": (C.c_ulonglong, []),
    "gsr_geometry_bytes": (C.c_size_t, [C.c_int]),
    "gsr_image_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "gsr_binning_bytes": (C.c_size_t, [C.c_longlong, C.c_int, C.c_int]),
    "gsr_rasterize_forward": (C.c_longlong, [
        ALLOC_FN, ALLOC_FN, ALLOC_FN, _P,
        C.c_int, C.c_int, C.c_int,
        _P, C.c_int, C.c_int,
        _P, _P, _P, _P,
        _P, C.c_float, _P, _P,
        _P, _P, _P,
        C.c_float, C.c_float, C.c_int,
        _P, _P, _P, _P, _P,
        C.c_int, _P]),
    "gsr_rasterize_forward_async": (C.c_int, [
        _P, _P, C.c_longlong, C.c_int, _P,
        C.c_int, C.c_int, C.c_int,
        _P, C.c_int, C.c_int,
        _P, _P, _P, _P,
        _P, C.c_float, _P, _P,
        _P, _P, _P,
        C.c_float, C.c_float,
        _P, _P, _P, _P, _P, _P, _P]),
    "gsr_build_cull_records": (C.c_int, [C.c_int, _P, _P, _P, _P, _P]),
    "gsr_read_counters": (C.c_int, [_P, C.c_int, _P, _P]),
    "gsr_clear_overflow": (C.c_int, [_P, C.c_int, _P]),
    "gsr_rasterize_backward": (C.c_int, [
        C.c_int, C.c_int, C.c_int, C.c_longlong,
        _P, C.c_int, C.c_int,
        _P, _P, _P, _P,
        _P, C.c_float, _P, _P,
        _P, _P, _P, _P,
        C.c_float, C.c_float, _P,
        _P, _P, _P,
        _P, _P, _P,
        _P, _P, _P, _P,
        _P, _P, _P, _P, _P,
        _P,
        C.c_int, _P]),
    "gsr_mark_visible": (C.c_int, [C.c_int, _P, _P, _P, _P, _P]),
    "gsr_export_state": (C.c_int, [C.c_int, C.c_longlong, C.c_int, C.c_int] + [_P] * 3 + [_P] * 11 + [_P]),
    "gsr_stage_timing": (None, [C.c_int]),
    "gsr_stage_times": (C.c_int, [_P, _P, C.c_int]),
    "gsr_l1_loss_grad": (C.c_int, [_P, _P, _P, C.c_longlong, C.c_float, _P, _P]),
    "gsr_l1_ssim_loss_grad": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P, _P, _P]),
    "gsr_map_adam_step": (C.c_int, [C.c_int] * 4 + [_P, _P, C.c_float, C.c_float, C.c_float] + [_P] * 13),
    "gsr_pack_gradient_rows": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "gsr_add_gradient_rows": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "gsr_pack_visible_rows": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, C.c_int, _P, _P]),
    "gsr_add_counted_rows": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P, _P, _P, _P]),
    "gsr_knn_workspace_bytes": (C.c_size_t, [C.c_longlong]),
    "gsr_dist2_knn3": (C.c_int, [_P, C.c_longlong, _P, _P, _P]),
    "gsr_depth_loss_grad": (C.c_int, [_P, _P, _P, C.c_longlong, C.c_float, C.c_float, C.c_float, _P, _P, _P, _P]),
    "gsr_tracking_loss_grad": (C.c_int, [_P] * 7 + [C.c_int, C.c_int, C.c_float, C.c_float, _P, _P, _P, _P, _P]),
    "gsr_exposure_adam_step": (C.c_int, [_P, _P, _P, _P, _P, C.c_float, _P]),
    "gsr_pose_adam_step": (C.c_int, [_P, _P, _P, _P, C.c_float, C.c_float, _P, _P, _P, _P, _P, _P, _P]),
    "gsr_sort_temp_bytes": (C.c_size_t, [C.c_longlong]),
    "gsr_sort_pairs": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_longlong, C.c_int, _P, _P]),
}


class GsrError(RuntimeError):
    pass


def load():
    """Load the CUDA library.  Raises if it has not been built — there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GsrError(
            f"{LIB_PATH} is missing: build it with `python -m gs_localization_b200.build` "
            "(or __graft_entry__.build()). gs_localization_b200 has no CPU / PyTorch fallback.")
    import torch  # noqa: F401  (makes libcudart.so.12 resident before dlopen)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.gsr_abi_version() != 2:
        raise GsrError("libgsr_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(code: int, what: str):
    if code < 0:
        msg = load().gsr_last_error().decode("utf-8", "replace")
        raise GsrError(f"{what} failed ({code}): {msg}")
    return code


STAGES = ("preprocess", "duplicate_with_keys", "radix_sort", "tile_ranges", "render", "render_backward",
          "preprocess_backward")


def stage_timing(enable: bool) -> None:
    load().gsr_stage_timing(int(bool(enable)))


def stage_times() -> dict:
    """{stage: mean ms per call} accumulated since stage_timing(True)."""
    n = len(STAGES)
    ms = (C.c_double * n)()
    calls = (C.c_ulonglong * n)()
    load().gsr_stage_times(ms, calls, n)
    return {STAGES[i]: (ms[i] / calls[i] if calls[i] else 0.0) for i in range(n)}


def launch_count() -> int:
    return int(load().gsr_launch_count())
