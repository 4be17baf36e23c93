"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/gsr_b200.h declares, the Python surface mirrors the reference package, and the product
package never reaches for the oracle or a CPU fallback."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gsr_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"GSR_API\s+[\w\s\*]+?\b(gsr_\w+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ("gsr_rasterize_forward", "gsr_rasterize_backward", "gsr_mark_visible", "gsr_geometry_bytes",
              "gsr_image_bytes", "gsr_binning_bytes", "gsr_sort_pairs", "gsr_export_state", "gsr_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from gs_localization_b200 import _lib
    lib = _lib.load()
    exported = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    for s in declared_symbols():
        assert re.search(rf"\bT {s}\b", exported), f"{s} not exported"
        assert hasattr(lib, s)
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(declared_symbols())
    assert lib.gsr_abi_version() == 2


def test_scratch_size_queries_need_no_gpu():
    from gs_localization_b200 import _lib
    lib = _lib.load()
    g1, g2 = lib.gsr_geometry_bytes(1000), lib.gsr_geometry_bytes(2000)
    assert 0 < g1 < g2
    assert lib.gsr_binning_bytes(0, 640, 480) > 0
    assert lib.gsr_binning_bytes(10_000, 640, 480) >= 10_000 * 12
    assert lib.gsr_image_bytes(640, 480) >= 640 * 480 * 4 + 1200 * 8


def test_library_contains_sm100a_code_only():
    from gs_localization_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_python_surface_matches_reference():
    import gs_localization_b200.diff_gaussian_rasterization as d
    assert d.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "prefiltered", "debug")          # reference __init__.py:160-172
    for name in ("rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible"):
        assert callable(getattr(d._C, name))                    # ext.cpp:15-19
    assert issubclass(d.GaussianRasterizer, torch.nn.Module)
    assert hasattr(d.GaussianRasterizer, "markVisible")
    import gs_localization_b200.diff_gaussian_rasterization_pose as p
    assert "projmatrix_raw" in p.GaussianRasterizationSettings._fields   # tools/__init__.py:67


def _settings(d):
    return d.GaussianRasterizationSettings(8, 8, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0,
                                           torch.zeros(3), False, False)


def test_xor_argument_checks_raise_like_the_reference():
    import gs_localization_b200.diff_gaussian_rasterization as d
    r = d.GaussianRasterizer(_settings(d))
    m = torch.zeros(4, 3)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(m, m, torch.zeros(4, 1), scales=m, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(m, m, torch.zeros(4, 1), shs=torch.zeros(4, 1, 3))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(m, m, torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), scales=m, rotations=torch.zeros(4, 4), cov3D_precomp=torch.zeros(4, 6))


def test_no_cpu_fallback():
    """CPU tensors are rejected loudly; nothing is computed on the host."""
    import gs_localization_b200.diff_gaussian_rasterization as d
    r = d.GaussianRasterizer(_settings(d))
    m = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        r(m, m, torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), scales=m, rotations=torch.zeros(4, 4))
    with pytest.raises(RuntimeError, match="num_points, 3"):
        d._C.rasterize_gaussians(torch.zeros(3), torch.zeros(4, 2), *([torch.Tensor([])] * 4), 1.0, torch.Tensor([]),
                                 torch.eye(4), torch.eye(4), 1.0, 1.0, 8, 8, torch.Tensor([]), 0, torch.zeros(3), False, False)


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "gs_localization_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "libgsr_oracle" not in src, f


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from gs_localization_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.GsrError, match="no CPU / PyTorch fallback"):
        _lib.load()


def test_compiled_binding_loads_and_matches_abi():
    """binding/torch_binding.cpp: the compiled `_C` glue links the same C-ABI library (no compute without a GPU)."""
    from gs_localization_b200 import _lib, build
    build.build_binding()
    lib = _lib.load()
    from gs_localization_b200 import _gsr_torch
    assert _gsr_torch.abi_version() == lib.gsr_abi_version()
    for name in ("forward", "backward", "mark_visible"):
        assert callable(getattr(_gsr_torch, name))
    import gs_localization_b200.diff_gaussian_rasterization._C as C
    assert C._B is _gsr_torch
    import pytest
    import torch
    with pytest.raises(RuntimeError):                      # CPU tensors: loud failure, not a fallback
        C.rasterize_gaussians(torch.zeros(3), torch.zeros(4, 3), torch.Tensor([]), torch.zeros(4, 1), torch.zeros(4, 3),
                              torch.zeros(4, 4), 1.0, torch.Tensor([]), torch.eye(4), torch.eye(4), 1.0, 1.0, 8, 8,
                              torch.zeros(4, 1, 3), 0, torch.zeros(3), False, False)


def test_committed_ncu_profile_matches_the_kernel_sources():
    """bench.py's roofline block takes warp-instruction and DRAM-byte counts from profiles/r2_roofline_profile.json and only if
    that capture was taken on exactly the rasterizer sources in the tree (bench.kernel_source_sha256).  A kernel change
    without a new capture must not go unnoticed: the bench line would silently lose those fields."""
    import json
    import sys
    sys.path.insert(0, ROOT)
    import bench
    prof, note = bench.load_profile()
    assert prof is not None, note
    d = json.load(open(bench.PROFILE_JSON))
    names = " ".join(k["name"] for k in d["kernels"])
    for kernel in ("render_fwd_kernel", "render_bwd_kernel", "preprocess_cull_kernel", "preprocess_fwd_kernel", "preprocess_bwd_kernel",
                   "scan_tiles_kernel", "scatter_kernel", "tile_sort_kernel", "color_fwd_kernel"):
        assert kernel in names, kernel
