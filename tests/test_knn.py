"""simple-knn distCUDA2: brute-force oracle vs an independent k-d tree on CPU; CUDA through the C ABI
bit-exact against the oracle and against the reference build (oracle/_ref/simple_knn) on the GPU."""
import os
import sys

import numpy as np
import pytest

from oracle import knn_oracle


def _cloud(P, seed, kind="surface"):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.random((P, 3)).astype(np.float32)
    # SfM-like: points on a few planes + clusters + outliers, very non-uniform density
    a = rng.random((P, 3)) * [6, 4, 0.01]
    b = rng.standard_normal((P, 3)) * 0.05 + rng.integers(-2, 3, (P, 1))
    c = rng.standard_normal((P, 3)) * 20
    pick = rng.random(P)
    return np.where(pick[:, None] < 0.6, a, np.where(pick[:, None] < 0.95, b, c)).astype(np.float32)


def test_knn_oracle_matches_kdtree():
    from scipy.spatial import cKDTree
    pts = _cloud(3000, 0)
    got = knn_oracle.dist2(pts)
    d, _ = cKDTree(pts.astype(np.float64)).query(pts.astype(np.float64), k=4)
    want = (d[:, 1:] ** 2).mean(axis=1)
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    assert np.allclose(got, want, rtol=1e-4, atol=1e-12)


def test_knn_oracle_small_and_duplicates():
    import math
    big = np.float32(3.4028234663852886e38)
    assert math.isinf(knn_oracle.dist2(np.zeros((1, 3)))[0])                       # no neighbours: FLT_MAX sums overflow
    two = knn_oracle.dist2(np.array([[0, 0, 0], [1, 0, 0]], np.float32))
    assert np.all(np.isinf(two))
    dup = knn_oracle.dist2(np.array([[0, 0, 0]] * 4 + [[2, 0, 0]], np.float32))
    assert np.array_equal(dup, np.float32([0, 0, 0, 0, 4.0])) and big > 0


def _cuda(pts):
    import torch
    from gs_localization_b200.simple_knn._C import distCUDA2
    return distCUDA2(torch.from_numpy(pts).cuda()).cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("P,kind,seed", [(1, "uniform", 0), (3, "uniform", 1), (4, "uniform", 2), (33, "surface", 3), (1000, "uniform", 4),
                                        (5000, "surface", 5), (20000, "surface", 6), (20000, "uniform", 7)])
def test_cuda_knn_bit_exact_vs_oracle(P, kind, seed):
    pts = _cloud(P, seed, kind)
    assert np.array_equal(_cuda(pts), knn_oracle.dist2(pts))


@pytest.mark.gpu
def test_cuda_knn_duplicates_and_degenerate_axes():
    pts = _cloud(4000, 9)
    pts[:500] = pts[0]                       # 500 coincident points
    pts[:, 2] = 1.5                          # zero extent along z
    assert np.array_equal(_cuda(pts), knn_oracle.dist2(pts))


@pytest.mark.gpu
def test_cuda_knn_bit_exact_vs_reference_build():
    import torch
    ref_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "simple_knn", "_C.so")):
        pytest.skip("oracle/_ref/simple_knn not built (reference sources absent at build time)")
    sys.path.insert(0, ref_dir)
    try:
        from simple_knn._C import distCUDA2 as ref_dist
    finally:
        sys.path.remove(ref_dir)
    for P, kind, seed in [(5000, "surface", 1), (300_000, "surface", 2), (200_000, "uniform", 3)]:
        pts = _cloud(P, seed, kind)
        want = ref_dist(torch.from_numpy(pts).cuda()).cpu().numpy()
        assert np.array_equal(_cuda(pts), want), (P, kind)
