"""Map / pose file formats: byte layout of the 3DGS PLY, load_ply's transposes, results_*.txt."""
import struct

import numpy as np
import torch

from gs_localization_b200 import io as gio
from gs_localization_b200 import synthetic as syn


def _raw(P=7, deg=2, seed=0):
    rng = np.random.default_rng(seed)
    M = (deg + 1) ** 2
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    return gio.RawGaussians(f(P, 3), f(P, 1, 3), f(P, M - 1, 3), f(P, 1), f(P, 3), f(P, 4))


def test_save_ply_byte_layout(tmp_path):
    """Header text and record layout as plyfile writes them for gaussian_model.py:192-208."""
    raw = _raw()
    path = str(tmp_path / "point_cloud" / "iteration_7" / "point_cloud.ply")
    gio.save_ply(path, raw)
    blob = open(path, "rb").read()
    head, body = blob.split(b"end_header\n", 1)
    lines = head.decode().splitlines()
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 7"]
    names = [l.split()[2] for l in lines[3:]]
    assert all(l.startswith("property float ") for l in lines[3:])
    assert names == ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{i}" for i in range(24)] + \
        ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    assert len(body) == 7 * len(names) * 4
    rec = struct.unpack("<%df" % len(names), body[2 * len(names) * 4:3 * len(names) * 4])      # Gaussian 2
    assert rec[:3] == tuple(raw.xyz[2]) and rec[3:6] == (0.0, 0.0, 0.0)
    assert rec[6:9] == tuple(raw.features_dc[2, 0])
    # channel-major f_rest: f_rest_{c*(M-1)+k} = features_rest[p, k, c]
    assert rec[9 + 1 * 8 + 5] == raw.features_rest[2, 5, 1]
    assert rec[33] == raw.opacity[2, 0] and rec[34:37] == tuple(raw.scaling[2]) and rec[37:41] == tuple(raw.rotation[2])


def test_ply_round_trip_and_activation(tmp_path):
    raw = _raw(P=100, deg=3, seed=1)
    path = str(tmp_path / "m.ply")
    gio.save_ply(path, raw)
    back = gio.load_ply_raw(path, max_sh_degree=3)
    for a, b in zip(raw, back):
        assert a.shape == b.shape and np.array_equal(a, b)
    m = gio.load_ply(path)
    assert m.sh_degree == 3 and m.shs.shape == (100, 16, 3)
    assert torch.equal(m.shs[:, 0], torch.from_numpy(raw.features_dc[:, 0]))
    assert torch.allclose(m.opacities, torch.sigmoid(torch.from_numpy(raw.opacity)))
    assert torch.allclose(m.scales, torch.exp(torch.from_numpy(raw.scaling)))
    assert torch.allclose(m.rotations.norm(dim=1), torch.ones(100))
    # rasterizer-side map -> file -> map
    g = syn.make_map(50, 1, 0.05, 1.0, seed=3)
    gio.save_ply(path, gio.deactivate(g))
    g2 = gio.load_ply(path)
    for a, b in zip(g[:5], g2[:5]):
        assert torch.allclose(a, b, atol=1e-5, rtol=1e-5)
    try:
        gio.load_ply_raw(path, max_sh_degree=3)
        assert False, "expected the reference's f_rest-count assertion"
    except AssertionError:
        pass


def test_load_ascii_big_endian_and_shuffled_properties(tmp_path):
    """A reader must not depend on property order or storage type (plyfile does not)."""
    names = ["opacity", "x", "z", "y", "rot_3", "rot_0", "rot_1", "rot_2", "scale_2", "scale_0", "scale_1", "f_dc_2", "f_dc_0", "f_dc_1"]
    rows = np.arange(2 * len(names), dtype=np.float64).reshape(2, -1) / 8
    p = tmp_path / "a.ply"
    p.write_text("ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 2\n" + "".join(f"property double {n}\n" for n in names) +
                 "end_header\n" + "\n".join(" ".join(repr(float(v)) for v in r) for r in rows) + "\n")
    a = gio.load_ply_raw(str(p))
    q = tmp_path / "b.ply"
    with open(q, "wb") as f:
        f.write(("ply\nformat binary_big_endian 1.0\nelement vertex 2\n" + "".join(f"property float {n}\n" for n in names) + "end_header\n").encode())
        f.write(rows.astype(">f4").tobytes())
    b = gio.load_ply_raw(str(q))
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    c = dict(zip(names, rows[1]))
    assert tuple(a.xyz[1]) == (c["x"], c["y"], c["z"]) and tuple(a.rotation[1]) == tuple(c[f"rot_{i}"] for i in range(4))
    assert tuple(a.scaling[1]) == tuple(c[f"scale_{i}"] for i in range(3)) and a.features_rest.shape == (2, 0, 3) and a.sh_degree == 0


def test_pose_results_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    poses = {}
    for i in range(5):
        w2c = syn.se3_exp(torch.from_numpy(rng.standard_normal(6))).numpy()
        poses[f"seq-0{5 - i}/frame-{i:06d}.color.png"] = w2c
    path = str(tmp_path / "results_dense.txt")
    gio.write_results(path, poses)
    back = gio.read_results(path)
    assert list(back) == sorted(poses)                      # name-sorted like the reference
    for k, v in back.items():
        assert np.abs(v.R - poses[k][:3, :3]).max() < 1e-12 and np.abs(v.T - poses[k][:3, 3]).max() < 1e-12
    # the reference's own quaternion convention: (w, x, y, z), 90 degrees about z
    R = gio.quat_to_rotmat([np.sqrt(0.5), 0, 0, np.sqrt(0.5)])
    assert np.allclose(R, [[0, -1, 0], [1, 0, 0], [0, 0, 1]])
    assert np.allclose(gio.rotmat_to_quat(R), [np.sqrt(0.5), 0, 0, np.sqrt(0.5)])


# ---------------------------------------------------------------------------------------------------- reference goldens
# tests/golden/ref_map.ply is what the REFERENCE's GaussianModel.save_ply wrote for the parameters stored in ref_map.npz,
# and ref_map.npz also holds what its load_ply / getters returned for that file (tests/golden/make_ply_golden.py).
import os

_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _golden_raw():
    z = np.load(os.path.join(_GOLD, "ref_map.npz"))
    return z, gio.RawGaussians(z["in_xyz"], z["in_f_dc"], z["in_f_rest"], z["in_opacity"], z["in_scaling"], z["in_rotation"])


def test_save_ply_writes_the_reference_bytes(tmp_path):
    """Same parameters -> the file the reference's save_ply wrote, byte for byte (gaussian_model.py:192-208)."""
    _, raw = _golden_raw()
    path = str(tmp_path / "ours.ply")
    gio.save_ply(path, raw)
    assert open(path, "rb").read() == open(os.path.join(_GOLD, "ref_map.ply"), "rb").read()


def test_load_ply_matches_the_reference_loader():
    """The reference-written file read by load_ply_raw / activate equals what the reference's load_ply and getters gave."""
    z, _ = _golden_raw()
    raw = gio.load_ply_raw(os.path.join(_GOLD, "ref_map.ply"), max_sh_degree=int(z["sh_degree"]))
    for ours, key in ((raw.xyz, "ld_xyz"), (raw.features_dc, "ld_f_dc"), (raw.features_rest, "ld_f_rest"), (raw.opacity, "ld_opacity"),
                      (raw.scaling, "ld_scaling"), (raw.rotation, "ld_rotation")):
        assert ours.dtype == np.float32 and ours.shape == z[key].shape, key
        assert np.array_equal(ours, z[key]), key
    g = gio.activate(raw)
    assert g.sh_degree == int(z["sh_degree"])
    assert np.array_equal(g.shs.numpy(), z["get_features"])
    # sigmoid / exp / normalize are torch's own ops in both; allow for vectorisation differences of one ulp
    np.testing.assert_allclose(g.opacities.numpy(), z["get_opacity"], rtol=2e-7, atol=0)
    np.testing.assert_allclose(g.scales.numpy(), z["get_scaling"], rtol=2e-7, atol=0)
    np.testing.assert_allclose(g.rotations.numpy(), z["get_rotation"], rtol=0, atol=2e-7)
    # the reference asserts on the f_rest count for a different SH degree (gaussian_model.py:230)
    import pytest
    with pytest.raises(AssertionError):
        gio.load_ply_raw(os.path.join(_GOLD, "ref_map.ply"), max_sh_degree=2)
