"""Map and pose files of LoGS, read and written without third-party packages.

* 3DGS PLY (gaussian_splatting/scene/gaussian_model.py:177-208 save_ply, :215-256 load_ply): one `vertex` element,
  all properties `float`, order x y z nx ny nz f_dc_* f_rest_* opacity scale_* rot_*; SH coefficients are stored
  channel-major (`[P,3,M-1]` flattened) and transposed on load; values are the raw (pre-activation) parameters.
  The reference goes through the `plyfile` package (binary_little_endian on x86); this module writes that byte
  layout directly and reads binary (either endianness) and ascii files with arbitrary property order/types.
* pose lists `results_*.txt` (gs_localization/pipelines/7scenes_localize_full_dslam.py:330-341): one line per query,
  `name qw qx qy qz tx ty tz`, world-to-camera.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict
from typing import NamedTuple

import numpy as np
import torch

from .synthetic import GaussianMap

_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
              "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
              "double": "f8", "float64": "f8"}


class RawGaussians(NamedTuple):
    """The optimiser-side (pre-activation) parameters exactly as GaussianModel holds them."""
    xyz: np.ndarray            # [P,3]
    features_dc: np.ndarray    # [P,1,3]
    features_rest: np.ndarray  # [P,M-1,3]
    opacity: np.ndarray        # [P,1]   (logit)
    scaling: np.ndarray        # [P,3]   (log)
    rotation: np.ndarray       # [P,4]   (wxyz, unnormalised)

    @property
    def sh_degree(self) -> int:
        return int(round(math.sqrt(self.features_rest.shape[1] + 1))) - 1


def attribute_names(n_dc: int, n_rest: int, n_scale: int = 3, n_rot: int = 4):
    """construct_list_of_attributes (gaussian_model.py:177-190)."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(n_dc)] + [f"f_rest_{i}" for i in range(n_rest)] + ["opacity"]
    names += [f"scale_{i}" for i in range(n_scale)] + [f"rot_{i}" for i in range(n_rot)]
    return names


def save_ply(path: str, raw: RawGaussians) -> None:
    """gaussian_model.py:192-208: normals are zeros; f_dc / f_rest are transposed to channel-major before flattening."""
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    P = raw.xyz.shape[0]
    f_dc = np.ascontiguousarray(np.transpose(raw.features_dc, (0, 2, 1))).reshape(P, -1)
    f_rest = np.ascontiguousarray(np.transpose(raw.features_rest, (0, 2, 1))).reshape(P, -1)
    table = np.concatenate([raw.xyz, np.zeros_like(raw.xyz), f_dc, f_rest, raw.opacity.reshape(P, 1), raw.scaling, raw.rotation],
                           axis=1).astype("<f4")
    names = attribute_names(f_dc.shape[1], f_rest.shape[1], raw.scaling.shape[1], raw.rotation.shape[1])
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % P
    header += "".join(f"property float {n}\n" for n in names) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(table.tobytes())


def _read_vertex_table(path: str):
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements, cur = None, [], None
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii").split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                cur = {"name": tok[1], "count": int(tok[2]), "props": []}
                elements.append(cur)
            elif tok[0] == "property":
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties are not part of the 3DGS map format")
                cur["props"].append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if not elements or elements[0]["name"] != "vertex":
            raise ValueError(f"{path}: first element must be 'vertex'")
        el = elements[0]
        if fmt == "ascii":
            data = np.loadtxt(f, dtype=np.float64, max_rows=el["count"], ndmin=2)
            if data.shape != (el["count"], len(el["props"])):
                raise ValueError(f"{path}: vertex table has shape {data.shape}")
            return {n: data[:, i] for i, (n, _) in enumerate(el["props"])}, el["count"]
        if fmt not in ("binary_little_endian", "binary_big_endian"):
            raise ValueError(f"{path}: unknown PLY format {fmt}")
        bo = "<" if fmt == "binary_little_endian" else ">"
        dt = np.dtype([(n, bo + t) for n, t in el["props"]])
        rec = np.fromfile(f, dtype=dt, count=el["count"])
        if rec.shape[0] != el["count"]:
            raise ValueError(f"{path}: truncated vertex data ({rec.shape[0]} of {el['count']})")
        return {n: rec[n] for n, _ in el["props"]}, el["count"]


def load_ply_raw(path: str, max_sh_degree: int | None = None) -> RawGaussians:
    """gaussian_model.py:215-256.  `max_sh_degree` reproduces the reference's assert on the f_rest count; None infers it."""
    col, P = _read_vertex_table(path)
    by_index = lambda prefix: sorted((n for n in col if n.startswith(prefix)), key=lambda n: int(n.split("_")[-1]))
    stack = lambda names: (np.stack([np.asarray(col[n], np.float32) for n in names], axis=1) if names
                           else np.zeros((P, 0), np.float32))
    xyz = stack(["x", "y", "z"])
    dc = stack(["f_dc_0", "f_dc_1", "f_dc_2"]).reshape(P, 3, 1)
    rest_names = by_index("f_rest_")
    if max_sh_degree is not None and len(rest_names) != 3 * (max_sh_degree + 1) ** 2 - 3:
        raise AssertionError(f"{path}: {len(rest_names)} f_rest properties, expected {3 * (max_sh_degree + 1) ** 2 - 3}")
    if len(rest_names) % 3:
        raise ValueError(f"{path}: f_rest count {len(rest_names)} is not a multiple of 3")
    rest = stack(rest_names).reshape(P, 3, len(rest_names) // 3)
    return RawGaussians(xyz, np.ascontiguousarray(dc.transpose(0, 2, 1)), np.ascontiguousarray(rest.transpose(0, 2, 1)),
                        np.asarray(col["opacity"], np.float32)[:, None], stack(by_index("scale_")), stack(by_index("rot")))


def activate(raw: RawGaussians, device="cpu") -> GaussianMap:
    """The getters the renderer reads (gaussian_model.py:96-115): sigmoid opacity, exp scaling, normalised rotation,
    features = cat(dc, rest) -> [P,M,3]."""
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(device)
    shs = torch.cat([t(raw.features_dc), t(raw.features_rest)], dim=1).contiguous()
    return GaussianMap(t(raw.xyz), shs, torch.sigmoid(t(raw.opacity)), torch.exp(t(raw.scaling)),
                       torch.nn.functional.normalize(t(raw.rotation)), raw.sh_degree)


def load_ply(path: str, device="cpu", max_sh_degree: int | None = None) -> GaussianMap:
    return activate(load_ply_raw(path, max_sh_degree), device)


def deactivate(gmap: GaussianMap) -> RawGaussians:
    """Inverse of `activate` (inverse_sigmoid / log), for writing a rasterizer-side map back to disk."""
    n = lambda x: x.detach().cpu().numpy().astype(np.float32)
    o = gmap.opacities.detach().cpu().double().clamp(1e-7, 1 - 1e-7)
    return RawGaussians(n(gmap.means3D), n(gmap.shs[:, :1]), n(gmap.shs[:, 1:]), torch.log(o / (1 - o)).float().numpy(),
                        np.log(n(gmap.scales)), n(gmap.rotations))


# ---------------------------------------------------------------------------------------------------- poses
def quat_to_rotmat(qvec) -> np.ndarray:
    """7scenes_localize_full_dslam.py:101-109 (w, x, y, z)."""
    w, x, y, z = np.array(qvec, dtype=float)
    return np.array([[1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * z * w, 2 * x * z + 2 * y * w],
                     [2 * x * y + 2 * z * w, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * x * w],
                     [2 * x * z - 2 * y * w, 2 * y * z + 2 * x * w, 1 - 2 * x * x - 2 * y * y]])


def rotmat_to_quat(R) -> np.ndarray:
    """Unit quaternion (w, x, y, z), w >= 0, of a rotation matrix (largest-pivot branch for stability)."""
    R = np.asarray(R, dtype=float)
    K = np.array([[R[0, 0] - R[1, 1] - R[2, 2], 0, 0, 0],
                  [R[0, 1] + R[1, 0], R[1, 1] - R[0, 0] - R[2, 2], 0, 0],
                  [R[0, 2] + R[2, 0], R[1, 2] + R[2, 1], R[2, 2] - R[0, 0] - R[1, 1], 0],
                  [R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1], R[0, 0] + R[1, 1] + R[2, 2]]]) / 3.0
    vals, vecs = np.linalg.eigh(K)
    x, y, z, w = vecs[:, np.argmax(vals)]
    q = np.array([w, x, y, z])
    return -q if q[0] < 0 else q


class Transformation:
    """7scenes_localize_full_dslam.py:96-99."""

    def __init__(self, R=None, T=None):
        self.R, self.T = R, T


def read_results(path: str) -> "OrderedDict[str, Transformation]":
    """`name qw qx qy qz tx ty tz` per line -> name-sorted OrderedDict (7scenes_localize_full_dslam.py:327-344)."""
    infos = OrderedDict()
    with open(path, "r") as f:
        for line in f:
            parts = line.strip().split()
            if not parts:
                continue
            infos[parts[0]] = Transformation(R=quat_to_rotmat(list(map(float, parts[1:5]))), T=np.array(list(map(float, parts[5:8]))))
    return OrderedDict(sorted(infos.items(), key=lambda item: item[0]))


def write_results(path: str, poses) -> None:
    """Inverse of read_results; `poses` maps name -> Transformation or (R, T) or a [4,4] world-to-camera matrix."""
    with open(path, "w") as f:
        for name, v in poses.items():
            if isinstance(v, Transformation):
                R, T = v.R, v.T
            elif isinstance(v, (tuple, list)):
                R, T = v
            else:
                M = np.asarray(v.detach().cpu() if torch.is_tensor(v) else v, dtype=float)
                R, T = M[:3, :3], M[:3, 3]
            q = rotmat_to_quat(np.asarray(R, dtype=float))
            f.write(" ".join([name] + [repr(float(c)) for c in list(q) + list(np.asarray(T, dtype=float).reshape(3))]) + "\n")
