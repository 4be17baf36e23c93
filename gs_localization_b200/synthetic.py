"""Deterministic synthetic maps, cameras and pose perturbations for tests and bench.py.

The shapes follow BASELINE.json's configs and SURVEY.md §8(d): a "room" of Gaussians
(80 % on the six walls, 20 % in the interior), pinhole cameras built the way LoGS builds
them (`getProjectionMatrix2`, gs_localization/pipelines/tools/graphics_utils.py:77-98;
`viewmatrix = W2C^T`, `projmatrix = viewmatrix @ projmatrix_raw`,
tools/camera_utils.py:144-158) and the left-multiplicative SE(3) update of
tools/pose_utils.py:54-122.  Everything is generated on the CPU with a seeded
torch.Generator so that every rank / box sees identical data.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import NamedTuple

import torch

# name -> (P, W, H, sh_degree, fx, box scale, sigma0)
CONFIGS = {
    "C1": dict(P=10_000, W=160, H=120, deg=0, f=131.25, box=1.0, sigma0=0.05, iters=50),
    "C2": dict(P=300_000, W=640, H=480, deg=3, f=525.0, box=1.0, sigma0=0.02, iters=200),
    "headline": dict(P=1_000_000, W=640, H=480, deg=3, f=525.0, box=1.0, sigma0=0.02, iters=50),
    "C3": dict(P=2_000_000, W=1024, H=576, deg=3, f=744.0, box=10.0, sigma0=0.15, iters=20),
    "C4": dict(P=3_000_000, W=1297, H=840, deg=3, f=1040.0, box=1.0, sigma0=0.015, iters=1),
    "C5": dict(P=6_000_000, W=1920, H=1080, deg=3, f=1000.0, box=1.0, sigma0=0.075, iters=1),
}
ROOM = (6.0, 4.0, 6.0)  # metres, centred at the origin


class GaussianMap(NamedTuple):
    means3D: torch.Tensor   # [P,3]
    shs: torch.Tensor       # [P,M,3]
    opacities: torch.Tensor  # [P,1]  (post-sigmoid)
    scales: torch.Tensor    # [P,3]  (post-exp)
    rotations: torch.Tensor  # [P,4]  wxyz, normalised
    sh_degree: int

    def to(self, device):
        return GaussianMap(*[t.to(device) if isinstance(t, torch.Tensor) else t for t in self])


def make_map(P: int, sh_degree: int = 3, sigma0: float = 0.02, box: float = 1.0, seed: int = 0) -> GaussianMap:
    g = torch.Generator(device="cpu").manual_seed(seed)
    ext = torch.tensor(ROOM) * box
    n_wall = int(P * 0.8)
    # walls: pick a face, uniform on it, 1 cm normal jitter
    face = torch.randint(0, 6, (n_wall,), generator=g)
    u = torch.rand(n_wall, 3, generator=g) - 0.5
    pts = u * ext
    axis = face // 2
    sign = (face % 2).float() * 2 - 1
    jitter = torch.randn(n_wall, generator=g) * 0.01
    pts[torch.arange(n_wall), axis] = sign * ext[axis] * 0.5 + jitter
    interior = (torch.rand(P - n_wall, 3, generator=g) - 0.5) * ext
    means = torch.cat([pts, interior], 0)
    perm = torch.randperm(P, generator=g)
    means = means[perm].contiguous()
    rot = torch.randn(P, 4, generator=g)
    rot = rot / rot.norm(dim=1, keepdim=True)
    opacity = torch.sigmoid(torch.randn(P, 1, generator=g) * 2 + 1)
    M = (sh_degree + 1) ** 2
    shs = torch.randn(P, M, 3, generator=g) * 0.05
    shs[:, 0, :] = (torch.rand(P, 3, generator=g) - 0.5) / 0.28209479
    scales = torch.exp(torch.randn(P, 3, generator=g) * 0.5 + math.log(sigma0))
    return GaussianMap(means.float(), shs.float().contiguous(), opacity.float(), scales.float(), rot.float(), sh_degree)


def projection_matrix2(znear, zfar, cx, cy, fx, fy, W, H) -> torch.Tensor:
    """getProjectionMatrix2 (tools/graphics_utils.py:77-98), untransposed."""
    left = ((2 * cx - W) / W - 1.0) * W / 2.0
    right = ((2 * cx - W) / W + 1.0) * W / 2.0
    top = ((2 * cy - H) / H + 1.0) * H / 2.0
    bottom = ((2 * cy - H) / H - 1.0) * H / 2.0
    left, right = znear / fx * left, znear / fx * right
    top, bottom = znear / fy * top, znear / fy * bottom
    Pm = torch.zeros(4, 4, dtype=torch.float64)
    Pm[0, 0] = 2.0 * znear / (right - left)
    Pm[1, 1] = 2.0 * znear / (top - bottom)
    Pm[0, 2] = (right + left) / (right - left)
    Pm[1, 2] = (top + bottom) / (top - bottom)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    return Pm


def skew(v: torch.Tensor) -> torch.Tensor:
    x, y, z = v.tolist()
    return torch.tensor([[0, -z, y], [z, 0, -x], [-y, x, 0]], dtype=v.dtype)


def so3_exp(theta: torch.Tensor) -> torch.Tensor:
    """tools/pose_utils.py:54-70"""
    Wm = skew(theta)
    W2 = Wm @ Wm
    a = float(theta.norm())
    I = torch.eye(3, dtype=theta.dtype)
    if a < 1e-5:
        return I + Wm + 0.5 * W2
    return I + (math.sin(a) / a) * Wm + ((1 - math.cos(a)) / a**2) * W2


def se3_exp(tau: torch.Tensor) -> torch.Tensor:
    """tools/pose_utils.py:73-102; tau = [rho(3), theta(3)]."""
    rho, theta = tau[:3], tau[3:]
    Wm = skew(theta)
    W2 = Wm @ Wm
    a = float(theta.norm())
    I = torch.eye(3, dtype=tau.dtype)
    if a < 1e-5:
        V = I + 0.5 * Wm + (1.0 / 6.0) * W2
    else:
        V = I + Wm * ((1.0 - math.cos(a)) / a**2) + W2 * ((a - math.sin(a)) / a**3)
    T = torch.eye(4, dtype=tau.dtype)
    T[:3, :3] = so3_exp(theta)
    T[:3, 3] = V @ rho
    return T


@dataclass
class Camera:
    """World-to-camera pose + pinhole intrinsics; produces the rasterizer's per-view constants."""
    w2c: torch.Tensor  # [4,4] float64
    W: int
    H: int
    fx: float
    fy: float
    cx: float
    cy: float
    znear: float = 0.01
    zfar: float = 100.0

    @property
    def tanfovx(self):
        return self.W / (2.0 * self.fx)

    @property
    def tanfovy(self):
        return self.H / (2.0 * self.fy)

    def matrices(self, device="cpu"):
        """(viewmatrix, projmatrix, projmatrix_raw, campos) float32, in the transposed
        storage the rasterizer expects (scene/cameras.py:56-59)."""
        view = self.w2c.t().contiguous()
        raw = projection_matrix2(self.znear, self.zfar, self.cx, self.cy, self.fx, self.fy, self.W, self.H).t().contiguous()
        full = view @ raw
        campos = torch.linalg.inv(view)[3, :3]
        f = lambda t: t.float().contiguous().to(device)
        return f(view), f(full), f(raw), f(campos)

    def perturbed(self, tau: torch.Tensor) -> "Camera":
        """T_w2c <- exp(tau) @ T_w2c (tools/pose_utils.py:105-122)."""
        return Camera(se3_exp(tau.double()) @ self.w2c, self.W, self.H, self.fx, self.fy, self.cx, self.cy, self.znear, self.zfar)


def look_at_w2c(eye: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    eye, target = eye.double(), target.double()
    fwd = target - eye
    fwd = fwd / fwd.norm()
    up = torch.tensor([0.0, -1.0, 0.0], dtype=torch.float64)  # image y points down
    if abs(float(fwd @ up)) > 0.99:
        up = torch.tensor([1.0, 0.0, 0.0], dtype=torch.float64)
    right = torch.linalg.cross(up, fwd)  # x to the right, y down, z forward (COLMAP convention)
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    Rm = torch.stack([right, down, fwd], 0)
    T = torch.eye(4, dtype=torch.float64)
    T[:3, :3] = Rm
    T[:3, 3] = -Rm @ eye
    return T


def make_camera(cfg: dict, query: int = 0) -> Camera:
    """Ground-truth pose of query q: position U(0.6*box), look at a random wall point, seed 1000+q."""
    g = torch.Generator(device="cpu").manual_seed(1000 + query)
    ext = torch.tensor(ROOM, dtype=torch.float64) * cfg["box"]
    eye = (torch.rand(3, generator=g, dtype=torch.float64) - 0.5) * ext * 0.6
    face = int(torch.randint(0, 6, (1,), generator=g))
    tgt = (torch.rand(3, generator=g, dtype=torch.float64) - 0.5) * ext
    tgt[face // 2] = (1 if face % 2 else -1) * ext[face // 2] * 0.5
    f = cfg["f"]
    return Camera(look_at_w2c(eye, tgt), cfg["W"], cfg["H"], f, f, cfg["W"] / 2.0, cfg["H"] / 2.0)


def initial_perturbation(query: int = 0, trans_m: float = 0.05, rot_deg: float = 2.0) -> torch.Tensor:
    """tau0 with |rho| = 5 cm and |theta| = 2 deg in random directions (PnP-level error)."""
    g = torch.Generator(device="cpu").manual_seed(5000 + query)
    rho = torch.randn(3, generator=g, dtype=torch.float64)
    theta = torch.randn(3, generator=g, dtype=torch.float64)
    rho = rho / rho.norm() * trans_m
    theta = theta / theta.norm() * math.radians(rot_deg)
    return torch.cat([rho, theta])


def pose_error(w2c_a: torch.Tensor, w2c_b: torch.Tensor):
    """(translation error in metres of the camera centre, rotation error in degrees)."""
    a, b = w2c_a.double(), w2c_b.double()
    ca = -a[:3, :3].t() @ a[:3, 3]
    cb = -b[:3, :3].t() @ b[:3, 3]
    dR = a[:3, :3] @ b[:3, :3].t()
    # atan2 of (sin, cos) from the antisymmetric part: well conditioned for tiny angles, unlike acos(trace)
    A = dR - dR.t()
    sin = 0.5 * float(torch.stack([A[2, 1], A[0, 2], A[1, 0]]).norm())
    cos = (float(torch.trace(dR)) - 1) / 2
    return float((ca - cb).norm()), math.degrees(math.atan2(sin, cos))
