"""Seeded synthetic scenes and cameras for the parity tests, smoke() and bench.py.

Distributions and camera rig are the ones fixed in BASELINE.md §2.3 / SURVEY.md §8d; the camera
attributes carry the names render() reads from the reference's Camera
(scene/cameras.py:62-74: world_view_transform, projection_matrix, full_proj_transform,
camera_center, FoVx/FoVy, image_height/width).  Matrices are stored the way the reference stores
them: the tensor is the TRANSPOSE of the column-vector matrix (row-vector convention), so flat
index m[4*c + r] is element (r, c).  Everything here is host-side numpy/torch on CPU; callers move
the tensors.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch


def world_to_view(R: np.ndarray, t: np.ndarray) -> np.ndarray:
    """4x4 world->camera matrix from a camera-to-world rotation R and world->camera translation t
    (convention of utils/graphics_utils.py:42-53 with translate=0, scale=1)."""
    m = np.eye(4, dtype=np.float64)
    m[:3, :3] = R.T
    m[:3, 3] = t
    return m.astype(np.float32)


def projection_matrix(znear: float, zfar: float, fovx: float, fovy: float) -> np.ndarray:
    """Perspective matrix with z in [0,1] and w = +z (utils/graphics_utils.py:56-76)."""
    tx, ty = math.tan(fovx / 2), math.tan(fovy / 2)
    top, right = ty * znear, tx * znear
    p = np.zeros((4, 4), dtype=np.float32)
    p[0, 0] = 2.0 * znear / (2 * right)
    p[1, 1] = 2.0 * znear / (2 * top)
    p[3, 2] = 1.0
    p[2, 2] = zfar / (zfar - znear)
    p[2, 3] = -(zfar * znear) / (zfar - znear)
    return p


def fov2focal(fov: float, pixels: int) -> float:
    return pixels / (2 * math.tan(fov / 2))


def focal2fov(focal: float, pixels: int) -> float:
    return 2 * math.atan(pixels / (2 * focal))


@dataclass
class SynthCamera:
    """Duck-types the attributes gaussian_renderer.render() reads (gaussian_renderer/__init__.py:56-70)."""
    FoVx: float
    FoVy: float
    image_height: int
    image_width: int
    world_view_transform: torch.Tensor  # [4,4] = W2C^T
    projection_matrix: torch.Tensor     # [4,4] = P^T
    full_proj_transform: torch.Tensor   # [4,4] = (P W2C)^T
    camera_center: torch.Tensor         # [3]
    znear: float = 0.01
    zfar: float = 100.0

    def to(self, device):
        return SynthCamera(self.FoVx, self.FoVy, self.image_height, self.image_width,
                           self.world_view_transform.to(device), self.projection_matrix.to(device),
                           self.full_proj_transform.to(device), self.camera_center.to(device), self.znear,
                           self.zfar)


def make_camera(R: np.ndarray, T: np.ndarray, fovx: float, fovy: float, H: int, W: int, znear=0.01,
                zfar=100.0) -> SynthCamera:
    wvt = torch.tensor(world_to_view(R, T)).transpose(0, 1).contiguous()
    proj = torch.tensor(projection_matrix(znear, zfar, fovx, fovy)).transpose(0, 1).contiguous()
    full = (wvt.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0).contiguous()
    center = wvt.inverse()[3, :3].contiguous()
    return SynthCamera(fovx, fovy, H, W, wvt, proj, full, center, znear, zfar)


def look_at_camera(eye: np.ndarray, target: np.ndarray, fovx: float, fovy: float, H: int, W: int) -> SynthCamera:
    """OpenCV-convention camera (x right, y down, z forward) at `eye` looking at `target`."""
    fwd = target - eye
    fwd = fwd / np.linalg.norm(fwd)
    up = np.array([0.0, 0.0, 1.0])
    if abs(np.dot(fwd, up)) > 0.999:
        up = np.array([0.0, 1.0, 0.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    c2w_R = np.stack([right, down, fwd], axis=1)  # columns = camera axes in world
    T = -c2w_R.T @ eye                            # world->camera translation
    return make_camera(c2w_R, T, fovx, fovy, H, W)


def orbit_camera(k: int, H: int = 800, W: int = 800, camera_angle_x: float = 0.6911112, radius: float = 3.359,
                 elevation_deg: float = 30.0, fovy: float | None = None) -> SynthCamera:
    """Blender-lego-like rig: azimuth 45deg*k, elevation 30deg, radius 3.359."""
    az = math.radians(45.0 * k)
    el = math.radians(elevation_deg)
    eye = radius * np.array([math.cos(el) * math.cos(az), math.cos(el) * math.sin(az), math.sin(el)])
    fovx = camera_angle_x
    if fovy is None:
        fovy = focal2fov(fov2focal(fovx, W), H)
    return look_at_camera(eye, np.zeros(3), fovx, fovy, H, W)


def make_scene(P: int, seed: int, sh_coeffs: int = 16, scale_mult: float = 1.0, extent: float = 1.3,
               precomp_rgb: bool = False) -> dict:
    """Seeded splats with the BASELINE.md §2.3 distributions (CPU float32 tensors)."""
    g = torch.Generator().manual_seed(seed)
    means3D = (torch.rand(P, 3, generator=g) * 2 - 1) * extent
    lo, hi = math.log(0.005), math.log(0.03)
    scales = torch.exp(torch.rand(P, 3, generator=g) * (hi - lo) + lo) * scale_mult
    q = torch.randn(P, 4, generator=g)
    rotations = q / q.norm(dim=1, keepdim=True)
    opacities = torch.sigmoid(torch.randn(P, 1, generator=g) * 1.5)
    out = dict(means3D=means3D, scales=scales, rotations=rotations, opacities=opacities)
    if precomp_rgb:
        out["colors_precomp"] = torch.rand(P, 3, generator=g)
    else:
        shs = torch.randn(P, sh_coeffs, 3, generator=g) * 0.05
        shs[:, 0, :] = torch.randn(P, 3, generator=g) * 0.5
        out["shs"] = shs
    return out


# The BASELINE.json configs as (P, H, W, seed, scale_mult, precomp_rgb, camera kwargs)
CONFIGS = {
    "plumbing_256": dict(P=256, H=64, W=64, seed=0, scale_mult=1.0, precomp_rgb=False),
    "lego_100k": dict(P=100_000, H=800, W=800, seed=1, scale_mult=1.0, precomp_rgb=False),
    "lego_1m": dict(P=1_000_000, H=800, W=800, seed=2, scale_mult=1.0, precomp_rgb=False),
    "dtu_500k": dict(P=500_000, H=1200, W=1600, seed=3, scale_mult=2.0, precomp_rgb=False),
    "owlii_2m": dict(P=2_000_000, H=1080, W=1920, seed=4, scale_mult=1.0, precomp_rgb=True),
}


def config_camera(name: str, k: int = 0) -> SynthCamera:
    c = CONFIGS[name]
    if name == "plumbing_256":
        # R = I, T = (0,0,4), FoVx = FoVy = 0.69
        return make_camera(np.eye(3), np.array([0.0, 0.0, 4.0]), 0.69, 0.69, c["H"], c["W"])
    if name == "dtu_500k":
        fovx = focal2fov(2892.0, 1600)
        fovy = focal2fov(2892.0, 1200)
        return orbit_camera(k, c["H"], c["W"], camera_angle_x=fovx, fovy=fovy)
    return orbit_camera(k, c["H"], c["W"])


def frame_offset(P: int, t: float, seed: int = 4, amplitude: float = 0.05) -> torch.Tensor:
    """Deterministic per-frame displacement of the means for the 4D configuration (BASELINE.json configs[4],
    SURVEY.md §8d): offset_i(t) = amplitude * sin(2 pi t + phi_i), phi_i ~ U(0, 2 pi) per coordinate (seeded).
    Stands in for the deformation network of scene/deform_model.py, which is the reference's model, not this path."""
    g = torch.Generator().manual_seed(10_000 + seed)
    phi = torch.rand(P, 3, generator=g) * (2.0 * math.pi)
    return amplitude * torch.sin(2.0 * math.pi * float(t) + phi)


def view_time_jobs(n_frames: int, n_views: int, rank: int = 0, world: int = 1) -> list:
    """(frame, view) jobs of one rank: job j = frame * n_views + view goes to rank j % world (round-robin).
    With world ranks the jobs of round r are j = r * world .. r * world + world - 1; a rank without a job in the
    last round gets None there (it still takes part in that round's gradient exchange, contributing zeros)."""
    total = n_frames * n_views
    rounds = (total + world - 1) // world
    out = []
    for r in range(rounds):
        j = r * world + rank
        out.append((j // n_views, j % n_views) if j < total else None)
    return out
