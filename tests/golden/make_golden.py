"""Generate tests/golden/reference_helpers.npz by IMPORTING the reference's own Python helpers
(/root/reference, read-only) — the only executable statement of the hot path's conventions that the
reference ships (its rasterizer is an un-vendored CUDA dependency with no tests).

Run in the authoring container only (`python tests/golden/make_golden.py`); the GPU box has no
/root/reference and only reads the committed .npz.

Pinned here:
  * getWorld2View2 / getProjectionMatrix / Camera matrices        utils/graphics_utils.py:42-76, scene/cameras.py:62-74
  * geom_transform_points (row-vector, w + 1e-7)                 utils/graphics_utils.py:24-31
  * eval_sh degrees 0..3                                          utils/sh_utils.py:57-112
  * build_rotation / build_scaling_rotation / strip_symmetric     utils/general_utils.py:122-171
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
from utils import graphics_utils as GU   # noqa: E402
from utils import sh_utils as SH         # noqa: E402

# utils/general_utils.py imports trimesh (absent here) and hard-codes device="cuda" in the three
# helpers we need; stub the import and strip the device keyword so the reference code runs on CPU.
sys.modules.setdefault("trimesh", types.ModuleType("trimesh"))
from utils import general_utils as GEN   # noqa: E402

_zeros = torch.zeros


def _cpu_zeros(*a, **k):
    k.pop("device", None)
    return _zeros(*a, **k)


def load_camera_cls():
    spec = importlib.util.spec_from_file_location("ref_cameras", os.path.join(REF, "scene", "cameras.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.Camera


def main():
    g = torch.Generator().manual_seed(1234)
    out = {}
    Camera = load_camera_cls()
    cams = []
    for k in range(4):
        # random rotation (QR) + translation, OpenCV-ish
        A = torch.randn(3, 3, generator=g, dtype=torch.float64).numpy()
        Q, _ = np.linalg.qr(A)
        if np.linalg.det(Q) < 0:
            Q[:, 0] *= -1
        T = np.array([0.1 * k, -0.2, 3.0 + k])
        fovx, fovy = 0.6 + 0.1 * k, 0.5 + 0.07 * k
        cam = Camera(colmap_id=k, R=Q, T=T, FoVx=fovx, FoVy=fovy, image=None, gt_alpha_mask=None,
                     image_name=str(k), uid=k, data_device="cpu", fid=0.0, image_width=64 + 16 * k,
                     image_height=48 + 16 * k)
        cams.append(dict(R=Q, T=T, fovx=fovx, fovy=fovy, W=64 + 16 * k, H=48 + 16 * k,
                         wvt=cam.world_view_transform.numpy(), proj=cam.projection_matrix.numpy(),
                         full=cam.full_proj_transform.numpy(), center=cam.camera_center.numpy()))
    for k, c in enumerate(cams):
        for name, v in c.items():
            out[f"cam{k}_{name}"] = np.asarray(v)
    out["w2v2"] = GU.getWorld2View2(cams[0]["R"], cams[0]["T"])
    out["projm"] = GU.getProjectionMatrix(0.01, 100.0, 0.7, 0.6).numpy()

    # points through the full projection (NDC)
    pts = (torch.rand(64, 3, generator=g) * 2 - 1)
    out["pts"] = pts.numpy()
    for k in range(4):
        out[f"ndc{k}"] = GU.geom_transform_points(pts, torch.tensor(cams[k]["full"])).numpy()
        out[f"view{k}"] = GU.geom_transform_points(pts, torch.tensor(cams[k]["wvt"])).numpy()

    # SH: reference layout for eval_sh is [..., C, coeffs]; the rasterizer takes [P, coeffs, C]
    shs = torch.randn(64, 16, 3, generator=g) * 0.3
    dirs = torch.randn(64, 3, generator=g)
    dirs = dirs / dirs.norm(dim=1, keepdim=True)
    out["shs"] = shs.numpy()
    out["dirs"] = dirs.numpy()
    for deg in range(4):
        out[f"sh_rgb{deg}"] = SH.eval_sh(deg, shs.transpose(1, 2), dirs).numpy()

    # covariance
    scales = torch.exp(torch.randn(64, 3, generator=g) * 0.5 - 3.0)
    q = torch.randn(64, 4, generator=g)
    torch.zeros = _cpu_zeros
    try:
        R = GEN.build_rotation(q)
        L = GEN.build_scaling_rotation(1.0 * scales, q)
        cov6 = GEN.strip_symmetric(L @ L.transpose(1, 2))
    finally:
        torch.zeros = _zeros
    out["scales"] = scales.numpy()
    out["quats"] = q.numpy()
    out["rotmats"] = R.numpy()
    out["cov6"] = cov6.numpy()

    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_helpers.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
