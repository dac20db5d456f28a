"""Small invocations of every kernel added after the last sanitizer pass (activate, knn, sh_grad_combine, factored
geometry backward) for `compute-sanitizer --tool memcheck|racecheck python scripts/sanitize_new_rows.py`."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from splatfields_b200 import _lib, activate_parameters, distCUDA2, rasterizer, synth
from splatfields_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    P, M = 3001, 16
    mk = lambda *s: torch.randn(*s, generator=g).to(dev).requires_grad_(True)
    xyz, rs, rr, ro, dc, rest = mk(P, 3), mk(P, 3), mk(P, 4), mk(P, 1), mk(P, 1, 3), mk(P, M - 1, 3)
    d = activate_parameters(xyz, rs, rr, ro, dc, rest)
    sum(v.sum() for k, v in d.items() if k.startswith("gaussian_")).backward()
    for n in (1, 5, 33, 4097):
        distCUDA2(torch.randn(n, 3, generator=g).to(dev))
    H, W = 96, 128
    sc = synth.make_scene(P, 3, scale_mult=3.0)
    cam = synth.orbit_camera(0, H, W).to(dev)
    t = {k: v.to(dev).requires_grad_(True) for k, v in sc.items()}
    rset = GaussianRasterizationSettings(H, W, math.tan(cam.FoVx / 2), math.tan(cam.FoVy / 2), torch.ones(3, device=dev),
                                         1.0, cam.world_view_transform, cam.full_proj_transform, 3, cam.camera_center,
                                         False, False)
    lib = _lib.load()
    for _ in range(1):
        color, radii, depth = GaussianRasterizer(rset)(means3D=t["means3D"], means2D=torch.zeros(P, 3, device=dev,
                                                       requires_grad=True), opacities=t["opacities"], shs=t["shs"],
                                                       scales=t["scales"], rotations=t["rotations"])
        dcol = torch.empty(P, 3, device=dev)
        rasterizer.set_grad_arena(None, None, dcol)
        color.sum().backward()
        rasterizer.set_grad_arena(None, None)
        out = torch.empty(P, 16, 3, device=dev)
        rasterizer.sh_grad_combine(t["means3D"].detach(), cam.camera_center.reshape(1, 3).contiguous(),
                                   dcol.reshape(1, P, 3), 3, out)
    torch.cuda.synchronize()
    print("sanitize_new_rows: done")


if __name__ == "__main__":
    main()
