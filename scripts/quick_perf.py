"""Quick device-side timing of one config (development aid; bench.py is the contract harness)."""
import argparse
import json
import math
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from splatfields_b200 import _lib, synth
from splatfields_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="lego_1m")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--cam", type=int, default=0)
    a = ap.parse_args()
    cfg = synth.CONFIGS[a.config]
    dev = torch.device("cuda:0")
    sc = synth.make_scene(cfg["P"], cfg["seed"], scale_mult=cfg["scale_mult"], precomp_rgb=cfg["precomp_rgb"])
    cam = synth.config_camera(a.config, a.cam).to(dev)
    H, W = cfg["H"], cfg["W"]
    t = {k: v.to(dev).requires_grad_(True) for k, v in sc.items()}
    deg = 0 if cfg["precomp_rgb"] else 3
    rs = GaussianRasterizationSettings(H, W, math.tan(cam.FoVx / 2), math.tan(cam.FoVy / 2),
                                       torch.ones(3, device=dev), 1.0, cam.world_view_transform,
                                       cam.full_proj_transform, deg, cam.camera_center, False, False)
    rast = GaussianRasterizer(rs)
    G = torch.randn(3, H, W, device=dev)
    m2d = torch.zeros_like(t["means3D"], requires_grad=True)

    def step():
        for v in t.values():
            v.grad = None
        color, radii, depth = rast(means3D=t["means3D"], means2D=m2d, opacities=t["opacities"], shs=t.get("shs"),
                                   colors_precomp=t.get("colors_precomp"), scales=t["scales"], rotations=t["rotations"])
        return color, radii

    ef = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf, tb = [], []
    for i in range(a.warmup + a.iters):
        ef[0].record()
        color, radii = step()
        ef[1].record()
        color.backward(G)
        ef[2].record()
        torch.cuda.synchronize()
        if i >= a.warmup:
            tf.append(ef[0].elapsed_time(ef[1]))
            tb.append(ef[1].elapsed_time(ef[2]))
    R = color.grad_fn.num_rendered if color.grad_fn is not None else -1
    _lib.profile_enable(True)
    color, radii = step()
    color.backward(G)
    torch.cuda.synchronize()
    def agg(rows):      # several launches share a name (the sort passes): report the per-step sum
        d = {}
        for k, v in rows:
            d[k] = d.get(k, 0.0) + v
        return d
    pf, pb = agg(_lib.profile_read(0)), agg(_lib.profile_read(1))
    _lib.profile_enable(False)
    # render(): the reference's two rasterizer passes vs the fused alpha channel
    from splatfields_b200 import render
    gd = dict(means3D=t["means3D"], active_sh_degree=deg, gaussian_opacity=t["opacities"], gaussian_scales=t["scales"],
              gaussian_rotations=t["rotations"])
    if "shs" in t:
        gd["gaussian_features"] = t["shs"]
    else:
        gd["gaussian_rgb"] = t["colors_precomp"]
    Ga = torch.randn(1, H, W, device=dev)
    rtimes = {}
    for fused in (False, True):
        ts = []
        for i in range(a.warmup + a.iters):
            for v in t.values():
                v.grad = None
            ef[0].record()
            o = render(cam, gd, None, torch.ones(3, device=dev), return_opacity=True, fused_alpha=fused)
            ((o["render"] * G).sum() + (o["opacity"] * Ga).sum()).backward()
            ef[2].record()
            torch.cuda.synchronize()
            if i >= a.warmup:
                ts.append(ef[0].elapsed_time(ef[2]))
        rtimes["fused" if fused else "two_pass"] = float(np.median(ts))
    med = lambda x: float(np.median(x))
    T = ((W + 15) // 16) * ((H + 15) // 16)
    res = dict(config=a.config, P=cfg["P"], H=H, W=W, R=R, visible=int((radii > 0).sum()), mean_list=R / T,
               fwd_ms=med(tf), bwd_ms=med(tb), total_ms=med(tf) + med(tb),
               msplats_s=cfg["P"] / (med(tf) + med(tb)) / 1e3, render_with_opacity_ms=rtimes, fwd_stages=pf,
               bwd_stages=pb)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
