"""A/B of the opt-in SFB_TIGHT_RECT=1 build path (preprocess clips each tile rectangle to the splat's alpha >= 1/255
footprint box; DESIGN.md §4 "next"): the image, depth and every gradient must not change, R and the step time
should drop.  The knob is read once per process, so each arm runs in its own child process.
    python scripts/ab_tight_rect.py [--config lego_1m]"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(config, out):
    import numpy as np
    import torch
    from splatfields_b200 import synth
    from tests.helpers import run_cuda
    cfg = synth.CONFIGS[config]
    sc = synth.make_scene(cfg["P"], cfg["seed"], scale_mult=cfg["scale_mult"], precomp_rgb=cfg["precomp_rgb"])
    cam = synth.config_camera(config, 0)
    H, W = cfg["H"], cfg["W"]
    dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(1)).numpy()
    deg = 0 if cfg["precomp_rgb"] else 3
    c, g = run_cuda(sc, cam, H, W, (1, 1, 1), deg, dL=dL)
    # timing: forward + backward through the autograd surface, CUDA events
    from splatfields_b200.host_api import ViewParallelRasterizer
    vp = ViewParallelRasterizer(sc, cam, H, W, deg, device="cuda")
    G = torch.as_tensor(dL).cuda()
    for _ in range(5):
        vp.step(G)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        vp.step(G)
    e1.record()
    torch.cuda.synchronize()
    np.savez(out, color=c["color"], depth=c["depth"], radii=c["radii"], R=c["num_rendered"],
             ms=e0.elapsed_time(e1) / 50, **{"g_" + k: v for k, v in g.items()})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="lego_1m")
    ap.add_argument("--child", default=None)
    a = ap.parse_args()
    if a.child:
        return child(a.config, a.child)
    import numpy as np
    res = {}
    with tempfile.TemporaryDirectory() as td:
        for arm, env in (("default", {}), ("tight", {"SFB_TIGHT_RECT": "1"})):
            out = os.path.join(td, arm + ".npz")
            subprocess.check_call([sys.executable, __file__, "--config", a.config, "--child", out],
                                  env={**os.environ, **env})
            res[arm] = dict(np.load(out))
    d, t = res["default"], res["tight"]
    line = {"config": a.config, "R_default": int(d["R"]), "R_tight": int(t["R"]), "ms_default": float(d["ms"]),
            "ms_tight": float(t["ms"]), "radii_equal": bool(np.array_equal(d["radii"], t["radii"])),
            "max_abs_color": float(np.abs(d["color"] - t["color"]).max()),
            "max_abs_depth": float(np.abs(d["depth"] - t["depth"]).max())}
    for k in d:
        if k.startswith("g_"):
            ref = d[k].astype(np.float64)
            line["normrel_" + k[2:]] = float(np.linalg.norm(t[k] - ref) / max(np.linalg.norm(ref), 1e-30))
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
