"""How far inside the parity tolerances (1e-5 abs image/depth, 1e-3 rel gradients) does the loaded library variant sit?
Prints the worst errors against the CPU oracle on BASELINE config 1 (lego_100k) and a dense small scene.
(development aid: decides whether the fast-exp build variant is safe to ship)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import oracle as O
from splatfields_b200 import synth
from tests.helpers import run_cuda, run_oracle

O.build()
for name, P, H, W, seed, sm in (("lego_100k", 100_000, 800, 800, 1, 1.0), ("dense_20k", 20_000, 256, 256, 33, 3.0)):
    sc = synth.make_scene(P, seed, scale_mult=sm)
    cam = synth.orbit_camera(0, H, W)
    dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(1)).numpy()
    f, _ = run_oracle(O, sc, cam, H, W, (1, 1, 1), 3)
    ok = f["margin"] > 1e-4
    dL[:, ~ok] = 0
    f, b = run_oracle(O, sc, cam, H, W, (1, 1, 1), 3, dL=dL)
    c, g = run_cuda(sc, cam, H, W, (1, 1, 1), 3, dL=dL)
    res = dict(case=name, variant=os.environ.get("SFB_LIB_VARIANT", "default"),
               keys_equal=bool(np.array_equal(c["point_list_keys"], f["point_list_keys"])),
               max_abs_rgb=float(np.abs(c["color"] - f["color"])[:, ok].max()),
               max_abs_depth=float(np.abs(c["depth"][0] - f["depth"].reshape(H, W))[ok].max()),
               pixels_excluded=int((~ok).sum()))
    for k in g:
        ref = b[k].reshape(g[k].shape).astype(np.float64)
        res["normrel_" + k] = float(np.linalg.norm(g[k] - ref) / max(np.linalg.norm(ref), 1e-30))
    print(json.dumps(res), flush=True)
