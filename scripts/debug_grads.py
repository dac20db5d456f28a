"""Per-gradient error report of the CUDA path against the oracle on a small scene (debugging aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.helpers import run_oracle, run_cuda, scene_and_camera
from oracle import oracle as O_mod


def main():
    print("start", flush=True)
    P, H, W, deg = int(os.environ.get("DBG_P", 4000)), 96, 128, 3
    sc, cam = scene_and_camera(P, H, W, 7, sh_degree=deg)
    O_mod.build()
    O = O_mod
    rng = np.random.default_rng(3)
    dL = rng.standard_normal((3, H, W)).astype(np.float32)
    bg = (0.3, 0.6, 0.1)
    f, b = run_oracle(O, sc, cam, H, W, bg, deg, dL=dL)
    print("oracle done", flush=True)
    c, g = run_cuda(sc, cam, H, W, bg, deg, dL=dL)
    print("R", c["num_rendered"], "img err", float(np.abs(c["color"] - f["color"]).max()))
    for k in g:
        got = np.asarray(g[k], np.float64)
        ref = np.asarray(b[k], np.float64).reshape(got.shape)
        nan = int(np.isnan(got).sum())
        gz = np.nan_to_num(got)
        print(f"{k:16s} nan={nan:6d} normrel={np.linalg.norm(gz - ref) / max(np.linalg.norm(ref), 1e-30):.3e} "
              f"|ref|={np.linalg.norm(ref):.3e} |got|={np.linalg.norm(gz):.3e}")


main()
if os.environ.get("SFB_MMA_DBG"):
    import ctypes as C
    from splatfields_b200 import _lib
    lib = _lib.load()
    if hasattr(lib, "sfb_debug_read"):
        buf = (C.c_float * 640)()
        print("debug_read rc", lib.sfb_debug_read(buf))
        a = np.array(buf[:]).reshape(32, 20)
        np.set_printoptions(linewidth=250, precision=5, suppress=False)
        print("cols: top nsweep gn st.x st.y dS0 dS1 dS2 dS3 dW0 dW1 dW2 dW3 dirS0 dirSX dirC0 bS0 dlp list dLp0")
        print(a)
