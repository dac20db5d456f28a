"""Generate tests/golden/activations.npz by RUNNING the reference's own GaussianModel getters (read-only
/root/reference) for SURVEY.md §8f-3: get_scaling / get_rotation / get_opacity / get_features
(scene/gaussian_model.py:64-86) as packed by get_gaussian_dict's static branch (train.py:42-50), forward and — through
autograd with fixed random cotangents — backward.  Authoring container only
(`python tests/golden/make_activations_golden.py`); the GPU box reads the committed .npz."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
for name in ("trimesh", "plyfile", "simple_knn", "simple_knn._C"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["plyfile"].PlyData = sys.modules["plyfile"].PlyElement = object
sys.modules["simple_knn._C"].distCUDA2 = None
_zeros = torch.zeros


def _cpu_zeros(*a, **k):
    k.pop("device", None)
    return _zeros(*a, **k)


torch.zeros = _cpu_zeros
spec = importlib.util.spec_from_file_location("ref_gaussian_model", os.path.join(REF, "scene", "gaussian_model.py"))
GM = importlib.util.module_from_spec(spec)
spec.loader.exec_module(GM)


def case(key, P, M, seed, iso, dtype, out):
    g = torch.Generator().manual_seed(seed)
    m = GM.GaussianModel(3)
    m.use_isotropic = iso
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    raw = {
        "xyz": r(P, 3),
        "scaling": r(P, 1 if iso else 3) * 1.5 - 4.0,
        "rotation": r(P, 4) * torch.exp(r(P, 1) * 2.0),        # norms spread over several decades
        "opacity": r(P, 1) * 4.0,                              # saturating sigmoids on both sides
        "f_dc": r(P, 1, 3),
        "f_rest": r(P, M - 1, 3) * 0.1,
    }
    if P >= 4:
        raw["rotation"][1] = 0.0                               # |x| < eps: F.normalize divides by eps
        raw["rotation"][2] = torch.tensor([1e-14, 0.0, 0.0, 0.0], dtype=torch.float64)
    raw = {k: v.to(dtype).requires_grad_(True) for k, v in raw.items()}
    m._xyz, m._scaling, m._rotation, m._opacity = raw["xyz"], raw["scaling"], raw["rotation"], raw["opacity"]
    m._features_dc, m._features_rest = raw["f_dc"], raw["f_rest"]
    # train.py:42-50
    d = {"gaussian_opacity": m.get_opacity, "gaussian_features": m.get_features, "gaussian_scales": m.get_scaling,
         "gaussian_rotations": m.get_rotation}
    cot = {k: torch.randn(v.shape, generator=g, dtype=torch.float64).to(dtype) for k, v in d.items()}
    sum((v * cot[k]).sum() for k, v in d.items()).backward()
    for k, v in raw.items():
        out[f"{key}.raw.{k}"] = v.detach().numpy()
        if k != "xyz":
            out[f"{key}.grad.{k}"] = v.grad.numpy()
    for k, v in d.items():
        out[f"{key}.out.{k}"] = v.detach().numpy()
        out[f"{key}.cot.{k}"] = cot[k].numpy()
    out[f"{key}.iso"] = np.int32(iso)


def main():
    out = {}
    case("a64", 96, 16, 1, False, torch.float64, out)
    case("a32", 96, 16, 1, False, torch.float32, out)
    case("iso32", 65, 16, 2, True, torch.float32, out)
    case("m4_32", 67, 4, 3, False, torch.float32, out)      # max_sh_degree 1: rows of 12 floats, odd tail
    case("m1_32", 33, 1, 4, False, torch.float32, out)       # DC only: empty f_rest
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "activations.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, f"{os.path.getsize(dst) / 1024:.0f} KiB,", len(out), "arrays")


if __name__ == "__main__":
    main()
