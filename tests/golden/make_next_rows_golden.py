"""Generate tests/golden/next_rows.npz by RUNNING the reference's own Python code (read-only /root/reference) for
the SURVEY.md §8f rows: the photometric loss and the densification bookkeeping.  Authoring container only
(`python tests/golden/make_next_rows_golden.py`); the GPU box reads the committed .npz.

Pinned here:
  * l1_loss, ssim (and autograd through them, fp64 and fp32)        utils/loss_utils.py:18-19, :33-76; train.py:183-184
  * mask term  F.l1_loss(clamp(opacity, 0, 1), gt_mask)              train.py:189-193
  * GaussianModel.add_densification_stats + max_radii2D update       scene/gaussian_model.py:427-430; train.py:280-282
  * GaussianModel.densify_and_prune selection masks                  scene/gaussian_model.py:355-425
    (the real methods run; densification_postfix / prune_points are intercepted to record the masks)
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
sys.path.insert(0, REF)

# absent third-party modules the imports of scene/gaussian_model.py drag in (never called here)
for name in ("trimesh", "plyfile", "simple_knn", "simple_knn._C"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["plyfile"].PlyData = sys.modules["plyfile"].PlyElement = object
sys.modules["simple_knn._C"].distCUDA2 = None

# the reference hard-codes device="cuda" in a few constructors; run them on CPU
_zeros = torch.zeros


def _cpu_zeros(*a, **k):
    k.pop("device", None)
    return _zeros(*a, **k)


torch.zeros = _cpu_zeros

from utils import loss_utils as LU            # noqa: E402
import importlib.util                          # noqa: E402

spec = importlib.util.spec_from_file_location("ref_gaussian_model", os.path.join(REF, "scene", "gaussian_model.py"))
GM = importlib.util.module_from_spec(spec)
spec.loader.exec_module(GM)


def loss_case(key, C, H, W, lam, seed, dtype, with_mask=False, lam_mask=0.1, out=None):
    g = torch.Generator().manual_seed(seed)
    # a smooth-ish "render" and a perturbed "ground truth", plus flat regions (the ill-conditioned case for SSIM)
    base = torch.rand(C, H, W, generator=g, dtype=torch.float64)
    img = (0.6 * base + 0.2).clone()
    gt = (img + 0.1 * torch.randn(C, H, W, generator=g, dtype=torch.float64)).clamp(0, 1)
    if H >= 12 and W >= 12:
        img[:, : H // 3, : W // 3] = 1.0      # white background block in both
        gt[:, : H // 3, : W // 3] = 1.0
        gt[:, -3:, :] = img[:, -3:, :]        # exact ties: sign(0) = 0
    img = img.to(dtype).requires_grad_(True)
    gt = gt.to(dtype)
    l1 = LU.l1_loss(img, gt)
    ss = LU.ssim(img, gt)
    loss = (1.0 - lam) * l1 + lam * (1.0 - ss)
    res = {"img": img.detach(), "gt": gt, "l1": l1.detach(), "ssim": ss.detach()}
    if with_mask:
        op = (torch.rand(1, H, W, generator=g, dtype=torch.float64) * 1.4 - 0.2).to(dtype).requires_grad_(True)
        mask = (torch.rand(1, H, W, generator=g, dtype=torch.float64) > 0.5).to(dtype)
        opacity_image = torch.clamp(op, 0.0, 1.0)
        lm = F.l1_loss(opacity_image.view(-1), mask.view(-1))
        loss = loss + lam_mask * lm
        res.update({"opacity": op.detach(), "mask": mask, "mask_l1": lm.detach()})
    loss.backward()
    res["loss"] = loss.detach()
    res["dL_dimg"] = img.grad
    if with_mask:
        res["dL_dopacity"] = op.grad
    for k, v in res.items():
        out[f"{key}.{k}"] = v.numpy()
    out[f"{key}.lambda"] = np.float64(lam)
    out[f"{key}.lambda_mask"] = np.float64(lam_mask if with_mask else 0.0)


class Probe(GM.GaussianModel):
    """The reference model with the optimizer surgery intercepted: records what the selection code selected."""

    def densification_postfix(self, new_xyz, *rest):
        self.rec.append(("postfix", new_xyz.detach().clone()))

    def prune_points(self, mask):
        self.rec.append(("prune", mask.detach().clone()))


def densify_case(key, P, seed, max_screen_size, out):
    g = torch.Generator().manual_seed(seed)
    m = Probe(3)
    m.rec = []
    m.percent_dense = 0.01
    extent = 4.7
    ids = torch.arange(P, dtype=torch.float32)
    m._xyz = torch.stack([ids, torch.zeros(P), torch.zeros(P)], dim=1)       # row id rides in x: masks are recoverable
    m._features_dc = torch.zeros(P, 1, 3)
    m._features_rest = torch.zeros(P, 15, 3)
    m._scaling = torch.log(torch.exp(torch.rand(P, 3, generator=g) * 6.0 - 7.0))     # raw = log(scale), scale in e^[-7,-1]
    m._rotation = F.normalize(torch.randn(P, 4, generator=g))
    m._opacity = torch.randn(P, 1, generator=g) * 3.0
    m.xyz_gradient_accum = torch.zeros(P, 1)
    m.denom = torch.zeros(P, 1)
    m.max_radii2D = torch.zeros(P)
    out[f"{key}.raw_scaling"] = m._scaling.numpy().copy()
    out[f"{key}.raw_opacity"] = m._opacity.numpy().copy()
    # three views of statistics (train.py:280-282 + :306-307)
    for v in range(3):
        radii = (torch.rand(P, generator=g) * 40 - 8).clamp(min=0).to(torch.int32)
        vis = radii > 0
        vsp = types.SimpleNamespace(grad=torch.randn(P, 3, generator=g) * 3e-4)
        m.max_radii2D[vis] = torch.max(m.max_radii2D[vis], radii[vis])
        m.add_densification_stats(vsp, vis)
        out[f"{key}.view{v}.radii"] = radii.numpy()
        out[f"{key}.view{v}.grad"] = vsp.grad.numpy()
        out[f"{key}.view{v}.accum"] = m.xyz_gradient_accum.numpy().copy()
        out[f"{key}.view{v}.denom"] = m.denom.numpy().copy()
        out[f"{key}.view{v}.max_radii2D"] = m.max_radii2D.numpy().copy()
    torch.manual_seed(seed)       # densify_and_split draws torch.normal samples
    m.densify_and_prune(0.0002, 0.005, extent, max_screen_size)
    kinds = [k for k, _ in m.rec]
    assert kinds == ["postfix", "postfix", "prune", "prune"], kinds
    clone = torch.zeros(P, dtype=torch.bool)
    clone[m.rec[0][1][:, 0].long()] = True
    split = m.rec[2][1][:P]
    prune = m.rec[3][1]
    out[f"{key}.clone"] = clone.numpy()
    out[f"{key}.split"] = split.numpy()
    out[f"{key}.prune"] = prune.numpy()
    out[f"{key}.params"] = np.array([0.0002, 0.01, extent, 0.005, float(max_screen_size or 0)], dtype=np.float64)


def main():
    out = {}
    out["window_1d"] = LU.gaussian(11, 1.5).numpy()
    out["window_2d"] = LU.create_window(11, 1)[0, 0].numpy()
    loss_case("a64", 3, 40, 56, 0.2, 1, torch.float64, out=out)
    loss_case("b64", 3, 16, 16, 0.2, 2, torch.float64, with_mask=True, out=out)
    loss_case("c64", 1, 23, 31, 1.0, 3, torch.float64, out=out)
    loss_case("d64", 3, 9, 7, 0.5, 4, torch.float64, out=out)          # smaller than the window
    loss_case("e64", 2, 33, 17, 0.0, 5, torch.float64, with_mask=True, lam_mask=0.25, out=out)   # l1 only
    loss_case("a32", 3, 40, 56, 0.2, 1, torch.float32, out=out)        # what the reference computes in practice
    densify_case("dn0", 5000, 11, 20, out)
    densify_case("dn1", 777, 12, None, out)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "next_rows.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, f"{os.path.getsize(dst) / 1024:.0f} KiB,", len(out), "arrays")


if __name__ == "__main__":
    main()
