"""Factored exchange of the SH gradient (include/splat_b200.h SFB_BWD_SH_FACTORED,
sfb_sh_grad_combine).

CPU: the identity behind it — per view, dL_dsh is the rank-1 block basis(dir) (x) (clamp-masked dL_dcolour) — checked
on the pinned C oracle's own backward, over several views (= the accumulation of the serial loop at train.py:169-242).
GPU: the CUDA path in factored mode + sfb_sh_grad_combine against its own direct dL_dsh rows and against the oracle."""
import math

import numpy as np
import pytest
import torch

from oracle import torch_naive as TN
from splatfields_b200 import synth
from tests.helpers import run_oracle


def _views(P, H, W, seed, V):
    sc = synth.make_scene(P, seed, scale_mult=3.0)
    sc["shs"][::7, 0, 1] = -3.0               # SH colour below 0 in one channel: the max(0, .) clamp is exercised
    sc["shs"][::11, 0, :] = -2.5
    cams = [synth.orbit_camera(k, H, W) for k in range(V)]
    Gs = [torch.randn(3, H, W, generator=torch.Generator().manual_seed(50 + k)) for k in range(V)]
    return sc, cams, Gs


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_oracle_sh_gradient_is_rank1_per_view(oracle, deg):
    P, H, W, V = 400, 48, 64, 3
    sc, cams, Gs = _views(P, H, W, 11, V)
    total = np.zeros((P, 16, 3), np.float64)
    masked, campos = [], []
    n_clamped = 0
    for cam, G in zip(cams, Gs):
        f, b = run_oracle(oracle, sc, cam, H, W, (1.0, 1.0, 1.0), deg, dL=G.numpy(), want_margin=False)
        total += b["dL_dsh"].astype(np.float64)
        cl = f["clamped"].reshape(P, 3).astype(bool)
        n_clamped += int(cl.sum())
        g = b["dL_dcolors"].astype(np.float64).copy()
        g[cl] = 0.0
        g[f["radii"] <= 0] = 0.0
        masked.append(g)
        campos.append(cam.camera_center.numpy())
    assert n_clamped > 0                      # the clamp mask is exercised
    ref = TN.sh_grad_combine_ref(sc["means3D"], np.stack(campos), np.stack(masked), deg, 16).numpy()
    scale = np.abs(total).max()
    assert scale > 0
    assert np.abs(ref - total).max() <= 2e-6 * scale
    assert np.all(ref[:, (deg + 1) ** 2:, :] == 0)


# ------------------------------------------------------------------------------------------------ GPU
def _cuda_backward(sc, cam, H, W, deg, G, factored, dev):
    from splatfields_b200 import rasterizer
    from splatfields_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer
    t = {k: v.to(dev).clone().requires_grad_(True) for k, v in sc.items()}
    camd = cam.to(dev)
    rs = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5),
        bg=torch.ones(3, device=dev), scale_modifier=1.0, viewmatrix=camd.world_view_transform,
        projmatrix=camd.full_proj_transform, sh_degree=deg, campos=camd.camera_center, prefiltered=False, debug=False)
    means2D = torch.zeros_like(t["means3D"], requires_grad=True)
    color, radii, depth = GaussianRasterizer(rs)(means3D=t["means3D"], means2D=means2D, opacities=t["opacities"],
                                                 shs=t["shs"], scales=t["scales"], rotations=t["rotations"])
    P = t["means3D"].shape[0]
    dcol = torch.full((P, 3), float("nan"), device=dev) if factored else None
    if factored:
        rasterizer.set_grad_arena(None, None, dcol)
    try:
        (color * G.to(dev)).sum().backward()
    finally:
        rasterizer.set_grad_arena(None, None)
    torch.cuda.synchronize()
    grads = {k: (None if v.grad is None else v.grad.detach().clone()) for k, v in t.items()}
    grads["means2D"] = means2D.grad.detach().clone()
    return grads, dcol, camd.camera_center.detach().reshape(3).clone(), t["means3D"].detach()


@pytest.mark.gpu
@pytest.mark.parametrize("deg,P", [(3, 20000), (2, 5000), (1, 5000), (0, 5000)])
def test_factored_backward_single_view_matches_direct(cuda_lib, deg, P):
    from splatfields_b200 import rasterizer
    dev = torch.device("cuda")
    H, W = 160, 208
    sc, cams, Gs = _views(P, H, W, 21, 1)
    direct, _, _, _ = _cuda_backward(sc, cams[0], H, W, deg, Gs[0], False, dev)
    fact, dcol, campos, means = _cuda_backward(sc, cams[0], H, W, deg, Gs[0], True, dev)
    assert fact["shs"] is None                          # autograd hands out no SH rows in factored mode
    assert not torch.isnan(dcol).any()                  # every row of the colour gradient was written
    for k in ("means3D", "opacities", "scales", "rotations", "means2D"):
        # same kernels, same inputs; only the atomics' arrival order differs between two runs
        a, b = fact[k], direct[k]
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6 * float(b.abs().max())), k
    out = torch.full((P, 16, 3), float("nan"), device=dev)
    rasterizer.sh_grad_combine(means, campos.reshape(1, 3).contiguous(), dcol.reshape(1, P, 3), deg, out)
    torch.cuda.synchronize()
    ref = direct["shs"]
    assert not torch.isnan(out).any()
    assert torch.all(out[:, (deg + 1) ** 2:, :] == 0)
    scale = float(ref.abs().max())
    assert scale > 0
    assert float((out - ref).abs().max()) <= 1e-4 * scale


@pytest.mark.gpu
def test_sh_grad_combine_multi_view_sum(cuda_lib):
    """V views rendered one after the other on one device: combine(factored colour gradients) == sum of the direct
    per-view dL_dsh rows (what V ranks + all-reduce would produce), and == the fp64 restatement."""
    from splatfields_b200 import rasterizer
    dev = torch.device("cuda")
    P, H, W, V, deg = 30000, 200, 200, 5, 3
    sc, cams, Gs = _views(P, H, W, 22, V)
    total = torch.zeros(P, 16, 3, device=dev, dtype=torch.float64)
    dcols, campos = [], []
    for cam, G in zip(cams, Gs):
        direct, _, _, _ = _cuda_backward(sc, cam, H, W, deg, G, False, dev)
        total += direct["shs"].double()
        _, dcol, cp, means = _cuda_backward(sc, cam, H, W, deg, G, True, dev)
        dcols.append(dcol)
        campos.append(cp)
    dc = torch.stack(dcols).contiguous()
    cps = torch.stack(campos).contiguous()
    out = torch.empty(P, 16, 3, device=dev)
    rasterizer.sh_grad_combine(means, cps, dc, deg, out)
    torch.cuda.synchronize()
    scale = float(total.abs().max())
    assert float((out.double() - total).abs().max()) <= 1e-4 * scale
    ref = TN.sh_grad_combine_ref(means.cpu(), cps.cpu(), dc.cpu(), deg, 16)
    assert float((out.cpu().double() - ref).abs().max()) <= 2e-6 * float(ref.abs().max())
    # unaligned / odd-sized output rows take the scalar store path: M = 9 at degree 2
    out9 = torch.empty(P, 9, 3, device=dev)
    rasterizer.sh_grad_combine(means, cps, dc, 2, out9)
    ref9 = TN.sh_grad_combine_ref(means.cpu(), cps.cpu(), dc.cpu(), 2, 9)
    assert float((out9.cpu().double() - ref9).abs().max()) <= 2e-6 * float(ref9.abs().max())


@pytest.mark.gpu
def test_sh_grad_combine_argument_errors(cuda_lib):
    from splatfields_b200 import rasterizer
    dev = torch.device("cuda")
    means = torch.zeros(8, 3, device=dev)
    with pytest.raises(Exception):
        rasterizer.sh_grad_combine(means, torch.zeros(65, 3, device=dev), torch.zeros(65, 8, 3, device=dev), 3,
                                   torch.empty(8, 16, 3, device=dev))          # V > 64
    with pytest.raises(Exception):
        rasterizer.sh_grad_combine(means, torch.zeros(1, 3, device=dev), torch.zeros(1, 8, 3, device=dev), 3,
                                   torch.empty(8, 9, 3, device=dev))           # M < (deg+1)^2
