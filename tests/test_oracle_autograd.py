"""CPU: the C oracle's analytic backward against fp64 autograd through an independent dense
restatement (oracle/torch_naive.py), and its forward against the same.  Config 0 of BASELINE.json
(256 Gaussians, 64x64, pure PyTorch on CPU) lives here."""
import math

import numpy as np
import pytest
import torch

from oracle import torch_naive as TN
from splatfields_b200 import synth


def _run(oracle, sc, cam, H, W, deg, bg, depth_cotangent=False):
    tfx, tfy = math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5)
    kw = dict(bg=bg.numpy(), viewmatrix=cam.world_view_transform.numpy(), projmatrix=cam.full_proj_transform.numpy(),
              campos=cam.camera_center.numpy(), tanfovx=tfx, tanfovy=tfy, H=H, W=W, sh_degree=deg)
    n = lambda k: sc[k].numpy() if k in sc else None
    f = oracle.forward(n("means3D"), n("opacities"), n("scales"), n("rotations"), shs=n("shs"),
                       colors_precomp=n("colors_precomp"), cov3D_precomp=n("cov3D_precomp"), want_margin=True, **kw)
    leaf = {k: v.clone().double().requires_grad_(True) for k, v in sc.items()}
    col, dep, radii, aux = TN.render_dense(
        leaf["means3D"], leaf["opacities"], leaf.get("scales"), leaf.get("rotations"), shs=leaf.get("shs"),
        colors_precomp=leaf.get("colors_precomp"), cov3D_precomp=leaf.get("cov3D_precomp"), bg=bg,
        viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, campos=cam.camera_center,
        tanfovx=tfx, tanfovy=tfy, H=H, W=W, sh_degree=deg)
    aux["ndc"].retain_grad()
    ok = f["margin"] > 1e-4          # pixels whose skip/stop decisions are not within 1e-4 of a threshold
    assert ok.mean() > 0.95
    assert np.array_equal(radii.numpy(), f["radii"])
    assert np.abs(col.detach().numpy() - f["color"])[:, ok].max() < 1e-5
    assert np.abs(dep.detach().numpy() - f["depth"])[:, ok].max() < 5e-5
    G = torch.randn(3, H, W, generator=torch.Generator().manual_seed(5))
    G[:, torch.tensor(~ok)] = 0
    Gd = None
    loss = (col * G.double()).sum()
    if depth_cotangent:       # a loss on the depth image as well (train.py:217-229, lambda_depth > 0)
        Gd = torch.randn(1, H, W, generator=torch.Generator().manual_seed(6))
        Gd[:, torch.tensor(~ok)] = 0
        loss = loss + (dep * Gd.double()).sum()
    loss.backward()
    b = oracle.backward(f, G.numpy(), n("means3D"), n("scales"), n("rotations"), shs=n("shs"),
                        cov3D_precomp=n("cov3D_precomp"), viewmatrix=kw["viewmatrix"], projmatrix=kw["projmatrix"],
                        campos=kw["campos"], tanfovx=tfx, tanfovy=tfy, sh_degree=deg,
                        dL_ddepth=None if Gd is None else Gd.numpy())
    return f, b, leaf, aux


def _rel(a, r):
    a, r = np.asarray(a, np.float64), r.numpy()
    return np.abs(a - r).max() / max(np.abs(r).max(), 1e-30)


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_plumbing_config_sh(oracle, deg):
    cfg = synth.CONFIGS["plumbing_256"]
    sc = synth.make_scene(cfg["P"], cfg["seed"], scale_mult=6.0, extent=1.0)
    cam = synth.config_camera("plumbing_256")
    f, b, leaf, aux = _run(oracle, sc, cam, cfg["H"], cfg["W"], deg, torch.tensor([1.0, 0.5, 0.2]))
    assert f["num_rendered"] > 500
    assert _rel(b["dL_dmeans3D"], leaf["means3D"].grad) < 1e-4
    assert _rel(b["dL_dopacity"], leaf["opacities"].grad.reshape(-1, 1)) < 1e-4
    assert _rel(b["dL_dscales"], leaf["scales"].grad) < 1e-4
    assert _rel(b["dL_drotations"], leaf["rotations"].grad) < 1e-4
    assert _rel(b["dL_dsh"], leaf["shs"].grad) < 1e-4
    assert _rel(b["dL_dmeans2D"][:, :2], aux["ndc"].grad) < 1e-4
    nb = (deg + 1) ** 2
    assert np.all(b["dL_dsh"][:, nb:, :] == 0)


def test_depth_cotangent_matches_autograd(oracle):
    """The depth image is differentiable when its cotangent is given: colour + depth loss against fp64
    autograd through the dense restatement."""
    cfg = synth.CONFIGS["plumbing_256"]
    sc = synth.make_scene(cfg["P"], cfg["seed"], scale_mult=6.0, extent=1.0)
    cam = synth.config_camera("plumbing_256")
    f, b, leaf, aux = _run(oracle, sc, cam, cfg["H"], cfg["W"], 2, torch.tensor([1.0, 0.5, 0.2]), depth_cotangent=True)
    assert _rel(b["dL_dmeans3D"], leaf["means3D"].grad) < 1e-4
    assert _rel(b["dL_dopacity"], leaf["opacities"].grad.reshape(-1, 1)) < 1e-4
    assert _rel(b["dL_dscales"], leaf["scales"].grad) < 1e-4
    assert _rel(b["dL_drotations"], leaf["rotations"].grad) < 1e-4
    assert _rel(b["dL_dsh"], leaf["shs"].grad) < 1e-4
    assert _rel(b["dL_dmeans2D"][:, :2], aux["ndc"].grad) < 1e-4
    # and the depth term is not a no-op: the same scene without it gives different mean gradients
    _, b0, leaf0, _ = _run(oracle, sc, cam, cfg["H"], cfg["W"], 2, torch.tensor([1.0, 0.5, 0.2]))
    assert np.abs(b["dL_dmeans3D"] - b0["dL_dmeans3D"]).max() > 1e-3 * np.abs(b0["dL_dmeans3D"]).max()


def test_plumbing_config_precomputed_rgb_and_cov(oracle):
    cfg = synth.CONFIGS["plumbing_256"]
    sc = synth.make_scene(cfg["P"], cfg["seed"], scale_mult=5.0, extent=1.0, precomp_rgb=True)
    sc["cov3D_precomp"] = TN.cov3d_6(sc["scales"].double(), sc["rotations"].double()).float()
    del sc["scales"], sc["rotations"]
    cam = synth.config_camera("plumbing_256")
    f, b, leaf, aux = _run(oracle, sc, cam, cfg["H"], cfg["W"], 0, torch.tensor([0.0, 0.0, 0.0]))
    assert _rel(b["dL_dmeans3D"], leaf["means3D"].grad) < 1e-4
    assert _rel(b["dL_dcolors"], leaf["colors_precomp"].grad) < 1e-4
    assert _rel(b["dL_dcov3D"], leaf["cov3D_precomp"].grad) < 1e-4
    assert _rel(b["dL_dopacity"], leaf["opacities"].grad.reshape(-1, 1)) < 1e-4


def test_oracle_binning_invariants(oracle):
    sc = synth.make_scene(3000, 11, scale_mult=2.0)
    cam = synth.orbit_camera(1, 128, 160)
    from tests.helpers import run_oracle
    f, _ = run_oracle(oracle, sc, cam, 128, 160, [1, 1, 1], 3)
    keys, pl, R = f["point_list_keys"], f["point_list"], f["num_rendered"]
    assert R == int(f["tiles_touched"].sum()) and R > 0
    assert np.all(keys[1:] >= keys[:-1])
    # ties in (tile, depth) keep ascending Gaussian index
    same = keys[1:] == keys[:-1]
    assert np.all(pl[1:][same] > pl[:-1][same])
    # the low 32 bits are the float32 depth bits of the listed Gaussian
    assert np.array_equal((keys & np.uint64(0xFFFFFFFF)).astype(np.uint32), f["depths"][pl].view(np.uint32))
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    rg = f["ranges"].astype(np.int64)
    for t in np.unique(tiles):
        idx = np.nonzero(tiles == t)[0]
        assert rg[t, 0] == idx[0] and rg[t, 1] == idx[-1] + 1
    empty = np.setdiff1d(np.arange(rg.shape[0]), np.unique(tiles))
    assert np.all(rg[empty] == 0)
    # independent construction: every (Gaussian, tile of its rectangle) instance, ordered by (tile, depth bits, index)
    gx, gy = (160 + 15) // 16, (128 + 15) // 16
    inst = []
    for i in np.nonzero(f["radii"] > 0)[0]:
        x, y, r = f["means2D"][i, 0], f["means2D"][i, 1], np.float32(f["radii"][i])
        x0, y0 = min(gx, max(0, int((x - r) / 16))), min(gy, max(0, int((y - r) / 16)))
        x1, y1 = min(gx, max(0, int((x + r + 15) / 16))), min(gy, max(0, int((y + r + 15) / 16)))
        d = int(f["depths"][i:i + 1].view(np.uint32)[0])
        inst += [((ty * gx + tx) << 32 | d, i) for ty in range(y0, y1) for tx in range(x0, x1)]
    inst.sort()
    assert len(inst) == R
    assert np.array_equal(np.array([k for k, _ in inst], np.uint64), keys)
    assert np.array_equal(np.array([i for _, i in inst], np.uint32), pl)


def test_empty_and_invisible_inputs(oracle):
    cam = synth.orbit_camera(0, 32, 48)
    from tests.helpers import run_oracle
    sc = synth.make_scene(64, 3)
    sc["means3D"] = sc["means3D"] + torch.tensor([100.0, 100.0, 100.0])   # everything behind / off screen
    f, b = run_oracle(oracle, sc, cam, 32, 48, [0.2, 0.4, 0.6], 3, dL=np.ones((3, 32, 48), np.float32))
    assert f["num_rendered"] == 0 and np.all(f["radii"] == 0)
    assert np.allclose(f["color"], np.array([0.2, 0.4, 0.6], np.float32)[:, None, None])
    assert all(np.all(v == 0) for v in b.values())


def test_oracle_parallel_sort_matches_numpy_stable_sort(oracle):
    """The multi-threaded path of the oracle's radix sort (>= 65536 instances, >= 2 threads): same list as numpy's
    stable sort of the emission-order keys."""
    from tests.helpers import run_oracle
    oracle.set_num_threads(4)
    H, W = 256, 320
    sc = synth.make_scene(20000, 13, scale_mult=2.5)
    f, _ = run_oracle(oracle, sc, synth.orbit_camera(2, H, W), H, W, [1, 1, 1], 3, want_margin=False)
    R = f["num_rendered"]
    assert R >= (1 << 16)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    vis = np.nonzero(f["radii"] > 0)[0]
    x, y, r = f["means2D"][vis, 0], f["means2D"][vis, 1], f["radii"][vis].astype(np.float32)
    cl = lambda v, hi: np.clip(v.astype(np.int64), 0, hi)
    x0, y0 = cl((x - r) / np.float32(16), gx), cl((y - r) / np.float32(16), gy)
    x1, y1 = cl((x + r + np.float32(15)) / np.float32(16), gx), cl((y + r + np.float32(15)) / np.float32(16), gy)
    d = f["depths"][vis].view(np.uint32).astype(np.uint64)
    keys, vals = [], []
    for k, i in enumerate(vis):                 # emission order: Gaussian index, then rows, then columns
        ty, tx = np.meshgrid(np.arange(y0[k], y1[k]), np.arange(x0[k], x1[k]), indexing="ij")
        t = (ty * gx + tx).reshape(-1).astype(np.uint64)
        keys.append((t << np.uint64(32)) | d[k])
        vals.append(np.full(t.size, i, np.uint32))
    keys, vals = np.concatenate(keys), np.concatenate(vals)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(keys[order], f["point_list_keys"])
    assert np.array_equal(vals[order], f["point_list"])
