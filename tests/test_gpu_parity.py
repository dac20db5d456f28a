"""GPU (-m gpu): the sm_100a library, called through the reference-facing GaussianRasterizer -> C ABI,
against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): bit-exact tile key / index buffers (and everything integer that feeds
them), 1e-5 abs on RGB / depth, 1e-3 rel on every per-splat gradient.  Pixels whose skip/stop decision
sits within 1e-4 (relative) of a threshold in the oracle are excluded from the image comparison: which
side of alpha = 1/255 they fall on depends on the last ulp of expf (glibc vs CUDA libdevice), exactly as
it would between the reference's CUDA build and any CPU restatement."""
import numpy as np
import pytest
import torch

from tests.helpers import IMG_ATOL, grad_close, run_cuda, run_oracle, scene_and_camera

pytestmark = pytest.mark.gpu

CASES = [
    # name,            P,      H,   W,  deg, kwargs
    ("small_sh3",      5_000,  128, 160, 3, dict(seed=21, scale_mult=2.0)),
    ("ragged_sh2",     8_000,  117, 203, 2, dict(seed=22, scale_mult=2.5)),
    ("sh1",            4_000,  96,  96,  1, dict(seed=23, scale_mult=3.0)),
    ("sh0",            4_000,  96,  96,  0, dict(seed=24, scale_mult=3.0)),
    ("rgb_precomp",    6_000,  144, 176, 0, dict(seed=25, scale_mult=2.0, precomp_rgb=True)),
    ("cov_precomp",    6_000,  144, 176, 3, dict(seed=26, scale_mult=2.0, cov_precomp=True)),
    ("dense_long",     30_000, 256, 256, 3, dict(seed=27, scale_mult=4.0)),
]


def _compare(O, sc, cam, H, W, deg, bg=(1.0, 1.0, 1.0), check_grads=True, ill_conditioned=False, min_ok=0.99):
    gen = torch.Generator().manual_seed(77)
    dL = torch.randn(3, H, W, generator=gen).numpy()
    f, _ = run_oracle(O, sc, cam, H, W, bg, deg)
    ok = f["margin"] > 1e-4
    dL_m = dL.copy()
    dL_m[:, ~ok] = 0
    f, b = run_oracle(O, sc, cam, H, W, bg, deg, dL=dL_m)
    c, g = run_cuda(sc, cam, H, W, bg, deg, dL=dL_m if check_grads else None)

    # ---- integer / key-feeding state: bit exact ----
    assert np.array_equal(c["radii"], f["radii"])
    assert np.array_equal(c["tiles_touched"], f["tiles_touched"])
    vis = f["radii"] > 0
    assert np.array_equal(c["depths"][vis].view(np.uint32), f["depths"][vis].view(np.uint32))
    assert np.array_equal(c["means2D"][vis].view(np.uint32), f["means2D"][vis].view(np.uint32))
    assert np.array_equal(c["conic_opacity"][vis].view(np.uint32), f["conic_opacity"][vis].view(np.uint32))
    assert np.array_equal(c["cov3D"][vis].view(np.uint32), f["cov3D"][vis].view(np.uint32))
    assert np.array_equal(c["rgb"][vis].view(np.uint32), f["rgb"][vis].view(np.uint32))
    assert np.array_equal(c["clamped"][vis], f["clamped"][vis])
    assert c["num_rendered"] == f["num_rendered"]
    assert np.array_equal(c["point_list_keys"], f["point_list_keys"])
    assert np.array_equal(c["point_list"], f["point_list"])
    assert np.array_equal(c["ranges"], f["ranges"])

    # ---- image: 1e-5 abs away from decision thresholds ----
    assert ok.mean() >= min_ok, f"too many threshold-sensitive pixels: {1 - ok.mean():.4f}"
    assert np.abs(c["color"] - f["color"])[:, ok].max() <= IMG_ATOL
    assert np.abs(c["depth"] - f["depth"])[:, ok].max() <= IMG_ATOL * max(1.0, float(f["depth"].max()))
    assert np.abs(c["final_T"] - f["final_T"])[ok].max() <= IMG_ATOL
    assert np.array_equal(c["n_contrib"][ok], f["n_contrib"][ok])
    # sensitive pixels may flip one contributor, never more than a 1/255-alpha step
    assert np.abs(c["color"] - f["color"]).max() <= 2.0 / 255.0 * max(1.0, float(np.abs(f["rgb"]).max()))

    # ---- gradients: 1e-3 rel ----
    if check_grads:
        for k in g:
            ref = b[k] if k != "dL_dopacity" else b[k].reshape(g[k].shape)
            if ill_conditioned and k in ("dL_dmeans3D", "dL_dscales", "dL_drotations", "dL_dcov3D"):
                # These go through the inverse of a near-singular 2x2 covariance (condition number 1e3..1e5
                # for needle-like splats): a 2e-7 relative (one-ulp) perturbation of the pixel sums moves them
                # by 2-3 % in the oracle itself, so fp32 summation order alone exceeds 1e-3 here.  The pixel
                # sums themselves (means2D, opacity, SH / colour gradients) are still held to 1e-3 below.
                got, r = np.asarray(g[k], np.float64), np.asarray(ref, np.float64)
                assert np.linalg.norm(got - r) / max(np.linalg.norm(r), 1e-30) < 5e-2, k
                continue
            grad_close(k, g[k], ref)
    return f, c


def _mut_aniso(sc):
    """needle-like splats: lambda1/lambda2 of the screen covariance far above 1e4 -> the never-cull path"""
    sc["scales"] = sc["scales"] * torch.tensor([0.03, 1.0, 25.0])


def _mut_opacity(sc):
    """opacities below the 1/255 visibility floor (never contribute), exactly 0, and 1 (0.99 clamp, fast saturation)"""
    o = sc["opacities"]
    n = o.shape[0]
    o[: n // 4] = 0.003
    o[n // 4: n // 3] = 0.0
    o[n // 3: n // 2] = 1.0
    o[n // 2: n // 2 + 50] = 1.0 / 255.0


def _mut_huge(sc):
    sc["scales"] = sc["scales"] * 40.0


EDGE_CASES = [
    ("needles",          4_000, 128, 160, 3, dict(seed=61, scale_mult=2.0), _mut_aniso),
    ("opacity_extremes", 6_000, 128, 160, 3, dict(seed=62, scale_mult=3.0), _mut_opacity),
    ("huge_splats",        300, 96,  112, 2, dict(seed=63, scale_mult=1.0), _mut_huge),
    ("tiny_image",         500, 17,  17,  3, dict(seed=64, scale_mult=6.0), None),
    ("one_pixel_row",      800, 1,   300, 1, dict(seed=65, scale_mult=6.0), None),
]


@pytest.mark.parametrize("name,P,H,W,deg,kw,mut", EDGE_CASES, ids=[c[0] for c in EDGE_CASES])
def test_parity_edge_cases(oracle, cuda_lib, name, P, H, W, deg, kw, mut):
    kw = dict(kw)
    seed = kw.pop("seed")
    sc, cam = scene_and_camera(P, H, W, seed, sh_degree=deg, **kw)
    if mut is not None:
        mut(sc)
    f, c = _compare(oracle, sc, cam, H, W, deg, bg=(0.2, 0.9, 0.4), ill_conditioned=name in ("needles", "huge_splats"))
    assert f["num_rendered"] > 0


@pytest.mark.parametrize("name,P,H,W,deg,kw", CASES, ids=[c[0] for c in CASES])
def test_parity_small(oracle, cuda_lib, name, P, H, W, deg, kw):
    kw = dict(kw)
    seed = kw.pop("seed")
    sc, cam = scene_and_camera(P, H, W, seed, sh_degree=deg, **kw)
    _compare(oracle, sc, cam, H, W, deg, bg=(1.0, 0.5, 0.2) if name != "rgb_precomp" else (0.0, 0.0, 0.0))


def test_parity_config1_lego_100k(oracle, cuda_lib):
    """BASELINE.json configs[1]: 100k Gaussians, 800x800, SH degree 3, fwd+bwd."""
    from splatfields_b200 import synth
    cfg = synth.CONFIGS["lego_100k"]
    sc = synth.make_scene(cfg["P"], cfg["seed"])
    cam = synth.config_camera("lego_100k", 0)
    _compare(oracle, sc, cam, cfg["H"], cfg["W"], 3)


def test_clamped_fov_and_near_plane(oracle, cuda_lib):
    """Gaussians far outside the frustum (tan-fov clamp active) and around the 0.2 near plane."""
    sc, cam = scene_and_camera(6_000, 128, 128, 31, scale_mult=6.0, extent=3.5)
    f, c = _compare(oracle, sc, cam, 128, 128, 3)
    assert (f["radii"] == 0).sum() > 100 and (f["radii"] > 0).sum() > 100


def test_scale_modifier_and_background(oracle, cuda_lib):
    import math
    from splatfields_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer
    sc, cam = scene_and_camera(3_000, 96, 128, 41, scale_mult=2.0)
    H, W = 96, 128
    n = lambda k: sc[k].numpy()
    from tests.helpers import cam_kwargs
    kw = cam_kwargs(cam, H, W, (0.3, 0.6, 0.9))
    f = oracle.forward(n("means3D"), n("opacities"), n("scales"), n("rotations"), shs=n("shs"), sh_degree=3,
                       scale_modifier=1.7, want_margin=True, **kw)
    dev = torch.device("cuda")
    camd = cam.to(dev)
    rs = GaussianRasterizationSettings(H, W, math.tan(cam.FoVx / 2), math.tan(cam.FoVy / 2),
                                       torch.tensor([0.3, 0.6, 0.9], device=dev), 1.7, camd.world_view_transform,
                                       camd.full_proj_transform, 3, camd.camera_center, False, True)
    t = {k: v.to(dev) for k, v in sc.items()}
    color, radii, depth = GaussianRasterizer(rs)(means3D=t["means3D"], means2D=torch.zeros_like(t["means3D"]),
                                                 opacities=t["opacities"], shs=t["shs"], scales=t["scales"],
                                                 rotations=t["rotations"])
    ok = f["margin"] > 1e-4
    assert np.array_equal(radii.cpu().numpy(), f["radii"])
    assert np.abs(color.cpu().numpy() - f["color"])[:, ok].max() <= IMG_ATOL
    assert depth.shape == (1, H, W) and radii.dtype == torch.int32


def test_fused_alpha_matches_two_reference_passes(oracle, cuda_lib):
    """render() runs the whole rasterizer twice when return_opacity (colour pass, then colours = 1
    / bg = 0).  The fused path must give the second pass's image and the SUM of both passes' gradients."""
    import math
    from splatfields_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer
    from tests.helpers import cam_kwargs
    P, H, W, deg = 6_000, 144, 176, 3
    sc, cam = scene_and_camera(P, H, W, 51, scale_mult=2.5)
    bg = (1.0, 0.5, 0.2)
    gen = torch.Generator().manual_seed(78)
    dLc = torch.randn(3, H, W, generator=gen).numpy()
    dLa = torch.randn(1, H, W, generator=gen).numpy()
    n = lambda k: sc[k].numpy()
    kw = cam_kwargs(cam, H, W, bg)
    f1 = oracle.forward(n("means3D"), n("opacities"), n("scales"), n("rotations"), shs=n("shs"), sh_degree=deg,
                        want_margin=True, **kw)
    ok = f1["margin"] > 1e-4
    dLc[:, ~ok] = 0
    dLa[:, ~ok] = 0
    kw0 = dict(kw)
    kw0["bg"] = np.zeros(3, np.float32)
    ones = np.ones((P, 3), np.float32)
    f2 = oracle.forward(n("means3D"), n("opacities"), n("scales"), n("rotations"), colors_precomp=ones,
                        sh_degree=0, **kw0)
    bk = dict(viewmatrix=kw["viewmatrix"], projmatrix=kw["projmatrix"], campos=kw["campos"], tanfovx=kw["tanfovx"],
              tanfovy=kw["tanfovy"])
    b1 = oracle.backward(f1, dLc, n("means3D"), n("scales"), n("rotations"), shs=n("shs"), sh_degree=deg, **bk)
    dLa3 = np.concatenate([dLa, np.zeros((2, H, W), np.float32)], 0)   # reference uses channel 0 of the alpha pass
    b2 = oracle.backward(f2, dLa3, n("means3D"), n("scales"), n("rotations"), sh_degree=0, **bk)

    dev = torch.device("cuda")
    camd = cam.to(dev)
    rs = GaussianRasterizationSettings(H, W, math.tan(cam.FoVx / 2), math.tan(cam.FoVy / 2),
                                       torch.tensor(bg, device=dev), 1.0, camd.world_view_transform,
                                       camd.full_proj_transform, deg, camd.camera_center, False, False)
    t = {k: v.to(dev).clone().requires_grad_(True) for k, v in sc.items()}
    m2d = torch.zeros(P, 3, device=dev, requires_grad=True)
    color, radii, depth, alpha = GaussianRasterizer(rs)(
        means3D=t["means3D"], means2D=m2d, opacities=t["opacities"], shs=t["shs"], scales=t["scales"],
        rotations=t["rotations"], with_alpha=True)
    assert alpha.shape == (1, H, W)
    assert np.abs(color.detach().cpu().numpy() - f1["color"])[:, ok].max() <= IMG_ATOL
    assert np.abs(alpha.detach().cpu().numpy()[0] - f2["color"][0])[ok].max() <= IMG_ATOL
    # coverage = 1 - final transmittance (telescoping sum), to rounding
    assert np.abs(alpha.detach().cpu().numpy()[0] - (1.0 - f1["final_T"]))[ok].max() <= 1e-5
    ((color * torch.as_tensor(dLc, device=dev)).sum() + (alpha * torch.as_tensor(dLa, device=dev)).sum()).backward()
    got = dict(dL_dmeans3D=t["means3D"].grad, dL_dmeans2D=m2d.grad, dL_dopacity=t["opacities"].grad,
               dL_dsh=t["shs"].grad, dL_dscales=t["scales"].grad, dL_drotations=t["rotations"].grad)
    for k, v in got.items():
        ref = b1[k].astype(np.float64) + (b2[k].astype(np.float64) if k != "dL_dsh" else 0.0)
        grad_close(k, v.detach().cpu().numpy(), ref.reshape(v.shape))
    # and the mirror of render(): fused == two-pass
    from splatfields_b200 import render
    gd = dict(means3D=t["means3D"].detach(), active_sh_degree=deg, gaussian_opacity=t["opacities"].detach(),
              gaussian_scales=t["scales"].detach(), gaussian_rotations=t["rotations"].detach(),
              gaussian_features=t["shs"].detach())
    o2 = render(camd, gd, None, torch.tensor(bg, device=dev), return_opacity=True, fused_alpha=False)
    o1 = render(camd, gd, None, torch.tensor(bg, device=dev), return_opacity=True, fused_alpha=True)
    assert torch.equal(o1["render"], o2["render"]) and torch.equal(o1["radii"], o2["radii"])
    assert float((o1["opacity"] - o2["opacity"]).abs().max()) <= 1e-6


@pytest.mark.parametrize("precomp_rgb", [False, True], ids=["sh3", "rgb"])
def test_depth_cotangent_matches_the_oracle(oracle, cuda_lib, precomp_rgb):
    """A loss on the depth image (train.py:195-229) reaches the Gaussians: colour + depth cotangents against the oracle
    (whose depth backward is checked against fp64 autograd on the CPU, tests/test_oracle_autograd.py)."""
    P, H, W = 8_000, 160, 208
    deg = 0 if precomp_rgb else 3
    sc, cam = scene_and_camera(P, H, W, 61, scale_mult=2.5, precomp_rgb=precomp_rgb)
    bg = (0.3, 0.6, 0.9)
    gen = torch.Generator().manual_seed(79)
    dL = torch.randn(3, H, W, generator=gen).numpy()
    dLd = torch.randn(1, H, W, generator=gen).numpy()
    f, _ = run_oracle(oracle, sc, cam, H, W, bg, deg)
    ok = f["margin"] > 1e-4
    dL[:, ~ok] = 0
    dLd[:, ~ok] = 0
    f, b = run_oracle(oracle, sc, cam, H, W, bg, deg, dL=dL, dLd=dLd)
    _, b0 = run_oracle(oracle, sc, cam, H, W, bg, deg, dL=dL)
    c, g = run_cuda(sc, cam, H, W, bg, deg, dL=dL, dLd=dLd)
    for k in g:
        ref = b[k] if k != "dL_dopacity" else b[k].reshape(g[k].shape)
        grad_close(k, g[k], ref)
    # the depth term matters (not a no-op), and without a depth cotangent nothing changes
    assert np.abs(b["dL_dmeans3D"] - b0["dL_dmeans3D"]).max() > 1e-2 * np.abs(b0["dL_dmeans3D"]).max()
    _, g0 = run_cuda(sc, cam, H, W, bg, deg, dL=dL)
    for k in g0:
        ref = b0[k] if k != "dL_dopacity" else b0[k].reshape(g0[k].shape)
        grad_close(k, g0[k], ref)
