"""GPU (-m gpu): size-independent properties at BASELINE.json's full sizes, and the edge cases of the
boundary (empty input, nothing visible, ragged image sizes, the two-pass render() contract)."""
import math

import numpy as np
import pytest
import torch

from splatfields_b200 import synth
from tests.helpers import run_cuda

pytestmark = pytest.mark.gpu


def _full(name, k=0, with_grads=False):
    cfg = synth.CONFIGS[name]
    sc = synth.make_scene(cfg["P"], cfg["seed"], scale_mult=cfg["scale_mult"], precomp_rgb=cfg["precomp_rgb"])
    cam = synth.config_camera(name, k)
    H, W = cfg["H"], cfg["W"]
    dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(9)).numpy() if with_grads else None
    return sc, cam, H, W, dL


@pytest.mark.parametrize("name", ["lego_1m", "dtu_500k", "owlii_2m"])
def test_full_size_binning_properties(cuda_lib, name):
    sc, cam, H, W, _ = _full(name)
    deg = 0 if "colors_precomp" in sc else 3
    c, _ = run_cuda(sc, cam, H, W, (1, 1, 1), deg)
    keys, pl, R = c["point_list_keys"], c["point_list"], c["num_rendered"]
    assert R == int(c["tiles_touched"].astype(np.int64).sum()) and R > 0
    assert (c["radii"] > 0).sum() == (c["tiles_touched"] > 0).sum()
    # sortedness on the full 64-bit key and stability (ties keep ascending Gaussian index)
    assert np.all(keys[1:] >= keys[:-1])
    same = keys[1:] == keys[:-1]
    assert np.all(pl[1:][same] > pl[:-1][same])
    # key low bits are the depth bits of the listed Gaussian; list is a permutation of the emitted instances
    assert np.array_equal((keys & np.uint64(0xFFFFFFFF)).astype(np.uint32), c["depths"][pl].view(np.uint32))
    assert np.array_equal(np.bincount(pl, minlength=c["radii"].shape[0]).astype(np.uint32), c["tiles_touched"])
    # ranges tile the list exactly
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    T = ((W + 15) // 16) * ((H + 15) // 16)
    assert tiles.max() < T
    counts = np.bincount(tiles, minlength=T)
    rg = c["ranges"].astype(np.int64)
    nz = counts > 0
    assert np.array_equal((rg[:, 1] - rg[:, 0])[nz], counts[nz]) and np.all(rg[~nz] == 0)
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
    assert np.array_equal(rg[nz, 0], starts[nz])
    # image sanity: transmittance in [0,1], colour = C + T*bg bounded, depth >= 0
    assert np.isfinite(c["color"]).all() and np.isfinite(c["depth"]).all()
    assert c["final_T"].min() >= 0 and c["final_T"].max() <= 1
    assert c["depth"].min() >= 0


def test_full_size_forward_is_deterministic_and_backward_is_linear(cuda_lib):
    sc, cam, H, W, dL = _full("lego_1m", k=3, with_grads=True)
    c1, g1 = run_cuda(sc, cam, H, W, (1, 1, 1), 3, dL=dL)
    c2, g2 = run_cuda(sc, cam, H, W, (1, 1, 1), 3, dL=2.0 * dL)
    assert np.array_equal(c1["color"], c2["color"]) and np.array_equal(c1["depth"], c2["depth"])
    assert np.array_equal(c1["point_list"], c2["point_list"])
    for k in g1:
        a, b = 2.0 * g1[k].astype(np.float64), g2[k].astype(np.float64)
        nrm = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
        assert nrm < 1e-5, (k, nrm)
    # gradients only where the splat was visible
    inv = c1["radii"] == 0
    for k in g1:
        assert np.all(g1[k][inv] == 0), k
        assert np.isfinite(g1[k]).all(), k


def test_empty_input(cuda_lib):
    from splatfields_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer
    dev = torch.device("cuda")
    cam = synth.orbit_camera(0, 40, 56).to(dev)
    rs = GaussianRasterizationSettings(40, 56, 0.3, 0.3, torch.ones(3, device=dev), 1.0, cam.world_view_transform,
                                       cam.full_proj_transform, 0, cam.camera_center, False, False)
    z = lambda *s: torch.zeros(*s, device=dev)
    color, radii, depth = GaussianRasterizer(rs)(means3D=z(0, 3), means2D=z(0, 3), opacities=z(0, 1),
                                                 colors_precomp=z(0, 3), scales=z(0, 3), rotations=z(0, 4))
    assert color.shape == (3, 40, 56) and radii.shape == (0,) and depth.shape == (1, 40, 56)
    assert torch.all(color == 0) and torch.all(depth == 0)   # the reference returns zero-filled outputs for P == 0


def test_nothing_visible_and_single_splat(cuda_lib):
    sc = synth.make_scene(50, 5)
    cam = synth.orbit_camera(2, 33, 47)
    far = dict(sc)
    far["means3D"] = sc["means3D"] + torch.tensor([50.0, 50.0, 50.0])
    dL = np.ones((3, 33, 47), np.float32)
    c, g = run_cuda(far, cam, 33, 47, (0.1, 0.2, 0.3), 3, dL=dL)
    assert c["num_rendered"] == 0 and np.all(c["radii"] == 0)
    assert np.allclose(c["color"], np.array([0.1, 0.2, 0.3], np.float32)[:, None, None])
    assert all(np.all(v == 0) for v in g.values())
    one = {k: v[:1].clone() for k, v in sc.items()}
    one["means3D"] = torch.zeros(1, 3)
    one["scales"] = torch.full((1, 3), 0.2)
    one["opacities"] = torch.full((1, 1), 0.9)
    c, g = run_cuda(one, cam, 33, 47, (0, 0, 0), 3, dL=dL)
    assert c["radii"][0] > 0 and c["num_rendered"] == c["tiles_touched"][0]
    assert c["color"].max() > 0.1 and np.abs(g["dL_dmeans3D"]).max() > 0


def test_mark_visible_matches_near_plane(cuda_lib):
    from splatfields_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer
    dev = torch.device("cuda")
    sc = synth.make_scene(20_000, 6, extent=4.0)
    cam = synth.orbit_camera(1, 64, 64)
    camd = cam.to(dev)
    rs = GaussianRasterizationSettings(64, 64, 0.3, 0.3, torch.ones(3, device=dev), 1.0, camd.world_view_transform,
                                       camd.full_proj_transform, 0, camd.camera_center, False, False)
    vis = GaussianRasterizer(rs).markVisible(sc["means3D"].to(dev)).cpu().numpy()
    hom = torch.cat([sc["means3D"], torch.ones(20_000, 1)], 1) @ cam.world_view_transform
    z = hom[:, 2].numpy()
    sure = np.abs(z - 0.2) > 1e-5
    assert np.array_equal(vis[sure], (z > 0.2)[sure]) and vis.dtype == np.bool_


def test_render_contract_two_passes(cuda_lib):
    """render(): keys, order and shapes of the reference's gaussian_renderer.render(); the alpha pass with
    colours = 1, bg = 0 renders 1 - final_T; both passes accumulate into the same viewspace_points."""
    from splatfields_b200 import render
    dev = torch.device("cuda")
    sc = synth.make_scene(20_000, 8, scale_mult=2.0)
    cam = synth.orbit_camera(5, 200, 264).to(dev)
    gd = dict(means3D=sc["means3D"].to(dev).requires_grad_(True), active_sh_degree=3,
              gaussian_opacity=sc["opacities"].to(dev).requires_grad_(True),
              gaussian_scales=sc["scales"].to(dev).requires_grad_(True),
              gaussian_rotations=sc["rotations"].to(dev).requires_grad_(True),
              gaussian_features=sc["shs"].to(dev).requires_grad_(True))
    out = render(cam, gd, None, torch.ones(3, device=dev), return_opacity=True)
    assert set(out) == {"render", "viewspace_points", "visibility_filter", "radii", "opacity", "depth"}
    assert out["render"].shape == (3, 200, 264) and out["opacity"].shape == (1, 200, 264)
    assert out["depth"].shape == (1, 200, 264) and out["radii"].dtype == torch.int32
    assert torch.equal(out["visibility_filter"], out["radii"] > 0)
    # with white bg: render = C + T*1  and opacity = 1 - T  =>  every channel <= C + T
    T = 1.0 - out["opacity"]
    assert float(T.min()) >= -1e-5 and float(T.max()) <= 1 + 1e-5
    loss = out["render"].sum() + 0.1 * out["opacity"].sum()
    loss.backward()
    vg = out["viewspace_points"].grad
    assert vg is not None and vg.shape == (20_000, 3) and torch.all(vg[:, 2] == 0)
    assert float(vg[out["visibility_filter"]].abs().sum()) > 0
    assert torch.all(vg[~out["visibility_filter"]] == 0)
    assert gd["means3D"].grad is not None and torch.isfinite(gd["means3D"].grad).all()
    # rgb-function path of the reference (gaussian_rgb_fnc) goes through colors_precomp
    gd2 = {k: v for k, v in gd.items() if k != "gaussian_features"}
    gd2["gaussian_rgb_fnc"] = lambda d: (d * 0.5 + 0.5)
    out2 = render(cam, gd2, None, torch.zeros(3, device=dev), return_opacity=False)
    assert out2["opacity"] is None and torch.isfinite(out2["render"]).all()


@pytest.mark.parametrize("P,H,W,scale,precomp", [(3_000, 64, 64, 4.0, False), (50_000, 240, 320, 2.0, False),
                                                  (100_000, 800, 800, 1.0, False), (300_000, 400, 608, 3.0, True),
                                                  (700, 16, 16, 30.0, False)])
def test_scratch_red_zones_intact(cuda_lib, P, H, W, scale, precomp):
    """debug=True fills a 256-byte red zone behind every carved scratch array and verifies all of them after
    the forward and after the backward: any kernel writing past its array (but inside the caller's allocation,
    where compute-sanitizer cannot see it) turns into an error here."""
    sc = synth.make_scene(P, 77, scale_mult=scale, precomp_rgb=precomp)
    cam = synth.orbit_camera(3, H, W)
    dL = np.ones((3, H, W), np.float32)
    c, g = run_cuda(sc, cam, H, W, (0.5, 0.5, 0.5), 0 if precomp else 3, dL=dL, debug=True)
    assert c["num_rendered"] == int(c["tiles_touched"].astype(np.int64).sum())
    assert all(np.isfinite(v).all() for v in g.values())


def test_host_pipeline_matches_synchronous_path(cuda_lib):
    """HostPipeline (3 streams, double-buffered) returns what the synchronous host call returns."""
    from splatfields_b200.host_api import HostPipeline, ViewParallelRasterizer, forward_backward_host
    dev = torch.device("cuda")
    P, H, W = 30_000, 160, 208
    sc = synth.make_scene(P, 9, scale_mult=2.0)
    cam = synth.orbit_camera(2, H, W)
    G = torch.randn(3, H, W, generator=torch.Generator().manual_seed(3))
    vp = ViewParallelRasterizer(sc, cam, H, W, 3, device=dev)
    host_in, ref_out = vp.pinned_host_buffers(sc, G)
    forward_backward_host(vp, host_in, ref_out)
    ref = {k: v.clone() for k, v in ref_out.items()}
    pipe = HostPipeline(vp)
    outs = [{k: torch.empty_like(v).pin_memory() for k, v in ref_out.items()} for _ in range(3)]
    tickets = [pipe.submit(host_in, outs[i]) for i in range(3)]
    for t in tickets:
        pipe.wait(t)
    pipe.drain()
    for o in outs:
        assert torch.equal(o["color"], ref["color"]) and torch.equal(o["radii"], ref["radii"])
        assert torch.equal(o["depth"], ref["depth"])
        rel = (o["grads"] - ref["grads"]).norm() / ref["grads"].norm()
        assert float(rel) < 1e-5          # atomics: order differs run to run
    # the gradient slab IS the parameters' .grad storage (no flatten copy)
    g = vp.grads()
    assert vp.params["shs"].grad.data_ptr() == g["shs"].data_ptr()
    assert vp.params["means3D"].grad.data_ptr() == g["means3D"].data_ptr()


def test_c_abi_argument_errors(cuda_lib):
    from splatfields_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer
    dev = torch.device("cuda")
    cam = synth.orbit_camera(0, 32, 32).to(dev)
    z = lambda *s: torch.zeros(*s, device=dev)
    rs = GaussianRasterizationSettings(32, 32, 0.3, 0.3, torch.ones(3, device=dev), 1.0, cam.world_view_transform,
                                       cam.full_proj_transform, 3, cam.camera_center, False, False)
    with pytest.raises(Exception, match="sh_degree / M mismatch"):      # degree 3 needs 16 coefficients
        GaussianRasterizer(rs)(means3D=z(8, 3), means2D=z(8, 3), opacities=z(8, 1), shs=z(8, 4, 3),
                               scales=z(8, 3) + 0.1, rotations=z(8, 4) + 0.5)
    with pytest.raises(Exception, match="must have dimensions"):
        GaussianRasterizer(rs)(means3D=z(8, 2), means2D=z(8, 2), opacities=z(8, 1), colors_precomp=z(8, 3),
                               scales=z(8, 3), rotations=z(8, 4))


def test_second_backward_on_the_same_buffers(cuda_lib):
    """The forward render leaves the per-splat gradient accumulators cleared and the FIRST backward relies on it
    (SFB_BWD_ACC_FRESH); a second backward on the same saved buffers (retain_graph) must clear them itself and give
    the same gradients — not the sum of both runs."""
    from splatfields_b200 import render
    dev = torch.device("cuda:0")
    P, H, W = 30_000, 160, 208
    sc = {k: v.to(dev) for k, v in synth.make_scene(P, 21, scale_mult=2.0).items()}
    cam = synth.orbit_camera(2, H, W).to(dev)
    leaves = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
    gd = dict(means3D=leaves["means3D"], active_sh_degree=3, gaussian_opacity=leaves["opacities"],
              gaussian_features=leaves["shs"], gaussian_scales=leaves["scales"], gaussian_rotations=leaves["rotations"])
    out = render(cam, gd, None, torch.ones(3, device=dev), return_opacity=False)
    G = torch.randn(3, H, W, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
    loss = (out["render"] * G).sum()
    g1 = torch.autograd.grad(loss, list(leaves.values()), retain_graph=True)
    g2 = torch.autograd.grad(loss, list(leaves.values()), retain_graph=False)
    for name, a, b in zip(leaves, g1, g2):
        scale = a.abs().max().item()
        assert scale > 0, name
        # float atomics reorder between runs: equal up to summation noise, nowhere near a factor of two
        assert (a - b).abs().max().item() <= 1e-4 * scale, name


def test_tma_row_gather_delivers_the_indexed_rows(cuda_lib):
    """The staging primitive of both render kernels (cp.async.bulk.tensor ... tile::gather4 on a tensor map over the
    48-byte record table, completion on an mbarrier): every gathered 64-byte row is the indexed record + 4 zeros, for
    ragged counts (not a multiple of 4 / 256), repeated and extreme indices."""
    from splatfields_b200 import _lib
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(3)
    for P, n in ((1, 1), (7, 5), (1000, 256), (100_000, 1023), (1_000_000, 4099)):
        table = torch.randn(P, 12, generator=g).to(dev)
        idx = torch.randint(0, P, (n,), generator=g, dtype=torch.int64)
        idx[0], idx[-1] = P - 1, 0
        idx = idx.to(torch.int32).to(dev)
        out = torch.full((n, 16), float("nan"), device=dev)
        _lib.check(cuda_lib.sfb_debug_gather_rows(P, table.data_ptr(), n, idx.data_ptr(), out.data_ptr(), None))
        torch.cuda.synchronize()
        assert torch.equal(out[:, :12], table[idx.long()]), (P, n)
        assert torch.all(out[:, 12:] == 0), (P, n)
