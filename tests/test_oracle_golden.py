"""CPU: pin the oracle (and the synthetic-camera helpers) against vectors produced by the reference's
own Python helpers (tests/golden/make_golden.py -> reference_helpers.npz)."""
import math
import os

import numpy as np
import torch

from splatfields_b200 import synth

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_helpers.npz"))


def test_world_to_view_and_projection_match_reference():
    # the reference inverts twice (C2W and back), so allow its round-off
    np.testing.assert_allclose(synth.world_to_view(G["cam0_R"], G["cam0_T"]), G["w2v2"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(synth.projection_matrix(0.01, 100.0, 0.7, 0.6), G["projm"], rtol=0, atol=1e-7)


def test_synth_camera_matches_reference_camera_class():
    for k in range(4):
        cam = synth.make_camera(G[f"cam{k}_R"], G[f"cam{k}_T"], float(G[f"cam{k}_fovx"]), float(G[f"cam{k}_fovy"]),
                                int(G[f"cam{k}_H"]), int(G[f"cam{k}_W"]))
        np.testing.assert_allclose(cam.world_view_transform.numpy(), G[f"cam{k}_wvt"], rtol=0, atol=1e-7)
        np.testing.assert_allclose(cam.projection_matrix.numpy(), G[f"cam{k}_proj"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(cam.full_proj_transform.numpy(), G[f"cam{k}_full"], rtol=0, atol=1e-5)
        np.testing.assert_allclose(cam.camera_center.numpy(), G[f"cam{k}_center"], rtol=0, atol=1e-5)


def _pre(oracle, k, pts, **extra):
    H, W = int(G[f"cam{k}_H"]), int(G[f"cam{k}_W"])
    P = pts.shape[0]
    kw = dict(scales=np.full((P, 3), 0.01, np.float32), rotations=np.tile(np.array([1, 0, 0, 0], np.float32), (P, 1)),
              shs=None, colors_precomp=np.ones((P, 3), np.float32), cov3D_precomp=None)
    kw.update(extra)
    return oracle.preprocess(pts, np.full(P, 0.5, np.float32), kw["scales"], kw["rotations"], kw["shs"],
                             kw["colors_precomp"], kw["cov3D_precomp"], G[f"cam{k}_wvt"], G[f"cam{k}_full"],
                             G[f"cam{k}_center"], math.tan(float(G[f"cam{k}_fovx"]) / 2),
                             math.tan(float(G[f"cam{k}_fovy"]) / 2), H, W, extra.get("sh_degree", 0)), H, W


def test_projection_matches_geom_transform_points(oracle):
    pts = G["pts"]
    for k in range(4):
        g, H, W = _pre(oracle, k, pts)
        vis = g["radii"] > 0
        assert vis.sum() > 10
        ndc = G[f"ndc{k}"]
        px = ((ndc[:, 0].astype(np.float64) + 1.0) * W - 1.0) * 0.5
        py = ((ndc[:, 1].astype(np.float64) + 1.0) * H - 1.0) * 0.5
        np.testing.assert_allclose(g["means2D"][vis, 0], px[vis], rtol=0, atol=2e-4)
        np.testing.assert_allclose(g["means2D"][vis, 1], py[vis], rtol=0, atol=2e-4)
        # depth = view-space z of the reference transform (w of that affine matrix is 1)
        np.testing.assert_allclose(g["depths"][vis], G[f"view{k}"][vis, 2], rtol=1e-6, atol=1e-6)
        # near cull: exactly the points with view z > 0.2 can be visible
        assert not (vis & (G[f"view{k}"][:, 2] <= 0.2)).any()


def test_sh_colour_matches_eval_sh(oracle):
    pts, shs = G["pts"], G["shs"]
    k = 0
    # `campos` only enters the SH view direction, the view matrix only the projection: put a fake
    # camera position 3 units in front of the real camera and the points on a small sphere around it,
    # so that normalize(p - campos) == golden dirs while every point stays on screen.
    fwd = G[f"cam{k}_wvt"][:3, 2]
    cam_center = (G[f"cam{k}_center"] + 3.0 * fwd).astype(np.float32)
    p = (cam_center[None, :].astype(np.float64) + 0.25 * G["dirs"]).astype(np.float32)
    for deg in range(4):
        H, W = int(G[f"cam{k}_H"]), int(G[f"cam{k}_W"])
        P = p.shape[0]
        g = oracle.preprocess(p, np.full(P, 0.5, np.float32), np.full((P, 3), 0.01, np.float32),
                              np.tile(np.array([1, 0, 0, 0], np.float32), (P, 1)), shs, None, None,
                              G[f"cam{k}_wvt"], G[f"cam{k}_full"], cam_center,
                              math.tan(float(G[f"cam{k}_fovx"]) / 2), math.tan(float(G[f"cam{k}_fovy"]) / 2), H, W, deg)
        vis = g["radii"] > 0
        assert vis.sum() >= 40
        ref = G[f"sh_rgb{deg}"] + 0.5
        np.testing.assert_allclose(g["rgb"][vis], np.maximum(ref[vis], 0), rtol=0, atol=5e-6)
        assert np.array_equal(g["clamped"][vis].astype(bool), ref[vis] < 0) or \
            np.abs(ref[vis][g["clamped"][vis].astype(bool) != (ref[vis] < 0)]).max() < 1e-6


def test_covariance_matches_build_scaling_rotation(oracle):
    q = G["quats"] / np.linalg.norm(G["quats"], axis=1, keepdims=True)   # build_rotation normalises
    pts = np.tile(np.array([[0.0, 0.0, 0.0]], np.float32), (q.shape[0], 1))
    g, H, W = _pre(oracle, 1, pts, scales=G["scales"], rotations=q.astype(np.float32))
    assert (g["radii"] > 0).all()
    np.testing.assert_allclose(g["cov3D"], G["cov6"], rtol=2e-5, atol=1e-9)


def test_torch_naive_follows_reference_helpers():
    from oracle import torch_naive as TN
    q = torch.tensor(G["quats"]).double()
    qn = q / q.norm(dim=1, keepdim=True)
    np.testing.assert_allclose(TN.quat_to_rot(qn).numpy(), G["rotmats"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(TN.cov3d_6(torch.tensor(G["scales"]).double(), qn).numpy(), G["cov6"], rtol=1e-5, atol=1e-9)
    d = torch.tensor(G["dirs"]).double()
    for deg in range(4):
        nb = (deg + 1) ** 2
        rgb = torch.einsum("nk,nkc->nc", TN.sh_basis(deg, d), torch.tensor(G["shs"]).double()[:, :nb])
        np.testing.assert_allclose(rgb.numpy(), G[f"sh_rgb{deg}"], rtol=0, atol=1e-6)
