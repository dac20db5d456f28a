"""SURVEY.md §8f-3 (fused GaussianModel activations) and §8f-5 (distCUDA2 replacement).

CPU: the torch restatement of the getters (oracle/torch_next_rows.gaussian_dict_static) against golden vectors made by
the reference's own GaussianModel (tests/golden/make_activations_golden.py -> activations.npz; this row is PINNED), and
the brute-force C kNN oracle against an independent scipy cKDTree (fp64).
GPU: the CUDA kernels, called through the reference-facing Python mirror -> C ABI, against both."""
import os

import numpy as np
import pytest
import torch

from oracle import torch_next_rows as TR

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "activations.npz"))
CASES = ["a64", "a32", "iso32", "m4_32", "m1_32"]
OUT_KEYS = ["gaussian_opacity", "gaussian_features", "gaussian_scales", "gaussian_rotations"]
RAW_KEYS = ["scaling", "rotation", "opacity", "f_dc", "f_rest"]


def _raw(key, dtype=None, device="cpu"):
    raw = {k: torch.from_numpy(GOLD[f"{key}.raw.{k}"]) for k in ["xyz"] + RAW_KEYS}
    if dtype is not None:
        raw = {k: v.to(dtype) for k, v in raw.items()}
    return {k: v.to(device).requires_grad_(k != "xyz") for k, v in raw.items()}


def _relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)) if b.size else 0.0


@pytest.mark.parametrize("key", CASES)
def test_activation_oracle_matches_reference_getters(key):
    raw = _raw(key)
    iso = bool(GOLD[f"{key}.iso"])
    d = TR.gaussian_dict_static(raw["xyz"], raw["scaling"], raw["rotation"], raw["opacity"], raw["f_dc"], raw["f_rest"],
                                use_isotropic=iso)
    for k in OUT_KEYS:
        assert np.array_equal(d[k].detach().numpy(), GOLD[f"{key}.out.{k}"]), k      # same torch ops: bit for bit
    sum((d[k] * torch.from_numpy(GOLD[f"{key}.cot.{k}"])).sum() for k in OUT_KEYS).backward()
    for k in RAW_KEYS:
        assert np.array_equal(raw[k].grad.numpy(), GOLD[f"{key}.grad.{k}"]), k


def _clouds():
    rs = np.random.RandomState(7)
    uni = rs.rand(6000, 3).astype(np.float32) * 2.6 - 1.3                    # dataset_readers.py:598-like init cloud
    clustered = np.concatenate([rs.randn(3000, 3) * 0.01 + 0.7, rs.randn(2000, 3) * 0.5, rs.rand(40, 3) * 50 - 25]
                               ).astype(np.float32)                          # dense blob + halo + far outliers
    planar = np.concatenate([rs.rand(3000, 2), np.zeros((3000, 1))], 1).astype(np.float32)   # zero z extent
    dup = np.repeat(rs.rand(700, 3).astype(np.float32), 4, axis=0)          # every point 4x: 3 neighbours at distance 0
    return {"uniform": uni, "clustered": clustered, "planar": planar, "duplicates": dup}


@pytest.mark.parametrize("name", ["uniform", "clustered", "planar", "duplicates"])
def test_knn_oracle_vs_kdtree(oracle, name):
    from scipy.spatial import cKDTree
    pts = _clouds()[name]
    got = oracle.knn3_mean_dist2(pts)
    if name == "duplicates":
        assert np.all(got == 0.0)
        return
    p64 = pts.astype(np.float64)
    dd, ii = cKDTree(p64).query(p64, k=4)
    assert np.array_equal(ii[:, 0], np.arange(len(pts)))        # no exact duplicates in these clouds: self comes first
    ref = (dd[:, 1:] ** 2).mean(1)
    # fp32 coordinate differences: the error scales with |coordinate|^2 * 2^-24, not with the distance itself
    tol = 4e-7 * (np.abs(p64).max(1) ** 2 + ref) + 1e-6 * ref
    assert np.all(np.abs(got - ref) <= tol)


def test_knn_oracle_tiny_inputs(oracle):
    FLT_MAX = np.finfo(np.float32).max
    assert oracle.knn3_mean_dist2(np.zeros((0, 3), np.float32)).shape == (0,)
    one = oracle.knn3_mean_dist2(np.zeros((1, 3), np.float32))
    assert np.isinf(one[0])                                      # (FLT_MAX + FLT_MAX + FLT_MAX) / 3 overflows
    three = oracle.knn3_mean_dist2(np.eye(3, dtype=np.float32))
    assert np.all(three == np.float32((np.float32(2.0) + np.float32(2.0) + FLT_MAX)) / np.float32(3.0))


# ------------------------------------------------------------------------------------------------ GPU
def _grad_floor_np(k, cot, raw_rotation):
    """Absolute error floor of an fp32 evaluation of the chain rule, per Gaussian [P, 1] — it scales with the
    cotangent, not with the (possibly cancelling) result:
      opacity   g * o * (1 - o): 1 - o carries an absolute error of one ulp of 1 (saturated sigmoids)
      rotation  (g - y (y.g)) / |x|: the subtraction cancels when g is nearly parallel to y
      scaling   g * exp(x): no cancellation (floor 0)"""
    if k == "opacity":
        return 3e-7 * np.abs(cot).reshape(cot.shape[0], -1)
    if k == "rotation":
        ln = np.maximum(np.linalg.norm(raw_rotation.astype(np.float64), axis=1, keepdims=True), 1e-12)
        return 1e-6 * np.abs(cot).max(1, keepdims=True) / ln
    return 0.0


def _grad_floor(k, gold, key):
    name = {"opacity": "gaussian_opacity", "rotation": "gaussian_rotations", "scaling": "gaussian_scales"}[k]
    return _grad_floor_np(k, gold[f"{key}.cot.{name}"], gold[f"{key}.raw.rotation"])


@pytest.mark.gpu
@pytest.mark.parametrize("key", CASES)
def test_cuda_activations_match_reference_getters(cuda_lib, key):
    from splatfields_b200 import activate_parameters
    dev = torch.device("cuda")
    raw = _raw(key, torch.float32, dev)
    d = activate_parameters(raw["xyz"], raw["scaling"], raw["rotation"], raw["opacity"], raw["f_dc"],
                            raw["f_rest"] if raw["f_rest"].shape[1] > 0 else None)
    assert d["means3D"] is raw["xyz"]
    for k in OUT_KEYS:
        ref = GOLD[f"{key}.out.{k}"]
        got = d[k].detach().cpu().numpy()
        assert got.shape == ref.shape, k
        if k == "gaussian_features":
            assert np.array_equal(got, ref.astype(np.float32))                     # a copy: bit for bit
        else:
            assert np.allclose(got, ref, rtol=2e-6, atol=1e-30), (k, _relerr(got, ref))
    cot = {k: torch.from_numpy(GOLD[f"{key}.cot.{k}"]).float().to(dev) for k in OUT_KEYS}
    sum((d[k] * cot[k]).sum() for k in OUT_KEYS).backward()
    torch.cuda.synchronize()
    for k in RAW_KEYS:
        ref = GOLD[f"{key}.grad.{k}"]
        if ref.size == 0:
            continue
        got = raw[k].grad.cpu().numpy()
        assert got.shape == ref.shape, k
        if k in ("f_dc", "f_rest"):
            assert np.array_equal(got, ref.astype(np.float32)), k
        else:
            g2, r2 = got.reshape(got.shape[0], -1), ref.reshape(ref.shape[0], -1)
            assert np.all(np.abs(g2 - r2) <= 2e-5 * np.abs(r2) + _grad_floor(k, GOLD, key)), (k, _relerr(got, ref))


@pytest.mark.gpu
def test_cuda_activations_scale_offset_and_large(cuda_lib):
    """Dynamic branch epilogue (train.py:73) and a size where every block shape occurs (P not a multiple of 256)."""
    from splatfields_b200 import activate_parameters
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(3)
    P, M = 100_003, 16
    mk = lambda *s: torch.randn(*s, generator=g).to(dev).requires_grad_(True)
    xyz, rs, rr, ro, dc, rest, off = mk(P, 3), mk(P, 3), mk(P, 4), mk(P, 1), mk(P, 1, 3), mk(P, M - 1, 3), mk(P, 3)
    d = activate_parameters(xyz, rs, rr, ro, dc, rest, scale_offset=off)
    ref = TR.gaussian_dict_static(xyz, rs, rr, ro, dc, rest, scale_offset=off)
    for k in OUT_KEYS:
        assert torch.allclose(d[k], ref[k], rtol=2e-6, atol=0), k
    cot = {k: torch.randn(d[k].shape, generator=g).to(dev) for k in OUT_KEYS}
    leaves = (rs, rr, ro, dc, rest, off)
    got = torch.autograd.grad(sum((d[k] * cot[k]).sum() for k in OUT_KEYS), leaves)
    want = torch.autograd.grad(sum((ref[k] * cot[k]).sum() for k in OUT_KEYS), leaves)
    cots = {"scaling": "gaussian_scales", "rotation": "gaussian_rotations", "opacity": "gaussian_opacity"}
    for a, b, name in zip(got, want, ("scaling", "rotation", "opacity", "f_dc", "f_rest", "scale_offset")):
        a2, b2 = a.reshape(P, -1).cpu().numpy(), b.reshape(P, -1).cpu().numpy()
        if name in cots:
            floor = _grad_floor_np(name, cot[cots[name]].cpu().numpy(), rr.detach().cpu().numpy())
            assert np.all(np.abs(a2 - b2) <= 2e-5 * np.abs(b2) + floor), name
        else:
            assert np.array_equal(a2, b2), name            # copies


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["uniform", "clustered", "planar", "duplicates"])
def test_cuda_dist2_matches_bruteforce_bitwise(cuda_lib, oracle, name):
    from splatfields_b200 import distCUDA2
    pts = _clouds()[name]
    ref = oracle.knn3_mean_dist2(pts)
    got = distCUDA2(torch.from_numpy(pts).cuda()).cpu().numpy()
    assert got.shape == ref.shape and got.dtype == np.float32
    assert np.array_equal(got, ref), f"{(got != ref).sum()} of {ref.size} differ; worst rel {_relerr(got, ref):.2e}"


@pytest.mark.gpu
def test_cuda_dist2_tiny_and_ragged_sizes(cuda_lib, oracle):
    from splatfields_b200 import distCUDA2
    rs = np.random.RandomState(5)
    assert distCUDA2(torch.zeros(0, 3, device="cuda")).shape == (0,)
    for P in (1, 2, 3, 4, 5, 31, 32, 33, 1023, 1024, 1025, 32769):
        pts = rs.randn(P, 3).astype(np.float32)
        ref = oracle.knn3_mean_dist2(pts)
        got = distCUDA2(torch.from_numpy(pts).cuda()).cpu().numpy()
        assert np.array_equal(got, ref), P            # inf == inf for P < 4


@pytest.mark.gpu
def test_cuda_dist2_full_size_properties(cuda_lib):
    """1M points (BASELINE config 2's cloud): size-independent properties — translation invariance on a power-of-two
    shift (exact in fp32), permutation equivariance, and agreement with a cKDTree on a random sample."""
    from scipy.spatial import cKDTree
    from splatfields_b200 import distCUDA2
    g = torch.Generator().manual_seed(2)
    pts = ((torch.rand(1_000_000, 3, generator=g) * 2 - 1) * 1.3).cuda()
    d0 = distCUDA2(pts)
    perm = torch.randperm(pts.shape[0], generator=g).cuda()
    assert torch.equal(distCUDA2(pts[perm]), d0[perm])
    grid = torch.round(pts * 4096) / 4096                      # coordinates on a 2^-12 grid: +8 is exact in fp32
    assert torch.equal(distCUDA2(grid + 8.0), distCUDA2(grid))
    sample = torch.randint(0, pts.shape[0], (2000,), generator=g)
    p64 = pts.cpu().double().numpy()
    dd, _ = cKDTree(p64).query(p64[sample.numpy()], k=4)
    ref = (dd[:, 1:] ** 2).mean(1)
    got = d0.cpu().numpy()[sample.numpy()]
    assert np.all(np.abs(got - ref) <= 4e-7 * (1.3 ** 2 * 3 + ref) + 1e-6 * ref)
