"""GPU: the SURVEY §8f rows (fused photometric loss, densification bookkeeping) through the C ABI / Python mirror,
against (1) golden vectors produced by the reference's own code, (2) the numpy oracle at BASELINE image sizes,
(3) the same statements executed by PyTorch on the same GPU.
Tolerances: loss scalars 2e-6 abs; loss gradient 2e-4 of its largest magnitude (fp32 filtering vs the fp64
reference; the reference's own fp32 run differs from fp64 by more, see test_next_rows_oracle); statistics and masks
bit-exact."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "next_rows.npz"))


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _t(a, dev, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    return t.requires_grad_(True) if grad else t


@pytest.mark.parametrize("key", ["a64", "b64", "c64", "d64", "e64"])
def test_loss_matches_reference_golden(key, cuda_lib):
    from splatfields_b200 import losses
    dev = _dev()
    img, gt = _t(GOLD[f"{key}.img"], dev, True), _t(GOLD[f"{key}.gt"], dev)
    lam, lam_m = float(GOLD[f"{key}.lambda"]), float(GOLD[f"{key}.lambda_mask"])
    has_mask = f"{key}.opacity" in GOLD.files
    op = _t(GOLD[f"{key}.opacity"], dev, True) if has_mask else None
    mk = _t(GOLD[f"{key}.mask"], dev) if has_mask else None
    loss, l1, ss, ml1 = losses.photometric_loss(img, gt, lam, op, mk, lam_m)
    loss.backward()
    assert abs(float(l1) - float(GOLD[f"{key}.l1"])) < 2e-6
    if lam != 0.0:
        assert abs(float(ss) - float(GOLD[f"{key}.ssim"])) < 2e-6
    assert abs(float(loss) - float(GOLD[f"{key}.loss"])) < 2e-6
    ref = GOLD[f"{key}.dL_dimg"]
    err = np.abs(img.grad.cpu().numpy().astype(np.float64) - ref).max()
    assert err <= 2e-4 * np.abs(ref).max(), (err, np.abs(ref).max())
    if has_mask:
        assert abs(float(ml1) - float(GOLD[f"{key}.mask_l1"])) < 2e-6
        refo = GOLD[f"{key}.dL_dopacity"]
        assert np.abs(op.grad.cpu().numpy() - refo).max() <= 1e-7 * max(1.0, np.abs(refo).max())


def test_l1_and_ssim_mirror_names(cuda_lib):
    """utils/loss_utils.l1_loss / ssim as separate calls (train.py:183-184 written the reference's way)."""
    from splatfields_b200 import losses
    dev = _dev()
    img, gt = _t(GOLD["a64.img"], dev, True), _t(GOLD["a64.gt"], dev)
    lam = 0.2
    loss = (1.0 - lam) * losses.l1_loss(img, gt) + lam * (1.0 - losses.ssim(img, gt))
    loss.backward()
    assert abs(float(loss) - float(GOLD["a64.loss"])) < 2e-6
    ref = GOLD["a64.dL_dimg"]
    assert np.abs(img.grad.cpu().numpy() - ref).max() <= 2e-4 * np.abs(ref).max()
    with pytest.raises(NotImplementedError):
        losses.ssim(img, gt, window_size=7)
    with pytest.raises(Exception):
        losses.l1_loss(img.detach().cpu(), gt.cpu())         # no CPU fallback


def test_loss_full_size_vs_oracle_and_torch(cuda_lib):
    """BASELINE image size (3 x 800 x 800): the fused kernels against the fp64 numpy oracle and against the
    reference's statements executed by PyTorch on the same GPU; the scalars are bit-reproducible run to run."""
    from oracle import next_rows as NR
    from oracle import torch_next_rows as TR
    from splatfields_b200 import losses
    dev = _dev()
    g = torch.Generator().manual_seed(7)
    H = W = 800
    # blurred noise + a white-background corner: mixes textured and perfectly flat regions
    base = torch.rand(3, H // 8, W // 8, generator=g)
    img = torch.nn.functional.interpolate(base[None], size=(H, W), mode="bilinear", align_corners=False)[0]
    img = (img + 0.05 * torch.rand(3, H, W, generator=g)).clamp(0, 1)
    gt = (img + 0.08 * torch.randn(3, H, W, generator=g)).clamp(0, 1)
    img[:, :200, :300] = 1.0
    gt[:, :200, :300] = 1.0
    op = torch.rand(1, H, W, generator=g) * 1.2 - 0.1
    mk = (torch.rand(1, H, W, generator=g) > 0.5).float()
    lam, lam_m = 0.2, 0.1
    o = NR.l1_ssim_loss(img.numpy(), gt.numpy(), lam, op[0].numpy(), mk[0].numpy(), lam_m)
    a, b = img.to(dev).requires_grad_(True), op.to(dev).requires_grad_(True)
    loss, l1, ss, ml1 = losses.photometric_loss(a, gt.to(dev), lam, b, mk.to(dev), lam_m)
    loss.backward()
    assert abs(float(loss) - o["loss"]) < 2e-6 and abs(float(ss) - o["ssim"]) < 2e-6 and abs(float(l1) - o["l1"]) < 2e-6
    ref = o["dL_dimg"]
    err = np.abs(a.grad.cpu().numpy() - ref).max()
    assert err <= 2e-4 * np.abs(ref).max(), (err, np.abs(ref).max())
    assert np.abs(b.grad.cpu().numpy()[0] - o["dL_dopacity"]).max() <= 1e-7 * np.abs(o["dL_dopacity"]).max() + 1e-12
    # same statements, PyTorch eager on the GPU (fp32 convolutions: looser — it is the noisier of the two)
    a2, b2 = img.to(dev).requires_grad_(True), op.to(dev).requires_grad_(True)
    tl, _ = TR.photometric_loss(a2, gt.to(dev), lam, b2, mk.to(dev), lam_m)
    tl.backward()
    assert abs(float(tl) - float(loss)) < 5e-6
    assert (a2.grad - a.grad).abs().max().item() <= 3e-3 * np.abs(ref).max()
    # determinism of the reduction
    again = losses.photometric_loss(img.to(dev), gt.to(dev), lam, op.to(dev), mk.to(dev), lam_m)[0]
    assert float(again) == float(loss)


def test_loss_linearity_in_upstream_gradient(cuda_lib):
    from splatfields_b200 import losses
    dev = _dev()
    img, gt = _t(GOLD["c64.img"], dev, True), _t(GOLD["c64.gt"], dev)
    (losses.photometric_loss(img, gt, 0.2)[0] * 0.125).backward()       # 1/8 views, train.py:242
    ref = GOLD["c64.dL_dimg"]
    # c64 was generated with lambda = 1; rebuild the lambda = 0.2 reference from the oracle
    from oracle import next_rows as NR
    o = NR.l1_ssim_loss(GOLD["c64.img"], GOLD["c64.gt"], 0.2, grad_scale=0.125)
    assert np.abs(img.grad.cpu().numpy() - o["dL_dimg"]).max() <= 2e-4 * np.abs(o["dL_dimg"]).max()
    assert ref.shape == o["dL_dimg"].shape


@pytest.mark.parametrize("key", ["dn0", "dn1"])
def test_densify_matches_reference_golden(key, cuda_lib):
    from splatfields_b200 import densify
    dev = _dev()
    P = GOLD[f"{key}.raw_scaling"].shape[0]
    acc = torch.zeros(P, 1, device=dev)
    den = torch.zeros(P, 1, device=dev)
    mr = torch.zeros(P, device=dev)
    for v in range(3):
        radii = torch.from_numpy(GOLD[f"{key}.view{v}.radii"]).to(dev)
        grad = torch.from_numpy(GOLD[f"{key}.view{v}.grad"]).to(dev)
        if v == 1:      # explicit filter instead of radii > 0
            densify.add_densification_stats(acc, den, grad, update_filter=radii > 0, radii=radii, max_radii2D=mr)
        else:
            densify.add_densification_stats(acc, den, grad, radii=radii, max_radii2D=mr)
        # CPU-torch golden: the norm may round one ulp differently from the GPU formula (fma vs separate products)
        ga = GOLD[f"{key}.view{v}.accum"]
        assert np.abs(acc.cpu().numpy() - ga).max() <= 2.5e-7 * np.abs(ga).max()
        assert np.array_equal(den.cpu().numpy(), GOLD[f"{key}.view{v}.denom"])
        assert np.array_equal(mr.cpu().numpy(), GOLD[f"{key}.view{v}.max_radii2D"])
    thr, pd, ext, mino, mss = (float(x) for x in GOLD[f"{key}.params"])
    # the predicates are compared on the golden statistics themselves (a one-ulp difference in accum could flip a
    # Gaussian sitting exactly on the gradient threshold)
    acc = torch.from_numpy(GOLD[f"{key}.view2.accum"]).to(dev)
    den = torch.from_numpy(GOLD[f"{key}.view2.denom"]).to(dev)
    raw_s, raw_o = torch.from_numpy(GOLD[f"{key}.raw_scaling"]).to(dev), torch.from_numpy(GOLD[f"{key}.raw_opacity"]).to(dev)
    for raw in (True, False):
        s = raw_s if raw else torch.exp(raw_s)
        o = raw_o if raw else torch.sigmoid(raw_o)
        clone, split, prune, counts = densify.densify_masks(acc, den, s, o, mr, thr, pd, ext, mino, mss or None, raw=raw)
        assert np.array_equal(clone.cpu().numpy(), GOLD[f"{key}.clone"])
        assert np.array_equal(split.cpu().numpy(), GOLD[f"{key}.split"])
        assert np.array_equal(prune.cpu().numpy(), GOLD[f"{key}.prune"])
        assert counts.cpu().tolist() == [int(GOLD[f"{key}.clone"].sum()), int(GOLD[f"{key}.split"].sum()),
                                         int(GOLD[f"{key}.prune"].sum())]


def test_densify_stats_vs_torch_statements_200k(cuda_lib):
    """The fused update against the reference's own statements run by PyTorch on the GPU, fed by a real rasterizer
    backward (viewspace_points.grad and radii of a lego-like scene through render())."""
    from oracle import torch_next_rows as TR
    from splatfields_b200 import densify, render, synth
    dev = _dev()
    P, H, W = 200_000, 400, 400
    sc = {k: v.to(dev) for k, v in synth.make_scene(P, 5).items()}
    cam = synth.orbit_camera(1, H, W).to(dev)
    gd = dict(means3D=sc["means3D"].requires_grad_(True), active_sh_degree=3, gaussian_opacity=sc["opacities"],
              gaussian_features=sc["shs"], gaussian_scales=sc["scales"], gaussian_rotations=sc["rotations"])
    out = render(cam, gd, None, torch.ones(3, device=dev), return_opacity=False)
    (out["render"] * torch.rand(3, H, W, device=dev)).sum().backward()
    grad, radii = out["viewspace_points"].grad, out["radii"]
    assert grad is not None and grad.abs().sum().item() > 0
    acc_a, den_a, mr_a = torch.zeros(P, 1, device=dev), torch.zeros(P, 1, device=dev), torch.zeros(P, device=dev)
    acc_b, den_b, mr_b = acc_a.clone(), den_a.clone(), mr_a.clone()
    for _ in range(2):
        densify.add_densification_stats(acc_a, den_a, grad, radii=radii, max_radii2D=mr_a)
        TR.add_densification_stats(acc_b, den_b, mr_b, grad, radii)
    # bit for bit what the reference's statements give on this GPU (torch.norm on CUDA rounds the products separately)
    assert torch.equal(acc_a, acc_b) and torch.equal(den_a, den_b) and torch.equal(mr_a, mr_b)
    assert int(den_a.sum().item()) == 2 * int((radii > 0).sum().item()) > 0
