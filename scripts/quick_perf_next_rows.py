"""Device-side timing of the SURVEY §8f rows next to the rasterizer (development aid / profiles/ evidence):
the fused photometric loss and the densification bookkeeping against the reference's own statements executed by
PyTorch on the same GPU (oracle/torch_next_rows.py).  One JSON line per case."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import torch_next_rows as TR            # baseline leg only
from splatfields_b200 import _lib, densify, losses

PEAK = None


def peak_gbs():
    global PEAK
    if PEAK is None:
        try:
            PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            PEAK = 6500.0
    return PEAK


def timed(fn, iters, warmup, flush):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for i in range(warmup + iters):
        flush.add_(1.0)                 # 512 MB read-modify-write: evicts the 126 MB L2 between iterations
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= warmup:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def loss_case(C, H, W, iters, warmup, flush, with_mask):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    img = torch.rand(C, H, W, device=dev, generator=g)
    gt = (img + 0.1 * torch.randn(C, H, W, device=dev, generator=g)).clamp(0, 1)
    op = torch.rand(1, H, W, device=dev, generator=g) if with_mask else None
    mk = (torch.rand(1, H, W, device=dev, generator=g) > 0.5).float() if with_mask else None
    lam, lam_m = 0.2, 0.1

    def ours():
        a = img.detach().requires_grad_(True)
        b = op.detach().requires_grad_(True) if with_mask else None
        losses.photometric_loss(a, gt, lam, b, mk, lam_m)[0].backward()

    def ref():
        a = img.detach().requires_grad_(True)
        b = op.detach().requires_grad_(True) if with_mask else None
        TR.photometric_loss(a, gt, lam, b, mk, lam_m)[0].backward()

    t_ours = timed(ours, iters, warmup, flush)
    t_ref = timed(ref, iters, warmup, flush)
    # per-kernel device times: the loss kernels append their event pairs to the library's current record list
    _lib.profile_enable(True)           # (starts a fresh record list)
    ours()
    torch.cuda.synchronize()
    rows = [r for r in _lib.profile_read(0) if r[0].startswith("loss.")]
    _lib.profile_enable(False)
    kern = {k: round(v * 1e3, 2) for k, v in rows}
    N = C * H * W
    alg = 44 * N + (16 * H * W if with_mask else 0)      # bytes: 2+3 floats (stats) + 5+1 floats (grad) per pixel-channel
    k_us = sum(kern.values())
    return dict(case="l1_ssim_loss", C=C, H=H, W=W, with_mask=with_mask, fused_ms=round(t_ours, 4),
                torch_eager_ms=round(t_ref, 4), speedup=round(t_ref / t_ours, 2), kernels_us=kern,
                algorithmic_bytes=alg, kernels_GBps=round(alg / (k_us * 1e-6) / 1e9, 1) if k_us > 0 else None,
                frac_of_measured_hbm=round(alg / (k_us * 1e-6) / 1e9 / peak_gbs(), 3) if k_us > 0 else None)


def densify_case(P, iters, warmup, flush):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(2)
    grad = torch.randn(P, 3, device=dev, generator=g) * 3e-4
    radii = (torch.rand(P, device=dev, generator=g) * 40 - 8).clamp(min=0).to(torch.int32)
    acc, den, mr = torch.zeros(P, 1, device=dev), torch.zeros(P, 1, device=dev), torch.zeros(P, device=dev)
    raw_s, raw_o = torch.randn(P, 3, device=dev, generator=g) - 4.0, torch.randn(P, 1, device=dev, generator=g) * 3

    def ours():
        densify.add_densification_stats(acc, den, grad, radii=radii, max_radii2D=mr)

    def ref():
        TR.add_densification_stats(acc, den, mr, grad, radii)

    def ours_masks():
        densify.densify_masks(acc, den, raw_s, raw_o, mr, 0.0002, 0.01, 4.7, 0.005, 20, raw=True)

    def ref_masks():      # scene/gaussian_model.py:411-422 + :360-363 + :397-400, predicates only
        grads = acc / den
        grads[grads.isnan()] = 0.0
        sc, opa = torch.exp(raw_s), torch.sigmoid(raw_o)
        smax = torch.max(sc, dim=1).values
        c = torch.logical_and(torch.norm(grads, dim=-1) >= 0.0002, smax <= 0.01 * 4.7)
        s = torch.logical_and(grads.squeeze() >= 0.0002, smax > 0.01 * 4.7)
        p = torch.logical_or(torch.logical_or((opa < 0.005).squeeze(), mr > 20), smax > 0.1 * 4.7)
        return c, s, p

    t1, t2 = timed(ours, iters, warmup, flush), timed(ref, iters, warmup, flush)
    t3, t4 = timed(ours_masks, iters, warmup, flush), timed(ref_masks, iters, warmup, flush)
    return dict(case="densify", P=P, stats_fused_ms=round(t1, 4), stats_torch_eager_ms=round(t2, 4),
                stats_speedup=round(t2 / t1, 2), masks_fused_ms=round(t3, 4), masks_torch_eager_ms=round(t4, 4),
                masks_speedup=round(t4 / t3, 2), stats_GBps=round(36.0 * P / (t1 * 1e-3) / 1e9, 1),
                masks_GBps=round(31.0 * P / (t3 * 1e-3) / 1e9, 1))


def activations_case(P, iters, warmup, flush):
    """§8f-3: GaussianModel getters (exp / normalize / sigmoid / cat) + their autograd, fused vs torch eager."""
    from splatfields_b200 import activate_parameters
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(3)
    mk = lambda *s: torch.randn(*s, device=dev, generator=g).requires_grad_(True)
    xyz, rs, rr, ro, dc, rest = mk(P, 3), mk(P, 3), mk(P, 4), mk(P, 1), mk(P, 1, 3), mk(P, 15, 3)
    keys = ("gaussian_opacity", "gaussian_features", "gaussian_scales", "gaussian_rotations")
    cot = None

    def run(fn):
        def f():
            nonlocal cot
            d = fn(xyz, rs, rr, ro, dc, rest)
            if cot is None:
                cot = {k: torch.randn_like(d[k]) for k in keys}
            torch.autograd.grad([d[k] for k in keys], (rs, rr, ro, dc, rest), [cot[k] for k in keys])
        return f
    t_ours = timed(run(activate_parameters), iters, warmup, flush)
    t_ref = timed(run(TR.gaussian_dict_static), iters, warmup, flush)
    alg = P * (8 + 48) * 4 * 2 * 2            # (8 + 3M) floats in and out, forward and backward
    return dict(case="activations", P=P, fused_fwd_bwd_ms=round(t_ours, 4), torch_eager_fwd_bwd_ms=round(t_ref, 4),
                speedup=round(t_ref / t_ours, 2), algorithmic_bytes=alg,
                GBps=round(alg / (t_ours * 1e-3) / 1e9, 1), frac_of_measured_hbm=round(alg / (t_ours * 1e-3) / 1e9 / peak_gbs(), 3))


def knn_case(P, iters, warmup, flush, clustered=False):
    """§8f-5: distCUDA2 replacement on an init-like cloud (uniform in the +-1.3 cube, dataset_readers.py:598)."""
    from splatfields_b200 import distCUDA2
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(4)
    pts = (torch.rand(P, 3, device=dev, generator=g) * 2 - 1) * 1.3
    if clustered:       # half of the points in 64 tight blobs: strongly non-uniform density
        c = (torch.rand(64, 3, device=dev, generator=g) * 2 - 1) * 1.3
        pts[: P // 2] = c[torch.randint(0, 64, (P // 2,), device=dev, generator=g)] + \
            torch.randn(P // 2, 3, device=dev, generator=g) * 0.004
    t = timed(lambda: distCUDA2(pts), iters, warmup, flush)
    _lib.profile_enable(True)
    distCUDA2(pts)
    torch.cuda.synchronize()
    rows = [r for r in _lib.profile_read(0) if r[0].startswith("knn.")]
    _lib.profile_enable(False)
    kern = {}
    for k, v in rows:
        kern[k] = round(kern.get(k, 0.0) + v * 1e3, 2)
    return dict(case="distCUDA2", P=P, clustered=clustered, ms=round(t, 4), Mpoints_s=round(P / (t * 1e-3) / 1e6, 1),
                kernels_us=kern)


def combine_case(P, V, iters, warmup, flush):
    """View-parallel SH-gradient rebuild (sfb_sh_grad_combine): 12 + 12 V bytes in, 192 out per splat."""
    from splatfields_b200 import rasterizer
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    means = (torch.rand(P, 3, device=dev, generator=g) * 2 - 1) * 1.3
    campos = torch.randn(V, 3, device=dev, generator=g) * 3.0
    dcol = torch.randn(V, P, 3, device=dev, generator=g)
    dcol[:, ::5] = 0.0                       # ~20% of the splats culled in every view
    out = torch.empty(P, 16, 3, device=dev)
    t = timed(lambda: rasterizer.sh_grad_combine(means, campos, dcol, 3, out), iters, warmup, flush)
    alg = P * (12 + 12 * V + 192)
    return dict(case="sh_grad_combine", P=P, V=V, ms=round(t, 4), algorithmic_bytes=alg,
                GBps=round(alg / (t * 1e-3) / 1e9, 1), frac_of_measured_hbm=round(alg / (t * 1e-3) / 1e9 / peak_gbs(), 3))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    a = ap.parse_args()
    flush = torch.zeros(128 * 1024 * 1024, device="cuda:0")
    for (C, H, W, m) in ((3, 800, 800, False), (3, 800, 800, True), (3, 1200, 1600, False), (3, 1080, 1920, True)):
        print(json.dumps(loss_case(C, H, W, a.iters, a.warmup, flush, m)), flush=True)
    for P in (1_000_000, 2_000_000):
        print(json.dumps(densify_case(P, a.iters, a.warmup, flush)), flush=True)
    for P in (1_000_000, 2_000_000):
        print(json.dumps(activations_case(P, a.iters, a.warmup, flush)), flush=True)
    for P, cl in ((100_000, False), (1_000_000, False), (1_000_000, True), (2_000_000, False)):
        print(json.dumps(knn_case(P, a.iters, a.warmup, flush, cl)), flush=True)
    for V in (2, 4, 8):
        print(json.dumps(combine_case(1_000_000, V, a.iters, a.warmup, flush)), flush=True)


if __name__ == "__main__":
    main()
