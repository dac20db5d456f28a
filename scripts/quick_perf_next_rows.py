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
    n0 = _lib.load().sfb_profile_count(0)
    _lib.profile_enable(True)
    ours()
    torch.cuda.synchronize()
    rows = [r for r in _lib.profile_read(0)[n0:] if r[0].startswith("loss.")]
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


if __name__ == "__main__":
    main()
