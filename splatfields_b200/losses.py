"""Host-side mirror of the reference's photometric loss: utils/loss_utils.py:18-19 `l1_loss`,
:45-54 `ssim` (11x11 Gaussian window, sigma 1.5, zero padding, size_average=True) and their combination at
train.py:183-184 (+ the mask term of :189-193).  Same names and argument meaning; the compute is ONE fused forward
pass + ONE gradient pass of the sm_100a library (csrc/loss.cu) instead of five depthwise convolutions, ~15
elementwise kernels and their autograd replay.  CUDA tensors only — there is no CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib
from .rasterizer import _prep, _ptr


def _chw(t):
    if t.dim() == 2:
        t = t.unsqueeze(0)
    if t.dim() == 4 and t.shape[0] == 1:
        t = t[0]
    if t.dim() != 3:
        raise Exception("expected an image of shape [C, H, W] (or [1, C, H, W] / [H, W])")
    return t


def _run(img, gt, lam, opacity, mask, lam_mask, want_grad):
    lib = _lib.load()
    if not img.is_cuda:
        raise _lib.SplatB200Error("the fused loss runs on CUDA tensors only (no CPU fallback)")
    dev = img.device
    x, y = _prep(_chw(img)), _prep(_chw(gt).to(dev))
    if x.shape != y.shape:
        raise Exception(f"image / ground-truth shapes differ: {tuple(x.shape)} vs {tuple(y.shape)}")
    C, H, W = x.shape
    o = m = None
    if opacity is not None:
        o, m = _prep(opacity.reshape(-1)), _prep(mask.to(dev).reshape(-1))
        if o.numel() != H * W or m.numel() != H * W:
            raise Exception("opacity / gt_mask must have H*W elements")
    f32 = dict(dtype=torch.float32, device=dev)
    scal = torch.empty(4, **f32)
    g_img = torch.empty_like(x) if want_grad else None
    g_op = torch.empty((H * W,), **f32) if (want_grad and o is not None) else None
    scratch = torch.empty(int(lib.sfb_loss_scratch_bytes(C, H, W)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = lib.sfb_l1_ssim_loss(C, H, W, _ptr(x), _ptr(y), float(lam), _ptr(o), _ptr(m), float(lam_mask), 1.0,
                                  _ptr(scal), _ptr(g_img), _ptr(g_op), _ptr(scratch), stream)
    _lib.check(rc)
    return scal, g_img, g_op


class _FusedLoss(torch.autograd.Function):
    """(img, gt, opacity | None, gt_mask | None, lambda_dssim, lambda_mask) -> (loss, l1, ssim, mask_l1) as 0-dim
    tensors; only `loss` is differentiable (w.r.t. img and opacity)."""

    @staticmethod
    def forward(ctx, img, gt, opacity, gt_mask, lambda_dssim, lambda_mask):
        if gt.requires_grad or (gt_mask is not None and gt_mask.requires_grad):
            raise NotImplementedError("the fused loss differentiates w.r.t. the rendered image / opacity only")
        want = img.requires_grad or (opacity is not None and opacity.requires_grad)
        scal, g_img, g_op = _run(img.detach(), gt, lambda_dssim, None if opacity is None else opacity.detach(), gt_mask,
                                 lambda_mask, want)
        ctx.shapes = (img.shape, None if opacity is None else opacity.shape)
        ctx.save_for_backward(*(t for t in (g_img, g_op) if t is not None))
        ctx.has = (g_img is not None, g_op is not None)
        loss, l1, ss, ml1 = scal[3], scal[0], scal[1], scal[2]
        if float(lambda_dssim) == 0.0:     # the SSIM map is not evaluated then: do not report a made-up 0
            ss = torch.full_like(ss, float("nan"))
        ctx.mark_non_differentiable(l1, ss, ml1)
        return loss, l1, ss, ml1

    @staticmethod
    def backward(ctx, g_loss, _g1, _g2, _g3):
        saved = list(ctx.saved_tensors)
        g_img = saved.pop(0) if ctx.has[0] else None
        g_op = saved.pop(0) if ctx.has[1] else None
        gi = None if g_img is None else (g_img * g_loss).reshape(ctx.shapes[0])
        go = None if g_op is None else (g_op * g_loss).reshape(ctx.shapes[1])
        return gi, None, go, None, None, None


def photometric_loss(image, gt_image, lambda_dssim, opacity=None, gt_mask=None, lambda_mask=0.0):
    """train.py:183-184 (+ :189-193 when opacity / gt_mask are given):
        loss = (1 - lambda_dssim) * l1_loss(image, gt) + lambda_dssim * (1 - ssim(image, gt))
               [+ lambda_mask * F.l1_loss(clamp(opacity, 0, 1), gt_mask)]
    Returns (loss, Ll1, ssim, mask_l1): 0-dim tensors; `loss` carries the gradient.  ssim is NaN when
    lambda_dssim == 0 (it is not evaluated then)."""
    if (opacity is None) != (gt_mask is None):
        raise Exception("opacity and gt_mask go together")
    return _FusedLoss.apply(image, gt_image, opacity, gt_mask, float(lambda_dssim), float(lambda_mask))


def l1_loss(network_output, gt):
    """utils/loss_utils.py:18-19."""
    return _FusedLoss.apply(network_output, gt, None, None, 0.0, 0.0)[0]


def ssim(img1, img2, window_size=11, size_average=True):
    """utils/loss_utils.py:45-54 for the configuration the reference uses (window 11, size_average=True)."""
    if window_size != 11 or not size_average:
        raise NotImplementedError("the fused SSIM implements window_size=11, size_average=True (train.py:184)")
    return 1.0 - _FusedLoss.apply(img1, img2, None, None, 1.0, 0.0)[0]
