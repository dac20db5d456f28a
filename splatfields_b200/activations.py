"""Host-side mirror of the step right before the rasterizer: get_gaussian_dict's static branch
(train.py:42-50), i.e. the GaussianModel getters get_scaling / get_rotation / get_opacity / get_features
(scene/gaussian_model.py:64-86), and the `ret['scales'] + scaling` epilogue of the dynamic branch (train.py:73).
One kernel forward, one backward (csrc/activate.cu) instead of exp + normalize + sigmoid + cat (and their four
autograd nodes), each a full pass over its tensor."""
from __future__ import annotations

import torch

from . import _lib
from .rasterizer import _ptr


def _c(t):
    """contiguous fp32 with a 16-byte aligned base"""
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if t.data_ptr() % 16 != 0:
        t = t.clone()
    return t


class _Activate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw_scaling, raw_rotation, raw_opacity, f_dc, f_rest, scale_offset):
        lib = _lib.load()
        if not raw_scaling.is_cuda:
            raise _lib.SplatB200Error("the fused activations run on CUDA tensors only (no CPU fallback)")
        P = raw_rotation.shape[0]
        if raw_rotation.dim() != 2 or raw_rotation.shape[1] != 4:
            raise Exception("raw_rotation must have dimensions (num_points, 4)")
        if raw_scaling.shape[0] != P or raw_scaling.shape[-1] not in (1, 3):
            raise Exception("raw_scaling must have dimensions (num_points, 3) or (num_points, 1)")
        iso = int(raw_scaling.shape[-1] == 1)
        dev = raw_scaling.device
        rs, rr, ro = _c(raw_scaling), _c(raw_rotation), _c(raw_opacity)
        dc, rest, off = _c(f_dc), _c(f_rest), _c(scale_offset)
        M = 0 if dc is None else 1 + (0 if rest is None else rest.shape[1])
        f32 = dict(dtype=torch.float32, device=dev)
        scales, rot, op = torch.empty((P, 3), **f32), torch.empty((P, 4), **f32), torch.empty((P, 1), **f32)
        feats = torch.empty((P, M, 3), **f32) if M > 0 else None
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.sfb_activate_forward(P, M, iso, _ptr(rs), _ptr(rr), _ptr(ro), _ptr(dc), _ptr(rest), _ptr(off),
                                                _ptr(scales), _ptr(rot), _ptr(op), _ptr(feats), stream))
        ctx.save_for_backward(rs, rr, ro)
        ctx.meta = (P, M, iso, raw_scaling.shape, raw_opacity.shape,
                    None if f_dc is None else f_dc.shape, None if f_rest is None else f_rest.shape,
                    scale_offset is not None)
        if feats is None:
            feats = scales.new_empty(0)
        return scales, rot, op, feats

    @staticmethod
    def backward(ctx, g_scales, g_rot, g_op, g_feats):
        lib = _lib.load()
        rs, rr, ro = ctx.saved_tensors
        P, M, iso, sh_scaling, sh_op, sh_dc, sh_rest, has_off = ctx.meta
        dev = rs.device
        f32 = dict(dtype=torch.float32, device=dev)
        need = ctx.needs_input_grad
        gs = _c(g_scales) if (g_scales is not None and (need[0] or (need[5] and has_off))) else None
        gr = _c(g_rot) if (g_rot is not None and need[1]) else None
        go = _c(g_op) if (g_op is not None and need[2]) else None
        gf = _c(g_feats) if (g_feats is not None and M > 0 and (need[3] or need[4])) else None
        d_rs = torch.empty(sh_scaling, **f32) if (need[0] and gs is not None) else None
        d_rr = torch.empty((P, 4), **f32) if gr is not None else None
        d_ro = torch.empty(sh_op, **f32) if go is not None else None
        d_dc = torch.empty(sh_dc, **f32) if (gf is not None and need[3]) else None
        d_rest = torch.empty(sh_rest, **f32) if (gf is not None and need[4] and sh_rest is not None) else None
        if P > 0 and any(t is not None for t in (d_rs, d_rr, d_ro, d_dc, d_rest)):
            with torch.cuda.device(dev):
                stream = torch.cuda.current_stream(dev).cuda_stream
                _lib.check(lib.sfb_activate_backward(
                    P, M, iso, _ptr(rs), _ptr(rr), _ptr(ro), _ptr(gs), _ptr(gr), _ptr(go), _ptr(gf),
                    _ptr(d_rs), _ptr(d_rr), _ptr(d_ro), _ptr(d_dc), _ptr(d_rest), stream))
        d_off = g_scales if (has_off and need[5]) else None       # d(exp(raw) + offset) / d offset = identity
        return d_rs, d_rr, d_ro, d_dc, d_rest, d_off


def activate_parameters(xyz, raw_scaling, raw_rotation, raw_opacity, features_dc=None, features_rest=None,
                        scale_offset=None, active_sh_degree=0) -> dict:
    """The reference's `gaussian_dict` (train.py:42-50) from the raw GaussianModel parameters, in one fused kernel:
        means3D = xyz; gaussian_scales = exp(raw_scaling) [.repeat(1,3) if [P,1]] [+ scale_offset];
        gaussian_rotations = normalize(raw_rotation); gaussian_opacity = sigmoid(raw_opacity);
        gaussian_features = cat(features_dc, features_rest, dim=1).
    Differentiable w.r.t. every raw parameter (and scale_offset)."""
    scales, rot, op, feats = _Activate.apply(raw_scaling, raw_rotation, raw_opacity, features_dc, features_rest,
                                             scale_offset)
    out = {"means3D": xyz, "active_sh_degree": active_sh_degree, "gaussian_opacity": op, "gaussian_scales": scales,
           "gaussian_rotations": rot}
    if features_dc is not None:
        out["gaussian_features"] = feats
    return out
