"""Host-side mirror of the `diff_gaussian_rasterization` Python surface the reference imports at
gaussian_renderer/__init__.py:14 and drives at :59-72, :74, :94-102, :106-114:

    GaussianRasterizationSettings (NamedTuple, 12 fields)
    GaussianRasterizer(nn.Module).forward(means3D, means2D, opacities, shs=None, colors_precomp=None,
                                          scales=None, rotations=None, cov3D_precomp=None)
        -> (color [3,H,W], radii [P] int32, depth [1,H,W])
    GaussianRasterizer.markVisible(positions) -> bool [P]

Same names, argument meaning, return order and error messages; the compute is the sm_100a library behind
the C ABI (include/splat_b200.h).  torch is used for device memory, streams and autograd glue only.
"""
from __future__ import annotations

import os
import threading
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _lib


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


# Optional gradient arena: when set, backward writes the parameter gradients straight into slices of one
# caller-owned flat fp32 slab (the buffer a view-parallel trainer hands to NCCL all-reduce) instead of
# allocating separate tensors.  fields = ((input name, floats per splat), ...) in slab order.
_ARENA = None
# Factored SH gradient (view-parallel exchange, host_api.ViewParallelRasterizer): when a [P, 3] tensor is set here,
# backward writes the clamp-masked colour gradient into it instead of the [P, M, 3] SH gradient rows
# (SFB_BWD_SH_FACTORED); `shs` then gets no .grad from autograd — the caller rebuilds the multi-view sum with
# sh_grad_combine().
_SH_COLOR_OUT = None
# Exchange over NVLink (host_api.ViewParallelRasterizer exchange="nvlink"): (_lib.XchgDesc, epoch[, fused]) for the NEXT
# backward.  The parameter gradients then leave through the symmetric buffers and autograd gets no .grad for the
# parameters — only means2D's (a per-view quantity).  fused=True: the backward call itself sums them over the ranks into
# the arena slab (one persistent kernel: geometry backward + exchange); otherwise sfb_xchg_finish does, afterwards.
_XCHG = None
_WARNED_DEPTH = False
PROPAGATE_DEPTH_GRAD = os.environ.get("SFB_DEPTH_GRAD", "1") != "0"


def set_grad_arena(slab, fields, sh_color_out=None, xchg=None):
    global _ARENA, _SH_COLOR_OUT, _XCHG
    _ARENA = None if slab is None else (slab, tuple(fields))
    _SH_COLOR_OUT = sh_color_out
    _XCHG = xchg


def sh_grad_combine(means3D, campos_views, dcolor_views, sh_degree, out):
    """out[P, M, 3] = sum over views v of basis(normalize(means3D - campos_views[v])) (x) dcolor_views[v]
    (include/splat_b200.h: sfb_sh_grad_combine).  campos_views [V, 3], dcolor_views [V, P, 3]: the factored
    per-view SH gradients; out: a contiguous fp32 CUDA tensor (e.g. the `shs` slice of the gradient slab)."""
    lib = _lib.load()
    if not means3D.is_cuda:
        raise _lib.SplatB200Error("sh_grad_combine runs on CUDA tensors only (no CPU fallback)")
    P = means3D.shape[0]
    V = campos_views.shape[0]
    if dcolor_views.numel() != V * P * 3 or out.numel() % (3 * max(P, 1)) != 0:
        raise Exception("dcolor_views must hold [V, P, 3] floats and out [P, M, 3]")
    M = out.numel() // (3 * P) if P > 0 else 0
    for t, name in ((means3D, "means3D"), (campos_views, "campos_views"), (dcolor_views, "dcolor_views"), (out, "out")):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise Exception(f"{name} must be a contiguous fp32 tensor")
    with torch.cuda.device(means3D.device):
        stream = torch.cuda.current_stream(means3D.device).cuda_stream
        _lib.check(lib.sfb_sh_grad_combine(P, V, int(sh_degree), int(M), _ptr(means3D), _ptr(campos_views),
                                           _ptr(dcolor_views), _ptr(out), stream))
    return out


def _arena_out(name, P, shape, **f32):
    if _ARENA is not None:
        slab, fields = _ARENA
        off = 0
        for fname, n in fields:
            if fname == name:
                numel = 1
                for d in shape:
                    numel *= d
                if numel == n * P and (off + n) * P <= slab.numel():
                    return slab[off * P:(off + n) * P].view(shape)
                break
            off += n
    return torch.empty(shape, **f32)


def _ptr(t):
    return None if t is None or t.numel() == 0 else t.data_ptr()


def _prep(t, name=None):
    """contiguous fp32 CUDA tensor with a 16-byte aligned base (the kernels use 128-bit accesses)."""
    if t is None or t.numel() == 0:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if t.data_ptr() % 16 != 0:
        t = t.clone()
    return t


class _Scratch:
    """The three caller-owned scratch buffers; torch owns the bytes, the library asks via callbacks.
    The ctypes callback objects are created once per thread and reused (building a CFUNCTYPE thunk per
    forward call costs tens of microseconds of pure host time on a ~1 ms step)."""

    _tls = threading.local()

    def __init__(self, device):
        self.device = device
        self.bufs = {}
        tls = _Scratch._tls
        if not hasattr(tls, "cbs"):
            tls.cbs = {name: _lib.ALLOC_FN(_Scratch._make(name)) for name in ("geom", "binning", "img")}
        tls.current = self

    @staticmethod
    def _make(name):
        def alloc(_user, nbytes):
            cur = _Scratch._tls.current
            t = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=cur.device)
            cur.bufs[name] = t
            return t.data_ptr()
        return alloc

    def cb(self, name):
        return _Scratch._tls.cbs[name]


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings, with_alpha=False):
    out = _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                    cov3Ds_precomp, raster_settings, with_alpha)
    return out if with_alpha else out[:3]


def _forward_impl(means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs, with_alpha=False):
    lib = _lib.load()
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise Exception("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise _lib.SplatB200Error("the rasterizer runs on CUDA tensors only (no CPU fallback)")
    dev = means3D.device
    P = means3D.shape[0]
    H, W = int(rs.image_height), int(rs.image_width)
    means3D_c = _prep(means3D)
    sh_c, col_c = _prep(sh), _prep(colors_precomp)
    op_c, sc_c, rot_c, cov_c = _prep(opacities), _prep(scales), _prep(rotations), _prep(cov3Ds_precomp)
    bg, vm, pm, cp = _prep(rs.bg.to(dev)), _prep(rs.viewmatrix.to(dev)), _prep(rs.projmatrix.to(dev)), _prep(rs.campos.to(dev))
    M = 0 if sh_c is None else (sh_c.shape[1] if sh_c.dim() == 3 else sh_c.numel() // (3 * P))
    color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
    depth = torch.empty((1, H, W), dtype=torch.float32, device=dev)
    radii = torch.empty((P,), dtype=torch.int32, device=dev)
    alpha = torch.empty((1, H, W), dtype=torch.float32, device=dev) if with_alpha else None
    scratch = _Scratch(dev)
    import ctypes as C
    nr = C.c_int(0)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = lib.sfb_rasterize_forward(
            P, int(rs.sh_degree), int(M), W, H,
            _ptr(bg), _ptr(means3D_c), _ptr(sh_c), _ptr(col_c), _ptr(op_c), _ptr(sc_c), float(rs.scale_modifier),
            _ptr(rot_c), _ptr(cov_c), _ptr(vm), _ptr(pm), _ptr(cp), float(rs.tanfovx), float(rs.tanfovy),
            int(bool(rs.prefiltered)), _ptr(color), _ptr(depth), _ptr(alpha), _ptr(radii),
            scratch.cb("geom"), None, scratch.cb("binning"), None, scratch.cb("img"), None,
            C.byref(nr), int(bool(rs.debug)), stream)
    _lib.check(rc)
    empty = torch.empty(0, dtype=torch.uint8, device=dev)
    geom, binning, img = (scratch.bufs.get(k, empty) for k in ("geom", "binning", "img"))
    saved = (means3D_c, sh_c, col_c, sc_c, rot_c, cov_c, bg, vm, pm, cp)
    return int(nr.value), color, depth, radii, geom, binning, img, M, saved, alpha


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings, with_alpha=False):
        rs = raster_settings
        args = (means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs)
        if rs.debug:
            try:
                out = _forward_impl(*args, with_alpha=with_alpha)
            except Exception as ex:
                torch.save(tuple(a.detach().cpu() if torch.is_tensor(a) else a for a in args[:-1]), "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            out = _forward_impl(*args, with_alpha=with_alpha)
        num_rendered, color, depth, radii, geom, binning, img, M, saved, alpha = out
        ctx.with_alpha = bool(with_alpha)
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.M = M
        ctx.shapes = tuple(None if t is None else t.shape for t in
                           (means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp))
        ctx.acc_fresh = [True]      # the forward left the gradient accumulators cleared; true for ONE backward
        # the prepared inputs go through save_for_backward (None entries are allowed): autograd's version counters then
        # catch an in-place update between forward and backward instead of silently using the new values
        ctx.save_for_backward(radii, geom, binning, img, *saved)
        ctx.mark_non_differentiable(radii)
        ctx.set_materialize_grads(False)     # unused outputs arrive as None in backward (no zero tensors to allocate)
        if with_alpha:
            return color, radii, depth, alpha
        # always four outputs so that backward has a fixed arity; the wrapper drops the placeholder
        return color, radii, depth, color.new_empty(0)

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii, _grad_depth, grad_out_alpha=None):
        # The depth image is differentiable: its cotangent goes to the kernels as one more composited channel (the
        # reference's depth losses, train.py:195-229).  Whether the pinned depth-diff-gaussian-rasterization build
        # propagates it is not checkable from the reference tree (every shipped recipe keeps
        # lambda_depth = 0, where both behaviours coincide).  SFB_DEPTH_GRAD=0 restores a forward-only depth (the
        # cotangent is dropped with a one-time warning).
        global _WARNED_DEPTH
        if _grad_depth is not None and not PROPAGATE_DEPTH_GRAD:
            if not _WARNED_DEPTH:
                _WARNED_DEPTH = True
                import warnings
                warnings.warn("splatfields_b200: SFB_DEPTH_GRAD=0 - the gradient of a depth loss is NOT propagated "
                              "to the Gaussians", RuntimeWarning, stacklevel=2)
            _grad_depth = None
        ctx.grad_depth = _grad_depth
        lib = _lib.load()
        rs = ctx.raster_settings
        radii, geom, binning, img, means3D, sh, col, sc, rot, cov, bg, vm, pm, cp = ctx.saved_tensors
        sh_means3D, sh_means2D, sh_sh, sh_col, sh_op, sh_sc, sh_rot, sh_cov = ctx.shapes
        dev = means3D.device
        P = means3D.shape[0]
        H, W = int(rs.image_height), int(rs.image_width)
        M = ctx.M
        f32 = dict(dtype=torch.float32, device=dev)
        if _XCHG is not None:
            return _RasterizeGaussians._backward_exchange(ctx, grad_out_color, grad_out_alpha, lib)
        dL_dmeans3D = _arena_out("means3D", P, (P, 3), **f32)
        dL_dmeans2D = torch.empty((P, 3), **f32)
        dL_dcolors = _arena_out("colors_precomp", P, (P, 3), **f32) if col is not None else None
        dL_dopacity = _arena_out("opacities", P, (P, 1), **f32)
        dL_dcov3D = _arena_out("cov3D_precomp", P, (P, 6), **f32) if cov is not None else None
        factored = sh is not None and _SH_COLOR_OUT is not None
        if factored:
            dL_dcolors, dL_dsh = _SH_COLOR_OUT, None
            if dL_dcolors.numel() != 3 * P or dL_dcolors.dtype != torch.float32 or not dL_dcolors.is_contiguous():
                raise Exception("sh_color_out must be a contiguous fp32 tensor of P * 3 elements")
        else:
            dL_dsh = _arena_out("shs", P, (P, M, 3), **f32) if sh is not None else None
        dL_dscales = _arena_out("scales", P, (P, 3), **f32) if cov is None else None
        dL_drot = _arena_out("rotations", P, (P, 4), **f32) if cov is None else None
        g = _prep(grad_out_color)
        ga = _prep(grad_out_alpha) if (ctx.with_alpha and grad_out_alpha is not None) else None
        gd = _prep(ctx.grad_depth)
        if g is None:      # only the alpha / depth image was used downstream
            g = torch.zeros((3, H, W), **f32)
        flags = (_lib.BWD_ACC_FRESH if ctx.acc_fresh[0] else 0) | (_lib.BWD_SH_FACTORED if factored else 0)
        ctx.acc_fresh[0] = False    # a second backward on the same buffers (retain_graph) must clear them itself
        if P > 0:
            with torch.cuda.device(dev):
                stream = torch.cuda.current_stream(dev).cuda_stream
                rc = lib.sfb_rasterize_backward(
                    P, int(rs.sh_degree), int(M), int(ctx.num_rendered), W, H,
                    _ptr(bg), _ptr(means3D), _ptr(sh), _ptr(col), _ptr(sc), float(rs.scale_modifier), _ptr(rot),
                    _ptr(cov), _ptr(vm), _ptr(pm), _ptr(cp), float(rs.tanfovx), float(rs.tanfovy), _ptr(radii),
                    _ptr(geom), _ptr(binning), _ptr(img), _ptr(g), _ptr(ga), _ptr(gd),
                    _ptr(dL_dmeans2D), _ptr(dL_dcolors), _ptr(dL_dopacity), _ptr(dL_dmeans3D), _ptr(dL_dcov3D),
                    _ptr(dL_dsh), _ptr(dL_dscales), _ptr(dL_drot), int(bool(rs.debug)), flags, None, 0, stream)
            _lib.check(rc)

        def shaped(t, shape):
            return None if (t is None or shape is None) else t.reshape(shape)
        grads = (
            shaped(dL_dmeans3D, sh_means3D),
            shaped(dL_dmeans2D, sh_means2D),
            shaped(dL_dsh, sh_sh),
            shaped(dL_dcolors, sh_col) if col is not None else None,
            shaped(dL_dopacity, sh_op),
            shaped(dL_dscales, sh_sc),
            shaped(dL_drot, sh_rot),
            shaped(dL_dcov3D, sh_cov) if cov is not None else None,
            None,
            None,
        )
        return grads


    @staticmethod
    def _backward_exchange(ctx, grad_out_color, grad_out_alpha, lib):
        """Backward in exchange mode (_XCHG): packed gradient records + pushed colour gradients, no per-parameter
        outputs (include/splat_b200.h: sfb_xchg)."""
        import ctypes as C
        desc, epoch = _XCHG[0], _XCHG[1]
        fused = len(_XCHG) > 2 and bool(_XCHG[2])
        rs = ctx.raster_settings
        radii, geom, binning, img, means3D, sh, col, sc, rot, cov, bg, vm, pm, cp = ctx.saved_tensors
        dev = means3D.device
        P = means3D.shape[0]
        H, W = int(rs.image_height), int(rs.image_width)
        if cov is not None:
            raise Exception("the gradient exchange needs scales / rotations (no cov3D_precomp)")
        dL_dmeans2D = torch.empty((P, 3), dtype=torch.float32, device=dev)
        g = _prep(grad_out_color)
        if g is None:
            g = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
        ga = _prep(grad_out_alpha) if (ctx.with_alpha and grad_out_alpha is not None) else None
        gd = _prep(ctx.grad_depth)
        flags = _lib.BWD_ACC_FRESH if ctx.acc_fresh[0] else 0
        ctx.acc_fresh[0] = False
        o = dict.fromkeys(("means3D", "opacities", "scales", "rotations", "colors_precomp", "shs"))
        if fused:       # the sums over the ranks land in the caller's slab (set_grad_arena)
            if _ARENA is None:
                raise Exception("fused exchange needs a gradient arena (set_grad_arena)")
            f32 = dict(dtype=torch.float32, device=dev)
            o["means3D"] = _arena_out("means3D", P, (P, 3), **f32)
            o["opacities"] = _arena_out("opacities", P, (P, 1), **f32)
            o["scales"] = _arena_out("scales", P, (P, 3), **f32)
            o["rotations"] = _arena_out("rotations", P, (P, 4), **f32)
            if sh is not None:
                o["shs"] = _arena_out("shs", P, (P, ctx.M, 3), **f32)
            else:
                o["colors_precomp"] = _arena_out("colors_precomp", P, (P, 3), **f32)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = lib.sfb_rasterize_backward(
                P, int(rs.sh_degree), int(ctx.M), int(ctx.num_rendered), W, H,
                _ptr(bg), _ptr(means3D), _ptr(sh), _ptr(col), _ptr(sc), float(rs.scale_modifier), _ptr(rot),
                None, _ptr(vm), _ptr(pm), _ptr(cp), float(rs.tanfovx), float(rs.tanfovy), _ptr(radii),
                _ptr(geom), _ptr(binning), _ptr(img), _ptr(g), _ptr(ga), _ptr(gd),
                _ptr(dL_dmeans2D), _ptr(o["colors_precomp"]), _ptr(o["opacities"]), _ptr(o["means3D"]), None,
                _ptr(o["shs"]), _ptr(o["scales"]), _ptr(o["rotations"]), int(bool(rs.debug)), flags,
                C.byref(desc), int(epoch), stream)
        _lib.check(rc)
        sh_means2D = ctx.shapes[1]
        return (None, dL_dmeans2D.reshape(sh_means2D) if sh_means2D is not None else None,
                None, None, None, None, None, None, None, None)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            lib = _lib.load()
            pos = _prep(positions)
            P = positions.shape[0]
            out = torch.empty((P,), dtype=torch.uint8, device=positions.device)
            vm, pm = _prep(rs.viewmatrix.to(positions.device)), _prep(rs.projmatrix.to(positions.device))
            if P > 0:
                with torch.cuda.device(positions.device):
                    stream = torch.cuda.current_stream(positions.device).cuda_stream
                    _lib.check(lib.sfb_mark_visible(P, _ptr(pos), _ptr(vm), _ptr(pm), _ptr(out), stream))
            return out.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, with_alpha=False):
        """Reference signature plus one extension: `with_alpha=True` appends a fused coverage image
        alpha[1,H,W] = sum(alpha_i * T_i) to the outputs — the image the reference's render() obtains from
        a second full rasterizer call with colours = 1 and bg = 0 (gaussian_renderer/__init__.py:104-115)."""
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, rs, with_alpha=with_alpha)
