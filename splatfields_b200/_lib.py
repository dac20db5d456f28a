"""ctypes binding of libsplat_b200.so (the C ABI in include/splat_b200.h).

There is deliberately NO fallback: if the library is missing or does not load, importing the product
path raises.  (The CPU oracle under oracle/ is test infrastructure and is never imported from here.)
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SFB_LIB_VARIANT selects an experimental build variant of the same library (build.py VARIANTS); default: the product
_VARIANT = os.environ.get("SFB_LIB_VARIANT", "")
LIB_PATH = os.path.join(_HERE, "libsplat_b200" + ("_" + _VARIANT if _VARIANT else "") + ".so")

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)

_lib = None

# every symbol include/splat_b200.h declares
SYMBOLS = (
    "sfb_abi_version", "sfb_last_error", "sfb_rasterize_forward", "sfb_rasterize_backward", "sfb_mark_visible",
    "sfb_export_geom", "sfb_export_binning", "sfb_export_img", "sfb_debug_gather_rows", "sfb_last_launch_count",
    "sfb_profile_enable", "sfb_profile_count", "sfb_profile_read", "sfb_profile_name",
    "sfb_loss_scratch_bytes", "sfb_loss_window", "sfb_l1_ssim_loss", "sfb_densify_stats", "sfb_densify_masks",
    "sfb_sh_grad_combine", "sfb_xchg_bytes", "sfb_xchg_finish", "sfb_xchg_status", "sfb_xchg_timeline", "sfb_xchg_tune", "sfb_activate_forward", "sfb_activate_backward", "sfb_knn_scratch_bytes",
    "sfb_knn3_mean_dist2",
)


BWD_ACC_FRESH = 1     # include/splat_b200.h: SFB_BWD_ACC_FRESH
BWD_SH_FACTORED = 2   # include/splat_b200.h: SFB_BWD_SH_FACTORED


XCHG_MAX_RANKS = 16   # include/splat_b200.h: SFB_XCHG_MAX_RANKS


class XchgDesc(C.Structure):
    """include/splat_b200.h: sfb_xchg — the symmetric buffers of the view-parallel gradient exchange."""
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("P", C.c_int), ("ngeo", C.c_int), ("local", C.c_void_p),
                ("peers", C.c_void_p * XCHG_MAX_RANKS), ("mc", C.c_void_p), ("max_ctas", C.c_int),
                ("campos_views", C.c_void_p)]


class SplatB200Error(RuntimeError):
    pass


def load():
    """Load the library; raise loudly when it has not been built (python -m splatfields_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SplatB200Error(
            f"{LIB_PATH} not found: build it with `python -m splatfields_b200.build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the rasterizer.")
    lib = C.CDLL(LIB_PATH)
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    lib.sfb_abi_version.restype = ci
    lib.sfb_last_error.restype = C.c_char_p
    lib.sfb_last_launch_count.restype = ci
    lib.sfb_rasterize_forward.restype = ci
    lib.sfb_rasterize_forward.argtypes = [
        ci, ci, ci, ci, ci,                       # P, sh_degree, M, W, H
        vp, vp, vp, vp, vp, vp, cf, vp, vp,       # bg, means3D, shs, colors, opacities, scales, mod, rot, cov3D
        vp, vp, vp, cf, cf, ci,                   # view, proj, campos, tanfovx, tanfovy, prefiltered
        vp, vp, vp, vp,                           # out_color, out_depth, out_alpha (nullable), radii
        ALLOC_FN, vp, ALLOC_FN, vp, ALLOC_FN, vp,
        C.POINTER(ci), ci, vp]
    lib.sfb_rasterize_backward.restype = ci
    lib.sfb_rasterize_backward.argtypes = [
        ci, ci, ci, ci, ci, ci,                   # P, sh_degree, M, R, W, H
        vp, vp, vp, vp, vp, cf, vp, vp,           # bg, means3D, shs, colors, scales, mod, rot, cov3D
        vp, vp, vp, cf, cf, vp,                   # view, proj, campos, tanfovx, tanfovy, radii
        vp, vp, vp, vp, vp, vp,                   # geom, binning, img, dL_dout_color, dL_dout_alpha, dL_dout_depth (nullable)
        vp, vp, vp, vp, vp, vp, vp, vp,           # 8 gradient outputs
        ci, ci,                                   # debug, flags (SFB_BWD_ACC_FRESH | SFB_BWD_SH_FACTORED)
        C.POINTER(XchgDesc), C.c_uint, vp]        # xchg (nullable), xchg_epoch, stream
    lib.sfb_mark_visible.restype = ci
    lib.sfb_mark_visible.argtypes = [ci, vp, vp, vp, vp, vp]
    lib.sfb_export_geom.restype = ci
    lib.sfb_export_geom.argtypes = [ci, vp, vp, cf, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.sfb_export_binning.restype = ci
    lib.sfb_export_binning.argtypes = [ci, ci, ci, ci, vp, vp, vp, vp, vp, vp]
    lib.sfb_export_img.restype = ci
    lib.sfb_export_img.argtypes = [ci, ci, vp, vp, vp, vp]
    lib.sfb_debug_gather_rows.restype = ci
    lib.sfb_debug_gather_rows.argtypes = [ci, vp, ci, vp, vp, vp]
    lib.sfb_profile_enable.restype = None
    lib.sfb_profile_enable.argtypes = [ci]
    lib.sfb_profile_read.restype = ci
    lib.sfb_profile_read.argtypes = [ci, C.POINTER(cf), ci]
    lib.sfb_profile_count.restype = ci
    lib.sfb_profile_count.argtypes = [ci]
    lib.sfb_profile_name.restype = C.c_char_p
    lib.sfb_profile_name.argtypes = [ci, ci]
    lib.sfb_loss_scratch_bytes.restype = C.c_size_t
    lib.sfb_loss_scratch_bytes.argtypes = [ci, ci, ci]
    lib.sfb_loss_window.restype = None
    lib.sfb_loss_window.argtypes = [C.POINTER(cf)]
    lib.sfb_l1_ssim_loss.restype = ci
    lib.sfb_l1_ssim_loss.argtypes = [ci, ci, ci, vp, vp, cf, vp, vp, cf, cf, vp, vp, vp, vp, vp]
    lib.sfb_densify_stats.restype = ci
    lib.sfb_densify_stats.argtypes = [ci, vp, vp, vp, vp, vp, vp, vp]
    lib.sfb_densify_masks.restype = ci
    lib.sfb_densify_masks.argtypes = [ci, vp, vp, vp, vp, vp, ci, cf, cf, cf, cf, cf, vp, vp, vp, vp, vp]
    lib.sfb_sh_grad_combine.restype = ci
    lib.sfb_sh_grad_combine.argtypes = [ci, ci, ci, ci, vp, vp, vp, vp, vp]
    lib.sfb_xchg_bytes.restype = C.c_size_t
    lib.sfb_xchg_bytes.argtypes = [ci, ci, ci, ci]
    lib.sfb_xchg_finish.restype = ci
    lib.sfb_xchg_finish.argtypes = [C.POINTER(XchgDesc), C.c_uint, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.sfb_xchg_status.restype = ci
    lib.sfb_xchg_status.argtypes = [C.POINTER(XchgDesc), C.POINTER(C.c_uint), vp]
    lib.sfb_xchg_timeline.restype = ci
    lib.sfb_xchg_timeline.argtypes = [C.POINTER(XchgDesc), C.POINTER(C.c_ulonglong), vp]
    lib.sfb_xchg_tune.restype = None
    lib.sfb_xchg_tune.argtypes = [ci, ci]
    lib.sfb_activate_forward.restype = ci
    lib.sfb_activate_forward.argtypes = [ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.sfb_activate_backward.restype = ci
    lib.sfb_activate_backward.argtypes = [ci, ci, ci] + [vp] * 13
    lib.sfb_knn_scratch_bytes.restype = C.c_size_t
    lib.sfb_knn_scratch_bytes.argtypes = [ci]
    lib.sfb_knn3_mean_dist2.restype = ci
    lib.sfb_knn3_mean_dist2.argtypes = [ci, vp, vp, vp, vp]
    if lib.sfb_abi_version() != 6:
        raise SplatB200Error("libsplat_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().sfb_last_error().decode("utf-8", "replace")
        if rc == -2:
            raise Exception(msg)   # argument errors surface like the reference's Python exceptions
        raise SplatB200Error(f"libsplat_b200 error {rc}: {msg}")


def profile_enable(on: bool):
    load().sfb_profile_enable(int(bool(on)))


def profile_read(which: int) -> list:
    """[(kernel name, ms), ...] of the last forward (which=0) / backward (which=1) run with profiling on."""
    lib = load()
    buf = (C.c_float * 64)()
    n = lib.sfb_profile_read(which, buf, 64)
    if n < 0:
        check(n)
    return [(lib.sfb_profile_name(which, i).decode(), float(buf[i])) for i in range(n)]
