"""ctypes front-end of the CPU oracle (oracle/splat_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of splat_oracle.c.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; nothing under
splatfields_b200/ does.  PARITY UNPINNED for the rasterizer arithmetic itself (the reference's
rasterizer is an un-vendored third-party CUDA extension with no tests; README.md:28 of the
reference pins it), pinned for camera / SH / covariance conventions via tests/golden/.

All arrays are numpy, C-contiguous, float32 unless stated.  The argument names follow the
reference call site gaussian_renderer/__init__.py:94-102 and SURVEY.md §8b.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsplat_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc + OpenMP)."""
    src = os.path.join(_HERE, "splat_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.so_count_rendered.restype = C.c_int64
        _lib.so_num_threads.restype = C.c_int
    return _lib


def num_threads() -> int:
    return int(lib().so_num_threads())


def set_num_threads(n: int) -> None:
    lib().so_set_num_threads(C.c_int(int(n)))


def knn3_mean_dist2(points):
    """Brute-force restatement of simple_knn.distCUDA2 (see so_knn3_mean_dist2)."""
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    out = np.zeros(pts.shape[0], np.float32)
    lib().so_knn3_mean_dist2(C.c_int(pts.shape[0]), _p(pts), _p(out))
    return out


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def preprocess(means3D, opacities, scales, rotations, shs, colors_precomp, cov3D_precomp, viewmatrix,
               projmatrix, campos, tanfovx, tanfovy, H, W, sh_degree, scale_modifier=1.0):
    means3D = _f32(means3D).reshape(-1, 3)
    P = means3D.shape[0]
    opacities = _f32(opacities).reshape(-1)
    scales, rotations = _f32(scales), _f32(rotations)
    shs, colors_precomp, cov3D_precomp = _f32(shs), _f32(colors_precomp), _f32(cov3D_precomp)
    if (shs is None) == (colors_precomp is None):
        raise Exception("Please provide excatly one of either SHs or precomputed colors!")
    if ((scales is None or rotations is None) and cov3D_precomp is None) or (
            (scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
    M = 0 if shs is None else shs.reshape(P, -1, 3).shape[1]
    vm, pm, cp = _f32(viewmatrix).reshape(16), _f32(projmatrix).reshape(16), _f32(campos).reshape(3)
    g = dict(
        radii=np.zeros(P, np.int32), means2D=np.zeros((P, 2), np.float32), depths=np.zeros(P, np.float32),
        cov3D=np.zeros((P, 6), np.float32), rgb=np.zeros((P, 3), np.float32),
        conic_opacity=np.zeros((P, 4), np.float32), tiles_touched=np.zeros(P, np.uint32),
        clamped=np.zeros((P, 3), np.uint8))
    lib().so_preprocess(
        C.c_int(P), C.c_int(sh_degree), C.c_int(M), _p(means3D), _p(scales), C.c_float(scale_modifier),
        _p(rotations), _p(opacities), _p(shs), _p(colors_precomp), _p(cov3D_precomp), _p(vm), _p(pm), _p(cp),
        C.c_int(W), C.c_int(H), C.c_float(tanfovx), C.c_float(tanfovy), _p(g["radii"]), _p(g["means2D"]),
        _p(g["depths"]), _p(g["cov3D"]), _p(g["rgb"]), _p(g["conic_opacity"]), _p(g["tiles_touched"]),
        _p(g["clamped"]))
    return g


def mark_visible(means3D, viewmatrix):
    means3D = _f32(means3D).reshape(-1, 3)
    out = np.zeros(means3D.shape[0], np.uint8)
    lib().so_mark_visible(C.c_int(means3D.shape[0]), _p(means3D), _p(_f32(viewmatrix).reshape(16)), _p(out))
    return out.astype(bool)


def bin_tiles(g, H, W):
    P = g["radii"].shape[0]
    R = int(lib().so_count_rendered(C.c_int(P), _p(g["tiles_touched"])))
    T = ((W + 15) // 16) * ((H + 15) // 16)
    keys = np.zeros(max(R, 1), np.uint64)
    vals = np.zeros(max(R, 1), np.uint32)
    ranges = np.zeros((T, 2), np.uint32)
    lib().so_bin(C.c_int(P), C.c_int(W), C.c_int(H), _p(g["means2D"]), _p(g["depths"]), _p(g["radii"]),
                 _p(g["tiles_touched"]), C.c_int64(R), _p(keys), _p(vals), _p(ranges))
    return dict(num_rendered=R, point_list_keys=keys[:R], point_list=vals[:R], ranges=ranges)


def forward(means3D, opacities, scales=None, rotations=None, shs=None, colors_precomp=None,
            cov3D_precomp=None, *, bg, viewmatrix, projmatrix, campos, tanfovx, tanfovy, H, W, sh_degree=0,
            scale_modifier=1.0, want_margin=False):
    """Full forward: returns dict(color[3,H,W], depth[1,H,W], radii[P] + every intermediate buffer)."""
    g = preprocess(means3D, opacities, scales, rotations, shs, colors_precomp, cov3D_precomp, viewmatrix,
                   projmatrix, campos, tanfovx, tanfovy, H, W, sh_degree, scale_modifier)
    b = bin_tiles(g, H, W)
    bg = _f32(bg).reshape(3)
    color = np.zeros((3, H, W), np.float32)
    depth = np.zeros((1, H, W), np.float32)
    final_T = np.zeros((H, W), np.float32)
    n_contrib = np.zeros((H, W), np.uint32)
    margin = np.zeros((H, W), np.float32) if want_margin else None
    pl = b["point_list"] if b["num_rendered"] else np.zeros(1, np.uint32)
    lib().so_render_forward(C.c_int(W), C.c_int(H), _p(b["ranges"]), _p(pl), _p(g["means2D"]), _p(g["rgb"]),
                            _p(g["depths"]), _p(g["conic_opacity"]), _p(bg), _p(color), _p(depth), _p(final_T),
                            _p(n_contrib), _p(margin))
    out = dict(color=color, depth=depth, final_T=final_T, n_contrib=n_contrib, margin=margin, bg=bg, H=H, W=W)
    out.update(g)
    out.update(b)
    return out


def backward(fwd, dL_dcolor, means3D, scales=None, rotations=None, shs=None, cov3D_precomp=None, *,
             viewmatrix, projmatrix, campos, tanfovx, tanfovy, sh_degree=0, scale_modifier=1.0, dL_ddepth=None):
    """Backward for the cotangent dL_dcolor[3,H,W]; returns the gradients in the layout of the
    reference's rasterize_gaussians_backward: dL_dmeans2D[P,3], dL_dcolors[P,3],
    dL_dopacity[P,1], dL_dmeans3D[P,3], dL_dcov3D[P,6], dL_dsh[P,M,3], dL_dscales[P,3], dL_drot[P,4]."""
    H, W = fwd["H"], fwd["W"]
    means3D = _f32(means3D).reshape(-1, 3)
    P = means3D.shape[0]
    scales, rotations, shs, cov3D_precomp = _f32(scales), _f32(rotations), _f32(shs), _f32(cov3D_precomp)
    M = 0 if shs is None else shs.reshape(P, -1, 3).shape[1]
    dL = _f32(dL_dcolor).reshape(3, H, W)
    d_mean2D = np.zeros((P, 2), np.float64)
    d_conic = np.zeros((P, 3), np.float64)
    d_opac = np.zeros(P, np.float64)
    d_col = np.zeros((P, 3), np.float64)
    pl = fwd["point_list"] if fwd["num_rendered"] else np.zeros(1, np.uint32)
    # optional cotangent of the depth image [1,H,W] (the depth is one more composited channel without background)
    dLd = None if dL_ddepth is None else _f32(dL_ddepth).reshape(H, W)
    d_depth = None if dLd is None else np.zeros(P, np.float64)
    lib().so_render_backward(C.c_int(W), C.c_int(H), _p(fwd["ranges"]), _p(pl), _p(fwd["means2D"]),
                             _p(fwd["rgb"]), _p(fwd["conic_opacity"]), _p(fwd["bg"]), _p(fwd["final_T"]),
                             _p(fwd["n_contrib"]), _p(dL), _p(d_mean2D), _p(d_conic), _p(d_opac), _p(d_col),
                             _p(fwd["depths"]) if dLd is not None else None, _p(dLd), _p(d_depth))
    dz = None if d_depth is None else d_depth.astype(np.float32)
    m2 = d_mean2D.astype(np.float32)
    cn = d_conic.astype(np.float32)
    cl = d_col.astype(np.float32)
    d_means3D = np.zeros((P, 3), np.float32)
    d_cov3D = np.zeros((P, 6), np.float32)
    d_sh = np.zeros((P, M, 3), np.float32)
    d_scales = np.zeros((P, 3), np.float32)
    d_rot = np.zeros((P, 4), np.float32)
    vm, pm, cp = _f32(viewmatrix).reshape(16), _f32(projmatrix).reshape(16), _f32(campos).reshape(3)
    lib().so_preprocess_backward(
        C.c_int(P), C.c_int(sh_degree), C.c_int(M), _p(means3D), _p(fwd["radii"]), _p(shs), _p(fwd["clamped"]),
        _p(scales), _p(rotations), C.c_float(scale_modifier), _p(fwd["cov3D"]),
        C.c_int(0 if cov3D_precomp is None else 1), _p(vm), _p(pm), _p(cp), C.c_int(W), C.c_int(H),
        C.c_float(tanfovx), C.c_float(tanfovy), _p(m2), _p(cn), _p(cl), _p(d_means3D), _p(d_cov3D), _p(d_sh),
        _p(d_scales), _p(d_rot), _p(dz))
    d_means2D = np.zeros((P, 3), np.float32)
    d_means2D[:, :2] = m2
    return dict(dL_dmeans2D=d_means2D, dL_dcolors=cl, dL_dopacity=d_opac.astype(np.float32).reshape(P, 1),
                dL_dmeans3D=d_means3D, dL_dcov3D=d_cov3D, dL_dsh=d_sh, dL_dscales=d_scales, dL_drotations=d_rot,
                dL_dconic=cn)
