"""Shared helpers of the parity tests: run one scene through the CPU oracle and through the CUDA
library (via the Python mirror of the reference interface) and compare."""
from __future__ import annotations

import math

import numpy as np
import torch

from splatfields_b200 import synth


def scene_and_camera(P, H, W, seed, *, sh_degree=3, precomp_rgb=False, scale_mult=1.0, extent=1.3, cam_k=0,
                     cam_name=None, cov_precomp=False):
    sc = synth.make_scene(P, seed, scale_mult=scale_mult, extent=extent, precomp_rgb=precomp_rgb)
    if cam_name is not None:
        cam = synth.config_camera(cam_name, cam_k)
    else:
        cam = synth.orbit_camera(cam_k, H, W)
    if cov_precomp:
        from oracle import torch_naive as TN
        sc["cov3D_precomp"] = TN.cov3d_6(sc["scales"].double(), sc["rotations"].double()).float()
        del sc["scales"], sc["rotations"]
    return sc, cam


def cam_kwargs(cam, H, W, bg):
    return dict(bg=np.asarray(bg, np.float32), viewmatrix=cam.world_view_transform.numpy(),
                projmatrix=cam.full_proj_transform.numpy(), campos=cam.camera_center.numpy(),
                tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5), H=H, W=W)


def run_oracle(O, sc, cam, H, W, bg, sh_degree, dL=None, want_margin=True, dLd=None):
    kw = cam_kwargs(cam, H, W, bg)
    n = lambda k: sc[k].numpy() if k in sc else None
    fwd = O.forward(n("means3D"), n("opacities"), n("scales"), n("rotations"), shs=n("shs"),
                    colors_precomp=n("colors_precomp"), cov3D_precomp=n("cov3D_precomp"), sh_degree=sh_degree,
                    want_margin=want_margin, **kw)
    bwd = None
    if dL is not None:
        bwd = O.backward(fwd, dL, n("means3D"), n("scales"), n("rotations"), shs=n("shs"),
                         cov3D_precomp=n("cov3D_precomp"), viewmatrix=kw["viewmatrix"], projmatrix=kw["projmatrix"],
                         campos=kw["campos"], tanfovx=kw["tanfovx"], tanfovy=kw["tanfovy"], sh_degree=sh_degree,
                         dL_ddepth=dLd)
    return fwd, bwd


def run_cuda(sc, cam, H, W, bg, sh_degree, dL=None, device="cuda", debug=False, dLd=None):
    """Through GaussianRasterizer (the reference-facing surface) -> C ABI.  Returns numpy dicts."""
    import ctypes as C
    from splatfields_b200 import _lib
    from splatfields_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer, _RasterizeGaussians
    dev = torch.device(device)
    t = {k: v.to(dev).clone().requires_grad_(True) for k, v in sc.items()}
    camd = cam.to(dev)
    rs = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5),
        bg=torch.tensor(bg, dtype=torch.float32, device=dev), scale_modifier=1.0,
        viewmatrix=camd.world_view_transform, projmatrix=camd.full_proj_transform, sh_degree=sh_degree,
        campos=camd.camera_center, prefiltered=False, debug=debug)
    means2D = torch.zeros_like(t["means3D"], requires_grad=True)
    rast = GaussianRasterizer(rs)
    color, radii, depth = rast(means3D=t["means3D"], means2D=means2D, opacities=t["opacities"],
                               shs=t.get("shs"), colors_precomp=t.get("colors_precomp"), scales=t.get("scales"),
                               rotations=t.get("rotations"), cov3D_precomp=t.get("cov3D_precomp"))
    out = dict(color=color.detach().cpu().numpy(), depth=depth.detach().cpu().numpy(),
               radii=radii.cpu().numpy())
    # internals through the export entry points (needs the autograd ctx buffers)
    fn = color.grad_fn
    if fn is not None:
        radii_s, geom, binning, img = fn.saved_tensors[:4]
        saved_inputs = fn.saved_tensors[4:]
        lib = _lib.load()
        P = t["means3D"].shape[0]
        R = fn.num_rendered
        T = ((W + 15) // 16) * ((H + 15) // 16)
        e = dict(means2D=torch.zeros(P, 2, device=dev), depths=torch.zeros(P, device=dev),
                 cov3D=torch.zeros(P, 6, device=dev), conic_opacity=torch.zeros(P, 4, device=dev),
                 rgb=torch.zeros(P, 3, device=dev), clamped=torch.zeros(P, 3, dtype=torch.uint8, device=dev),
                 tiles_touched=torch.zeros(P, dtype=torch.int32, device=dev))
        dp = lambda k: saved_inputs[k].data_ptr() if saved_inputs[k] is not None else None
        # saved_inputs = (means3D, sh, col, scales, rotations, cov3D_precomp, bg, view, proj, campos)
        _lib.check(lib.sfb_export_geom(P, geom.data_ptr(), dp(3), 1.0, dp(4), dp(5),
                                       e["means2D"].data_ptr(), e["depths"].data_ptr(),
                                       e["cov3D"].data_ptr(), e["conic_opacity"].data_ptr(), e["rgb"].data_ptr(),
                                       e["clamped"].data_ptr(), e["tiles_touched"].data_ptr(), None))
        keys = torch.zeros(max(R, 1), dtype=torch.int64, device=dev)
        pl = torch.zeros(max(R, 1), dtype=torch.int32, device=dev)
        ranges = torch.zeros(T, 2, dtype=torch.int32, device=dev)
        _lib.check(lib.sfb_export_binning(P, R, W, H, geom.data_ptr(), binning.data_ptr(), keys.data_ptr(),
                                          pl.data_ptr(), ranges.data_ptr(), None))
        fT = torch.zeros(H, W, device=dev)
        nc = torch.zeros(H, W, dtype=torch.int32, device=dev)
        _lib.check(lib.sfb_export_img(W, H, img.data_ptr(), fT.data_ptr(), nc.data_ptr(), None))
        torch.cuda.synchronize()
        out.update({k: v.cpu().numpy() for k, v in e.items()})
        out["tiles_touched"] = out["tiles_touched"].astype(np.uint32)
        out.update(num_rendered=R, point_list_keys=keys[:R].cpu().numpy().astype(np.uint64),
                   point_list=pl[:R].cpu().numpy().astype(np.uint32),
                   ranges=ranges.cpu().numpy().astype(np.uint32), final_T=fT.cpu().numpy(),
                   n_contrib=nc.cpu().numpy().astype(np.uint32))
    grads = None
    if dL is not None:
        loss = (color * torch.as_tensor(dL, device=dev)).sum()
        if dLd is not None:      # a loss on the depth image as well
            loss = loss + (depth * torch.as_tensor(dLd, device=dev).reshape(depth.shape)).sum()
        loss.backward()
        torch.cuda.synchronize()
        grads = dict(dL_dmeans3D=t["means3D"].grad, dL_dmeans2D=means2D.grad, dL_dopacity=t["opacities"].grad)
        for k_in, k_out in (("shs", "dL_dsh"), ("colors_precomp", "dL_dcolors"), ("scales", "dL_dscales"),
                            ("rotations", "dL_drotations"), ("cov3D_precomp", "dL_dcov3D")):
            if k_in in t:
                grads[k_out] = t[k_in].grad
        grads = {k: v.detach().cpu().numpy() for k, v in grads.items()}
    return out, grads


# ---- tolerances (BASELINE.json north_star): image/depth 1e-5 abs, gradients 1e-3 rel ----
IMG_ATOL = 1e-5
GRAD_RTOL = 1e-3


def grad_close(name, got, ref, rtol=GRAD_RTOL, atol_frac=1e-5):
    """|got - ref| <= rtol*|ref| + atol, atol = atol_frac * max|ref|.  The absolute floor covers splats
    whose net gradient is a cancellation of large per-pixel terms (fp32 summation order differs from the
    oracle's fp64 accumulation, exactly as it differs between two runs of the reference's atomics)."""
    got = np.asarray(got, np.float64).reshape(-1)
    ref = np.asarray(ref, np.float64).reshape(-1)
    scale = np.abs(ref).max() if ref.size else 0.0
    err = np.abs(got - ref)
    tol = rtol * np.abs(ref) + atol_frac * scale
    bad = err > tol
    nrm = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)
    assert not bad.any(), (f"{name}: {bad.sum()} / {ref.size} elements off; worst err {err[bad].max():.3e} "
                           f"(ref scale {scale:.3e}), norm-rel {nrm:.3e}")
    assert nrm <= rtol, f"{name}: norm-wise relative error {nrm:.3e} > {rtol}"
    return nrm
