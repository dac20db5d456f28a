"""Host-side mirror of the densification bookkeeping that consumes the rasterizer's side outputs every training
iteration: GaussianModel.add_densification_stats (scene/gaussian_model.py:427-430), the
max_radii2D update (train.py:280-282) and the selection predicates of densify_and_prune / densify_and_clone /
densify_and_split (scene/gaussian_model.py:355-425).  One kernel each (csrc/densify.cu) instead of ~10 boolean-mask
indexing launches with a nonzero() host sync apiece.  The optimizer-state surgery that follows is the caller's."""
from __future__ import annotations

import torch

from . import _lib
from .rasterizer import _ptr


def _flat_f32(t, name):
    if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
        raise Exception(f"{name} must be a contiguous fp32 CUDA tensor (it is updated in place)")
    return t


def add_densification_stats(xyz_gradient_accum, denom, viewspace_grad, update_filter=None, radii=None, max_radii2D=None):
    """In place, for every i with update_filter[i] (default: radii[i] > 0, the reference's visibility_filter):
        xyz_gradient_accum[i] += ||viewspace_grad[i, :2]||;  denom[i] += 1;  max_radii2D[i] = max(max_radii2D[i], radii[i])
    viewspace_grad: viewspace_points.grad [P, 3]; radii: int32 [P] from the rasterizer."""
    lib = _lib.load()
    P = viewspace_grad.shape[0]
    acc, den = _flat_f32(xyz_gradient_accum, "xyz_gradient_accum"), _flat_f32(denom, "denom")
    g = viewspace_grad
    if g.dtype != torch.float32 or not g.is_contiguous():
        g = g.float().contiguous()
    if g.dim() != 2 or g.shape[1] != 3:
        raise Exception("viewspace_grad must have dimensions (num_points, 3)")
    if update_filter is None and radii is None:
        raise Exception("provide update_filter or radii")
    f = None
    if update_filter is not None:
        f = update_filter.contiguous()
        f = f.view(torch.uint8) if f.dtype == torch.bool else f.to(torch.uint8)
    r = None
    if radii is not None:
        r = radii if (radii.dtype == torch.int32 and radii.is_contiguous()) else radii.to(torch.int32).contiguous()
    mr = None if max_radii2D is None else _flat_f32(max_radii2D, "max_radii2D")
    if acc.numel() != P or den.numel() != P or (mr is not None and mr.numel() != P):
        raise Exception("statistics buffers must hold one value per Gaussian")
    with torch.cuda.device(g.device):
        stream = torch.cuda.current_stream(g.device).cuda_stream
        _lib.check(lib.sfb_densify_stats(P, _ptr(g), _ptr(r), _ptr(f), _ptr(acc), _ptr(den), _ptr(mr), stream))


def densify_masks(xyz_gradient_accum, denom, scaling, opacity, max_radii2D, grad_threshold, percent_dense, extent,
                  min_opacity, max_screen_size, raw=False):
    """(clone_mask, split_mask, prune_mask, counts[3]) for the P existing Gaussians — the predicates of
    densify_and_prune(max_grad, min_opacity, extent, max_screen_size).  scaling [P,3] / opacity [P,1]: activated
    values (get_scaling / get_opacity), or the raw parameters with raw=True (exp / sigmoid applied in the kernel)."""
    lib = _lib.load()
    P = scaling.shape[0]
    dev = scaling.device
    if scaling.dim() != 2 or scaling.shape[1] not in (1, 3):
        raise Exception("scaling must have dimensions (num_points, 3), or (num_points, 1) for an isotropic model")
    if scaling.shape[1] == 1:        # the reference's use_isotropic model keeps one raw scale per Gaussian
        scaling = scaling.expand(P, 3)   # (scene/gaussian_model.py:64-68 repeats it the same way)
    sc = scaling.detach().float().contiguous()
    op = opacity.detach().float().contiguous().reshape(-1)
    acc, den = xyz_gradient_accum.detach().float().contiguous(), denom.detach().float().contiguous()
    mr = None if max_radii2D is None else max_radii2D.detach().float().contiguous()
    masks = torch.empty((3, P), dtype=torch.uint8, device=dev)
    counts = torch.empty(3, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.sfb_densify_masks(
            P, _ptr(acc), _ptr(den), _ptr(sc), _ptr(op), _ptr(mr), int(bool(raw)), float(grad_threshold),
            float(percent_dense * extent), float(0.1 * extent), float(min_opacity), float(max_screen_size or 0.0),
            masks[0].data_ptr(), masks[1].data_ptr(), masks[2].data_ptr(), counts.data_ptr(), stream))
    m = masks.view(torch.bool)
    return m[0], m[1], m[2], counts
