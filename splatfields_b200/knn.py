"""Host-side mirror of simple_knn._C.distCUDA2 (the reference imports it at
scene/gaussian_model.py:23-25 and calls it once, at :105, to initialise the Gaussian scales from the point cloud):

    distCUDA2(points [P,3] float cuda) -> [P] float: mean squared distance to the 3 nearest other points

Same name, argument and result; the compute is csrc/knn.cu (Morton order + implicit 32-ary box hierarchy, exact)."""
from __future__ import annotations

import torch

from . import _lib


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    if not points.is_cuda:
        raise _lib.SplatB200Error("distCUDA2 runs on CUDA tensors only (no CPU fallback)")
    if points.dim() != 2 or points.shape[1] != 3:
        raise Exception("points must have dimensions (num_points, 3)")
    pts = points.detach()
    if pts.dtype != torch.float32 or not pts.is_contiguous():
        pts = pts.float().contiguous()
    P = pts.shape[0]
    out = torch.empty(P, dtype=torch.float32, device=pts.device)
    if P == 0:
        return out
    scratch = torch.empty(int(lib.sfb_knn_scratch_bytes(P)) + 256, dtype=torch.uint8, device=pts.device)
    base = (scratch.data_ptr() + 255) // 256 * 256
    with torch.cuda.device(pts.device):
        stream = torch.cuda.current_stream(pts.device).cuda_stream
        _lib.check(lib.sfb_knn3_mean_dist2(P, pts.data_ptr(), out.data_ptr(), base, stream))
    return out
