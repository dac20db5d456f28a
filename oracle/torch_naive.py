"""Dense fp64 PyTorch restatement of the splat rasterizer — an INDEPENDENT cross-check of the C
oracle's analytic backward (autograd through this forward), and the "pure-PyTorch projection +
naive blend" CPU leg that BASELINE.md §2.2 names.  O(P·H·W): small scenes only.

TEST INFRASTRUCTURE ONLY (same rule as splat_oracle.c): never imported by splatfields_b200/.

Follows the reference's in-tree Python math:
  * row-vector homogeneous transform, divide by (w + 1e-7): utils/graphics_utils.py:24-31
  * quaternion -> R, L = R S, Sigma = L L^T:               utils/general_utils.py:138-171
  * SH basis, constants and signs:                          utils/sh_utils.py:26-112
  * tanfov, settings packing:                               gaussian_renderer/__init__.py:56-72
and SURVEY.md Appendix A for what has no in-tree source (EWA, radius, rect, compositing rules).
"""
from __future__ import annotations

import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]


def sh_basis(deg: int, d: torch.Tensor) -> torch.Tensor:
    """[N, (deg+1)^2] real SH basis values at unit directions d[N,3]."""
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    b = [torch.full_like(x, C0)]
    if deg > 0:
        b += [-C1 * y, C1 * z, -C1 * x]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        b += [C2[0] * xy, C2[1] * yz, C2[2] * (2 * zz - xx - yy), C2[3] * xz, C2[4] * (xx - yy)]
    if deg > 2:
        b += [C3[0] * y * (3 * xx - yy), C3[1] * xy * z, C3[2] * y * (4 * zz - xx - yy),
              C3[3] * z * (2 * zz - 3 * xx - 3 * yy), C3[4] * x * (4 * zz - xx - yy), C3[5] * z * (xx - yy),
              C3[6] * x * (xx - 3 * yy)]
    return torch.stack(b, dim=1)


def sh_grad_combine_ref(means3D, campos_views, dcolor_views, deg: int, M: int) -> torch.Tensor:
    """fp64 restatement of sfb_sh_grad_combine: the multi-view sum of SH gradients from their factored form,
        dL_dsh[i, k, :] = sum_v basis_k(normalize(means3D[i] - campos_views[v])) * dcolor_views[v, i, :],
    i.e. what loss.backward() over the serial per-view loop of train.py:169-242 accumulates in features.grad when
    the colour of view v is eval_sh(deg, shs, dir_v) (utils/sh_utils.py:57-112; extract_geo.py:40-44) and
    dcolor_views[v] is its gradient before the max(0, .) clamp.  Returns [P, M, 3] (zeros above the active degree)."""
    means3D = torch.as_tensor(means3D, dtype=torch.float64)
    campos_views = torch.as_tensor(campos_views, dtype=torch.float64).reshape(-1, 3)
    dcolor_views = torch.as_tensor(dcolor_views, dtype=torch.float64).reshape(campos_views.shape[0], -1, 3)
    P = means3D.shape[0]
    out = torch.zeros(P, M, 3, dtype=torch.float64)
    nb = (deg + 1) ** 2
    for v in range(campos_views.shape[0]):
        d = means3D - campos_views[v]
        d = d / d.norm(dim=1, keepdim=True)
        out[:, :nb, :] += sh_basis(deg, d)[:, :, None] * dcolor_views[v][:, None, :]
    return out


def quat_to_rot(q: torch.Tensor) -> torch.Tensor:
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1)
    return R.reshape(-1, 3, 3)


def cov3d_6(scales, rotations, mod=1.0):
    L = quat_to_rot(rotations) * (mod * scales)[:, None, :]
    S = L @ L.transpose(1, 2)
    return torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], dim=1)


def render_dense(means3D, opacities, scales=None, rotations=None, shs=None, colors_precomp=None,
                 cov3D_precomp=None, *, bg, viewmatrix, projmatrix, campos, tanfovx, tanfovy, H, W,
                 sh_degree=0, scale_modifier=1.0, dtype=torch.float64):
    """Returns (color[3,H,W], depth[1,H,W], radii[P]).  All inputs torch CPU tensors; differentiable."""
    cast = lambda t: None if t is None else t.to(dtype)
    means3D, opacities = cast(means3D), cast(opacities).reshape(-1)
    scales, rotations, shs = cast(scales), cast(rotations), cast(shs)
    colors_precomp, cov3D_precomp = cast(colors_precomp), cast(cov3D_precomp)
    V, Pm, cam, bg = cast(viewmatrix), cast(projmatrix), cast(campos), cast(bg)
    P = means3D.shape[0]
    ones = torch.ones(P, 1, dtype=dtype)
    hom = torch.cat([means3D, ones], dim=1)
    p_view = hom @ V            # row-vector convention: tensors hold the transposed matrices
    p_hom = hom @ Pm
    p_w = 1.0 / (p_hom[:, 3] + 0.0000001)
    ndc = p_hom[:, :2] * p_w[:, None]
    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)

    cov6 = cov3D_precomp if cov3D_precomp is not None else cov3d_6(scales, rotations, scale_modifier)
    Sig = torch.stack([cov6[:, 0], cov6[:, 1], cov6[:, 2], cov6[:, 1], cov6[:, 3], cov6[:, 4], cov6[:, 2],
                       cov6[:, 4], cov6[:, 5]], dim=1).reshape(P, 3, 3)
    tz = p_view[:, 2]
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    txtz, tytz = p_view[:, 0] / tz, p_view[:, 1] / tz
    cx, cy = (txtz < -limx) | (txtz > limx), (tytz < -limy) | (tytz > limy)
    # the external backward treats a clamped t.x / t.y as a constant
    tx = torch.where(cx, (txtz.clamp(-limx, limx) * tz).detach(), p_view[:, 0])
    ty = torch.where(cy, (tytz.clamp(-limy, limy) * tz).detach(), p_view[:, 1])
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -(fx * tx) / (tz * tz), zero, fy / tz, -(fy * ty) / (tz * tz)], dim=1).reshape(P, 2, 3)
    Rw = V[:3, :3].t()          # mathematical world->camera rotation
    Mm = J @ Rw
    cov2 = Mm @ Sig @ Mm.transpose(1, 2)
    a, b, c = cov2[:, 0, 0] + 0.3, cov2[:, 0, 1], cov2[:, 1, 1] + 0.3
    det = a * c - b * b
    conA, conB, conC = c / det, -b / det, a / det
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam)).detach()
    px = ((ndc[:, 0] + 1.0) * W - 1.0) * 0.5
    py = ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5
    gx, gy = (W + 15) // 16, (H + 15) // 16

    def rect(p, lo_hi, g):
        v = (p.detach() + lo_hi) / 16.0
        return torch.clamp(torch.trunc(v), 0, g).to(torch.int64)
    x0, x1 = rect(px, -radius, gx), rect(px, radius + 15, gx)
    y0, y1 = rect(py, -radius, gy), rect(py, radius + 15, gy)
    visible = (tz.detach() > 0.2) & (det.detach() != 0) & ((x1 - x0) * (y1 - y0) > 0)
    radii = torch.where(visible, radius, torch.zeros_like(radius)).to(torch.int32)

    if colors_precomp is not None:
        rgb = colors_precomp
    else:
        d = means3D - cam[None]
        d = d / d.norm(dim=1, keepdim=True)
        nb = (sh_degree + 1) ** 2
        rgb = torch.einsum("nk,nkc->nc", sh_basis(sh_degree, d), shs[:, :nb, :]) + 0.5
        rgb = torch.clamp(rgb, min=0.0)  # clamp: zero gradient where negative, as `clamped` does

    # global depth order on the fp32 depth (ties by index), as the tile keys use float32 depth bits
    idx = torch.nonzero(visible).reshape(-1)
    order = torch.argsort(p_view[idx, 2].detach().to(torch.float32), stable=True)
    idx = idx[order]
    N = idx.numel()
    ys, xs = torch.meshgrid(torch.arange(H, dtype=dtype), torch.arange(W, dtype=dtype), indexing="ij")
    pxs, pys = xs.reshape(-1), ys.reshape(-1)
    tix, tiy = (pxs // 16).to(torch.int64), (pys // 16).to(torch.int64)
    dx = px[idx][:, None] - pxs[None, :]
    dy = py[idx][:, None] - pys[None, :]
    power = -0.5 * (conA[idx][:, None] * dx * dx + conC[idx][:, None] * dy * dy) - conB[idx][:, None] * dx * dy
    alpha = torch.clamp(opacities[idx][:, None] * torch.exp(power), max=0.99)
    in_rect = (tix[None, :] >= x0[idx][:, None]) & (tix[None, :] < x1[idx][:, None]) & \
              (tiy[None, :] >= y0[idx][:, None]) & (tiy[None, :] < y1[idx][:, None])
    live = in_rect & (power.detach() <= 0) & (alpha.detach() >= 1.0 / 255.0)
    a_eff = torch.where(live, alpha, torch.zeros_like(alpha))
    one_m = 1.0 - a_eff
    T_after = torch.cumprod(one_m, dim=0)
    T_before = torch.cat([torch.ones(1, H * W, dtype=dtype), T_after[:-1]], dim=0) if N > 0 else T_after
    # stop at the first live entry whose T' would drop below 1e-4; it and everything after is dropped
    stop = live & (T_after.detach() < 0.0001)
    dead = torch.cumsum(stop.to(torch.int64), dim=0) > 0
    w = torch.where(dead, torch.zeros_like(a_eff), a_eff * T_before)
    color = torch.einsum("np,nc->cp", w, rgb[idx])
    depth = torch.einsum("np,n->p", w, p_view[idx, 2])
    T_final = torch.prod(torch.where(dead, torch.ones_like(one_m), one_m), dim=0) if N > 0 else torch.ones(H * W, dtype=dtype)
    color = color + T_final[None, :] * bg[:, None]
    return color.reshape(3, H, W), depth.reshape(1, H, W), radii, dict(ndc=ndc, px=px, py=py)
