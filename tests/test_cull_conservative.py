"""CPU: the patch-culling predicate of the render kernels (csrc/common.cuh: patch_mask + refine_patch_mask) is
CONSERVATIVE — it never drops a (splat, 8x4 patch) pair in which some pixel passes the reference's own
`power <= 0 and alpha >= 1/255` test.  The device formulas are restated here in float32 numpy and compared with a
brute-force evaluation of all 32 pixels in the reference's op order, over random conics (condition number up to
1e4 — beyond that the kernels never cull), opacities and patch offsets.  (The GPU parity tests then check that
the composited images are unchanged.)"""
import numpy as np
import pytest

f32 = np.float32


def _cases(seed, N, smax, off):
    rng = np.random.default_rng(seed)
    s1 = np.exp(rng.uniform(np.log(0.3), np.log(smax), N))
    s2 = np.exp(rng.uniform(np.log(0.3), np.log(smax), N))
    th = rng.uniform(0, np.pi, N)
    c, s = np.cos(th), np.sin(th)
    a = (c * c * s1 * s1 + s * s * s2 * s2 + 0.3).astype(f32)       # dilated 2D covariance
    b = (c * s * (s1 * s1 - s2 * s2)).astype(f32)
    cc = (s * s * s1 * s1 + c * c * s2 * s2 + 0.3).astype(f32)
    det = a * cc - b * b
    A, B, C = (cc / det).astype(f32), (-b / det).astype(f32), (a / det).astype(f32)
    op = np.maximum((1 / (1 + np.exp(-rng.normal(0, 1.5, N)))).astype(f32), f32(1 / 255.0))
    ill = (np.maximum(s1, s2) ** 2 + 0.3) / (np.minimum(s1, s2) ** 2 + 0.3) > 1e4
    X0 = rng.uniform(-off, off - 10, N).astype(f32)
    Y0 = rng.uniform(-off, off - 10, N).astype(f32)
    return a, cc, A, B, C, op, ill, X0, Y0


@pytest.mark.parametrize("seed,smax,off", [(0, 60, 60), (1, 400, 300), (2, 8, 20)])
def test_patch_culling_never_drops_a_contributing_pair(seed, smax, off):
    N = 200_000
    a, cc, A, B, C, op, ill, X0, Y0 = _cases(seed, N, smax, off)
    X1, Y1 = X0 + f32(7), Y0 + f32(3)
    tau = (2 * np.log(255.0 * op)).astype(f32)
    # footprint box (preprocess.cu): 2 % + 0.5 px inflated AABB of {q <= tau}
    hx = f32(1.02) * np.sqrt(np.maximum(tau, 0) * a) + f32(0.5)
    hy = f32(1.02) * np.sqrt(np.maximum(tau, 0) * cc) + f32(0.5)
    box = (hx >= X0) & (-hx <= X1) & (hy >= Y0) & (-hy <= Y1)
    # exact refinement (common.cuh: refine_patch_mask)
    inside = (X0 <= 0) & (X1 >= 0) & (Y0 <= 0) & (Y1 >= 0)
    nBA, nBC = (-B / A).astype(f32), (-B / C).astype(f32)
    q = lambda dx, dy: (A * dx * dx + f32(2) * B * dx * dy + C * dy * dy).astype(f32)
    d0, d1 = np.clip(nBC * X0, Y0, Y1), np.clip(nBC * X1, Y0, Y1)
    e0, e1 = np.clip(nBA * Y0, X0, X1), np.clip(nBA * Y1, X0, X1)
    qmin = np.minimum(np.minimum(q(X0, d0), q(X1, d1)), np.minimum(q(e0, Y0), q(e1, Y1)))
    DX, DY = np.maximum(np.abs(X0), np.abs(X1)), np.maximum(np.abs(Y0), np.abs(Y1))
    M = A * DX * DX + C * DY * DY + f32(2) * np.abs(B) * DX * DY
    refined = inside | ~(qmin > tau * f32(1.0001) + f32(2e-3) + f32(8e-6) * M)
    keep = ill | (box & refined)
    # brute force over the 32 pixel centres, reference op order
    contrib = np.zeros(N, bool)
    for iy in range(4):
        for ix in range(8):
            dx, dy = -(X0 + f32(ix)), -(Y0 + f32(iy))
            power = (((A * dx) * dx + (C * dy) * dy) * f32(-0.5) - (B * dx) * dy).astype(f32)
            with np.errstate(over="ignore"):
                alpha = np.minimum(f32(0.99), op * np.exp(power).astype(f32))
            contrib |= (power <= 0) & (alpha >= f32(1 / 255.0))
    assert contrib.sum() > 1000
    assert not (contrib & ~keep).any(), "culling dropped a contributing (splat, patch) pair"
    # and it is tight: of the pairs it keeps (well-conditioned splats), at most a few percent never contribute
    kept_ok = keep & ~ill
    assert (kept_ok & ~contrib).sum() <= 0.05 * max(kept_ok.sum(), 1)


@pytest.mark.parametrize("seed,smax", [(3, 40), (4, 150), (5, 6)])
def test_tight_tile_rect_never_drops_a_contributing_tile(seed, smax):
    """The footprint box preprocess stores per splat (SplatRec::hx, hy; the render kernels cull 8x4 patches with it) is
    conservative at tile granularity too: clipping the reference's tile rectangle to it, restated in float32 numpy
    (floorf, +1 on the exclusive edge, clamps to the grid), never drops a tile in which some pixel passes the
    reference's alpha >= 1/255 test (brute force).  (Round 1 shipped that clip as an opt-in preprocess variant; it
    changed the key / index buffers and was removed — the property it rests on is what the culling relies on.)"""
    N, gx, gy = 4000, 50, 50                     # 800x800 image
    rng = np.random.default_rng(seed)
    a, cc, A, B, C, op, ill, _, _ = _cases(seed, N, smax, 60)
    op = np.where(rng.random(N) < 0.1, f32(0.002), op)            # some splats below 1/255 everywhere
    pix = rng.uniform(-40, 840, N).astype(f32)
    piy = rng.uniform(-40, 840, N).astype(f32)
    mid = f32(0.5) * (a + cc)
    det = a * cc - (B / (A * C - B * B)) ** 2                      # b^2 recovered from the conic (b = -B * det)
    lam1 = mid + np.sqrt(np.maximum(f32(0.1), mid * mid - det)).astype(f32)
    rad = np.ceil(f32(3.0) * np.sqrt(lam1)).astype(f32)
    clampi = lambda v, hi: np.clip(v, 0, hi)
    x0 = clampi(((pix - rad) / f32(16)).astype(np.int32), gx); y0 = clampi(((piy - rad) / f32(16)).astype(np.int32), gy)
    x1 = clampi(((pix + rad + f32(15)) / f32(16)).astype(np.int32), gx)
    y1 = clampi(((piy + rad + f32(15)) / f32(16)).astype(np.int32), gy)
    tau = np.maximum(f32(0), (f32(2) * np.log(f32(255) * op)).astype(f32))
    hx = np.where(ill, f32(3e38), f32(1.02) * np.sqrt(tau * a) + f32(0.5)).astype(f32)
    hy = np.where(ill, f32(3e38), f32(1.02) * np.sqrt(tau * cc) + f32(0.5)).astype(f32)
    below = op < f32(1 / 255.0)
    with np.errstate(over="ignore", invalid="ignore"):
        tx0 = np.maximum(x0, clampi(np.floor((pix - hx) / f32(16)).astype(np.int64), gx))
        ty0 = np.maximum(y0, clampi(np.floor((piy - hy) / f32(16)).astype(np.int64), gy))
        tx1 = np.minimum(x1, clampi(np.floor((pix + hx) / f32(16)).astype(np.int64) + 1, gx))
        ty1 = np.minimum(y1, clampi(np.floor((piy + hy) / f32(16)).astype(np.int64) + 1, gy))
    big = hx >= f32(1e30)
    tx0, ty0, tx1, ty1 = (np.where(big, o, t) for o, t in ((x0, tx0), (y0, ty0), (x1, tx1), (y1, ty1)))
    tx1 = np.where(below, tx0, tx1)                                 # area 0
    ox, oy = np.meshgrid(np.arange(16, dtype=np.float32), np.arange(16, dtype=np.float32))
    dropped = kept = contributing_dropped = 0
    for i in range(N):
        for ty in range(y0[i], y1[i]):
            for tx in range(x0[i], x1[i]):
                if tx0[i] <= tx < tx1[i] and ty0[i] <= ty < ty1[i]:
                    kept += 1
                    continue
                dropped += 1
                dx = pix[i] - (f32(16 * tx) + ox)
                dy = piy[i] - (f32(16 * ty) + oy)
                power = ((A[i] * dx) * dx + (C[i] * dy) * dy) * f32(-0.5) - (B[i] * dx) * dy
                with np.errstate(over="ignore"):
                    alpha = np.minimum(f32(0.99), op[i] * np.exp(power.astype(f32)))
                if ((power <= 0) & (alpha >= f32(1 / 255.0))).any():
                    contributing_dropped += 1
    assert dropped > 0.05 * (dropped + kept), "the clip should remove a visible share of the tiles"
    assert contributing_dropped == 0
