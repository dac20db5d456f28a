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
