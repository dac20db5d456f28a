"""CPU, world_size = 2 over gloo: the host logic of the view-parallel path — gradient slab carving,
1/world folded into the cotangent, ONE all-reduce(SUM) — reproduces the reference's serial multi-view
accumulation (train.py:169 loop, train.py:242 mean of the per-view losses).  The rasterizer itself is
replaced by a tiny differentiable stand-in (no GPU here); the CUDA path is covered by -m gpu tests."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from splatfields_b200 import synth
from splatfields_b200.host_api import SLAB_FIELDS_SH, ViewParallelRasterizer


class _StubRast(torch.nn.Module):
    """colour[3,H,W] as a smooth function of every parameter and of the camera (so views differ)."""

    def __init__(self, cam, H, W):
        super().__init__()
        self.cam, self.H, self.W = cam, H, W

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        hom = torch.cat([means3D, torch.ones_like(means3D[:, :1])], 1) @ self.cam.full_proj_transform
        w = torch.sigmoid(hom[:, :1]) * opacities * scales.sum(1, keepdim=True) * (rotations ** 2).sum(1, keepdim=True)
        feat = (shs * torch.linspace(0.5, 1.5, shs.shape[1])[None, :, None]).sum(1)        # [P,3]
        img = (w * feat).t() @ torch.sin(torch.arange(means3D.shape[0] * self.H * self.W, dtype=torch.float32)
                                         .reshape(means3D.shape[0], self.H * self.W) * 0.01)
        color = img.reshape(3, self.H, self.W) + 0.0 * means2D.sum()
        return color, torch.ones(means3D.shape[0], dtype=torch.int32), color[:1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P, H, W = 40, 8, 12
        sc = synth.make_scene(P, 5)
        cam = synth.orbit_camera(rank, H, W)
        vp = ViewParallelRasterizer(sc, cam, H, W, 3, device="cpu", world_size=world, exchange="allreduce")
        vp.rast = _StubRast(cam, H, W)
        G = torch.randn(3, H, W, generator=torch.Generator().manual_seed(100 + rank))
        vp.step(G)
        got = {k: v.clone() for k, v in vp.grads().items()}
        if rank == 0:
            # serial reference: loop over the views, mean of the losses, one backward
            leaves = {k: v.clone().requires_grad_(True) for k, v in sc.items()}
            losses = []
            for r in range(world):
                c = synth.orbit_camera(r, H, W)
                Gr = torch.randn(3, H, W, generator=torch.Generator().manual_seed(100 + r))
                col, _, _ = _StubRast(c, H, W)(leaves["means3D"], torch.zeros(P, 3), leaves["opacities"],
                                               shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
                losses.append((col * Gr).sum())
            (sum(losses) / world).backward()
            ok = all(torch.allclose(got[n], leaves[n].grad.reshape(-1), rtol=1e-4, atol=1e-5) for n, _ in SLAB_FIELDS_SH)
            ret.put(bool(ok))
    finally:
        dist.destroy_process_group()


class _StubRastSH(torch.nn.Module):
    """Stand-in whose colour is the SH polynomial of the view direction (no clamp), like the rasterizer's; in
    factored mode it behaves like the CUDA backward with SFB_BWD_SH_FACTORED: the colour gradient goes to the
    buffer set with set_grad_arena(..., sh_color_out) and `shs` receives no gradient from autograd."""

    def __init__(self, cam, H, W, deg):
        super().__init__()
        self.cam, self.H, self.W, self.deg = cam, H, W, deg

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        from oracle import torch_naive as TN
        from splatfields_b200 import rasterizer
        P = means3D.shape[0]
        d = means3D.detach() - self.cam.camera_center
        basis = TN.sh_basis(self.deg, d / d.norm(dim=1, keepdim=True))             # [P, nb]
        nb = basis.shape[1]
        if rasterizer._SH_COLOR_OUT is not None or getattr(self, "factored", False):
            col = (basis[:, :, None] * shs.detach()[:, :nb]).sum(1).requires_grad_(True)

            def hook(g):
                rasterizer._SH_COLOR_OUT.copy_(g)
            col.register_hook(hook)
        else:
            col = (basis[:, :, None] * shs[:, :nb]).sum(1)
        w = opacities * scales.sum(1, keepdim=True) * (rotations ** 2).sum(1, keepdim=True) * (1 + means3D[:, :1])
        pix = torch.sin(torch.arange(P * self.H * self.W, dtype=torch.float32).reshape(P, self.H * self.W) * 0.01)
        color = ((w * col).t() @ pix).reshape(3, self.H, self.W) + 0.0 * means2D.sum()
        return color, torch.ones(P, dtype=torch.int32), color[:1]


def _worker_factored(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import torch_naive as TN
        P, H, W, deg = 40, 8, 12, 2
        sc = synth.make_scene(P, 5)
        cam = synth.orbit_camera(rank, H, W)
        vp = ViewParallelRasterizer(sc, cam, H, W, deg, device="cpu", world_size=world, exchange="factored")
        assert vp.exchange == "factored" and vp.campos_views.shape == (world, 3)
        vp.rast = _StubRastSH(cam, H, W, deg)
        vp.rast.factored = True
        # CPU stand-in of sfb_sh_grad_combine (the CUDA kernel is covered by tests/test_sh_factored.py -m gpu)
        vp._combine = lambda means, campos, dcol, d, out: out.copy_(
            TN.sh_grad_combine_ref(means, campos, dcol.reshape(campos.shape[0], -1, 3), d, 16).float().reshape(-1))
        G = torch.randn(3, H, W, generator=torch.Generator().manual_seed(100 + rank))
        vp.step(G)
        got = {k: v.clone() for k, v in vp.grads().items()}
        if rank == 0:
            leaves = {k: v.clone().requires_grad_(True) for k, v in sc.items()}
            losses = []
            for r in range(world):
                c = synth.orbit_camera(r, H, W)
                Gr = torch.randn(3, H, W, generator=torch.Generator().manual_seed(100 + r))
                col, _, _ = _StubRastSH(c, H, W, deg)(leaves["means3D"], torch.zeros(P, 3), leaves["opacities"],
                                                     shs=leaves["shs"], scales=leaves["scales"],
                                                     rotations=leaves["rotations"])
                losses.append((col * Gr).sum())
            (sum(losses) / world).backward()
            ok = all(torch.allclose(got[n], leaves[n].grad.reshape(-1), rtol=1e-4, atol=1e-5) for n, _ in SLAB_FIELDS_SH)
            ok = ok and float(got["shs"].abs().max()) > 0
            ret.put(bool(ok))
    finally:
        dist.destroy_process_group()


def _spawn2(worker, world=2):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert ret.get(timeout=5) is True


def test_view_parallel_factored_exchange_world_3():
    """Odd world size: slot order of the all-gather, 1/world scaling and the rebuild over three views."""
    _spawn2(_worker_factored, world=3)


def test_view_parallel_allreduce_matches_serial_mean():
    _spawn2(_worker)


def test_view_parallel_factored_exchange_matches_serial_mean():
    """all-gather of the per-view colour gradients + all-reduce of the 11 geometry floats + local rebuild of the SH
    rows == the serial multi-view accumulation, SH rows included."""
    _spawn2(_worker_factored)


def test_slab_layout():
    sc = synth.make_scene(16, 1)
    vp = ViewParallelRasterizer(sc, synth.orbit_camera(0, 16, 16), 16, 16, 3, device="cpu")
    g = vp.grads()
    assert [k for k in g] == ["means3D", "opacities", "scales", "rotations", "shs"]
    assert vp.slab.numel() == 59 * 16 and sum(v.numel() for v in g.values()) == 59 * 16
    # slices tile the slab without gaps, in order
    off = 0
    for v in g.values():
        assert v.data_ptr() == vp.slab.data_ptr() + 4 * off
        off += v.numel()


def test_view_time_jobs_cover_every_job_once():
    for n_frames, n_views, world in ((300, 6, 8), (3, 2, 2), (5, 3, 4), (1, 1, 8)):
        seen = []
        rounds = None
        for r in range(world):
            jobs = synth.view_time_jobs(n_frames, n_views, r, world)
            rounds = len(jobs) if rounds is None else rounds
            assert len(jobs) == rounds                       # every rank takes part in every round
            seen += [j for j in jobs if j is not None]
        assert sorted(seen) == [(f, v) for f in range(n_frames) for v in range(n_views)]
    off0, off1 = synth.frame_offset(50, 0.0), synth.frame_offset(50, 1.0)
    assert off0.shape == (50, 3) and float(off0.abs().max()) <= 0.05 + 1e-7
    assert torch.allclose(off0, off1, atol=1e-6)             # period 1 in t
    assert not torch.allclose(off0, synth.frame_offset(50, 0.25), atol=1e-3)


def _worker_view_time(rank, world, port, ret):
    """BASELINE config 4 in miniature: 3 frames x 1 view = 3 jobs over 2 ranks (the last round has an idle rank);
    per-job camera and per-frame means; gradients summed per round == the serial loop over the round's jobs."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P, H, W = 30, 8, 12
        n_frames, n_views = 3, 1
        sc = synth.make_scene(P, 6)
        base = sc["means3D"].clone()
        vp = ViewParallelRasterizer(sc, synth.orbit_camera(0, H, W), H, W, 3, device="cpu", world_size=world,
                                    exchange="allreduce")
        vp.rast_factory = lambda settings, cam: _StubRast(cam, H, W)
        jobs = synth.view_time_jobs(n_frames, n_views, rank, world)
        ok = True
        for rnd, job in enumerate(jobs):
            if job is None:
                vp.idle_step()
            else:
                f, v = job
                vp.set_camera(synth.orbit_camera(v + f, H, W))
                vp.set_means(base + synth.frame_offset(P, f / n_frames))
                G = torch.randn(3, H, W, generator=torch.Generator().manual_seed(7 * f + v))
                vp.step(G)
            got = {k: t.clone() for k, t in vp.grads().items()}
            if rank == 0:
                leaves = {k: t.clone().requires_grad_(True) for k, t in sc.items()}
                losses = []
                for r in range(world):
                    jr = synth.view_time_jobs(n_frames, n_views, r, world)[rnd]
                    if jr is None:
                        continue
                    f, v = jr
                    G = torch.randn(3, H, W, generator=torch.Generator().manual_seed(7 * f + v))
                    m = leaves["means3D"] + synth.frame_offset(P, f / n_frames)
                    col, _, _ = _StubRast(synth.orbit_camera(v + f, H, W), H, W)(
                        m, torch.zeros(P, 3), leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"],
                        rotations=leaves["rotations"])
                    losses.append((col * G).sum())
                (sum(losses) / world).backward()
                ok = ok and all(torch.allclose(got[n], leaves[n].grad.reshape(-1), rtol=1e-4, atol=1e-5)
                                for n, _ in SLAB_FIELDS_SH)
        with_factored_guard = False
        try:
            vpf = ViewParallelRasterizer(sc, synth.orbit_camera(0, H, W), H, W, 3, device="cpu", world_size=world,
                                         exchange="factored")
            vpf.set_means(base)
        except Exception:
            with_factored_guard = True
        if rank == 0:
            ret.put(bool(ok and with_factored_guard))
    finally:
        dist.destroy_process_group()


def test_view_time_sharded_rounds_match_serial_loop():
    _spawn2(_worker_view_time)


def test_slab_layout_follows_sh_coefficient_count():
    sc = synth.make_scene(16, 1, sh_coeffs=9)          # max_sh_degree = 2
    vp = ViewParallelRasterizer(sc, synth.orbit_camera(0, 16, 16), 16, 16, 2, device="cpu")
    assert vp.floats_per_splat == 11 + 27 and vp.grads()["shs"].numel() == 27 * 16
    rgb = synth.make_scene(16, 1, precomp_rgb=True)
    vp = ViewParallelRasterizer(rgb, synth.orbit_camera(0, 16, 16), 16, 16, 0, device="cpu")
    assert [k for k in vp.grads()] == ["means3D", "opacities", "scales", "rotations", "colors_precomp"]
    assert vp.floats_per_splat == 14 and vp.exchange == "allreduce"


# ------------------------------------------------------------------------------------------------ side outputs
class _StubRastSide(_StubRast):
    """Stand-in whose radii and means2D gradient depend on the camera, like the rasterizer's per-view side outputs."""

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        color, _, depth = super().forward(means3D, means2D, opacities, shs=shs, scales=scales, rotations=rotations)
        hom = torch.cat([means3D, torch.ones_like(means3D[:, :1])], 1) @ self.cam.full_proj_transform
        radii = (hom[:, 0].detach() * 7).round().clamp(min=0).to(torch.int32)          # some splats invisible (0)
        color = color + (means2D[:, :2] * hom[:, :2].detach()).sum() * 1e-3
        return color, radii, depth


def _serial_side(sc, world, P, H, W):
    """The reference's serial loop (train.py:169-178): only the LAST view's radii / viewspace gradient survive."""
    leaves = {k: v.clone().requires_grad_(True) for k, v in sc.items()}
    per_view = []
    for r in range(world):
        c = synth.orbit_camera(r, H, W)
        Gr = torch.randn(3, H, W, generator=torch.Generator().manual_seed(100 + r))
        m2d = torch.zeros(P, 3, requires_grad=True)
        col, radii, _ = _StubRastSide(c, H, W)(leaves["means3D"], m2d, leaves["opacities"], shs=leaves["shs"],
                                               scales=leaves["scales"], rotations=leaves["rotations"])
        ((col * Gr).sum() / world).backward()
        per_view.append((radii, m2d.grad.clone()))
    return per_view


def _worker_side(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P, H, W = 40, 8, 12
        sc = synth.make_scene(P, 5)
        cam = synth.orbit_camera(rank, H, W)
        vp = ViewParallelRasterizer(sc, cam, H, W, 3, device="cpu", world_size=world, exchange="allreduce")
        vp.rast = _StubRastSide(cam, H, W)
        G = torch.randn(3, H, W, generator=torch.Generator().manual_seed(100 + rank))
        vp.step(G, keep=True)
        last = vp.side_outputs("last")
        allv = vp.side_outputs("all")
        per_view = _serial_side(sc, world, P, H, W)
        r_last, g_last = per_view[-1]
        ok = torch.equal(last["radii"], r_last) and torch.allclose(last["viewspace_grad"], g_last, rtol=1e-5, atol=1e-7)
        ok = ok and torch.equal(last["visibility_filter"], r_last > 0)
        rmax = torch.stack([r for r, _ in per_view]).max(0).values
        gsum = sum(g[:, :2].norm(dim=1) * (r > 0) for r, g in per_view)
        cnt = sum((r > 0).float() for r, _ in per_view)
        ok = ok and torch.equal(allv["radii"], rmax) and torch.allclose(allv["grad_norm"], gsum, rtol=1e-5, atol=1e-7)
        ok = ok and torch.equal(allv["count"], cnt) and bool((cnt == 0).any()) and bool((cnt > 0).any())
        # the bookkeeping that follows (scene/gaussian_model.py:427-430, train.py:280-282) gives every rank the same state
        acc, den, mr = torch.zeros(P), torch.zeros(P), torch.zeros(P)
        f = last["visibility_filter"]
        acc[f] += last["viewspace_grad"][f, :2].norm(dim=1)
        den[f] += 1
        mr[f] = torch.max(mr[f], last["radii"][f].float())
        state = torch.stack([acc, den, mr])
        ref = state.clone()
        dist.broadcast(ref, src=0)
        ok = ok and torch.equal(state, ref)
        ret.put(bool(ok))
    finally:
        dist.destroy_process_group()


def test_side_outputs_match_the_serial_loop_on_every_rank():
    """radii / viewspace gradient of the LAST view (train.py:178) reach every rank; the 'all' statistics are the
    max / sum over the views; the densification state built from them is identical across ranks."""
    world = 3
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker_side, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert [ret.get(timeout=5) for _ in range(world)] == [True] * world
