"""Host-side drivers above the rasterizer:

* ViewParallelRasterizer — the multi-GPU unit of BASELINE.json's metric: one process per GPU, splats
  replicated, each rank renders ITS camera (the reference renders the views of one iteration serially in
  the loop at train.py:169 and averages the losses at train.py:242); the per-splat gradients of all
  ranks are summed into a flat fp32 slab [59, P].  Three exchanges produce that sum:
  - "nvlink" (default on CUDA with more than one rank): the library's own kernels over symmetric (peer-mapped /
    NVSwitch-multicast) memory — the geometry backward pushes its colour gradients into every rank's table while
    it computes and leaves the 11 (14) parameter gradients as packed records, and ONE kernel (sfb_xchg_finish)
    sums those records through the switch (multimem.ld_reduce / multimem.st), rebuilds the SH rows and unpacks
    the sums into the slab.  No NCCL collective on the data path.
  - "factored": the same factorisation over torch.distributed collectives (all-gather of the [P, 3] colour
    gradients + all-reduce of the [P, 11] geometry gradients, SH rows rebuilt by sfb_sh_grad_combine); runs on
    any backend (the gloo tests use it).
  - "allreduce": ONE all-reduce of the whole slab (the plain formulation; per-rank means).
  The SH gradient of one view is the rank-1 block basis(dir) (x) dL_dcolour and dir is known to every rank,
  which is why 3 floats per splat and view travel instead of the 48 SH floats.
* forward_backward_host — the same step for callers that hold HOST buffers: pinned host -> device copies
  of every input, forward + backward, device -> pinned host copies of image, depth, radii and gradients.

torch is plumbing here (device memory, streams, torch.distributed); the compute is libsplat_b200.so.
"""
from __future__ import annotations

import math
import os

import torch

from . import _lib, rasterizer
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer

# (name, floats per splat) in slab order
SLAB_FIELDS_SH = (("means3D", 3), ("opacities", 1), ("scales", 3), ("rotations", 4), ("shs", 48))
SLAB_FIELDS_RGB = (("means3D", 3), ("opacities", 1), ("scales", 3), ("rotations", 4), ("colors_precomp", 3))


class ViewParallelRasterizer:
    """One rank of the view-parallel step: holds a replica of the splats, renders this rank's camera, and sums the
    gradient slab over the ranks (module docstring).  `step(cotangent)` = forward + backward + exchange;
    `grads()` = views into the summed slab, one per parameter."""

    def __init__(self, scene: dict, camera, H: int, W: int, sh_degree: int, device, world_size: int = 1,
                 bg=(1.0, 1.0, 1.0), exchange: str | None = None):
        self.device = torch.device(device)
        self.world = int(world_size)
        self.rank = 0
        if self.world > 1:
            import torch.distributed as dist
            self.rank = dist.get_rank()
        self.factored_output = None      # single rank: a [P*3] tensor here keeps the SH gradient factored (HostPipeline)
        self.sh_degree = int(sh_degree)
        exchange = exchange or os.environ.get("SFB_EXCHANGE") or ("nvlink" if self.device.type == "cuda" else "factored")
        if exchange not in ("nvlink", "factored", "allreduce"):
            raise Exception("exchange must be 'nvlink', 'factored' or 'allreduce'")
        # a single rank has nothing to exchange; the factored form exists for SH colours only
        if self.world <= 1 or (exchange == "factored" and "shs" not in scene):
            exchange = "allreduce"
        self.exchange = exchange
        self.H, self.W = H, W
        self.P = scene["means3D"].shape[0]
        self.params = {k: v.detach().to(self.device, copy=True).contiguous().requires_grad_(True)
                       for k, v in scene.items()}
        cam = camera.to(self.device)
        self.settings = GaussianRasterizationSettings(
            image_height=H, image_width=W, tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5),
            bg=torch.tensor(bg, dtype=torch.float32, device=self.device), scale_modifier=1.0,
            viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, sh_degree=sh_degree,
            campos=cam.camera_center, prefiltered=False, debug=False)
        self.rast_factory = lambda settings, cam: GaussianRasterizer(settings)   # (the CPU tests plug a stand-in in)
        self.rast = self.rast_factory(self.settings, cam)
        self.means2D = torch.zeros(self.P, 3, device=self.device, requires_grad=True)
        if "shs" in scene:       # 3 * M floats of SH per splat (M = 16 for the reference's max_sh_degree = 3)
            self.fields = SLAB_FIELDS_SH[:-1] + (("shs", 3 * int(scene["shs"].shape[1])),)
        else:
            self.fields = SLAB_FIELDS_RGB
        self.floats_per_splat = sum(n for _, n in self.fields)
        # the flat gradient slab: field-major [sum(n), P] so every field is one contiguous run
        self.slab = torch.empty(self.floats_per_splat * self.P, dtype=torch.float32, device=self.device)
        self.last = None
        self._combine = rasterizer.sh_grad_combine
        self.time_exchange = False       # bench: record CUDA events around the gradient exchange of every step
        self.exchange_events = []
        self.xchg = None
        self.exchange_fallback_reason = None
        if self.exchange == "nvlink":
            try:
                self._setup_nvlink("shs" in scene)
            except Exception as ex:      # no symmetric-memory support on this system: the NCCL formulation of the same sum
                import sys
                self.exchange_fallback_reason = f"{type(ex).__name__}: {ex}"
                self.exchange = "factored" if "shs" in scene else "allreduce"
                sys.stderr.write(f"splatfields_b200: NVLink exchange unavailable ({self.exchange_fallback_reason}); "
                                 f"using exchange='{self.exchange}' over torch.distributed\n")
        if self.exchange in ("factored", "nvlink") and "shs" in scene:
            self._gather_campos(cam)
        if self.exchange == "factored":
            assert self.fields[-1][0] == "shs"          # the SH rows are the tail of the slab
            self.geo_floats = self.floats_per_splat - self.fields[-1][1]
            self.dcolor_mine = torch.empty(self.P * 3, dtype=torch.float32, device=self.device)
            self.dcolor_views = torch.empty(self.world * self.P * 3, dtype=torch.float32, device=self.device)

    def _setup_nvlink(self, has_sh: bool) -> None:
        """Symmetric buffers of the NVLink exchange (include/splat_b200.h: sfb_xchg): torch allocates and maps them
        (torch.distributed._symmetric_memory — plumbing), the library's kernels do the rest."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        lib = _lib.load()
        ngeo = 12 if has_sh else 16
        nbytes = int(lib.sfb_xchg_bytes(self.P, self.world, ngeo, int(has_sh)))
        buf = symm.empty(nbytes, dtype=torch.uint8, device=self.device)
        buf.zero_()
        hdl = symm.rendezvous(buf, dist.group.WORLD)
        d = _lib.XchgDesc()
        d.rank, d.world, d.P, d.ngeo = int(hdl.rank), int(hdl.world_size), self.P, ngeo
        d.local = buf.data_ptr()
        for r, ptr in enumerate(hdl.buffer_ptrs):
            d.peers[r] = int(ptr)
        mc = int(hdl.multicast_ptr) if os.environ.get("SFB_XCHG_NO_MULTICAST", "0") != "1" else 0
        d.mc = mc if mc else None
        d.max_ctas = 0
        d.campos_views = None
        # SFB_XCHG_FUSED=1: one persistent kernel does the geometry backward AND the exchange chunk by chunk (measured
        # slower than backward + sfb_xchg_finish on B200: the per-chunk flag releases wait for the NVLink stores they
        # cover, profiles/r02d_*); default: the two kernels
        self.xchg_fused = os.environ.get("SFB_XCHG_FUSED", "0") == "1"
        self.xchg, self._xchg_buf, self._xchg_hdl = d, buf, hdl
        self.xchg_epoch = 0
        self.xchg_multicast = bool(mc)
        torch.cuda.synchronize(self.device)
        dist.barrier()                       # every rank's buffer is zeroed and mapped before the first step

    def _gather_campos(self, cam) -> None:
        """Every rank needs every camera centre (3 floats per view) to rebuild the SH rows: one tiny all-gather per
        camera change."""
        import torch.distributed as dist
        mine = cam.camera_center.detach().to(self.device, torch.float32).reshape(3).contiguous()
        allc = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allc, mine)
        self.campos_views = torch.stack(allc).contiguous()

    # -- per-job state for (view, time)-sharded runs (BASELINE.json configs[4]): a new camera and / or new means
    def set_camera(self, camera) -> None:
        """Render another camera from now on.  In factored mode every rank must call this in the same step (the
        camera centres are re-gathered)."""
        cam = camera.to(self.device)
        self.settings = self.settings._replace(
            tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5), viewmatrix=cam.world_view_transform,
            projmatrix=cam.full_proj_transform, campos=cam.camera_center)
        self.rast = self.rast_factory(self.settings, cam)
        if self.exchange in ("factored", "nvlink") and "shs" in self.params:
            self._gather_campos(cam)

    def set_means(self, means3D: torch.Tensor, same_on_all_ranks: bool = False) -> None:
        """Replace the means, e.g. canonical means + this job's per-frame offset.  The factored exchange rebuilds the
        SH rows from the means THIS rank holds, so it needs identical means on every rank; ranks that render
        different time steps must use exchange="allreduce" (precomputed colours, the 4D recipe's input, always do)."""
        if self.exchange in ("factored", "nvlink") and "shs" in self.params and not same_on_all_ranks:
            raise Exception("per-rank means with the factored SH exchange: construct with exchange='allreduce', "
                            "or pass same_on_all_ranks=True if every rank sets the same means")
        with torch.no_grad():
            self.params["means3D"].copy_(means3D.to(self.device))

    def idle_step(self) -> int:
        """A round in which this rank has no job: contribute zeros to the exchange (collectives stay symmetric)."""
        self.slab.zero_()
        if self.exchange == "factored":
            self.dcolor_mine.zero_()
        if self.exchange == "nvlink":
            # zero records and a zero slot in every rank's colour table: a backward over zero cotangents does that
            raise Exception("idle_step is not available with exchange='nvlink': give every rank a job per round")
        return self._exchange(0)

    # -- one fwd + bwd (+ gradient exchange); returns the number of library kernel launches issued
    def step(self, cotangent: torch.Tensor, keep: bool = False) -> int:
        p = self.params
        for v in p.values():
            v.grad = None
        self.means2D.grad = None
        lib = _lib.load()
        color, radii, depth = self.rast(means3D=p["means3D"], means2D=self.means2D, opacities=p["opacities"],
                                        shs=p.get("shs"), colors_precomp=p.get("colors_precomp"),
                                        scales=p["scales"], rotations=p["rotations"])
        n = lib.sfb_last_launch_count()
        factored = self.exchange == "factored"
        nvlink = self.exchange == "nvlink"
        fused = nvlink and self.xchg_fused
        if nvlink:
            self.xchg_epoch += 1
            if fused and "shs" in p:
                self.xchg.campos_views = self.campos_views.data_ptr()
        sh_out = self.dcolor_mine if factored else (self.factored_output if self.world == 1 else None)
        rasterizer.set_grad_arena(self.slab, self.fields, None if sh_out is None else sh_out.view(self.P, 3),
                                  (self.xchg, self.xchg_epoch, fused) if nvlink else None)
        try:
            # mean over views (train.py:242) folded into the cotangent: backward is linear in it
            color.backward(cotangent if self.world == 1 else cotangent * (1.0 / self.world))
        finally:
            rasterizer.set_grad_arena(None, None)
        n += lib.sfb_last_launch_count()
        # normally a no-op: the backward kernel already wrote into the slab slices (the .grad tensors ARE
        # those slices); copy only if autograd handed back separate storage.
        for name, dst in self.grads().items():
            if nvlink or (sh_out is not None and name == "shs"):
                continue                     # filled by the exchange below / kept factored
            g = p[name].grad
            if g is None:
                dst.zero_()
            elif g.data_ptr() != dst.data_ptr():
                dst.copy_(g.reshape(-1))
        n = self._exchange(n)
        if keep:
            self.last = (color.detach(), radii, depth.detach())
        return n

    def _exchange(self, n: int) -> int:
        """Sum the gradient slab over the ranks (module docstring); n = launch counter to continue."""
        p = self.params
        factored = self.exchange == "factored"
        ev = None
        if self.time_exchange and self.world > 1 and self.device.type == "cuda":
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        if self.exchange == "nvlink" and self.xchg_fused:
            pass                             # summed inside the backward call (fused kernel)
        elif self.exchange == "nvlink":
            lib = _lib.load()
            g = self.grads()
            has_sh = "shs" in p
            M = p["shs"].shape[1] if has_sh else 0
            gp = lambda name: g[name].data_ptr() if name in g else None
            with torch.cuda.device(self.device):
                stream = torch.cuda.current_stream(self.device).cuda_stream
                import ctypes as C
                _lib.check(lib.sfb_xchg_finish(
                    C.byref(self.xchg), int(self.xchg_epoch), self.sh_degree, int(M),
                    p["means3D"].data_ptr(), self.campos_views.data_ptr() if has_sh else None,
                    gp("means3D"), gp("opacities"), gp("scales"), gp("rotations"), gp("colors_precomp"), gp("shs"),
                    stream))
            n += 1
        elif factored:
            import torch.distributed as dist
            # all-gather 3 floats / splat / view, all-reduce the 11 geometry floats; the SH rows are rebuilt
            # locally while the all-reduce is still in flight (it only depends on the all-gather)
            h_ag = dist.all_gather_into_tensor(self.dcolor_views, self.dcolor_mine, async_op=True)
            h_ar = dist.all_reduce(self.slab[:self.geo_floats * self.P], op=dist.ReduceOp.SUM, async_op=True)
            h_ag.wait()
            sh_out = self.slab[self.geo_floats * self.P:]
            self._combine(p["means3D"].detach(), self.campos_views, self.dcolor_views, self.sh_degree, sh_out)
            h_ar.wait()
            n += 1                           # the rebuild kernel (the NCCL kernels are not counted as ours)
        elif self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.slab, op=dist.ReduceOp.SUM)
        if ev is not None:
            ev[1].record()
            self.exchange_events.append(ev)
        return n

    def side_outputs(self, mode: str = "last") -> dict:
        """The per-view side outputs the densification bookkeeping consumes (radii -> max_radii2D, train.py:280-286;
        viewspace_points.grad -> add_densification_stats, train.py:306-311, scene/gaussian_model.py:427-430), made
        IDENTICAL on every rank so that the replicated splat sets take the same clone / split / prune decisions.
        Call after step(..., keep=True).
          mode="last": the last view's values (rank world-1), i.e. exactly what the reference's serial loop keeps
                       (train.py:178 overwrites render_pkg per view): dict(radii [P] int32, visibility_filter [P] bool,
                       viewspace_grad [P,3]) — feed them to densify.add_densification_stats as they are.
          mode="all":  statistics over ALL views of the step: radii = max over views, grad_norm = sum over the views of
                       ||viewspace_grad[:, :2]|| where the splat was visible, count = number of such views
                       (xyz_gradient_accum += grad_norm; denom += count).
        Two small collectives (16 B/splat), off the critical path of the gradient exchange."""
        import torch.distributed as dist
        if self.last is None:
            raise Exception("side_outputs() needs step(..., keep=True)")
        radii = self.last[1].to(torch.int32).contiguous()
        g2 = self.means2D.grad
        g2 = torch.zeros(self.P, 3, device=self.device) if g2 is None else g2.detach().contiguous()
        if mode == "last":
            if self.world > 1:
                radii, g2 = radii.clone(), g2.clone()
                dist.broadcast(radii, src=self.world - 1)
                dist.broadcast(g2, src=self.world - 1)
            return dict(radii=radii, visibility_filter=radii > 0, viewspace_grad=g2)
        if mode != "all":
            raise Exception("mode must be 'last' or 'all'")
        vis = (radii > 0).to(torch.float32)
        stats = torch.stack([g2[:, :2].norm(dim=1) * vis, vis])           # [2, P]
        rmax = radii.clone()
        if self.world > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
            dist.all_reduce(rmax, op=dist.ReduceOp.MAX)
        return dict(radii=rmax, visibility_filter=rmax > 0, grad_norm=stats[0], count=stats[1])

    def exchange_status(self) -> int:
        """0, or the NVLink exchange's error word (include/splat_b200.h: sfb_xchg_status); synchronises the stream."""
        if self.xchg is None:
            return 0
        import ctypes as C
        st = C.c_uint(0)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().sfb_xchg_status(C.byref(self.xchg), C.byref(st),
                                                   torch.cuda.current_stream(self.device).cuda_stream))
        return int(st.value)

    def exchange_timeline(self) -> dict | None:
        """Device timeline (microseconds from the kernel's first CTA) of the last step's exchange kernel on this rank
        (include/splat_b200.h: sfb_xchg_timeline); synchronises the stream.  Fused kernel: when the last geometry chunk,
        the last NVLink unit, the last SH chunk and the last unpack finished; sfb_xchg_finish: its five phases."""
        if self.xchg is None:
            return None
        import ctypes as C
        t = (C.c_ulonglong * 12)()
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().sfb_xchg_timeline(C.byref(self.xchg), t,
                                                     torch.cuda.current_stream(self.device).cuda_stream))
        t = [int(v) for v in t]
        names = (("geometry_done", "nvlink_units_done", "sh_rows_done", "unpack_done", "kernel_end") if self.xchg_fused
                 else ("barrier_a", "slice_reduced", "sh_rows_done", "barrier_b", "kernel_end"))
        out = {n: (round((v - t[0]) / 1e3, 1) if v else None) for n, v in zip(names, t[1:6])}
        if self.xchg_fused and sum(t[6:]) > 0:      # share of the CTAs' time per kind of work
            tot = float(sum(t[6:]))
            out["cta_time_share"] = {n: round(v / tot, 3) for n, v in
                                     zip(("choosing_or_waiting", "geometry", "flag_release", "nvlink_units", "sh_rows", "unpack"), t[6:])}
        return out

    def exchange_ms(self) -> list:
        """Device time of the gradient exchange (collectives + SH rebuild) of the steps run with time_exchange."""
        torch.cuda.synchronize(self.device)
        out = [a.elapsed_time(b) for a, b in self.exchange_events]
        self.exchange_events = []
        return out

    def bytes_on_wire_per_splat(self) -> float:
        """Bytes each rank sends (= receives) per splat and step for the gradient exchange (ring / NVSwitch model:
        all-reduce 2(N-1)/N * size, all-gather (N-1) * size)."""
        N = self.world
        if N <= 1:
            return 0.0
        if self.exchange == "nvlink":     # received: the other views' colour gradients + this rank's share of the
            ngeo = self.xchg.ngeo         # switch-reduced records and everybody else's broadcast sums
            gc = (N - 1) * 12.0 if "shs" in self.params else 0.0
            return gc + 4.0 * ngeo * (1.0 / N + (N - 1.0) / N)
        if self.exchange == "factored":
            return 2.0 * (N - 1) / N * 4 * self.geo_floats + (N - 1) * 12.0
        return 2.0 * (N - 1) / N * 4 * self.floats_per_splat

    def grads(self) -> dict:
        out, off = {}, 0
        for name, n in self.fields:
            out[name] = self.slab[off * self.P:(off + n) * self.P]
            off += n
        return out

    # -- tile-list statistics of this rank's view (for the bench's roofline accounting)
    def list_stats(self) -> dict:
        lib = _lib.load()
        p = self.params
        color, radii, depth = self.rast(means3D=p["means3D"], means2D=self.means2D, opacities=p["opacities"],
                                        shs=p.get("shs"), colors_precomp=p.get("colors_precomp"),
                                        scales=p["scales"], rotations=p["rotations"])
        fn = color.grad_fn
        radii_s, geom, binning, img = fn.saved_tensors[:4]
        R = fn.num_rendered
        H, W = self.H, self.W
        gx, gy = (W + 15) // 16, (H + 15) // 16
        T = gx * gy
        ranges = torch.zeros(T, 2, dtype=torch.int32, device=self.device)
        _lib.check(lib.sfb_export_binning(self.P, R, W, H, geom.data_ptr(), binning.data_ptr(), None, None,
                                          ranges.data_ptr(), None))
        nc = torch.zeros(H, W, dtype=torch.int32, device=self.device)
        _lib.check(lib.sfb_export_img(W, H, img.data_ptr(), None, nc.data_ptr(), None))
        torch.cuda.synchronize()
        lens = (ranges[:, 1] - ranges[:, 0]).long()
        pad = torch.zeros(gy * 16, gx * 16, dtype=torch.int64, device=self.device)
        pad[:H, :W] = nc.long()
        tmax = pad.reshape(gy, 16, gx, 16).amax(dim=(1, 3)).reshape(-1)      # deepest contributor per tile
        # entries the forward sweeps: whole 256-batches until every pixel of the tile is done (<= list length);
        # entries the backward sweeps: up to the deepest last contributor.
        r_bwd = int(torch.minimum(tmax, lens).sum())
        r_fwd = int(torch.minimum(((tmax + 256) // 256) * 256, lens).sum())
        return dict(R=int(R), visible=int((radii > 0).sum()), mean_list=float(lens.float().mean()),
                    max_list=int(lens.max()), R_fwd=r_fwd, R_bwd=r_bwd)

    # -- pinned host mirrors of every input and output of one step
    def shard_rows(self):
        """Rows [r0, r1) of the P splats this rank moves between host and device in sharded host-buffer steps."""
        per = (self.P + self.world - 1) // self.world
        r = self.rank if self.world > 1 else 0
        return min(r * per, self.P), min((r + 1) * per, self.P), per

    def pinned_host_buffers(self, scene: dict, cotangent: torch.Tensor, shard: bool = False, factored: bool = False):
        """Pinned host mirrors of one step.  shard=True (world > 1): the job's ONE host copy of the splats is split by
        rows over the ranks — this rank's buffers hold rows shard_rows() of every parameter and of every gradient
        field.  factored=True (world == 1): the SH gradient stays in its factored form (grads = 11 geometry floats +
        3 colour-gradient floats per splat, rebuilt on demand by sh_rows_from_factored)."""
        r0, r1, _ = self.shard_rows() if shard else (0, self.P, self.P)
        with numa_local(self.device):       # first touch on the GPU's own NUMA node
            pin = lambda t: t.detach().cpu().contiguous().pin_memory()
            host_in = {k: pin(v[r0:r1]) for k, v in scene.items()}
            host_in["dL_dcolor"] = pin(cotangent)
            nfl = (self.floats_per_splat - (self.fields[-1][1] - 3 if factored else 0)) * (r1 - r0)
            host_out = dict(color=torch.empty(3, self.H, self.W).pin_memory(),
                            depth=torch.empty(1, self.H, self.W).pin_memory(),
                            radii=torch.empty(self.P, dtype=torch.int32).pin_memory(),
                            grads=torch.empty(nfl).pin_memory())
        self._cot_dev = torch.empty_like(cotangent, device=self.device)
        return host_in, host_out


class numa_local:
    """Context manager: run the calling thread on the CPUs next to `device` (NVML's ideal affinity), so that pinned
    host buffers allocated inside land on the GPU's own NUMA node (first touch); restores the affinity afterwards.
    A no-op when NVML or the affinity calls are unavailable."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.saved = None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                tok = vis.split(",")[idx].strip()
                h = pynvml.nvmlDeviceGetHandleByIndex(int(tok)) if tok.isdigit() else pynvml.nvmlDeviceGetHandleByUUID(tok)
            else:
                h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.saved = os.sched_getaffinity(0)
            pynvml.nvmlDeviceSetCpuAffinity(h)
            self.cpus = sorted(os.sched_getaffinity(0))
        except Exception:
            self.saved = None
        return self

    def __exit__(self, *exc):
        if self.saved is not None:
            try:
                os.sched_setaffinity(0, self.saved)
            except Exception:
                pass
        return False


def sh_rows_from_factored(means3D, campos, dcolor, sh_degree, M=16):
    """dL_dsh [P, M, 3] from the factored gradient of ONE view (host or device tensors, plain torch): the rows the
    rasterizer would have written — basis(normalize(means3D - campos)) (x) dcolor.  For host-side consumers of
    HostPipeline(factored=True) that need the rows of some splats."""
    d = means3D - campos.reshape(1, 3)
    d = d / d.norm(dim=1, keepdim=True)
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    C0, C1 = 0.28209479177387814, 0.4886025119029199
    C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
    C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
          1.445305721320277, -0.5900435899266435)
    b = [torch.full_like(x, C0)]
    if sh_degree > 0:
        b += [-C1 * y, C1 * z, -C1 * x]
    if sh_degree > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        b += [C2[0] * xy, C2[1] * yz, C2[2] * (2 * zz - xx - yy), C2[3] * xz, C2[4] * (xx - yy)]
    if sh_degree > 2:
        b += [C3[0] * y * (3 * xx - yy), C3[1] * xy * z, C3[2] * y * (4 * zz - xx - yy),
              C3[3] * z * (2 * zz - 3 * xx - 3 * yy), C3[4] * x * (4 * zz - xx - yy), C3[5] * z * (xx - yy),
              C3[6] * x * (xx - 3 * yy)]
    B = torch.stack(b, dim=1)                                     # [P, (deg+1)^2]
    out = torch.zeros(means3D.shape[0], M, 3, dtype=means3D.dtype, device=means3D.device)
    out[:, :B.shape[1], :] = B.unsqueeze(2) * dcolor.unsqueeze(1)
    return out


def forward_backward_host(vp: ViewParallelRasterizer, host_in: dict, host_out: dict) -> None:
    """HOST buffers in, HOST buffers out, nothing overlapped: H2D of all splat parameters + cotangent, fwd + bwd
    (+ exchange), D2H of image, depth, radii and the gradient slab; returns when the host buffers are valid.
    (Whole-scene buffers: pinned_host_buffers(shard=False).)"""
    with torch.no_grad():
        for k, v in vp.params.items():
            v.copy_(host_in[k], non_blocking=True)
        vp._cot_dev.copy_(host_in["dL_dcolor"], non_blocking=True)
    vp.step(vp._cot_dev, keep=True)
    color, radii, depth = vp.last
    host_out["color"].copy_(color, non_blocking=True)
    host_out["depth"].copy_(depth, non_blocking=True)
    host_out["radii"].copy_(radii, non_blocking=True)
    host_out["grads"].copy_(vp.slab, non_blocking=True)
    torch.cuda.current_stream(vp.device).synchronize()


class HostPipeline:
    """Asynchronous HOST-buffer front end: `submit(host_in, host_out)` / `wait(ticket)`.

    Every step moves all of its inputs host->device and all of its results device->host, on three streams with
    double-buffered device staging, so that the upload of step k+1, the kernels of step k and the download of step
    k-1 overlap (PCIe is full duplex; the kernels take ~1/4 of either copy).

    shard=True (the default with more than one rank): the job has ONE host copy of the splats, so every rank uploads
    rows shard_rows() of each parameter and the ranks all-gather the rest over NVLink (5 in-place NCCL all-gathers on
    the compute stream) instead of pushing the same 236 B/splat through the host N times; likewise every rank
    downloads its rows of the SUMMED gradient slab (identical on all ranks) plus its own image, depth and radii.
    factored=True (single rank): the SH gradient is downloaded in its factored form (3 floats per splat,
    SFB_BWD_SH_FACTORED; sh_rows_from_factored rebuilds rows on demand): 56 instead of 236 bytes per splat."""

    def __init__(self, vp: ViewParallelRasterizer, depth: int = 2, shard: bool | None = None, factored: bool = False):
        self.vp = vp
        dev = vp.device
        self.depth = depth
        self.shard = (vp.world > 1) if shard is None else bool(shard and vp.world > 1)
        self.factored = bool(factored)
        if self.factored and (vp.world > 1 or "shs" not in vp.params):
            raise Exception("factored host output is the single-rank SH path")
        self.r0, self.r1, self.per = vp.shard_rows() if self.shard else (0, vp.P, vp.P)
        self.s_h2d = torch.cuda.Stream(dev)
        self.s_comp = torch.cuda.Stream(dev)
        self.s_d2h = torch.cuda.Stream(dev)
        # staging: per parameter [world * per, ...] rows so that every rank's shard has the same size (in-place all-gather)
        rows = self.per * vp.world if self.shard else vp.P
        self.in_dev = [{k: torch.empty((rows,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev)
                        for k, v in vp.params.items()} for _ in range(depth)]
        self.cot_dev = [torch.empty(3, vp.H, vp.W, device=dev) for _ in range(depth)]
        self.slabs = [torch.empty_like(vp.slab) for _ in range(depth)]
        self.dcol = [torch.empty(vp.P * 3, dtype=torch.float32, device=dev) for _ in range(depth)] if self.factored else None
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_comp = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]
        self.live = [None] * depth       # keeps the per-step output tensors alive until their D2H is done
        self.n = 0
        for s in (self.s_h2d, self.s_comp, self.s_d2h):
            s.wait_stream(torch.cuda.current_stream(dev))

    def h2d_bytes(self, host_in: dict) -> int:
        return sum(t.numel() * t.element_size() for t in host_in.values())

    def d2h_bytes(self, host_out: dict) -> int:
        return sum(t.numel() * t.element_size() for t in host_out.values())

    def submit(self, host_in: dict, host_out: dict) -> int:
        vp, k = self.vp, self.n
        b = k % self.depth
        r0, r1 = self.r0, self.r1
        # upload into staging set b (free once the compute that last read it has finished)
        with torch.cuda.stream(self.s_h2d):
            if k >= self.depth:
                self.s_h2d.wait_event(self.ev_comp[b])
            for name, t in self.in_dev[b].items():
                t[r0:r1].copy_(host_in[name], non_blocking=True)
            self.cot_dev[b].copy_(host_in["dL_dcolor"], non_blocking=True)
            self.ev_in[b].record(self.s_h2d)
        # compute on the staged inputs, gradients into slab b (free once its previous download is done)
        with torch.cuda.stream(self.s_comp):
            self.s_comp.wait_event(self.ev_in[b])
            if k >= self.depth:
                self.s_comp.wait_event(self.ev_out[b])
            if self.shard:
                import torch.distributed as dist
                for name, t in self.in_dev[b].items():      # in place: this rank's rows sit where the gather puts them
                    flat = t.view(-1)
                    n = flat.numel() // vp.world
                    dist.all_gather_into_tensor(flat, flat[vp.rank * n:(vp.rank + 1) * n])
            for name, p in vp.params.items():
                p.data = self.in_dev[b][name][:vp.P]
            vp.slab = self.slabs[b]
            if self.factored:
                vp.factored_output = self.dcol[b]
            vp.step(self.cot_dev[b], keep=True)
            color, radii, depth = vp.last
            self.ev_comp[b].record(self.s_comp)
        with torch.cuda.stream(self.s_d2h):
            self.s_d2h.wait_event(self.ev_comp[b])
            for t in (color, depth, radii, self.slabs[b]):
                t.record_stream(self.s_d2h)      # allocated on the compute stream, read on this one
            host_out["color"].copy_(color, non_blocking=True)
            host_out["depth"].copy_(depth, non_blocking=True)
            host_out["radii"].copy_(radii, non_blocking=True)
            # gradient fields, rows [r0, r1) of each (the slab is field-major: every field is one [P, n] run)
            off_dev, off_host = 0, 0
            for name, nf in vp.fields:
                if self.factored and name == "shs":
                    src = self.dcol[b]
                    nf_out = 3
                else:
                    src = self.slabs[b][off_dev * vp.P:(off_dev + nf) * vp.P]
                    nf_out = nf
                cnt = (r1 - r0) * nf_out
                host_out["grads"][off_host:off_host + cnt].copy_(src[r0 * nf_out:r1 * nf_out], non_blocking=True)
                off_dev += nf
                off_host += cnt
            self.ev_out[b].record(self.s_d2h)
        self.live[b] = (color, radii, depth)
        self.n += 1
        return k

    def wait(self, ticket: int) -> None:
        """Returns when the host_out buffers given to submit(ticket) are valid."""
        self.ev_out[ticket % self.depth].synchronize()

    def drain(self) -> None:
        for s in (self.s_h2d, self.s_comp, self.s_d2h):
            s.synchronize()
