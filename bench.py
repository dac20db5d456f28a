#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the splat-rasterizer hot path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one forward + backward pass of the rasterizer over the workload BASELINE.json's metric is
quoted on: 1 000 000 synthetic Gaussians, 800x800, SH degree 3 (configs[2], "lego_1m"), one camera per
rank (view-parallel, splats replicated), followed for N > 1 by the exchange that sums the per-splat gradient
slab [P, 59] over the ranks (default "nvlink": the library's own kernels over symmetric memory — colour gradients
pushed to every rank by the geometry backward, packed [P, 11] records summed in the switch, SH rows rebuilt on every
rank; SFB_EXCHANGE=factored: the same factorisation over NCCL; SFB_EXCHANGE=allreduce: one all-reduce of the whole slab).
metric = Msplats/s = N * P / t_step.

Printed by rank 0 as one JSON line.  Extra objects: roofline (dominant kernel, live CUDA-event timing on
the launching stream), cpu_baseline (the oracle port timed on the host cores, N=1 only), e2e (same metric
through the public host-buffer API with the host<->device copies inside the timed region), clocks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

WORKLOAD = "lego_1m"
METRIC = "Msplats/s fwd+bwd @ 1M Gaussians 800\u00d7800; views/s at 1/2/4/8 B200"
UNIT = "Msplats/s"
SH_DEGREE = 3


# ----------------------------------------------------------------------------------------- helpers
def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self, wait_first_sample_s: float = 5.0):
        """Start polling and wait until nvidia-smi has delivered its first sample, so that its start-up
        (NVML initialisation takes the driver lock) never falls inside a timed region."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < wait_first_sample_s:
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def mark(self):
        """Samples taken from now on belong to the timed regions."""
        self.first = len(self.lines)

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines[getattr(self, "first", 0):]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_workload(rank):
    from splatfields_b200 import synth
    cfg = synth.CONFIGS[WORKLOAD]
    sc = synth.make_scene(cfg["P"], cfg["seed"], scale_mult=cfg["scale_mult"], precomp_rgb=cfg["precomp_rgb"])
    cam = synth.config_camera(WORKLOAD, rank % 8)
    G = torch.randn(3, cfg["H"], cfg["W"], generator=torch.Generator().manual_seed(1000 + rank))
    return cfg, sc, cam, G


# ----------------------------------------------------------------------------------------- reference arm
def oracle_step(O, sc, cam, cfg, G):
    """One fwd+bwd of the CPU oracle port on the full workload (all host threads)."""
    from tests.helpers import run_oracle
    t0 = time.perf_counter()
    f, b = run_oracle(O, sc, cam, cfg["H"], cfg["W"], (1.0, 1.0, 1.0), SH_DEGREE, dL=G.numpy(), want_margin=False)
    return time.perf_counter() - t0, f


def run_reference(args):
    rank, world, _ = _dist_env()
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    O.set_num_threads(cores)
    cfg, sc, cam, G = make_workload(0)
    # A CPU step takes about a second: the run is bounded in wall time (REF_ARM_BUDGET_S, default 150 s of timed
    # steps + at most 3 warm-up steps) so that any --steps K / --warmup W ends within a few minutes; the line says
    # how many full steps were actually timed.
    budget_s = float(os.environ.get("REF_ARM_BUDGET_S", "150"))
    for _ in range(min(args.warmup, 3)):
        oracle_step(O, sc, cam, cfg, G)
    ts = []
    for _ in range(args.steps):
        t, _f = oracle_step(O, sc, cam, cfg, G)
        ts.append(t)
        if sum(ts) > budget_s:
            break
    t_step = float(np.mean(ts))
    value = cfg["P"] / t_step / 1e6
    sample = (f"{len(ts)} full fwd+bwd steps of {WORKLOAD} (P={cfg['P']}, {cfg['H']}x{cfg['W']}, SH deg 3) timed"
              + (f"; stopped at the {budget_s:.0f} s wall-time bound, {args.steps} were asked for" if len(ts) < args.steps else ""))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "timed_steps": len(ts), "ms_per_step": t_step * 1e3,
        "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "P": cfg["P"], "H": cfg["H"], "W": cfg["W"], "sh_degree": SH_DEGREE,
                   "note": "the reference ships no CPU rasterizer and its CUDA rasterizer is an un-vendored "
                           "dependency: this arm is the C/OpenMP oracle port of that algorithm on the host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": O.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- our arm
def algorithmic_bytes(kernel, P, R, R_fwd, R_bwd, HW, n_visible):
    """Algorithmic bytes per launch of each kernel (per-unit figures from SURVEY.md §8d)."""
    Bg_pre, Bg_bwd = 236 + 83, 303 + 256
    table = {
        "preprocess": P * Bg_pre,
        "geom_backward": P * Bg_bwd,
        "geom_backward_exchange": P * Bg_bwd,
        "render_forward": R_fwd * 44 + HW * 24,
        "render_backward": R_bwd * (40 + 72) + HW * 20,
        "duplicate": R * 8 + P * 16,                       # write (tile, idx) pairs; read rect + count + rank
        "tile_ranges": R * 4,
        "zero_grad_acc": P * 48,
        "instance_block_sums": P * 8,
    }
    if kernel in table:
        return table[kernel]
    if kernel.endswith(".hist"):
        return (R if kernel.startswith("tile") else P) * 4
    if kernel.endswith(".scatter"):
        return (R if kernel.startswith("tile") else P) * 16   # read + write one 8-byte (key, value) pair
    return 0


def roofline_report(acc, nprof, stats, P, HW, ms_step):
    """(roofline, stages) objects of the bench line from the per-kernel event timings of the profiled steps.
    acc: {kernel name: [ms of every launch over nprof steps]}; stats: ViewParallelRasterizer.list_stats()."""
    # a kernel name can occur several times per step (radix passes): per-launch average duration
    per_step = {k: sum(v) / nprof for k, v in acc.items()}
    per_launch = {k: sum(v) / len(v) for k, v in acc.items()}
    dom = max(per_step, key=per_step.get)
    peak, peak_src = measured_peak_hbm()
    ab = algorithmic_bytes(dom, P, stats["R"], stats["R_fwd"], stats["R_bwd"], HW, stats["visible"])
    ach = ab / (per_launch[dom] * 1e-3) / 1e9
    traffic_tbl = {}
    try:     # DRAM bytes per launch from the committed `ncu --set full` capture (profiles/)
        traffic_tbl = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass
    ncu_stats = {}
    try:     # issue-slot utilisation / DRAM % of the same capture: what actually bounds each kernel
        ncu_stats = json.load(open(os.path.join(ROOT, "profiles", "ncu_kernel_stats.json")))
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic_tbl.get(dom), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(ab), "kernel_ms_per_launch": per_launch[dom],
                "kernel_share_of_step": per_step[dom] / max(sum(per_step.values()), 1e-9),
                "ncu": ncu_stats.get(dom),
                "note": "the two render kernels are issue-bound, not HBM-bound (ncu: 60-79% issue-active, "
                        "<3% DRAM; their splat records stay L2-resident), so their HBM fraction is low by "
                        "construction; see roofline_by_kernel for the HBM-bound kernels"}
    roofline_by_kernel = {}
    for kname in per_launch:
        abk = algorithmic_bytes(kname, P, stats["R"], stats["R_fwd"], stats["R_bwd"], HW, stats["visible"])
        if abk:
            a_k = abk / (per_launch[kname] * 1e-3) / 1e9
            roofline_by_kernel[kname] = {
                "achieved_GBps": round(a_k, 1), "frac": round(a_k / peak, 4),
                "ms_per_launch": round(per_launch[kname], 4), "traffic": traffic_tbl.get(kname),
                "ncu_issue_active_pct": (ncu_stats.get(kname) or {}).get("issue_active_pct"),
                "ncu_dram_pct_of_peak": (ncu_stats.get(kname) or {}).get("dram_pct_of_peak")}
    total_bytes = P * 878 + stats["R"] * 200 + HW * 44     # SURVEY §8d whole-path figure
    stages = {"roofline_by_kernel": roofline_by_kernel,
              "ms_per_step_by_kernel": {k: round(v, 4) for k, v in sorted(per_step.items(), key=lambda kv: -kv[1])},
              "num_rendered": stats["R"], "visible": stats["visible"], "mean_tile_list": stats["mean_list"],
              "max_tile_list": stats["max_list"],
              "whole_path_frac_of_hbm_roofline": total_bytes / (ms_step * 1e-3) / 1e9 / peak}
    return roofline, stages


EXCHANGE_TEXT = {
    "nvlink": "own kernels over NVLink symmetric memory - the geometry backward pushes its colour gradients into every "
              "rank's table while it computes (multimem.st) and leaves packed [P,11] records; one more kernel sums the "
              "records in the switch (multimem.ld_reduce) and broadcasts them, rebuilds the SH rows and unpacks - two ranks: "
              "records pushed straight into the peer's inbox and summed locally; no NCCL collective on the data path",
    "factored": "NCCL all-gather of [P,3] colour gradients + all-reduce of [P,11] geometry gradients, SH rows rebuilt per rank",
    "allreduce": "1 NCCL all-reduce of [P,59] fp32 grads",
}


def run_ours(args):
    rank, world, local = _dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the rasterizer has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from splatfields_b200 import _lib, synth
    from splatfields_b200.host_api import ViewParallelRasterizer, forward_backward_host
    _lib.load()

    cfg, sc, cam, G = make_workload(rank)
    P, H, W = cfg["P"], cfg["H"], cfg["W"]
    vp = ViewParallelRasterizer(sc, cam, H, W, SH_DEGREE, device=dev, world_size=world)
    Gd = G.to(dev)

    def sync_all():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: `value` ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        vp.step(Gd)
    sync_all()
    if world > 1 and vp.exchange == "nvlink":
        # the NVLink exchange's device-side waits are bounded: if one ever gave up during warm-up (error word, on any
        # rank), measure the NCCL formulation of the same sum instead and say so in the line
        st = torch.tensor([vp.exchange_status()], device=dev, dtype=torch.int64)
        dist.all_reduce(st, op=dist.ReduceOp.MAX)
        if int(st.item()) != 0:
            reason = f"nvlink exchange reported error word {int(st.item()):#x} during warm-up"
            del vp
            vp = ViewParallelRasterizer(sc, cam, H, W, SH_DEGREE, device=dev, world_size=world, exchange="factored")
            vp.exchange_fallback_reason = reason
            for _ in range(max(args.warmup, 3)):
                vp.step(Gd)
            sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    if rank == 0:
        sampler.mark()
    e0.record()
    launches = 0
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        launches += vp.step(Gd)
    e1.record()
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps   # CPU time to enqueue one step
    sync_all()
    ms_total = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if dist is not None:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = float(ms_total.item()) / args.steps
    value = world * P / (ms_step * 1e-3) / 1e6

    # ---- end-to-end through the host-buffer API: `e2e` ----
    # HostPipeline.submit/wait: every step uploads the job's inputs (all splat parameters + this rank's cotangent) from
    # pinned host memory and downloads image, depth, radii and the per-splat gradients into pinned host memory; upload
    # of step k+1, kernels of step k and download of step k-1 overlap on three streams.  With N ranks the job's ONE host
    # copy of the splats is split by rows over the ranks (each uploads 1/N, NCCL all-gathers the rest over NVLink) and
    # each rank downloads 1/N of the summed gradient slab.  Wall clock, max over ranks.
    from splatfields_b200.host_api import HostPipeline

    def time_pipeline(pipe, host_in, outs):
        for k in range(3):
            pipe.submit(host_in, outs[k % 2])
        pipe.drain()
        sync_all()
        t0 = time.perf_counter()
        first = pipe.n
        for k in range(args.steps):
            t = pipe.submit(host_in, outs[k % 2])
            if t > first:
                pipe.wait(t - 1)
        pipe.wait(pipe.n - 1)
        pipe.drain()
        wall = time.perf_counter() - t0
        sync_all()
        ms = torch.tensor([wall * 1e3], device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / args.steps

    slab0, params0 = vp.slab, {k: p.data for k, p in vp.params.items()}
    factored_host = world == 1       # single rank: SH gradient downloaded in factored form (56 instead of 236 B/splat)
    host_in, host_out = vp.pinned_host_buffers(sc, G, shard=world > 1, factored=factored_host)
    host_out2 = {k: torch.empty_like(v).pin_memory() for k, v in host_out.items()}
    pipe = HostPipeline(vp, factored=factored_host)
    ms_e2e_step = time_pipeline(pipe, host_in, (host_out, host_out2))
    h2d, d2h = pipe.h2d_bytes(host_in), pipe.d2h_bytes(host_out)
    vp.factored_output = None
    e2e = {"value": world * P / (ms_e2e_step * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_e2e_step,
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "api": "splatfields_b200.host_api.HostPipeline.submit/wait (pinned host buffers on the GPU's NUMA node, "
                  "3 streams, depth 2)",
           "gradient_format": ("factored SH gradient: [P,11] geometry + [P,3] colour gradients "
                               "(host_api.sh_rows_from_factored rebuilds [P,16,3] rows on demand)") if factored_host
           else f"full [P,59] slab, rows sharded over the {world} ranks (each rank moves 1/{world} of the splats "
                f"through its PCIe link; the parameters are all-gathered over NVLink)",
           "bytes_note": "per rank"}
    if world == 1:
        # context: the same pipeline with the full [P,16,3] SH gradient rows downloaded, and the fully synchronous call
        try:
            del pipe, host_out, host_out2
            host_in_f, host_out_f = vp.pinned_host_buffers(sc, G, shard=False, factored=False)
            host_out_f2 = {k: torch.empty_like(v).pin_memory() for k, v in host_out_f.items()}
            pipe_f = HostPipeline(vp, factored=False)
            ms_full = time_pipeline(pipe_f, host_in_f, (host_out_f, host_out_f2))
            e2e["full_rows"] = {"value": P / (ms_full * 1e-3) / 1e6, "ms_per_step": ms_full,
                                "d2h_bytes_per_step": int(pipe_f.d2h_bytes(host_out_f))}
            forward_backward_host(vp, host_in_f, host_out_f)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            nsync = max(3, min(args.steps, 10))
            for _ in range(nsync):
                forward_backward_host(vp, host_in_f, host_out_f)
            e2e["synchronous_ms_per_step"] = (time.perf_counter() - t0) * 1e3 / nsync
            del pipe_f, host_out_f, host_out_f2
        except Exception as ex:      # context numbers only: never lose the bench line over them
            e2e["context_error"] = f"{type(ex).__name__}: {ex}"
    vp.slab = slab0
    for k, p_ in vp.params.items():
        p_.data = params0[k]
    sync_all()
    clocks = sampler.stop() if rank == 0 else None     # samples span both timed regions (value + e2e)

    # ---- roofline of the dominant kernel (profiled steps, outside the timed regions) ----
    roofline, stages = None, None
    acc = {}
    nprof = 5
    # every rank runs the profiled steps (vp.step contains the all-reduce: collectives must stay symmetric);
    # only rank 0 records and reads the per-kernel events.
    if rank == 0:
        _lib.profile_enable(True)
    vp.time_exchange = True
    for _ in range(nprof):
        vp.step(Gd)
        torch.cuda.synchronize()
        if rank == 0:
            for which in (0, 1):
                for name, ms in _lib.profile_read(which):
                    acc.setdefault(name, []).append(ms)
    vp.time_exchange = False
    exch_ms = vp.exchange_ms()
    exch_timeline = vp.exchange_timeline() if (world > 1 and vp.exchange == "nvlink") else None
    if rank == 0:
        _lib.profile_enable(False)
    sync_all()
    if rank == 0:
        stats = vp.list_stats()
        roofline, stages = roofline_report(acc, nprof, stats, P, H * W, ms_step)

    # ---- exchange check: this exchange against the plain NCCL all-reduce of the whole slab, same views ----
    exchange_check = None
    if world > 1 and vp.exchange != "allreduce":
        vp.step(Gd)
        mine = {k: v.clone() for k, v in vp.grads().items()}
        vref = ViewParallelRasterizer(sc, cam, H, W, SH_DEGREE, device=dev, world_size=world, exchange="allreduce")
        vref.step(Gd)
        worst = 0.0
        for k, v in vref.grads().items():
            worst = max(worst, float((mine[k] - v).abs().max()) / max(float(v.abs().max()), 1e-30))
        wt = torch.tensor([worst], device=dev)
        dist.all_reduce(wt, op=dist.ReduceOp.MAX)
        exchange_check = {"max_abs_err_over_max_abs_vs_nccl_allreduce": float(wt.item()), "ok": bool(wt.item() < 2e-5)}
        del vref, mine

    # ---- CPU baseline: the oracle port on the host cores, bounded sample (N = 1 only) ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        O.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
        t, _f = oracle_step(O, sc, cam, cfg, G)
        cpu_baseline = {"value": P / t / 1e6, "unit": UNIT, "cores": O.num_threads(), "kind": "port",
                        "sample": f"1 full fwd+bwd step of {WORKLOAD} ({t:.2f} s of wall time on all host threads)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "P": P, "H": H, "W": W, "sh_degree": SH_DEGREE,
                       "parallelism": (f"view-parallel x{world} (one camera per GPU, splats replicated; gradient "
                                       f"exchange: " + EXCHANGE_TEXT[vp.exchange] + ")")
                       if world > 1 else "single view",
                       "l2": "inputs larger than L2 (236 MB of splat parameters + 0.5 GB of scratch per step "
                             "vs 126 MB L2)"},
            "views_per_s": world / (ms_step * 1e-3), "host_enqueue_ms_per_step": host_enqueue_ms,
            "exchange": None if world == 1 else {
                "mode": vp.exchange, "multicast": getattr(vp, "xchg_multicast", None),
                "fallback_reason": vp.exchange_fallback_reason,
                "fused_with_geometry_backward": bool(getattr(vp, "xchg_fused", False)) if vp.exchange == "nvlink" else False,
                "ms_per_step": (float(np.mean(exch_ms)) if exch_ms else None)
                if not (vp.exchange == "nvlink" and getattr(vp, "xchg_fused", False)) else None,
                "device_timeline_us_rank0": exch_timeline,
                "fused_kernel_ms_per_step": (stages or {}).get("ms_per_step_by_kernel", {}).get("geom_backward_exchange"),
                "nvlink_bytes_received_per_splat_per_rank": vp.bytes_on_wire_per_splat(),
                "nvlink_MB_received_per_rank_per_step": vp.bytes_on_wire_per_splat() * P / 1e6,
                "exchange_check": exchange_check,
                "note": "ms_per_step: device time from the end of the backward to the end of the exchange on rank 0 "
                        "(profiled steps); fused_kernel_ms_per_step: the kernel that does the geometry backward AND the "
                        "exchange (the plain geometry backward takes ~0.11 ms at N=1)"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "stages": stages,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def _watchdog(seconds):
    """Hard stop: a hung collective must never hold a multi-GPU box for the whole lease."""
    import signal

    def _die(signum, frame):
        sys.stderr.write(f"bench.py: watchdog fired after {seconds}s\n")
        sys.stderr.flush()
        os._exit(3)
    signal.signal(signal.SIGALRM, _die)
    signal.alarm(seconds)


def main():
    _watchdog(int(os.environ.get("BENCH_WATCHDOG_S", "900")))
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
