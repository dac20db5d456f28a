"""Multi-GPU measurement (torchrun): where the time of the NVLink exchange goes.  For a few settings of the tuning
hooks (share of CTAs that start on the slice reduction, round trips in flight per thread) runs lego_1m steps and reads
the device timeline of sfb_xchg_finish (sfb_xchg_timeline) on every rank.  Rank 0 prints one JSON line per setting:
per-phase durations in microseconds, median over the steps, max over the ranks."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from splatfields_b200 import _lib, synth
from splatfields_b200.host_api import ViewParallelRasterizer


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    cfg = synth.CONFIGS["lego_1m"]
    sc = synth.make_scene(cfg["P"], cfg["seed"], scale_mult=cfg["scale_mult"], precomp_rgb=cfg["precomp_rgb"])
    cam = synth.config_camera("lego_1m", rank % 8)
    G = torch.randn(3, cfg["H"], cfg["W"], generator=torch.Generator().manual_seed(1000 + rank)).to(dev)
    vp = ViewParallelRasterizer(sc, cam, cfg["H"], cfg["W"], 3, device=dev, world_size=world, exchange="nvlink")
    assert vp.exchange == "nvlink", vp.exchange_fallback_reason
    names = ["barrier_a", "reduce", "sh_rows_after_reduce", "barrier_b", "unpack", "total"]
    for nred, depth in ((4, 4), (4, 16), (2, 16), (1, 16), (8, 16), (2, 4)):
        lib.sfb_xchg_tune(nred, depth)
        rows = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(24):
            if it == 4:
                torch.cuda.synchronize(); dist.barrier(); e0.record()
            vp.step(G)
            if it >= 4 and it % 4 == 0:       # reading the timeline synchronises: only some of the steps
                t = (C.c_ulonglong * 12)()
                _lib.check(lib.sfb_xchg_timeline(C.byref(vp.xchg), t, torch.cuda.current_stream(dev).cuda_stream))
                t = [int(v) for v in t]
                rows.append([(t[1] - t[0]) / 1e3, (t[2] - t[1]) / 1e3, (t[3] - t[2]) / 1e3, (t[4] - max(t[2], t[3])) / 1e3,
                             (t[5] - t[4]) / 1e3, (t[5] - t[0]) / 1e3])
        e1.record(); torch.cuda.synchronize()
        med = torch.tensor(np.median(np.array(rows), axis=0), device=dev)
        allr = [torch.empty_like(med) for _ in range(world)]
        dist.all_gather(allr, med)
        ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            per_rank = torch.stack(allr).cpu().numpy()
            print(json.dumps({"world": world, "multicast": vp.xchg_multicast, "reduce_ctas_eighths": nred, "in_flight": depth,
                              "ms_per_step_with_timeline_reads": float(ms.item()),
                              "us_max_over_ranks": dict(zip(names, [round(float(v), 1) for v in per_rank.max(axis=0)])),
                              "us_min_over_ranks": dict(zip(names, [round(float(v), 1) for v in per_rank.min(axis=0)]))}),
                  flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
