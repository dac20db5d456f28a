"""BASELINE.json configs[4] ("owlii_2m": 2 M Gaussians, 1080x1920, precomputed RGB, 300 frames x 6 views, (view, time)
jobs sharded round-robin over the ranks): throughput of the rasterizer path over a bounded number of rounds.
torchrun (or a single process) on a GPU box; rank 0 prints one JSON line.  The per-frame displacement
synth.frame_offset stands in for the reference's deformation network (out of scope)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from splatfields_b200 import synth
from splatfields_b200.host_api import ViewParallelRasterizer


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=24)
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--views", type=int, default=6)
    ap.add_argument("--exchange", default=None, help="nvlink (default when every rank has a job in every round) | allreduce")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    cfg = synth.CONFIGS["owlii_2m"]
    P, H, W = cfg["P"], cfg["H"], cfg["W"]
    sc = synth.make_scene(P, cfg["seed"], scale_mult=cfg["scale_mult"], precomp_rgb=cfg["precomp_rgb"])
    base = sc["means3D"].to(dev)
    jobs = synth.view_time_jobs(a.frames, a.views, rank, world)[: a.rounds + 3]
    # ranks render different time steps: the [P, 14] parameter gradients are summed as they are (no SH factorisation);
    # the NVLink exchange needs a job on every rank in every round (no idle contributions)
    exchange = a.exchange or ("nvlink" if (a.frames * a.views) % world == 0 else "allreduce")
    vp = ViewParallelRasterizer(sc, synth.config_camera("owlii_2m", 0), H, W, 0, device=dev, world_size=world,
                                exchange=exchange)
    G = torch.randn(3, H, W, generator=torch.Generator().manual_seed(rank)).to(dev)
    offsets = {f: synth.frame_offset(P, f / a.frames, cfg["seed"]).to(dev) for f in {j[0] for j in jobs if j}}

    cams = {v: synth.config_camera("owlii_2m", v).to(dev) for v in range(a.views)}    # device-resident: no per-job H2D

    def run(job):
        if job is None:
            return vp.idle_step()
        f, v = job
        vp.set_camera(cams[v])
        vp.set_means(base + offsets[f])
        return vp.step(G)

    for job in jobs[:3]:
        run(job)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    done = 0
    for job in jobs[3:]:
        run(job)
        done += job is not None
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    njobs = torch.tensor([done], device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(njobs)
    if rank == 0:
        t = float(ms.item()) * 1e-3
        print(json.dumps({"config": "owlii_2m view x time", "world": world, "exchange": vp.exchange,
                          "multicast": getattr(vp, "xchg_multicast", None), "jobs": int(njobs.item()),
                          "rounds": len(jobs) - 3, "ms_per_round": t * 1e3 / max(len(jobs) - 3, 1),
                          "jobs_per_s": float(njobs.item()) / t, "msplats_s": float(njobs.item()) * P / t / 1e6}),
              flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
