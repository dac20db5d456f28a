"""Multi-GPU check (torchrun, NCCL): the factored gradient exchange (all-gather of colour gradients + all-reduce of
the geometry gradients + sfb_sh_grad_combine) against the plain all-reduce of the whole [59, P] slab, on the same
views.  Prints one JSON line from rank 0; exit code 1 on mismatch."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from splatfields_b200 import synth
from splatfields_b200.host_api import ViewParallelRasterizer


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    P, H, W, deg = 200_000, 400, 400, 3
    sc = synth.make_scene(P, 9, scale_mult=1.5)
    sc["shs"][::7, 0, 1] = -3.0           # exercise the colour clamp
    cam = synth.orbit_camera(rank, H, W)
    G = torch.randn(3, H, W, generator=torch.Generator().manual_seed(100 + rank)).to(dev)
    out = {}
    for mode in ("allreduce", "factored"):
        vp = ViewParallelRasterizer(sc, cam, H, W, deg, device=dev, world_size=world, exchange=mode)
        assert vp.exchange == mode
        vp.step(G)
        vp.step(G)                         # twice: buffers are reused across steps
        torch.cuda.synchronize()
        out[mode] = {k: v.clone() for k, v in vp.grads().items()}
    worst = {}
    ok = True
    for k in out["allreduce"]:
        a, b = out["factored"][k], out["allreduce"][k]
        scale = float(b.abs().max())
        err = float((a - b).abs().max())
        worst[k] = err / max(scale, 1e-30)
        ok = ok and scale > 0 and worst[k] < 2e-5
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print(json.dumps({"check": "factored exchange == all-reduce of the whole slab", "world": world, "P": P,
                          "max_err_over_max_abs": worst, "ok": bool(flag.item() == 0)}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
