"""Multi-GPU check (torchrun): the three gradient exchanges of host_api.ViewParallelRasterizer on the same views —
"nvlink" (the library's own kernels over symmetric memory: multimem through the switch, and again with the multicast
mapping withheld = peer loads / stores), "factored" (NCCL all-gather + all-reduce + sfb_sh_grad_combine) — against the
plain NCCL all-reduce of the whole [59, P] slab.  Also a precomputed-colour scene ([14, P] records, nvlink vs
all-reduce).  Prints one JSON line from rank 0; exit code 1 on mismatch."""
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
    info = {}
    # nvlink: backward (records + pushed colour gradients), then sfb_xchg_finish; _p2p: the same without the multicast
    # mapping; _fused: the one-kernel variant (geometry backward + exchange); _fused_p2p: that without multicast
    for mode in ("allreduce", "factored", "nvlink", "nvlink_p2p", "nvlink_fused", "nvlink_fused_p2p"):
        os.environ["SFB_XCHG_NO_MULTICAST"] = "1" if mode.endswith("_p2p") else "0"
        os.environ["SFB_XCHG_FUSED"] = "1" if "_fused" in mode else "0"
        vp = ViewParallelRasterizer(sc, cam, H, W, deg, device=dev, world_size=world, exchange=mode.split("_")[0])
        assert vp.exchange == mode.split("_")[0]
        if vp.exchange == "nvlink":
            info[mode] = {"multicast": vp.xchg_multicast, "fused": vp.xchg_fused}
        for _ in range(3):                 # several steps: buffers, flags and parities are reused across steps
            vp.step(G)
        torch.cuda.synchronize()
        out[mode] = {k: v.clone() for k, v in vp.grads().items()}
        del vp
    # precomputed colours: [14, P] records, plain sum
    sc2 = synth.make_scene(P, 10, scale_mult=1.5, precomp_rgb=True)
    for mode in ("allreduce", "nvlink", "nvlink_fused"):
        os.environ["SFB_XCHG_NO_MULTICAST"] = "0"
        os.environ["SFB_XCHG_FUSED"] = "1" if "_fused" in mode else "0"
        vp = ViewParallelRasterizer(sc2, cam, H, W, 0, device=dev, world_size=world, exchange=mode.split("_")[0])
        for _ in range(2):
            vp.step(G)
        torch.cuda.synchronize()
        out["rgb_" + mode] = {k: v.clone() for k, v in vp.grads().items()}
        del vp
    worst = {}
    ok = True
    os.environ["SFB_XCHG_FUSED"] = "0"
    for mode, base in (("factored", "allreduce"), ("nvlink", "allreduce"), ("nvlink_p2p", "allreduce"),
                       ("nvlink_fused", "allreduce"), ("nvlink_fused_p2p", "allreduce"),
                       ("rgb_nvlink", "rgb_allreduce"), ("rgb_nvlink_fused", "rgb_allreduce")):
        for k in out[base]:
            a, b = out[mode][k], out[base][k]
            scale = float(b.abs().max())
            err = float((a - b).abs().max())
            worst[mode + "." + k] = err / max(scale, 1e-30)
            ok = ok and scale > 0 and worst[mode + "." + k] < 2e-5
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print(json.dumps({"check": "nvlink (multicast / peer) and factored exchanges == NCCL all-reduce of the whole slab",
                          "world": world, "P": P, "nvlink": info, "max_err_over_max_abs": worst,
                          "ok": bool(flag.item() == 0)}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
