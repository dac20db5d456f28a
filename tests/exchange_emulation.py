"""Two (or more) ranks of the NVLink gradient exchange emulated on ONE GPU, in ONE process: every "rank" owns a buffer
of sfb_xchg_bytes() bytes on the same device, the peer table points at the other buffers, and the ranks'
sfb_xchg_finish kernels run concurrently on separate streams (each limited to its share of the SMs, so that all of
them are resident while they wait for each other's flags).  Exercises the unicast (non-multicast) protocol of
csrc/exchange.cu and geom_backward_kernel<PUSH> end to end; the multicast path needs NVSwitch peers
(scripts/check_exchange.py on >= 2 GPUs).

Run as a script by tests/test_gpu_exchange.py (a process of its own: a protocol bug traps the kernel, which poisons the
CUDA context).  Prints one JSON line; exit code 1 on mismatch."""
import ctypes as C
import json
import math
import os
import sys

# every stream gets a hardware queue of its own: the ranks' kernels wait for each other on the device, so they must
# never be serialised behind one another by queue aliasing (default: 8 connections shared by all streams)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from splatfields_b200 import _lib, rasterizer, synth
from splatfields_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer


def settings(cam, H, W, deg, dev):
    camd = cam.to(dev)
    return GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5),
        bg=torch.ones(3, device=dev), scale_modifier=1.0, viewmatrix=camd.world_view_transform,
        projmatrix=camd.full_proj_transform, sh_degree=deg, campos=camd.camera_center, prefiltered=False, debug=False)


def run_fused(world, P, H, W, deg, precomp_rgb, steps=3, M=None):
    """The fused backward + exchange kernel (one persistent kernel per rank and step): every emulated rank runs its
    forward and backward on a stream of its own, so that the ranks' kernels are resident side by side."""
    dev = torch.device("cuda:0")
    lib = _lib.load()
    sc = synth.make_scene(P, 5, scale_mult=2.5, precomp_rgb=precomp_rgb)
    if not precomp_rgb:
        sc["shs"][::7, 0, 1] = -3.0          # exercise the colour clamp
        if M is not None:
            sc["shs"] = sc["shs"][:, :M].contiguous()
    t = {k: v.to(dev) for k, v in sc.items()}
    cams = [synth.orbit_camera(r, H, W) for r in range(world)]
    ngeo = 16 if precomp_rgb else 12
    nbytes = int(lib.sfb_xchg_bytes(P, world, ngeo, int(not precomp_rgb)))
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=dev) for _ in range(world)]
    campos = torch.stack([c.camera_center for c in cams]).to(dev).contiguous()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    descs = []
    for r in range(world):
        d = _lib.XchgDesc()
        d.rank, d.world, d.P, d.ngeo = r, world, P, ngeo
        d.local = bufs[r].data_ptr()
        for q in range(world):
            d.peers[q] = bufs[q].data_ptr()
        d.mc = None
        d.max_ctas = max(8, sms // world)          # leave room for the other ranks' render kernels
        d.campos_views = None if precomp_rgb else campos.data_ptr()
        descs.append(d)
    names = ["means3D", "opacities", "scales", "rotations"] + (["colors_precomp"] if precomp_rgb else ["shs"])
    fields = tuple((n, t[n][0].numel()) for n in names)
    fps = sum(n for _, n in fields)
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    worst = {}
    ok = True

    def forward(r, leaf, m2d):
        return GaussianRasterizer(settings(cams[r], H, W, deg, dev))(
            means3D=leaf["means3D"], means2D=m2d, opacities=leaf["opacities"], shs=leaf.get("shs"),
            colors_precomp=leaf.get("colors_precomp"), scales=leaf["scales"], rotations=leaf["rotations"])

    for step in range(1, steps + 1):
        Gs = [torch.randn(3, H, W, generator=torch.Generator().manual_seed(100 * step + r)).to(dev) for r in range(world)]
        torch.cuda.synchronize()
        # reference: plain backward per view, on the rank's own stream (this also fills the stream's allocator pool: no
        # cudaMalloc may happen later while another rank's kernel is waiting on the device)
        refs = []
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                leaf = {k: v.clone().requires_grad_(True) for k, v in t.items()}
                m2d = torch.zeros(P, 3, device=dev, requires_grad=True)
                color, _, _ = forward(r, leaf, m2d)
                color.backward(Gs[r])
                refs.append({n: leaf[n].grad for n in names})
        torch.cuda.synchronize()
        ref = {n: sum(refs[r][n] for r in range(world)) if world != 2 else refs[0][n] + refs[1][n] for n in names}
        del refs
        # the ranks' forwards (each synchronises the host once), then every backward without host synchronisation
        state = []
        slabs = [torch.full((fps * P,), float("nan"), device=dev) for _ in range(world)]
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                leaf = {k: v.clone().requires_grad_(True) for k, v in t.items()}
                m2d = torch.zeros(P, 3, device=dev, requires_grad=True)
                color, _, _ = forward(r, leaf, m2d)
                state.append((leaf, m2d, color))
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                rasterizer.set_grad_arena(slabs[r], fields, None, (descs[r], step, True))
                try:
                    state[r][2].backward(Gs[r])
                finally:
                    rasterizer.set_grad_arena(None, None)
        torch.cuda.synchronize()
        for r in range(world):
            st = C.c_uint(0)
            _lib.check(lib.sfb_xchg_status(C.byref(descs[r]), C.byref(st), None))
            if st.value:
                print(f"fused, rank {r}: exchange error word {st.value:#x} (world {world}, step {step})", flush=True)
                return False, {"status": st.value}
            assert state[r][0]["means3D"].grad is None and state[r][1].grad is not None
        off = 0
        for n, nf in fields:
            scale = float(ref[n].abs().max())
            for r in range(world):
                got = slabs[r][off * P:(off + nf) * P].view(ref[n].shape)
                err = float((got - ref[n]).abs().max()) / max(scale, 1e-30)
                worst[n] = max(worst.get(n, 0.0), err)
                same = bool(torch.equal(got, slabs[0][off * P:(off + nf) * P].view(ref[n].shape)))
                ok = ok and scale > 0 and err < 2e-5 and same
            off += nf
    return ok, worst


def run(world, P, H, W, deg, precomp_rgb, steps=3):
    dev = torch.device("cuda:0")
    lib = _lib.load()
    sc = synth.make_scene(P, 5, scale_mult=2.5, precomp_rgb=precomp_rgb)
    if not precomp_rgb:
        sc["shs"][::7, 0, 1] = -3.0          # exercise the colour clamp
    t = {k: v.to(dev) for k, v in sc.items()}
    cams = [synth.orbit_camera(r, H, W) for r in range(world)]
    ngeo = 16 if precomp_rgb else 12
    nbytes = int(lib.sfb_xchg_bytes(P, world, ngeo, int(not precomp_rgb)))
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=dev) for _ in range(world)]
    descs = []
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    for r in range(world):
        d = _lib.XchgDesc()
        d.rank, d.world, d.P, d.ngeo = r, world, P, ngeo
        d.local = bufs[r].data_ptr()
        for q in range(world):
            d.peers[q] = bufs[q].data_ptr()
        d.mc = None
        d.max_ctas = (2 * sms) // world
        descs.append(d)
    campos = torch.stack([c.camera_center for c in cams]).to(dev).contiguous()
    M = 0 if precomp_rgb else t["shs"].shape[1]
    names = ["means3D", "opacities", "scales", "rotations"] + (["colors_precomp"] if precomp_rgb else ["shs"])
    worst = {}
    ok = True
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    for step in range(1, steps + 1):
        Gs = [torch.randn(3, H, W, generator=torch.Generator().manual_seed(100 * step + r)).to(dev) for r in range(world)]
        # reference: plain backward per view, summed in view order
        ref = {n: torch.zeros_like(t[n]) for n in names}
        for r in range(world):
            leaf = {k: v.clone().requires_grad_(True) for k, v in t.items()}
            m2d = torch.zeros(P, 3, device=dev, requires_grad=True)
            color, _, _ = GaussianRasterizer(settings(cams[r], H, W, deg, dev))(
                means3D=leaf["means3D"], means2D=m2d, opacities=leaf["opacities"], shs=leaf.get("shs"),
                colors_precomp=leaf.get("colors_precomp"), scales=leaf["scales"], rotations=leaf["rotations"])
            color.backward(Gs[r])
            for n in names:
                ref[n] += leaf[n].grad
        # exchange mode: every rank's backward (packed records + pushed colour gradients) ...
        for r in range(world):
            leaf = {k: v.clone().requires_grad_(True) for k, v in t.items()}
            m2d = torch.zeros(P, 3, device=dev, requires_grad=True)
            color, _, _ = GaussianRasterizer(settings(cams[r], H, W, deg, dev))(
                means3D=leaf["means3D"], means2D=m2d, opacities=leaf["opacities"], shs=leaf.get("shs"),
                colors_precomp=leaf.get("colors_precomp"), scales=leaf["scales"], rotations=leaf["rotations"])
            rasterizer.set_grad_arena(None, (), None, (descs[r], step))
            try:
                color.backward(Gs[r])
            finally:
                rasterizer.set_grad_arena(None, None)
            assert leaf["means3D"].grad is None and m2d.grad is not None
        torch.cuda.synchronize()
        # ... then the ranks' finish kernels side by side
        # (all output tensors exist before the first finish kernel starts: a cudaMalloc issued while rank 0's kernel is
        #  already spinning on rank 1's flag can wait for the device to drain, i.e. for the very kernel that waits for us)
        outs = [{n: torch.full_like(t[n], float("nan")) for n in names} for _ in range(world)]
        torch.cuda.synchronize()
        for r in range(world):
            o = outs[r]
            gp = lambda n: o[n].data_ptr() if n in o else None
            with torch.cuda.stream(streams[r]):
                _lib.check(lib.sfb_xchg_finish(
                    C.byref(descs[r]), step, deg, int(M), t["means3D"].data_ptr(),
                    None if precomp_rgb else campos.data_ptr(), gp("means3D"), gp("opacities"), gp("scales"),
                    gp("rotations"), gp("colors_precomp"), gp("shs"), streams[r].cuda_stream))
        torch.cuda.synchronize()
        for r in range(world):
            st = C.c_uint(0)
            _lib.check(lib.sfb_xchg_status(C.byref(descs[r]), C.byref(st), None))
            if st.value:
                flags = [bufs[q][:256].view(torch.int32).tolist() for q in range(world)]
                print(f"rank {r}: exchange error word {st.value:#x} (world {world}, step {step}); flag words "
                      f"A/B/done/ticket/err per rank: " + str([(f[0:world], f[16:16 + world], f[32:35]) for f in flags]),
                      flush=True)
                return False, {"status": st.value}
        for n in names:
            scale = float(ref[n].abs().max())
            for r in range(world):
                err = float((outs[r][n] - ref[n]).abs().max()) / max(scale, 1e-30)
                worst[n] = max(worst.get(n, 0.0), err)
                ok = ok and scale > 0 and err < 2e-5 and bool(torch.equal(outs[r][n], outs[0][n]))
    return ok, worst


def main():
    res = []
    allok = True
    for world, P, H, W, deg, rgb in ((2, 20_000, 160, 208, 3, False), (4, 9_001, 96, 128, 2, False),
                                     (2, 12_000, 128, 128, 0, True), (3, 7_777, 96, 96, 1, False)):
        ok, worst = run(world, P, H, W, deg, rgb)
        res.append({"world": world, "P": P, "deg": deg, "precomp_rgb": rgb, "ok": ok, "max_err_over_max_abs": worst})
        allok = allok and ok
    # the fused kernel: two ranks (direct sums), three and four ranks (owner reduces and broadcasts), ragged chunk
    # counts, an active degree below the allocated coefficients, precomputed colours
    for world, P, H, W, deg, rgb, M in ((2, 20_000, 160, 208, 3, False, None), (4, 9_001, 96, 128, 2, False, None),
                                        (2, 12_000, 128, 128, 0, True, None), (3, 7_777, 96, 96, 1, False, None),
                                        (3, 5_000, 96, 96, 0, True, None), (2, 3_000, 64, 64, 0, False, 16)):
        ok, worst = run_fused(world, P, H, W, deg, rgb, M=M)
        res.append({"fused": True, "world": world, "P": P, "deg": deg, "precomp_rgb": rgb, "ok": ok,
                    "max_err_over_max_abs": worst})
        allok = allok and ok
    print(json.dumps({"check": "emulated ranks on one GPU: NVLink exchange == sum of the views' plain backward passes",
                      "cases": res, "ok": allok}), flush=True)
    sys.exit(0 if allok else 1)


if __name__ == "__main__":
    main()
