"""Which fp32 formula does torch.norm(g[:, :2], dim=-1) use on the GPU?  (bit-level; decides densify.cu's rounding)"""
import torch
dev = torch.device("cuda:0")
g = torch.randn(1_000_000, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(0)) * 3e-4
ref = torch.norm(g[:, :2], dim=-1)
x, y = g[:, 0].double(), g[:, 1].double()
f32 = lambda t: t.float()
cands = {
    "sqrt(fl(x*x) + fl(y*y))": torch.sqrt(f32(f32(x * x).double() + f32(y * y).double())),
    "sqrt(fma(y,y,fl(x*x)))": torch.sqrt(f32(f32(x * x).double() + y * y)),
    "sqrt(fma(x,x,fl(y*y)))": torch.sqrt(f32(f32(y * y).double() + x * x)),
    "sqrt(fl(x*x + y*y)) exact sum": torch.sqrt(f32(x * x + y * y)),
    "fl(sqrt_f64(x*x + y*y))": f32(torch.sqrt(x * x + y * y)),
    "hypot": torch.hypot(g[:, 0], g[:, 1]),
}
for k, v in cands.items():
    print(f"{k:36s} mismatches: {(v != ref).sum().item()}")
m = torch.norm(g[:, :2], dim=-1, keepdim=True)
print("keepdim same:", torch.equal(m.squeeze(-1), ref))
mask = torch.rand(1_000_000, device=dev) > 0.3
print("masked-gather same:", torch.equal(torch.norm(g[mask, :2], dim=-1), ref[mask]))
