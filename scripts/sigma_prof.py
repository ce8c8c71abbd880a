"""Runs the fused sigma-MoE kernels once each at the C4 shape (for ncu): fwd, bwd, wgrad x2, combine, scatter_reduce."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from competesmoe_b200 import ops  # noqa: E402
dev = torch.device("cuda")
T, D, Dout, E, K, H = 8192, 1024, 1024, 64, 8, 128
g = torch.Generator().manual_seed(0)
x = torch.randn(T, D, generator=g).bfloat16().to(dev)
dout = torch.randn(T, Dout, generator=g).bfloat16().to(dev)
keys = (torch.randn(E, D, H, generator=g) * D ** -0.5).bfloat16().to(dev)
values = (torch.randn(E, H, Dout, generator=g) * H ** -0.5).bfloat16().to(dev)
sel = torch.stack([torch.randperm(E, generator=g)[:K] for _ in range(T)]).int().to(dev)
w = torch.rand(T, K, generator=g).to(dev) + 0.1
route = ops.route_build(sel, E, row_tile=128)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    y, h = ops.sigma_ffn_fwd(x, keys, values, None, route)
    dz, hw, dxr, dwp = ops.sigma_ffn_bwd(dout, keys, values, route, w, h)
    dv = ops.sigma_wgrad(hw, dout, E, route, transpose=False)
    dk = ops.sigma_wgrad(dz, x, E, route, transpose=True)
    out = ops.combine_fwd(y, route.slot_to_row, route.sel, w, T, K, round_w=True)
    dx = ops.scatter_reduce(dxr, route.slot_to_row, T, K)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()), float(dx.float().abs().mean()), float(dv.abs().mean()), float(dk.abs().mean()))
