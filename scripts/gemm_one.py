"""One plain ROWS launch at the bench shape, repeated (for ncu counter A/B runs; tuning aid)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from competesmoe_b200 import ops
dev = torch.device("cuda")
T, K, E, D, F = 4096, 2, 4, 3072, 8192
g = torch.Generator().manual_seed(1)
sel = torch.stack([torch.randperm(E, generator=g)[:K] for _ in range(T)]).int().to(dev)
route = ops.route_build(sel, E)
bf = dict(device=dev, dtype=torch.bfloat16)
xp = torch.randn(route.row_cap, D, **bf)
w1 = torch.randn(E, 2 * F, D, **bf) * 0.02
for _ in range(4):
    ops.gemm_rows(xp, w1, w_is_kn=False, route=route)
torch.cuda.synchronize()
