"""Times sigma_wgrad at the C4 shape (CSMOE_SIGMA_DBG selects which part of the kernel is switched off)."""
import os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from competesmoe_b200 import ops  # noqa: E402
dev = torch.device("cuda")
T, D, E, K, H = 8192, 1024, 64, 8, 128
g = torch.Generator().manual_seed(0)
x = torch.randn(T, D, generator=g).bfloat16().to(dev)
sel = torch.stack([torch.randperm(E, generator=g)[:K] for _ in range(T)]).int().to(dev)
route = ops.route_build(sel, E, row_tile=128)
dz = torch.randn(route.row_cap, H, generator=g).bfloat16().to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, n=10):
    for _ in range(2): fn()
    tot = 0.0
    for _ in range(n):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / n * 1e3
print(f"DBG={os.environ.get('CSMOE_SIGMA_DBG', '0'):>3s}  wgrad transpose {timeit(lambda: ops.sigma_wgrad(dz, x, E, route, True)):7.1f} us   direct {timeit(lambda: ops.sigma_wgrad(dz, x, E, route, False)):7.1f} us")
