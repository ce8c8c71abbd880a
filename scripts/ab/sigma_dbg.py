"""Times the fused sigma kernels at the C4 shape (CSMOE_SIGMA_DBG selects which part of the kernels is switched off)."""
import os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from competesmoe_b200 import ops  # noqa: E402
dev = torch.device("cuda")
T, D, E, K, H = 8192, 1024, 64, 8, 128
g = torch.Generator().manual_seed(0)
x = torch.randn(T, D, generator=g).bfloat16().to(dev)
keys = (torch.randn(E, D, H, generator=g) * D ** -0.5).bfloat16().to(dev)
values = (torch.randn(E, H, D, generator=g) * H ** -0.5).bfloat16().to(dev)
sel = torch.stack([torch.randperm(E, generator=g)[:K] for _ in range(T)]).int().to(dev)
w = torch.rand(T, K, generator=g).to(dev)
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
print(f"DBG={os.environ.get('CSMOE_SIGMA_DBG', '0'):>3s}  fwd {timeit(lambda: ops.sigma_ffn_fwd(x, keys, values, None, route)):7.1f} us  "
      f"bwd {timeit(lambda: ops.sigma_ffn_bwd(x, keys, values, route, w, dz)):7.1f} us  "
      f"wgrad(T) {timeit(lambda: ops.sigma_wgrad(dz, x, E, route, True)):7.1f} us  wgrad {timeit(lambda: ops.sigma_wgrad(dz, x, E, route, False)):7.1f} us")

if os.environ.get("CSMOE_SIGMA_STATS"):
    from competesmoe_b200 import _lib
    st = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
    _lib.load().csmoe_sigma_set_stats(st.data_ptr())
    ops.sigma_wgrad(dz, x, E, route, True)
    torch.cuda.synchronize()
    _lib.load().csmoe_sigma_set_stats(None)
    v = st.view(148, 8).double().mean(0).tolist()
    print(f"wgrad per CTA: producer(w0) wait empty {v[0]:.0f}, mma wait full {v[1]:.0f}, mma wait tempty {v[2]:.0f}, epi(w5) wait tfull {v[3]:.0f}, total {v[4]:.0f} cycles, k-blocks handled by warp 0: {v[5]:.0f}; warp 0 index phase {v[6]:.0f}, issue phase {v[7]:.0f} cycles")
