"""Where the GPU idles inside one MoE-layer step at the bench shape: kernel start/end times from torch.profiler
(kineto), printed in launch order with the idle gap before each kernel.  Usage: python scripts/timeline_step.py
[router|competition] [graph]."""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "router"
graph = len(sys.argv) > 2 and sys.argv[2] == "graph"
dev = torch.device("cuda", 0)
layer = bench.build_layer(dev, None)
params = list(layer.parameters())
g = torch.Generator().manual_seed(1235)
x = torch.randn(1, bench.TOKENS, bench.D_MODEL, generator=g).bfloat16().to(dev).requires_grad_(True)
dy = torch.randn(1, bench.TOKENS, bench.D_MODEL, generator=g).bfloat16().to(dev)
bench.set_branch(layer, mode == "competition")
if graph:
    layer.enable_cuda_graphs()
for _ in range(5):
    bench.one_step(layer, x, dy, params)
torch.cuda.synchronize()
N = 4
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        bench.one_step(layer, x, dy, params)
    torch.cuda.synchronize()
evs = []
for e in prof.events():
    if "cuda" in str(e.device_type).lower() and e.time_range is not None:
        evs.append((e.time_range.start, e.time_range.end, e.name))
evs.sort()
per = len(evs) // N
last = evs[-per:]                      # the last step
t0, t1 = last[0][0], last[-1][1]
busy = sum(e[1] - e[0] for e in last)
print(f"mode={mode} graph={graph}: {per} device activities / step, span {t1 - t0:.1f} us, busy {busy:.1f} us, idle {t1 - t0 - busy:.1f} us")
prev_end = t0
for s, e, n in last:
    print(f"  gap {s - prev_end:7.1f} us | {e - s:8.1f} us  {n[:100]}")
    prev_end = max(prev_end, e)
# step-to-step period
starts = [evs[i * per][0] for i in range(N)]
print("step period us:", [round(starts[i + 1] - starts[i], 1) for i in range(N - 1)])
