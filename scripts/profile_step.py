"""Kernel-level timeline of one MoE-layer step at the bench shape (torch.profiler, CUDA activities).
Usage: python scripts/profile_step.py [router|competition] -> table on stdout."""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "router"
dev = torch.device("cuda", 0)
layer = bench.build_layer(dev)
params = list(layer.parameters())
g = torch.Generator().manual_seed(1235)
x = torch.randn(1, bench.TOKENS, bench.D_MODEL, generator=g).bfloat16().to(dev).requires_grad_(True)
dy = torch.randn(1, bench.TOKENS, bench.D_MODEL, generator=g).bfloat16().to(dev)
bench.set_branch(layer, mode == "competition")
for _ in range(5):
    bench.one_step(layer, x, dy, params)
torch.cuda.synchronize()
N = 5
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        bench.one_step(layer, x, dy, params)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"] if hasattr(prof.key_averages()[0], "device_type") else prof.key_averages()
rows = sorted(((e.key, e.device_time_total / N, e.count / N) for e in prof.key_averages() if getattr(e, "device_time_total", 0) > 0 and "cuda" in str(getattr(e, "device_type", "")).lower()),
              key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"mode={mode}: {len(rows)} distinct kernels, {sum(r[2] for r in rows):.0f} launches/step, {tot:.1f} us of kernel time per step")
for k, t, c in rows[:45]:
    print(f"{t:9.1f} us  x{c:5.1f}  {k[:110]}")
