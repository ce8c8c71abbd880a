"""Kernel-level timeline of one MoE-layer step at the bench shape (torch.profiler, CUDA activities).
Usage: python scripts/profile_step.py [router|competition] -> table on stdout.
Under torchrun (WORLD_SIZE > 1) the layer runs expert-parallel over all ranks and rank 0 prints its own timeline."""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402

import os  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "router"
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
ep = None
if world > 1:
    import torch.distributed as dist
    from competesmoe_b200.ep import EPGroup
    dist.init_process_group("nccl", device_id=dev)
    ep = EPGroup(dist.group.WORLD, dev)
layer = bench.build_layer(dev, ep)
params = list(layer.parameters())
g = torch.Generator().manual_seed(1235 + rank)
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
if rank != 0:
    ep.close()
    dist.destroy_process_group()
    sys.exit(0)
print(f"mode={mode}: {len(rows)} distinct kernels, {sum(r[2] for r in rows):.0f} launches/step, {tot:.1f} us of kernel time per step")
for k, t, c in rows[:45]:
    print(f"{t:9.1f} us  x{c:5.1f}  {k[:110]}")
if ep is not None:
    ep.close()
    dist.destroy_process_group()
