"""NCCL check of competesmoe_b200.grad_sync.reduce_gradients on real GPUs (run under torchrun, >= 2 ranks).

Every rank holds its own tokens.  Reference: the FULL (unsharded) pretrain layer run on the concatenation of all ranks'
tokens -- its gradients are what a correct data-/expert-parallel step must reproduce.  Tested: (a) expert parallelism over
the whole world (expert shards need no reduction, replicated w_gate is summed over the world), (b) pure data parallelism
(ep_size 1: every gradient summed over the world), both through bucketed NCCL all-reduces.  Prints GRAD_SYNC_OK."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ep_worker import close, pt_args  # noqa: E402


def main():
    world, rank, local = (int(os.environ[k]) for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from competesmoe_b200.ep import EPGroup
    from competesmoe_b200.grad_sync import make_ep_dp_groups, reduce_gradients
    from competesmoe_b200.pretrain import CompeteSMoE
    E, K, D, H, B, N = 16, 4, 256, 128, 2, 256

    def build():
        torch.manual_seed(5)
        layer = CompeteSMoE(D, E, H, n_heads=K, args=pt_args(), activation=F.relu, selection_mode="gate",
                            log_interval=None).to(dev)
        layer.train()
        layer.regularization_present = False
        layer.set_total_steps(id_layer=0)
        layer.prob_flips_final[0] = torch.zeros_like(layer.prob_flips_final[0])
        layer.set_current_steps(1)
        return layer

    def step(layer, x, dy):
        for p in layer.parameters():
            p.grad = None
        x = x.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = layer(x, id_layer=0)
        (out.float() * dy).sum().backward()

    g = torch.Generator().manual_seed(50 + rank)
    x = torch.randn(B, N, D, generator=g).to(dev)
    dy = torch.randn(B, N, D, generator=g).to(dev)
    xs = [torch.empty_like(x) for _ in range(world)]
    dys = [torch.empty_like(dy) for _ in range(world)]
    dist.all_gather(xs, x)
    dist.all_gather(dys, dy)
    full = build()
    step(full, torch.cat(xs, 0), torch.cat(dys, 0))
    for ep_size in (world, 1):
        groups = make_ep_dp_groups(ep_size)
        layer = build()
        epg = None
        if ep_size > 1:
            epg = EPGroup(groups.ep, dev)
            layer.enable_expert_parallel(epg, max_tokens=B * N)
        step(layer, x, dy)
        counts = reduce_gradients(layer, dp_group=groups.dp)
        lo, El = (layer.ep_expert_offset, E // ep_size) if ep_size > 1 else (0, E)
        close(layer.w_gate.grad, full.w_gate.grad, 3e-2, f"ep{ep_size} d w_gate")
        close(layer.keys.grad, full.keys.grad[lo:lo + El], 3e-2, f"ep{ep_size} d keys")
        close(layer.values.grad, full.values.grad[lo:lo + El], 3e-2, f"ep{ep_size} d values")
        if rank == 0:
            print(f"grad_sync ep_size={ep_size} world={world}: ok, collectives {counts}", flush=True)
        if epg is not None:
            epg.close()
    dist.destroy_process_group()
    if rank == 0:
        print("GRAD_SYNC_OK", flush=True)


if __name__ == "__main__":
    main()
