"""Gradient synchronisation that knows which parameters are expert-parallel.

reference: moe_pretrain_model/framework/task/simple_task.py:403-413 issues one `all_reduce` per parameter over the whole
world; moe_model/train/train.py:1474-1480 registers whole MoE layers as ZeRO-3 leaves.  With expert parallelism
(competesmoe_b200/ep.py) that is both wrong and wasteful for the expert weights: rank r of an EP group owns experts
[r*E/P, (r+1)*E/P) and its gradient already contains the contribution of every token of the group (tokens were
dispatched to the owner), so it must only be summed over the *replicas* of the same shard (the DP group across EP
groups), while replicated parameters (gate / w_gate, o_bias, everything outside the MoE layers) are summed over the
world.  Gradients are flattened into a few large buckets per (group, dtype): NVSwitch collectives are latency-, not
link-bound, so bucket count is what matters.

    groups = make_ep_dp_groups(ep_size)                 # collective: every rank calls it
    layer.enable_expert_parallel(EPGroup(groups.ep, dev), ...)
    ...
    loss.backward()
    reduce_gradients(model, dp_group=groups.dp)          # instead of the per-parameter loop

Same arithmetic as the reference loop (SUM, no averaging) unless `average=True`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors


@dataclass
class EPDPGroups:
    ep: Optional[dist.ProcessGroup]    # ranks that share one set of experts (expert-parallel group of this rank)
    dp: Optional[dist.ProcessGroup]    # ranks that hold the same expert shard (one per EP group); None when there is one EP group
    ep_size: int
    dp_size: int


def make_ep_dp_groups(ep_size: int) -> EPDPGroups:
    """World of W ranks -> W / ep_size expert-parallel groups of consecutive ranks and ep_size data-parallel groups of
    ranks with equal position inside their EP group.  Collective: every rank creates every group, in the same order."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if world % ep_size != 0:
        raise ValueError(f"world size {world} is not a multiple of the expert-parallel degree {ep_size}")
    n_ep = world // ep_size
    my_ep = my_dp = None
    for g in range(n_ep):
        ranks = list(range(g * ep_size, (g + 1) * ep_size))
        pg = dist.new_group(ranks)
        if rank in ranks:
            my_ep = pg
    if n_ep > 1:
        for pos in range(ep_size):
            ranks = [g * ep_size + pos for g in range(n_ep)]
            pg = dist.new_group(ranks)
            if rank in ranks:
                my_dp = pg
    return EPDPGroups(my_ep, my_dp, ep_size, n_ep)


def expert_parallel_parameters(module: torch.nn.Module) -> Dict[int, torch.nn.Parameter]:
    """id -> parameter for every parameter that is sharded over an EP group: the expert weights of layers on which
    enable_expert_parallel() was called (multimodal: `experts.*`; pretrain: keys / values / bias)."""
    out: Dict[int, torch.nn.Parameter] = {}
    for m in module.modules():
        if getattr(m, "_ep", None) is None:
            continue
        if hasattr(m, "experts") and isinstance(getattr(m, "experts"), torch.nn.Module):
            for p in m.experts.parameters():
                out[id(p)] = p
        for name in ("keys", "values", "bias"):          # (the shared expert of deepseekv2/3 stays replicated)
            p = getattr(m, name, None)
            if isinstance(p, torch.nn.Parameter):
                out[id(p)] = p
    return out


def _buckets(grads: List[torch.Tensor], bucket_bytes: int) -> Iterable[List[torch.Tensor]]:
    by_dtype: Dict[torch.dtype, List[torch.Tensor]] = {}
    for g in grads:
        by_dtype.setdefault(g.dtype, []).append(g)
    for gs in by_dtype.values():
        cur, size = [], 0
        for g in gs:
            nbytes = g.numel() * g.element_size()
            if cur and size + nbytes > bucket_bytes:
                yield cur
                cur, size = [], 0
            cur.append(g)
            size += nbytes
        if cur:
            yield cur


def _all_reduce_bucketed(grads: List[torch.Tensor], group, bucket_bytes: int, scale: Optional[float]) -> int:
    work: List[Tuple[object, torch.Tensor, List[torch.Tensor]]] = []
    for bucket in _buckets(grads, bucket_bytes):
        flat = _flatten_dense_tensors(bucket)
        work.append((dist.all_reduce(flat, group=group, async_op=True), flat, bucket))
    for handle, flat, bucket in work:
        handle.wait()
        if scale is not None:
            flat.mul_(scale)
        for g, synced in zip(bucket, _unflatten_dense_tensors(flat, bucket)):
            g.copy_(synced)
    return len(work)


def reduce_gradients(module: torch.nn.Module, dp_group: Optional[dist.ProcessGroup] = None,
                     world_group: Optional[dist.ProcessGroup] = None, average: bool = False,
                     bucket_bytes: int = 64 << 20) -> Dict[str, int]:
    """Sum (or average) the gradients after backward: replicated parameters over `world_group` (default: the whole
    world), expert-parallel parameters over `dp_group` only (skipped when there is a single EP group).  Returns the
    number of collectives issued per class -- the reference loop issues one per parameter."""
    if not (dist.is_available() and dist.is_initialized()):
        return {"replicated": 0, "expert": 0}
    ep_params = expert_parallel_parameters(module)
    rep, exp = [], []
    for p in module.parameters():
        if p.grad is None:
            continue
        if not p.grad.is_contiguous():
            p.grad = p.grad.contiguous()
        (exp if id(p) in ep_params else rep).append(p.grad)
    n_world = dist.get_world_size(world_group)
    out = {"replicated": _all_reduce_bucketed(rep, world_group, bucket_bytes, 1.0 / n_world if average else None),
           "expert": 0}
    if exp and dp_group is not None and dist.get_world_size(dp_group) > 1:
        # averaging over tokens means dividing by the number of token shards, i.e. the world size, for experts too
        out["expert"] = _all_reduce_bucketed(exp, dp_group, bucket_bytes, 1.0 / n_world if average else None)
    elif exp and average:
        for g in exp:
            g.mul_(1.0 / n_world)
    return out
