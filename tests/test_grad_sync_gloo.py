"""EP-aware gradient reduction (competesmoe_b200/grad_sync.py) on CPU with gloo, world size 2 and 4: replicated
parameters are summed over the world, expert-parallel ones only over the replicas of the same shard, in a handful of
bucketed collectives instead of one per parameter (reference loop: framework/task/simple_task.py:403-413)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _FakeEPLayer(torch.nn.Module):
    """Parameter layout of the pretrain plugin under expert parallelism: w_gate replicated, keys / values sharded."""

    def __init__(self, el):
        super().__init__()
        self.w_gate = torch.nn.Parameter(torch.zeros(8, 16))
        self.keys = torch.nn.Parameter(torch.zeros(el, 16, 4))
        self.values = torch.nn.Parameter(torch.zeros(el, 4, 16))
        self._ep = object()


class _FakeMMLayer(torch.nn.Module):
    def __init__(self, el):
        super().__init__()
        self.gate = torch.nn.Linear(16, 8, bias=False)
        self.experts = torch.nn.ModuleList([torch.nn.Linear(16, 16) for _ in range(el)])
        self._ep = object()


def _worker(rank, world, ep_size, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from competesmoe_b200.grad_sync import expert_parallel_parameters, make_ep_dp_groups, reduce_gradients
        groups = make_ep_dp_groups(ep_size)
        assert groups.ep_size == ep_size and groups.dp_size == world // ep_size
        model = torch.nn.Sequential(torch.nn.Linear(16, 16), _FakeEPLayer(8 // ep_size), _FakeMMLayer(2),
                                    torch.nn.LayerNorm(16))
        n_exp = len(expert_parallel_parameters(model))
        assert n_exp == 2 + 4                       # keys, values + 2 experts x (weight, bias)
        ep_ids = set(expert_parallel_parameters(model))
        for i, p in enumerate(model.parameters()):
            p.grad = torch.full_like(p, float(rank + 1)) + i
        counts = reduce_gradients(model, dp_group=groups.dp, bucket_bytes=4096)
        n_params = sum(1 for _ in model.parameters())
        assert 1 <= counts["replicated"] < n_params - n_exp + 1
        ep_pos = rank % ep_size
        replicas = [g * ep_size + ep_pos for g in range(world // ep_size)]
        for i, p in enumerate(model.parameters()):
            if id(p) in ep_ids:
                want = sum(r + 1 for r in replicas) + i * len(replicas)
            else:
                want = sum(r + 1 for r in range(world)) + i * world
            assert torch.equal(p.grad, torch.full_like(p, float(want))), (i, float(p.grad.flatten()[0]), want)
        assert (counts["expert"] == 0) == (world == ep_size)
        # average=True divides every gradient by the number of token shards (the world size)
        for p in model.parameters():
            p.grad = torch.ones_like(p)
        reduce_gradients(model, dp_group=groups.dp, average=True)
        for p in model.parameters():
            want = (len(replicas) if id(p) in ep_ids else world) / world
            assert torch.allclose(p.grad, torch.full_like(p, want))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize("world,ep_size", [(2, 2), (2, 1), (4, 2)])
def test_reduce_gradients_gloo(world, ep_size):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, ep_size, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for r, msg in res:
        assert msg == "ok", f"rank {r}:\n{msg}"
