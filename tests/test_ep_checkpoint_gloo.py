"""Checkpoints under expert parallelism (host logic, gloo, CPU, world size 2): after the experts are sharded the plain
state_dict() holds only the local shard; full_state_dict() all-gathers it back to the REFERENCE layout (identical on
every rank, equal to the state dict before sharding) and load_full_state_dict() puts a full checkpoint back into the
sharded layer.  Both plugins."""
import os
import socket
from types import SimpleNamespace

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn
import torch.nn.functional as F


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _same(a, b):
    return set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from competesmoe_b200.multimodal import CompeteSMoE as MM
        from competesmoe_b200.pretrain import CompeteSMoE as PT
        from oracle import multimodal as om
        from oracle import pretrain as op
        group = SimpleNamespace(world=world, rank=rank, group=None)
        # ---- multimodal: experts.{e}.fc1/fc2
        torch.manual_seed(0)                      # same full model on every rank
        E = 4
        experts = nn.ModuleList([nn.Sequential(nn.Linear(8, 12), nn.GELU(), nn.Linear(12, 8)) for _ in range(E)])
        layer = MM(8, 8, E, 2, experts, om.default_args())
        full = {k: v.clone() for k, v in layer.state_dict().items()}
        layer._shard_experts(rank, world)
        local = layer.state_dict()
        assert {k for k in local if k.startswith("experts.")} == {f"experts.{i}.{j}.{n}" for i in range(E // world)
                                                                   for j in (0, 2) for n in ("weight", "bias")}
        assert torch.equal(local["experts.0.0.weight"], full[f"experts.{rank * (E // world)}.0.weight"])
        assert _same(layer.full_state_dict(group), full)
        # load a different full checkpoint into the sharded layer, gather it back
        other = {k: (v + 1.0 if v.is_floating_point() else v) for k, v in full.items()}
        layer.load_full_state_dict(other, group)
        assert _same(layer.full_state_dict(group), other)
        # ---- pretrain: stacked keys / values / bias
        torch.manual_seed(1)
        pl = PT(16, 8, 4, n_heads=2, args=op.default_args(), activation=F.relu, selection_mode="gate", log_interval=None,
                bias=True)
        with torch.no_grad():
            pl.bias.normal_()
        full = {k: v.clone() for k, v in pl.state_dict().items()}
        pl._shard_experts(rank, world)
        assert pl.keys.shape[0] == 8 // world and pl.state_dict()["w_gate"].shape[0] == 8
        assert _same(pl.full_state_dict(group), full)
        other = {k: v * 2.0 for k, v in full.items()}
        pl.load_full_state_dict(other, group)
        assert _same(pl.full_state_dict(group), other)
        # sharding after a backward pass (i.e. after an optimizer could exist) is refused
        pl2 = PT(16, 8, 4, n_heads=2, args=op.default_args(), activation=F.relu, selection_mode="gate", log_interval=None)
        pl2.keys.grad = torch.zeros_like(pl2.keys)
        try:
            pl2._shard_experts(rank, world)
            raise AssertionError("sharding with live gradients must raise")
        except RuntimeError:
            pass
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as exc:   # surface the failure in the parent
        import traceback
        q.put((rank, traceback.format_exc() + str(exc)))


def test_full_state_dict_round_trip_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert all(r[1] == "ok" for r in res), res
