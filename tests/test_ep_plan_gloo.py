"""Host logic of the expert-parallel exchange on CPU: two gloo ranks derive the layout from the gathered count matrix
(ep.plan_host = the specification of ep_exchange_plan_kernel), emulate dispatch / return with gloo collectives and check
that the result is exactly the padded expert-major layout of the globally stable-sorted slots."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from competesmoe_b200.ep import plan_host, recv_row_cap


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, E, K, tokens, row_tile, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        El = E // world
        g = torch.Generator().manual_seed(1000 + rank)
        T = tokens[rank]
        # top-k without repeats per token, skewed so that some experts stay empty
        scores = torch.rand(T, E, generator=g) + (torch.arange(E) < E // 2).float() * 0.5
        sel = scores.topk(K, dim=-1).indices.to(torch.int32)
        x = torch.arange(T, dtype=torch.float32).unsqueeze(1) + 1000.0 * rank       # row payload = (rank, token) id
        flat = sel.reshape(-1)
        n = flat.numel()
        counts = torch.bincount(flat.long(), minlength=E).to(torch.int32)
        # rank of each slot inside its expert, in slot order (what slot_to_row - pad_offsets is on the GPU)
        order = torch.sort(flat.long(), stable=True).indices
        first = torch.zeros(E + 1, dtype=torch.long)
        first[1:] = counts.long().cumsum(0)
        rank_in_e = torch.empty(n, dtype=torch.long)
        rank_in_e[order] = torch.arange(n) - first[flat.long()[order]]
        all_counts = [torch.zeros(E, dtype=torch.int32) for _ in range(world)]
        dist.all_gather(all_counts, counts)
        C = torch.stack(all_counts)
        dest_base, recv_counts, recv_pad = plan_host(C, rank, row_tile)
        # every rank derives the same totals
        assert torch.equal(recv_counts.long(), C.long().sum(0)[rank * El:(rank + 1) * El])
        assert all(int(v) % row_tile == 0 for v in recv_pad)
        cap = recv_row_cap(world, max(tokens) * K, El, row_tile)
        assert int(recv_pad[-1]) <= cap
        # ---- dispatch, emulated: (owner, row, payload, tag) per slot, routed with all_gather_object
        owner = flat.long() // El
        row = dest_base.long()[flat.long()] + rank_in_e
        msgs = [(int(owner[j]), int(row[j]), float(x[j // K, 0]), (rank << 32) | j) for j in range(n)]
        everything = [None] * world
        dist.all_gather_object(everything, msgs)
        recv = torch.full((cap,), float("nan"))
        tags = torch.full((cap,), -1, dtype=torch.long)
        for src_msgs in everything:
            for o, r, payload, tag in src_msgs:
                if o == rank:
                    assert tags[r] == -1, "two slots were sent to the same row"
                    recv[r], tags[r] = payload, tag
        # ---- the received layout is the padded expert-major layout of the global stable sort
        all_sel = [None] * world
        dist.all_gather_object(all_sel, flat.tolist())
        for el in range(El):
            e = rank * El + el
            want = [(s << 32) | j for s in range(world) for j, v in enumerate(all_sel[s]) if v == e]
            lo = int(recv_pad[el])
            got = tags[lo:lo + int(recv_counts[el])].tolist()
            assert got == want, f"expert {e}: rows are not in (source rank, source order) order"
            assert (tags[lo + int(recv_counts[el]):int(recv_pad[el + 1])] == -1).all(), "padding rows were written"
        # ---- return, emulated: row r goes back to slot (tag & 0xffffffff) of rank (tag >> 32)
        back = [(int(t) >> 32, int(t) & 0xFFFFFFFF, float(recv[r]) * 2.0) for r, t in enumerate(tags.tolist()) if t >= 0]
        allback = [None] * world
        dist.all_gather_object(allback, back)
        ret = torch.full((n,), float("nan"))
        for msgs_b in allback:
            for s, j, v in msgs_b:
                if s == rank:
                    ret[j] = v
        assert torch.equal(ret, 2.0 * x[:, 0].repeat_interleave(K)), "return trip does not restore slot order"
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize("E,K,tokens,row_tile", [(8, 2, (37, 5), 128), (4, 2, (300, 300), 256), (16, 4, (0, 64), 128)])
def test_ep_plan_two_ranks_gloo(E, K, tokens, row_tile):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, E, K, tokens, row_tile, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for r, msg in res:
        assert msg == "ok", f"rank {r}:\n{msg}"


def test_plan_host_single_rank_is_route_layout():
    counts = torch.tensor([[5, 0, 130, 256]])
    dest, rc, pad = plan_host(counts, 0, 128)
    assert dest.tolist() == [0, 128, 128, 384] and rc.tolist() == [5, 0, 130, 256] and pad.tolist() == [0, 128, 128, 384, 640]
