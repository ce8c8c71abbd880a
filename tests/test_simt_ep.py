"""The expert-parallel peer-memory kernels (competesmoe_b200/csrc/ep.cu) on the CPU, at group sizes 2, 4 and 8.

The kernels are compiled from their own source against the SIMT emulator (tests/simt/); the "ranks" are forked processes
and the symmetric buffers are anonymous shared mappings, which a forked child sees at the parent's addresses -- exactly
the property CUDA IPC gives the real thing (every rank holds a pointer to every peer's buffer).  The release / acquire
flag accesses of the device-side barrier become `__atomic` builtins, so the barrier protocol itself runs across processes.

Checked: the barrier (no rank passes before all arrived, epochs advance together), `ep_exchange_plan` against its host
specification `ep.plan_host`, dispatch -> row pointers -> return as a round trip (every slot gets back exactly the row it
sent, received rows are expert-major and ordered by source rank), cast + all-gather (`ep_gather_push`) and the
deterministic reduce-scatter (`ep_reduce_pull`, all three instantiations: P <= 2 / 4 / 8).  The group size 8 code paths
have not been run on hardware by the builder this round (profiles/r02_ep_scaling.md); this is their first execution.
TEST INFRASTRUCTURE: no NVLink, no memory-model fidelity beyond acquire / release, no timing.
"""
import ctypes as C
import mmap
import multiprocessing as mp
import os
import traceback

import numpy as np
import pytest
import torch

import simt_host

BF16, F32 = 1, 0
CTRL_BYTES = 4096 + 16 * 1024 * 4


class Shared:
    """Anonymous shared mapping: same address in every forked child."""

    def __init__(self, nbytes: int):
        self.nbytes = max(int(nbytes), 4096)
        self.mm = mmap.mmap(-1, self.nbytes)
        self.addr = C.addressof(C.c_char.from_buffer(self.mm))

    def array(self, dtype, count=-1, offset=0):
        return np.frombuffer(self.mm, dtype=dtype, count=count, offset=offset)


def peers(bufs, offset=0):
    return (C.c_void_p * len(bufs))(*[b.addr + offset for b in bufs])


def ptr(a):
    return None if a is None else a.ctypes.data


# numpy only: these run inside forked children, where torch's OpenMP pool (threads of the parent) must not be touched
def bf16_bits(x: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)          # round to nearest even (finite inputs)


def bf16_vals(bits: np.ndarray) -> np.ndarray:
    return (bits.astype(np.uint32) << 16).view(np.float32)


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    lib, stats = simt_host.build(tmp_path_factory.mktemp("simt_ep"), ep=True)
    assert stats["ep.cu"]["sys_flags"] == 3 and stats["ep.cu"]["launches"] >= 7
    return lib


def run_ranks(P, fn, *args, timeout=180):
    """fn(rank, *args) -> picklable result, in P forked processes; returns the list of results by rank."""
    ctx = mp.get_context("fork")
    q = ctx.Queue()

    def body(rank):
        try:
            q.put((rank, True, fn(rank, *args)))
        except BaseException:
            q.put((rank, False, traceback.format_exc()))

    procs = [ctx.Process(target=body, args=(r,)) for r in range(P)]
    for p in procs:
        p.start()
    out = {}
    try:
        for _ in range(P):
            rank, ok, res = q.get(timeout=timeout)
            assert ok, f"rank {rank} failed:\n{res}"
            out[rank] = res
    finally:
        for p in procs:
            p.join(timeout=5)
            if p.is_alive():
                p.kill()
    return [out[r] for r in range(P)]


def check(lib, rc, what):
    assert rc == 0, f"{what}: status {rc}: {lib.csmoe_last_error().decode(errors='replace')}"


# ------------------------------------------------------------------------------------------------ barrier
@pytest.mark.parametrize("P", [2, 4, 8])
def test_flag_barrier_orders_the_ranks(lib, P):
    ctrl = [Shared(CTRL_BYTES) for _ in range(P)]
    board = Shared(4096)                       # board[i * P + r] = 1 once rank r has finished phase i

    def rank_fn(rank):
        b = board.array(np.int32)
        seen = []
        for phase in range(4):
            b[phase * P + rank] = 1
            for ch in (0, 1):                  # two independent channels, as WeightExchange uses them
                check(lib, lib.csmoe_ep_barrier(peers(ctrl, 128 * ch), ctrl[rank].addr + 2048 + 64 * ch, rank, P, None), "barrier")
            seen.append(int(b[phase * P:(phase + 1) * P].sum()))   # after the barrier every rank's mark must be visible
        epoch = ctrl[rank].array(np.int32, 1, 2048)[0], ctrl[rank].array(np.int32, 1, 2048 + 64)[0]
        return seen, tuple(int(e) for e in epoch)

    for seen, epoch in run_ranks(P, rank_fn):
        assert seen == [P] * 4 and epoch == (4, 4)


# ------------------------------------------------------------------------------------------------ exchange plan
@pytest.mark.parametrize("P,E,row_tile", [(2, 4, 256), (4, 16, 128), (8, 8, 256), (8, 64, 128)])
def test_exchange_plan_matches_host_specification(lib, P, E, row_tile):
    from competesmoe_b200.ep import plan_host, recv_row_cap
    ctrl = [Shared(CTRL_BYTES) for _ in range(P)]
    rng = np.random.default_rng(P * 100 + E)
    counts_all = rng.integers(0, 700, size=(P, E)).astype(np.int32)
    counts_all[:, rng.integers(0, E)] = 0                      # an expert nobody routes to
    counts_all[rng.integers(0, P)] = 0                         # and a rank without tokens
    max_slots = int(counts_all.sum(1).max())
    El = E // P
    row_cap = recv_row_cap(P, max_slots, El, row_tile)

    def rank_fn(rank):
        counts = counts_all[rank].copy()
        dest_base, recv_counts = np.zeros(E, np.int32), np.zeros(El, np.int32)
        recv_pad, tile_expert = np.zeros(El + 1, np.int32), np.full(row_cap // 128, -7, np.int32)
        for _ in range(2):                                     # twice: the count matrix and the flags are reused
            check(lib, lib.csmoe_ep_exchange_plan(ptr(counts), peers(ctrl, 4096), peers(ctrl, 0), ctrl[rank].addr + 2048, rank, P, E,
                                                  row_tile, row_cap, ptr(dest_base), ptr(recv_counts), ptr(recv_pad),
                                                  ptr(tile_expert), None), "exchange_plan")
        return dest_base, recv_counts, recv_pad, tile_expert

    for rank, (dest_base, recv_counts, recv_pad, tile_expert) in enumerate(run_ranks(P, rank_fn)):
        want_base, want_counts, want_pad = plan_host(torch.from_numpy(counts_all), rank, row_tile)
        assert np.array_equal(dest_base, want_base.numpy()), f"rank {rank}: dest_base"
        assert np.array_equal(recv_counts, want_counts.numpy()) and np.array_equal(recv_pad, want_pad.numpy())
        want_te = np.full(row_cap // 128, -1, np.int32)
        for el in range(El):
            want_te[recv_pad[el] // 128:recv_pad[el + 1] // 128] = el
        assert np.array_equal(tile_expert, want_te), f"rank {rank}: tile_expert"
        assert recv_pad[El] <= row_cap


# ------------------------------------------------------------------------------------------------ dispatch / return round trip
@pytest.mark.parametrize("P,E,K,T,D", [(2, 4, 2, 150, 64), (4, 8, 2, 90, 128), (8, 8, 1, 40, 64), (8, 16, 4, 33, 64)])
def test_dispatch_and_return_round_trip(lib, P, E, K, T, D):
    from competesmoe_b200.ep import recv_row_cap
    row_tile, El = 128, E // P
    max_slots = T * K
    row_cap = recv_row_cap(P, max_slots, El, row_tile)
    ctrl = [Shared(CTRL_BYTES) for _ in range(P)]
    recv_x = [Shared(row_cap * D * 2) for _ in range(P)]
    tags = [Shared(row_cap * 8) for _ in range(P)]
    ret_y = [Shared(max_slots * D * 2) for _ in range(P)]
    rng = np.random.default_rng(7)
    tokens = [int(rng.integers(1, T + 1)) for _ in range(P)]            # ragged: every rank has its own token count
    tokens[0] = T
    sels = [np.stack([rng.permutation(E)[:K] for _ in range(t)]).astype(np.int32) for t in tokens]
    # token t of rank r is the row filled with the bf16-exact value 1 + (r * 37 + t) % 240
    xs = [np.repeat(((1 + (r * 37 + np.arange(t)) % 240).astype(np.float32))[:, None], D, axis=1) for r, t in enumerate(tokens)]

    def rank_fn(rank):
        Tn = tokens[rank]
        n = Tn * K
        sel = np.ascontiguousarray(sels[rank].reshape(-1))
        x = bf16_bits(xs[rank])
        # routing maps of this rank's own slots (the same library, local launch)
        lcap = int(lib.csmoe_route_row_cap(n, E, row_tile))
        ws = np.zeros(max(int(lib.csmoe_route_workspace_bytes(n, E)) // 4, 1), np.int32)
        counts, offsets, pad = np.zeros(E, np.int32), np.zeros(E + 1, np.int32), np.zeros(E + 1, np.int32)
        ssel, sidx, s2r = np.zeros(n, np.int32), np.zeros(n, np.int64), np.zeros(n, np.int32)
        r2s, te = np.zeros(lcap, np.int32), np.zeros(lcap // 128, np.int32)
        check(lib, lib.csmoe_route_build(ptr(sel), n, E, row_tile, lcap, ptr(counts), ptr(offsets), ptr(pad), ptr(ssel), ptr(sidx),
                                         ptr(s2r), ptr(r2s), ptr(te), ptr(ws), None), "route_build")
        dest_base, recv_counts = np.zeros(E, np.int32), np.zeros(El, np.int32)
        recv_pad, tile_expert = np.zeros(El + 1, np.int32), np.zeros(row_cap // 128, np.int32)
        flags, epoch = peers(ctrl, 0), ctrl[rank].addr + 2048
        check(lib, lib.csmoe_ep_exchange_plan(ptr(counts), peers(ctrl, 4096), flags, epoch, rank, P, E, row_tile, row_cap,
                                              ptr(dest_base), ptr(recv_counts), ptr(recv_pad), ptr(tile_expert), None), "plan")
        check(lib, lib.csmoe_ep_dispatch(ptr(x), BF16, D, K, n, ptr(sel), ptr(s2r), ptr(pad), ptr(dest_base), El, None,
                                         peers(recv_x), peers(tags), rank, P, None), "dispatch")
        check(lib, lib.csmoe_ep_barrier(flags, epoch, rank, P, None), "barrier")
        c_rows = np.zeros(row_cap, np.uint64)
        check(lib, lib.csmoe_ep_row_ptrs(tags[rank].addr, ptr(tile_expert), ptr(recv_counts), ptr(recv_pad), El, row_cap,
                                         peers(ret_y), D, BF16, P, ptr(c_rows), recv_x[rank].addr, D, None), "row_ptrs")
        received = bf16_vals(recv_x[rank].array(np.uint16, row_cap * D)).reshape(row_cap, D).copy()
        # the "expert": identity.  Every received row goes back to the slot it came from.
        check(lib, lib.csmoe_ep_push_rows(recv_x[rank].addr, BF16, D, D, row_cap, ptr(c_rows), None), "push_rows")
        check(lib, lib.csmoe_ep_barrier(flags, epoch, rank, P, None), "barrier")
        back = bf16_vals(ret_y[rank].array(np.uint16, n * D)).reshape(n, D).copy()
        return back, received, recv_counts, recv_pad, c_rows != 0

    res = run_ranks(P, rank_fn)
    for rank, (back, received, recv_counts, recv_pad, routed) in enumerate(res):
        want = np.repeat(xs[rank], K, axis=0)                       # slot j = (t, k) carries token t = j // K
        assert np.array_equal(back, want), f"rank {rank}: a slot did not get its own row back"
        # receive side: expert-major, inside an expert ordered by source rank, inside a source rank by slot order
        for el in range(El):
            e = rank * El + el
            rows = [xs[s][t] for s in range(P) for t, k in zip(*np.nonzero(sels[s] == e))]
            assert recv_counts[el] == len(rows)
            seg = received[recv_pad[el]:recv_pad[el] + len(rows)]
            assert np.array_equal(seg, np.stack(rows) if rows else seg), f"rank {rank} expert {e}: received rows out of order"
            assert bool(routed[recv_pad[el]:recv_pad[el] + len(rows)].all())
            pad_rows = received[recv_pad[el] + len(rows):recv_pad[el + 1]]
            assert not pad_rows.any() and not routed[recv_pad[el] + len(rows):recv_pad[el + 1]].any()   # padding zeroed, no return address


# ------------------------------------------------------------------------------------------------ weights move, rows stay
@pytest.mark.parametrize("P", [2, 3, 4, 8])
def test_gather_push_and_reduce_pull(lib, P):
    n = 8 * 1037                                                     # per-rank shard elements (vectors of 8)
    op_copy = [Shared(P * n * 2) for _ in range(P)]                  # bf16 operand copy of ALL experts on every rank
    grads = [Shared(P * n * 4) for _ in range(P)]                    # full-size fp32 gradient buffer of every rank
    ctrl = [Shared(CTRL_BYTES) for _ in range(P)]
    shard = lambda r: ((np.arange(n) * 7 + r * 13) % 251).astype(np.float32) - 125.0        # bf16-exact integers
    grad = lambda r: ((np.arange(P * n) * 3 + r * 5) % 1021).astype(np.float32) - 510.0     # exact in fp32 sums

    def rank_fn(rank):
        src = shard(rank)
        flags, epoch = peers(ctrl, 0), ctrl[rank].addr + 2048
        check(lib, lib.csmoe_ep_gather_push(ptr(src), F32, n, peers(op_copy), BF16, rank * n, rank, P, None), "gather_push")
        check(lib, lib.csmoe_ep_barrier(flags, epoch, rank, P, None), "barrier")
        full = bf16_vals(op_copy[rank].array(np.uint16, P * n)).copy()
        grads[rank].array(np.float32, P * n)[:] = grad(rank)
        check(lib, lib.csmoe_ep_barrier(flags, epoch, rank, P, None), "barrier")
        out32 = np.zeros(n, np.float32)
        check(lib, lib.csmoe_ep_reduce_pull(peers(grads), rank * n, n, ptr(out32), F32, P, None), "reduce_pull f32")
        out16 = np.zeros(n, np.uint16)
        check(lib, lib.csmoe_ep_reduce_pull(peers(grads), rank * n, n, ptr(out16), BF16, P, None), "reduce_pull bf16")
        check(lib, lib.csmoe_ep_barrier(flags, epoch, rank, P, None), "barrier")
        return full, out32, bf16_vals(out16)

    want_full = np.concatenate([shard(r) for r in range(P)])
    for rank, (full, out32, out16) in enumerate(run_ranks(P, rank_fn)):
        assert np.array_equal(full, want_full), f"rank {rank}: operand copy incomplete"
        acc = np.zeros(n, np.float32)
        for r in range(P):                                           # ascending rank order, fp32: the kernel's order
            acc += grad(r)[rank * n:(rank + 1) * n]
        assert np.array_equal(out32, acc), f"rank {rank}: reduced slice"
        assert np.array_equal(out16, bf16_vals(bf16_bits(acc)))


# ------------------------------------------------------------------------------------------------ whole layers, sharded
@pytest.mark.parametrize("world", [4] + ([2, 8] if os.environ.get("CSMOE_SIMT_FULL", "0") == "1" else []))
def test_sharded_layers_match_unsharded_layers_on_emulated_ranks(tmp_path_factory, world):
    """tests/ep_worker.py's parity cases (what `bench.py --gpus N` reports as `ep_parity`) on `world` emulated ranks:
    spawned processes, gloo, the peer-memory kernels on shared memory.  Both plugins, both steps, both exchange modes of
    the pretrain layer, ragged token counts.  The builder's GPU budget ended before an 8-GPU run of this code:
    CSMOE_SIMT_FULL=1 adds group sizes 2 and 8 (8 ranks: 130 s on 8 cores, run and green at every change of ep.py /
    ep.cu / ep_worker.py) and the pretrain competition step; the peer-memory kernels themselves run at group size 8 in
    the default suite (tests above)."""
    import socket
    import simt_ep_worker
    workdir = tmp_path_factory.mktemp(f"simt_ep_layers{world}")
    simt_host.build(workdir, ref_gemm=True, ep=True)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=simt_ep_worker.rank_main, args=(r, world, port, str(workdir / "libcsmoe_simt.so"), q))
             for r in range(world)]
    for p in procs:
        p.start()
    try:
        for _ in range(world):
            rank, ok, res = q.get(timeout=900)
            assert ok, f"rank {rank} failed:\n{res}"
            assert len(res) >= 5 and res[-1] == "ragged glu top-1", res
    finally:
        for p in procs:
            p.join(timeout=10)
            if p.is_alive():
                p.kill()
