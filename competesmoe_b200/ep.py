"""Expert parallelism (stage 6): experts sharded over the ranks of a process group, tokens stay data-parallel.

Rank r of P owns experts [r*E/P, (r+1)*E/P).  `torch.distributed` is used for plumbing only (exchanging CUDA IPC
handles once, all-gather / reduce-scatter of expert weights for the rare competition step); the data path is
libcsmoe kernels that store straight into peer HBM over NVLink (csrc/ep.cu):

    forward   route_build -> exchange_plan (counts to all peers + barrier + layout) -> dispatch (permute + send)
              -> barrier -> row_ptrs -> grouped GEMM 1 -> grouped GEMM 2 whose epilogue writes every output row into the
              source rank's return buffer -> barrier -> local gate-weighted combine
    backward  dispatch(w * dout) -> barrier -> wgrad / dgrad GEMMs, the last of which returns dx rows the same way
              -> barrier -> local reduce over k

No host synchronisation and no host-side split sizes: buffer capacities are static worst cases (every slot of every
rank routed to one rank), so the step stays CUDA-graph capturable.  The reference has no counterpart (data-parallel
only).
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.distributed as dist
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib, ops
from ._lib import ROW_TILE, check
from .functional import FFNSpec, _bf16, _ffn_first, _FUSE_ACT_BIAS, _FUSE_BWD, sigma_fused_ok

MAX_EXPERTS = 1024
# WeightExchange: run the gather under the router kernels and the gradient reduction under the dx reduction (side stream)
_WX_OVERLAP = os.environ.get("CSMOE_WX_OVERLAP", "1") != "0"
_CTRL_BYTES = 4096 + 16 * MAX_EXPERTS * 4   # flags, epoch, counts_all[P<=16][E<=1024]


# ------------------------------------------------------------------------------------------------ host mirror of the plan
def plan_host(counts_all: torch.Tensor, rank: int, row_tile: int):
    """Pure-integer mirror of ep_exchange_plan_kernel (csrc/ep.cu), used by the CPU tests and as its specification.

    counts_all [P, E]: rows rank s sends to expert e.  Returns (dest_base [E], recv_counts [E/P], recv_pad_offsets [E/P+1]).
    Every rank derives the same layout from the same matrix: inside owner o's receive space expert e starts at the
    row_tile-aligned prefix sum of the totals of o's experts, and inside an expert rows are ordered by source rank."""
    P, E = counts_all.shape
    El = E // P
    c = counts_all.to(torch.int64)
    total = c.sum(0)
    before = c[:rank].sum(0)
    padded = (total + row_tile - 1) // row_tile * row_tile
    start = torch.zeros(E, dtype=torch.int64)
    end = torch.zeros(P, dtype=torch.int64)
    for o in range(P):
        run = 0
        for el in range(El):
            start[o * El + el] = run
            run += int(padded[o * El + el])
        end[o] = run
    dest_base = start + before
    recv_counts = total[rank * El:(rank + 1) * El]
    recv_pad = torch.cat([start[rank * El:(rank + 1) * El], end[rank:rank + 1]])
    return dest_base.to(torch.int32), recv_counts.to(torch.int32), recv_pad.to(torch.int32)


def recv_row_cap(world: int, max_slots_per_rank: int, experts_per_rank: int, row_tile: int) -> int:
    """Static capacity of a receive buffer: every slot of every rank lands on this rank."""
    n = world * max_slots_per_rank
    worst = n + experts_per_rank * (row_tile - 1)
    return (worst + row_tile - 1) // row_tile * row_tile


# ------------------------------------------------------------------------------------------------ symmetric memory
class _CAI:
    """Minimal __cuda_array_interface__ holder so torch can alias memory that libcsmoe allocated."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class SymmetricBuffer:
    """The same allocation on every rank of the group, each mapped into every other rank's address space."""

    def __init__(self, local_ptr: int, peer_ptrs: List[int], nbytes: int, device: torch.device):
        self.local_ptr, self.peer_ptrs, self.nbytes, self.device = local_ptr, peer_ptrs, nbytes, device
        self._bytes = torch.as_tensor(_CAI(local_ptr, nbytes), device=device)

    def peers(self, offset: int = 0):
        """Host array of P device pointers (entry r = this buffer on rank r, shifted by `offset` bytes)."""
        arr = (C.c_void_p * len(self.peer_ptrs))()
        for i, p in enumerate(self.peer_ptrs):
            arr[i] = p + offset
        return arr

    def tensor(self, offset: int, shape, dtype: torch.dtype) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= s
        nb = n * torch.empty((), dtype=dtype).element_size()
        assert offset + nb <= self.nbytes, "symmetric buffer too small"
        return self._bytes[offset:offset + nb].view(dtype).view(*shape)


class EPGroup:
    """Process-group-wide state: rank / world, the control block (barrier flags, epoch, count matrix) and scratch."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None, device: Optional[torch.device] = None):
        self.group = group
        on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if on else 1
        self.rank = dist.get_rank(group) if on else 0
        assert self.world <= 16, "libcsmoe expert parallelism supports up to 16 ranks per group"
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self._owned: List[SymmetricBuffer] = []
        self.ctrl = self.alloc(_CTRL_BYTES)
        self._flags = self.ctrl.peers(0)
        self._counts_all = self.ctrl.peers(4096)
        self._epoch_ptr = self.ctrl.local_ptr + 2048
        self._scratch = {}
        self._identity = {}
        self._flags_ch = {}

    # ---- memory
    def alloc(self, nbytes: int) -> SymmetricBuffer:
        """Collective: every rank of the group must call this in the same order with the same size."""
        lib = _lib.load()
        nbytes = (int(nbytes) + 255) // 256 * 256
        hb = lib.csmoe_ep_ipc_handle_bytes()
        handle = C.create_string_buffer(hb)
        ptr = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.csmoe_ep_alloc(nbytes, C.byref(ptr), handle), "csmoe_ep_alloc")
            peers = [0] * self.world
            peers[self.rank] = ptr.value
            if self.world > 1:
                handles: List[Optional[bytes]] = [None] * self.world
                dist.all_gather_object(handles, bytes(handle.raw), group=self.group)
                for r, h in enumerate(handles):
                    if r == self.rank:
                        continue
                    p = C.c_void_p()
                    check(lib.csmoe_ep_open(C.create_string_buffer(h, hb), C.byref(p)), "csmoe_ep_open")
                    peers[r] = p.value
                dist.barrier(group=self.group)
        buf = SymmetricBuffer(ptr.value, peers, nbytes, self.device)
        self._owned.append(buf)
        return buf

    def scratch(self, name: str, nbytes: int) -> SymmetricBuffer:
        """Group-wide scratch (consumed inside one forward or one backward); grown collectively on demand."""
        buf = self._scratch.get(name)
        if buf is None or buf.nbytes < nbytes:
            buf = self.alloc(nbytes)
            self._scratch[name] = buf
        return buf

    def identity(self, n: int) -> torch.Tensor:
        t = self._identity.get(n)
        if t is None:
            t = torch.arange(n, dtype=torch.int32, device=self.device)
            self._identity[n] = t
        return t

    def close(self):
        lib = _lib.load()
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)
        for buf in self._owned:
            for r, p in enumerate(buf.peer_ptrs):
                if r != self.rank and p:
                    lib.csmoe_ep_close(p)
        if self.world > 1:
            dist.barrier(group=self.group)
        for buf in self._owned:
            lib.csmoe_ep_free(buf.local_ptr)
        self._owned.clear()

    # ---- kernels
    def barrier(self, channel: int = 0):
        """Device-side barrier on the current stream.  Channels have their own flag vectors and epoch counters: two
        streams of one rank may each run a barrier sequence as long as every rank issues the same sequence per channel."""
        assert 0 <= channel < 8
        if channel == 0:
            flags = self._flags
        else:
            flags = self._flags_ch.get(channel)
            if flags is None:
                flags = self._flags_ch[channel] = self.ctrl.peers(128 * channel)
        ops._call("csmoe_ep_barrier", flags, self._epoch_ptr + 64 * channel, self.rank, self.world, ops._stream())

    def exchange_plan(self, counts: torch.Tensor, num_experts: int, row_tile: int, row_cap: int) -> "EPPlan":
        assert num_experts % self.world == 0 and num_experts <= MAX_EXPERTS
        El = num_experts // self.world
        i32 = dict(dtype=torch.int32, device=counts.device)
        dest_base = torch.empty(num_experts, **i32)
        recv_counts = torch.empty(El, **i32)
        recv_pad = torch.empty(El + 1, **i32)
        tile_expert = torch.empty(row_cap // ROW_TILE, **i32)
        ops._call("csmoe_ep_exchange_plan", counts.data_ptr(), self._counts_all, self._flags, self._epoch_ptr, self.rank,
                  self.world, num_experts, row_tile, row_cap, dest_base.data_ptr(), recv_counts.data_ptr(),
                  recv_pad.data_ptr(), tile_expert.data_ptr(), ops._stream())
        return EPPlan(num_experts, El, row_tile, row_cap, dest_base, recv_counts, recv_pad, tile_expert)

    def dispatch(self, src: torch.Tensor, top_k: int, route: ops.Route, plan: "EPPlan", recv, tags=None,
                 slot_w: Optional[torch.Tensor] = None):
        src = src.contiguous()
        if slot_w is not None:
            slot_w = slot_w.reshape(-1).contiguous()
            assert slot_w.dtype == torch.float32
        ops._call("csmoe_ep_dispatch", src.data_ptr(), ops._dt(src), src.shape[1], top_k, route.n_slots,
                  route.sel.data_ptr(), route.slot_to_row.data_ptr(), route.pad_offsets.data_ptr(),
                  plan.dest_base.data_ptr(), plan.experts_per_rank, ops._p(slot_w), recv, tags, self.rank, self.world,
                  ops._stream())

    def row_ptrs(self, tags: Optional[torch.Tensor], plan: "EPPlan", ret, ret_ld: int, dtype: torch.dtype,
                 recv: Optional[torch.Tensor], want_ptrs: bool = True) -> Optional[torch.Tensor]:
        c_rows = torch.empty(plan.row_cap, dtype=torch.int64, device=self.device) if want_ptrs else None
        dt = _lib.BF16 if dtype == torch.bfloat16 else _lib.F32
        ops._call("csmoe_ep_row_ptrs", ops._p(tags), plan.tile_expert.data_ptr(), plan.recv_counts.data_ptr(),
                  plan.recv_pad_offsets.data_ptr(), plan.experts_per_rank, plan.row_cap, ret, ret_ld, dt, self.world,
                  ops._p(c_rows), ops._p(recv), recv.shape[1] if recv is not None else 0, ops._stream())
        return c_rows


@dataclass
class EPPlan:
    num_experts: int
    experts_per_rank: int
    row_tile: int
    row_cap: int
    dest_base: torch.Tensor
    recv_counts: torch.Tensor
    recv_pad_offsets: torch.Tensor
    tile_expert: torch.Tensor

    def local_route(self, n_slots_hint: int) -> ops.Route:
        """The received rows as a Route the grouped GEMM understands (only the fields it reads are filled)."""
        return ops.Route(self.experts_per_rank, 1, n_slots_hint, self.row_cap, None, self.recv_counts, None,
                         self.recv_pad_offsets, None, None, None, None, self.tile_expert, self.row_tile)

    def fused_route(self, c_rows: torch.Tensor) -> ops.Route:
        """The received rows as the fused sigma-MoE kernels see them: one "slot" per received row (row r is slot r when
        it holds a routed row, -1 on padding), so the kernels' validity masks and row gathers work on the receive buffer
        directly.  c_rows (the return addresses) is 0 exactly on padding rows."""
        rows = torch.arange(self.row_cap, dtype=torch.int32, device=c_rows.device)
        rmap = torch.where(c_rows != 0, rows, torch.full_like(rows, -1))
        return ops.Route(self.experts_per_rank, 1, self.row_cap, self.row_cap, None, self.recv_counts, None,
                         self.recv_pad_offsets, None, None, None, rmap, self.tile_expert, self.row_tile)


class EPLayerState:
    """Per-layer exchange buffers that must live from forward to backward (the received tokens are the activations
    autograd would have saved anyway).  One outstanding forward per layer."""

    def __init__(self, group: EPGroup, num_experts: int, top_k: int, d_in: int, d_out: int, max_tokens: int,
                 row_tile: int = 256):
        assert num_experts % group.world == 0, "the number of experts must be a multiple of the EP group size"
        self.group, self.E, self.K, self.D, self.Dout = group, num_experts, top_k, d_in, d_out
        self.El = num_experts // group.world
        self.max_tokens = max_tokens
        self.row_tile = row_tile
        self.max_slots = max_tokens * top_k
        self.row_cap = recv_row_cap(group.world, self.max_slots, self.El, row_tile)
        self.recv_x = group.alloc(self.row_cap * d_in * 2)
        self.tags = group.alloc(self.row_cap * 8)
        self.ret_y = group.alloc(self.max_slots * d_out * 2)
        # backward-only scratch is shared by every layer of the group (consumed inside one backward call)
        group.scratch("recv_dy", self.row_cap * d_out * 2)
        group.scratch("ret_dx", self.max_slots * d_in * 2)



# ------------------------------------------------------------------------------------------------ weights move, tokens stay
class WeightExchange:
    """Expert parallelism for SMALL experts (the sigma-MoE shapes of the pretraining plugin): parameters, gradients and
    optimizer state stay sharded (rank r owns experts [r*E/P, (r+1)*E/P)), but for compute every rank holds a full-size
    operand copy of all experts and runs the ordinary single-GPU kernels on its own tokens.  Per layer step and rank
    this moves (2 + 4) * |experts| * (P-1)/P bytes (bf16 operands out, fp32 gradients in) instead of the
    4 * T*K*D*2 * (P-1)/P bytes of dispatching the K-fold expanded token rows and their gradients: at BASELINE
    configs[3] (d=1024, H=128, E=64, K=8, 8192 tokens per GPU) 88 MB against 470 MB.  `prefer_weights()` is the rule.

        forward    gather_push: cast + all-gather in one kernel, each rank stores its shard into every rank's copy
                   -> barrier -> local fused expert kernels on the full copy
        backward   weight-gradient GEMMs write full-size fp32 buffers -> barrier -> reduce_pull: the owner sums its
                   slice of every rank's buffer in ascending rank order (deterministic)

    Everything is kernels over IPC-mapped peer memory plus the device-side flag barrier (csrc/ep.cu), so both the router
    and the competition step stay CUDA-graph capturable.  Hazards: the operand copy of a layer is rewritten by the next
    forward of that layer -- a peer can only get there after the barrier of this layer's backward, i.e. after every
    rank's last read; without a backward in between (evaluation, activation recomputation) the parameters have not
    changed and the rewrite stores identical bytes.  The gradient buffers are per layer and rewritten one step later,
    with that step's forward barrier in between.  No counterpart in the reference (data-parallel all-reduce of every
    gradient, simple_task.py:403-413; ZeRO in moe_model/train/train.py:1474-1480)."""

    def __init__(self, group: EPGroup, shards: dict):
        """shards: role ('w1', 'b1', 'w2', 'b2') -> this rank's parameter shard [E/P, ...] (or None).  Collective."""
        self.group = group
        self._op = {}       # (role, dtype) -> (SymmetricBuffer, full shape)
        self._grad = {}     # role -> (SymmetricBuffer, full shape)
        self._have = {}     # role -> full operand tensor gathered in the current layer step
        self.consumers = 0
        self._prefetched = None
        self._side = None
        for role in ("w1", "b1", "w2", "b2"):
            t = shards.get(role)
            if t is None:
                continue
            assert t.numel() % 8 == 0, "weight exchange moves 8-element vectors: the shard size must be a multiple of 8"
            full = (group.world * t.shape[0], *t.shape[1:])
            n = group.world * t.numel()
            odt = torch.bfloat16 if role in ("w1", "w2") else t.dtype
            self._op[(role, odt)] = (group.alloc(n * (2 if odt == torch.bfloat16 else 4)), full)
            self._grad[role] = (group.alloc(n * 4), full)

    @staticmethod
    def prefer_weights(num_experts: int, expert_params: int, max_tokens: int, top_k: int, d_in: int, d_out: int) -> bool:
        """True when exchanging the expert weights moves fewer bytes per layer step than dispatching the token rows."""
        weights = (2 + 4) * num_experts * expert_params
        tokens = 2 * max_tokens * top_k * (d_in + d_out) * 2
        return weights < tokens

    def begin_step(self):
        self._have = {}
        self.consumers = 0
        self._prefetched = None

    def _operand_buffer(self, role: str, shard: torch.Tensor, dtype: torch.dtype):
        key = (role, dtype)
        if key not in self._op:       # fp32-accurate mode outside autocast: allocated on first use (collective)
            full = (self.group.world * shard.shape[0], *shard.shape[1:])
            self._op[key] = (self.group.alloc(self.group.world * shard.numel() * (2 if dtype == torch.bfloat16 else 4)), full)
        return self._op[key]

    def _side_stream(self) -> torch.cuda.Stream:
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.group.device)
        return self._side

    def _push(self, shards: dict, op_dtype: torch.dtype) -> bool:
        g = self.group
        pushed = False
        for role, t in shards.items():
            if t is None:
                continue
            dt = op_dtype if role in ("w1", "w2") else t.dtype
            have = self._have.get(role)
            if have is not None and have.dtype == dt:
                continue
            buf, full = self._operand_buffer(role, t, dt)
            src = t.detach().contiguous()
            ops._call("csmoe_ep_gather_push", src.data_ptr(), ops._dt(src), src.numel(), buf.peers(),
                      _lib.BF16 if dt == torch.bfloat16 else _lib.F32, g.rank * src.numel(), g.rank, g.world, ops._stream())
            self._have[role] = buf.tensor(0, full, dt)
            pushed = True
        return pushed

    def prefetch(self, shards: dict, op_dtype: torch.dtype):
        """Start a layer step: publish this rank's shards from a side stream, so that the transfer runs under the
        router GEMM / top-k / routing-map kernels the layer issues next.  operands() joins and runs the barrier."""
        self.begin_step()
        if self.group.world == 1 or not _WX_OVERLAP:
            return
        main, side = torch.cuda.current_stream(self.group.device), self._side_stream()
        side.wait_stream(main)            # the optimizer's update of the shards is ordered on the main stream
        with torch.cuda.stream(side):
            self._prefetched = self._push(shards, op_dtype)

    def operands(self, shards: dict, op_dtype: torch.dtype) -> dict:
        g = self.group
        self.consumers += 1
        pushed = False
        if self._prefetched is not None:
            torch.cuda.current_stream(g.device).wait_stream(self._side_stream())
            pushed, self._prefetched = self._prefetched, None
        pushed = self._push(shards, op_dtype) or pushed
        if pushed:
            g.barrier()
        return {role: (None if t is None else self._have[role]) for role, t in shards.items()}

    def grad_out(self, role: str, shape) -> torch.Tensor:
        buf, full = self._grad[role]
        assert tuple(shape) == tuple(full), f"gradient of {role}: {tuple(shape)} != {tuple(full)}"
        return buf.tensor(0, full, torch.float32)

    def reduce_begin(self, grads: dict, early: bool = False):
        """grads: role -> (full-size gradient or None, parameter shard or None for 'dtype of the gradient').  Gradients
        that were not written into grad_out() already (bias gradients) are copied there.  Barrier, then the owner's
        slices are summed on a side stream: what the caller issues before reduce_end() (the dx reduction) overlaps.
        early=True: the barrier itself runs on the side stream too (its own barrier channel), so the caller's next
        kernels -- the other weight-gradient GEMM -- overlap the wait and the transfer."""
        g = self.group
        todo = []
        for role, (full, param) in grads.items():
            if full is None:
                todo.append(None)
                continue
            buf, shape = self._grad[role]
            dst = buf.tensor(0, shape, torch.float32)
            if full.data_ptr() != dst.data_ptr():
                dst.copy_(full)
            dt = param.dtype if param is not None else full.dtype
            todo.append((buf, torch.empty((shape[0] // g.world, *shape[1:]), dtype=dt, device=g.device)))
        overlap = g.world > 1 and _WX_OVERLAP
        early = early and overlap
        main = torch.cuda.current_stream(g.device)
        if not early:
            g.barrier()
        if overlap:
            side = self._side_stream()
            side.wait_stream(main)
        with torch.cuda.stream(side) if overlap else contextlib.nullcontext():
            if early:
                g.barrier(channel=1)
            for item in todo:
                if item is None:
                    continue
                buf, local = item
                n = local.numel()
                ops._call("csmoe_ep_reduce_pull", buf.peers(), g.rank * n, n, local.data_ptr(), ops._dt(local), g.world, ops._stream())
        return todo, overlap

    def reduce_end(self, *pending):
        """-> the local gradient slices of every reduce_begin() handed in, in order (None where no gradient was given)."""
        g = self.group
        if any(ov for _, ov in pending):
            torch.cuda.current_stream(g.device).wait_stream(self._side_stream())
        if self.consumers > 1:
            g.barrier()     # another function of this layer step writes the same gradient buffers next
        return [None if item is None else item[1] for todo, _ in pending for item in todo]

    def reduce(self, grads: dict):
        return self.reduce_end(self.reduce_begin(grads))


# ------------------------------------------------------------------------------------------------ weights for the dense step
class AllGatherExpertsFn(Function):
    """[E/P, ...] local expert parameters -> [E, ...] on every rank; backward = reduce-scatter of the gradient.
    Used by the (rare) competition step, where every expert sees every token and moving 2*E*F_e bytes of weights is
    cheaper than moving the tokens K-fold."""

    @staticmethod
    def forward(ctx, w_local, group: EPGroup):
        ctx.group = group
        if group.world == 1:
            return w_local
        out = torch.empty((group.world * w_local.shape[0], *w_local.shape[1:]), dtype=w_local.dtype, device=w_local.device)
        dist.all_gather_into_tensor(out, w_local.contiguous(), group=group.group)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        group = ctx.group
        if group.world == 1:
            return g, None
        out = torch.empty((g.shape[0] // group.world, *g.shape[1:]), dtype=g.dtype, device=g.device)
        dist.reduce_scatter_tensor(out, g.contiguous(), op=dist.ReduceOp.SUM, group=group.group)
        return out, None


def gather_experts(w_local: Optional[torch.Tensor], group: EPGroup) -> Optional[torch.Tensor]:
    return None if w_local is None else AllGatherExpertsFn.apply(w_local, group)


# ------------------------------------------------------------------------------------------------ checkpoints under EP
_PRETRAIN_SHARDED = ("keys", "values", "bias")


def full_state_dict(layer, group=None) -> dict:
    """COLLECTIVE over the layer's expert-parallel group.  The layer's state dict in the REFERENCE layout -- multimodal:
    `experts.{global e}.<child>.{weight,bias}`; pretrain: `keys / values / bias` with all E experts along dim 0 -- with
    the expert shards all-gathered, identical on every rank.  This is what a rank-0 `torch.save` must write: after
    `enable_expert_parallel` the plain `state_dict()` holds only this rank's experts, renumbered from 0."""
    g = group or layer._ep.group
    P = g.world
    sd = {k: v.detach() for k, v in layer.state_dict().items()}
    if P == 1:
        return sd

    def gather(t):
        parts = [torch.empty_like(t) for _ in range(P)]
        dist.all_gather(parts, t.contiguous(), group=g.group)
        return parts

    out = {}
    if hasattr(layer, "experts"):                                  # multimodal plugin: one module per local expert
        El = len(layer.experts)
        for k in sorted(sd):
            if not k.startswith("experts."):
                out[k] = sd[k]
                continue
            _, i, rest = k.split(".", 2)
            for r, part in enumerate(gather(sd[k])):
                out[f"experts.{r * El + int(i)}.{rest}"] = part
    else:                                                          # pretrain plugin: stacked [E/P, ...] parameters
        for k in sorted(sd):
            out[k] = torch.cat(gather(sd[k]), dim=0) if k in _PRETRAIN_SHARDED else sd[k]
    return out


def load_full_state_dict(layer, state_dict: dict, group=None, strict: bool = True):
    """Load a reference-layout (all experts) state dict into an expert-parallel layer: every rank keeps its own slice.
    No communication; every rank must be handed the same full dict."""
    g = group or layer._ep.group
    P, r = g.world, g.rank
    if P == 1:
        return layer.load_state_dict(state_dict, strict=strict)
    local = {}
    if hasattr(layer, "experts"):
        El = len(layer.experts)
        for k, v in state_dict.items():
            if k.startswith("experts."):
                _, e, rest = k.split(".", 2)
                if r * El <= int(e) < (r + 1) * El:
                    local[f"experts.{int(e) - r * El}.{rest}"] = v
            else:
                local[k] = v
    else:
        for k, v in state_dict.items():
            if k in _PRETRAIN_SHARDED:
                El = v.shape[0] // P
                local[k] = v[r * El:(r + 1) * El]
            else:
                local[k] = v
    return layer.load_state_dict(local, strict=strict)


# ------------------------------------------------------------------------------------------------ sparse experts, EP
class EPSparseFFNFn(Function):
    """out[t] = sum_k w[t,k] * FFN_{sel[t,k]}(x[t]) with the experts sharded over the group (w1/b1/w2/b2 hold the
    LOCAL experts only, sel indexes GLOBAL experts)."""

    @staticmethod
    def forward(ctx, x, w, sel, w1, b1, w2, b2, spec: FFNSpec, st: EPLayerState):
        g = st.group
        T, K = sel.shape
        assert T <= st.max_tokens and K == st.K, f"EP layer sized for {st.max_tokens} tokens x top-{st.K}, got {T} x {K}"
        xb = _bf16(x)
        w1b, w2b = _bf16(w1), _bf16(w2)
        route = ops.route_build(sel, st.E, row_tile=ROW_TILE)
        plan = g.exchange_plan(route.counts, st.E, st.row_tile, st.row_cap)
        g.dispatch(xb, K, route, plan, st.recv_x.peers(), st.tags.peers())
        g.barrier()
        xp = st.recv_x.tensor(0, (st.row_cap, st.D), torch.bfloat16)
        tags = st.tags.tensor(0, (st.row_cap,), torch.int64)
        c_rows = g.row_ptrs(tags, plan, st.ret_y.peers(), st.Dout, torch.bfloat16, xp)
        lr = plan.local_route(T * K)
        fused = b2 is None and sigma_fused_ok(xb, w1, w2, spec, torch.bfloat16)
        if fused:
            # expert size 128 + ReLU: both projections in one kernel on the received rows (tiled loads: the rows are
            # already expert-major here), then the output rows are pushed to their source ranks
            fr = plan.fused_route(c_rows)
            y_loc, h = ops.sigma_ffn_fwd(xp, w1b, w2b, b1, fr, slots_per_row=1, xp=xp)
            ops._call("csmoe_ep_push_rows", y_loc.data_ptr(), ops.BF16, y_loc.shape[1], y_loc.shape[1], y_loc.shape[0],
                      c_rows.data_ptr(), ops._stream())
            z = h
        else:
            z, h = _ffn_first(xp, w1b, b1, spec, route=lr)
            ops.gemm_rows(h, w2b, w_is_kn=spec.kn_layout, bias=b2, route=lr, c_rows=c_rows)   # output rows land on their source ranks
        g.barrier()
        y = st.ret_y.tensor(0, (T * K, st.Dout), torch.bfloat16)
        ident = g.identity(T * K)
        out = ops.combine_fwd(y, ident, route.sel, w, T, K, round_each=spec.round_each, round_w=spec.round_w)
        ctx.st, ctx.spec, ctx.route, ctx.plan = st, spec, route, plan
        ctx.fused = fused
        ctx.x_dtype, ctx.has_b = x.dtype, (b1 is not None, b2 is not None)
        ctx.save_for_backward(z, h, w, w1, w2)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        z, h, w, w1, w2 = ctx.saved_tensors
        st, spec, route, plan = ctx.st, ctx.spec, ctx.route, ctx.plan
        g = st.group
        T, K, El = route.n_slots // route.top_k, route.top_k, st.El
        w1b, w2b = _bf16(w1), _bf16(w2)
        dout = _bf16(dout.contiguous())
        ident = g.identity(T * K)
        y = st.ret_y.tensor(0, (T * K, st.Dout), torch.bfloat16)
        xp = st.recv_x.tensor(0, (st.row_cap, st.D), torch.bfloat16)
        tags = st.tags.tensor(0, (st.row_cap,), torch.int64)
        dw = ops.combine_bwd_w(y, dout, ident, T, K) if ctx.needs_input_grad[1] else None
        wu = w.to(torch.bfloat16).float() if spec.round_w else w
        recv_dy = g.scratch("recv_dy", st.row_cap * st.Dout * 2)
        ret_dx = g.scratch("ret_dx", st.max_slots * st.D * 2)
        g.dispatch(dout, K, route, plan, recv_dy.peers(), None, slot_w=wu)
        g.barrier()
        dyp = recv_dy.tensor(0, (st.row_cap, st.Dout), torch.bfloat16)
        c_rows = g.row_ptrs(tags, plan, ret_dx.peers(), st.D, torch.bfloat16, dyp)
        lr = plan.local_route(T * K)
        if ctx.fused:
            # fused dgrad on the received (already weighted) upstream rows, d x rows pushed back to their source ranks,
            # weight gradients straight from the receive buffers (identity row map)
            fr = plan.fused_route(c_rows)
            ones = torch.ones(st.row_cap, dtype=torch.float32, device=dyp.device)
            dz, _hw, dxr, _ = ops.sigma_ffn_bwd(dyp, w1b, w2b, fr, ones, h, slots_per_row=1, dyp=dyp)
            dw2 = ops.sigma_wgrad(h, dyp, El, fr, transpose=False, out_dtype=w2.dtype, slots_per_row=1)
            dw1 = ops.sigma_wgrad(dz, xp, El, fr, transpose=True, out_dtype=w1.dtype, slots_per_row=1)
            db1 = ops.bias_grad(dz, El, route=lr, out_dtype=w1.dtype) if ctx.has_b[0] else None
            ops._call("csmoe_ep_push_rows", dxr.data_ptr(), ops.BF16, dxr.shape[1], dxr.shape[1], dxr.shape[0],
                      c_rows.data_ptr(), ops._stream())
            g.barrier()
            dx = None
            if ctx.needs_input_grad[0]:
                dxs = ret_dx.tensor(0, (T * K, st.D), torch.bfloat16)
                dx = ops.scatter_reduce(dxs, ident, T, K).to(ctx.x_dtype)
            return dx, dw, None, dw1, db1, dw2, None, None, None
        db2 = ops.bias_grad(dyp, El, route=lr, out_dtype=w2.dtype) if ctx.has_b[1] else None
        if spec.kn_layout:
            dw2 = ops.gemm_reduce(h, dyp, El, route=lr, out_dtype=w2.dtype)
        else:
            dw2 = ops.gemm_reduce(dyp, h, El, route=lr, out_dtype=w2.dtype)
        db1 = None
        if _FUSE_BWD:
            dz = ops.gemm_rows(dyp, w2b, w_is_kn=not spec.kn_layout, route=lr, act_bwd=spec.act, aux=z)
            db1 = ops.bias_grad(dz, El, route=lr, out_dtype=w1.dtype) if ctx.has_b[0] else None
        else:
            dh = ops.gemm_rows(dyp, w2b, w_is_kn=not spec.kn_layout, route=lr)
            if ctx.has_b[0] and _FUSE_ACT_BIAS and spec.act not in (ops.ACT_NONE, ops.ACT_SILU_GLU):
                dz, db1 = ops.act_bwd_bias(z, dh, spec.act, El, route=lr, out_dtype=w1.dtype)
            else:
                dz = dh if spec.act == ops.ACT_NONE else ops.act_bwd(z, dh, spec.act, lr)
                db1 = ops.bias_grad(dz, El, route=lr, out_dtype=w1.dtype) if ctx.has_b[0] else None
        if spec.kn_layout:
            dw1 = ops.gemm_reduce(xp, dz, El, route=lr, out_dtype=w1.dtype)
        else:
            dw1 = ops.gemm_reduce(dz, xp, El, route=lr, out_dtype=w1.dtype)
        # dx rows go straight back to their source ranks (always executed: the barrier sequence must match on all ranks)
        ops.gemm_rows(dz, w1b, w_is_kn=not spec.kn_layout, route=lr, c_rows=c_rows)
        g.barrier()
        dx = None
        if ctx.needs_input_grad[0]:
            dxs = ret_dx.tensor(0, (T * K, st.D), torch.bfloat16)
            dx = ops.scatter_reduce(dxs, ident, T, K).to(ctx.x_dtype)
        return dx, dw, None, dw1, db1, dw2, db2, None, None
