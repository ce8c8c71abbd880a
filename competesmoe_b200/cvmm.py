"""Drop-in for the CVMM op boundary (reference: moe_pretrain_model/layers/cvmm.py).

Same public names and argument meaning -- `CVMMSel`, `cvmm_prepare_sel`, `cvmm_prepare_sel2`, `cvmm(x, sel, keys)` --
but the sort is a stable counting sort on the GPU (csmoe_route_build) and the conditional matmul is the tcgen05 grouped
GEMM over TMA-staged, expert-major rows instead of Triton pointer gathers; the weight gradient is a deterministic
grouped GEMM instead of split-K fp32 atomics (cvmm.py:194-345).

    cvmm(x, sel, keys)[out_index[i]] = x[sel_index[i]] @ keys[sel.sel[i]]      (+ weighted reduction over K, :481-483)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Union

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import ops


@dataclass
class CVMMSel:
    raw_sel: torch.Tensor
    sel: torch.Tensor
    sel_index: torch.Tensor
    out_index: Optional[torch.Tensor] = None
    reduction_weight: Optional[torch.Tensor] = None
    _route: Optional[ops.Route] = None   # permutation maps of the padded expert-major space (built once per selection)
    _checked: Optional[tuple] = None     # identity of the (sel_index, out_index) pair last verified against _route

    def clone(self) -> "CVMMSel":
        return CVMMSel(self.raw_sel, self.sel, self.sel_index, self.out_index, self.reduction_weight, self._route,
                       self._checked)


# The grouped GEMM addresses rows through the maps of `_route` (built from raw_sel), not through sel_index / out_index.
# Those two tensors are public and callers do rewrite them (competesmoe.py:516-521, full_moe_relative_attention.py:453-458),
# so before they are ignored they are verified to be one of the layouts the maps can express:
#   out_index is None : sel_index == sort_index                 (every slot reads its own input row, output in slot order)
#   out_index given   : out_index == sort_index, sel_index == sort_index // c   (c slots share one input row)
# Anything else raises NotImplementedError instead of silently computing the wrong rows.  The check costs one
# device->host sync per distinct (sel_index, out_index) pair; it is skipped while a CUDA graph is being captured and can
# be switched off with `competesmoe_b200.cvmm.CHECK_SEL_INDEX = False` once a call site is known to be good.
CHECK_SEL_INDEX = True


def _ident(t: Optional[torch.Tensor]):
    return None if t is None else (t.data_ptr(), t._version, tuple(t.shape))


def _verify_index_layout(sel: CVMMSel, slots_per_row: int) -> None:
    if not CHECK_SEL_INDEX or torch.cuda.is_current_stream_capturing():
        return
    key = (_ident(sel.sel_index), _ident(sel.out_index), slots_per_row)
    if sel._checked == key:
        return
    si = sel._route.sort_index
    ok = sel.sel_index is not None and sel.sel_index.numel() == si.numel()
    if ok and sel.out_index is not None:
        ok = sel.out_index.numel() == si.numel() and bool(
            (sel.out_index.reshape(-1) == si).all() & (sel.sel_index.reshape(-1) == si // slots_per_row).all())
    elif ok:
        ok = bool((sel.sel_index.reshape(-1) == si).all())
    if not ok:
        raise NotImplementedError(
            "cvmm: sel_index / out_index are not the stable-sort maps of raw_sel (`pos // c` with out_index = pos, or "
            "sel_index = pos with out_index = None); arbitrary index tensors are not supported by the grouped-GEMM path")
    sel._checked = key


def _num_experts_hint(sel: torch.Tensor, n_experts: Optional[int]) -> int:
    if n_experts is None:
        raise ValueError("the number of experts is needed to build the routing maps")
    return n_experts


def cvmm_prepare_sel(sel: torch.Tensor, n_experts: int) -> CVMMSel:
    """cvmm.py:23-26: one selection per row."""
    route = ops.route_build(sel.reshape(-1, 1), n_experts)
    return CVMMSel(sel, route.sorted_sel.view_as(sel), route.sort_index, None, None, route,
                   (_ident(route.sort_index), None, 1))


def cvmm_prepare_sel2(sel: torch.Tensor, w: Optional[torch.Tensor] = None, n_experts: Optional[int] = None) -> CVMMSel:
    """cvmm.py:580-592: K selections per row; sel_index = sorted position // K, out_index = sorted position.
    The reference infers nothing about E here; pass `n_experts` or let `cvmm` rebuild the maps from `keys.shape[0]`."""
    k = sel.shape[-1]
    if n_experts is None:
        return CVMMSel(sel, None, None, None, w, None)  # completed lazily by cvmm() once E is known
    route = ops.route_build(sel.reshape(-1, k), n_experts)
    in_index = route.sort_index // k
    return CVMMSel(sel, route.sorted_sel.view_as(sel), in_index, route.sort_index, w, route,
                   (_ident(in_index), _ident(route.sort_index), k))   # built here: correct by construction


def _complete(sel: CVMMSel, n_experts: int) -> CVMMSel:
    if sel._route is not None and sel._route.num_experts == n_experts:
        return sel
    k = sel.raw_sel.shape[-1]
    route = ops.route_build(sel.raw_sel.reshape(-1, k), n_experts)
    sel._route = route
    if sel.sel is None:
        sel.sel = route.sorted_sel.view_as(sel.raw_sel)
        sel.sel_index = route.sort_index // k
        sel.out_index = route.sort_index
        sel._checked = (_ident(sel.sel_index), _ident(sel.out_index), k)
    return sel


def _out_dtype(x: torch.Tensor) -> torch.dtype:
    """cvmm.py:29-32 get_dtype(): the autocast dtype when autocast is active, else fp32 -- whatever the input dtype."""
    if torch.is_autocast_enabled():
        return torch.get_autocast_dtype('cuda')
    return torch.float32


class CVMM(Function):
    """cvmm.py:460-551."""

    @staticmethod
    def forward(ctx, x, keys, route: ops.Route, slots_per_row: int, reduction_weight, out_dtype):
        x2 = x.flatten(end_dim=-2)
        # Outside autocast the reference computes in fp32 (get_dtype(), cvmm.py:29-32; allow_tf32=False, :395): fp32
        # operands then take the fp32-accurate products (six split-bf16 tensor-core GEMMs, fp32 accumulation) instead of
        # one bf16 product -- rtol 1e-4 against the reference instead of 1e-2.
        ctx.fp32 = fp32 = out_dtype == torch.float32 and (x2.dtype == torch.float32 or keys.dtype == torch.float32)
        if fp32:
            xp = ops.gather_rows(x2.float(), route, slots_per_src_row=slots_per_row)
            yp = ops.gemm_rows_f32(xp, keys.float(), w_is_kn=True, route=route)
        else:
            xb = ops.cast_bf16(x2) if x2.dtype != torch.bfloat16 else x2
            kb = ops.cast_bf16(keys) if keys.dtype != torch.bfloat16 else keys
            xp = ops.gather_rows(xb, route, slots_per_src_row=slots_per_row)
            yp = ops.gemm_rows(xp, kb, w_is_kn=True, route=route, out_dtype=out_dtype)
        n_slots = route.n_slots
        if reduction_weight is None:
            out = ops.scatter_reduce(yp, route.slot_to_row, n_slots, 1)           # back to slot order, no reduction
        else:
            K = reduction_weight.shape[-1]           # slots reduced into one output row (top-k, or heads x top-k)
            out = ops.combine_fwd(yp, route.slot_to_row, route.sel, reduction_weight.float(), n_slots // K, K,
                                  round_w=out_dtype == torch.bfloat16)
        ctx.route, ctx.slots_per_row = route, slots_per_row
        ctx.x_shape, ctx.x_dtype, ctx.out_dtype = x.shape, x.dtype, out_dtype
        ctx.save_for_backward(xp, yp, keys, reduction_weight)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        xp, yp, keys, rw = ctx.saved_tensors
        route = ctx.route
        n_slots, E = route.n_slots, route.num_experts
        K = rw.shape[-1] if rw is not None else 1
        g2 = g.reshape(-1, g.shape[-1]).contiguous()
        if ctx.fp32:
            gb, kb = g2.float(), keys.float()
        else:
            gb = ops.cast_bf16(g2) if g2.dtype != torch.bfloat16 else g2
            kb = ops.cast_bf16(keys) if keys.dtype != torch.bfloat16 else keys
        drw = None
        if rw is None:
            gp = ops.gather_rows(gb, route, slots_per_src_row=1)
        else:
            ypb = yp if yp.dtype == gb.dtype else yp.to(gb.dtype)
            drw = ops.combine_bwd_w(ypb, gb, route.slot_to_row, n_slots // K, K).view_as(rw).to(rw.dtype)
            wv = rw.float()
            if ctx.out_dtype == torch.bfloat16:
                wv = wv.bfloat16().float()
            gp = ops.gather_rows(gb, route, slot_w=wv, slots_per_src_row=K)
        if ctx.fp32:
            dkeys = ops.gemm_reduce_f32(xp, gp, E, route=route).to(keys.dtype)
        else:
            dkeys = ops.gemm_reduce(xp, gp, E, route=route, out_dtype=keys.dtype)    # [E, k_in, n]
        dx = None
        if ctx.needs_input_grad[0]:
            dxp = ops.gemm_rows_f32(gp, kb, w_is_kn=False, route=route) if ctx.fp32 else \
                ops.gemm_rows(gp, kb, w_is_kn=False, route=route)                   # gp @ keys^T
            rows_in = n_slots // ctx.slots_per_row
            dx = ops.scatter_reduce(dxp, route.slot_to_row, rows_in, ctx.slots_per_row).view(ctx.x_shape).to(ctx.x_dtype)
        return dx, dkeys, None, None, drw, None


def cvmm(x: torch.Tensor, sel: Union[torch.Tensor, CVMMSel], keys: torch.Tensor) -> torch.Tensor:
    """cvmm.py:555-577.  `keys` is [E, k_in, n]."""
    if not isinstance(sel, CVMMSel):
        sel = cvmm_prepare_sel(sel, keys.shape[0])
    sel = _complete(sel, keys.shape[0])
    # prepare_sel2 pattern: sel_index = pos // K with out_index = pos (K slots share one input row);
    # after `sel_index = out_index; out_index = None` (or prepare_sel) every slot has its own input row.
    # MoE attention (full_moe_relative_attention.py:453-458) re-divides: sel_index = out_index // k on per-head inputs
    # with the reduction weight flattened over (heads, k) -- any uniform `pos // c` pattern is recovered from the
    # number of input rows, and the reduction group is the last dimension of reduction_weight.
    n_slots = sel._route.n_slots
    if sel.out_index is None:
        slots_per_row = 1
    else:
        rows_in = x.numel() // x.shape[-1]
        if rows_in == 0 or n_slots % rows_in != 0:
            raise ValueError(f"cvmm: {rows_in} input rows do not divide the {n_slots} selection slots")
        slots_per_row = n_slots // rows_in
    _verify_index_layout(sel, slots_per_row)
    rw = sel.reduction_weight
    if rw is not None and (rw.shape[-1] > 8 or n_slots % rw.shape[-1] != 0):
        raise ValueError(f"cvmm: reduction over {rw.shape[-1]} slots per output row is not supported (1..8, dividing {n_slots})")
    out = CVMM.apply(x, keys, sel._route, slots_per_row, rw, _out_dtype(x))
    if sel.reduction_weight is None:
        return out.view(*sel.raw_sel.shape, keys.shape[-1])
    return out.view(*sel.reduction_weight.shape[:-1], keys.shape[-1])


# ------------------------------------------------------------------------------------------------ torch.library op
# The reference also exposes the forward kernel launcher as a dispatcher op so that torch.compile can trace through it
# (cvmm.py:348-416):  mylib::cvmm_triton(Tensor x, Tensor sel_index, Tensor sel, Tensor keys, ScalarType out_dtype,
# Tensor out_index) -> Tensor, with   out[out_index[i]] = x[sel_index[i]] @ keys[sel[i]]   for the SORTED expert ids
# `sel` (out_index = tensor(-1): row i of the output is sorted row i).  This entry point takes the index tensors as
# they are -- any sel_index / out_index, not only the two layouts `cvmm()` recognises -- because the rows are first
# brought into sorted order with the caller's own indices and only then handed to the grouped GEMM.
CVMM_TRITON_SCHEMA = "(Tensor x, Tensor sel_index, Tensor sel, Tensor keys, ScalarType out_dtype, Tensor out_index) -> Tensor"


def cvmm_triton(x: torch.Tensor, sel_index: torch.Tensor, sel: torch.Tensor, keys: torch.Tensor, out_dtype: torch.dtype,
                out_index: torch.Tensor) -> torch.Tensor:
    x2 = x.flatten(end_dim=-2)
    assert x2.shape[-1] == keys.shape[1]
    sel_shape = sel.shape
    fsel = sel.flatten()
    M, (E, _, N) = fsel.shape[0], keys.shape
    xs = x2.index_select(0, sel_index.flatten().long())                       # sorted row i <- x[sel_index[i]]
    route = ops.route_build(fsel.to(torch.int32).view(-1, 1), E)              # already sorted: slot i is sorted row i
    if out_dtype == torch.float32 and (xs.dtype == torch.float32 or keys.dtype == torch.float32):
        # the reference's kernel converts the operands to out_dtype (cvmm.py:389-395): fp32 in, fp32 arithmetic
        yp = ops.gemm_rows_f32(ops.gather_rows(xs.float().contiguous(), route, slots_per_src_row=1), keys.float(),
                               w_is_kn=True, route=route)
    else:
        xb = ops.cast_bf16(xs) if xs.dtype != torch.bfloat16 else xs.contiguous()
        kb = ops.cast_bf16(keys) if keys.dtype != torch.bfloat16 else keys
        xp = ops.gather_rows(xb, route, slots_per_src_row=1)
        yp = ops.gemm_rows(xp, kb, w_is_kn=True, route=route,
                           out_dtype=out_dtype if out_dtype in (torch.bfloat16, torch.float32) else torch.float32)
    ys = ops.scatter_reduce(yp, route.slot_to_row, M, 1)
    if ys.dtype != out_dtype:
        ys = ys.to(out_dtype)
    # out_index "is None" is spelled tensor(-1) (cvmm.py:385-387); a 1-element index with M > 1 can only be that marker
    if out_index.numel() == 1 and (M != 1 or int(out_index) == -1):
        out = ys
    else:
        out = torch.empty_like(ys).index_copy_(0, out_index.flatten().long(), ys)
    return out.view(*sel_shape, N)


def _cvmm_triton_fake(x, sel_index, sel, keys, out_dtype, out_index):
    return torch.empty((*sel.shape, keys.shape[-1]), device=x.device, dtype=out_dtype)


def _register_library_op():
    """Define mylib::cvmm_triton unless the reference's own module already did (both in one process: the reference's
    definition stays, ours is reachable as csmoe::cvmm_triton and through `cvmm_triton_call`)."""
    lib = torch.library
    for ns in ("mylib", "csmoe"):
        try:
            lib.define(f"{ns}::cvmm_triton", CVMM_TRITON_SCHEMA)
        except RuntimeError:
            continue                                   # already defined in this process
        lib.impl(f"{ns}::cvmm_triton", "CUDA")(cvmm_triton)
        (getattr(lib, "register_fake", None) or lib.impl_abstract)(f"{ns}::cvmm_triton")(_cvmm_triton_fake)
        return getattr(getattr(torch.ops, ns), "cvmm_triton")
    return cvmm_triton


cvmm_triton_call = _register_library_op()
