"""Tensor-level wrappers over the C ABI (include/csmoe.h).

Each function takes CUDA torch tensors, allocates outputs with torch (PyTorch owns device memory) and launches the
corresponding libcsmoe kernel on the current stream.  No computation happens here and nothing falls back to PyTorch or
the CPU: a non-CUDA tensor or a failing call raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import (ACT_GELU, ACT_GELU_TANH, ACT_NONE, ACT_RELU, ACT_SILU, ACT_SILU_GLU, BF16, F32, GEMM_REDUCE,
                   GEMM_ROWS, ROW_TILE, GemmArgs, check)

__all__ = [
    "Route", "route_build", "router_fwd", "router_from_logits", "topk_renorm", "gather_rows", "combine_fwd", "combine_bwd_w",
    "scatter_reduce", "gemm_rows", "gemm_reduce", "act_fwd", "act_bwd", "act_bwd_bias", "bias_grad", "cast_bf16", "affinity_fwd",
    "affinity_bwd", "affinity_from_rowsum", "diversity_fwd", "compete_bwd", "ACT_NONE", "ACT_RELU", "ACT_GELU", "ACT_GELU_TANH", "ACT_SILU", "ACT_SILU_GLU",
]

launch_count = 0  # number of libcsmoe kernels launched so far (bench.py reports the delta over its timed region)
gemm_timing = None  # when set to a list, every grouped GEMM is bracketed by CUDA events: (start, end, flops, tag)


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    raise TypeError(f"libcsmoe supports bfloat16 and float32 tensors, got {t.dtype}")


def _cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("libcsmoe kernels need CUDA tensors; there is no CPU path in this package")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


_NO_STORAGE = {}


def _p(t: Optional[torch.Tensor]):
    """Device address of a tensor.  A tensor with zero elements has no storage (data_ptr() == 0), which the library
    would reject as a NULL pointer although every entry point returns early on zero sizes: such a tensor gets a valid
    address that is never dereferenced (a rank without tokens, an empty micro-batch)."""
    if t is None:
        return None
    p = t.data_ptr()
    if p == 0:
        d = _NO_STORAGE.get(t.device)
        if d is None:
            d = _NO_STORAGE[t.device] = torch.zeros(64, dtype=torch.uint8, device=t.device)
        return d.data_ptr()
    return p


def _call(name: str, *args, kernels: int = 1) -> None:
    global launch_count
    launch_count += kernels
    check(getattr(_lib.load(), name)(*args), name)


# ----------------------------------------------------------------------------------------------- routing metadata
@dataclass
class Route:
    """Permutation maps between slot order (t*K + k) and the padded expert-major row space."""
    num_experts: int
    top_k: int
    n_slots: int
    row_cap: int
    sel: torch.Tensor          # [n_slots] int32 expert of each slot
    counts: torch.Tensor       # [E] int32
    offsets: torch.Tensor      # [E+1] int32 (exclusive scan of counts)
    pad_offsets: torch.Tensor  # [E+1] int32 (segment starts, multiples of ROW_TILE)
    sorted_sel: torch.Tensor   # [n_slots] int32  (reference: ssel)
    sort_index: torch.Tensor   # [n_slots] int64  (reference: out_index; in_index = sort_index // K)
    slot_to_row: torch.Tensor  # [n_slots] int32
    row_to_slot: torch.Tensor  # [row_cap] int32, -1 on padding rows
    tile_expert: torch.Tensor  # [row_cap / ROW_TILE] int32, -1 past the end
    row_tile: int = ROW_TILE   # alignment of the expert segments (256 lets the CTA-pair GEMM run)


def route_row_cap(n_slots: int, num_experts: int, row_tile: int = ROW_TILE) -> int:
    return int(_lib.load().csmoe_route_row_cap(n_slots, num_experts, row_tile))


def default_row_tile(n_slots: int, num_experts: int) -> int:
    """256-row expert segments (CTA-pair GEMM tiles) unless the padding would cost more than ~6% extra rows."""
    return 256 if num_experts * 128 * 16 <= max(n_slots, 1) else ROW_TILE


def route_build(sel: torch.Tensor, num_experts: int, row_tile: Optional[int] = None) -> Route:
    """sel: [T, K] (or flat) integer expert ids on CUDA."""
    _cuda(sel)
    top_k = sel.shape[-1] if sel.dim() > 1 else 1
    flat = sel.reshape(-1)
    if flat.dtype != torch.int32:
        flat = flat.to(torch.int32)
    flat = flat.contiguous()
    n = flat.numel()
    dev = flat.device
    lib = _lib.load()
    if row_tile is None:
        row_tile = default_row_tile(n, num_experts)
    row_cap = int(lib.csmoe_route_row_cap(n, num_experts, row_tile))
    ws = torch.empty(max(int(lib.csmoe_route_workspace_bytes(n, num_experts)) // 4, 1), dtype=torch.int32, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    counts = torch.empty(num_experts, **i32)
    offsets = torch.empty(num_experts + 1, **i32)
    pad_offsets = torch.empty(num_experts + 1, **i32)
    sorted_sel = torch.empty(n, **i32)
    sort_index = torch.empty(n, dtype=torch.int64, device=dev)
    slot_to_row = torch.empty(n, **i32)
    row_to_slot = torch.empty(row_cap, **i32)
    tile_expert = torch.empty(row_cap // ROW_TILE, **i32)
    _call("csmoe_route_build", _p(flat), n, num_experts, row_tile, row_cap, _p(counts), _p(offsets), _p(pad_offsets),
          _p(sorted_sel), _p(sort_index), _p(slot_to_row), _p(row_to_slot), _p(tile_expert), _p(ws), _stream(),
          kernels=3 if n > 0 else 2)
    return Route(num_experts, top_k, n, row_cap, flat, counts, offsets, pad_offsets, sorted_sel, sort_index,
                 slot_to_row, row_to_slot, tile_expert, row_tile)


# ----------------------------------------------------------------------------------------------- router
_ROUTER_GEMM = __import__("os").environ.get("CSMOE_ROUTER_GEMM", "1") != "0"


def _router_gemm_ok(T: int, D: int, E: int, dtype: torch.dtype) -> bool:
    """Gate GEMM on the tensor cores: worth it (and legal for the grouped GEMM) from ~16 experts on bf16 activations."""
    return _ROUTER_GEMM and dtype == torch.bfloat16 and E >= 16 and E % 8 == 0 and T >= 256 and T % ROW_TILE == 0 and D % 64 == 0


def _renorm_dt(renorm_dtype: Optional[torch.dtype], default: torch.dtype) -> int:
    """dtype the reference rounds the routing-weight denominator to (`.to(x.dtype)`, x = the layer's input)."""
    rd = default if renorm_dtype is None else renorm_dtype
    return BF16 if rd == torch.bfloat16 else F32


def router_fwd(x: torch.Tensor, wg: torch.Tensor, top_k: int, renorm_dtype: Optional[torch.dtype] = None):
    """x [T, D], wg [E, D] (same dtype) -> logits [T,E] (x dtype), probs [T,E] f32, topk_w [T,K] f32, topk_idx [T,K] i32.
    renorm_dtype: dtype of the layer input whose `.to(x.dtype)` rounds the top-k sum (default: x's own dtype; fp32 for
    fp32 inputs that were cast to bf16 under autocast)."""
    _cuda(x, wg)
    assert x.dim() == 2 and wg.dim() == 2 and x.shape[1] == wg.shape[1] and x.dtype == wg.dtype
    x, wg = x.contiguous(), wg.contiguous()
    T, D = x.shape
    E = wg.shape[0]
    probs = torch.empty(T, E, dtype=torch.float32, device=x.device)
    tw = torch.empty(T, top_k, dtype=torch.float32, device=x.device)
    ti = torch.empty(T, top_k, dtype=torch.int32, device=x.device)
    if _router_gemm_ok(T, D, E, x.dtype):
        # many experts: logits = x . Wg^T as a one-"expert" dense grouped GEMM, then softmax / top-k from the logits
        logits = gemm_rows(x, wg.unsqueeze(0), w_is_kn=False, dense_rows=T, a_expert_rows=0)
        return (logits,) + router_from_logits(logits, top_k, out=(probs, tw, ti), renorm_dtype=renorm_dtype)
    logits = torch.empty(T, E, dtype=x.dtype, device=x.device)
    _call("csmoe_router_fwd", _p(x), _p(wg), _dt(x), T, D, E, top_k, _renorm_dt(renorm_dtype, x.dtype), _p(logits), _p(probs),
          _p(tw), _p(ti), _stream())
    return logits, probs, tw, ti


def router_from_logits(logits: torch.Tensor, top_k: int, out=None, renorm_dtype: Optional[torch.dtype] = None):
    """logits [T, E] (activation dtype, already rounded) -> probs [T,E] f32, topk_w [T,K] f32, topk_idx [T,K] i32: the
    softmax / top-k / renormalisation half of router_fwd, bit-identical to it on the same logits."""
    _cuda(logits)
    logits = logits.contiguous()
    T, E = logits.shape
    if out is None:
        out = (torch.empty(T, E, dtype=torch.float32, device=logits.device),
               torch.empty(T, top_k, dtype=torch.float32, device=logits.device),
               torch.empty(T, top_k, dtype=torch.int32, device=logits.device))
    probs, tw, ti = out
    _call("csmoe_router_from_logits", _p(logits), _dt(logits), T, E, top_k, _renorm_dt(renorm_dtype, logits.dtype), _p(probs),
          _p(tw), _p(ti), _stream())
    return probs, tw, ti


def router_aux_fwd(logits: torch.Tensor, probs: torch.Tensor, topk_idx: torch.Tensor, batch: int):
    """Balance + z losses of the router step in one pass.  Returns (losses [2] f32 = (balance, z), cnt [B,E], lse [T])."""
    _cuda(logits, probs, topk_idx)
    T, E = probs.shape
    K = topk_idx.shape[1]
    assert T % batch == 0
    N = T // batch
    dev = probs.device
    lib = _lib.load()
    ws = torch.empty(int(lib.csmoe_router_aux_workspace_bytes(batch, N, E)) // 4, dtype=torch.float32, device=dev)
    psum = torch.empty(batch, E, dtype=torch.float32, device=dev)
    cnt = torch.empty(batch, E, dtype=torch.float32, device=dev)
    lse = torch.empty(T, dtype=torch.float32, device=dev)
    losses = torch.empty(2, dtype=torch.float32, device=dev)
    _call("csmoe_router_aux_fwd", _p(logits), _dt(logits), _p(probs), _p(topk_idx), batch, N, E, K, _p(psum), _p(cnt),
          _p(lse), _p(losses), _p(ws), _stream(), kernels=2)
    return losses, cnt, lse


def router_bwd(x: torch.Tensor, wg: torch.Tensor, probs: torch.Tensor, topk_w: torch.Tensor, topk_idx: torch.Tensor,
               batch: int, *, dtw=None, dprobs=None, dlogits=None, lse=None, cnt=None, g_losses=None,
               need_dx: bool = True, need_dwg: bool = True, wg_dtype: Optional[torch.dtype] = None,
               renorm_dtype: Optional[torch.dtype] = None):
    """Fused router backward -> (dx [T,D] x.dtype or None, dwg [E,D] or None)."""
    _cuda(x, wg, probs, topk_w, topk_idx, dtw, dprobs, dlogits, lse, cnt, g_losses)
    T, D = x.shape
    E = wg.shape[0]
    K = topk_idx.shape[1]
    N = T // batch
    dev = x.device
    f32 = lambda t: None if t is None else t.contiguous().float()  # noqa: E731
    dtw, dprobs, dlogits, g_losses = f32(dtw), f32(dprobs), f32(dlogits), f32(g_losses)
    wg_dtype = wg_dtype or wg.dtype
    rdt = _renorm_dt(renorm_dtype, x.dtype)
    dl = torch.empty(T, E, dtype=torch.float32, device=dev)
    if _router_gemm_ok(T, D, E, x.dtype) and wg_dtype in (torch.bfloat16, torch.float32):
        # many experts: only d logits comes from the fused kernel; dx = dl . Wg and dWg = dl^T . x run on the tensor
        # cores with dl rounded to the activation dtype (what autograd hands the reference's bf16 gate Linear)
        _call("csmoe_router_bwd", _p(x), _p(wg), _dt(x), _p(probs), _p(topk_w), _p(topk_idx), _p(dtw), _p(dprobs),
              _p(dlogits), _p(lse), _p(cnt), _p(g_losses), batch, N, D, E, K, rdt, _p(dl), None, None, F32, None, _stream())
        dlb = dl.to(x.dtype)
        dx = gemm_rows(dlb, wg.unsqueeze(0), w_is_kn=True, dense_rows=T, a_expert_rows=0) if need_dx else None
        dwg = None
        if need_dwg:
            # dWg [E, D] = dl^T . x contracts over all T tokens into ONE 64..-row output block: a handful of tiles.  Split
            # the tokens into S row ranges ("experts" of a dense REDUCE launch), one partial [E, D] each, summed in a
            # fixed order afterwards (deterministic): 53 -> ~8 us at T = 8192, E = 64, D = 1024.
            S = 1
            while S < 32 and T % (2 * S * ROW_TILE) == 0 and T // (2 * S) >= 2 * ROW_TILE:
                S *= 2
            if S == 1:
                dwg = gemm_reduce(dlb, x, 1, dense_rows=T, out_dtype=wg_dtype)[0]
            else:
                part = gemm_reduce(dlb, x, S, dense_rows=T // S, a_expert_rows=T // S, b_expert_rows=T // S,
                                   out_dtype=torch.float32)
                dwg = part.sum(0).to(wg_dtype)
        return dx, dwg
    dx = torch.empty(T, D, dtype=x.dtype, device=dev) if need_dx else None
    dwg = torch.empty(E, D, dtype=wg_dtype, device=dev) if need_dwg else None
    ws = None
    if need_dwg:
        ws = torch.empty(int(_lib.load().csmoe_router_bwd_workspace_bytes(T, D, E)) // 4, dtype=torch.float32, device=dev)
    _call("csmoe_router_bwd", _p(x), _p(wg), _dt(x), _p(probs), _p(topk_w), _p(topk_idx), _p(dtw), _p(dprobs),
          _p(dlogits), _p(lse), _p(cnt), _p(g_losses), batch, N, D, E, K, rdt, _p(dl), _p(dx), _p(dwg),
          _dt(dwg) if dwg is not None else F32, _p(ws), _stream(), kernels=3 if need_dwg else 1)
    return dx, dwg


def topk_renorm(scores: torch.Tensor, top_k: int, sigmoid: bool = False, round_dtype: torch.dtype = torch.float32,
                round_out: bool = False):
    """scores [T, E] f32 -> (w [T,K] f32, idx [T,K] i32); w = topk / round(sum topk)."""
    _cuda(scores)
    scores = scores.contiguous()
    assert scores.dtype == torch.float32 and scores.dim() == 2
    T, E = scores.shape
    tw = torch.empty(T, top_k, dtype=torch.float32, device=scores.device)
    ti = torch.empty(T, top_k, dtype=torch.int32, device=scores.device)
    mode = (1 if sigmoid else 0) | (2 if round_out else 0)
    rd = BF16 if round_dtype == torch.bfloat16 else F32
    _call("csmoe_topk_renorm", _p(scores), T, E, top_k, mode, rd, _p(tw), _p(ti), _stream())
    return tw, ti


# ----------------------------------------------------------------------------------------------- permutation
def gather_rows(src: torch.Tensor, route: Route, slot_w: Optional[torch.Tensor] = None,
                slots_per_src_row: Optional[int] = None) -> torch.Tensor:
    """src [T, D] -> [row_cap, D] in the padded expert-major space (zeros on padding rows).
    Row r reads src[row_to_slot[r] // slots_per_src_row] (default: the route's top_k, i.e. src is token-major)."""
    _cuda(src, slot_w)
    src = src.contiguous()
    T, D = src.shape
    k_div = route.top_k if slots_per_src_row is None else slots_per_src_row
    assert T * k_div == route.n_slots, f"gather_rows: {T} source rows x {k_div} != {route.n_slots} slots"
    dst = torch.empty(route.row_cap, D, dtype=src.dtype, device=src.device)
    if slot_w is not None:
        slot_w = slot_w.reshape(-1).contiguous()
        assert slot_w.dtype == torch.float32 and slot_w.numel() == route.n_slots
    _call("csmoe_gather_rows", _p(src), _dt(src), T, D, k_div, _p(route.row_to_slot), route.row_cap, _p(slot_w),
          _p(dst), _stream())
    return dst


def combine_fwd(y: torch.Tensor, slot_to_row: torch.Tensor, sel: torch.Tensor, w: torch.Tensor, T: int, top_k: int,
                round_each: bool = False, round_w: bool = False) -> torch.Tensor:
    _cuda(y, slot_to_row, sel, w)
    D = y.shape[-1]
    out = torch.empty(T, D, dtype=y.dtype, device=y.device)
    w = w.reshape(-1).contiguous()
    assert w.dtype == torch.float32 and slot_to_row.dtype == torch.int32 and sel.dtype == torch.int32
    flags = (1 if round_each else 0) | (2 if round_w else 0)
    _call("csmoe_combine_fwd", _p(y), _dt(y), T, D, top_k, _p(slot_to_row), _p(sel), _p(w), flags, _p(out), _stream())
    return out


def combine_bwd_w(y: torch.Tensor, dout: torch.Tensor, slot_to_row: torch.Tensor, T: int, top_k: int) -> torch.Tensor:
    _cuda(y, dout, slot_to_row)
    dout = dout.contiguous()
    assert y.dtype == dout.dtype
    D = y.shape[-1]
    dw = torch.empty(T, top_k, dtype=torch.float32, device=y.device)
    _call("csmoe_combine_bwd_w", _p(y), _p(dout), _dt(y), T, D, top_k, _p(slot_to_row), _p(dw), _stream())
    return dw


def scatter_reduce(g: torch.Tensor, slot_to_row: torch.Tensor, T: int, top_k: int,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dx[t] = sum_k g[row(t,k)]; adds into `out` when given."""
    _cuda(g, slot_to_row, out)
    D = g.shape[-1]
    acc = out is not None
    if out is None:
        out = torch.empty(T, D, dtype=g.dtype, device=g.device)
    _call("csmoe_scatter_reduce", _p(g), _dt(g), T, D, top_k, _p(slot_to_row), 1 if acc else 0, _p(out), _stream())
    return out


# ----------------------------------------------------------------------------------------------- grouped GEMM
def _gemm(args: GemmArgs, flops: float = 0.0, tag: str = "") -> None:
    if gemm_timing is None:
        _call("csmoe_grouped_gemm", C.byref(args), _stream())
        return
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    _call("csmoe_grouped_gemm", C.byref(args), _stream())
    end.record()
    gemm_timing.append((start, end, flops, tag))


def gemm_rows(a: torch.Tensor, w: torch.Tensor, *, w_is_kn: bool, route: Optional[Route] = None,
              dense_rows: int = 0, a_expert_rows: int = 0, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
              want_preact: bool = False, out_dtype: Optional[torch.dtype] = None, act_bwd: int = ACT_NONE,
              aux: Optional[torch.Tensor] = None, c_rows: Optional[torch.Tensor] = None, sum_experts: bool = False,
              rowsum_softplus: Optional[bool] = None, bias_after_round: bool = False,
              accumulate_into: Optional[torch.Tensor] = None):
    """C[row] = A[row] . W[expert(row)] with a fused epilogue.

    a: [rows, k] bf16.  w: [E, n, k] (w_is_kn=False, nn.Linear layout) or [E, k, n] (w_is_kn=True).
    route given: rows are the padded expert-major space.  dense_rows > 0: every expert processes `dense_rows` rows of
    `a` (shared when a_expert_rows == 0) and C is [E * dense_rows, n].
    Forward epilogue: + bias, activation; want_preact also returns the pre-activation.  act = ACT_SILU_GLU: w is
    [E, 2F, k]; returns (h [rows, F], z [rows, 2F]).
    Backward epilogue (act_bwd, aux = saved z): C = (A . W) * act'(z); ACT_SILU_GLU returns dz [rows, 2F] from dh [rows, F].
    sum_experts (dense, a_expert_rows > 0): C [dense_rows, n] = sum_e A[e] . W[e] in one launch (k loop over experts).
    rowsum_softplus (None = off; True / False = round every softplus to bf16 or not): additionally returns
    rowsum [rows, ceil(n / 64)] f32, the per-64-column sums of softplus(C) (the competition score, see affinity_from_rowsum).
    c_rows [rows] int64 (expert-parallel return): output row r is stored at address c_rows[r] (0 = skipped) instead of a
    local C; nothing is returned.
    Returns C, or (C, preact) when want_preact.
    """
    _cuda(a, w, bias, aux)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16, "grouped GEMM operands must be bfloat16"
    assert a.dim() == 2 and w.dim() == 3 and a.stride(1) == 1 and w.stride(2) == 1
    E = w.shape[0]
    k = a.shape[1]
    n = w.shape[2] if w_is_kn else w.shape[1]
    assert (w.shape[1] if w_is_kn else w.shape[2]) == k, f"contraction mismatch: a {tuple(a.shape)} w {tuple(w.shape)}"
    out_dtype = out_dtype or a.dtype
    glu_fwd = act == ACT_SILU_GLU
    glu_bwd = act_bwd == ACT_SILU_GLU
    g = GemmArgs()
    g.mode, g.b_layout, g.num_experts = GEMM_ROWS, 1 if w_is_kn else 0, E
    if dense_rows:
        assert dense_rows % ROW_TILE == 0
        m = dense_rows if sum_experts else E * dense_rows
        g.dense, g.dense_rows, g.a_expert_rows = 1, dense_rows, a_expert_rows
        g.sum_experts = 1 if sum_experts else 0
    else:
        assert route is not None and a.shape[0] == route.row_cap
        m = route.row_cap
        g.tile_expert = _p(route.tile_expert)
        g.pad_offsets = _p(route.pad_offsets)      # expert-aligned tile raster of the pair kernels
        g.row_tile = route.row_tile
    g.m, g.n, g.k = m, n, k
    g.a, g.lda = _p(a), a.stride(0)
    g.b, g.ldb, g.b_expert_stride = _p(w), w.stride(1), w.stride(0)
    c_cols = n // 2 if glu_fwd else (2 * n if glu_bwd else n)
    if c_rows is not None:
        assert c_rows.dtype == torch.int64 and c_rows.numel() >= m and act == ACT_NONE and act_bwd == ACT_NONE
        assert not want_preact and out_dtype in (torch.bfloat16, torch.float32)
        c = None
        g.c, g.ldc, g.c_dtype, g.c_rows = None, c_cols, BF16 if out_dtype == torch.bfloat16 else F32, _p(c_rows)
    elif accumulate_into is not None:    # C += A . W (fp32 C; the fp32-accurate path sums six bf16 products this way)
        c = accumulate_into
        assert c.dtype == torch.float32 and c.shape == (m, c_cols) and c.is_contiguous() and not glu_fwd and not glu_bwd
        g.c, g.ldc, g.c_dtype, g.accumulate = _p(c), c_cols, F32, 1
        out_dtype = torch.float32
    else:
        c = torch.empty(m, c_cols, dtype=out_dtype, device=a.device)
        g.c, g.ldc, g.c_dtype = _p(c), c_cols, _dt(c)
    g.act = act
    pre = None
    if want_preact or glu_fwd:
        pre = torch.empty(m, n, dtype=out_dtype, device=a.device)
        g.preact, g.ldpre = _p(pre), n
    if bias is not None:
        bias = bias.contiguous()
        assert bias.shape == (E, n)
        g.bias, g.bias_dtype = _p(bias), _dt(bias)
        g.bias_after_round = 1 if bias_after_round else 0   # bf16(acc) + bias instead of bf16(acc + bias)
    if act_bwd != ACT_NONE:
        assert aux is not None and aux.dtype == torch.bfloat16 and aux.stride(1) == 1 and aux.shape[0] == m
        assert aux.shape[1] == (2 * n if glu_bwd else n)
        g.act_bwd, g.aux, g.ldaux = act_bwd, _p(aux), aux.stride(0)
    rs = None
    if rowsum_softplus is not None:
        assert c is not None and not glu_fwd and act_bwd == ACT_NONE and not want_preact
        rs = torch.empty(m, (n + 63) // 64, dtype=torch.float32, device=a.device)
        g.rowsum, g.rowsum_round = _p(rs), 1 if rowsum_softplus else 0
    algo_rows = E * dense_rows if dense_rows else route.n_slots
    _gemm(g, 2.0 * algo_rows * n * k, "rows_kn" if w_is_kn else "rows_nk")
    if rs is not None:
        return c, rs
    return (c, pre) if (want_preact or glu_fwd) else c


def gemm_reduce(a: torch.Tensor, b: torch.Tensor, num_experts: int, *, route: Optional[Route] = None,
                dense_rows: int = 0, a_expert_rows: int = 0, b_expert_rows: int = 0,
                out_dtype: torch.dtype = torch.float32, accumulate_into: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C[e] = A[rows of e]^T . B[rows of e]  -> [E, a.shape[1], b.shape[1]] (wgrad).  `out`: write into this contiguous
    [E, m, n] tensor (e.g. a symmetric gradient buffer of ep.WeightExchange) instead of allocating."""
    _cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    m, n = a.shape[1], b.shape[1]
    g = GemmArgs()
    g.mode, g.num_experts = GEMM_REDUCE, num_experts
    if dense_rows:
        assert dense_rows % ROW_TILE == 0
        g.dense, g.dense_rows, g.a_expert_rows, g.b_expert_stride = 1, dense_rows, a_expert_rows, b_expert_rows
        g.k = dense_rows
    else:
        assert route is not None and a.shape[0] == route.row_cap and b.shape[0] == route.row_cap
        g.pad_offsets = _p(route.pad_offsets)
        g.k = route.row_cap
    g.m, g.n = m, n
    g.a, g.lda = _p(a), a.stride(0)
    g.b, g.ldb = _p(b), b.stride(0)
    if accumulate_into is not None:
        c = accumulate_into
        assert c.dtype == torch.float32 and c.shape == (num_experts, m, n) and c.is_contiguous()
        g.accumulate = 1
    elif out is not None:
        c = out
        assert c.shape == (num_experts, m, n) and c.is_contiguous() and c.dtype in (torch.float32, torch.bfloat16)
    else:
        c = torch.empty(num_experts, m, n, dtype=out_dtype, device=a.device)
    g.c, g.ldc, g.c_expert_stride, g.c_dtype = _p(c), n, m * n, _dt(c)
    algo_rows = num_experts * dense_rows if dense_rows else route.n_slots
    _gemm(g, 2.0 * algo_rows * m * n, "reduce")
    return c


# ----------------------------------------------------------------------------------------------- elementwise
def _tile_map(route: Optional[Route], rows: int):
    """The route's tile map when `rows` is its padded row space (row tiles past the routed rows are skipped)."""
    if route is None or route.tile_expert is None or rows != route.row_cap:
        return None
    return route.tile_expert.data_ptr()


def act_fwd(z: torch.Tensor, act: int, route: Optional[Route] = None) -> torch.Tensor:
    _cuda(z)
    assert z.dim() == 2 and z.stride(1) == 1
    rows = z.shape[0]
    cols = z.shape[1] // 2 if act == ACT_SILU_GLU else z.shape[1]
    h = torch.empty(rows, cols, dtype=z.dtype, device=z.device)
    _call("csmoe_act_fwd", _p(z), _dt(z), rows, cols, z.stride(0), act, _p(h), cols, _tile_map(route, rows), _stream())
    return h


def act_bwd(z: torch.Tensor, dh: torch.Tensor, act: int, route: Optional[Route] = None) -> torch.Tensor:
    _cuda(z, dh)
    dh = dh.contiguous()
    rows = z.shape[0]
    cols = z.shape[1] // 2 if act == ACT_SILU_GLU else z.shape[1]
    assert dh.shape == (rows, cols) and dh.dtype == z.dtype
    dz = torch.empty_like(z)
    _call("csmoe_act_bwd", _p(z), _p(dh), _dt(z), rows, cols, z.stride(0), cols, act, _p(dz), _tile_map(route, rows),
          _stream())
    return dz


def bias_grad(g: torch.Tensor, num_experts: int, *, route: Optional[Route] = None, dense_rows: int = 0,
              out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    _cuda(g)
    n = g.shape[1]
    out_dtype = out_dtype or g.dtype
    db = torch.empty(num_experts, n, dtype=out_dtype, device=g.device)
    po = None if dense_rows else _p(route.pad_offsets)
    ws = torch.empty(int(_lib.load().csmoe_bias_grad_workspace_bytes(n, num_experts)) // 4, dtype=torch.float32,
                     device=g.device)
    _call("csmoe_bias_grad", _p(g), _dt(g), g.stride(0), n, num_experts, po, 1 if dense_rows else 0, dense_rows, _p(db),
          _dt(db), _p(ws), _stream(), kernels=2)
    return db


def act_bwd_bias(z: torch.Tensor, dh: torch.Tensor, act: int, num_experts: int, *, route: Optional[Route] = None,
                 dense_rows: int = 0, out_dtype: Optional[torch.dtype] = None):
    """(dz, dbias): dz = dh * act'(z) and its per-expert column sums in one pass (activation backward + bias gradient of
    the first projection)."""
    _cuda(z, dh)
    assert z.dtype == dh.dtype and z.shape == dh.shape and z.stride(1) == 1 and dh.stride(1) == 1
    n = z.shape[1]
    out_dtype = out_dtype or z.dtype
    dz = torch.empty_like(z)
    db = torch.empty(num_experts, n, dtype=out_dtype, device=z.device)
    po = None if dense_rows else _p(route.pad_offsets)
    ws = torch.empty(int(_lib.load().csmoe_bias_grad_workspace_bytes(n, num_experts)) // 4, dtype=torch.float32,
                     device=z.device)
    _call("csmoe_act_bwd_bias", _p(z), _p(dh), _dt(z), z.stride(0), dh.stride(0), n, num_experts, po,
          1 if dense_rows else 0, dense_rows, act, _p(dz), _p(db), _dt(db), _p(ws), _stream(), kernels=2)
    return dz, db


def split_bf16x3(src: torch.Tensor):
    """fp32 tensor -> (hi, mid, lo) bf16 tensors of the same shape with hi + mid + lo == src to 24 bits."""
    _cuda(src)
    assert src.dtype == torch.float32
    src = src.contiguous()
    hi, mid, lo = (torch.empty(src.shape, dtype=torch.bfloat16, device=src.device) for _ in range(3))
    _call("csmoe_split_f32_bf16x3", _p(src), _p(hi), _p(mid), _p(lo), src.numel(), _stream())
    return hi, mid, lo


# hi.hi, hi.mid, mid.hi, mid.mid, hi.lo, lo.hi: every product term above 2^-24 of the result
_X3_PAIRS = ((0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0))


def gemm_rows_f32(a: torch.Tensor, w: torch.Tensor, *, bias=None, act: int = ACT_NONE, want_preact: bool = False,
                  bias_after_round: bool = False, **where):
    """gemm_rows for fp32 operands with fp32-accurate results (six split-bf16 tensor-core products, fp32 accumulation).
    Same arguments as gemm_rows (layout / row-space keywords in **where); bias, activation and the saved pre-activation
    are applied by the last product's epilogue."""
    A = a if isinstance(a, tuple) else split_bf16x3(a)
    W = w if isinstance(w, tuple) else split_bf16x3(w)
    c = None
    for n, (i, j) in enumerate(_X3_PAIRS):
        last = n == len(_X3_PAIRS) - 1
        kw = dict(where)
        if last:
            kw.update(bias=bias, act=act, want_preact=want_preact)
        if c is None:
            c = gemm_rows(A[i], W[j], out_dtype=torch.float32, **kw)
        else:
            r = gemm_rows(A[i], W[j], accumulate_into=c, **kw)
            if last and want_preact:
                return r
    return c


def gemm_reduce_f32(a: torch.Tensor, b: torch.Tensor, num_experts: int, **where) -> torch.Tensor:
    """gemm_reduce for fp32 operands, fp32-accurate (see gemm_rows_f32)."""
    A = a if isinstance(a, tuple) else split_bf16x3(a)
    B = b if isinstance(b, tuple) else split_bf16x3(b)
    where.pop("out_dtype", None)
    c = None
    for i, j in _X3_PAIRS:
        c = gemm_reduce(A[i], B[j], num_experts, out_dtype=torch.float32, **where) if c is None else \
            gemm_reduce(A[i], B[j], num_experts, accumulate_into=c, **where)
    return c


def cast_bf16(src: torch.Tensor) -> torch.Tensor:
    """fp32 -> bf16 copy (parameters under autocast)."""
    _cuda(src)
    if src.dtype == torch.bfloat16:
        return src
    src = src.contiguous()
    dst = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    _call("csmoe_cast_f32_bf16", _p(src), _p(dst), src.numel(), _stream())
    return dst


# ----------------------------------------------------------------------------------------------- competition
def affinity_fwd(y: torch.Tensor, num_experts: int, T: int, t_pad: int, eager_bf16: bool) -> torch.Tensor:
    """y [E * t_pad, D] -> aff [T, E] f32.  eager_bf16: round each softplus and the mean to bf16 (eager bf16 modules);
    otherwise keep fp32 (autocast semantics: softplus and mean run in fp32)."""
    _cuda(y)
    D = y.shape[-1]
    aff = torch.empty(T, num_experts, dtype=torch.float32, device=y.device)
    _call("csmoe_affinity_fwd", _p(y), _dt(y), num_experts, T, t_pad, D, BF16 if eager_bf16 else F32, _p(aff), _stream())
    return aff


def affinity_from_rowsum(rowsum: torch.Tensor, num_experts: int, T: int, t_pad: int, D: int, eager_bf16: bool) -> torch.Tensor:
    """rowsum [E * t_pad, ceil(D / 64)] (from gemm_rows(..., rowsum_softplus=...)) -> aff [T, E] f32."""
    _cuda(rowsum)
    assert rowsum.dtype == torch.float32 and rowsum.is_contiguous() and rowsum.shape == (num_experts * t_pad, (D + 63) // 64)
    aff = torch.empty(T, num_experts, dtype=torch.float32, device=rowsum.device)
    _call("csmoe_affinity_from_rowsum", _p(rowsum), rowsum.shape[1], num_experts, T, t_pad, D, BF16 if eager_bf16 else F32,
          _p(aff), _stream())
    return aff


def affinity_bwd(y: torch.Tensor, daff: torch.Tensor, num_experts: int, T: int, t_pad: int,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(y, daff, out)
    D = y.shape[-1]
    daff = daff.contiguous().float()
    acc = out is not None
    if out is None:
        out = torch.zeros_like(y) if t_pad != T else torch.empty_like(y)
    _call("csmoe_affinity_bwd", _p(y), _p(daff), _dt(y), num_experts, T, t_pad, D, 1 if acc else 0, _p(out), _stream())
    return out


def diversity_fwd(y: torch.Tensor, sel: torch.Tensor, T: int, t_pad: int):
    """y [E * t_pad, D], sel [T, K] i32 -> (loss [] f32, inv_norm [T,K] f32, sim [T,K,K] f32): mean off-diagonal cosine
    similarity between each token's K selected expert outputs (competesmoe.py:180-218)."""
    _cuda(y, sel)
    assert sel.dtype == torch.int32 and sel.is_contiguous()
    K = sel.shape[1]
    D = y.shape[-1]
    f32 = dict(dtype=torch.float32, device=y.device)
    inv_norm = torch.empty(T, K, **f32)
    sim = torch.empty(T, K, K, **f32)
    partial = torch.empty(max(T, 1), **f32)
    loss = torch.empty((), **f32)
    _call("csmoe_diversity_fwd", _p(y), _dt(y), T, t_pad, D, K, _p(sel), _p(inv_norm), _p(sim), _p(partial), _p(loss),
          _stream(), kernels=2 if T > 0 else 1)
    return loss, inv_norm, sim


def compete_bwd(y: torch.Tensor, num_experts: int, T: int, t_pad: int, sel: torch.Tensor, *,
                daff: Optional[torch.Tensor] = None, w: Optional[torch.Tensor] = None,
                dout: Optional[torch.Tensor] = None, inv_norm: Optional[torch.Tensor] = None,
                sim: Optional[torch.Tensor] = None, g_div: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Gradient of the dense expert outputs y [E * t_pad, D] from the three consumers of a competition step (score,
    combine, diversity loss) in one pass; see csmoe_compete_bwd in include/csmoe.h."""
    _cuda(y, sel, daff, w, dout, inv_norm, sim, g_div)
    K = sel.shape[1]
    D = y.shape[-1]
    f32c = lambda t: None if t is None else t.contiguous().float()  # noqa: E731
    daff, w, g_div = f32c(daff), f32c(w), f32c(g_div)
    if dout is not None:
        dout = dout.contiguous()
        assert dout.dtype == y.dtype and dout.shape == (T, D)
    dy = torch.empty_like(y)
    _call("csmoe_compete_bwd", _p(y), _dt(y), num_experts, T, t_pad, D, K, _p(daff), _p(sel), _p(w), _p(dout),
          _p(inv_norm), _p(sim), _p(g_div), _p(dy), _stream())
    return dy


# ----------------------------------------------------------------------------------------------- losses
def losses_fwd(p: torch.Tensor, aff: torch.Tensor, aff_idx: torch.Tensor, gate_idx: Optional[torch.Tensor], batch: int):
    """Competition-step losses (csmoe_losses_fwd) -> (q [T,E], losses [5], cnt [B,E], colr [B,E])."""
    _cuda(p, aff, aff_idx, gate_idx)
    T, E = p.shape
    K = aff_idx.shape[1]
    assert T % batch == 0 and p.dtype == torch.float32 and aff.dtype == torch.float32 and aff_idx.dtype == torch.int32
    N = T // batch
    p, aff, aff_idx = p.contiguous(), aff.contiguous(), aff_idx.contiguous()
    gate_idx = None if gate_idx is None else gate_idx.contiguous()
    f32 = dict(dtype=torch.float32, device=p.device)
    ws = torch.empty(int(_lib.load().csmoe_losses_workspace_bytes(batch, N, E)) // 4, **f32)
    q = torch.empty(T, E, **f32)
    colq, cnt, colr = (torch.empty(batch, E, **f32) for _ in range(3))
    losses = torch.empty(5, **f32)
    _call("csmoe_losses_fwd", _p(p), _p(aff), _p(aff_idx), _p(gate_idx), batch, N, E, K, _p(q), _p(colq), _p(cnt), _p(colr),
          _p(losses), _p(ws), _stream(), kernels=2)
    return q, losses, cnt, colr


def losses_bwd(p, q, aff_idx, gate_idx, cnt, colr, g: torch.Tensor, batch: int):
    """-> (dp [T,E], daff [T,E]); g [5] f32 on the device."""
    _cuda(p, q, aff_idx, gate_idx, cnt, colr, g)
    T, E = p.shape
    K = aff_idx.shape[1]
    g = g.contiguous().float()
    dp = torch.empty_like(p)
    daff = torch.empty_like(p)
    _call("csmoe_losses_bwd", _p(p), _p(q), _p(aff_idx), _p(gate_idx), _p(cnt), _p(colr), _p(g), batch, T // batch, E, K,
          _p(dp), _p(daff), _stream())
    return dp, daff


def entropy_balance_fwd(probs: torch.Tensor, batch: int):
    """-> (loss [] f32, colr [B,E]) with loss = mean_b sum_e m log m, m = mean_n probs."""
    _cuda(probs)
    T, E = probs.shape
    assert T % batch == 0 and probs.dtype == torch.float32
    probs = probs.contiguous()
    N = T // batch
    f32 = dict(dtype=torch.float32, device=probs.device)
    ws = torch.empty(int(_lib.load().csmoe_losses_workspace_bytes(batch, N, E)) // 4, **f32)
    colr = torch.empty(batch, E, **f32)
    loss = torch.empty(1, **f32)
    _call("csmoe_entropy_balance_fwd", _p(probs), batch, N, E, _p(colr), _p(loss), _p(ws), _stream(), kernels=2)
    return loss[0], colr


def entropy_balance_bwd(colr: torch.Tensor, g: torch.Tensor, batch: int, N: int) -> torch.Tensor:
    _cuda(colr, g)
    E = colr.shape[1]
    g = g.reshape(1).contiguous().float()
    dprobs = torch.empty(batch * N, E, dtype=torch.float32, device=colr.device)
    _call("csmoe_entropy_balance_bwd", _p(colr), _p(g), batch, N, E, _p(dprobs), _stream())
    return dprobs


def topk_renorm_bwd(scores, w, idx, dw, sigmoid: bool, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """d scores [T,E] of csmoe_topk_renorm; added into `out` when given."""
    _cuda(scores, w, idx, dw, out)
    T, E = scores.shape
    K = idx.shape[1]
    dw = dw.contiguous().float()
    acc = out is not None
    if out is None:
        out = torch.empty_like(scores)
    assert out.dtype == torch.float32 and out.is_contiguous()
    _call("csmoe_topk_renorm_bwd", _p(scores), _p(w), _p(idx), _p(dw), T, E, K, 1 if sigmoid else 0, 1 if acc else 0, _p(out),
          _stream())
    return out


def dense_rows(idx: torch.Tensor, t_pad: int) -> torch.Tensor:
    """[T,K] i32 expert ids -> flat i32 row indices idx * t_pad + t into the dense [E * t_pad, D] outputs."""
    _cuda(idx)
    T, K = idx.shape
    rows = torch.empty(T * K, dtype=torch.int32, device=idx.device)
    _call("csmoe_dense_rows", _p(idx.contiguous()), T, K, t_pad, _p(rows), _stream())
    return rows


# ----------------------------------------------------------------------------------------------- fused sigma-MoE FFN
def sigma_ffn_supported(D: int, H: int, Dout: int) -> bool:
    return bool(_lib.load().csmoe_sigma_ffn_supported(D, H, Dout))


def sigma_ffn_fwd(x: torch.Tensor, keys: torch.Tensor, values: torch.Tensor, bias: Optional[torch.Tensor], route: Route,
                  slots_per_row: Optional[int] = None, xp: Optional[torch.Tensor] = None):
    """x [T, D] bf16 token-major, keys [E, D, H], values [E, H, Dout] bf16 -> (y [row_cap, Dout], h [row_cap, H]) bf16 in
    the padded expert-major row space; h = relu(x . keys + bias) is zero on padding rows."""
    _cuda(x, keys, values, bias, xp)
    assert x.dtype == keys.dtype == values.dtype == torch.bfloat16 and x.is_contiguous() and keys.is_contiguous() and values.is_contiguous()
    T, D = x.shape
    E, _, H = keys.shape
    Dout = values.shape[2]
    k = route.top_k if slots_per_row is None else slots_per_row
    h = torch.empty(route.row_cap, H, dtype=torch.bfloat16, device=x.device)
    y = torch.empty(route.row_cap, Dout, dtype=torch.bfloat16, device=x.device)
    if bias is not None:
        bias = bias.contiguous()
    _call("csmoe_sigma_ffn_fwd", _p(x), T, D, Dout, E, _p(keys), _p(values), _p(bias), _dt(bias) if bias is not None else F32,
          _p(route.row_to_slot), _p(route.tile_expert), route.row_cap, k, _p(xp), _p(h), _p(y), _stream())
    return y, h


def sigma_ffn_bwd(dout: torch.Tensor, keys: torch.Tensor, values: torch.Tensor, route: Route, slot_w: torch.Tensor,
                  h: torch.Tensor, slots_per_row: Optional[int] = None, dyp: Optional[torch.Tensor] = None):
    """-> (dz [row_cap, H], hw [row_cap, H], dxr [row_cap, D] bf16, dw_part [2, n_slots] f32)."""
    _cuda(dout, keys, values, slot_w, h, dyp)
    assert dout.dtype == torch.bfloat16 and dout.is_contiguous() and slot_w.dtype == torch.float32
    T, Dout = dout.shape
    E, D, H = keys.shape
    k = route.top_k if slots_per_row is None else slots_per_row
    dev = dout.device
    dz = torch.empty(route.row_cap, H, dtype=torch.bfloat16, device=dev)
    hw = torch.empty(route.row_cap, H, dtype=torch.bfloat16, device=dev)
    dxr = torch.empty(route.row_cap, D, dtype=torch.bfloat16, device=dev)
    dw_part = torch.empty(2, route.n_slots, dtype=torch.float32, device=dev)
    slot_w = slot_w.reshape(-1).contiguous()
    _call("csmoe_sigma_ffn_bwd", _p(dout), T, D, Dout, E, _p(keys), _p(values), _p(route.row_to_slot), _p(route.tile_expert),
          route.row_cap, k, _p(slot_w), route.n_slots, _p(h), _p(dyp), _p(dz), _p(hw), _p(dxr), _p(dw_part), _stream())
    return dz, hw, dxr, dw_part


def sigma_wgrad(a: torch.Tensor, g: torch.Tensor, num_experts: int, route: Route, transpose: bool,
                out_dtype: torch.dtype = torch.float32, slots_per_row: Optional[int] = None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """c[e] = a[rows of e]^T . gather(g)[rows of e]; a [row_cap, 128], g [T, N] token-major.
    transpose=False -> [E, 128, N]; True -> [E, N, 128].  `out`: write into this contiguous tensor of that shape."""
    _cuda(a, g)
    assert a.dtype == g.dtype == torch.bfloat16 and a.is_contiguous() and g.is_contiguous() and a.shape[0] == route.row_cap
    T, N = g.shape
    H = a.shape[1]
    k = route.top_k if slots_per_row is None else slots_per_row
    shape = (num_experts, N, H) if transpose else (num_experts, H, N)
    if out is not None:
        assert tuple(out.shape) == shape and out.is_contiguous()
        c = out
    else:
        c = torch.empty(shape, dtype=out_dtype, device=a.device)
    _call("csmoe_sigma_wgrad", _p(a), _p(g), T, N, num_experts, _p(route.row_to_slot), _p(route.pad_offsets), route.row_cap, k,
          1 if transpose else 0, _p(c), _dt(c), _stream())
    return c


# ----------------------------------------------------------------------------------------------- block tail
def layernorm_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, out_dtype: torch.dtype):
    """x [T, D] (fp32 / bf16) -> (y [T, D] out_dtype, mean [T], rstd [T]); fp32 statistics."""
    _cuda(x, gamma, beta)
    x = x.contiguous()
    T, D = x.shape
    y = torch.empty(T, D, dtype=out_dtype, device=x.device)
    mean = torch.empty(T, dtype=torch.float32, device=x.device)
    rstd = torch.empty(T, dtype=torch.float32, device=x.device)
    _call("csmoe_layernorm_fwd", _p(x), _dt(x), T, D, _p(gamma.float().contiguous()), _p(beta.float().contiguous()), float(eps),
          _p(y), _dt(y), _p(mean), _p(rstd), _stream())
    return y, mean, rstd


def layernorm_bwd(dy: torch.Tensor, x: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor, gamma: torch.Tensor):
    """-> (dx [T, D] x.dtype, dgamma [D] f32, dbeta [D] f32)."""
    _cuda(dy, x, mean, rstd, gamma)
    dy, x = dy.contiguous(), x.contiguous()
    T, D = x.shape
    dx = torch.empty_like(x)
    dgamma = torch.empty(D, dtype=torch.float32, device=x.device)
    dbeta = torch.empty(D, dtype=torch.float32, device=x.device)
    ws = torch.empty(int(_lib.load().csmoe_layernorm_bwd_workspace_bytes(T, D)) // 4, dtype=torch.float32, device=x.device)
    _call("csmoe_layernorm_bwd", _p(dy), _dt(dy), _p(x), _dt(x), _p(mean), _p(rstd), _p(gamma.float().contiguous()), T, D, _p(dx),
          _p(dgamma), _p(dbeta), _p(ws), _stream(), kernels=2)
    return dx, dgamma, dbeta


def combine_residual_fwd(y, slot_to_row, sel, w, T: int, top_k: int, residual: torch.Tensor, p: float, seed: int,
                         round_each: bool = False, round_w: bool = False) -> torch.Tensor:
    """out [T, D] (residual's dtype) = residual + dropout_p(sum_k w[t,k] * y[row(t,k)])."""
    _cuda(y, slot_to_row, sel, w, residual)
    D = y.shape[-1]
    residual = residual.reshape(T, D).contiguous()
    out = torch.empty_like(residual)
    w = w.reshape(-1).contiguous()
    flags = (1 if round_each else 0) | (2 if round_w else 0)
    _call("csmoe_combine_residual_fwd", _p(y), _dt(y), T, D, top_k, _p(slot_to_row), _p(sel), _p(w), flags, _p(residual),
          _dt(residual), float(p), int(seed), _p(out), _stream())
    return out


def residual_dropout_fwd(v: torch.Tensor, residual: torch.Tensor, p: float, seed: int) -> torch.Tensor:
    _cuda(v, residual)
    v = v.contiguous()
    T, D = v.shape
    residual = residual.reshape(T, D).contiguous()
    out = torch.empty_like(residual)
    _call("csmoe_residual_dropout_fwd", _p(v), _dt(v), _p(residual), _dt(residual), T, D, float(p), int(seed), _p(out), _stream())
    return out


def dropout_bwd(g: torch.Tensor, p: float, seed: int, out_dtype: torch.dtype) -> torch.Tensor:
    _cuda(g)
    g = g.contiguous()
    dv = torch.empty(g.shape, dtype=out_dtype, device=g.device)
    _call("csmoe_dropout_bwd", _p(g), _dt(g), g.numel(), float(p), int(seed), _p(dv), _dt(dv), _stream())
    return dv
