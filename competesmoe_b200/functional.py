"""Autograd functions that compose the libcsmoe kernels into the differentiable pieces of the MoE layer.

  GateFn        router GEMM + softmax + top-k (csmoe_router_fwd)         reference: competesmoe.py:301-320
  SparseFFNFn   permute -> grouped GEMMs -> combine, and its backward     reference: moe.py:172-213 / cvmm.py:460-551
  DenseFFNFn    every expert on every token (competition step)           reference: competesmoe.py:240-245 / :399-403
  AffinityFn    mean softplus score of each (token, expert)              reference: competesmoe.py:243 / :403
  SelectCombineFn  gate-weighted sum of the selected dense outputs        reference: recomputed by compute_moe (:374)
  CompeteLossesFn  softmax(affinity) + distillation MSE variants + balance / entropy balance on the affinity, one kernel
                   pair forward, one backward   reference: competesmoe.py:322-335,350-371 / layers/moe/competesmoe.py:541-593
  EntropyBalanceFn router-step entropy balance of the pretrain layer from the router's probabilities (moe.py:323-332)
  CompeteTailFn everything downstream of the dense expert outputs of a competition step (score, top-k, combine,
                diversity loss) with ONE backward kernel for d(dense outputs)  reference: competesmoe.py:219-259,:371-374

Stacked expert weights use one of two layouts: "nk" = [E, n, k] (nn.Linear.weight) or "kn" = [E, k, n] (sigma-MoE
keys / values).  All GEMM operands are bfloat16; fp32 accumulation happens in TMEM.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

import os

from . import ops
from ._lib import ROW_TILE

# Epilogue fusion switches (measured on B200 at the bench shape, profiles/r01b_*):
#   forward  GLU in the GEMM-1 epilogue: same time as GEMM + separate kernel, one launch and one HBM pass fewer -> on.
#   backward act'(z) in the dgrad-2 epilogue: the row-strided z reads make that epilogue longer than its main loop
#   (0.57 ms vs 0.33 + 0.17 ms) -> off by default until the epilogue I/O is staged through shared memory.
_FUSE_FWD = os.environ.get("CSMOE_FUSE_EPILOGUE_FWD", "1") != "0"
_FUSE_BWD = os.environ.get("CSMOE_FUSE_EPILOGUE_BWD", "0") != "0"
# activation backward + bias gradient of the first projection in one kernel (csmoe_act_bwd_bias)
_FUSE_ACT_BIAS = os.environ.get("CSMOE_FUSE_ACT_BIAS", "1") != "0"
#   competition score (mean softplus of the dense outputs) reduced in the down projection's epilogue: "1" always,
#   "0" never (stand-alone csmoe_affinity_fwd re-reads y), "auto" = only where that GEMM's main loop is long enough to
#   hide the extra epilogue math (contraction >= 1024; the sigma-MoE shapes with H = 128 are epilogue bound).
_SCORE_EPILOGUE = os.environ.get("CSMOE_SCORE_EPILOGUE", "auto")
_SIGMA_FUSED = os.environ.get("CSMOE_SIGMA_FUSED", "1") != "0"


@dataclass(frozen=True)
class FFNSpec:
    """Static description of the expert FFN."""
    act: int                 # ops.ACT_*
    kn_layout: bool = False  # False: w1 [E,F,D], w2 [E,Dout,F] (nn.Linear);  True: w1 [E,D,H], w2 [E,H,Dout] (sigma-MoE)
    round_each: bool = True  # combine: round the running sum to the activation dtype after every expert (moe.py:204)
    round_w: bool = False    # combine: round the routing weight to the activation dtype first (cvmm.py:483)
    fp32: bool = False       # fp32 operands outside autocast: fp32-accurate products (six split-bf16 MMAs, ops.gemm_rows_f32)
    bias_after_round: bool = False  # first projection: bf16(x . W) + b (pretrain: fp32 bias added to the bf16 cvmm result)
    return_hidden: bool = False  # SparseFFNFn also returns (h = act(z) [row_cap, H], row_to_slot) for the relu_pass_rate log

    @property
    def glu(self) -> bool:
        return self.act == ops.ACT_SILU_GLU


def _bf16(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None or t.dtype == torch.bfloat16:
        return t
    return ops.cast_bf16(t)


def _op(t: Optional[torch.Tensor], spec: "FFNSpec") -> Optional[torch.Tensor]:
    """GEMM operand in the arithmetic the spec asks for: bf16 (cast when needed) or, in fp32 mode, the fp32 tensor itself."""
    if t is None:
        return None
    if spec.fp32:
        return t if t.dtype == torch.float32 else t.float()
    return _bf16(t)


def _mm_rows(a, w, spec: "FFNSpec", **kw):
    if not spec.fp32:
        return ops.gemm_rows(a, w, **kw)
    for k in ("rowsum_softplus", "act_bwd", "aux", "c_rows"):
        assert kw.pop(k, None) in (None, ops.ACT_NONE), f"fp32-accurate path has no fused '{k}' epilogue"
    kw.pop("out_dtype", None)
    return ops.gemm_rows_f32(a, w, **kw)


def _mm_reduce(a, b, num_experts, spec: "FFNSpec", **kw):
    if not spec.fp32:
        return ops.gemm_reduce(a, b, num_experts, **kw)
    out_dtype = kw.pop("out_dtype", torch.float32)
    out = kw.pop("out", None)
    c = ops.gemm_reduce_f32(a, b, num_experts, **kw)
    return c.to(out_dtype) if out is None else out.copy_(c)


def _wx_operands(wx, spec: "FFNSpec", w1, b1, w2, b2=None):
    """Expert parameters as GEMM operands.  wx = None: this rank holds every expert (cast when needed).  wx =
    ep.WeightExchange: w1 / b1 / w2 are this rank's SHARDS; the exchange returns the full-size operand copies (cast +
    all-gather over peer memory, once per layer step)."""
    if wx is None:
        return _op(w1, spec), b1, _op(w2, spec), b2
    full = wx.operands({"w1": w1, "b1": b1, "w2": w2, "b2": b2}, torch.float32 if spec.fp32 else torch.bfloat16)
    return full["w1"], full["b1"], full["w2"], full["b2"]


def _wx_grad_out(wx, name: str, like: torch.Tensor):
    """Where a weight-gradient GEMM writes: a fresh tensor, or (expert parallel, weights exchanged) the full-size fp32
    gradient buffer the owners will reduce their slices from."""
    return None if wx is None else wx.grad_out(name, like.shape)


def _pad_rows(x: torch.Tensor, rows: int) -> torch.Tensor:
    if x.shape[0] == rows:
        return x
    pad = torch.zeros(rows - x.shape[0], x.shape[1], dtype=x.dtype, device=x.device)
    return torch.cat([x, pad], dim=0)


def _ffn_first(xp, w1, b1, spec: FFNSpec, **where):
    """z (pre-activation, with bias) and h = act(z) for the first projection."""
    if spec.glu and not spec.kn_layout and b1 is None and _FUSE_FWD and not spec.fp32:
        h, z = ops.gemm_rows(xp, w1, w_is_kn=False, act=ops.ACT_SILU_GLU, **where)   # GLU fused in the epilogue
        return z, h
    bar = spec.bias_after_round and b1 is not None and not spec.fp32
    if spec.glu:
        z = _mm_rows(xp, w1, spec, w_is_kn=spec.kn_layout, bias=b1, bias_after_round=bar, **where)
        return z, ops.act_fwd(z, spec.act, where.get("route"))
    if spec.act == ops.ACT_NONE:
        z = _mm_rows(xp, w1, spec, w_is_kn=spec.kn_layout, bias=b1, bias_after_round=bar, **where)
        return z, z
    h, z = _mm_rows(xp, w1, spec, w_is_kn=spec.kn_layout, bias=b1, act=spec.act, want_preact=True, bias_after_round=bar, **where)
    return z, h


# ------------------------------------------------------------------------------------------------ router
class GateFn(Function):
    """(x [T,D], wg [E,D]) -> logits [T,E] (x dtype), probs [T,E] f32, topk_w [T,K] f32, topk_idx [T,K] i32,
    losses [2] f32 = (balance loss, z-loss) of the router step (zeros unless want_aux).
    renorm_dtype: dtype of the LAYER's input, to which the reference rounds the top-k sum (`.to(x.dtype)`); None = x's.

    Forward: csmoe_router_fwd (+ csmoe_router_aux_fwd).  Backward: one fused csmoe_router_bwd call folds the routing
    weight gradient, any incoming d probs / d logits and the two aux-loss gradients into d logits, then dx and dWg."""

    @staticmethod
    def forward(ctx, x, wg, top_k: int, batch: int, want_aux: bool, renorm_dtype=None):
        wgx = wg if wg.dtype == x.dtype else wg.to(x.dtype)
        logits, probs, tw, ti = ops.router_fwd(x, wgx, top_k, renorm_dtype)
        cnt = lse = None
        if want_aux:
            losses, cnt, lse = ops.router_aux_fwd(logits, probs, ti, batch)
        else:
            losses = torch.zeros(2, dtype=torch.float32, device=x.device)
        ctx.save_for_backward(x, wgx, probs, tw, ti, cnt, lse)
        ctx.batch, ctx.wg_dtype, ctx.want_aux, ctx.renorm_dtype = batch, wg.dtype, want_aux, renorm_dtype
        ctx.mark_non_differentiable(ti)
        return logits, probs, tw, ti, losses

    @staticmethod
    @once_differentiable
    def backward(ctx, dlogits, dprobs, dtw, _, dlosses):
        x, wgx, probs, tw, ti, cnt, lse = ctx.saved_tensors
        dx, dwg = ops.router_bwd(x, wgx, probs, tw, ti, ctx.batch, dtw=dtw, dprobs=dprobs, dlogits=dlogits,
                                 lse=lse if ctx.want_aux else None, cnt=cnt if ctx.want_aux else None,
                                 g_losses=dlosses if ctx.want_aux else None, need_dx=ctx.needs_input_grad[0],
                                 need_dwg=ctx.needs_input_grad[1], wg_dtype=ctx.wg_dtype, renorm_dtype=ctx.renorm_dtype)
        return dx, dwg, None, None, None, None


# ------------------------------------------------------------------------------------------------ sparse experts
class SparseFFNFn(Function):
    """out[t] = sum_k w[t,k] * FFN_{sel[t,k]}(x[t])   with x [T,D] bf16, w [T,K] f32, sel [T,K] i32."""

    @staticmethod
    def forward(ctx, x, w, sel, w1, b1, w2, b2, spec: FFNSpec, wx=None):
        T, K = sel.shape
        xb = _op(x, spec)
        w1b, b1, w2b, b2 = _wx_operands(wx, spec, w1, b1, w2, b2)
        E = w1b.shape[0]
        route = ops.route_build(sel, E)
        xp = ops.gather_rows(xb, route)
        z, h = _ffn_first(xp, w1b, b1, spec, route=route)
        y = _mm_rows(h, w2b, spec, w_is_kn=spec.kn_layout, bias=b2, route=route)
        out = ops.combine_fwd(y, route.slot_to_row, route.sel, w, T, K, round_each=spec.round_each, round_w=spec.round_w)
        ctx.route, ctx.spec = route, spec
        ctx.x_dtype = x.dtype
        ctx.has_b = (b1 is not None, b2 is not None)
        ctx.save_for_backward(xp, z, h, y, w, w1, w2)
        # fp32 master weights (pretrain under autocast): keep the bf16 copies of this step for the backward pass
        ctx.wb = (w1b if (w1b is not w1 and (wx is not None or not spec.fp32)) else None,
                  w2b if (w2b is not w2 and (wx is not None or not spec.fp32)) else None)
        ctx.wx = wx
        if spec.return_hidden:
            hd = h.detach()
            ctx.mark_non_differentiable(hd, route.row_to_slot)
            return out, hd, route.row_to_slot
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout, _dh=None, _dmap=None):
        xp, z, h, y, w, w1, w2 = ctx.saved_tensors
        route, spec, wx = ctx.route, ctx.spec, ctx.wx
        T, K, E = route.n_slots // route.top_k, route.top_k, route.num_experts
        w1b = ctx.wb[0] if ctx.wb[0] is not None else _op(w1, spec)
        w2b = ctx.wb[1] if ctx.wb[1] is not None else _op(w2, spec)
        dout = _op(dout.contiguous(), spec)
        need = ctx.needs_input_grad
        dw = ops.combine_bwd_w(y, dout, route.slot_to_row, T, K) if need[1] else None
        wu = w.to(torch.bfloat16).float() if spec.round_w else w
        dyp = ops.gather_rows(dout, route, slot_w=wu)                          # w * dout in expert-major order
        db2 = ops.bias_grad(dyp, E, route=route, out_dtype=w2.dtype) if ctx.has_b[1] else None
        if spec.kn_layout:
            dw2 = _mm_reduce(h, dyp, E, spec, route=route, out_dtype=w2.dtype, out=_wx_grad_out(wx, "w2", w2b))  # [E, H, Dout]
        else:
            dw2 = _mm_reduce(dyp, h, E, spec, route=route, out_dtype=w2.dtype, out=_wx_grad_out(wx, "w2", w2b))  # [E, Dout, F]
        # dgrad of the second projection with the activation backward fused into its epilogue: dz = (dy W2) * act'(z)
        fused_db1 = False
        if _FUSE_BWD and not spec.fp32:
            dz = ops.gemm_rows(dyp, w2b, w_is_kn=not spec.kn_layout, route=route, act_bwd=spec.act, aux=z)
        else:
            dh = _mm_rows(dyp, w2b, spec, w_is_kn=not spec.kn_layout, route=route)
            if ctx.has_b[0] and _FUSE_ACT_BIAS and spec.act not in (ops.ACT_NONE, ops.ACT_SILU_GLU):
                # activation backward + bias gradient in one pass over dh / z
                dz, db1 = ops.act_bwd_bias(z, dh, spec.act, E, route=route, out_dtype=w1.dtype)
                fused_db1 = True
            else:
                dz = dh if spec.act == ops.ACT_NONE else ops.act_bwd(z, dh, spec.act, route)
        if not fused_db1:
            db1 = ops.bias_grad(dz, E, route=route, out_dtype=w1.dtype) if ctx.has_b[0] else None
        if spec.kn_layout:
            dw1 = _mm_reduce(xp, dz, E, spec, route=route, out_dtype=w1.dtype, out=_wx_grad_out(wx, "w1", w1b))  # [E, D, H]
        else:
            dw1 = _mm_reduce(dz, xp, E, spec, route=route, out_dtype=w1.dtype, out=_wx_grad_out(wx, "w1", w1b))  # [E, F, D]
        dx = None
        if need[0]:
            dxp = _mm_rows(dz, w1b, spec, w_is_kn=not spec.kn_layout, route=route)
            dx = ops.scatter_reduce(dxp, route.slot_to_row, T, K).to(ctx.x_dtype)
        if wx is not None:     # every rank's full-size gradients are complete: the owners sum their slices
            dw1, db1, dw2, db2 = wx.reduce({"w1": (dw1, w1), "b1": (db1, None), "w2": (dw2, w2), "b2": (db2, None)})
        return dx, dw, None, dw1, db1, dw2, db2, None, None


class SigmaFFNFn(Function):
    """SparseFFNFn for the sigma-MoE layout with expert size 128 and ReLU (the pretrain plugin's compute_moe_main,
    competesmoe.py:510-522) on the fused kernels of csrc/sigma_ffn.cu: one kernel gathers the token rows, runs both
    projections with the hidden activations kept on chip and writes y; backward is one fused dgrad kernel (dh, relu',
    d routing weight, dx rows) plus two gathered weight-gradient GEMMs.  Saved for backward: x (bf16) and h only."""

    @staticmethod
    def forward(ctx, x, w, sel, keys, bias, values, spec: FFNSpec, residual=None, drop_p: float = 0.0, drop_seed: int = 0,
                wx=None):
        """residual [T, Dout] given: returns residual + dropout(layer output) (the block tail of
        relative_moe_transformer.py:157) from the combine kernel's epilogue, in the residual's dtype."""
        T, K = sel.shape
        xb = _bf16(x).contiguous()
        kb, bias_full, vb, _ = _wx_operands(wx, spec, keys, bias, values)
        E = kb.shape[0]
        route = ops.route_build(sel, E, row_tile=ROW_TILE)
        y, h = ops.sigma_ffn_fwd(xb, kb, vb, bias_full, route)
        if residual is None:
            out = ops.combine_fwd(y, route.slot_to_row, route.sel, w, T, K, round_each=spec.round_each, round_w=spec.round_w)
        else:
            out = ops.combine_residual_fwd(y, route.slot_to_row, route.sel, w, T, K, residual, drop_p, drop_seed,
                                           round_each=spec.round_each, round_w=spec.round_w)
        ctx.tail = None if residual is None else (drop_p, drop_seed)
        ctx.route, ctx.spec, ctx.x_dtype = route, spec, x.dtype
        ctx.save_for_backward(xb, h, w, keys, values, bias)
        ctx.wb = (kb if kb is not keys else None, vb if vb is not values else None)
        ctx.wx = wx
        if spec.return_hidden:
            hd = h.detach()
            ctx.mark_non_differentiable(hd, route.row_to_slot)
            return out, hd, route.row_to_slot
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout, _dh=None, _dmap=None):
        xb, h, w, keys, values, bias = ctx.saved_tensors
        route, spec, wx = ctx.route, ctx.spec, ctx.wx
        T, K, E = route.n_slots // route.top_k, route.top_k, route.num_experts
        kb = ctx.wb[0] if ctx.wb[0] is not None else _bf16(keys)
        vb = ctx.wb[1] if ctx.wb[1] is not None else _bf16(values)
        dres = None
        if ctx.tail is not None:      # d(block output): identity into the residual, dropout mask regenerated for the layer
            dres = dout
            dout = ops.dropout_bwd(dout.reshape(-1, dout.shape[-1]), ctx.tail[0], ctx.tail[1], torch.bfloat16)
        dout = _bf16(dout.contiguous())
        wu = w.to(torch.bfloat16).float() if spec.round_w else w
        dz, hw, dxr, dw_part = ops.sigma_ffn_bwd(dout, kb, vb, route, wu, h)
        dw = dw_part.sum(0).view(T, K) if ctx.needs_input_grad[1] else None
        dvalues = ops.sigma_wgrad(hw, dout, E, route, transpose=False, out_dtype=values.dtype,
                                  out=_wx_grad_out(wx, "w2", vb))                                 # [E, H, Dout]
        if wx is not None:      # the owners start summing d values while d keys is still being computed
            pend_v = wx.reduce_begin({"w2": (dvalues, values)}, early=True)
        dkeys = ops.sigma_wgrad(dz, xb, E, route, transpose=True, out_dtype=keys.dtype,
                                out=_wx_grad_out(wx, "w1", kb))                                   # [E, D, H]
        dbias = ops.bias_grad(dz, E, route=route, out_dtype=bias.dtype) if bias is not None else None
        if wx is not None:      # ... and d keys / d bias while dx is reduced over k
            pend_k = wx.reduce_begin({"w1": (dkeys, keys), "b1": (dbias, None)})
        dx = ops.scatter_reduce(dxr, route.slot_to_row, T, K).to(ctx.x_dtype) if ctx.needs_input_grad[0] else None
        if wx is not None:
            dvalues, dkeys, dbias = wx.reduce_end(pend_v, pend_k)
        return dx, dw, None, dkeys, dbias, dvalues, None, dres, None, None, None


def sigma_fused_ok(x: torch.Tensor, keys: torch.Tensor, values: torch.Tensor, spec: FFNSpec, cdt: torch.dtype) -> bool:
    """The fused path covers the shapes the reference's pretraining sweeps use: expert size 128, ReLU, bf16 compute,
    model dims that are multiples of 128.  Everything else takes SparseFFNFn.  CSMOE_SIGMA_FUSED=0 switches it off."""
    return (_SIGMA_FUSED and spec.kn_layout and spec.act == ops.ACT_RELU and cdt == torch.bfloat16
            and keys.shape[1] % 128 == 0 and ops.sigma_ffn_supported(keys.shape[1], keys.shape[2], values.shape[2]))


# ------------------------------------------------------------------------------------------------ dense experts
class DenseFFNFn(Function):
    """y[e, t] = FFN_e(x[t]) for every expert and token -> [E * t_pad, Dout] (t_pad = T rounded up to the row tile).

    score_round (None / bool): also return rowsum [E * t_pad, ceil(Dout / 64)] = per-64-column sums of softplus(y),
    reduced in the epilogue of the down projection (north-star stage 2: the neural-response score never re-reads y);
    the bool selects eager-bf16 rounding of every softplus.  The row sums carry no gradient: CompeteTailFn
    differentiates the score through y."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, spec: FFNSpec, score_round: Optional[bool] = None, wx=None):
        T = x.shape[0]
        t_pad = (T + 2 * ROW_TILE - 1) // (2 * ROW_TILE) * (2 * ROW_TILE)   # 256: CTA-pair GEMM tiles
        xb = _pad_rows(_op(x, spec), t_pad)
        w1b, b1, w2b, b2 = _wx_operands(wx, spec, w1, b1, w2, b2)
        z, h = _ffn_first(xb, w1b, b1, spec, dense_rows=t_pad, a_expert_rows=0)
        fuse_score = score_round is not None and not spec.fp32 and \
            (_SCORE_EPILOGUE == "1" or (_SCORE_EPILOGUE == "auto" and h.shape[1] >= 1024))
        y = _mm_rows(h, w2b, spec, w_is_kn=spec.kn_layout, bias=b2, dense_rows=t_pad, a_expert_rows=t_pad,
                     rowsum_softplus=score_round if fuse_score else None)
        ctx.spec, ctx.T, ctx.t_pad, ctx.x_dtype = spec, T, t_pad, x.dtype
        ctx.has_b = (b1 is not None, b2 is not None)
        ctx.save_for_backward(xb, z, h, w1, w2)
        ctx.wx, ctx.wb = wx, ((w1b, w2b) if wx is not None else None)
        if score_round is None:
            return y
        if not fuse_score:
            return y, None
        y, rowsum = y
        ctx.mark_non_differentiable(rowsum)
        return y, rowsum

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, _drowsum=None):
        xb, z, h, w1, w2 = ctx.saved_tensors
        spec, T, t_pad, wx = ctx.spec, ctx.T, ctx.t_pad, ctx.wx
        w1b, w2b = ctx.wb if wx is not None else (_op(w1, spec), _op(w2, spec))
        E = w1b.shape[0]
        dy = _op(dy.contiguous(), spec)
        db2 = ops.bias_grad(dy, E, dense_rows=t_pad, out_dtype=w2.dtype) if ctx.has_b[1] else None
        o2 = _wx_grad_out(wx, "w2", w2b)
        if spec.kn_layout:
            dw2 = _mm_reduce(h, dy, E, spec, dense_rows=t_pad, a_expert_rows=t_pad, b_expert_rows=t_pad, out_dtype=w2.dtype, out=o2)
        else:
            dw2 = _mm_reduce(dy, h, E, spec, dense_rows=t_pad, a_expert_rows=t_pad, b_expert_rows=t_pad, out_dtype=w2.dtype, out=o2)
        fused_db1 = False
        if _FUSE_BWD and not spec.fp32:
            dz = ops.gemm_rows(dy, w2b, w_is_kn=not spec.kn_layout, dense_rows=t_pad, a_expert_rows=t_pad,
                               act_bwd=spec.act, aux=z)
        else:
            dh = _mm_rows(dy, w2b, spec, w_is_kn=not spec.kn_layout, dense_rows=t_pad, a_expert_rows=t_pad)
            if ctx.has_b[0] and _FUSE_ACT_BIAS and spec.act not in (ops.ACT_NONE, ops.ACT_SILU_GLU):
                dz, db1 = ops.act_bwd_bias(z, dh, spec.act, E, dense_rows=t_pad, out_dtype=w1.dtype)
                fused_db1 = True
            else:
                dz = dh if spec.act == ops.ACT_NONE else ops.act_bwd(z, dh, spec.act)
        if not fused_db1:
            db1 = ops.bias_grad(dz, E, dense_rows=t_pad, out_dtype=w1.dtype) if ctx.has_b[0] else None
        o1 = _wx_grad_out(wx, "w1", w1b)
        if spec.kn_layout:
            dw1 = _mm_reduce(xb, dz, E, spec, dense_rows=t_pad, a_expert_rows=0, b_expert_rows=t_pad, out_dtype=w1.dtype, out=o1)
        else:
            dw1 = _mm_reduce(dz, xb, E, spec, dense_rows=t_pad, a_expert_rows=t_pad, b_expert_rows=0, out_dtype=w1.dtype, out=o1)
        dx = None
        if ctx.needs_input_grad[0]:
            # dx[t] = sum_e dz[e, t] . W1[e]: one GEMM whose k loop runs over (expert, hidden) -- no [E, T, D] intermediate
            if dz.shape[1] % 64 == 0:
                dx = _mm_rows(dz, w1b, spec, w_is_kn=not spec.kn_layout, dense_rows=t_pad, a_expert_rows=t_pad,
                              sum_experts=True)[:T].to(ctx.x_dtype)
            else:
                dxe = _mm_rows(dz, w1b, spec, w_is_kn=not spec.kn_layout, dense_rows=t_pad, a_expert_rows=t_pad)
                dx = dxe.view(E, t_pad, -1)[:, :T].float().sum(0).to(ctx.x_dtype)
        if wx is not None:
            dw1, db1, dw2, db2 = wx.reduce({"w1": (dw1, w1), "b1": (db1, None), "w2": (dw2, w2), "b2": (db2, None)})
        return dx, dw1, db1, dw2, db2, None, None, None


class AffinityFn(Function):
    """aff[t, e] = mean_d softplus(y[e, t, d]) (fp32 tensor holding values rounded to y.dtype)."""

    @staticmethod
    def forward(ctx, y, num_experts: int, T: int, t_pad: int, eager_bf16: bool):
        aff = ops.affinity_fwd(y, num_experts, T, t_pad, eager_bf16)
        ctx.save_for_backward(y)
        ctx.dims = (num_experts, T, t_pad)
        return aff

    @staticmethod
    @once_differentiable
    def backward(ctx, daff):
        (y,) = ctx.saved_tensors
        E, T, t_pad = ctx.dims
        return ops.affinity_bwd(y, daff, E, T, t_pad), None, None, None, None


class SelectCombineFn(Function):
    """out[t] = sum_k w[t,k] * y[sel[t,k] * t_pad + t]  (ascending-expert order, rounded like compute_moe)."""

    @staticmethod
    def forward(ctx, y, w, sel, t_pad: int, spec: FFNSpec):
        T, K = sel.shape
        rows = (sel.long() * t_pad + torch.arange(T, device=sel.device).unsqueeze(1)).to(torch.int32).reshape(-1)
        sel_flat = sel.reshape(-1).contiguous()
        out = ops.combine_fwd(y, rows, sel_flat, w, T, K, round_each=spec.round_each, round_w=spec.round_w)
        ctx.save_for_backward(y, w, rows)
        ctx.dims = (T, K, spec)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        y, w, rows = ctx.saved_tensors
        T, K, spec = ctx.dims
        dout = dout.contiguous().to(y.dtype)
        dw = ops.combine_bwd_w(y, dout, rows, T, K)
        wu = w.to(torch.bfloat16).float() if spec.round_w else w
        dy = torch.zeros_like(y)
        # each (expert, token) row is selected at most once: plain indexed store of w * dout
        contrib = (wu.reshape(T, K, 1) * dout.float().unsqueeze(1)).to(y.dtype).reshape(T * K, -1)
        dy.index_copy_(0, rows.long(), contrib)
        return dy, dw, None, None, None


class GatherRowsFn(Function):
    """topk_out[t, k] = y[sel[t,k] * t_pad + t]  -> [T, K, D]  (input of the diversity loss, competesmoe.py:255-258)."""

    @staticmethod
    def forward(ctx, y, sel, t_pad: int):
        T, K = sel.shape
        rows = (sel.long() * t_pad + torch.arange(T, device=sel.device).unsqueeze(1)).reshape(-1)
        ctx.save_for_backward(rows)
        ctx.shape = y.shape
        return y.index_select(0, rows).view(T, K, -1)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (rows,) = ctx.saved_tensors
        dy = torch.zeros(ctx.shape, dtype=g.dtype, device=g.device)
        dy.index_copy_(0, rows, g.reshape(rows.numel(), -1))
        return dy, None, None


class CompeteTailFn(Function):
    """y [E * t_pad, D] -> (aff [T,E] f32, w [T,K] f32, idx [T,K] i32, out [T,D], diversity loss [] f32).

    The competition step's consumers of the dense expert outputs in one autograd node: neural-response score
    (competesmoe.py:243), top-k + renormalisation (:249-254, weights stay attached to the scores), gate-weighted sum of
    the selected outputs (what compute_moe recomputes at :374) and the diversity loss of the selected outputs (:180-218).
    Backward joins the three gradient sources of y in a single kernel (csmoe_compete_bwd) instead of three autograd
    branches summed in the activation dtype."""

    @staticmethod
    def forward(ctx, y, num_experts: int, T: int, t_pad: int, top_k: int, sigmoid: bool, x_dtype: torch.dtype,
                spec: FFNSpec, rowsum: Optional[torch.Tensor] = None):
        eager_bf16 = x_dtype == torch.bfloat16
        if rowsum is not None:   # the score was reduced in the down projection's epilogue: y is not read again
            aff = ops.affinity_from_rowsum(rowsum, num_experts, T, t_pad, y.shape[1], eager_bf16)
        else:
            aff = ops.affinity_fwd(y, num_experts, T, t_pad, eager_bf16)
        w, idx = ops.topk_renorm(aff, top_k, sigmoid=sigmoid, round_dtype=x_dtype, round_out=eager_bf16)
        rows = ops.dense_rows(idx, t_pad)
        out = ops.combine_fwd(y, rows, idx.reshape(-1), w, T, top_k, round_each=spec.round_each, round_w=spec.round_w)
        div, inv_norm, sim = ops.diversity_fwd(y, idx, T, t_pad)
        ctx.save_for_backward(y, aff, w, idx, rows, inv_norm, sim)
        ctx.dims = (num_experts, T, t_pad, top_k, sigmoid, spec)
        ctx.mark_non_differentiable(idx)
        return aff, w, idx, out, div

    @staticmethod
    @once_differentiable
    def backward(ctx, daff, dw_ext, _, dout, ddiv):
        y, aff, w, idx, rows, inv_norm, sim = ctx.saved_tensors
        E, T, t_pad, K, sigmoid, spec = ctx.dims
        dw = dw_ext
        if dout is not None:
            dout = dout.contiguous().to(y.dtype)
            dwc = ops.combine_bwd_w(y, dout, rows, T, K)
            dw = dwc if dw is None else dw + dwc
        if dw is not None:
            # top-k renormalisation backward (w = v / sum(v), v = scores or sigmoid(scores) at idx), added to the incoming
            # d aff in the same kernel
            daff = ops.topk_renorm_bwd(aff, w, idx, dw, sigmoid, out=None if daff is None else daff.contiguous().float().clone())
        wu = w.to(torch.bfloat16).float() if spec.round_w else w
        dy = ops.compete_bwd(y, E, T, t_pad, idx, daff=daff, w=wu if dout is not None else None, dout=dout,
                             inv_norm=inv_norm if ddiv is not None else None, sim=sim if ddiv is not None else None,
                             g_div=ddiv)
        return dy, None, None, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------------ losses
class CompeteLossesFn(Function):
    """(p [T,E] gate softmax, aff [T,E] scores, aff_idx [T,K] i32, gate_idx [T,K] i32 or None) -> (q [T,E], losses [5]).

    q = softmax(aff) (not differentiable here: its only consumers are the losses below);
    losses = (MSE(p, q), MSE at aff_idx, MSE at gate_idx, multimodal balance(aff_idx, q), pretrain entropy_balance(q)).
    Gradients: p from the three MSE terms (q detached, as in the reference), aff from the two balance terms."""

    @staticmethod
    def forward(ctx, p, aff, aff_idx, gate_idx, batch: int):
        q, losses, cnt, colr = ops.losses_fwd(p, aff, aff_idx, gate_idx, batch)
        ctx.save_for_backward(p, q, aff_idx, gate_idx, cnt, colr)
        ctx.batch = batch
        ctx.mark_non_differentiable(q)
        return q, losses

    @staticmethod
    @once_differentiable
    def backward(ctx, _dq, dlosses):
        p, q, aff_idx, gate_idx, cnt, colr = ctx.saved_tensors
        if dlosses is None:
            return None, None, None, None, None
        dp, daff = ops.losses_bwd(p, q, aff_idx, gate_idx, cnt, colr, dlosses, ctx.batch)
        return (dp if ctx.needs_input_grad[0] else None), (daff if ctx.needs_input_grad[1] else None), None, None, None


class EntropyBalanceFn(Function):
    """probs [T,E] f32 (batch-major, T = batch * N) -> mean_b sum_e m log m with m = mean_n probs: the pretrain layer's
    `entropy_balance(gate_logits)` (layers/moe/moe.py:323-332; log_softmax(logits) is the log of these probabilities)."""

    @staticmethod
    def forward(ctx, probs, batch: int):
        loss, colr = ops.entropy_balance_fwd(probs, batch)
        ctx.save_for_backward(colr)
        ctx.dims = (batch, probs.shape[0] // batch)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (colr,) = ctx.saved_tensors
        return ops.entropy_balance_bwd(colr, g, *ctx.dims), None


# ------------------------------------------------------------------------------------------------ block tail
class LayerNormCastFn(Function):
    """y = LayerNorm(x) written in `out_dtype` (fp32 statistics): norm2 of the pre-LN block plus the autocast cast that
    follows it, one pass (relative_moe_transformer.py:150-153)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps: float, out_dtype: torch.dtype):
        x2 = x.reshape(-1, x.shape[-1])
        y, mean, rstd = ops.layernorm_fwd(x2, gamma, beta, eps, out_dtype)
        ctx.save_for_backward(x2, mean, rstd, gamma)
        ctx.shape = x.shape
        return y.view(*x.shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x2, mean, rstd, gamma = ctx.saved_tensors
        dx, dgamma, dbeta = ops.layernorm_bwd(dy.reshape(x2.shape), x2, mean, rstd, gamma)
        return dx.view(ctx.shape), dgamma.to(gamma.dtype), dbeta.to(gamma.dtype), None, None


class ResidualDropoutFn(Function):
    """out = residual + dropout_p(v) in one pass (`src = src + self.dropout(src3)`, relative_moe_transformer.py:157); the
    mask comes from a counter-based Philox stream and is regenerated in backward."""

    @staticmethod
    def forward(ctx, v, residual, p: float, seed: int):
        out = ops.residual_dropout_fwd(v.reshape(-1, v.shape[-1]), residual, p, seed)
        ctx.tail, ctx.v_dtype = (p, seed), v.dtype
        return out.view(*residual.shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        dv = ops.dropout_bwd(g.reshape(-1, g.shape[-1]), ctx.tail[0], ctx.tail[1], ctx.v_dtype)
        return dv.view(*g.shape), g, None, None
