"""Oracle (CPU, plain PyTorch) for the language-pretraining CompeteSMoE layer and its CVMM op.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates /root/reference/moe_pretrain_model/layers/cvmm.py (index algebra and the CVMM autograd function, with the
Triton kernels replaced by per-expert matmuls) and layers/moe/{moe.py,competesmoe.py}.  Mixed precision is explicit:
`op_dtype` plays the role of the CUDA autocast dtype (cvmm.py:29-32); ops that CUDA autocast runs in fp32 (softmax,
log_softmax, softplus, mse_loss, normalize) are computed in fp32 here as well.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F

from .multimodal import _topk_or_forced, stable_topk


def default_args(**kw) -> SimpleNamespace:
    """Attributes read by layers/moe/competesmoe.py (:100-121,:456-490,:540-605)."""
    base = dict(warm_up=0.0, rate_flip=0.07, stop_after=100, max_compete_in_iter=3, is_cosine=False,
                is_norm_weight=False, norm_sigmoid=False, scale_weight=1.0, hybrid=False, tribrid=False, in_topk=False,
                balance_affinity=False, balance_loss_coef=0.01, balance_loss_coef_comp=0.01, router_loss_coef=0.01,
                router_theta=1.0, test_only=False)
    base.update(kw)
    return SimpleNamespace(**base)


# ------------------------------------------------------------------------------------------------ CVMM
@dataclass
class Sel:
    """cvmm.py:11-20 (CVMMSel)."""
    raw_sel: torch.Tensor
    sel: torch.Tensor
    sel_index: torch.Tensor
    out_index: Optional[torch.Tensor] = None
    reduction_weight: Optional[torch.Tensor] = None

    def clone(self) -> "Sel":
        return Sel(self.raw_sel, self.sel, self.sel_index, self.out_index, self.reduction_weight)


def prepare_sel2(sel: torch.Tensor, w: Optional[torch.Tensor] = None) -> Sel:
    """cvmm.py:580-592, with the sort made stable (the reference's `sort()` is unspecified on ties)."""
    k = sel.shape[-1]
    fsel = sel.flatten()
    ssel, sel_index = torch.sort(fsel, stable=True)
    return Sel(sel, ssel.view_as(sel), sel_index // k, sel_index, w)


def cvmm(x: torch.Tensor, sel: Sel, keys: torch.Tensor, op_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """cvmm.py:555-577 + CVMM.forward :464-488.  out[out_index[i]] = x[sel_index[i]] @ keys[ssel[i]] for sorted row i,
    then the optional weighted reduction over the K selections.  Differentiable through torch autograd, which yields the
    same gradients as CVMM.backward (:491-551)."""
    x2 = x.flatten(end_dim=-2)
    ssel = sel.sel.flatten()
    M = ssel.shape[0]
    E, _, N = keys.shape
    dest = sel.sel_index if sel.out_index is None else sel.out_index
    bounds = torch.searchsorted(ssel.contiguous(), torch.arange(E + 1, dtype=ssel.dtype, device=ssel.device))
    pieces = []
    for e in range(E):
        lo, hi = int(bounds[e]), int(bounds[e + 1])
        a = x2[sel.sel_index[lo:hi]].to(op_dtype)
        pieces.append(a @ keys[e].to(op_dtype))
    stacked = torch.cat(pieces, dim=0)
    out = torch.zeros(M, N, dtype=op_dtype, device=stacked.device).index_copy(0, dest, stacked)
    out = out.view(*sel.sel.shape, N)
    if sel.reduction_weight is not None:
        rw = sel.reduction_weight
        out = out.view(*rw.shape, N)
        out = (rw.unsqueeze(-2).type_as(out) @ out).squeeze(-2)
    return out


# ------------------------------------------------------------------------------------------------ losses
def entropy_balance(logits: torch.Tensor) -> torch.Tensor:
    """moe.py:323-332 with utils/entropy.py:21-22 and utils/distributed_ops.py:47-58 (sync_distributed=False)."""
    s = logits.flatten(1, -2)
    ls = F.log_softmax(s.float(), dim=-1)
    lm = ls.logsumexp(-2) - math.log(ls.shape[-2])
    return -(-(lm * lm.exp()).sum(-1)).mean()


def experts_diversity_loss(topk_outputs: torch.Tensor) -> torch.Tensor:
    """competesmoe.py:330-372."""
    B, N, K, D = topk_outputs.shape
    nrm = F.normalize(topk_outputs.float(), p=2, dim=-1).view(B * N, K, D)
    sim = torch.bmm(nrm, nrm.transpose(1, 2)) * (1 - torch.eye(K, device=nrm.device))
    return sim.mean()


def router_loss(gate_softmax: torch.Tensor, affinity_softmax: torch.Tensor) -> torch.Tensor:
    return F.mse_loss(gate_softmax.float(), affinity_softmax.float())


# ------------------------------------------------------------------------------------------------ the layer
def compute_gate(x, w_gate, args, op_dtype):
    """competesmoe.py:456-464."""
    if getattr(args, "is_cosine", False) and not getattr(args, "is_norm_weight", False):
        return F.linear(F.normalize(x.float(), p=2.0, dim=-1).to(op_dtype), F.normalize(w_gate.float(), p=2.0, dim=-1).to(op_dtype))
    if getattr(args, "is_norm_weight", False):
        return F.linear(x.to(op_dtype), F.normalize(w_gate.float(), p=2.0, dim=-1).to(op_dtype))
    return F.linear(x.to(op_dtype), w_gate.to(op_dtype))


def router_policy(x, w_gate, k, args, op_dtype, forced=None):
    """competesmoe.py:465-490.  `forced`: see oracle.multimodal._topk_or_forced."""
    logits = compute_gate(x, w_gate, args, op_dtype)
    if getattr(args, "norm_sigmoid", False):
        gate_softmax = F.softmax(logits, dim=-1, dtype=torch.float32)
        weights, selected = _topk_or_forced(logits, k, forced)
        weights = torch.sigmoid(weights / getattr(args, "scale_weight", 1.0))
    else:
        gate_softmax = F.softmax(logits, dim=-1, dtype=torch.float32)
        weights, selected = _topk_or_forced(gate_softmax, k, forced)
    weights = weights / torch.sum(weights, dim=-1, keepdim=True).to(x.dtype)
    return weights, selected, gate_softmax, logits


def competition_policy(x, keys, values, k, activation, op_dtype, forced=None):
    """competesmoe.py:381-414 (competition_policy_mlp_faster): dense all-expert pass, score = mean softplus."""
    B, N, D = x.shape
    eo = torch.matmul(x.reshape(-1, D).to(op_dtype), keys.to(op_dtype))          # [E, T, H]
    eo = activation(eo)
    eo = torch.matmul(eo, values.to(op_dtype))                                   # [E, T, Dv]
    eo = eo.transpose(1, 0)                                                      # [T, E, Dv]
    aff = torch.mean(F.softplus(eo.float()), dim=-1).view(B, N, -1)
    aff_softmax = F.softmax(aff, dim=-1, dtype=torch.float32)
    weights, selected = _topk_or_forced(aff, k, forced)
    weights = weights / torch.sum(weights, dim=-1, keepdim=True).to(x.dtype)
    eo = eo.reshape(B, N, *eo.shape[1:])
    idx = selected.unsqueeze(-1).expand(B, N, k, eo.size(-1))
    return weights, selected, aff_softmax, aff, torch.gather(eo, dim=2, index=idx)


def compute_moe_main(x, selected, weights, keys, values, activation, op_dtype, bias=None):
    """competesmoe.py:510-522 with moe.py:397-416 (compute_scores)."""
    s = prepare_sel2(selected.int())
    scores = cvmm(x, s, keys, op_dtype)
    if bias is not None:
        scores = scores + bias[s.raw_sel.long()]
    scores = activation(scores)
    s2 = s.clone()
    s2.reduction_weight = weights
    s2.sel_index = s2.out_index
    s2.out_index = None
    return cvmm(scores, s2, values, op_dtype)


def competesmoe_forward(x: torch.Tensor, w_gate: torch.Tensor, keys: torch.Tensor, values: torch.Tensor, k: int,
                        args: SimpleNamespace, competition: bool, activation: Callable = F.relu,
                        op_dtype: torch.dtype = torch.float32, bias=None, o_bias=None, forced_selected=None):
    """competesmoe.py:524-616.  Returns (output [B,N,Dv], regs: name -> loss as passed to add_reg, debug).
    `forced_selected` [B,N,K]: evaluate the branch taken under that routing decision (parity of values under identical
    routing); debug["own_selected"] keeps the oracle's own decision."""
    regs: Dict[str, torch.Tensor] = {}
    gw, gsel, gsoft, glogits = router_policy(x, w_gate, k, args, op_dtype, None if competition else forced_selected)
    debug = {"gate_selected": gsel, "gate_weights": gw, "gate_softmax": gsoft, "gate_logits": glogits}
    if competition:
        aw, asel, asoft, aff, topk_out = competition_policy(x, keys, values, k, activation, op_dtype, forced_selected)
        out = compute_moe_main(x, asel, aw, keys, values, activation, op_dtype, bias)
        regs["mlp_comp_diver_loss"] = experts_diversity_loss(topk_out) * args.balance_loss_coef_comp / 2
        if args.balance_affinity:
            regs["mlp_comp_ebalance"] = entropy_balance(asoft) * args.balance_loss_coef_comp / 2
        if args.in_topk:
            rl = router_loss(torch.gather(gsoft, -1, asel), torch.gather(asoft, -1, asel).detach())
        elif args.hybrid:
            rl = router_loss(gsoft, asoft.detach()) + \
                router_loss(torch.gather(gsoft, -1, asel), torch.gather(asoft, -1, asel).detach()) * args.router_theta
        elif args.tribrid:
            rl = router_loss(gsoft, asoft.detach()) + \
                router_loss(torch.gather(gsoft, -1, asel), torch.gather(asoft, -1, asel).detach()) * args.router_theta + \
                router_loss(torch.gather(gsoft, -1, gsel), torch.gather(asoft, -1, gsel).detach()) * args.router_theta
        else:
            rl = router_loss(gsoft, asoft.detach())
        regs["mlp_router_loss"] = rl * args.router_loss_coef
        debug.update(selected=asel, weights=aw, affinity=aff, affinity_softmax=asoft,
                     own_selected=stable_topk(aff.detach(), k)[1])
    else:
        out = compute_moe_main(x, gsel, gw, keys, values, activation, op_dtype, bias)
        regs["mlp_ebalance"] = entropy_balance(glogits) * (args.balance_loss_coef / 1)
        debug.update(selected=gsel, weights=gw, own_selected=stable_topk(
            (glogits if getattr(args, "norm_sigmoid", False) else gsoft).detach(), k)[1])
    res = out.view(*x.shape[:-1], values.shape[-1])
    if o_bias is not None:
        res = res + o_bias
    return res, regs, debug
