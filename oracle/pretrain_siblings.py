"""Oracle (CPU, plain PyTorch) for the sibling routers of the language-pretraining plugin.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates /root/reference/moe_pretrain_model/layers/moe/{smoe.py, smoeut_norm.py, xmoe.py, smoe_perturbed.py,
deepseekv2.py, deepseekv3.py}: the sigma-MoE expert path (two CVMM calls, oracle/pretrain.py compute_moe_main) behind
different gates.  `params` holds the reference's parameter names:
    all           w_gate [E, D], keys [E, D, H], values [E, H, Dv]
    xmoe / smoe_perturbed   + expert_embeddings [E, E/2], expert_sel [E/2, D]
    deepseekv2 / deepseekv3 + keys_shared [1, D, H], values_shared [1, H, Dv]   (v3: + e_score_correction_bias [E], unused)
Ties in top-k: lowest index first (DESIGN.md "routing parity").
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Callable, Dict

import torch
import torch.nn.functional as F

from .multimodal import stable_topk
from .pretrain import compute_moe_main, entropy_balance

SIBLINGS = ("smoe", "smoe_sigmoid", "xmoe", "smoe_perturbed", "deepseekv2", "deepseekv3")
REG_NAME = {"smoe_sigmoid": "mlp_balance"}          # smoeut_norm.py:142; every other variant logs "mlp_ebalance"


def cosine_logits(x, expert_sel, emb, theta: float, op_dtype):
    """xmoe.py:117-153 / smoe_perturbed.py:123-159; rescales `emb` in place like the reference."""
    reduced = F.linear(x.to(op_dtype), expert_sel.to(op_dtype))
    with torch.no_grad():
        n = emb.norm(p=2.0, dim=-1, keepdim=True)
        emb.mul_(1.5 / (n + theta) if theta else 1.5 / n)
    if theta:
        m1 = reduced.float() / (reduced.norm(p=2, dim=-1, keepdim=True) + theta)
    else:
        m1 = F.normalize(reduced.float(), p=2.0, dim=-1, eps=1e-4)
    logits = torch.matmul(m1, emb.float().transpose(0, 1)).type_as(reduced)
    ok = logits.isfinite()
    if not ok.all():
        logits = torch.where(ok, logits, logits[ok].min())
    return logits


def sibling_forward(name: str, x: torch.Tensor, params: Dict[str, torch.Tensor], k: int, args: SimpleNamespace,
                    activation: Callable = F.relu, op_dtype: torch.dtype = torch.float32, theta: float = 0.1):
    """Returns (output [B, N, Dv], regs {name: loss}, debug)."""
    keys, values = params["keys"], params["values"]
    if name in ("xmoe", "smoe_perturbed"):
        logits = cosine_logits(x, params["expert_sel"], params["expert_embeddings"],
                               theta if name == "smoe_perturbed" else 0.0, op_dtype)
        softmax = F.softmax(logits / 0.3, dim=-1, dtype=torch.float).to(x.dtype)        # xmoe.py:161
        weights, selected = stable_topk(softmax, k)
        weights = torch.softmax(weights, dim=-1)                                          # _keepTopk
        scores_for_margin = softmax
    else:
        logits = F.linear(x.to(op_dtype), params["w_gate"].to(op_dtype))
        if name == "smoe":                                                                # smoe.py:233-239
            softmax = F.softmax(logits, dim=-1, dtype=torch.float32)
            weights, selected = stable_topk(softmax, k)
            weights = weights / torch.sum(weights, dim=-1, keepdim=True).to(x.dtype)
            scores_for_margin = softmax
        elif name == "smoe_sigmoid":                                                      # smoeut_norm.py:95-126
            sig = torch.sigmoid(logits)
            weights, selected = stable_topk(sig, k)
            weights = weights / torch.sum(weights, dim=-1, keepdim=True).to(x.dtype)
            scores_for_margin = sig
        elif name == "deepseekv2":                                                        # deepseekv2.py:135-141
            weights, selected = stable_topk(logits, k)
            weights = F.softmax(weights, dim=-1).to(x.dtype)
            scores_for_margin = logits
        elif name == "deepseekv3":                                                        # deepseekv3.py:142-150
            sig = torch.sigmoid(logits)
            weights, selected = stable_topk(sig, k)
            weights = weights / (weights.sum(dim=-1, keepdim=True) + 1e-20)
            weights = weights * 1                                                         # routed_scaling_factor
            scores_for_margin = sig
        else:
            raise ValueError(name)
    out = compute_moe_main(x, selected, weights, keys, values, activation, op_dtype)
    if name in ("deepseekv2", "deepseekv3"):                                              # shared expert: every token, weight 1
        sel0 = torch.zeros(*selected.shape[:-1], 1, dtype=selected.dtype)
        one = torch.ones(*selected.shape[:-1], 1)
        out = out + compute_moe_main(x, sel0, one, params["keys_shared"], params["values_shared"], activation, op_dtype)
    res = out.view(*x.shape[:-1], values.shape[-1])
    regs = {REG_NAME.get(name, "mlp_ebalance"): entropy_balance(logits) * (args.balance_loss_coef / 1)}
    return res, regs, {"selected": selected, "weights": weights, "scores": scores_for_margin, "gate_logits": logits}


def att_projection(x: torch.Tensor, params: Dict[str, torch.Tensor], n_experts: int, n_copies: int, k: int,
                   op_dtype: torch.dtype = torch.float32, theta: float = 0.1):
    """smoe_perturbed.py:199-226 (`att_forward` + `compute_moe` of a layer built with is_att=True, the expert projections
    of FullMoeRopeAttention, full_moe_relative_attention.py:267-296,351-389): per head, softmax(cosine gate / 0.3) over that
    head's experts in x's dtype, top-k, softmax of the kept values as weights; the projection is x @ experts[head * E + e]
    ([D, d_head] each) summed over the k selections with those weights (CVMM with reduction_weight, cvmm.py:481-483:
    weight rounded to the op dtype, fp32 accumulate, one rounding).
    params: expert_sel [E_total / 2, D], expert_embeddings [E_total, E_total / 2], experts [E_total, D, d_head].
    Returns (out [..., n_copies, d_head], debug)."""
    logits = cosine_logits(x, params["expert_sel"], params["expert_embeddings"], theta, op_dtype)
    logits = logits.view(*logits.shape[:-1], n_copies, -1)
    softmax = F.softmax(logits / 0.3, dim=-1, dtype=torch.float).to(x.dtype)
    _, idx = stable_topk(softmax.detach(), k)
    val = torch.softmax(torch.gather(softmax, -1, idx), dim=-1)
    flat = torch.arange(n_copies, device=idx.device).view(*([1] * (idx.dim() - 2)), n_copies, 1) * n_experts + idx
    w = params["experts"].to(op_dtype)[flat]                                   # [..., heads, k, D, d_head]
    proj = torch.einsum("...d,...hkde->...hke", x.to(op_dtype), w)
    out = (val.to(op_dtype).unsqueeze(-2).float() @ proj.float()).squeeze(-2).to(op_dtype)
    return out, {"selected": idx, "weights": val, "scores": softmax, "gate_logits": logits}
