"""Oracle (CPU, plain PyTorch) for the sibling routers of the multimodal plugin: the baselines CompeteSMoE is compared
against, which share compute_moe and the loss helpers and differ only in the gate.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates /root/reference/moe_model/model/moe/{smoe.py, smoe_sigmoidgating.py, xmoe.py, smoe_perturbed.py, shard_smoe.py,
deepseekv3.py}.  Experts are weight dictionaries as in oracle/multimodal.py; gate parameters come in a dict:
    smoe, smoe_sigmoidgating           {"gate_w": [E, D]}
    xmoe, smoe_perturbed               {"inp_reduction_w": [E/2, D], "expert_embeddings": [E, E/2]}
    smoe_share, deepseekv3             {"gate_w": [E-1, D]}  (the last expert is the shared one)
Ties in top-k are broken lowest-index-first, like the CUDA kernels (DESIGN.md "routing parity").
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Sequence

import torch
import torch.nn.functional as F

from .multimodal import ExpertW, balanceloss, compute_moe, expert_forward, stable_topk, zloss

SIBLINGS = ("smoe", "smoe_sigmoidgating", "xmoe", "smoe_perturbed", "smoe_share", "deepseekv3")
XMOE_TEMPERATURE = 0.3          # xmoe.py:27, smoe_perturbed.py:26
PERTURBED_THETA = 0.1           # smoe_perturbed.py:11 (constructor default)


def combine_loss(selected, gate_softmax, gate_logits, num_experts: int, args: SimpleNamespace):
    """moe.py:214-226 with acitve_zloss=True."""
    balance = balanceloss(selected, gate_softmax, num_experts)
    z = zloss(gate_logits)
    return balance * args.balance_loss_coef + z * args.router_z_loss_coef, balance, z


def renorm_embeddings_(emb: torch.Tensor, theta: float = 0.0) -> None:
    """xmoe.py:81-85 / smoe_perturbed.py:126-130: the forward rescales the parameter in place (a side effect the
    drop-in keeps): every expert embedding gets norm 1.5 (1.5 * n / (n + theta) for the perturbed gate)."""
    with torch.no_grad():
        n = emb.norm(p=2.0, dim=-1, keepdim=True)
        emb.mul_(1.5 / (n + theta) if theta else 1.5 / n)


def cosine_gate(x, inp_reduction_w, emb, theta: float):
    """xmoe.py:51-67,79-93 (theta = 0: F.normalize with eps 1e-4) and smoe_perturbed.py:96-112,124-135 (theta > 0:
    divide by norm + theta).  Returns (gate_logits in x.dtype, gate_softmax in x.dtype)."""
    reduced = F.linear(x, inp_reduction_w)
    if theta:
        m1 = reduced.float() / (reduced.norm(p=2, dim=-1, keepdim=True) + theta)
    else:
        m1 = F.normalize(reduced.float(), p=2.0, dim=-1, eps=1e-4)
    logits = torch.matmul(m1, emb.float().transpose(0, 1)).type_as(reduced)
    ok = logits.isfinite()
    if not ok.all():
        logits = torch.where(ok, logits, logits[ok].min())
    softmax = F.softmax(logits / XMOE_TEMPERATURE, dim=-1, dtype=torch.float).to(x.dtype)
    return logits, softmax


def sibling_forward(name: str, x: torch.Tensor, gate: Dict[str, torch.Tensor], experts: Sequence[ExpertW], k: int,
                    out_dim: int, args: SimpleNamespace, return_id_experts: bool = False):
    """Returns (output, auxiliary_loss, None, infor_aux, debug).  `k` and len(experts) are the constructor's
    num_selected / num_of_experts (the shared-expert variants route k-1 of E-1 and always add the last expert)."""
    E = len(experts)
    want_aux = x.requires_grad
    scale_sel = scale_shared = None
    if name in ("smoe", "smoe_sigmoidgating"):
        logits = F.linear(x, gate["gate_w"])
        softmax = F.softmax(logits, dim=-1, dtype=torch.float32)
        if name == "smoe":                                   # smoe.py:20-42
            weights, selected = stable_topk(softmax, k)
            want_aux = x.requires_grad or return_id_experts  # smoe.py:48
        else:                                                # smoe_sigmoidgating.py:17-42
            weights, selected = stable_topk(torch.sigmoid(logits), k)
        weights = weights / torch.sum(weights, dim=-1, keepdim=True).to(x.dtype)
        routed, n_routed = experts, E
    elif name in ("xmoe", "smoe_perturbed"):                 # xmoe.py:76-104, smoe_perturbed.py:120-144
        theta = PERTURBED_THETA if name == "smoe_perturbed" else 0.0
        renorm_embeddings_(gate["expert_embeddings"], theta)
        logits, softmax = cosine_gate(x, gate["inp_reduction_w"], gate["expert_embeddings"], theta)
        weights, selected = stable_topk(softmax, k)
        weights = torch.softmax(weights, dim=2)              # _keepTopk: softmax over the k kept probabilities
        routed, n_routed = experts, E
    elif name in ("smoe_share", "deepseekv3"):               # shard_smoe.py:37-67, deepseekv3.py:37-56
        n_routed, k = E - 1, k - 1
        logits = F.linear(x, gate["gate_w"])
        softmax = F.softmax(logits, dim=-1, dtype=torch.float32)
        weights, selected = stable_topk(softmax, k)          # moe.py:113-132 topk_expert
        weights = weights / torch.sum(weights, dim=-1, keepdim=True).to(x.dtype)
        routed = experts[:n_routed]
        scale_sel, scale_shared = (0.5, 0.5) if name == "smoe_share" else (1.0, 1.0)
        if name == "deepseekv3":
            want_aux = True                                  # deepseekv3.py:47: the loss is computed unconditionally
    else:
        raise ValueError(name)
    output = compute_moe(x, routed, selected, weights, out_dim)
    if scale_sel is not None:
        shared = expert_forward(experts[n_routed], x)
        if name == "smoe_share":
            output = torch.zeros_like(output) + (shared * 0.5 + output * 0.5)     # shard_smoe.py:55
        else:
            output = torch.zeros_like(output) + (shared + output)                 # deepseekv3.py:45
    aux = torch.tensor(0.0, dtype=x.dtype)
    info: Dict[str, torch.Tensor] = {}
    if want_aux:
        aux, balance, z = combine_loss(selected, softmax, logits, n_routed, args)
        info = {"balance_loss": balance.clone().detach(), "router_z_loss": z.clone().detach()}
    debug = {"selected": selected, "weights": weights, "gate_softmax": softmax, "gate_logits": logits}
    return output, aux, None, info, debug
