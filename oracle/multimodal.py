"""Oracle (CPU, plain PyTorch) for the multimodal CompeteSMoE layer.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, function by function, /root/reference/moe_model/model/moe/{moe.py,competesmoe.py}.  Experts are described by
plain weight dictionaries instead of nn.Modules so that the same numbers can be fed to the CUDA path:
    {"kind": "mlp", "act": "gelu"|"gelu_tanh"|"relu"|"silu", "w1": [F, Din], "b1": [F]|None, "w2": [Dout, F], "b2": ...}
    {"kind": "glu", "act": "silu", "w1": [2F, Din] (gate rows then up rows), "w2": [Dout, F]}     (Phi3MLP)

One deliberate difference, documented in DESIGN.md "routing parity": torch.topk leaves the order of exactly tied scores
unspecified (and CPU and CUDA differ); the oracle defines ties as lowest-expert-index-first, which is what the CUDA
kernels implement.  Tokens whose k-th/(k+1)-th score margin is below 1e-3 are reported separately by the tests.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

ExpertW = Dict[str, object]


def default_args(**kw) -> SimpleNamespace:
    """The attributes the reference layer reads from `args` (competesmoe.py:13-30,350-371; moe.py:214-226)."""
    base = dict(rate_flip=0.05, warm_up=0.0, max_compete_in_iter=3, hybrid=False, router_theta=1.0,
                router_loss_coef=0.01, diversity_loss_coef=0.01, bal_comp_loss_coef=0.01, balance_loss_coef=0.01,
                router_z_loss_coef=0.001, norm_sigmoid=False, init_weight=True, moe_name="competesmoe")
    base.update(kw)
    return SimpleNamespace(**base)


# ------------------------------------------------------------------------------------------------ expert modules
def _activation(name: str, z: torch.Tensor) -> torch.Tensor:
    if name == "gelu":
        return F.gelu(z)
    if name == "gelu_tanh":
        return F.gelu(z, approximate="tanh")
    if name == "relu":
        return F.relu(z)
    if name == "silu":
        return F.silu(z)
    raise ValueError(name)


def expert_forward(ew: ExpertW, x: torch.Tensor) -> torch.Tensor:
    """SiglipMLP / Sequential(Linear, GELU, Linear) (siglip_smoe.py:85-97, multimodal_projector/builder.py:56-66) and
    Phi3MLP (transformers: down(up * act(gate)))."""
    if ew["kind"] == "mlp":
        h = _activation(ew["act"], F.linear(x, ew["w1"], ew.get("b1")))
        return F.linear(h, ew["w2"], ew.get("b2"))
    if ew["kind"] == "glu":
        gu = F.linear(x, ew["w1"])
        gate, up = gu.chunk(2, dim=-1)
        return F.linear(up * _activation(ew["act"], gate), ew["w2"])
    raise ValueError(ew["kind"])


# ------------------------------------------------------------------------------------------------ selection
def stable_topk(scores: torch.Tensor, k: int):
    """topk with deterministic tie-break (lowest index first); values identical to torch.topk (moe.py:130)."""
    order = torch.sort(scores, dim=-1, descending=True, stable=True).indices[..., :k]
    return torch.gather(scores, -1, order), order


def _topk_or_forced(scores: torch.Tensor, k: int, forced: Optional[torch.Tensor]):
    """stable_topk, or -- parity tests that compare values under IDENTICAL routing -- the scores at `forced` indices
    (the routing decision itself is compared separately, with the north-star's low-margin rule)."""
    if forced is None:
        return stable_topk(scores, k)
    forced = forced.long().view(*scores.shape[:-1], k)
    return torch.gather(scores, -1, forced), forced


def router_policy(x: torch.Tensor, gate_w: torch.Tensor, k: int, forced: Optional[torch.Tensor] = None):
    """competesmoe.py:301-320 + moe.py:113-132."""
    gate_logits = F.linear(x, gate_w)
    gate_softmax = F.softmax(gate_logits, dim=-1, dtype=torch.float32)
    weights, selected = _topk_or_forced(gate_softmax, k, forced)
    weights = weights / torch.sum(weights, dim=-1, keepdim=True).to(x.dtype)
    return weights, selected, gate_softmax, gate_logits


def competition_policy(x: torch.Tensor, experts: Sequence[ExpertW], k: int, norm_sigmoid: bool = False,
                       forced: Optional[torch.Tensor] = None):
    """competesmoe.py:219-259: run every expert on every token, score = mean softplus(output)."""
    B, N, _ = x.shape
    E = len(experts)
    affinity = torch.zeros(B, N, E, dtype=x.dtype, device=x.device)
    outs = []
    for i, ew in enumerate(experts):
        out_i = expert_forward(ew, x)
        affinity[:, :, i] = torch.mean(F.softplus(out_i), dim=-1)
        outs.append(out_i.unsqueeze(2))
    expert_outputs = torch.cat(outs, dim=2)
    affinity_softmax = F.softmax(affinity, dim=-1, dtype=torch.float32)
    if norm_sigmoid:
        weights, selected = _topk_or_forced(torch.sigmoid(affinity), k, forced)
    else:
        weights, selected = _topk_or_forced(affinity, k, forced)
    weights = weights / torch.sum(weights, dim=-1, keepdim=True).to(x.dtype)
    idx = selected.unsqueeze(-1).expand(B, N, k, expert_outputs.size(-1))
    topk_outputs = torch.gather(expert_outputs, dim=2, index=idx)
    return weights, selected, affinity_softmax, affinity, topk_outputs


# ------------------------------------------------------------------------------------------------ dispatch/combine
def compute_moe(x: torch.Tensor, experts: Sequence[ExpertW], selected: torch.Tensor, weights: torch.Tensor,
                out_dim: int) -> torch.Tensor:
    """moe.py:172-213: loop over experts in ascending id, gather rows, run the expert, weighted in-place add."""
    B, N, _ = x.shape
    results = torch.zeros(B, N, out_dim, dtype=x.dtype, device=x.device)
    for i, ew in enumerate(experts):
        b_idx, t_idx, k_idx = torch.where(selected == i)
        out = expert_forward(ew, x[b_idx, t_idx])
        results[b_idx, t_idx] += weights[b_idx, t_idx, k_idx].unsqueeze(0).T * out
    return results


# ------------------------------------------------------------------------------------------------ losses
def zloss(gate_logits: torch.Tensor) -> torch.Tensor:
    """moe.py:71-88."""
    return torch.square(torch.logsumexp(gate_logits, dim=-1)).mean()


def balanceloss(selected: torch.Tensor, probs: torch.Tensor, num_experts: int) -> torch.Tensor:
    """moe.py:90-110: only the top-1 choice enters the density."""
    density_proxy = probs.mean(dim=-2)                                     # '... n e -> ... e'
    top1 = F.one_hot(selected[..., 0], num_experts).float()                # rearrange('... k -> k ...')[0]
    density = top1.mean(dim=-2)
    return (density_proxy * density).mean() * float(num_experts ** 2)


def experts_diversity_loss(topk_outputs: torch.Tensor) -> torch.Tensor:
    """competesmoe.py:180-218: mean of the off-diagonal cosine-similarity matrix (diagonal zeroed, mean over K*K)."""
    eo = topk_outputs.to(torch.float32)
    B, N, K, D = eo.shape
    nrm = F.normalize(eo, p=2, dim=-1).view(B * N, K, D)
    sim = torch.bmm(nrm, nrm.transpose(1, 2))
    sim = sim * (1 - torch.eye(K, device=sim.device))
    return sim.mean()


def router_loss(gate_softmax: torch.Tensor, affinity_softmax: torch.Tensor) -> torch.Tensor:
    """competesmoe.py:322-335."""
    return F.mse_loss(gate_softmax, affinity_softmax)


# ------------------------------------------------------------------------------------------------ the layer
def competesmoe_forward(x: torch.Tensor, gate_w: torch.Tensor, experts: Sequence[ExpertW], k: int, out_dim: int,
                        args: SimpleNamespace, competition: bool, return_id_experts: bool = False,
                        forced_selected: Optional[torch.Tensor] = None):
    """competesmoe.py:337-415.  `competition` stands for the schedule test at :347
    (x.requires_grad and current_steps >= step_warm and prob_flips[...] == 1).
    Returns (output, auxiliary_loss, None, infor_aux, debug) where debug holds the routing tensors for parity tests.
    `forced_selected` [B,N,K]: evaluate the step under that routing decision (of the branch taken) instead of the
    oracle's own top-k; debug["own_selected"] still holds the oracle's decision."""
    E = len(experts)
    gate_weights, gate_sel, gate_softmax, gate_logits = router_policy(
        x, gate_w, k, None if competition else forced_selected)
    auxiliary_loss = torch.tensor(0.0, dtype=x.dtype, device=x.device)
    infor_aux: Dict[str, torch.Tensor] = {}
    debug = {"gate_selected": gate_sel, "gate_weights": gate_weights, "gate_softmax": gate_softmax,
             "gate_logits": gate_logits}
    if competition:
        aff_w, aff_sel, aff_softmax, aff, topk_out = competition_policy(x, experts, k, getattr(args, "norm_sigmoid", False),
                                                                        forced_selected)
        if getattr(args, "hybrid", False):
            g_topk = torch.gather(gate_softmax, dim=-1, index=aff_sel)
            a_topk = torch.gather(aff_softmax, dim=-1, index=aff_sel)
            routerloss = router_loss(gate_softmax, aff_softmax.detach()) + \
                router_loss(g_topk, a_topk.detach()) * args.router_theta
        else:
            routerloss = router_loss(gate_softmax, aff_softmax.detach())
        diversity = experts_diversity_loss(topk_out)
        balance = balanceloss(aff_sel, aff_softmax, E)
        auxiliary_loss = routerloss * args.router_loss_coef + diversity * args.diversity_loss_coef + \
            balance * args.bal_comp_loss_coef
        output = compute_moe(x, experts, aff_sel, aff_w, out_dim)
        infor_aux = {"balance_loss": balance.detach().clone(), "diversity_loss": diversity.detach().clone(),
                     "routerloss": routerloss.detach().clone()}
        debug.update(selected=aff_sel, weights=aff_w, affinity=aff, affinity_softmax=aff_softmax)
        sc = torch.sigmoid(aff) if getattr(args, "norm_sigmoid", False) else aff
        debug["own_selected"] = stable_topk(sc.detach(), k)[1]
    else:
        output = compute_moe(x, experts, gate_sel, gate_weights, out_dim)
        if x.requires_grad or return_id_experts:
            balance = balanceloss(gate_sel, gate_softmax, E)
            z = zloss(gate_logits)
            auxiliary_loss = balance * args.balance_loss_coef + z * args.router_z_loss_coef   # moe.py:214-226
            infor_aux = {"balance_loss": balance.detach().clone(), "router_z_loss": z.detach().clone()}
        debug.update(selected=gate_sel, weights=gate_weights)
        debug["own_selected"] = stable_topk(gate_softmax.detach(), k)[1]
    return output, auxiliary_loss, None, infor_aux, debug


def topk_margin(scores: torch.Tensor, k: int) -> torch.Tensor:
    """Smallest gap between consecutive scores among the k+1 largest per token.  A token whose margin is below 1e-3 may
    legitimately differ from the reference in which experts it picks (gap k|k+1) or in their order (gaps inside the
    top-k; the order decides the top-1 used by the balance loss); such tokens are exempt from the bit-exact routing
    comparison (BASELINE.json north_star) and counted separately."""
    s = torch.sort(scores.float(), dim=-1, descending=True).values
    top = s[..., : min(k + 1, s.shape[-1])]
    if top.shape[-1] < 2:
        return torch.full(s.shape[:-1], float("inf"))
    return (top[..., :-1] - top[..., 1:]).min(dim=-1).values


# ------------------------------------------------------------------------------------------------ schedule
def build_flip_schedule(flip_steps: int, rate_flip: float, max_compete_in_iter: int,
                        prior: Optional[List[torch.Tensor]] = None, draws: Optional[torch.Tensor] = None) -> torch.Tensor:
    """competesmoe.py:96-138 (create_balanced_flip_current): one uniform draw per step; a step whose cumulative
    competition count over earlier layers is already `max_compete_in_iter` pushes its flag left, else right.
    `draws` lets a test replay a fixed RNG stream; otherwise torch.rand(1) per step like the reference."""
    freq = torch.zeros(flip_steps, dtype=torch.int32)
    for p in prior or []:
        freq += p.int()
    cur = [False] * flip_steps
    for i in range(flip_steps):
        u = draws[i].item() if draws is not None else torch.rand(1).item()
        if u < rate_flip:
            if freq[i] < max_compete_in_iter:
                cur[i] = True
                freq[i] += 1
            else:
                placed = False
                for j in range(i - 1, -1, -1):
                    if freq[j] < max_compete_in_iter and not cur[j]:
                        cur[j] = True
                        freq[j] += 1
                        placed = True
                        break
                if not placed:
                    for j in range(i + 1, flip_steps):
                        if freq[j] < max_compete_in_iter and not cur[j]:
                            cur[j] = True
                            freq[j] += 1
                            break
    return torch.tensor(cur, dtype=torch.bool)
