"""Sibling routers of the multimodal plugin on the same kernels: the baselines CompeteSMoE is
compared against.  Only the gate differs; dispatch, expert FFN and combine are the libcsmoe path of MoeLayer.

reference: moe_model/model/moe/{smoe.py, smoe_sigmoidgating.py, xmoe.py, smoe_perturbed.py, shard_smoe.py, deepseekv3.py}
Same registry names, constructor, forward signature, return tuple, parameter names (gate.weight, expert_embeddings,
inp_reduction.weight, experts.{e}.*) and side effects (the cosine gates rescale `expert_embeddings` in place every
forward) as the reference classes.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .functional import DenseFFNFn, GateFn
from .multimodal import MoeLayer, TopkRenormFn, register_moe


def _aux_info(balance_loss, router_z_loss):
    return {"balance_loss": balance_loss.detach().clone(), "router_z_loss": router_z_loss.detach().clone()}


@register_moe("smoe")
class SMoeLayer(MoeLayer):
    """reference: smoe.py:12-64 -- softmax gate, top-k renormalised; balance + z loss when training or asked for ids."""

    def __init__(self, in_embed_dim=768, out_embed_dim=768, num_of_experts=4, num_selected=2, expert=None, args=None):
        super().__init__(in_embed_dim, out_embed_dim, num_of_experts, num_selected, expert, args)
        self.log_metrics = {}
        self.is_vision = False
        self.init_gate_weights()

    def forward(self, x, return_id_experts=False, is_vision=False):
        self.is_vision = is_vision
        B, N, D = x.shape
        x2 = x.reshape(B * N, D)
        lay, w1, b1, w2, b2 = self._stacked_weights()
        want_aux = bool(x.requires_grad or return_id_experts)
        logits, probs, gw, gidx, losses = GateFn.apply(x2, self.gate.weight, self.num_selected, B, want_aux)
        out = self._sparse_ffn(x2, gw, gidx, w1, b1, w2, b2, self._spec(lay))
        aux, info = x.new_zeros(()), {}
        if want_aux:
            aux = losses[0] * self.args.balance_loss_coef + losses[1] * self.args.router_z_loss_coef
            info = _aux_info(losses[0], losses[1])
            self.log_metrics.update(weights=gw.view(B, N, -1), gate_softmax=probs.view(B, N, -1),
                                    selected_experts=gidx.view(B, N, -1))
        self.last_routing = (gidx.view(B, N, -1), gw.detach().view(B, N, -1))
        return out.view(B, N, self.out_embed_dim).to(x.dtype), aux, None, info


@register_moe("smoe_sigmoidgating")
class SMoESigmoidGating(MoeLayer):
    """reference: smoe_sigmoidgating.py:8-58 -- top-k of sigmoid(logits) renormalised; the losses use the softmax."""

    def __init__(self, in_embed_dim=768, out_embed_dim=768, num_of_experts=4, num_selected=2, expert=None, args=None):
        super().__init__(in_embed_dim, out_embed_dim, num_of_experts, num_selected, expert, args)
        self.sigmoid = nn.Sigmoid()
        self.init_gate_weights()

    def forward(self, x, return_id_experts=False, is_vision=False):
        B, N, D = x.shape
        x2 = x.reshape(B * N, D)
        lay, w1, b1, w2, b2 = self._stacked_weights()
        want_aux = bool(x.requires_grad)
        # sigmoid is monotonic: the softmax top-1 that the fused balance loss counts is the sigmoid top-1
        logits, probs, _, _, losses = GateFn.apply(x2, self.gate.weight, self.num_selected, B, want_aux)
        gw, gidx = TopkRenormFn.apply(logits.float(), self.num_selected, True, x.dtype)
        out = self._sparse_ffn(x2, gw, gidx, w1, b1, w2, b2, self._spec(lay))
        aux, info = x.new_zeros(()), {}
        if want_aux:
            aux = losses[0] * self.args.balance_loss_coef + losses[1] * self.args.router_z_loss_coef
            info = _aux_info(losses[0], losses[1])
        self.last_routing = (gidx.view(B, N, -1), gw.detach().view(B, N, -1))
        return out.view(B, N, self.out_embed_dim).to(x.dtype), aux, None, info


class _CosineGateLayer(MoeLayer):
    """Shared body of xmoe.py:11-104 and smoe_perturbed.py:9-144: tokens are projected to E/2 dimensions, compared by
    (perturbed) cosine similarity with per-expert embeddings of norm 1.5, softmax at temperature 0.3, top-k, softmax
    over the kept probabilities."""
    theta = 0.0
    _inplace_params = ("expert_embeddings",)

    def _init_cosine_gate(self, in_embed_dim, num_of_experts):
        self.register_parameter("expert_embeddings", nn.Parameter(torch.empty(num_of_experts, int(num_of_experts / 2))))
        self.inp_reduction = nn.Linear(in_embed_dim, int(num_of_experts / 2), bias=False)
        self.temperature = 0.3

    def init_gate_weights(self):
        if getattr(self.args, "init_weight", True) is False:
            return
        gen = torch.Generator(device=self.expert_embeddings.device)
        gen.manual_seed(42)
        nn.init.normal_(self.expert_embeddings, mean=0.0, std=0.02, generator=gen)

    def _cosine(self, mat1, mat2, eps=1e-4):
        if self.theta:
            m1 = mat1.float() / (mat1.norm(p=2, dim=-1, keepdim=True) + self.theta)
        else:
            m1 = F.normalize(mat1.float(), p=2.0, dim=-1, eps=eps)
        return torch.matmul(m1, mat2.float().transpose(0, 1)).type_as(mat1)

    def forward(self, x, return_id_experts=False, is_vision=False):
        B, N, D = x.shape
        E, K = self.num_of_experts, self.num_selected
        x2 = x.reshape(B * N, D)
        lay, w1, b1, w2, b2 = self._stacked_weights()
        # the D -> E/2 projection runs in the router kernel (logits output = x . W^T rounded to x.dtype, like nn.Linear)
        reduced = GateFn.apply(x2, self.inp_reduction.weight, 1, B, False)[0]
        with torch.no_grad():
            norm = self.expert_embeddings.norm(p=2.0, dim=-1, keepdim=True)
            self.expert_embeddings.mul_(1.5 / (norm + self.theta) if self.theta else 1.5 / norm)
        gate_logits = self._cosine(reduced, self.expert_embeddings)
        ok = gate_logits.isfinite()
        gate_logits = torch.where(ok, gate_logits, gate_logits.masked_fill(~ok, float("inf")).min())
        gate_softmax = F.softmax(gate_logits / self.temperature, dim=-1, dtype=torch.float).to(x.dtype)
        # indices from the top-k kernel (ties: lowest index first), values gathered so that autograd sees them
        _, gidx = ops.topk_renorm(gate_softmax.detach().float(), K)
        kept = torch.gather(gate_softmax, 1, gidx.long())
        gw = torch.softmax(kept, dim=-1).float()
        out = self._sparse_ffn(x2, gw, gidx, w1, b1, w2, b2, self._spec(lay))
        aux, info = x.new_zeros(()), {}
        if x.requires_grad:
            aux, bal, z = self.combine_loss(gidx.view(B, N, K), gate_softmax.view(B, N, E), gate_logits.view(B, N, E))
            info = _aux_info(bal, z)
        self.last_routing = (gidx.view(B, N, K), gw.detach().view(B, N, K))
        return out.view(B, N, self.out_embed_dim).to(x.dtype), aux, None, info


@register_moe("xmoe")
class XMOE(_CosineGateLayer):
    """reference: xmoe.py:11-104."""

    def __init__(self, in_embed_dim=768, out_embed_dim=768, num_of_experts=4, num_selected=2, expert=None, args=None):
        MoeLayer.__init__(self, in_embed_dim, out_embed_dim, num_of_experts, num_selected, expert, args)
        self.gate = nn.Linear(in_embed_dim, num_of_experts, bias=True)   # present (and unused) in the reference: xmoe.py:18
        self._init_cosine_gate(in_embed_dim, num_of_experts)
        self.bias = None
        self.init_gate_weights()


@register_moe("smoe_perturbed")
class MoEPerturbedCosingGating(_CosineGateLayer):
    """reference: smoe_perturbed.py:9-144 (perturbed cosine gate, theta added to both norms)."""

    def __init__(self, in_embed_dim=768, out_embed_dim=768, num_of_experts=4, num_selected=2, expert=None, args=None,
                 theta=0.1):
        MoeLayer.__init__(self, in_embed_dim, out_embed_dim, num_of_experts, num_selected, expert, args)
        self.theta = theta
        self._init_cosine_gate(in_embed_dim, num_of_experts)
        self.init_gate_weights()


class _SharedExpertLayer(MoeLayer):
    """shard_smoe.py:12-67 / deepseekv3.py:12-56: the last expert sees every token (a dense 1-expert grouped GEMM), the
    other E-1 are routed top-(k-1); the gate has E-1 outputs and the losses count E-1 experts."""
    shared_scale = 1.0
    routed_scale = 1.0
    always_aux = False

    def __init__(self, in_embed_dim=768, out_embed_dim=768, num_of_experts=4, num_selected=2, expert=None, args=None):
        # the reference calls MoeLayer.__init__() with defaults and then overrides every attribute
        MoeLayer.__init__(self, in_embed_dim, out_embed_dim, num_of_experts, num_selected, expert, args)
        self.num_selected, self.num_of_experts = self.num_selected - 1, self.num_of_experts - 1
        self.gate = nn.Linear(in_embed_dim, self.num_of_experts, bias=False)
        self.init_gate_weights()

    def forward(self, x, return_id_experts=False, is_vision=False):
        B, N, D = x.shape
        E, K = self.num_of_experts, self.num_selected       # routed experts / routed choices
        x2 = x.reshape(B * N, D)
        lay, w1, b1, w2, b2 = self._stacked_weights()       # [E + 1, ...]: routed experts then the shared one
        spec = self._spec(lay)
        want_aux = bool(self.always_aux or x.requires_grad)
        logits, probs, gw, gidx, losses = GateFn.apply(x2, self.gate.weight, K, B, want_aux)
        sl = lambda t, a, b: None if t is None else t[a:b]   # noqa: E731
        routed = self._sparse_ffn(x2, gw, gidx, sl(w1, 0, E), sl(b1, 0, E), sl(w2, 0, E), sl(b2, 0, E), spec)
        shared = DenseFFNFn.apply(x2, sl(w1, E, E + 1), sl(b1, E, E + 1), sl(w2, E, E + 1), sl(b2, E, E + 1), spec)[:B * N]
        if self.shared_scale == 1.0:
            out = shared + routed                                                     # deepseekv3.py:45
        else:
            out = shared * self.shared_scale + routed * self.routed_scale            # shard_smoe.py:55
        aux, info = x.new_zeros(()), {}
        if want_aux:
            aux = losses[0] * self.args.balance_loss_coef + losses[1] * self.args.router_z_loss_coef
            info = _aux_info(losses[0], losses[1])
        self.last_routing = (gidx.view(B, N, K), gw.detach().view(B, N, K))
        out = out.view(B, N, self.out_embed_dim).to(x.dtype)
        if return_id_experts and self.always_aux:
            return out, aux, probs.view(B, N, E)                                      # deepseekv3.py:53-54
        return out, aux, None, info


@register_moe("smoe_share")
class MoEShareLayer(_SharedExpertLayer):
    """reference: shard_smoe.py:12-67 (output = 0.5 * shared + 0.5 * routed)."""
    shared_scale = 0.5
    routed_scale = 0.5


@register_moe("deepseekv3")
class MoEShareLayerV3(_SharedExpertLayer):
    """reference: deepseekv3.py:12-56 (output = shared + routed; the losses are computed on every call)."""
    always_aux = True

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.routed_scaling_factor = 2.5
