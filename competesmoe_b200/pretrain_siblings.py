"""Sibling routers of the language-pretraining plugin on the same kernels.

reference: moe_pretrain_model/layers/moe/{smoe.py, smoeut_norm.py, xmoe.py, smoe_perturbed.py, deepseekv2.py, deepseekv3.py}
Same registry names (smoe, smoe_sigmoid, xmoe, smoe_perturbed, deepseekv2, deepseekv3), constructor keywords, parameter
names (w_gate, keys, values, expert_embeddings, expert_sel, keys_shared, values_shared, e_score_correction_bias) and
regulariser names as the reference classes; the two CVMM calls become the fused permute -> grouped GEMM -> activation
-> grouped GEMM -> combine path of pretrain.MoE.compute_moe_main.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops
from .functional import DenseFFNFn, GateFn
from .multimodal import TopkRenormFn
from .pretrain import MoE, register_moe


class _SiblingBase(MoE):
    reg_suffix = "_ebalance"

    def _finish(self, x, out, gate_logits, selected, weights, all_probs):
        """Common tail of every sibling's forward (smoe.py:246-262)."""
        self.layer += 1
        self.was_training = self.training
        res = out.view(*x.shape[:-1], self.v_dim)
        if self.o_bias is not None:
            res = res + self.o_bias
        lg = gate_logits.view(*x.shape[:-1], -1)
        self.add_reg(lambda: self.entropy_balance(lg) * (self.args.balance_loss_coef / self.div), f"{self.name_moe}{self.reg_suffix}")
        lead = x.shape[:-1]
        self.last_routing = (selected.view(*lead, -1), weights.detach().view(*lead, -1))
        if getattr(self.args, "test_only", False):
            self.add_dist_experts(selection=selected)
            self.add_dist_weight(weight=weights)
            self.add_dist_weight(weight=all_probs, is_all=True)
        return res


@register_moe("smoe")
class SMoeLayer(_SiblingBase):
    """reference: smoe.py:38-264 -- softmax gate, top-k renormalised (the base sigma-MoE forward)."""

    def forward(self, x, return_id_experts=False, return_full=True, *args, **kwargs):
        cdt = self._compute_dtype(x)
        x2 = x.reshape(-1, x.shape[-1])
        logits, probs, gw, gidx = self.compute_gate(x2, cdt, self._x_dtype or x.dtype)     # smoe.py:238 `.to(x.dtype)`
        out = self.compute_moe_main(x2, gidx, gw, cdt)
        return self._finish(x, out, logits, gidx, gw, probs)


@register_moe("smoe_sigmoid")
class SMoEUTNorm(_SiblingBase):
    """reference: smoeut_norm.py:36-150 -- top-k of sigmoid(logits), renormalised; the regulariser is logged as
    `mlp_balance` (sic)."""
    reg_suffix = "_balance"

    def forward(self, x, return_id_experts=False, return_full=True, *args, **kwargs):
        cdt = self._compute_dtype(x)
        x2 = x.reshape(-1, x.shape[-1])
        logits, probs, _, _ = self.compute_gate(x2, cdt)
        gw, gidx = TopkRenormFn.apply(logits.float(), self.num_selected, True, self._x_dtype or x.dtype)
        out = self.compute_moe_main(x2, gidx, gw, cdt)
        return self._finish(x, out, logits, gidx, gw, torch.sigmoid(logits.float()).softmax(dim=-1))


class _CosineGate(_SiblingBase):
    """Shared body of xmoe.py:37-197 and smoe_perturbed.py:39-197 (cosine gate over an E/2-dimensional projection)."""
    theta = 0.0
    _inplace_params = ("expert_embeddings",)   # rescaled by every forward call (restored after graph capture)

    def _init_cosine(self, sel_bias: bool):
        self.reduction_dim = int(self.n_experts / 2)
        emb = torch.empty(self.num_of_experts, self.reduction_dim)
        torch.nn.init.orthogonal_(emb, gain=0.32)
        self.register_parameter("expert_embeddings", torch.nn.Parameter(emb))
        self.temperature = 0.3
        self.bias = None
        self.expert_sel = torch.nn.Parameter(torch.empty(self.reduction_dim, self.k_vec_dim))
        self.sel_bias = torch.nn.Parameter(torch.zeros(self.reduction_dim)) if sel_bias else None
        torch.nn.init.normal_(self.expert_sel, std=self.k_vec_dim ** -0.5 * self.sel_weight_scale)

    def _cosine(self, mat1, mat2, eps=1e-4):
        if self.theta:
            m1 = mat1.float() / (mat1.norm(p=2, dim=-1, keepdim=True) + self.theta)
        else:
            m1 = F.normalize(mat1.float(), p=2.0, dim=-1, eps=eps)
        return torch.matmul(m1, mat2.float().transpose(0, 1)).type_as(mat1)

    def _cosine_gate_logits(self, x2, cdt):
        """compute_gate of xmoe.py / smoe_perturbed.py:149-160: cosine of the reduced input with the (rescaled in place)
        expert embeddings, non-finite scores replaced by the smallest finite one."""
        reduced = GateFn.apply(self._cast(x2, cdt), self.expert_sel, 1, 1, False)[0]          # D -> E/2 projection in the router kernel
        if self.sel_bias is not None:
            reduced = reduced + self.sel_bias.to(reduced.dtype)
        with torch.no_grad():
            norm = self.expert_embeddings.norm(p=2.0, dim=-1, keepdim=True)
            self.expert_embeddings.mul_(1.5 / (norm + self.theta) if self.theta else 1.5 / norm)
        gate_logits = self._cosine(reduced, self.expert_embeddings)
        ok = gate_logits.isfinite()
        return torch.where(ok, gate_logits, gate_logits.masked_fill(~ok, float("inf")).min())

    def att_forward(self, x, n_experts, n_copies, return_full=True, *args, **kwargs):
        """smoe_perturbed.py:199-223 -- the one att_forward the reference ships live (the base class's is commented out,
        moe.py:456-486): per-head softmax(cosine gate / temperature) in the layer input's dtype, top-k per head, softmax
        of the kept probabilities as weights, CVMM selection over the shifted (head, expert) indices.  The reference also
        appends the gate tensors to three lists that `before_loss` empties unread (moe.py:259-267, :341-358): dropped."""
        from .cvmm import cvmm_prepare_sel2
        from .pretrain import Selection
        assert self.is_att, "att_forward needs a layer built with is_att=True"
        if self.selection_dropout > 0 and self.training:
            x = F.dropout(x, self.selection_dropout)
        cdt = self._compute_dtype(x)
        lead = x.shape[:-1]
        gate_logits = self._cosine_gate_logits(x.reshape(-1, x.shape[-1]), cdt).view(*lead, n_copies, -1)
        gate_softmax = F.softmax(gate_logits / self.temperature, dim=-1, dtype=torch.float).to(self._x_dtype or x.dtype)
        with torch.no_grad():      # torch.topk(sorted=False) leaves ties open; here: highest first, lowest index on ties
            _, idx = ops.topk_renorm(gate_softmax.detach().float().reshape(-1, gate_softmax.shape[-1]), self.num_selected)
            idx = idx.view(*lead, n_copies, self.num_selected).long()
        val = torch.softmax(torch.gather(gate_softmax, -1, idx), dim=-1)
        shift = (torch.arange(n_copies, device=idx.device, dtype=idx.dtype) * n_experts).unsqueeze(-1)
        sel_pp = cvmm_prepare_sel2((shift + idx).flatten(-2, -1).int(), val, n_experts=self.n_experts)
        return Selection(gate_logits, val, idx, sel_pp)

    def forward(self, x, return_id_experts=False, return_full=True, *args, **kwargs):
        cdt = self._compute_dtype(x)
        K = self.num_selected
        x2 = x.reshape(-1, x.shape[-1])
        gate_logits = self._cosine_gate_logits(x2, cdt)
        gate_softmax = F.softmax(gate_logits / self.temperature, dim=-1, dtype=torch.float).to(self._x_dtype or x.dtype)
        _, gidx = ops.topk_renorm(gate_softmax.detach().float(), K)
        kept = torch.gather(gate_softmax, 1, gidx.long())
        gw = torch.softmax(kept, dim=-1).float()
        out = self.compute_moe_main(x2, gidx, gw, cdt)
        return self._finish(x, out, gate_logits, gidx, gw, gate_softmax)


@register_moe("xmoe")
class XMOE(_CosineGate):
    def __init__(self, *a, sel_bias: bool = False, **kw):
        super().__init__(*a, sel_bias=sel_bias, **kw)
        self._init_cosine(sel_bias)


@register_moe("smoe_perturbed")
class MoEPerturbedCosingGating(_CosineGate):
    theta = 0.1

    def __init__(self, *a, sel_bias: bool = False, **kw):
        super().__init__(*a, sel_bias=sel_bias, **kw)
        self._init_cosine(sel_bias)


class _SharedExpert(_SiblingBase):
    """deepseekv2.py:96-181 / deepseekv3.py:96-190: one shared sigma-MoE expert (keys_shared / values_shared) sees every
    token with weight 1 -- a dense one-expert grouped GEMM pair -- next to the routed experts."""

    def _init_shared(self, dmodel, weight_scale, bias):
        self.n_shared_experts = 1
        hs = self.expert_size * self.n_shared_experts
        self.values_shared = torch.nn.Parameter(torch.empty(1, hs, self.v_dim))
        torch.nn.init.normal_(self.values_shared, std=hs ** -0.5 * weight_scale)
        self.keys_shared = torch.nn.Parameter(torch.empty(1, self.k_vec_dim, hs))
        torch.nn.init.normal_(self.keys_shared, std=dmodel ** -0.5 * weight_scale)
        self.bias_shared = torch.nn.Parameter(torch.zeros(1, hs)) if bias else None

    def _shared_out(self, x2, cdt):
        y = DenseFFNFn.apply(self._cast(x2, cdt), self.keys_shared, self.bias_shared, self.values_shared, None, self._spec(cdt))
        return y[:x2.shape[0]]


@register_moe("deepseekv2")
class DeepSeekV2(_SharedExpert):
    """reference: deepseekv2.py:38-181 -- top-k of the raw logits, softmax over the kept logits."""

    def __init__(self, dmodel, *a, weight_scale: float = 1.0, bias: bool = False, **kw):
        super().__init__(dmodel, *a, weight_scale=weight_scale, bias=bias, **kw)
        self._init_shared(dmodel, weight_scale, bias)

    def forward(self, x, return_id_experts=False, return_full=True, *args, **kwargs):
        cdt = self._compute_dtype(x)
        x2 = x.reshape(-1, x.shape[-1])
        logits, probs, _, _ = self.compute_gate(x2, cdt)
        _, gidx = ops.topk_renorm(logits.detach().float(), self.num_selected)
        gw = F.softmax(torch.gather(logits, 1, gidx.long()), dim=-1).to(self._x_dtype or x.dtype).float()
        out = self.compute_moe_main(x2, gidx, gw, cdt) + self._shared_out(x2, cdt)
        return self._finish(x, out, logits, gidx, gw, probs)


@register_moe("deepseekv3")
class DeepSeekV3(_SharedExpert):
    """reference: deepseekv3.py:38-190 -- top-k of sigmoid(logits) divided by their sum (+1e-20), scaling factor 1;
    `e_score_correction_bias` exists as a parameter and is not used by the forward."""

    def __init__(self, dmodel, *a, weight_scale: float = 1.0, bias: bool = False, **kw):
        super().__init__(dmodel, *a, weight_scale=weight_scale, bias=bias, **kw)
        self._init_shared(dmodel, weight_scale, bias)
        self.e_score_correction_bias = torch.nn.Parameter(torch.zeros(self.n_experts))
        self.n_group, self.topk_group, self.routed_scaling_factor, self.e_score_correction_bias_scale = 8, 4, 1, 0.001

    def forward(self, x, return_id_experts=False, return_full=True, *args, **kwargs):
        cdt = self._compute_dtype(x)
        x2 = x.reshape(-1, x.shape[-1])
        logits, probs, _, _ = self.compute_gate(x2, cdt)
        sig = torch.sigmoid(logits)
        _, gidx = ops.topk_renorm(sig.detach().float(), self.num_selected)
        kept = torch.gather(sig, 1, gidx.long())
        gw = (kept / (kept.sum(dim=-1, keepdim=True) + 1e-20) * self.routed_scaling_factor).float()
        out = self.compute_moe_main(x2, gidx, gw, cdt) + self._shared_out(x2, cdt)
        return self._finish(x, out, logits, gidx, gw, probs)
