"""Drop-in for the multimodal MoE plugin (reference: moe_model/model/moe/{moe.py,competesmoe.py,register.py}).

Same constructor, forward signature, return 4-tuple, `args` attribute names, schedule methods and checkpoint layout
(`gate.weight`, `experts.{e}.<child>.{weight,bias}`, buffer `prob_flips`) as the reference; the arithmetic runs in the
libcsmoe CUDA kernels (no eager expert loop, no host syncs per expert).  See INTEGRATION.md for how the class is
registered under the reference's own `MOE_REGISTRY`.
"""
from __future__ import annotations

import copy
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import experts as X
from . import ops
from .functional import (AffinityFn, CompeteLossesFn, CompeteTailFn, DenseFFNFn, FFNSpec, GatherRowsFn, GateFn,
                         SelectCombineFn, SparseFFNFn)
from .graphs import capture_guard
from .schedule import make_layer_schedule

MOE_REGISTRY: Dict[str, type] = {}


def register_moe(*names):
    """Same contract as moe_model/model/moe/register.py:5-16, except that the decorated class is returned."""
    def decorate(cls):
        for name in names:
            if name in MOE_REGISTRY and MOE_REGISTRY[name] != cls:
                raise AssertionError(f"Model named '{name}' conflicts with existing model! {cls} vs {MOE_REGISTRY[name]}")
            MOE_REGISTRY[name] = cls
        return cls
    return decorate


def get_moe(model_name):
    try:
        return MOE_REGISTRY[model_name]
    except KeyError:
        raise ValueError(f"Attempted to load moe method'{model_name}', but no model for this name found! "
                         f"Supported model names: {', '.join(MOE_REGISTRY.keys())}")


class TopkRenormFn(Function):
    """(scores [T,E] f32) -> (w [T,K] f32, idx [T,K] i32): top-k of the affinity scores, renormalised
    (competesmoe.py:249-254).  Differentiable w.r.t. the scores: the weights are *not* detached in the reference."""

    @staticmethod
    def forward(ctx, scores, k: int, sigmoid: bool, dtype: torch.dtype):
        w, idx = ops.topk_renorm(scores, k, sigmoid=sigmoid, round_dtype=dtype, round_out=dtype == torch.bfloat16)
        ctx.save_for_backward(scores, w, idx)
        ctx.sigmoid = sigmoid
        ctx.mark_non_differentiable(idx)
        return w, idx

    @staticmethod
    @once_differentiable
    def backward(ctx, dw, _):
        scores, w, idx = ctx.saved_tensors
        return ops.topk_renorm_bwd(scores, w, idx, dw, ctx.sigmoid), None, None, None


class _CombineRowsFn(Function):
    """out[t] = sum_k w[t, k] * y[rows[t * K + k]] in ascending-expert order (compute_moe with the per-slot outputs given)."""

    @staticmethod
    def forward(ctx, y, w, sel, rows, spec: FFNSpec):
        T, K = sel.shape
        out = ops.combine_fwd(y.contiguous(), rows, sel.reshape(-1).contiguous(), w, T, K, round_each=spec.round_each,
                              round_w=spec.round_w)
        ctx.save_for_backward(y, w, rows)
        ctx.dims = (T, K, spec)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        y, w, rows = ctx.saved_tensors
        T, K, spec = ctx.dims
        dout = dout.contiguous().to(y.dtype)
        dw = ops.combine_bwd_w(y.contiguous(), dout, rows, T, K)
        wu = w.to(torch.bfloat16).float() if spec.round_w else w
        dy = (wu.reshape(T * K, 1) * dout.float().repeat_interleave(K, dim=0)).to(y.dtype)       # rows = identity here
        return dy, dw, None, None, None


class MoeLayer(nn.Module):
    """Base layer: gate + experts + the loss helpers every router shares (reference: moe.py:8-132,214-226)."""

    def __init__(self, in_embed_dim=768, out_embed_dim=768, num_of_experts=4, num_selected=2, expert=None, args=None):
        nn.Module.__init__(self)   # explicit: under integrate.bind_multimodal the next class in the MRO is the reference's MoeLayer
        self.in_embed_dim = in_embed_dim
        self.out_embed_dim = out_embed_dim
        self.num_of_experts = num_of_experts
        self.num_selected = num_selected
        if expert is None:
            self.experts = nn.ModuleList([
                nn.Sequential(nn.Linear(in_embed_dim, out_embed_dim), nn.GELU(), nn.Linear(out_embed_dim, out_embed_dim))
                for _ in range(num_of_experts)])
        elif isinstance(expert, nn.ModuleList):
            self.experts = expert
        else:
            self.experts = nn.ModuleList([copy.deepcopy(expert) for _ in range(num_of_experts)])
        self.gate = nn.Linear(in_embed_dim, num_of_experts, bias=False)
        self.args = args
        self.is_vision = False
        self.log_metrics = {}
        self._layout: Optional[X.ExpertLayout] = None
        self._ep = None            # ep.EPLayerState once enable_expert_parallel() was called
        self.ep_expert_offset = 0  # global index of experts[0] under expert parallelism
        self._graphs = None        # {key: captured step} once enable_cuda_graphs() was called

    # ---- CUDA graphs (no counterpart in the reference, whose step has a host sync per expert and per layer)
    def enable_cuda_graphs(self, enabled: bool = True):
        """Opt in: training-mode calls with a CUDA input that requires grad are replayed from captured CUDA graphs (one
        forward graph + one backward graph per (branch, shape, dtype), torch.cuda.make_graphed_callables), removing the
        per-kernel launch gaps of the ~40-launch step; CUDA autocast state is part of the graph key.  Everything else (eval,
        no-grad, expert parallelism, return_id_experts) takes the normal path.  Parameters may change value between calls, not storage."""
        self._graphs = {} if enabled else None
        # layers without their own graph dispatch (the sibling routers) get it through an instance-level wrapper
        if not hasattr(type(self), "_forward_impl"):
            if enabled and "_eager_forward" not in self.__dict__:
                self._eager_forward = self.forward
                self.forward = self._wrapped_graph_forward
            elif not enabled and "_eager_forward" in self.__dict__:
                self.forward = self._eager_forward
                del self._eager_forward
        return self

    def __getstate__(self):
        """copy.deepcopy / pickling: captured CUDA graphs belong to this instance's storage and are not copied -- the
        copy keeps graph mode switched on and captures its own graphs on first use."""
        state = self.__dict__.copy()
        if state.get("_graphs") is not None:
            state["_graphs"] = {}
        return state

    _inplace_params = ()     # parameters a forward call rescales in place (xmoe / smoe_perturbed: expert_embeddings)

    def _wrapped_graph_forward(self, x, return_id_experts=False, is_vision=False):
        if not self._graph_eligible(x, return_id_experts) or self._ep is not None:
            return self._eager_forward(x, return_id_experts, is_vision)
        self._stacked_weights()
        params = tuple(p for p in self.parameters() if p.requires_grad)
        ac_on, ac_dt = self._autocast_state()
        key = (bool(is_vision), tuple(x.shape), x.dtype, ac_on, ac_dt, tuple(p.data_ptr() for p in params))
        entry = self._graphs.get(key)
        if entry is None:
            names: List[str] = []
            # capture runs the forward several times: put back what it rescales so that the first replay is call no. 1
            keep = {n: getattr(self, n).detach().clone() for n in self._inplace_params}

            def fn(xx, *_params):
                with torch.autocast("cuda", dtype=ac_dt, enabled=ac_on, cache_enabled=False):
                    out, aux, _, info = self._eager_forward(xx, False, is_vision)
                names[:] = sorted(info)
                return (out, aux) + tuple(info[k] for k in names)

            sample = x.detach().clone().requires_grad_(True)
            with capture_guard(), torch.autocast("cuda", dtype=ac_dt, enabled=ac_on, cache_enabled=False):
                graphed = torch.cuda.make_graphed_callables(fn, (sample,) + params, allow_unused_input=True)
            with torch.no_grad():
                for n, v in keep.items():
                    getattr(self, n).copy_(v)
            entry = (graphed, list(names), self.last_routing)
            self._graphs[key] = entry
        graphed, names, routing = entry
        res = graphed(x, *params)
        self.last_routing = routing
        return res[0], res[1], None, dict(zip(names, res[2:]))

    def _graph_eligible(self, x, return_id_experts) -> bool:
        return (self._graphs is not None and self.training and x.is_cuda and x.requires_grad and x.numel() > 0
                and torch.is_grad_enabled() and not return_id_experts
                and not torch.cuda.is_current_stream_capturing())

    @staticmethod
    def _autocast_state():
        """(enabled, dtype) of CUDA autocast: part of the graph key, and re-entered (weight cache off) around the capture
        so that the captured forward sees what the eager call would."""
        return torch.is_autocast_enabled(), torch.get_autocast_dtype("cuda")

    def _graphed_call(self, x, branch: bool):
        self._stacked_weights()          # storage fusing re-points expert .data once: do it before keying on pointers
        params = tuple(p for p in self.parameters() if p.requires_grad)
        ac_on, ac_dt = self._autocast_state()
        key = (branch, tuple(x.shape), x.dtype, ac_on, ac_dt, tuple(p.data_ptr() for p in params))
        entry = self._graphs.get(key)
        if entry is None:
            names: List[str] = []

            def fn(xx, *_params):
                with torch.autocast("cuda", dtype=ac_dt, enabled=ac_on, cache_enabled=False):
                    out, aux, _, info = self._forward_impl(xx, False)
                names[:] = sorted(info)
                return (out, aux) + tuple(info[k] for k in names)

            sample = x.detach().clone().requires_grad_(True)
            with capture_guard(), torch.autocast("cuda", dtype=ac_dt, enabled=ac_on, cache_enabled=False):
                graphed = torch.cuda.make_graphed_callables(fn, (sample,) + params)
            entry = (graphed, names, self.last_routing)
            self._graphs[key] = entry
        graphed, names, routing = entry
        res = graphed(x, *params)
        self.last_routing = routing     # static tensors of this graph, refreshed by the replay
        return res[0], res[1], None, dict(zip(names, res[2:]))

    # ---- expert parallelism (no counterpart in the reference, which is data-parallel only)
    def enable_expert_parallel(self, group, max_tokens: int, row_tile: int = 256):
        """Shard the experts over `group` (a competesmoe_b200.ep.EPGroup): this rank keeps experts
        [rank*E/P, (rank+1)*E/P) as `experts.{0..E/P-1}`; the gate stays replicated.  `max_tokens` = the largest B*N this
        rank will ever pass to forward (sizes the peer-mapped exchange buffers).  Call after loading a full checkpoint and
        BEFORE the optimizer is built (the other experts' parameters leave the module).  Afterwards `state_dict()` holds
        this rank's experts only, renumbered from 0; use `full_state_dict()` / `load_full_state_dict()` for checkpoints
        in the reference layout (`experts.{global e}...`)."""
        from .ep import EPLayerState
        self._shard_experts(group.rank, group.world)
        self._ep = EPLayerState(group, self.num_of_experts, self.num_selected, self.in_embed_dim, self.out_embed_dim,
                                max_tokens, row_tile)
        return self

    def _shard_experts(self, rank: int, world: int):
        E = self.num_of_experts
        if E % world != 0:
            raise ValueError(f"{E} experts cannot be split over an expert-parallel group of {world} ranks")
        El = E // world
        self.ep_expert_offset = rank * El
        self.experts = nn.ModuleList([self.experts[i] for i in range(self.ep_expert_offset, self.ep_expert_offset + El)])
        self._layout = None

    def full_state_dict(self, group=None):
        """Collective: the reference-layout state dict with every expert (ep.full_state_dict)."""
        from .ep import full_state_dict
        return full_state_dict(self, group)

    def load_full_state_dict(self, state_dict, group=None, strict: bool = True):
        from .ep import load_full_state_dict
        return load_full_state_dict(self, state_dict, group, strict)

    def _sparse_ffn(self, x2, gw, gidx, w1, b1, w2, b2, spec):
        if self._ep is None:
            return SparseFFNFn.apply(x2, gw, gidx, w1, b1, w2, b2, spec)
        from .ep import EPSparseFFNFn
        return EPSparseFFNFn.apply(x2, gw, gidx, w1, b1, w2, b2, spec, self._ep)

    def _all_expert_weights(self, w1, b1, w2, b2):
        """Competition step under expert parallelism: every rank needs every expert (all-gather, grads reduce-scatter)."""
        if self._ep is None or self._ep.group.world == 1:
            return w1, b1, w2, b2
        from .ep import gather_experts
        g = self._ep.group
        return gather_experts(w1, g), gather_experts(b1, g), gather_experts(w2, g), gather_experts(b2, g)

    # ---- initialisation (moe.py:50-70)
    def init_gate_weights(self, std=0.02):
        if getattr(self.args, "init_weight", True) is False:
            return
        device = self.gate.weight.device if self.gate.weight.device != torch.device("meta") else torch.device("cpu")
        gen = torch.Generator(device=device)
        gen.manual_seed(42)
        nn.init.normal_(self.gate.weight, mean=0.0, std=std, generator=gen)

    # ---- expert parameters as stacked tensors
    def _expert_layout(self) -> X.ExpertLayout:
        if self._layout is None:
            layouts = [X.describe_expert(m) for m in self.experts]
            if any(l != layouts[0] for l in layouts):
                raise NotImplementedError("all experts of a layer must share one architecture")
            self._layout = layouts[0]
        return self._layout

    def _stacked_weights(self):
        lay = self._expert_layout()
        lin1 = [m.get_submodule(lay.first) for m in self.experts]
        lin2 = [m.get_submodule(lay.second) for m in self.experts]
        groups = [[l.weight for l in lin1], [l.bias for l in lin1], [l.weight for l in lin2], [l.bias for l in lin2]]
        for g in groups:
            if g[0] is not None:
                X.fuse_storage(g)
        w1, b1, w2, b2 = (X.stack_params(g) for g in groups)
        return lay, w1, b1, w2, b2

    # ---- losses (moe.py:71-110); inputs are [B, N, E] / [B, N, K]
    def zloss(self, gate_logits, gate_softmax=None):
        return torch.square(torch.logsumexp(gate_logits.float(), dim=-1)).mean()

    def balanceloss(self, selected_experts, gate_softmax):
        proxy = gate_softmax.mean(dim=-2)
        top1 = F.one_hot(selected_experts[..., 0].long(), self.num_of_experts).to(gate_softmax.dtype)
        return (proxy * top1.mean(dim=-2)).mean() * float(self.num_of_experts ** 2)

    def combine_loss(self, selected_experts, gate_softmax, gate_logits, acitve_zloss=True):
        balance_loss = self.balanceloss(selected_experts=selected_experts, gate_softmax=gate_softmax)
        router_z_loss = torch.zeros((), device=gate_softmax.device)
        if acitve_zloss:
            router_z_loss = self.zloss(gate_logits, gate_softmax)
            aux = balance_loss * self.args.balance_loss_coef + router_z_loss * self.args.router_z_loss_coef
        else:
            aux = balance_loss * self.args.balance_loss_coef
        return aux, balance_loss, router_z_loss

    def experts_diversity_loss(self, expert_outputs):
        """[B*N (or B, N), K, D] -> mean of the off-diagonal cosine similarities (competesmoe.py:180-218)."""
        eo = expert_outputs.to(torch.float32)
        K, D = eo.shape[-2:]
        nrm = F.normalize(eo, p=2, dim=-1).reshape(-1, K, D)
        sim = torch.bmm(nrm, nrm.transpose(1, 2))
        sim = sim * (1 - torch.eye(K, device=eo.device))
        return sim.mean()

    def _spec(self, lay: X.ExpertLayout, x: Optional[torch.Tensor] = None) -> FFNSpec:
        # fp32 activations outside autocast (an fp32 module called as is): fp32-accurate expert products, what the eager
        # reference computes; everything else runs the bf16 tensor-core path
        fp32 = x is not None and x.dtype == torch.float32 and not torch.is_autocast_enabled()
        return FFNSpec(act=lay.act, kn_layout=False, round_each=True, round_w=False, fp32=fp32)

    # ---- the reference's policy-level methods, same names and signatures, on the kernels.  The layers' own forward does
    # not go through them (it uses the fused forms of the same steps); they are here for callers and subclasses that do.
    def topk_expert(self, gate_logits):
        """moe.py:113-132: (top-k softmax probabilities -- not renormalised --, their indices, the full softmax).
        Ties: lowest expert index first (torch.topk leaves them open)."""
        gate_softmax = F.softmax(gate_logits, dim=-1, dtype=torch.float32)
        _, idx = ops.topk_renorm(gate_softmax.detach().reshape(-1, gate_softmax.shape[-1]).contiguous(), self.num_selected)
        selected_experts = idx.long().view(*gate_softmax.shape[:-1], self.num_selected)
        return torch.gather(gate_softmax, -1, selected_experts), selected_experts, gate_softmax

    def compute_moe(self, selected_experts, weights, results, x, expert_outputs=None, return_topk_outputs=False):
        """moe.py:172-213: results += sum_k weights[..., k] * expert_{selected[..., k]}(x), experts visited in ascending
        order with the running sum rounded to x's dtype (one fused dispatch / grouped GEMM / combine pass here).
        expert_outputs: per-expert outputs on all tokens ([E][B, N, D_out]) to combine instead of running the experts.
        return_topk_outputs: also the diversity loss of the selected experts' outputs, returned as (results, loss) when x
        requires grad and logged to `log_metrics['diver_loss']` otherwise, like the reference."""
        B, N, D = x.shape
        T, K = B * N, selected_experts.shape[-1]
        sel = selected_experts.reshape(T, K).to(torch.int32).contiguous()
        w = weights.reshape(T, K).float()
        x2 = x.reshape(T, D)
        lay, w1, b1, w2, b2 = self._stacked_weights()
        spec = self._spec(lay, x2)
        topk_out = None
        if expert_outputs is not None:
            y = (torch.stack(list(expert_outputs), 0) if not torch.is_tensor(expert_outputs) else expert_outputs)
            y = y.reshape(self.num_of_experts * T, -1).contiguous()
            out = SelectCombineFn.apply(y, w, sel, T, spec)
            if return_topk_outputs:
                topk_out = GatherRowsFn.apply(y, sel, T)
        elif return_topk_outputs:
            # the selected experts' outputs themselves are needed: one sparse pass per selection rank with weight 1
            one = torch.ones(T, 1, dtype=torch.float32, device=x.device)
            cols = [self._sparse_ffn(x2, one, sel[:, k:k + 1].contiguous(), w1, b1, w2, b2, spec) for k in range(K)]
            topk_out = torch.stack(cols, 1)                                         # [T, K, D_out]
            rows = torch.arange(T * K, dtype=torch.int32, device=x.device)
            out = _CombineRowsFn.apply(topk_out.reshape(T * K, -1), w, sel, rows, spec)
        else:
            out = self._sparse_ffn(x2, w, sel, w1, b1, w2, b2, spec)
        results += out.view(B, N, -1).to(results.dtype)
        if topk_out is not None:
            diver_loss = self.experts_diversity_loss(topk_out.view(B, N, K, -1))
            if not x.requires_grad:
                if not hasattr(self, "log_metrics"):
                    self.log_metrics = {}
                self.log_metrics["diver_loss"] = diver_loss.item()
            else:
                return results, diver_loss
        return results

    def forward(self, x, return_id_experts=False):
        """Plain sparse MoE (moe.py:228-246)."""
        B, N, D = x.shape
        x2 = x.reshape(B * N, D)
        lay, w1, b1, w2, b2 = self._stacked_weights()
        logits, probs, gw, gidx, losses = GateFn.apply(x2, self.gate.weight, self.num_selected, B, True)
        out = self._sparse_ffn(x2, gw, gidx, w1, b1, w2, b2, self._spec(lay, x2)).view(B, N, self.out_embed_dim)
        balance_loss, router_z_loss = losses[0], losses[1]
        aux = balance_loss * self.args.balance_loss_coef + router_z_loss * self.args.router_z_loss_coef
        infor_aux = {"balance_loss": balance_loss.detach().clone(), "router_z_loss": router_z_loss.detach().clone()}
        if return_id_experts:
            return out, aux, probs.view(B, N, -1)
        return out, aux, None, infor_aux


@register_moe("competesmoe", "competesmoe_b200")
class CompeteSMoE(MoeLayer):
    """CompeteSMoE layer (reference: moe_model/model/moe/competesmoe.py:9-415)."""

    def __init__(self, in_embed_dim=768, out_embed_dim=768, num_of_experts=4, num_selected=2, expert=None, args=None):
        MoeLayer.__init__(self, in_embed_dim, out_embed_dim, num_of_experts, num_selected, expert, args)
        if args is None or not hasattr(args, "rate_flip"):
            raise ValueError("The 'args' parameter must have the attribute 'rate_flip'.")
        if not hasattr(args, "warm_up"):
            raise ValueError("The 'args' parameter must include 'warm_up'.")
        self.warm_up = args.warm_up
        self.rate_flip = args.rate_flip
        self.total_steps = None
        self.current_steps = 0
        self.step_warm = None
        self.is_prob_flips = True
        self.register_buffer("prob_flips", torch.zeros(15801))
        self._flips_host = None
        self._flips_key = None
        self.last_routing = None   # (selected [B,N,K] i32, weights [B,N,K] f32) of the last forward, for tests / logging
        self.init_gate_weights()

    # ---- schedule (competesmoe.py:35-179)
    def set_total_steps(self, total_steps, id_layer, prob_flips_final):
        assert id_layer is not None, "You must setup id layer is not None"
        assert prob_flips_final is not None, "You must setup prob_flips_final is not None"
        self.total_steps = total_steps
        self.step_warm, flags = make_layer_schedule(total_steps, self.warm_up, self.rate_flip,
                                                    self.args.max_compete_in_iter, prob_flips_final)
        self.flip_steps = total_steps - self.step_warm
        prob_flips_final[id_layer] = flags
        self.prob_flips = flags
        self.is_prob_flips = False
        return prob_flips_final

    def set_current_steps(self, step):
        self.current_steps = step

    def _is_competition_step(self, x) -> bool:
        """competesmoe.py:347, without a device->host sync per call: the flag vector is mirrored on the host and
        refreshed only when the buffer object or its version changes (load_state_dict, set_total_steps)."""
        if not x.requires_grad or self.step_warm is None or self.current_steps < self.step_warm:
            return False
        pf = self.prob_flips
        key = (id(pf), pf._version, pf.data_ptr())
        if key != self._flips_key:
            self._flips_host = pf.detach().to("cpu").ne(0).tolist()
            self._flips_key = key
        return bool(self._flips_host[self.current_steps - self.step_warm])

    # ---- policies
    def _gate(self, x2, batch, want_aux):
        return GateFn.apply(x2, self.gate.weight, self.num_selected, batch, want_aux)

    def router_policy(self, x):
        """competesmoe.py:301-320, the reference's signature: x [B, N, D] -> (renormalised top-k weights [B, N, K] fp32,
        selected experts [B, N, K] int64, gate softmax [B, N, E] fp32, gate logits [B, N, E] in x's dtype)."""
        B, N, D = x.shape
        logits, probs, gw, gidx, _ = self._gate(x.reshape(B * N, D), B, False)
        K, E = self.num_selected, self.num_of_experts
        return gw.view(B, N, K), gidx.long().view(B, N, K), probs.view(B, N, E), logits.view(B, N, E)

    def competition_policy(self, x):
        """competesmoe.py:219-259, the reference's signature: every expert on every token, affinity = mean softplus(output)
        held in x's dtype, top-k of it (of its sigmoid under args.norm_sigmoid) renormalised.  Returns (weights [B, N, K],
        selected experts [B, N, K] int64, softmax(affinity) [B, N, E] fp32, affinity [B, N, E] in x's dtype, the selected
        experts' outputs [B, N, K, D_out])."""
        B, N, D = x.shape
        T, E, K = B * N, self.num_of_experts, self.num_selected
        x2 = x.reshape(T, D)
        lay, w1, b1, w2, b2 = self._stacked_weights()
        w1, b1, w2, b2 = self._all_expert_weights(w1, b1, w2, b2)
        spec = self._spec(lay, x2)
        y_all = DenseFFNFn.apply(x2, w1, b1, w2, b2, spec)                          # [E * t_pad, D_out]
        t_pad = y_all.shape[0] // E
        aff = AffinityFn.apply(y_all, E, T, t_pad, x.dtype == torch.bfloat16)       # fp32 tensor of values in x's dtype
        w, idx = TopkRenormFn.apply(aff, K, bool(getattr(self.args, "norm_sigmoid", False)), x.dtype)
        topk_out = GatherRowsFn.apply(y_all, idx, t_pad)
        aff3 = aff.view(B, N, E)
        return (w.to(x.dtype).view(B, N, K), idx.long().view(B, N, K), F.softmax(aff3, dim=-1, dtype=torch.float32),
                aff3.to(x.dtype), topk_out.view(B, N, K, -1))

    def topk_expert_softmax(self, gate_logits):
        """competesmoe.py:262-278: top-k of the logits, softmax over the kept ones."""
        gate_softmax = F.softmax(gate_logits, dim=-1, dtype=torch.float32)
        _, idx = ops.topk_renorm(gate_logits.detach().float().reshape(-1, gate_logits.shape[-1]).contiguous(), self.num_selected)
        selected_experts = idx.long().view(*gate_logits.shape[:-1], self.num_selected)
        return F.softmax(torch.gather(gate_logits, -1, selected_experts), dim=-1, dtype=torch.float), selected_experts, gate_softmax

    def topk_expert_sigmoid(self, gate_logits):
        """competesmoe.py:279-295: top-k of sigmoid(logits), not renormalised."""
        gate_softmax = F.softmax(gate_logits, dim=-1, dtype=torch.float32)
        gate_sigmoid = torch.sigmoid(gate_logits)
        _, idx = ops.topk_renorm(gate_sigmoid.detach().float().reshape(-1, gate_logits.shape[-1]).contiguous(), self.num_selected)
        selected_experts = idx.long().view(*gate_logits.shape[:-1], self.num_selected)
        return torch.gather(gate_sigmoid, -1, selected_experts), selected_experts, gate_softmax

    def router_loss(self, gate_softmax, affinity_softmax):
        return F.mse_loss(gate_softmax, affinity_softmax)

    def forward(self, x, return_id_experts=False, is_vision=False):
        if self._graph_eligible(x, return_id_experts):
            branch = self._is_competition_step(x)
            # under expert parallelism only the router step is captured (kernels + device-side barriers); the competition
            # step gathers the expert weights with NCCL and stays eager
            if self._ep is None or not branch:
                return self._graphed_call(x, branch)
        return self._forward_impl(x, return_id_experts)

    def _forward_empty(self, x, return_id_experts):
        """No tokens (an empty micro-batch): nothing to launch.  The reference runs its ops on the empty tensors: an empty
        output, and every loss it computes on this branch is a mean over zero tokens, i.e. NaN -- reproduced as is."""
        B, N, _ = x.shape
        compete = self._is_competition_step(x)
        out = x.new_zeros(B, N, self.out_embed_dim) + x.sum() * 0          # keeps the output attached to x's graph
        nan = x.new_full((), float("nan"))
        self.last_routing = (torch.zeros(B, N, self.num_selected, dtype=torch.int32, device=x.device),
                             torch.zeros(B, N, self.num_selected, dtype=torch.float32, device=x.device))
        if compete:
            return out, nan, None, {"balance_loss": nan.clone(), "diversity_loss": nan.clone(), "routerloss": nan.clone()}
        if x.requires_grad or return_id_experts:
            return out, nan, None, {"balance_loss": nan.clone(), "router_z_loss": nan.clone()}
        return out, x.new_zeros(()), None, {}

    def _forward_impl(self, x, return_id_experts=False):
        B, N, D = x.shape
        T, E, K = B * N, self.num_of_experts, self.num_selected
        if T == 0 and self._ep is None:      # under expert parallelism a rank without tokens still takes part in the exchange
            return self._forward_empty(x, return_id_experts)
        x2 = x.reshape(T, D)
        lay, w1, b1, w2, b2 = self._stacked_weights()
        spec = self._spec(lay, x2)
        compete = self._is_competition_step(x)
        want_aux = (not compete) and (x.requires_grad or return_id_experts)
        gate_logits, gate_softmax, gate_w, gate_idx, gate_losses = self._gate(x2, B, want_aux)
        auxiliary_loss = x.new_zeros(())
        infor_aux = {}
        if compete:
            w1, b1, w2, b2 = self._all_expert_weights(w1, b1, w2, b2)
            # dense pass; the per-row softplus sums of the score come out of the down projection's epilogue
            y_all, score_sums = DenseFFNFn.apply(x2, w1, b1, w2, b2, spec, x.dtype == torch.bfloat16)   # [E * t_pad, Dout]
            t_pad = y_all.shape[0] // E
            # score, top-k, combine (the selected experts' outputs are reused from the dense pass instead of being
            # recomputed as compute_moe does at competesmoe.py:374) and diversity loss: one autograd node
            aff, aff_w, aff_idx, out, diversity_loss = CompeteTailFn.apply(
                y_all, E, T, t_pad, K, bool(getattr(self.args, "norm_sigmoid", False)), x.dtype, spec, score_sums)
            # softmax(affinity), distillation MSE (+ the hybrid top-k term) and the balance loss on the affinity: one kernel
            # pair forward and one backward (competesmoe.py:350-371, moe.py:90-110)
            _, comp_losses = CompeteLossesFn.apply(gate_softmax, aff, aff_idx, None, B)
            if getattr(self.args, "hybrid", False):
                routerloss = comp_losses[0] + comp_losses[1] * self.args.router_theta
            else:
                routerloss = comp_losses[0]
            balance_loss = comp_losses[3]
            auxiliary_loss = routerloss * self.args.router_loss_coef + diversity_loss * self.args.diversity_loss_coef + \
                balance_loss * self.args.bal_comp_loss_coef
            self.last_routing = (aff_idx.view(B, N, K), aff_w.detach().view(B, N, K))
            infor_aux = {"balance_loss": balance_loss.detach().clone(), "diversity_loss": diversity_loss.detach().clone(),
                         "routerloss": routerloss.detach().clone()}
        else:
            out = self._sparse_ffn(x2, gate_w, gate_idx, w1, b1, w2, b2, spec)
            self.last_routing = (gate_idx.view(B, N, K), gate_w.detach().view(B, N, K))
            if want_aux:
                # balance + z losses come out of one fused reduction kernel (moe.py:214-226 combine_loss)
                balance_loss, router_z_loss = gate_losses[0], gate_losses[1]
                auxiliary_loss = balance_loss * self.args.balance_loss_coef + router_z_loss * self.args.router_z_loss_coef
                infor_aux = {"balance_loss": balance_loss.detach().clone(), "router_z_loss": router_z_loss.detach().clone()}
        return out.view(B, N, self.out_embed_dim).to(x.dtype), auxiliary_loss, None, infor_aux
