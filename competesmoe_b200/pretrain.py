"""Drop-in for the language-pretraining MoE plugin (reference: moe_pretrain_model/layers/moe/{moe.py,competesmoe.py,
register.py} and the three framework mixins the layer inherits, framework/layers/{regularized,logging,once_per_iter}_layer.py).

Same constructor keywords, `forward(x, *, id_layer)`, schedule hooks, regulariser names (`mlp_ebalance`,
`mlp_router_loss`, `mlp_comp_diver_loss`, `mlp_comp_ebalance`) and checkpoint layout (`w_gate [E,D]`, `keys [E,D,H]`,
`values [E,H,D]`, optional `bias [E,H]`, `o_bias [D]`).  The two CVMM calls + eager bmm of the reference's
`compute_moe_main` are replaced by one fused permute -> grouped GEMM -> activation -> grouped GEMM -> combine path
(functional.SparseFFNFn); parameters stay fp32 and are cast to bf16 per step, as under the reference's autocast.
"""
from __future__ import annotations

import collections
import dataclasses
import math
from typing import Any, Callable, Dict, Optional

import torch
import torch.nn.functional as F

from . import ops
from .functional import (CompeteLossesFn, CompeteTailFn, DenseFFNFn, EntropyBalanceFn, FFNSpec, GateFn, SigmaFFNFn,
                         SparseFFNFn, sigma_fused_ok)
from .graphs import capture_guard
from .multimodal import TopkRenormFn
from .schedule import make_layer_schedule

MOE_REGISTRY: Dict[str, type] = {}


def register_moe(*names):
    def decorate(cls):
        for name in names:
            if name in MOE_REGISTRY and MOE_REGISTRY[name] != cls:
                raise AssertionError(f"Model named '{name}' conflicts with existing model!")
            MOE_REGISTRY[name] = cls
        return cls
    return decorate


def get_moe(model_name):
    try:
        return MOE_REGISTRY[model_name]
    except KeyError:
        raise ValueError(f"Attempted to load moe method'{model_name}', but no model for this name found! "
                         f"Supported model names: {', '.join(MOE_REGISTRY.keys())}")


# ------------------------------------------------------------------------------------------------ mixins
class RegularizedLayer:
    """framework/layers/regularized_layer.py:9-62: named regularisers, averaged per name on read, then reset."""

    def __init__(self) -> None:
        self.reg_accumulated = {}
        self.reg_counts_n = {}
        self.regularization_present = False

    @property
    def reg_enabled(self) -> bool:
        return self.training and self.regularization_present

    def add_reg(self, loss_fn: Callable[[], torch.Tensor], name: str = "reg"):
        if self.reg_enabled:
            v = loss_fn()
            if name in self.reg_accumulated:
                self.reg_accumulated[name] = self.reg_accumulated[name] + v
                self.reg_counts_n[name] += 1
            else:
                self.reg_accumulated[name] = v
                self.reg_counts_n[name] = 1

    def get_reg_loss(self) -> Dict[str, torch.Tensor]:
        out = {n: self.reg_accumulated[n] / self.reg_counts_n[n] for n in self.reg_accumulated}
        self.reg_accumulated = {}
        self.reg_counts_n = {}
        return out


class LoggingLayer:
    """framework/layers/logging_layer.py:9-52 (running sums of logged scalars)."""

    def __init__(self) -> None:
        self._logs = {}
        self._log_counts = {}

    def log(self, name: str, value: Any, drop_old: bool = False):
        if torch.is_tensor(value):
            value = value.detach()
        if name not in self._logs or drop_old or not isinstance(value, (torch.Tensor, float, int)):
            self._logs[name], self._log_counts[name] = value, 1
        else:
            self._logs[name] = self._logs[name] + value
            self._log_counts[name] += 1

    def get_logs(self) -> Dict[str, Any]:
        res = {k: (v / self._log_counts[k] if isinstance(v, (torch.Tensor, float, int)) else v) for k, v in self._logs.items()}
        self._logs, self._log_counts = {}, {}
        return res


class OncePerIterLayer:
    """framework/layers/once_per_iter_layer.py:1-10."""

    def pre_train_forward(self):
        pass

    def post_train_forward(self):
        pass

    def before_loss(self):
        pass


def _activation_code(fn: Callable) -> int:
    """The reference passes the activation as a callable (tasks/transformer_lm_mixin.py:121-122); identify it by value."""
    probe = torch.tensor([-2.0, -0.5, 0.0, 0.75, 3.0])
    got = fn(probe.clone())
    for code, ref in ((ops.ACT_RELU, F.relu(probe)), (ops.ACT_GELU, F.gelu(probe)),
                      (ops.ACT_GELU_TANH, F.gelu(probe, approximate="tanh")), (ops.ACT_SILU, F.silu(probe)),
                      (ops.ACT_NONE, probe)):
        if torch.allclose(got, ref, atol=1e-6):
            return code
    raise NotImplementedError("expert activation is not one of relu / gelu / gelu-tanh / silu / identity")


Selection = collections.namedtuple("Selection", ["raw_sel", "sel_val", "raw_sel_index", "sel_index"])   # moe.py:33


# ------------------------------------------------------------------------------------------------ base layer
class MoE(LoggingLayer, RegularizedLayer, OncePerIterLayer, torch.nn.Module):
    """sigma-MoE layout MoE MLP (reference: layers/moe/moe.py:35-138,323-332,373-440)."""

    def __init__(self, dmodel: int, n_experts: int, expert_size: int, n_heads: int, std_gate: float = 1.0,
                 std_expert: float = 1.0, topk=2, dropout: float = 0, weight_scale: float = 1.0,
                 selection_mode: str = "sigmoid", perplexity_reg: float = 0.0, perplexity_reg_mode: str = "step",
                 activation_after_topk: bool = False, activation=F.relu, sel_bias: bool = False, bias: bool = False,
                 v_dim: Optional[int] = None, expert_dropout: float = 0.0, sync_distributed: bool = False,
                 selection_dropout: float = 0.0, log_interval: Optional[int] = 100, args=None, is_att=False,
                 out_dmodel=None, inp_expert=None, out_expert=None):
        # explicit initialisation (no cooperative super().__init__): under integrate.bind_pretrain the reference's
        # MoE follows in the MRO and its __init__ must not run
        torch.nn.Module.__init__(self)
        LoggingLayer.__init__(self)
        RegularizedLayer.__init__(self)
        self.is_att = bool(is_att)
        self.iter = 0
        self.k_dim = self.k_vec_dim = dmodel
        self.v_dim = v_dim if v_dim is not None else dmodel
        self.n_experts = self.num_experts = self.num_of_experts = n_experts
        self.expert_size = expert_size
        self.size = n_experts * expert_size
        self.n_heads = n_heads
        self.num_selected = n_heads                      # moe.py:128: top-k is pkm.n_heads, the `topk` argument is ignored
        self.dropout, self.expert_dropout, self.selection_dropout = dropout, expert_dropout, selection_dropout
        self.selection_mode, self.perplexity_reg, self.perplexity_reg_mode = selection_mode, perplexity_reg, perplexity_reg_mode
        self.activation_after_topk = activation_after_topk
        self.activation = activation
        self._act_code = _activation_code(activation)
        self.weight_scale = self.sel_weight_scale = weight_scale
        self.layer = 0
        self.was_training = True
        self.sync_distributed = sync_distributed and torch.distributed.is_initialized()
        self.log_interval = log_interval
        self.out_dmodel = out_dmodel if out_dmodel is not None else dmodel
        self.div = 1
        self.real_n_experts = 1
        self.name_moe = "mlp"
        self.args = args
        self.training = False
        if self.is_att:
            # expert projections of SwitchHead-style MoE attention (moe.py:111-117; created by
            # full_moe_relative_attention.py:267-296 with n_experts = experts x heads): a gate over all (head, expert)
            # pairs and one [inp_expert, out_expert] matrix per pair; top-k is the `topk` argument here (moe.py:105)
            self.num_selected = topk
            self.w_gate = torch.nn.Parameter(torch.randn(n_experts, dmodel) * std_gate)
            self.renorm_rows(self.w_gate)
            self.div = 10
            self.real_n_experts = n_heads
            self.register_parameter("experts", torch.nn.Parameter(torch.randn(n_experts, inp_expert, out_expert) * std_expert))
            self.keys = self.values = None
        else:
            self.w_gate = torch.nn.Parameter(torch.empty(n_experts, dmodel))
            torch.nn.init.normal_(self.w_gate, std=dmodel ** -0.5 * weight_scale)
            self.register_parameter("values", torch.nn.Parameter(torch.empty(n_experts, expert_size, self.v_dim)))
            self.register_parameter("keys", torch.nn.Parameter(torch.empty(n_experts, dmodel, expert_size)))
            torch.nn.init.normal_(self.keys, std=dmodel ** -0.5 * weight_scale)
            torch.nn.init.normal_(self.values, std=self.size ** -0.5 * weight_scale)
        if bias:
            self.bias = torch.nn.Parameter(torch.zeros(n_experts, expert_size))
            self.o_bias = torch.nn.Parameter(torch.zeros(self.v_dim))
        else:
            self.bias = None
            self.o_bias = None
        self.dist_experts = None
        self.entropy_expert_selected, self.entropy_expert_all = [], []
        self.last_routing = None
        self.pre_train_forward()

    gate = property(lambda self: (lambda x: F.linear(x, self.w_gate, None)))

    def renorm_rows(self, x: torch.Tensor):
        """moe.py:140-144."""
        with torch.no_grad():
            std_t = x.std(dim=-1, keepdim=True)
            x.div_(x.norm(dim=-1, keepdim=True))
            x.mul_(std_t / x.std())

    # ---- MoE-attention projections (is_att): layers/transformer/full_moe_relative_attention.py:351-389,453-458
    def att_forward(self, x, n_experts, n_copies, return_full=True, *args, **kwargs):
        """Per-head expert selection for one attention projection (moe.py:456-486, the sigmoid-gated SwitchHead form the
        attention module calls at full_moe_relative_attention.py:374): gate logits [.., heads, experts], top-k per head,
        sigmoid of the selected logits as weights, and the CVMM selection over the shifted (head, expert) indices."""
        from .cvmm import cvmm_prepare_sel2
        assert self.is_att, "att_forward needs a layer built with is_att=True"
        if self.selection_dropout > 0 and self.training:
            x = F.dropout(x, self.selection_dropout)
        sel = F.linear(x, self.w_gate.to(x.dtype) if not torch.is_autocast_enabled() else self.w_gate, None)
        sel = sel.view(*sel.shape[:-1], n_copies, -1)
        with torch.no_grad():
            if self.expert_dropout > 0 and self.training:
                sel2 = sel.masked_fill(torch.rand_like(sel) < self.expert_dropout, float("-inf"))
            else:
                sel2 = sel
            _, sel_index = sel2.topk(self.num_selected, dim=-1, sorted=False)
        sel_val = torch.gather(sel, -1, sel_index).sigmoid()
        if self.training is False:
            self.add_dist_experts(selection=sel_index)
        shift = (torch.arange(n_copies, device=sel_index.device, dtype=sel_index.dtype) * n_experts).unsqueeze(-1)
        sel_pp = cvmm_prepare_sel2((shift + sel_index).flatten(-2, -1).int(), sel_val, n_experts=self.n_experts)
        return Selection(sel, sel_val, sel_index, sel_pp)

    def compute_moe(self, x: torch.Tensor, sel: "Selection") -> torch.Tensor:
        """moe.py:488-489: the projection itself, one grouped GEMM over the (head, expert) pairs."""
        from .cvmm import cvmm
        return cvmm(x, sel.sel_index, self.experts)

    # ---- expert parallelism (no counterpart in the reference, which is data-parallel only)
    _ep = None

    def enable_expert_parallel(self, group, max_tokens: int, row_tile: int = 128, exchange: str = "auto"):
        """Shard keys / values (/ bias) over `group` (competesmoe_b200.ep.EPGroup): this rank keeps the slices
        [rank*E/P, (rank+1)*E/P) along dim 0; w_gate (and o_bias) stay replicated.  `max_tokens` = the largest number of
        tokens this rank passes to forward.  Call after loading a full checkpoint and BEFORE the optimizer is built: the
        sharded tensors are new Parameters, an optimizer created earlier would keep updating the old full-size ones.
        Afterwards `state_dict()` holds this rank's shard only; use `full_state_dict()` / `load_full_state_dict()` for
        checkpoints in the reference layout.

        `exchange`: what crosses NVLink in a layer step.  "tokens": every (token, expert) row travels to the expert's
        owner and back (ep.EPSparseFFNFn).  "weights": the owners publish bf16 copies of their experts, every rank
        computes on its own tokens and the weight gradients are reduced onto the owners (ep.WeightExchange) -- fewer
        bytes whenever the experts are small next to the K-fold expanded batch, which is the case for every sigma-MoE
        configuration of the reference's sweeps.  "auto" compares the two byte counts."""
        from .ep import EPLayerState, WeightExchange
        if exchange not in ("auto", "tokens", "weights"):
            raise ValueError(f"exchange must be 'auto', 'tokens' or 'weights', got {exchange!r}")
        if exchange == "auto":
            per_expert = self.k_vec_dim * self.expert_size + self.expert_size * self.v_dim
            exchange = "weights" if WeightExchange.prefer_weights(self.n_experts, per_expert, max_tokens, self.num_selected,
                                                                  self.k_vec_dim, self.v_dim) else "tokens"
        self._shard_experts(group.rank, group.world)
        if exchange == "weights":
            self._ep = _WeightsEP(group, WeightExchange(group, {"w1": self.keys, "b1": self.bias, "w2": self.values}))
        else:
            self._ep = EPLayerState(group, self.n_experts, self.num_selected, self.k_vec_dim, self.v_dim, max_tokens, row_tile)
        return self

    @property
    def _wx(self):
        """The layer's ep.WeightExchange when its experts are sharded with exchange="weights", else None."""
        return getattr(self._ep, "wx", None)

    def _shard_experts(self, rank: int, world: int):
        E = self.n_experts
        if E % world != 0:
            raise ValueError(f"{E} experts cannot be split over an expert-parallel group of {world} ranks")
        if any(getattr(self, n) is not None and getattr(self, n).grad is not None for n in ("keys", "values", "bias")):
            raise RuntimeError("enable_expert_parallel() must run before training starts (and before the optimizer is "
                               "created): the expert parameters are replaced by their local shards")
        El = E // world
        lo = rank * El
        for name in ("keys", "values", "bias"):
            p = getattr(self, name)
            if p is not None:
                setattr(self, name, torch.nn.Parameter(p.detach()[lo:lo + El].clone(), requires_grad=p.requires_grad))
        self.ep_expert_offset = lo

    def full_state_dict(self, group=None):
        """Collective: the reference-layout state dict with every expert (ep.full_state_dict)."""
        from .ep import full_state_dict
        return full_state_dict(self, group)

    def load_full_state_dict(self, state_dict, group=None, strict: bool = True):
        from .ep import load_full_state_dict
        return load_full_state_dict(self, state_dict, group, strict)

    def _all_expert_weights(self):
        if self._ep is None or self._ep.group.world == 1 or self._wx is not None:
            return self.keys, self.bias, self.values
        from .ep import gather_experts
        g = self._ep.group
        return gather_experts(self.keys, g), gather_experts(self.bias, g), gather_experts(self.values, g)

    # ---- block tail (pretrain_block.FusedPreLNMoEBlock): (residual, dropout p, seed) for the combine epilogue
    _tail = None
    _tail_done = False
    _x_dtype = None     # dtype of the block's LayerNorm output when the fused pre-LN already cast the input

    # ---- CUDA graphs (no counterpart in the reference, whose cvmm path re-tunes and syncs on the host)
    _graphs = None
    _graphable = True
    _inplace_params = ()         # parameters a forward call rescales in place (xmoe / smoe_perturbed: expert_embeddings)

    def enable_cuda_graphs(self, enabled: bool = True):
        """Opt in: training-mode calls with a CUDA input that requires grad are replayed from captured CUDA graphs (one
        forward + one backward graph per (step kind, shape, dtype, autocast state); torch.cuda.make_graphed_callables
        behind the unchanged `layer(x, id_layer=...)` call).  At the sigma-MoE shapes (H = 128) the ~80 launches of a step
        take longer to issue from Python than to run, so this is worth 1.5-4x there.  Regularisers still arrive through
        `add_reg` under their usual names.  Eval / no-grad / test_only calls, and the competition step under expert
        parallelism (NCCL weight gather), take the normal path."""
        if enabled and self._graphs is None:
            self._graphs = {}
            self._eager_forward = self.forward
            self.forward = self._graph_forward
        elif not enabled and self._graphs is not None:
            self._graphs = None
            self.forward = self._eager_forward
        return self

    def __getstate__(self):
        """copy.deepcopy / pickling: captured CUDA graphs belong to this instance's storage and are not copied -- the
        copy keeps graph mode switched on and captures its own graphs on first use."""
        state = self.__dict__.copy()
        if state.get("_graphs") is not None:
            state["_graphs"] = {}
        return state

    def _graph_forward(self, x, *args, **kwargs):
        eligible = (self._graphable and self.training and torch.is_tensor(x) and x.is_cuda
                    and x.requires_grad and torch.is_grad_enabled() and not getattr(self.args, "test_only", False)
                    and not torch.cuda.is_current_stream_capturing() and not args
                    and set(kwargs) <= {"id_layer"})
        if not eligible:
            return self._eager_forward(x, *args, **kwargs)
        autocast = torch.is_autocast_enabled()
        adt = torch.get_autocast_dtype("cuda")
        id_layer = kwargs.get("id_layer")
        probe = getattr(self, "_is_competition_step", None)
        branch = bool(probe(x, id_layer)) if probe is not None else False
        # expert parallelism: the router step is kernels + device-side barriers over peer memory, which capture like any
        # other launch (every rank replays the same sequence); the competition step gathers the expert weights with NCCL
        # and stays eager
        if self._ep is not None and branch and self._wx is None:
            return self._eager_forward(x, *args, **kwargs)
        params = tuple(p for p in self.parameters() if p.requires_grad)
        reg_on = self.reg_enabled
        key = (branch, id_layer, tuple(x.shape), x.dtype, autocast, adt, reg_on,
               tuple(p.data_ptr() for p in params))
        entry = self._graphs.get(key)
        if entry is None:
            names = []
            keep = (self.layer, getattr(self, "nb_diver", 0))
            # capture runs the forward several times: put back what it rescales so that the first replay is call no. 1
            keep_p = {n: getattr(self, n).detach().clone() for n in self._inplace_params}

            def fn(xx, *_params):
                collected = []
                self.add_reg = (lambda loss_fn, name="reg": collected.append((name, loss_fn()))) if reg_on else \
                    (lambda loss_fn, name="reg": None)
                try:
                    with torch.autocast("cuda", dtype=adt, enabled=autocast, cache_enabled=False):
                        out = self._eager_forward(xx, **kwargs)
                finally:
                    del self.add_reg
                names[:] = [n for n, _ in collected]
                return (out,) + tuple(t for _, t in collected)

            sample = x.detach().clone().requires_grad_(True)
            with capture_guard(), torch.autocast("cuda", dtype=adt, enabled=autocast, cache_enabled=False):
                graphed = torch.cuda.make_graphed_callables(fn, (sample,) + params, allow_unused_input=True)
            self.layer, self.nb_diver = keep
            with torch.no_grad():
                for n, v in keep_p.items():
                    getattr(self, n).copy_(v)
            entry = (graphed, list(names), self.last_routing)
            self._graphs[key] = entry
        graphed, names, routing = entry
        res = graphed(x, *params)
        for name, t in zip(names, res[1:]):
            self.add_reg(lambda t=t: t, name)
        self.last_routing = routing          # static tensors of this graph, refreshed by the replay
        self.layer += 1
        self.was_training = self.training
        return res[0]

    # ---- bookkeeping hooks
    def pre_train_forward(self):
        self.total_selections, self.total_gate_softmax, self.total_gate_logits = [], [], []

    def before_loss(self):
        self.pre_train_forward()
        if self.training:
            self.iter += 1

    # ---- expert-usage statistics of `args.test_only` runs (moe.py:145-183; read by the evaluation harness)
    def entropy(self, prob_dist):
        return -torch.sum(prob_dist * torch.log(prob_dist + 1e-18), dim=-1)

    def add_dist_experts(self, selection=None):
        assert selection is not None, "Selection must to not None"
        n = self.num_of_experts // self.real_n_experts          # per-head expert count for MoE-attention projections
        sel = selection.reshape(-1, selection.shape[-1]).long()
        hist = F.one_hot(sel, num_classes=n).reshape(-1, n).sum(-2)
        self.dist_experts = hist if self.dist_experts is None else self.dist_experts + hist

    def get_dist_experts(self):
        return self.dist_experts

    def add_dist_weight(self, weight, is_all=False):
        ent = self.entropy(weight).mean()
        (self.entropy_expert_all if is_all else self.entropy_expert_selected).append(ent)

    def get_weight_dist(self):
        return {"entropy_all": torch.stack(self.entropy_expert_all).mean().item(),
                "entropy_topk": torch.stack(self.entropy_expert_selected).mean().item()}

    # ---- losses
    def entropy_balance(self, sel: torch.Tensor) -> torch.Tensor:
        """moe.py:323-332: minus the entropy of the sequence-averaged routing distribution, averaged over the batch."""
        s = sel.flatten(1, -2)
        ls = F.log_softmax(s.float(), dim=-1)
        lm = ls.logsumexp(-2) - math.log(ls.shape[-2])
        return (lm * lm.exp()).sum(-1).mean()

    # ---- compute
    def _compute_dtype(self, x: torch.Tensor) -> torch.dtype:
        if torch.is_autocast_enabled():
            return torch.get_autocast_dtype('cuda')
        return x.dtype

    _cast_memo = None

    def _cast(self, x2: torch.Tensor, cdt: torch.dtype) -> torch.Tensor:
        """x in the compute dtype, cast once per forward: the gate and the expert path (and the dense competition pass)
        share one copy and one autograd node instead of a cast each (autocast would cast per consumer too)."""
        if x2.dtype == cdt:
            return x2
        m = self._cast_memo
        if m is not None and m[0] is x2 and m[1] == cdt:
            return m[2]
        y = x2.to(cdt)
        self._cast_memo = (x2, cdt, y)
        return y

    def _spec(self, cdt: torch.dtype) -> FFNSpec:
        # CVMM.forward reduces with `reduction_weight.type_as(res) @ res` (cvmm.py:481-483): weight rounded to the op
        # dtype, fp32 accumulation, one rounding at the end.
        return FFNSpec(act=self._act_code, kn_layout=True, round_each=False, round_w=cdt == torch.bfloat16,
                       bias_after_round=True, fp32=cdt == torch.float32)

    def compute_gate(self, x2: torch.Tensor, cdt: Optional[torch.dtype] = None, x_dtype: Optional[torch.dtype] = None):
        """x_dtype: the layer input's dtype, to which the reference rounds the sum of the top-k weights (`.to(x.dtype)`).
        Called as the reference calls it -- compute_gate(x), moe.py:395-396 -- it returns the gate logits [..., E]."""
        if cdt is None:
            return self._reference_style_logits(x2)
        return GateFn.apply(self._cast(x2, cdt), self.w_gate, self.num_selected, 1, False, x_dtype)[:4]

    # ---- the reference's policy-level methods under their own names and signatures (moe.py:273-322,373-416;
    # competesmoe.py:381-455,465-490,510-522), on the kernels.  The layers' forward uses the fused forms of the same steps;
    # these are for callers and subclasses that go through the reference's method names.
    def _reference_style_logits(self, x):
        cdt = self._compute_dtype(x)
        return self.compute_gate(x.reshape(-1, x.shape[-1]), cdt, self._x_dtype or x.dtype)[0].view(*x.shape[:-1], -1)

    def topk_expert(self, gate_logits):
        """moe.py:373-393: (top-k softmax probabilities, not renormalised; their indices; the full softmax).  Ties: lowest
        expert index first."""
        gate_softmax = F.softmax(gate_logits, dim=-1, dtype=torch.float32)
        _, idx = ops.topk_renorm(gate_softmax.detach().reshape(-1, gate_softmax.shape[-1]).contiguous(), self.num_selected)
        selected_experts = idx.long().view(*gate_softmax.shape[:-1], self.num_selected)
        return torch.gather(gate_softmax, -1, selected_experts), selected_experts, gate_softmax

    def topk_expert_softmax(self, gate_logits):
        """competesmoe.py:415-434: top-k of the logits, softmax over the kept ones."""
        gate_softmax = F.softmax(gate_logits, dim=-1, dtype=torch.float32)
        _, idx = ops.topk_renorm(gate_logits.detach().float().reshape(-1, gate_logits.shape[-1]).contiguous(), self.num_selected)
        selected_experts = idx.long().view(*gate_logits.shape[:-1], self.num_selected)
        return F.softmax(torch.gather(gate_logits, -1, selected_experts), dim=-1, dtype=torch.float), selected_experts, gate_softmax

    def update_aux_statistics(self, gate_logits, gate_softmax, selected_experts):
        """moe.py:259-267: kept for callers; the three lists are emptied unread by `before_loss` (moe.py:341-358)."""
        self.total_selections.append(selected_experts)
        self.total_gate_logits.append(gate_logits)
        self.total_gate_softmax.append(gate_softmax)

    def zloss(self, gate_logits, gate_softmax=None):
        """moe.py:273-290."""
        return torch.square(torch.logsumexp(gate_logits, dim=-1)).mean()

    def balanceloss(self, selected_experts, gate_softmax):
        """moe.py:292-321 (top-1 density x mean probability; per head for MoE-attention projections)."""
        if self.is_att:
            k = gate_softmax.shape[-1]
            proxy = gate_softmax.mean(dim=1)                                                  # b n h k -> b h k
            density = F.one_hot(selected_experts, num_classes=k).float()[:, :, :, 0, :].mean(dim=1)
            return (proxy * density).mean() * float(k ** 2)
        proxy = gate_softmax.mean(dim=-2)
        density = F.one_hot(selected_experts[..., 0], self.num_of_experts // self.real_n_experts).float().mean(dim=-2)
        return (proxy * density).mean() * float(self.num_of_experts ** 2)

    def add_perplexity_reg(self):
        """moe.py:341-358: everything but the reset of the per-iteration lists is commented out in the reference."""
        self.pre_train_forward()

    def competition_policy_mlp_faster(self, x):
        """competesmoe.py:381-414, the reference's signature: every expert on every token (without the hidden bias),
        affinity = mean softplus(output) in fp32, top-k renormalised in x's dtype.  Returns (weights [B, N, K], selected
        experts [B, N, K] int64, softmax(affinity) [B, N, E], affinity [B, N, E], selected outputs [B, N, K, D_v])."""
        from .functional import AffinityFn, GatherRowsFn
        lead = x.shape[:-1]
        cdt = self._compute_dtype(x)
        xdt = self._x_dtype or x.dtype
        x2 = x.reshape(-1, x.shape[-1])
        T, E, K = x2.shape[0], self.n_experts, self.num_selected
        self._wx_prefetch(cdt)
        self._wx_open = False
        keys, _, values = self._all_expert_weights()
        y_all = DenseFFNFn.apply(self._cast(x2, cdt), keys, None, values, None, self._spec(cdt), None, self._wx)
        t_pad = y_all.shape[0] // E
        aff = AffinityFn.apply(y_all, E, T, t_pad, xdt == torch.bfloat16)
        w, idx = TopkRenormFn.apply(aff, K, False, xdt)
        topk_out = GatherRowsFn.apply(y_all, idx, t_pad)
        aff3 = aff.view(*lead, E)
        return (w.view(*lead, K), idx.long().view(*lead, K), F.softmax(aff3, dim=-1, dtype=torch.float32), aff3,
                topk_out.view(*lead, K, -1))

    def compute_scores(self, input: torch.Tensor, index) -> torch.Tensor:
        """moe.py:397-416: activation(cvmm(input, index, keys) + bias[index.raw_sel]) through the public op."""
        from .cvmm import cvmm
        scores = cvmm(input, index, self.keys)
        if self.bias is not None:
            scores = scores + self.bias[index.raw_sel.long()]
        scores = self.activation(scores)
        if self._plot_training():
            with torch.no_grad():
                self.log("relu_pass_rate", (scores > 0).float().sum() / scores.numel())
        return scores

    def _plot_training(self) -> bool:
        """moe.py:405: `self.train and log_interval is not None and iter % log_interval == 0` (`self.train` is the bound
        method there, i.e. always true).  Not logged from inside a CUDA-graph capture (the flag changes per iteration)."""
        return (self.log_interval is not None and self.iter % self.log_interval == 0
                and not (torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()))

    def _log_relu_pass_rate(self, h, row_to_slot, n_slots: int):
        """moe.py:405-414: fraction of positive hidden activations.  h is the padded expert-major [row_cap, H] buffer:
        only rows that hold a routed slot count (padding rows inside a tile are zero, tiles past the end are never
        written), and the denominator is the number of routed activations."""
        with torch.no_grad():
            routed = (row_to_slot >= 0).unsqueeze(1)
            self.log("relu_pass_rate", ((h > 0) & routed).sum().float() / float(n_slots * h.shape[1]))

    def _wx_prefetch(self, cdt: torch.dtype):
        """Expert parallelism with exchanged weights: publish this rank's expert shards now, from a side stream, so
        that the transfer runs under the router kernels of this forward call (ep.WeightExchange.prefetch)."""
        wx = self._wx
        if wx is not None:
            wx.prefetch({"w1": self.keys, "b1": self.bias, "w2": self.values},
                        torch.bfloat16 if cdt == torch.bfloat16 else torch.float32)
            self._wx_open = True

    _wx_open = False

    def compute_moe_main(self, x2, selected, weights, cdt: Optional[torch.dtype] = None, same_step: bool = False):
        if cdt is None:
            # the reference's signature (competesmoe.py:510-522): x [..., D], selected_experts / weights [..., K]
            lead, K = x2.shape[:-1], selected.shape[-1]
            out = self.compute_moe_main(x2.reshape(-1, x2.shape[-1]), selected.reshape(-1, K).to(torch.int32).contiguous(),
                                        weights.reshape(-1, K).float(), self._compute_dtype(x2))
            return out.view(*lead, self.v_dim)
        wx = self._wx
        if wx is not None and not same_step and not self._wx_open:   # callers that did not prefetch (sibling routers)
            wx.begin_step()
        self._wx_open = False
        if self._ep is not None and wx is None:
            from .ep import EPSparseFFNFn
            return EPSparseFFNFn.apply(self._cast(x2, cdt), weights, selected, self.keys, self.bias, self.values, None,
                                       self._spec(cdt), self._ep)
        spec = self._spec(cdt)
        xc = self._cast(x2, cdt)
        fused = sigma_fused_ok(xc, self.keys, self.values, spec, cdt)      # expert size 128 + ReLU: csrc/sigma_ffn.cu
        tail = self._tail
        if tail is not None and fused and self.o_bias is None and not self._plot_training():
            # called from pretrain_block.FusedPreLNMoEBlock: `src + dropout(layer output)` leaves the combine epilogue
            residual, p, seed = tail
            self._tail_done = True
            return SigmaFFNFn.apply(xc, weights, selected, self.keys, self.bias, self.values, spec,
                                    residual.reshape(-1, residual.shape[-1]), p, seed, wx)
        if self._plot_training():
            spec = dataclasses.replace(spec, return_hidden=True)
            if fused:
                out, h, row_to_slot = SigmaFFNFn.apply(xc, weights, selected, self.keys, self.bias, self.values, spec,
                                                       None, 0.0, 0, wx)
            else:
                out, h, row_to_slot = SparseFFNFn.apply(xc, weights, selected, self.keys, self.bias, self.values, None, spec, wx)
            self._log_relu_pass_rate(h, row_to_slot, selected.numel())
            return out
        if fused:
            return SigmaFFNFn.apply(xc, weights, selected, self.keys, self.bias, self.values, spec, None, 0.0, 0, wx)
        return SparseFFNFn.apply(xc, weights, selected, self.keys, self.bias, self.values, None, spec, wx)

    def forward(self, x, return_id_experts=False, return_full=True, *args, **kwargs):
        """Plain sigma-MoE forward (moe.py:418-449)."""
        B = x.shape[:-1]
        cdt = self._compute_dtype(x)
        self._wx_prefetch(cdt)
        x2 = x.reshape(-1, x.shape[-1])
        logits, probs, _, gidx = self.compute_gate(x2, cdt)
        gw = torch.gather(probs, 1, gidx.long())     # moe.py:373-393 topk_expert: the raw top-k probabilities, no renormalisation
        if self.training is False:
            self.add_dist_experts(selection=gidx)
        out = self.compute_moe_main(x2, gidx, gw, cdt)
        self.layer += 1
        self.was_training = self.training
        res = out.view(*B, self.v_dim)
        if self.o_bias is not None:
            res = res + self.o_bias
        lg = logits.view(*B, -1)
        self.add_reg(lambda: self.entropy_balance(lg) * (self.args.balance_loss_coef / self.div), f"{self.name_moe}_ebalance")
        return res


class _WeightsEP:
    """`layer._ep` of a layer whose experts are sharded with exchange="weights": the group and the exchange, no token buffers."""

    def __init__(self, group, wx):
        self.group, self.wx = group, wx


# ------------------------------------------------------------------------------------------------ CompeteSMoE
@register_moe("competesmoe", "competesmoe_b200")
class CompeteSMoE(MoE):
    """reference: layers/moe/competesmoe.py:37-616."""

    def __init__(self, dmodel: int, n_experts: int, expert_size: int, n_heads: int, std_gate=1.0, std_expert=1.0, topk=2,
                 dropout: float = 0, weight_scale: float = 1.0, selection_mode: str = "sigmoid",
                 perplexity_reg: float = 0.0, perplexity_reg_mode: str = "step", activation_after_topk: bool = False,
                 activation=F.relu, sel_bias: bool = False, bias: bool = False, v_dim: Optional[int] = None,
                 expert_dropout: float = 0.0, sync_distributed: bool = False, selection_dropout: float = 0.0,
                 log_interval: Optional[int] = 100, args=None, std=1, out_dmodel=None, is_att=False, inp_expert=None,
                 out_expert=None):
        MoE.__init__(self, dmodel=dmodel, n_experts=n_experts, expert_size=expert_size, n_heads=n_heads, topk=topk,
                         dropout=dropout, weight_scale=weight_scale, selection_mode=selection_mode,
                         perplexity_reg=perplexity_reg, perplexity_reg_mode=perplexity_reg_mode,
                         activation_after_topk=activation_after_topk, activation=activation, sel_bias=sel_bias, bias=bias,
                         v_dim=v_dim, expert_dropout=expert_dropout, sync_distributed=sync_distributed,
                         selection_dropout=selection_dropout, log_interval=log_interval, args=args,
                         out_dmodel=out_dmodel, is_att=is_att, out_expert=out_expert, inp_expert=inp_expert,
                         std_gate=std_gate, std_expert=std_expert)
        self.warm_up = args.warm_up
        self.rate_flip = args.rate_flip
        self.current_steps = 0
        self.step_warm = None
        self.is_prob_flips = True
        assert args.stop_after > 0, f"Warning: stop_after {args.stop_after} < 1, You must setting stop_after > 0"
        self.total_steps = args.stop_after
        self.prob_flips_final = {}
        self.max_compete_in_iter = args.max_compete_in_iter
        self.nb_diver = 0
        self._flips_host = {}

    # ---- schedule (competesmoe.py:123-273, :328-329)
    def set_total_steps(self, id_layer=0):
        self.step_warm, flags = make_layer_schedule(self.total_steps, self.warm_up, self.rate_flip,
                                                    self.max_compete_in_iter, self.prob_flips_final)
        self.flip_steps = self.total_steps - self.step_warm
        self.prob_flips_final[id_layer] = flags
        self.is_prob_flips = False
        return self.prob_flips_final

    def set_current_steps(self, step):
        self.current_steps = step

    def _is_competition_step(self, x, id_layer) -> bool:
        """competesmoe.py:528 without the per-call device sync: flags are mirrored on the host per tensor version."""
        if not x.requires_grad or self.step_warm is None or self.current_steps < self.step_warm:
            return False
        pf = self.prob_flips_final[id_layer]
        key = (id(pf), pf._version, pf.data_ptr())
        cached = self._flips_host.get(id_layer)
        if cached is None or cached[0] != key:
            cached = (key, pf.detach().to("cpu").ne(0).tolist())
            self._flips_host[id_layer] = cached
        return bool(cached[1][self.current_steps - self.step_warm])

    def pre_train_forward(self):
        super().pre_train_forward()
        self.total_router_gate, self.total_router_affinity = [], []

    def add_perplexity_reg(self):
        self.pre_train_forward()

    def before_loss(self):
        self.add_perplexity_reg()
        if self.training:
            self.iter += 1

    # ---- policies
    def compute_gate(self, x2: torch.Tensor, cdt: Optional[torch.dtype] = None, x_dtype: Optional[torch.dtype] = None):
        """competesmoe.py:456-464: plain, cosine, or weight-normalised gate.  x_dtype: dtype of the layer input, which the
        reference rounds the top-k sum to (`.to(x.dtype)`, :489): fp32 for fp32 inputs under autocast, so no rounding.
        Called as compute_gate(x) it returns the gate logits [..., E] like the reference's."""
        if cdt is None:
            return self._reference_style_logits(x2)
        a = self.args
        if getattr(a, "is_cosine", False) and not getattr(a, "is_norm_weight", False):
            return GateFn.apply(F.normalize(x2.float(), p=2.0, dim=-1).to(cdt), F.normalize(self.w_gate, p=2.0, dim=-1),
                                self.num_selected, 1, False, x_dtype)[:4]
        if getattr(a, "is_norm_weight", False):
            return GateFn.apply(self._cast(x2, cdt), F.normalize(self.w_gate, p=2.0, dim=-1), self.num_selected, 1, False,
                                x_dtype)[:4]
        return GateFn.apply(self._cast(x2, cdt), self.w_gate, self.num_selected, 1, False, x_dtype)[:4]

    def router_policy(self, x2, cdt: Optional[torch.dtype] = None, x_dtype: Optional[torch.dtype] = None,
                      is_normal_mode: bool = False):
        """competesmoe.py:465-490.  Called as the reference calls it -- router_policy(x, is_normal_mode=...) with
        x [B, N, D] -- it returns (weights [B, N, K], selected experts [B, N, K] int64, gate softmax [B, N, E], gate logits)."""
        if cdt is None:
            lead, K = x2.shape[:-1], self.num_selected
            gw, gidx, probs, logits = self.router_policy(x2.reshape(-1, x2.shape[-1]), self._compute_dtype(x2),
                                                         self._x_dtype or x2.dtype)
            return gw.view(*lead, K), gidx.long().view(*lead, K), probs.view(*lead, -1), logits.view(*lead, -1)
        a = self.args
        assert not (getattr(a, "is_cosine", False) and getattr(a, "is_norm_weight", False)), \
            "Can not active  both  Cosine and Norm Weigh. Just use one method - Cosine or Norm Weigh to Normalization"
        logits, probs, gw, gidx = self.compute_gate(x2, cdt, x_dtype)
        if getattr(a, "norm_sigmoid", False):
            scale = float(getattr(a, "scale_weight", 1.0))
            gw, gidx = TopkRenormFn.apply(logits.float() / scale, self.num_selected, True, x_dtype)
        return gw, gidx, probs, logits

    def router_loss(self, gate_softmax, affinity_softmax):
        return F.mse_loss(gate_softmax, affinity_softmax)

    def experts_diversity_loss(self, expert_outputs):
        """competesmoe.py:330-372 on [T, K, D]."""
        eo = expert_outputs.float()
        K, D = eo.shape[-2:]
        nrm = F.normalize(eo, p=2, dim=-1).reshape(-1, K, D)
        sim = torch.bmm(nrm, nrm.transpose(1, 2)) * (1 - torch.eye(K, device=eo.device))
        self.nb_diver += K * (K - 1) * nrm.shape[0]
        return sim.mean()

    def forward(self, x, return_id_experts=False, return_full=True, *args, **kwargs):
        id_layer = kwargs["id_layer"]
        assert id_layer is not None, "Layer Id must to not None"
        a = self.args
        lead = x.shape[:-1]
        cdt = self._compute_dtype(x)
        self._wx_prefetch(cdt)
        x2 = x.reshape(-1, x.shape[-1])
        T, E, K = x2.shape[0], self.n_experts, self.num_selected
        is_comp = self._is_competition_step(x, id_layer)
        if T == 0 and self._ep is None:      # (under expert parallelism a rank without tokens still takes part in the exchange)
            # The reference cannot train on an empty batch: entropy_balance takes math.log(0) (moe.py:323-332 -> ValueError)
            # and the competition's `.view(B, N, -1)` is ambiguous (competesmoe.py:399 -> RuntimeError).  Without
            # regularisers its router branch returns an empty result, and so does this.
            if is_comp or self.reg_enabled:
                raise ValueError("CompeteSMoE.forward: no tokens in the batch (math domain error in the reference's "
                                 "entropy_balance / ambiguous view in its competition step)")
            res = x.new_zeros(*lead, self.v_dim) + x.sum() * 0
            return res + self.o_bias if self.o_bias is not None else res
        # dtype the reference layer would see: a fused pre-LN hands the input over already cast (pretrain_block.py), but
        # the `.to(x.dtype)` roundings of the reference refer to the LayerNorm's output dtype
        xdt = self._x_dtype or x.dtype
        gate_w, gate_idx, gate_softmax, gate_logits = self.router_policy(x2, cdt, xdt)
        if is_comp:
            spec = self._spec(cdt)
            self._wx_open = False
            keys, bias, values = self._all_expert_weights()
            # competition_policy_mlp_faster (:381-414) scores every expert WITHOUT the hidden bias; compute_moe_main then
            # recomputes the selected experts with it (moe.py:400-401).  Without a bias the two coincide and the selected
            # outputs are reused from the dense pass; with one the sparse path runs on the competition's selection.
            y_all, score_sums = DenseFFNFn.apply(self._cast(x2, cdt), keys, None, values, None, spec,
                                                 xdt == torch.bfloat16, self._wx)    # [E * t_pad, Dv]
            t_pad = y_all.shape[0] // E
            aff, aff_w, aff_idx, out, diver = CompeteTailFn.apply(y_all, E, T, t_pad, K, False, xdt, spec, score_sums)
            self.nb_diver += K * (K - 1) * T
            if self.bias is not None:
                out = self.compute_moe_main(x2, aff_idx, aff_w, cdt, same_step=True)
            # softmax(affinity), every router-loss variant and the entropy balance on the affinity from one kernel pair
            # (competesmoe.py:541-593): losses = (MSE, MSE at the competition's top-k, MSE at the router's top-k, -, ebalance)
            _, cl = CompeteLossesFn.apply(gate_softmax, aff, aff_idx, gate_idx if a.tribrid and not (a.in_topk or a.hybrid) else None,
                                          lead[0] if len(lead) > 1 else 1)
            self.add_reg(lambda: diver * a.balance_loss_coef_comp / 2, self.name_moe + "_comp_diver_loss")
            if a.balance_affinity:
                self.add_reg(lambda: cl[4] * a.balance_loss_coef_comp / 2, f"{self.name_moe}_comp_ebalance")
            if a.in_topk:
                rl = cl[1]
            elif a.hybrid:
                rl = cl[0] + cl[1] * a.router_theta
            elif a.tribrid:
                rl = cl[0] + cl[1] * a.router_theta + cl[2] * a.router_theta
            else:
                rl = cl[0]
            self.add_reg(lambda: rl * a.router_loss_coef, f"{self.name_moe}_router_loss")
            self.last_routing = (aff_idx.view(*lead, K), aff_w.detach().view(*lead, K))
        else:
            out = self.compute_moe_main(x2, gate_idx, gate_w, cdt)
            if self.reg_enabled:   # entropy_balance(gate_logits) (:603-605) from the router kernel's probabilities
                eb = EntropyBalanceFn.apply(gate_softmax, lead[0] if len(lead) > 1 else 1)
                self.add_reg(lambda: eb * (a.balance_loss_coef / self.div), f"{self.name_moe}_ebalance")
            self.last_routing = (gate_idx.view(*lead, K), gate_w.detach().view(*lead, K))
        self.layer += 1
        if a.test_only:
            self.add_dist_experts(selection=gate_idx)
            self.add_dist_weight(weight=gate_w)
            self.add_dist_weight(weight=gate_softmax, is_all=True)
        self.was_training = self.training
        res = out.view(*lead, self.v_dim)
        if self.o_bias is not None:
            res = res + self.o_bias
        return res
