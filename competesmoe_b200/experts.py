"""Expert-module introspection and zero-copy stacking of per-expert parameters.

The reference keeps experts as real nn.Module children (`experts.{e}.fc1.weight`, ... -- SURVEY.md 8b: checkpoint
layout, sparse upcycling and optimizer grouping all depend on those names), while the grouped GEMM wants one
[E, n, k] tensor.  `fuse_storage` re-points every expert's parameter at a slice of one flat buffer (names, shapes and
state_dict unchanged); `StackParamsFn` then exposes the [E, ...] view without copying and hands each expert its slice
of the gradient.  Parameters that cannot be re-pointed (DeepSpeed ZeRO-3 placeholders) are stacked with a copy instead.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
from torch.autograd import Function

from . import ops


@dataclass
class ExpertLayout:
    kind: str                    # "mlp" | "glu"
    act: int                     # ops.ACT_*
    first: str                   # attribute path of the first Linear inside an expert ("fc1", "0", "gate_up_proj")
    second: str                  # ... of the second Linear ("fc2", "2", "down_proj")
    has_bias1: bool
    has_bias2: bool


def _act_code(m) -> int:
    name = type(m).__name__
    if isinstance(m, nn.GELU):
        return ops.ACT_GELU_TANH if m.approximate == "tanh" else ops.ACT_GELU
    if isinstance(m, nn.ReLU) or name == "ReLU":
        return ops.ACT_RELU
    if isinstance(m, nn.SiLU) or name in ("SiLUActivation", "SiLU"):
        return ops.ACT_SILU
    if name in ("GELUTanh", "PytorchGELUTanh", "NewGELUActivation", "FastGELUActivation"):
        return ops.ACT_GELU_TANH
    if name == "GELUActivation":
        return ops.ACT_GELU
    if isinstance(m, nn.Identity):
        return ops.ACT_NONE
    raise NotImplementedError(f"libcsmoe has no kernel for activation {name}; supported: GELU(erf/tanh), ReLU, SiLU")


def describe_expert(m: nn.Module) -> ExpertLayout:
    """Recognise the expert architectures the reference instantiates (siglip_smoe.py:85-97 SiglipMLP / CLIPMLP,
    multimodal_projector/builder.py:56-66 Sequential(Linear, GELU, Linear), Phi3MLP-style gate_up/down GLU)."""
    if isinstance(getattr(m, "fc1", None), nn.Linear) and isinstance(getattr(m, "fc2", None), nn.Linear):
        act = _act_code(getattr(m, "activation_fn", None) or getattr(m, "act", None) or nn.GELU())
        return ExpertLayout("mlp", act, "fc1", "fc2", m.fc1.bias is not None, m.fc2.bias is not None)
    if isinstance(m, nn.Sequential) and len(m) == 3 and isinstance(m[0], nn.Linear) and isinstance(m[2], nn.Linear):
        return ExpertLayout("mlp", _act_code(m[1]), "0", "2", m[0].bias is not None, m[2].bias is not None)
    if isinstance(getattr(m, "gate_up_proj", None), nn.Linear) and isinstance(getattr(m, "down_proj", None), nn.Linear):
        act = _act_code(getattr(m, "activation_fn", None) or nn.SiLU())
        if act != ops.ACT_SILU:
            raise NotImplementedError("gate_up/down experts are supported with SiLU gating only")
        if m.gate_up_proj.bias is not None or m.down_proj.bias is not None:
            raise NotImplementedError("gate_up/down experts with bias are not supported")
        return ExpertLayout("glu", ops.ACT_SILU_GLU, "gate_up_proj", "down_proj", False, False)
    raise NotImplementedError(
        f"expert module {type(m).__name__} is not a recognised 2-layer MLP (fc1/fc2, Sequential(Linear, act, Linear) "
        f"or gate_up_proj/down_proj); the B200 layer has no generic eager fallback")


def _is_stacked_view(ps: Sequence[torch.Tensor]) -> bool:
    p0 = ps[0]
    if not p0.is_contiguous():
        return False
    step = p0.numel() * p0.element_size()
    base = p0.data_ptr()
    for i, p in enumerate(ps):
        if p.dtype != p0.dtype or p.shape != p0.shape or not p.is_contiguous() or p.data_ptr() != base + i * step:
            return False
    try:
        s0 = p0.untyped_storage().data_ptr()
        return all(p.untyped_storage().data_ptr() == s0 for p in ps)
    except Exception:
        return False


def fuse_storage(params: Sequence[nn.Parameter]) -> bool:
    """Make params[e].data a slice of one contiguous [E, ...] buffer.  Returns False when that is not allowed."""
    if _is_stacked_view(params):
        return True
    if any(hasattr(p, "ds_id") or p.device.type == "meta" for p in params):
        return False
    with torch.no_grad():
        flat = torch.stack([p.data for p in params])
        for i, p in enumerate(params):
            p.data = flat[i]
    return True


class StackParamsFn(Function):
    """[p_0, ..., p_{E-1}] -> [E, *p.shape]; zero-copy when the parameters already share one buffer."""

    @staticmethod
    def forward(ctx, *ps):
        if _is_stacked_view(ps):
            p0 = ps[0].detach()
            return torch.as_strided(p0, (len(ps), *p0.shape), (p0.numel(), *p0.stride()), p0.storage_offset())
        return torch.stack([p.detach() for p in ps])

    @staticmethod
    def backward(ctx, g):
        return tuple(g[i] for i in range(g.shape[0]))


def stack_params(ps: List[Optional[nn.Parameter]]) -> Optional[torch.Tensor]:
    if ps[0] is None:
        return None
    return StackParamsFn.apply(*ps)
