"""The steps either side of the MoE layer in the reference's pre-LN transformer block, fused around the layer
(reference: moe_pretrain_model/layers/transformer/relative_moe_transformer.py:150-159):

    src2 = self.norm2(mlp_input)                 LayerNorm + the autocast cast of its output: one kernel (csmoe_layernorm_fwd)
    src3 = self.pkm(src2, id_layer=id_layer)
    src  = src + self.dropout(src3)              residual add + dropout in the combine kernel's epilogue
                                                 (csmoe_combine_residual_fwd) -- or one pass over the finished layer output
                                                 (csmoe_residual_dropout_fwd) where the layer cannot take the tail

`FusedPreLNMoEBlock(norm2, pkm, dropout)` holds the SAME `norm2` LayerNorm and `pkm` layer objects the reference block
holds (parameters, state-dict keys and the layer's regulariser side effects are untouched) and replaces the three lines
above; `patch_block(block)` rebinds them on a reference `RelativeMoeTransformerEncoderLayer`-shaped object.
"""
from __future__ import annotations

from typing import Optional

import torch

from .functional import LayerNormCastFn, ResidualDropoutFn

_MASK64 = (1 << 63) - 1


class FusedPreLNMoEBlock(torch.nn.Module):
    def __init__(self, norm: torch.nn.LayerNorm, pkm: torch.nn.Module, dropout: float):
        super().__init__()
        self.norm2 = norm
        self.pkm = pkm
        self.p = float(dropout.p if isinstance(dropout, torch.nn.Dropout) else dropout)
        self._calls = 0

    def _next_seed(self) -> int:
        """A fresh dropout stream per call, derived on the host from torch's seed (no device sync, reproducible under
        torch.manual_seed)."""
        self._calls += 1
        return (torch.initial_seed() * 0x9E3779B97F4A7C15 + self._calls * 0xD1B54A32D192ED03 + id(self) % 65536) & _MASK64

    def forward(self, src: torch.Tensor, id_layer: Optional[int] = None) -> torch.Tensor:
        n = self.norm2
        D = src.shape[-1]
        cdt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else src.dtype
        fusable = (src.is_cuda and isinstance(n, torch.nn.LayerNorm) and n.elementwise_affine and n.bias is not None
                   and D % 8 == 0 and D <= 2048 and src.dtype in (torch.float32, torch.bfloat16)
                   and (src.dtype, cdt) != (torch.bfloat16, torch.float32))
        src2 = LayerNormCastFn.apply(src, n.weight, n.bias, n.eps, cdt) if fusable else n(src)
        p = self.p if self.training else 0.0
        seed = self._next_seed() if p > 0.0 else 0
        pkm = self.pkm
        # the layer applies the tail itself (combine epilogue) when it runs the fused expert kernels; not from inside a
        # CUDA-graph replay, where the residual pointer and the seed would be frozen into the graph
        offer = fusable and getattr(pkm, "_graphs", None) is None and hasattr(pkm, "_tail")
        if offer:
            pkm._tail, pkm._tail_done = (src, p, seed), False
        if fusable and hasattr(pkm, "_x_dtype"):
            pkm._x_dtype = src.dtype          # what `norm2(src)` would have handed the layer
        try:
            out = pkm(src2, id_layer=id_layer)
            done = offer and pkm._tail_done
        finally:
            if offer:
                pkm._tail, pkm._tail_done = None, False
            if hasattr(pkm, "_x_dtype"):
                pkm._x_dtype = None
        if done:
            return out.view(*src.shape)
        if src.is_cuda and D % 8 == 0 and out.dtype in (torch.float32, torch.bfloat16) and \
                (out.dtype, src.dtype) != (torch.float32, torch.bfloat16):
            return ResidualDropoutFn.apply(out.contiguous(), src.contiguous(), p, seed)
        return src + torch.nn.functional.dropout(out, p, self.training)


def patch_block(block: torch.nn.Module) -> torch.nn.Module:
    """Give a reference-shaped pre-LN block (attributes `norm2`, `pkm`, `dropout`, `preln`) the fused MLP half: returns a
    FusedPreLNMoEBlock sharing the block's own submodules, to be called as `src = fused(src, id_layer=...)` in place of
    relative_moe_transformer.py:150-157."""
    if not getattr(block, "preln", True):
        raise NotImplementedError("post-LN blocks normalise after the residual: nothing to fuse before the layer")
    return FusedPreLNMoEBlock(block.norm2, block.pkm, block.dropout)
