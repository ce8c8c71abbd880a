"""ctypes binding of libcsmoe.so (include/csmoe.h).

There is no fallback: if the library is missing it is built with nvcc once (in-tree); if that fails, or a call
returns a non-zero status, a RuntimeError is raised.  Nothing in this package computes the hot path on the CPU.
"""
from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "lib" / "libcsmoe.so"
ABI_VERSION = 2      # include/csmoe.h CSMOE_ABI_VERSION: bumped whenever a signature or csmoe_gemm_args changes

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU, ACT_GELU_TANH, ACT_SILU, ACT_SILU_GLU = range(6)
GEMM_ROWS, GEMM_REDUCE = 0, 1
ROW_TILE = 128

vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
f32c, u64 = C.c_float, C.c_uint64


class GemmArgs(C.Structure):
    _fields_ = [
        ("mode", i32), ("b_layout", i32), ("num_experts", i32), ("dense", i32),
        ("m", i64), ("n", i64), ("k", i64), ("dense_rows", i64),
        ("a", vp), ("lda", i64), ("a_expert_rows", i64),
        ("b", vp), ("ldb", i64), ("b_expert_stride", i64),
        ("c", vp), ("ldc", i64), ("c_expert_stride", i64), ("c_dtype", i32), ("act", i32),
        ("bias", vp), ("bias_dtype", i32), ("accumulate", i32),
        ("preact", vp), ("ldpre", i64),
        ("tile_expert", vp), ("pad_offsets", vp),
        ("max_ctas", i32), ("act_bwd", i32),
        ("aux", vp), ("ldaux", i64),
        ("row_tile", i32), ("sum_experts", i32),
        ("c_rows", vp),
        ("rowsum", vp), ("rowsum_round", i32), ("bias_after_round", i32),
    ]


# name -> (restype, argtypes); mirrors include/csmoe.h one to one
_SIGNATURES = {
    "csmoe_abi_version": (i32, []),
    "csmoe_last_error": (C.c_char_p, []),
    "csmoe_device_supported": (i32, []),
    "csmoe_route_row_cap": (i64, [i64, i32, i32]),
    "csmoe_route_workspace_bytes": (i64, [i64, i32]),
    "csmoe_route_build": (i32, [vp, i64, i32, i32, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "csmoe_router_fwd": (i32, [vp, vp, i32, i64, i32, i32, i32, i32, vp, vp, vp, vp, vp]),
    "csmoe_router_aux_workspace_bytes": (i64, [i64, i64, i32]),
    "csmoe_router_aux_fwd": (i32, [vp, i32, vp, vp, i64, i64, i32, i32, vp, vp, vp, vp, vp, vp]),
    "csmoe_router_bwd_workspace_bytes": (i64, [i64, i32, i32]),
    "csmoe_router_bwd": (i32, [vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, i32, i32, i32, i32, vp, vp, vp, i32, vp, vp]),
    "csmoe_router_from_logits": (i32, [vp, i32, i64, i32, i32, i32, vp, vp, vp, vp]),
    "csmoe_topk_renorm": (i32, [vp, i64, i32, i32, i32, i32, vp, vp, vp]),
    "csmoe_gather_rows": (i32, [vp, i32, i64, i32, i32, vp, i64, vp, vp, vp]),
    "csmoe_combine_fwd": (i32, [vp, i32, i64, i32, i32, vp, vp, vp, i32, vp, vp]),
    "csmoe_combine_bwd_w": (i32, [vp, vp, i32, i64, i32, i32, vp, vp, vp]),
    "csmoe_scatter_reduce": (i32, [vp, i32, i64, i32, i32, vp, i32, vp, vp]),
    "csmoe_grouped_gemm": (i32, [C.POINTER(GemmArgs), vp]),
    "csmoe_act_fwd": (i32, [vp, i32, i64, i64, i64, i32, vp, i64, vp, vp]),
    "csmoe_act_bwd": (i32, [vp, vp, i32, i64, i64, i64, i64, i32, vp, vp, vp]),
    "csmoe_bias_grad_workspace_bytes": (i64, [i32, i32]),
    "csmoe_bias_grad": (i32, [vp, i32, i64, i32, i32, vp, i32, i64, vp, i32, vp, vp]),
    "csmoe_act_bwd_bias": (i32, [vp, vp, i32, i64, i64, i32, i32, vp, i32, i64, i32, vp, vp, i32, vp, vp]),
    "csmoe_cast_f32_bf16": (i32, [vp, vp, i64, vp]),
    "csmoe_split_f32_bf16x3": (i32, [vp, vp, vp, vp, i64, vp]),
    "csmoe_affinity_fwd": (i32, [vp, i32, i32, i64, i64, i32, i32, vp, vp]),
    "csmoe_affinity_from_rowsum": (i32, [vp, i32, i32, i64, i64, i32, i32, vp, vp]),
    "csmoe_affinity_bwd": (i32, [vp, vp, i32, i32, i64, i64, i32, i32, vp, vp]),
    "csmoe_diversity_fwd": (i32, [vp, i32, i64, i64, i32, i32, vp, vp, vp, vp, vp, vp]),
    "csmoe_compete_bwd": (i32, [vp, i32, i32, i64, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "csmoe_sigma_ffn_supported": (i32, [i64, i32, i64]),
    "csmoe_sigma_set_stats": (i32, [vp]),
    "csmoe_sigma_ffn_fwd": (i32, [vp, i64, i32, i32, i32, vp, vp, vp, i32, vp, vp, i64, i32, vp, vp, vp, vp]),
    "csmoe_sigma_ffn_bwd": (i32, [vp, i64, i32, i32, i32, vp, vp, vp, vp, i64, i32, vp, i64, vp, vp, vp, vp, vp, vp, vp]),
    "csmoe_sigma_wgrad": (i32, [vp, vp, i64, i32, i32, vp, vp, i64, i32, i32, vp, i32, vp]),
    "csmoe_losses_workspace_bytes": (i64, [i64, i64, i32]),
    "csmoe_losses_fwd": (i32, [vp, vp, vp, vp, i64, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp]),
    "csmoe_losses_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, i64, i64, i32, i32, vp, vp, vp]),
    "csmoe_entropy_balance_fwd": (i32, [vp, i64, i64, i32, vp, vp, vp, vp]),
    "csmoe_entropy_balance_bwd": (i32, [vp, vp, i64, i64, i32, vp, vp]),
    "csmoe_topk_renorm_bwd": (i32, [vp, vp, vp, vp, i64, i32, i32, i32, i32, vp, vp]),
    "csmoe_dense_rows": (i32, [vp, i64, i32, i64, vp, vp]),
    "csmoe_layernorm_fwd": (i32, [vp, i32, i64, i32, vp, vp, f32c, vp, i32, vp, vp, vp]),
    "csmoe_layernorm_bwd_workspace_bytes": (i64, [i64, i32]),
    "csmoe_layernorm_bwd": (i32, [vp, i32, vp, i32, vp, vp, vp, i64, i32, vp, vp, vp, vp, vp]),
    "csmoe_combine_residual_fwd": (i32, [vp, i32, i64, i32, i32, vp, vp, vp, i32, vp, i32, f32c, u64, vp, vp]),
    "csmoe_residual_dropout_fwd": (i32, [vp, i32, vp, i32, i64, i32, f32c, u64, vp, vp]),
    "csmoe_dropout_bwd": (i32, [vp, i32, i64, f32c, u64, vp, i32, vp]),
    "csmoe_ep_ipc_handle_bytes": (i32, []),
    "csmoe_ep_alloc": (i32, [i64, C.POINTER(vp), vp]),
    "csmoe_ep_open": (i32, [vp, C.POINTER(vp)]),
    "csmoe_ep_close": (i32, [vp]),
    "csmoe_ep_free": (i32, [vp]),
    "csmoe_ep_barrier": (i32, [vp, vp, i32, i32, vp]),
    "csmoe_ep_exchange_plan": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i64, vp, vp, vp, vp, vp]),
    "csmoe_ep_dispatch": (i32, [vp, i32, i32, i32, i64, vp, vp, vp, vp, i32, vp, vp, vp, i32, i32, vp]),
    "csmoe_ep_row_ptrs": (i32, [vp, vp, vp, vp, i32, i64, vp, i64, i32, i32, vp, vp, i32, vp]),
    "csmoe_ep_push_rows": (i32, [vp, i32, i64, i32, i64, vp, vp]),
    "csmoe_ep_gather_push": (i32, [vp, i32, i64, vp, i32, i64, i32, i32, vp]),
    "csmoe_ep_reduce_pull": (i32, [vp, i64, i64, vp, i32, i32, vp]),
}

_lock = threading.Lock()
_lib = None


def exported_symbols() -> list[str]:
    """Every symbol include/csmoe.h declares (used by the CPU-side load test)."""
    return sorted(_SIGNATURES)


def load() -> C.CDLL:
    """Load (building first if needed) libcsmoe.so.  Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        # build() is a digest compare of csrc/ + include/ against lib/libcsmoe.stamp when the library exists, and a
        # rebuild when they differ: a stale binary (lib/ is git-ignored and survives checkouts) would pass the symbol
        # check below and then misread csmoe_gemm_args.  CSMOE_SKIP_BUILD_CHECK=1 loads whatever is there.
        import os
        if not LIB_PATH.exists() or os.environ.get("CSMOE_SKIP_BUILD_CHECK", "0") != "1":
            from . import build as _build

            _build.build()
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library is stale / incomplete
            fn.restype = res
            fn.argtypes = args
        if lib.csmoe_abi_version() != ABI_VERSION:
            raise RuntimeError("libcsmoe.so ABI version mismatch; rebuild with `python -m competesmoe_b200.build --force`")
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().csmoe_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed with status {rc}: {msg}")
