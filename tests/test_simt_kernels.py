"""The bandwidth-bound CUDA kernels (routing maps, router, permute / combine, activations, losses, competition tail,
block tail) run on the CPU from their own source, on a SIMT emulator (tests/simt/simt.h, tests/simt_host.py): the shipped
.cu files are compiled with g++, every CUDA thread is an OS thread, and the C-ABI entry points are called with host
pointers through the product's own wrappers (competesmoe_b200/ops.py with its CUDA guards lifted for the test).

The assertions are the GPU tests' own (tests/test_gpu_kernels.py and friends are called with DEV = "cpu" on smaller
shapes), so what is compared with the oracle here is exactly what is compared on the B200 -- index algebra, reduction
orders, rounding points and tie-breaks of the kernels as written.  TEST INFRASTRUCTURE: the package cannot reach the
emulator (it lives under tests/, and ops.py refuses non-CUDA tensors unless a test lifts the guard), the tensor-core /
TMA kernels are not covered, and a green run here is not a GPU result.
"""
import pytest
import torch
import torch.nn.functional as F

import simt_host
import test_gpu_kernels as gk

from oracle import multimodal as om
from oracle import pretrain as op
from helpers import assert_close_rms


@pytest.fixture(scope="module")
def ops(tmp_path_factory):
    lib, stats = simt_host.build(tmp_path_factory.mktemp("simt"))
    assert stats["bound_symbols"] >= 40, stats
    from competesmoe_b200 import _lib
    from competesmoe_b200 import ops as _ops
    mp = pytest.MonkeyPatch()
    mp.setattr(_lib, "load", lambda: lib)             # ops.py and _lib.check() ask _lib.load() for the library
    mp.setattr(_ops, "_cuda", lambda *ts: None)       # the "CUDA tensors only" guard
    mp.setattr(_ops, "_stream", lambda: None)
    mp.setattr(_ops, "_ROUTER_GEMM", False)           # the tensor-core gate GEMM is not part of the emulated library
    mp.setattr(gk, "DEV", "cpu")
    yield _ops
    mp.undo()


# ------------------------------------------------------------------------------------------------ stage 3: routing maps
@pytest.mark.parametrize("row_tile", [128, 256])
@pytest.mark.parametrize("T,K,E", [(1, 1, 1), (7, 2, 4), (700, 2, 4), (1500, 2, 8), (600, 8, 64), (2100, 1, 3), (513, 3, 1000)])
def test_route_build_bit_exact(ops, T, K, E, row_tile):
    gk.test_route_build_bit_exact(ops, T, K, E, row_tile)


def test_route_build_empty_input(ops):
    gk.test_route_build_empty_input(ops)


def test_route_build_clamps_out_of_range_ids(ops):
    sel = torch.tensor([[0, 9], [-3, 1], [2, 2]], dtype=torch.int32)
    r = ops.route_build(sel, 3)
    assert r.counts.tolist() == [2, 1, 3] and int(r.counts.sum()) == sel.numel()


# ------------------------------------------------------------------------------------------------ stage 1: router
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("T,D,E,K", [(67, 64, 4, 2), (96, 256, 64, 8), (40, 512, 8, 2), (33, 128, 33, 5), (16, 3072, 4, 2)])
def test_router_matches_oracle(ops, dtype, T, D, E, K):
    gk.test_router_matches_oracle(ops, dtype, T, D, E, K)


def test_router_tie_break_is_lowest_index(ops):
    gk.test_router_tie_break_is_lowest_index(ops)


@pytest.mark.parametrize("E,K", [(4, 2), (64, 8), (33, 5)])
def test_topk_renorm(ops, E, K):
    gk.test_topk_renorm(ops, E, K)


def test_router_aux_and_backward_match_autograd(ops):
    gk.test_router_aux_and_backward_match_autograd(ops)


# ------------------------------------------------------------------------------------------------ stages 3 / 5: permute, combine
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_gather_combine_scatter(ops, dtype):
    gk.test_gather_combine_scatter(ops, dtype)


def test_combine_order_matches_reference_rounding(ops):
    gk.test_combine_order_matches_reference_rounding(ops)


# ------------------------------------------------------------------------------------------------ elementwise
@pytest.mark.parametrize("act,fn", [("ACT_RELU", F.relu), ("ACT_GELU", F.gelu), ("ACT_GELU_TANH", lambda z: F.gelu(z, approximate="tanh")),
                                    ("ACT_SILU", F.silu)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_activation_fwd_bwd(ops, act, fn, dtype):
    gk.test_activation_fwd_bwd(ops, act, fn, dtype)


def test_glu_fwd_bwd(ops):
    gk.test_glu_fwd_bwd(ops)
