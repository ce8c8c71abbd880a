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
import test_gpu_block as gb
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
    mp.setattr(gb, "DEV", "cpu")
    yield _ops
    mp.undo()


# ------------------------------------------------------------------------------------------------ stage 3: routing maps
@pytest.mark.parametrize("row_tile", [128, 256])
@pytest.mark.parametrize("T,K,E", [(1, 1, 1), (7, 2, 4), (700, 2, 4), (1500, 2, 8), (600, 8, 64), (2100, 1, 3), (513, 3, 1000)])
def test_route_build_bit_exact(ops, T, K, E, row_tile):
    gk.test_route_build_bit_exact(ops, T, K, E, row_tile)


def test_route_build_empty_input(ops):
    gk.test_route_build_empty_input(ops)


def test_ops_accept_empty_inputs(ops):
    gk.test_ops_accept_empty_inputs(ops)


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


@pytest.mark.parametrize("B,N,D,E,K", [(2, 60, 128, 64, 4), (1, 50, 64, 33, 2)])
def test_router_backward_with_more_experts_than_column_lanes(ops, B, N, D, E, K):
    gk.test_router_aux_and_backward_match_autograd(ops, B, N, D, E, K)


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


def test_bias_grad_and_cast(ops):
    gk.test_bias_grad_and_cast(ops)


@pytest.mark.parametrize("act,dense", [("relu", False), ("gelu_tanh", True), ("gelu", False), ("silu", True)])
def test_act_bwd_bias_matches_separate_kernels(ops, act, dense):
    gk.test_act_bwd_bias_matches_separate_kernels(ops, act, dense)


def test_split_bf16x3_is_exact_to_24_bits(ops):
    """csmoe_split_f32_bf16x3 (the operands of the fp32-accurate products): hi + mid + lo == x to fp32 precision, each
    term a bf16 number, hi the round-to-nearest of x."""
    g = torch.Generator().manual_seed(6)
    x = torch.randn(4099, generator=g) * torch.logspace(-6, 6, 4099)
    hi, mid, lo = ops.split_bf16x3(x)
    assert hi.dtype == mid.dtype == lo.dtype == torch.bfloat16
    assert torch.equal(hi, x.bfloat16())
    total = hi.double() + mid.double() + lo.double()
    assert float(((total - x.double()).abs() / x.double().abs().clamp_min(1e-30)).max()) <= 2.0 ** -22


# ------------------------------------------------------------------------------------------------ stage 2: competition tail
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_affinity_fwd_bwd(ops, dtype):
    gk.test_affinity_fwd_bwd(ops, dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("E,K,T,t_pad,D", [(4, 2, 70, 256, 1152), (8, 1, 5, 256, 64), (64, 8, 20, 256, 1024), (6, 3, 40, 256, 264)])
def test_diversity_and_fused_competition_backward(ops, dtype, E, K, T, t_pad, D):
    gk.test_diversity_and_fused_competition_backward(ops, dtype, E, K, T, t_pad, D)


@pytest.mark.parametrize("sigmoid", [False, True])
def test_topk_renorm_bwd_and_dense_rows(sigmoid, ops):
    gk.test_topk_renorm_bwd_and_dense_rows(sigmoid, ops)


def test_topk_is_total_on_non_finite_scores(ops):
    gk.test_topk_is_total_on_non_finite_scores(ops)


# ------------------------------------------------------------------------------------------------ losses
@pytest.mark.parametrize("B,N,E,K", [(1, 300, 4, 2), (3, 200, 8, 2), (2, 130, 64, 8), (2, 77, 33, 5)])
def test_compete_losses_kernels_match_torch(ops, B, N, E, K):
    gk.test_compete_losses_kernels_match_torch(B, N, E, K)


@pytest.mark.parametrize("B,N,E", [(1, 512, 4), (3, 100, 8), (2, 96, 64)])
def test_entropy_balance_kernel_matches_pretrain_formula(ops, B, N, E):
    gk.test_entropy_balance_kernel_matches_pretrain_formula(B, N, E)


# ------------------------------------------------------------------------------------------------ more than 64 experts
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("T,D,E,K", [(40, 256, 128, 4), (24, 512, 128, 8), (33, 128, 200, 8), (20, 64, 256, 2), (30, 128, 65, 3)])
def test_router_matches_oracle_wide(ops, dtype, T, D, E, K):
    gk.test_router_matches_oracle(ops, dtype, T, D, E, K)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("T,D,E,K", [(77, 64, 8, 2), (50, 128, 64, 8), (40, 128, 128, 8), (21, 64, 256, 5), (33, 64, 97, 1)])
def test_router_from_logits_equals_fused_router(ops, dtype, T, D, E, K):
    gk.test_router_from_logits_equals_fused_router(ops, dtype, T, D, E, K)


def test_router_tie_break_is_lowest_index_wide(ops):
    gk.test_router_tie_break_is_lowest_index_wide(ops)


@pytest.mark.parametrize("B,N,D,E,K", [(1, 131, 128, 8, 1), (2, 90, 64, 8, 2), (1, 24, 128, 128, 1), (1, 64, 64, 4, 4)])
def test_router_backward_reproduces_autograd_through_the_rounded_denominator(ops, B, N, D, E, K):
    gk.test_router_backward_reproduces_autograd_through_the_rounded_denominator(ops, B, N, D, E, K)


def test_router_renorm_dtype_is_the_layer_inputs(ops):
    gk.test_router_renorm_dtype_is_the_layer_inputs(ops)


@pytest.mark.parametrize("E,K", [(65, 3), (128, 8), (200, 5), (256, 8)])
def test_topk_renorm_wide(ops, E, K):
    gk.test_topk_renorm(ops, E, K, T=60)


@pytest.mark.parametrize("B,N,D,E,K", [(2, 20, 128, 128, 4), (1, 24, 64, 200, 8), (1, 40, 64, 72, 2)])
def test_router_aux_and_backward_match_autograd_wide(ops, B, N, D, E, K):
    gk.test_router_aux_and_backward_match_autograd(ops, B, N, D, E, K)


def test_topk_is_total_on_non_finite_scores_wide(ops):
    gk.test_topk_is_total_on_non_finite_scores_wide(ops)


@pytest.mark.parametrize("B,N,E,K", [(2, 130, 128, 8), (3, 77, 200, 5), (1, 150, 65, 2), (2, 64, 256, 8)])
def test_compete_losses_kernels_match_torch_wide(ops, B, N, E, K):
    gk.test_compete_losses_kernels_match_torch(B, N, E, K)


@pytest.mark.parametrize("B,N,E", [(2, 200, 128), (3, 96, 256), (1, 130, 100)])
def test_entropy_balance_kernel_matches_pretrain_formula_wide(ops, B, N, E):
    gk.test_entropy_balance_kernel_matches_pretrain_formula(B, N, E)


# ------------------------------------------------------------------------------------------------ block tail
def test_layernorm_cast_kernel_matches_torch(ops):
    gb.test_layernorm_cast_kernel_matches_torch()


def test_residual_dropout_mask_is_consistent_between_forward_and_backward(ops):
    gb.test_residual_dropout_mask_is_consistent_between_forward_and_backward()


def test_fused_combine_tail_applies_dropout_like_the_stand_alone_kernel(ops):
    gb.test_fused_combine_tail_applies_dropout_like_the_stand_alone_kernel()
