"""Whole layers on the CPU: the plugins' host code (competesmoe_b200/{multimodal,pretrain,siblings,pretrain_siblings,
functional,cvmm}.py) driving the shipped non-GEMM kernels on the SIMT emulator (tests/simt/), checked against the golden
fixtures of the UNMODIFIED reference with the GPU tests' own assertions (the `-m gpu` test functions are called with
DEV = "cpu").

Two stand-ins, both test infrastructure and both stated here so that nobody reads more into a green run than it says:
  * csmoe_grouped_gemm is tests/simt/ref_gemm.cpp, a plain-loop statement of the entry point's contract with the
    kernel's rounding points -- NOT the tcgen05 kernel (which, like the fused sigma-MoE kernels, is tested on the GPU
    only; the fused path is switched off here and the layers take the grouped-GEMM path);
  * CUDA autocast does not exist on a CPU-only torch, so `torch.autocast / is_autocast_enabled / get_autocast_dtype` are
    replaced by a flag for the duration of the module: the pretrain layer reads exactly these to pick its compute dtype.
What IS verified without a GPU: routing decisions and maps, router / loss / competition-tail / combine kernels as shipped,
every autograd.Function's wiring (saved tensors, gradient routing, regulariser names), the schedule, checkpoint layout --
against reference outputs, losses and gradients.
"""
import contextlib

import pytest
import torch

import simt_host
import test_gpu_edge_cases as ge
import test_gpu_fp32 as gf
import test_gpu_multimodal as gm
import test_gpu_pretrain as gp
import test_gpu_pretrain_siblings as gps
import test_gpu_siblings as gs


class _Fresh(dict):
    """Fixture dict whose tensors come out as fresh copies: on the GPU `fx["x"].to("cuda")` is a copy, on the CPU it is
    the fixture tensor itself, and `requires_grad_()` on it would leak into the next reader."""

    def __getitem__(self, k):
        v = dict.__getitem__(self, k)
        return v.clone() if torch.is_tensor(v) else v


@pytest.fixture(scope="module", autouse=True)
def emulated(tmp_path_factory):
    lib, stats = simt_host.build(tmp_path_factory.mktemp("simt_layers"), ref_gemm=True)
    from competesmoe_b200 import _lib, functional, ops
    mp = pytest.MonkeyPatch()
    mp.setattr(_lib, "load", lambda: lib)
    mp.setattr(ops, "_cuda", lambda *ts: None)
    mp.setattr(ops, "_stream", lambda: None)
    mp.setattr(ops, "_ROUTER_GEMM", False)
    mp.setattr(functional, "_SIGMA_FUSED", False)
    state = {"on": False, "dtype": torch.bfloat16}

    @contextlib.contextmanager
    def autocast(device_type, dtype=torch.bfloat16, enabled=True, cache_enabled=None):
        old = dict(state)
        state.update(on=bool(enabled), dtype=dtype)
        try:
            yield
        finally:
            state.update(old)

    mp.setattr(torch, "autocast", autocast)
    mp.setattr(torch, "is_autocast_enabled", lambda *a: state["on"])
    mp.setattr(torch, "get_autocast_dtype", lambda dev: state["dtype"])
    mp.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
    for mod in (gm, gp, gs, gps, gf, ge):
        mp.setattr(mod, "DEV", "cpu")
        if hasattr(mod, "load_golden"):
            mp.setattr(mod, "load_golden", (lambda f: lambda name: _Fresh(f(name)))(mod.load_golden))
    yield
    mp.undo()


# ------------------------------------------------------------------------------------------------ multimodal plugin
@pytest.mark.parametrize("name", gm.CASES)
def test_multimodal_layer_matches_reference_golden(name):
    gm.test_layer_matches_reference_golden(name)


def test_multimodal_upcycled_experts_analytic_kats():
    gm.test_upcycled_experts_analytic_kats()


def test_multimodal_checkpoint_layout_and_fused_storage():
    gm.test_checkpoint_layout_and_fused_storage()


def test_multimodal_inference_path_takes_router_branch_without_aux():
    gm.test_inference_path_takes_router_branch_without_aux()


@pytest.mark.parametrize("competition", [False, True])
def test_multimodal_skewed_routing_hot_and_empty_experts(competition):
    gm.test_skewed_routing_hot_and_empty_experts_against_oracle(competition)


@pytest.mark.parametrize("kind", ["mlp", "glu"])
def test_multimodal_policy_level_methods_match_the_oracle(kind):
    gm.test_policy_level_methods_match_the_oracle(kind)


@pytest.mark.parametrize("name", gs.SIB)
def test_multimodal_sibling_matches_reference_golden(name):
    gs.test_sibling_matches_reference_golden(name)


# ------------------------------------------------------------------------------------------------ pretrain plugin
@pytest.mark.parametrize("name", gp.PT)
def test_pretrain_layer_matches_reference_golden(name):
    gp.test_pretrain_layer_matches_reference_golden(name)


@pytest.mark.parametrize("name", gp.PT_WIDE)
def test_pretrain_layer_with_128_experts_matches_reference_golden(name):
    gp.test_pretrain_layer_with_128_experts_matches_reference_golden(name)


@pytest.mark.parametrize("autocast,variant", [(True, {}), (True, {"norm_sigmoid": True, "scale_weight": 2.0}), (False, {}),
                                              (True, {"is_cosine": True})])
def test_pretrain_policy_level_methods_match_the_oracle(autocast, variant):
    gp.test_policy_level_methods_match_the_oracle(autocast, variant)


def test_cvmm_op_both_call_patterns():
    gp.test_cvmm_op_both_call_patterns()


def test_cvmm_moe_attention_call_patterns():
    gp.test_cvmm_moe_attention_call_patterns()


def test_cvmm_rejects_index_tensors_it_cannot_express():
    gp.test_cvmm_rejects_index_tensors_it_cannot_express()


def test_cvmm_triton_library_op_is_fp32_accurate_for_fp32_operands():
    gp.test_cvmm_triton_library_op_is_fp32_accurate_for_fp32_operands(direct=True)


def test_moe_attention_projection_layer_is_att():
    gp.test_moe_attention_projection_layer_is_att()


@pytest.mark.parametrize("name", gps.PTSIB)
def test_pretrain_sibling_matches_reference_golden(name):
    gps.test_pretrain_sibling_matches_reference_golden(name)


@pytest.mark.parametrize("autocast", [True, False])
def test_moe_attention_projection_matches_reference_golden(autocast):
    gps.test_moe_attention_projection_matches_reference_golden(autocast)


# ------------------------------------------------------------------------------------------------ fp32 callers (rtol 1e-4)
@pytest.mark.parametrize("name", ["mm_siglip_router_f32", "mm_glu_router_f32", "mm_siglip_comp_f32", "mm_projector_comp_f32"])
def test_multimodal_fp32_module_matches_reference_at_1e4(name):
    gf.test_multimodal_fp32_module_matches_reference_at_1e4(name)


def test_cvmm_op_fp32_without_autocast_matches_the_reference_kernels_at_1e4():
    gf.test_cvmm_op_fp32_without_autocast_matches_the_reference_kernels_at_1e4()


@pytest.mark.parametrize("name", ["pt_router_f32", "pt_comp_tribrid_f32", "pt_router_e128_f32", "pt_comp_e128_f32"])
def test_pretrain_fp32_without_autocast_matches_reference_at_1e4(name):
    gf.test_pretrain_fp32_without_autocast_matches_reference_at_1e4(name)


# ------------------------------------------------------------------------------------------------ edge shapes, both plugins
@pytest.mark.parametrize("competition", [False, True])
@pytest.mark.parametrize("case", list(ge.MM_EDGE))
def test_multimodal_edge_shape_matches_oracle(case, competition):
    ge.test_multimodal_edge_shape_matches_oracle(case, competition)


@pytest.mark.parametrize("competition", [False, True])
def test_multimodal_non_contiguous_input_without_gradient_and_zero_tokens(competition):
    ge.test_multimodal_non_contiguous_input(competition)
    ge.test_multimodal_input_without_gradient(competition)
    ge.test_multimodal_zero_tokens(competition)


@pytest.mark.parametrize("competition", [False, True])
@pytest.mark.parametrize("case", list(ge.PT_EDGE))
def test_pretrain_edge_shape_matches_oracle(case, competition):
    ge.test_pretrain_edge_shape_matches_oracle(case, competition)


@pytest.mark.parametrize("competition", [False, True])
def test_pretrain_non_contiguous_input_and_zero_tokens(competition):
    ge.test_pretrain_non_contiguous_input(competition)
    ge.test_pretrain_zero_tokens(competition)


# ------------------------------------------------------------------------------------------------ gate GEMM branch of the router ops
@pytest.mark.parametrize("T,D,E,K,renorm", [(256, 64, 16, 2, None), (256, 64, 16, 2, torch.float32), (256, 64, 72, 4, None)])
def test_router_ops_gate_gemm_branch_matches_the_fused_kernels(T, D, E, K, renorm, monkeypatch):
    """ops.router_fwd / router_bwd take the tensor-core gate GEMM for E >= 16 on full row tiles (csmoe_router_from_logits,
    dx / dWg as grouped GEMMs).  The host wiring of that branch -- here over the plain-loop GEMM stand-in -- against the
    fused CUDA-core kernels on the same inputs: same routing, same weights, gradients to bf16 rounding."""
    from competesmoe_b200 import ops
    g = torch.Generator().manual_seed(T + E)
    x = torch.randn(T, D, generator=g).bfloat16()
    wg = (torch.randn(E, D, generator=g) * 0.1).bfloat16()
    dtw = torch.randn(T, K, generator=g)
    res = {}
    for flag in (False, True):
        monkeypatch.setattr(ops, "_ROUTER_GEMM", flag)
        assert ops._router_gemm_ok(T, D, E, x.dtype) == flag
        logits, probs, tw, ti = ops.router_fwd(x, wg, K, renorm_dtype=renorm)
        dx, dwg = ops.router_bwd(x, wg, probs, tw, ti, 1, dtw=dtw, renorm_dtype=renorm)
        res[flag] = (logits, probs, tw, ti, dx, dwg)
    a, b = res[False], res[True]
    gf.assert_close_rms(b[0], a[0], 1e-2, "logits")
    same = (a[3] == b[3]).all(-1)
    assert float(same.float().mean()) > 0.98                      # a bf16 ulp on a logit may swap a near-tie
    torch.testing.assert_close(b[2][same], a[2][same], rtol=2e-2, atol=1e-4)
    if bool(same.all()):
        gf.assert_close_rms(b[4], a[4], 2e-2, "dx")
        gf.assert_close_rms(b[5], a[5], 2e-2, "d gate")


# ------------------------------------------------------------------------------------------------ `args.test_only` statistics
@pytest.mark.parametrize("name", ["competesmoe", "smoe"])
def test_pretrain_test_only_statistics_match_the_reference(name):
    """Expert-usage histogram and routing entropies the evaluation harness reads after a `-test_only` run
    (layers/moe/moe.py:145-183, competesmoe.py:607-611): the unmodified reference class and this layer on the same weights
    and two eval batches.  Needs the reference tree (present in the build container, where this tier runs)."""
    import importlib
    import os
    from pathlib import Path
    import torch.nn.functional as F
    if not Path("/root/reference/moe_pretrain_model").exists():
        pytest.skip("the reference tree is not on this machine")
    import sys
    import types
    from oracle import gen_golden as gg
    from oracle import pretrain as op
    import competesmoe_b200.pretrain_siblings  # noqa: F401
    from competesmoe_b200.pretrain import get_moe
    # The reference's layers/cvmm.py defines the torch.library op mylib::cvmm_triton at import: once per process, and this
    # process may hold the package's definition already.  Its Triton kernels need a GPU anyway, so `layers.cvmm` is the
    # oracle's restatement of the op here (pinned on the reference's kernels by cvmm_triton_interp); the LAYER classes --
    # what this test is about -- are the unmodified reference's.
    cv = types.ModuleType("layers.cvmm")
    cv.CVMMSel, cv.cvmm_prepare_sel2 = op.Sel, op.prepare_sel2
    cv.cvmm = lambda x, sel, keys: op.cvmm(x, sel, keys, torch.float32)

    def _no_sel(*a, **k):
        raise NotImplementedError("cvmm_prepare_sel is not used by the layers under test")

    cv.cvmm_prepare_sel = _no_sel
    for k in [k for k in sys.modules if k == "layers" or k == "framework" or k.startswith(("layers.", "framework."))]:
        sys.modules.pop(k)
    sys.modules["layers.cvmm"] = cv
    pm = gg.load_pretrain_reference()
    if name != "competesmoe":
        with gg.quiet():
            importlib.import_module("layers.moe." + gg.PT_SIBLING_MODULES[name])
    args = gg.pt_args(test_only=True)
    D, E, H, K = 64, 8, 32, 2
    torch.manual_seed(3)
    cwd = os.getcwd()
    os.chdir("/tmp")
    try:
        with gg.quiet():
            ref = pm["get_moe"](name)(D, E, H, n_heads=K, args=args, activation=F.relu, selection_mode="gate", log_interval=None)
            if name == "competesmoe":
                ref.set_total_steps(id_layer=0)
    finally:
        os.chdir(cwd)
    ours = get_moe(name)(D, E, H, n_heads=K, args=args, activation=F.relu, selection_mode="gate", log_interval=None)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ref.eval(), ours.eval()
    g = torch.Generator().manual_seed(9)
    with torch.no_grad():
        for _ in range(2):
            x = torch.randn(2, 40, D, generator=g)
            a = ref(x, id_layer=0) if name == "competesmoe" else ref(x)
            b = ours(x, id_layer=0) if name == "competesmoe" else ours(x)
            gf.assert_close_rms(b, a, 1e-4, "output")
    assert torch.equal(ours.get_dist_experts().cpu(), ref.get_dist_experts())
    assert int(ours.get_dist_experts().sum()) == 2 * 2 * 40 * K
    wr, wo = ref.get_weight_dist(), ours.get_weight_dist()
    assert set(wr) == set(wo)
    for k in wr:
        assert abs(wo[k] - wr[k]) <= 1e-5 * abs(wr[k]) + 1e-7, (k, wo[k], wr[k])
    for k in [k for k in sys.modules if k == "layers" or k == "framework" or k.startswith(("layers.", "framework."))]:
        sys.modules.pop(k)
