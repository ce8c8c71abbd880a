"""Edge shapes of both drop-in layers against the CPU oracle: one token, top-k = number of experts, top-1, a single
expert, token counts that are not a multiple of anything, non-contiguous inputs, inputs without gradient, zero tokens.
The reference's own tests do not go there, its semantics do: each case is what
`moe_model/model/moe/competesmoe.py:337-415` / `moe_pretrain_model/layers/moe/competesmoe.py:595-620` compute on that
input, restated by oracle/multimodal.py and oracle/pretrain.py.

Written after the round's GPU budget was spent: green on the SIMT emulator (tests/test_simt_layers.py), first hardware
run pending (marker first_hw_run, tests/conftest.py)."""
from types import SimpleNamespace

import pytest
import torch
import torch.nn.functional as F

from oracle import multimodal as om
from oracle import pretrain as op

from helpers import assert_close_rms, build_multimodal_layer, expert_linears

pytestmark = [pytest.mark.gpu, pytest.mark.first_hw_run]
DEV = "cuda"


# ------------------------------------------------------------------------------------------------ multimodal plugin
def _mm_case(B, N, D, Fh, E, K, kind, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "glu":
        exps = [{"kind": "glu", "act": "silu", "w1": (torch.randn(2 * Fh, D, generator=g) * D ** -0.5).bfloat16(),
                 "w2": (torch.randn(D, Fh, generator=g) * Fh ** -0.5).bfloat16()} for _ in range(E)]
    else:
        exps = [{"kind": "mlp", "act": "gelu_tanh", "w1": (torch.randn(Fh, D, generator=g) * D ** -0.5).bfloat16(),
                 "b1": (torch.randn(Fh, generator=g) * 0.1).bfloat16(), "w2": (torch.randn(D, Fh, generator=g) * Fh ** -0.5).bfloat16(),
                 "b2": (torch.randn(D, generator=g) * 0.1).bfloat16()} for _ in range(E)]
    x = torch.randn(B, N, D, generator=g).bfloat16()
    gate_w = (torch.randn(E, D, generator=g) * 0.3).bfloat16()
    dy = torch.randn(B, N, D, generator=g).bfloat16()
    return exps, x, gate_w, dy


def _mm_check(B, N, D, Fh, E, K, kind, competition, seed, x_grad=True, noncontig=False):
    exps, x, gate_w, dy = _mm_case(B, N, D, Fh, E, K, kind, seed)
    args = om.default_args()
    fx = {"meta": dict(d_in=D, d_out=D, E=E, K=K, competition=competition, args=vars(args)), "experts": exps,
          "gate_w": gate_w, "x": x}
    layer = build_multimodal_layer(fx, DEV, torch.bfloat16)
    # ---- oracle
    xr = x.clone().requires_grad_(x_grad)
    gw = gate_w.clone().requires_grad_(True)
    ex = [{k: (v.clone().requires_grad_(True) if torch.is_tensor(v) else v) for k, v in e.items()} for e in exps]
    # the schedule test at competesmoe.py:347 needs x.requires_grad: without it the reference takes the router branch
    o_out, o_aux, _, o_info, dbg = om.competesmoe_forward(xr, gw, ex, K, D, args, competition and x_grad)
    o_loss = (o_out.float() * dy.float()).sum() + o_aux.float()
    if o_loss.requires_grad:
        o_loss.backward()
    # ---- layer
    if noncontig:       # [N, B, D] storage viewed as [B, N, D]
        xg = x.transpose(0, 1).contiguous().to(DEV).transpose(0, 1).requires_grad_(x_grad)
        assert not xg.is_contiguous() or B == 1 or N == 1
    else:
        xg = x.to(DEV).requires_grad_(x_grad)
    out, aux, none, info = layer(xg)
    assert none is None and out.shape == (B, N, D) and out.dtype == torch.bfloat16
    loss = (out.float() * dy.to(DEV).float()).sum() + aux.float()
    if loss.requires_grad:
        loss.backward()
    if B * N == 0:
        # the reference's losses on this branch are means over zero tokens: NaN there, NaN here
        assert out.numel() == 0 and (xg.grad is None or xg.grad.numel() == 0)
        assert bool(torch.isnan(o_aux)) and bool(torch.isnan(aux)) and set(info) == set(o_info)
        return
    sel, w = layer.last_routing
    took_comp = competition and x_grad
    margin = om.topk_margin(dbg["affinity"] if took_comp else dbg["gate_softmax"], K) if K < E else torch.ones(B, N)
    got_sel = sel.cpu().long().view(B, N, K)
    ref_sel = dbg["selected"]
    agree = (got_sel == ref_sel).all(-1) if K < E else (got_sel.sort(-1).values == ref_sel.sort(-1).values).all(-1)
    assert bool((margin[~agree] < 1e-3).all()), "routing differs on a token with a clear margin"
    n_ex = int((~agree).sum())
    assert_close_rms(out[agree.to(DEV)], o_out.detach()[agree], 2e-2, "output")
    assert set(info) == set(o_info)
    if n_ex:
        return
    assert_close_rms(aux, o_aux.detach(), 2e-2, "aux loss")
    if x_grad:
        assert_close_rms(xg.grad, xr.grad, 3e-2, "dx")
    else:
        assert xg.grad is None
    if gw.grad is not None and K == 1 and not took_comp:
        # Top-1 routing weight of a bf16 model: w = p / bf16(p) = 1 +- 2^-9, whose gradient
        # dtw / r + bf16(-dtw (w / r)) is what is left of two terms that cancel to their last bf16 digit -- a sawtooth of
        # dtw with a period of one bf16 ulp.  Given the SAME dtw the kernel reproduces autograd's value bit for bit
        # (test_gpu_kernels.py::test_router_backward_reproduces_autograd_through_the_rounded_denominator); dtw itself
        # (<dy, expert output>) agrees with the oracle's to ~1e-3, a full period, so only the magnitude can be compared.
        got, ref = layer.gate.weight.grad.float().cpu(), gw.grad.float()
        assert float((got - ref).norm()) <= 2.0 * float(ref.norm()), "d gate (top-1): magnitude"
    elif gw.grad is not None:
        assert_close_rms(layer.gate.weight.grad, gw.grad, 3e-2, "d gate", outliers=0.02)
    for e, (mod, ref) in enumerate(zip(layer.experts, ex)):
        l1, l2 = expert_linears(mod)
        pairs = [(l1.weight, ref["w1"], "w1"), (l2.weight, ref["w2"], "w2")]
        if kind != "glu":
            pairs += [(l1.bias, ref["b1"], "b1"), (l2.bias, ref["b2"], "b2")]
        for p, r, nm in pairs:
            if r.grad is None or not bool(r.grad.any()):
                assert p.grad is None or not bool(p.grad.any()), f"expert {e} {nm}: gradient of an expert without tokens"
            else:
                assert_close_rms(p.grad, r.grad, 3e-2, f"expert {e} d {nm}", outliers=1e-3)


MM_EDGE = {
    "one token": dict(B=1, N=1, D=64, Fh=128, E=4, K=2, kind="mlp"),
    "top-k = E": dict(B=2, N=37, D=64, Fh=136, E=4, K=4, kind="mlp"),
    "top-1": dict(B=1, N=131, D=128, Fh=264, E=8, K=1, kind="mlp"),
    "single expert": dict(B=2, N=19, D=64, Fh=128, E=1, K=1, kind="glu"),
    "odd sizes glu": dict(B=3, N=43, D=72, Fh=104, E=5, K=3, kind="glu"),
    "129 tokens": dict(B=1, N=129, D=64, Fh=128, E=4, K=2, kind="mlp"),
}


@pytest.mark.parametrize("competition", [False, True])
@pytest.mark.parametrize("case", list(MM_EDGE))
def test_multimodal_edge_shape_matches_oracle(case, competition):
    _mm_check(**MM_EDGE[case], competition=competition, seed=100 + len(case))


@pytest.mark.parametrize("competition", [False, True])
def test_multimodal_non_contiguous_input(competition):
    _mm_check(B=3, N=50, D=64, Fh=128, E=4, K=2, kind="mlp", competition=competition, seed=7, noncontig=True)


@pytest.mark.parametrize("competition", [False, True])
def test_multimodal_input_without_gradient(competition):
    """Frozen upstream (x.requires_grad False) in training mode: the reference's schedule test (competesmoe.py:347) then
    takes the router branch whatever the flip says, and the parameters still get their gradients."""
    _mm_check(B=2, N=40, D=64, Fh=128, E=4, K=2, kind="mlp", competition=competition, seed=8, x_grad=False)


@pytest.mark.parametrize("competition", [False, True])
def test_multimodal_zero_tokens(competition):
    _mm_check(B=2, N=0, D=64, Fh=128, E=4, K=2, kind="mlp", competition=competition, seed=9)


# ------------------------------------------------------------------------------------------------ pretrain plugin
def _pt_check(B, N, D, H, E, K, competition, seed, noncontig=False, **argkw):
    from competesmoe_b200.pretrain import CompeteSMoE
    base = dict(warm_up=0.0, rate_flip=0.07, stop_after=8, max_compete_in_iter=3, is_cosine=False, is_norm_weight=False,
                norm_sigmoid=False, scale_weight=1.0, hybrid=False, tribrid=False, in_topk=False, balance_affinity=False,
                balance_loss_coef=0.01, balance_loss_coef_comp=0.01, router_loss_coef=0.01, router_theta=1.0, test_only=False)
    base.update(argkw)
    args = SimpleNamespace(**base)
    torch.manual_seed(seed)
    layer = CompeteSMoE(D, E, H, n_heads=K, args=args, activation=F.relu, selection_mode="gate", log_interval=None)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        layer.w_gate.copy_(torch.randn(E, D, generator=g) * 0.3)
        layer.keys.copy_(torch.randn(E, D, H, generator=g) * D ** -0.5)
        layer.values.copy_(torch.randn(E, H, D, generator=g) * H ** -0.5)
    w_gate, keys, values = (p.detach().clone() for p in (layer.w_gate, layer.keys, layer.values))
    layer = layer.to(DEV)
    layer.train()
    layer.regularization_present = True
    layer.step_warm = 0
    layer.prob_flips_final = {0: torch.full((8,), bool(competition), device=DEV)}
    layer.set_current_steps(1)
    x = torch.randn(B, N, D, generator=g)
    dy = torch.randn(B, N, D, generator=g)
    xr, wg, ks, vs = (t.clone().requires_grad_(True) for t in (x, w_gate, keys, values))
    o_out, o_regs, dbg = op.competesmoe_forward(xr, wg, ks, vs, K, args, competition, op_dtype=torch.bfloat16)
    ((o_out.float() * dy).sum() + sum(o_regs.values())).backward()
    if noncontig:
        xg = x.transpose(0, 1).contiguous().to(DEV).transpose(0, 1).requires_grad_(True)
    else:
        xg = x.to(DEV).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = layer(xg, id_layer=0)
        regs = layer.get_reg_loss()
    assert out.shape == (B, N, D) and set(regs) == set(o_regs)
    ((out.float() * dy.to(DEV)).sum() + sum(regs.values())).backward()
    sel, w = layer.last_routing
    margin = om.topk_margin(dbg["affinity"] if competition else dbg["gate_softmax"], K) if K < E else torch.ones(B, N)
    got_sel, ref_sel = sel.cpu().long().view(B, N, K), dbg["selected"]
    agree = (got_sel == ref_sel).all(-1) if K < E else (got_sel.sort(-1).values == ref_sel.sort(-1).values).all(-1)
    assert bool((margin[~agree] < 1e-3).all()), "routing differs on a token with a clear margin"
    assert_close_rms(out[agree.to(DEV)], o_out.detach()[agree], 2e-2, "output")
    if int((~agree).sum()):
        return
    for k in regs:
        got, ref = float(regs[k].detach()), float(o_regs[k].detach())
        assert abs(got - ref) <= 3e-2 * abs(ref) + 2e-5, (k, got, ref)
    assert_close_rms(xg.grad, xr.grad, 3e-2, "dx")
    assert_close_rms(layer.keys.grad, ks.grad, 3e-2, "dkeys", outliers=1e-3)
    assert_close_rms(layer.values.grad, vs.grad, 3e-2, "dvalues", outliers=1e-3)
    assert_close_rms(layer.w_gate.grad, wg.grad, 3e-2, "dw_gate", outliers=0.02)


PT_EDGE = {
    "one token": dict(B=1, N=1, D=64, H=32, E=8, K=2),
    "top-k = E": dict(B=2, N=21, D=64, H=16, E=4, K=4),
    "top-1": dict(B=1, N=77, D=64, H=32, E=8, K=1),
    "single expert": dict(B=2, N=9, D=64, H=32, E=1, K=1),
    "odd sizes": dict(B=3, N=43, D=72, H=24, E=5, K=3),
    "expert size 128": dict(B=2, N=65, D=128, H=128, E=8, K=2),       # the fused sigma-MoE kernels on the GPU
}


@pytest.mark.parametrize("competition", [False, True])
@pytest.mark.parametrize("case", list(PT_EDGE))
def test_pretrain_edge_shape_matches_oracle(case, competition):
    _pt_check(**PT_EDGE[case], competition=competition, seed=200 + len(case))


@pytest.mark.parametrize("competition", [False, True])
def test_pretrain_non_contiguous_input(competition):
    _pt_check(B=3, N=30, D=64, H=32, E=8, K=2, competition=competition, seed=17, noncontig=True)


@pytest.mark.parametrize("competition", [False, True])
def test_pretrain_zero_tokens(competition):
    """The reference cannot train on an empty batch (math.log(0) in entropy_balance, moe.py:323-332; an ambiguous view in
    the competition step, competesmoe.py:399): an error there, a ValueError here; in eval mode both return an empty result."""
    from competesmoe_b200.pretrain import CompeteSMoE
    args = SimpleNamespace(warm_up=0.0, rate_flip=0.07, stop_after=8, max_compete_in_iter=3, is_cosine=False, is_norm_weight=False,
                           norm_sigmoid=False, scale_weight=1.0, hybrid=False, tribrid=False, in_topk=False, balance_affinity=False,
                           balance_loss_coef=0.01, balance_loss_coef_comp=0.01, router_loss_coef=0.01, router_theta=1.0, test_only=False)
    layer = CompeteSMoE(64, 8, 32, n_heads=2, args=args, activation=F.relu, selection_mode="gate", log_interval=None).to(DEV)
    layer.train()
    layer.regularization_present = True
    layer.step_warm = 0
    layer.prob_flips_final = {0: torch.full((8,), bool(competition), device=DEV)}
    layer.set_current_steps(1)
    x = torch.zeros(2, 0, 64, device=DEV, requires_grad=True)
    with pytest.raises(ValueError, match="no tokens"):
        layer(x, id_layer=0)
    layer.eval()
    with torch.no_grad():
        out = layer(torch.zeros(2, 0, 64, device=DEV), id_layer=0)
    assert out.shape == (2, 0, 64)
