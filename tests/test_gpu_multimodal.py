"""GPU parity of the multimodal drop-in layer against the golden vectors (outputs of the unmodified reference) and, at
other shapes, against the CPU oracle.  Tolerances: bf16 rtol 2e-2 (plus an atol scaled by the tensor RMS); routing
indices bit-exact except tokens with a top-k margin below 1e-3, which are counted and reported."""
from types import SimpleNamespace

import pytest
import torch

from oracle import multimodal as om

from conftest import load_golden
from helpers import assert_close_rms, build_multimodal_layer, expert_linears

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = ["mm_siglip_router_bf16", "mm_siglip_comp_bf16", "mm_siglip_router_f32", "mm_projector_router_f32",
         "mm_glu_router_f32", "mm_siglip_comp_f32", "mm_projector_comp_f32", "mm_glu_comp_f32",
         "mm_siglip_comp_hybrid_f32", "mm_siglip_comp_normsigmoid_f32"]


def run_layer(layer, fx, dtype):
    x = fx["x"].to(DEV, dtype).requires_grad_(True)
    dy = fx["dy"].to(DEV, dtype)
    out, aux, none, info = layer(x)
    assert none is None and out.dtype == dtype and out.shape == fx["out"].shape
    ((out.float() * dy.float()).sum() + aux.float()).backward()
    return x, out, aux, info


@pytest.mark.parametrize("name", CASES)
def test_layer_matches_reference_golden(name):
    fx = load_golden(name)
    m = fx["meta"]
    dtype = torch.bfloat16            # the tensor-core path computes in bf16 whatever the storage dtype of the fixture
    layer = build_multimodal_layer(fx, DEV, dtype)
    x, out, aux, info = run_layer(layer, fx, dtype)
    sel, w = layer.last_routing
    # ---- routing decisions vs the reference's own
    scores_src = "affinity" if m["competition"] else "gate_softmax"
    args = SimpleNamespace(**m["args"])
    exps = [{k: (v.to(dtype) if torch.is_tensor(v) else v) for k, v in e.items()} for e in fx["experts"]]
    _, _, _, _, dbg = om.competesmoe_forward(fx["x"].to(dtype), fx["gate_w"].to(dtype), exps, m["K"], m["d_out"], args,
                                             m["competition"])
    scores, thr = dbg[scores_src], 1e-3
    if m["competition"] and getattr(args, "norm_sigmoid", False):
        # the selection runs on sigmoid(score) (competesmoe.py:243-247), which squeezes the scores into a range where bf16
        # resolves 2^-8: an fp32 reference run and this bf16 path then differ on tokens within two bf16 steps of a tie
        scores = torch.sigmoid(scores)
        thr = 1e-3 + (2 * 2.0 ** -8 if "float32" in m["dtype"] else 0.0)
    margin = om.topk_margin(scores, m["K"])
    agree = (sel.cpu().long() == fx["selected"]).all(-1)
    n_ex = int((~agree).sum())
    assert bool((margin[~agree] <= thr).all()), "routing differs from the reference on a token with a clear margin"
    print(f"{name}: {n_ex}/{agree.numel()} low-margin tokens exempt from bit-exact routing")
    # ---- values.  bf16 fixtures: the north-star's bf16 rtol 2e-2.  f32 fixtures are reference outputs computed in
    # fp32 while this path computes in bf16 on the tensor cores, so they get twice that.
    rt = 2e-2 if "bfloat16" in m["dtype"] else 4e-2
    assert_close_rms(out[agree.to(DEV)], fx["out"][agree], rt, "output")
    assert_close_rms(w.cpu()[agree], fx["weights"][agree], rt, "routing weights")
    if n_ex == 0:
        assert_close_rms(aux, fx["aux"], rt, "aux loss")
        assert set(info) == set(fx["info"])
        for k in info:
            # the diversity loss is a mean of signed cosines near 0: compare with an absolute floor
            got, ref = float(info[k]), float(fx["info"][k])
            assert abs(got - ref) <= rt * abs(ref) + 2e-3, (k, got, ref)
        assert_close_rms(x.grad[agree.to(DEV)], fx["dx"][agree], 1.5 * rt, "dx")
        if fx["dgate_w"] is not None:
            assert_close_rms(layer.gate.weight.grad, fx["dgate_w"], 1.5 * rt, "dgate")
        for e, (mod, ref) in enumerate(zip(layer.experts, fx["dexperts"])):
            ref_list = list(ref.values())
            l1, l2 = expert_linears(mod)
            params = [l1.weight] + ([l1.bias] if l1.bias is not None else []) + [l2.weight] + ([l2.bias] if l2.bias is not None else [])
            assert len(params) == len(ref_list)
            for p, r in zip(params, ref_list):
                assert p.grad is not None
                assert_close_rms(p.grad, r, 1.5 * rt, f"expert {e} grad {tuple(r.shape)}")


def test_upcycled_experts_analytic_kats():
    """identical experts => output == the dense expert, diversity == 1 - 1/K, balance == 1."""
    fx = load_golden("mm_siglip_comp_upcycled_f32")
    m = fx["meta"]
    layer = build_multimodal_layer(fx, DEV, torch.bfloat16)
    x, out, aux, info = run_layer(layer, fx, torch.bfloat16)
    dense = om.expert_forward({k: (v.bfloat16() if torch.is_tensor(v) else v) for k, v in fx["experts"][0].items()},
                              fx["x"].bfloat16())
    assert_close_rms(out, dense, 2e-2, "upcycled output")
    assert abs(float(info["diversity_loss"]) - (1 - 1 / m["K"])) < 5e-3
    sel, _ = layer.last_routing
    assert sel.cpu().tolist() == [[[0, 1]] * m["N"]] * m["B"]          # exact ties -> lowest index first


def test_checkpoint_layout_and_fused_storage():
    fx = load_golden("mm_siglip_router_bf16")
    layer = build_multimodal_layer(fx, DEV, torch.bfloat16)
    keys = set(layer.state_dict().keys())
    expect = {"gate.weight", "prob_flips"} | {f"experts.{e}.{n}" for e in range(4) for n in
                                               ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias")}
    assert keys == expect
    sd = {k: v.clone() for k, v in layer.state_dict().items()}
    x = fx["x"].to(DEV, torch.bfloat16)
    with torch.no_grad():
        o1 = layer(x)[0]
    # parameters now alias one flat buffer; loading a state dict must still take effect (sparse upcycling path)
    w = [e.fc1.weight for e in layer.experts]
    assert all(w[i].data_ptr() == w[0].data_ptr() + i * w[0].numel() * 2 for i in range(4))
    sd2 = {k: (torch.zeros_like(v) if "experts.1." in k else v) for k, v in sd.items()}
    layer.load_state_dict(sd2)
    with torch.no_grad():
        o2 = layer(x)[0]
    assert not torch.equal(o1, o2)
    layer.load_state_dict(sd)
    with torch.no_grad():
        o3 = layer(x)[0]
    assert torch.equal(o1, o3)                                         # deterministic, bit-identical re-run


def test_inference_path_takes_router_branch_without_aux():
    fx = load_golden("mm_siglip_comp_bf16")
    layer = build_multimodal_layer(fx, DEV, torch.bfloat16).eval()
    with torch.no_grad():
        out, aux, none, info = layer(fx["x"].to(DEV, torch.bfloat16))
    assert info == {} and float(aux) == 0.0 and none is None


@pytest.mark.parametrize("competition", [False, True])
def test_siglip_shape_against_oracle(competition):
    """C2' SigLIP MoE MLP shape (1152 -> 4304 -> 1152, GELU-tanh, bias; intermediate size not a multiple of 64)."""
    torch.manual_seed(0)
    B, N, D, Fh, E, K = 2, 320, 1152, 4304, 4, 2
    g = torch.Generator().manual_seed(1236)
    exps = [{"kind": "mlp", "act": "gelu_tanh", "w1": (torch.randn(Fh, D, generator=g) * D ** -0.5).bfloat16(),
             "b1": (torch.randn(Fh, generator=g) * 0.1).bfloat16(), "w2": (torch.randn(D, Fh, generator=g) * Fh ** -0.5).bfloat16(),
             "b2": (torch.randn(D, generator=g) * 0.1).bfloat16()} for _ in range(E)]
    gate_w = (torch.randn(E, D, generator=g) * 0.02).bfloat16()
    x = torch.randn(B, N, D, generator=g).bfloat16()
    dy = torch.randn(B, N, D, generator=g).bfloat16()
    args = om.default_args()
    fx = {"meta": dict(d_in=D, d_out=D, E=E, K=K, competition=competition, args=vars(args)), "experts": exps,
          "gate_w": gate_w, "x": x, "dy": dy, "out": torch.empty(B, N, D)}
    layer = build_multimodal_layer(fx, DEV, torch.bfloat16)
    xr = x.clone().requires_grad_(True)
    o_out, o_aux, _, o_info, dbg = om.competesmoe_forward(xr, gate_w, exps, K, D, args, competition)
    ((o_out.float() * dy.float()).sum() + o_aux.float()).backward()
    xg, out, aux, info = run_layer(layer, fx, torch.bfloat16)
    sel, w = layer.last_routing
    margin = om.topk_margin(dbg["affinity"] if competition else dbg["gate_softmax"], K)
    agree = (sel.cpu().long() == dbg["selected"]).all(-1)
    assert bool((margin[~agree] < 1e-3).all())
    print(f"siglip-shape competition={competition}: {int((~agree).sum())}/{agree.numel()} low-margin tokens exempt")
    assert_close_rms(out[agree.to(DEV)], o_out.detach()[agree], 2e-2, "output")
    assert_close_rms(xg.grad[agree.to(DEV)], xr.grad[agree], 4e-2, "dx")


@pytest.mark.first_hw_run
@pytest.mark.parametrize("competition", [False, True])
def test_skewed_routing_hot_and_empty_experts_against_oracle(competition):
    """Written after the round's GPU budget was spent: green on the SIMT emulator (tests/test_simt_layers.py), first
    hardware run pending (marker first_hw_run, tests/conftest.py).
    SURVEY.md 8(d) skewed-routing variant: +2.0 on the gate logit of expert 0 for every token (a hot expert) and two
    experts the router never picks (empty segments in the row space, zero weight gradients).  Layer vs oracle: routing,
    output, dx, gate gradient, every expert's weight gradients."""
    from helpers import expert_linears
    B, N, D, Fh, E, K = 2, 96, 128, 264, 8, 2
    g = torch.Generator().manual_seed(4242)
    exps = [{"kind": "mlp", "act": "gelu_tanh", "w1": (torch.randn(Fh, D, generator=g) * D ** -0.5).bfloat16(),
             "b1": (torch.randn(Fh, generator=g) * 0.1).bfloat16(), "w2": (torch.randn(D, Fh, generator=g) * Fh ** -0.5).bfloat16(),
             "b2": (torch.randn(D, generator=g) * 0.1).bfloat16()} for _ in range(E)]
    x = torch.randn(B, N, D, generator=g)
    x[..., 0] = 4.0                                           # a constant feature: gate_w[e, 0] acts as a per-expert bias
    x = x.bfloat16()
    gate_w = torch.randn(E, D, generator=g) * 0.02
    gate_w[:, 0] = 0.0
    gate_w[0, 0] = 0.5                                        # +2.0 on logit 0
    gate_w[6:, 0] = -2.0                                      # -8.0: experts 6 and 7 are never selected by the router
    gate_w = gate_w.bfloat16()
    dy = torch.randn(B, N, D, generator=g).bfloat16()
    args = om.default_args()
    fx = {"meta": dict(d_in=D, d_out=D, E=E, K=K, competition=competition, args=vars(args)), "experts": exps,
          "gate_w": gate_w, "x": x, "dy": dy, "out": torch.empty(B, N, D)}
    layer = build_multimodal_layer(fx, DEV, torch.bfloat16)
    xr = x.clone().requires_grad_(True)
    gw = gate_w.clone().requires_grad_(True)
    ex = [{k: (v.clone().requires_grad_(True) if torch.is_tensor(v) else v) for k, v in e.items()} for e in exps]
    o_out, o_aux, _, o_info, dbg = om.competesmoe_forward(xr, gw, ex, K, D, args, competition)
    ((o_out.float() * dy.float()).sum() + o_aux.float()).backward()
    xg, out, aux, info = run_layer(layer, fx, torch.bfloat16)
    sel, w = layer.last_routing
    if not competition:
        counts = torch.bincount(dbg["selected"].flatten(), minlength=E)
        assert int(counts[0]) == B * N and int(counts[6:].sum()) == 0      # expert 0 takes every token, 6 and 7 none
    margin = om.topk_margin(dbg["affinity"] if competition else dbg["gate_softmax"], K)
    agree = (sel.cpu().long() == dbg["selected"]).all(-1)
    assert bool((margin[~agree] < 1e-3).all())
    n_ex = int((~agree).sum())
    print(f"skewed routing competition={competition}: {n_ex}/{agree.numel()} low-margin tokens exempt")
    assert_close_rms(out[agree.to(DEV)], o_out.detach()[agree], 2e-2, "output")
    if n_ex == 0:
        assert_close_rms(aux, o_aux.detach(), 2e-2, "aux loss")
        assert_close_rms(xg.grad, xr.grad, 3e-2, "dx")
        assert_close_rms(layer.gate.weight.grad, gw.grad, 3e-2, "d gate", outliers=0.02)
        for e, (mod, ref) in enumerate(zip(layer.experts, ex)):
            l1, l2 = expert_linears(mod)
            for p, r, nm in ((l1.weight, ref["w1"], "w1"), (l1.bias, ref["b1"], "b1"), (l2.weight, ref["w2"], "w2"), (l2.bias, ref["b2"], "b2")):
                if r.grad is None or not bool(r.grad.any()):
                    assert p.grad is None or not bool(p.grad.any()), f"expert {e} {nm}: gradient of an expert without tokens"
                else:
                    # bf16 gradients summed over every token (dense pass): the oracle rounds dz op by op, this path once;
                    # seen on the emulator: 1 of 33 792 elements 4 bf16 ulps off (0.125 at |ref| ~ 4) -> 0.1 % outliers
                    assert_close_rms(p.grad, r.grad, 3e-2, f"expert {e} d {nm}", outliers=1e-3)


@pytest.mark.parametrize("name", ["mm_siglip_router_bf16", "mm_siglip_comp_bf16", "mm_glu_router_f32"])
def test_whole_step_cuda_graph_replay_matches_eager(name):
    """competesmoe_b200.graphs.GraphedStep: forward + backward of the layer captured once and replayed (the step has no
    host sync, so it is capturable).  Replays on fresh inputs must reproduce the eager step bit for bit."""
    from competesmoe_b200.graphs import GraphedStep
    fx = load_golden(name)
    dtype = torch.bfloat16
    layer = build_multimodal_layer(fx, DEV, dtype)
    comp = bool(fx["meta"]["competition"])
    g = torch.Generator().manual_seed(7)
    step = GraphedStep(layer, fx["x"].to(DEV, dtype))
    for trial in range(2):      # second trial: new inputs through the same captured graph
        x_cpu = fx["x"] if trial == 0 else torch.randn(fx["x"].shape, generator=g)
        dy_cpu = fx["dy"] if trial == 0 else torch.randn(fx["dy"].shape, generator=g)
        out_g, aux_g = step.run(x_cpu.to(DEV, dtype), dy_cpu.to(DEV, dtype), branch=comp)
        out_g, aux_g, dx_g = out_g.clone(), aux_g.clone(), step.dx.clone()
        grads_g = {n: p.grad.clone() for n, p in layer.named_parameters() if p.grad is not None}
        for p in layer.parameters():
            p.grad = None
        x = x_cpu.to(DEV, dtype).requires_grad_(True)
        out, aux, _, _ = layer(x)
        torch.autograd.backward((out, aux), (dy_cpu.to(DEV, dtype), torch.ones_like(aux)))
        assert torch.equal(out_g, out) and torch.equal(aux_g, aux) and torch.equal(dx_g, x.grad)
        assert set(grads_g) == {n for n, p in layer.named_parameters() if p.grad is not None}
        for n, p in layer.named_parameters():
            if p.grad is not None:
                assert torch.equal(grads_g[n], p.grad), n


@pytest.mark.parametrize("name", ["mm_siglip_router_bf16", "mm_siglip_comp_bf16", "mm_glu_comp_f32"])
def test_layer_cuda_graph_mode_matches_eager(name):
    """layer.enable_cuda_graphs(): the unchanged nn.Module call, replayed from captured forward / backward graphs, gives
    the eager call's outputs, losses, routing and gradients bit for bit -- on the capture inputs and on fresh ones."""
    fx = load_golden(name)
    dtype = torch.bfloat16
    eager = build_multimodal_layer(fx, DEV, dtype)
    graphed = build_multimodal_layer(fx, DEV, dtype).enable_cuda_graphs()
    g = torch.Generator().manual_seed(11)
    for trial in range(3):
        x_cpu = fx["x"] if trial == 0 else torch.randn(fx["x"].shape, generator=g)
        dy = (fx["dy"] if trial == 0 else torch.randn(fx["dy"].shape, generator=g)).to(DEV, dtype)
        res = []
        for layer in (eager, graphed):
            for p in layer.parameters():
                p.grad = None
            x = x_cpu.to(DEV, dtype).requires_grad_(True)
            out, aux, none, info = layer(x)
            assert none is None
            torch.autograd.backward((out, aux), (dy, torch.ones_like(aux)))
            res.append((out.clone(), aux.clone(), x.grad.clone(), {k: v.clone() for k, v in info.items()},
                        layer.last_routing[0].clone(), {n: p.grad.clone() for n, p in layer.named_parameters() if p.grad is not None}))
        (o0, a0, dx0, i0, r0, g0), (o1, a1, dx1, i1, r1, g1) = res
        assert torch.equal(o0, o1) and torch.equal(a0, a1) and torch.equal(dx0, dx1) and torch.equal(r0, r1)
        assert set(i0) == set(i1) and all(torch.equal(i0[k], i1[k]) for k in i0)
        assert set(g0) == set(g1) and all(torch.equal(g0[k], g1[k]) for k in g0)
    assert len(graphed._graphs) == 1
    # eval / no-grad calls bypass the graphs
    graphed.eval()
    with torch.no_grad():
        out, aux, _, _ = graphed(fx["x"].to(DEV, dtype))
    assert out.shape == fx["out"].shape


def test_layer_cuda_graph_mode_under_autocast_matches_eager():
    """fp32 module called under torch.autocast(bf16): graph mode keys on the autocast state and replays what the eager
    call computes, bit for bit."""
    fx = load_golden("mm_siglip_router_f32")
    eager = build_multimodal_layer(fx, DEV, torch.float32)
    graphed = build_multimodal_layer(fx, DEV, torch.float32).enable_cuda_graphs()
    g = torch.Generator().manual_seed(13)
    for trial in range(2):
        x_cpu = fx["x"] if trial == 0 else torch.randn(fx["x"].shape, generator=g)
        dy = fx["dy"].to(DEV)
        res = []
        for layer in (eager, graphed):
            for p in layer.parameters():
                p.grad = None
            x = x_cpu.to(DEV).requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out, aux, _, _ = layer(x)
            torch.autograd.backward((out, aux), (dy.to(out.dtype), torch.ones_like(aux)))
            res.append((out.clone(), aux.clone(), x.grad.clone()))
        assert all(torch.equal(a, b) for a, b in zip(*res))
    assert len(graphed._graphs) == 1


@pytest.mark.first_hw_run
@pytest.mark.parametrize("kind", ["mlp", "glu"])
def test_policy_level_methods_match_the_oracle(kind):
    """The reference's policy-level methods under their own names and signatures -- router_policy(x), topk_expert(logits),
    compute_moe(selected, weights, results, x[, expert_outputs][, return_topk_outputs]), competition_policy(x)
    (competesmoe.py:219-259,301-320; moe.py:113-132,172-213) -- against the oracle's restatement of each, values and
    gradients.  Written after the round's GPU budget was spent (green on the SIMT emulator)."""
    import torch.nn.functional as F
    from test_gpu_edge_cases import _mm_case
    B, N, D, Fh, E, K = 2, 45, 64, 136, 4, 2
    exps, x, gate_w, dy = _mm_case(B, N, D, Fh, E, K, kind, seed=77)
    args = om.default_args()
    layer = build_multimodal_layer({"meta": dict(d_in=D, d_out=D, E=E, K=K, competition=False, args=vars(args)),
                                    "experts": exps, "gate_w": gate_w, "x": x}, DEV, torch.bfloat16)
    ex = [{k: (v.clone().requires_grad_(True) if torch.is_tensor(v) else v) for k, v in e.items()} for e in exps]
    fresh = lambda: x.detach().clone()       # (on the CPU tier `.to(DEV)` is the tensor itself)
    xr = fresh().requires_grad_(True)
    xg = fresh().to(DEV).requires_grad_(True)
    # ---- router_policy / topk_expert
    ow, osel, oprobs, ologits = om.router_policy(xr, gate_w, K)
    w, sel, probs, logits = layer.router_policy(xg)
    assert sel.dtype == torch.int64 and sel.shape == (B, N, K) and probs.shape == (B, N, E) and logits.dtype == torch.bfloat16
    agree = (sel.cpu() == osel).all(-1)
    assert bool((om.topk_margin(oprobs, K)[~agree] < 1e-3).all())
    assert_close_rms(logits, ologits.detach(), 2e-2, "gate logits")
    assert_close_rms(probs, oprobs.detach(), 2e-2, "gate softmax")
    assert_close_rms(w.cpu()[agree], ow.detach()[agree], 2e-2, "routing weights")
    tw, tsel, tprobs = layer.topk_expert(logits.detach())
    ref_p = F.softmax(logits.detach().cpu(), dim=-1, dtype=torch.float32)
    rv, ri = om.stable_topk(ref_p, K)
    assert torch.equal(tsel.cpu(), ri)
    torch.testing.assert_close(tprobs.cpu(), ref_p, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(tw.cpu(), rv, rtol=1e-5, atol=1e-7)
    # ---- compute_moe under the oracle's routing: values, dx, return_topk_outputs, expert_outputs
    o_out = om.compute_moe(xr, ex, osel, ow.detach(), D)
    (o_out.float() * dy.float()).sum().backward()
    res = torch.zeros(B, N, D, dtype=torch.bfloat16, device=DEV)
    out = layer.compute_moe(osel.to(DEV), ow.detach().to(DEV), res, xg)
    assert out is res
    assert_close_rms(out, o_out.detach(), 2e-2, "compute_moe")
    (out.float() * dy.to(DEV).float()).sum().backward()
    assert_close_rms(xg.grad, xr.grad, 3e-2, "compute_moe dx")
    with torch.no_grad():
        dense = [om.expert_forward(e, fresh()) for e in exps]                                  # [E][B, N, D]
        idx = osel.unsqueeze(-1).expand(B, N, K, D)
        div_ref = om.experts_diversity_loss(torch.gather(torch.stack(dense, 2), 2, idx))
    xg2 = fresh().to(DEV).requires_grad_(True)
    out2, diver = layer.compute_moe(osel.to(DEV), ow.detach().to(DEV), torch.zeros_like(res), xg2, return_topk_outputs=True)
    assert_close_rms(out2, o_out.detach(), 2e-2, "compute_moe (return_topk_outputs)")
    assert abs(float(diver) - float(div_ref)) <= 2e-2 * abs(float(div_ref)) + 2e-3
    (out2.float().sum() + diver).backward()
    assert xg2.grad is not None and bool(torch.isfinite(xg2.grad).all())
    out3 = layer.compute_moe(osel.to(DEV), ow.detach().to(DEV), torch.zeros_like(res), fresh().to(DEV),
                             expert_outputs=[d.to(DEV) for d in dense])
    assert_close_rms(out3, o_out.detach(), 2e-2, "compute_moe (expert_outputs)")
    # ---- competition_policy
    xr2 = fresh().requires_grad_(True)
    cw, csel, csoft, caff, ctop = om.competition_policy(xr2, ex, K)
    xg3 = fresh().to(DEV).requires_grad_(True)
    w2, sel2, soft2, aff2, top2 = layer.competition_policy(xg3)
    assert sel2.dtype == torch.int64 and aff2.dtype == torch.bfloat16 and top2.shape == (B, N, K, D)
    agree2 = (sel2.cpu() == csel).all(-1)
    assert bool((om.topk_margin(caff, K)[~agree2] < 1e-3).all())
    assert_close_rms(aff2, caff.detach(), 2e-2, "affinity")
    assert_close_rms(soft2, csoft.detach(), 2e-2, "softmax(affinity)")
    assert_close_rms(w2.cpu()[agree2], cw.detach()[agree2], 2e-2, "competition weights")
    assert_close_rms(top2.cpu()[agree2], ctop.detach()[agree2], 2e-2, "selected experts' outputs")
    if bool(agree2.all()):
        ((ctop.float() * dy.float().unsqueeze(2)).sum() + (cw.float() * 3).sum()).backward()
        ((top2.float() * dy.to(DEV).float().unsqueeze(2)).sum() + (w2.float() * 3).sum()).backward()
        assert_close_rms(xg3.grad, xr2.grad, 3e-2, "competition_policy dx")
