"""The oracle against the committed golden vectors (outputs of the unmodified reference, see oracle/gen_golden.py).
CPU only; this is what pins the oracle."""
from types import SimpleNamespace

import pytest
import torch
import torch.nn.functional as F

from oracle import multimodal as om
from oracle import pretrain as op

from conftest import load_golden

MM = ["mm_siglip_router_f32", "mm_projector_router_f32", "mm_glu_router_f32", "mm_siglip_comp_f32",
      "mm_projector_comp_f32", "mm_glu_comp_f32", "mm_siglip_comp_hybrid_f32", "mm_siglip_router_bf16",
      "mm_siglip_comp_bf16", "mm_siglip_comp_upcycled_f32", "mm_siglip_comp_normsigmoid_f32"]
PT = ["pt_router_f32", "pt_comp_f32", "pt_comp_hybrid_bal_f32", "pt_comp_intopk_f32", "pt_comp_tribrid_f32",
      "pt_router_cosine_f32", "pt_router_normweight_f32", "pt_router_normsigmoid_f32", "pt_comp_cosine_f32",
      "pt_router_bias_f32", "pt_comp_bias_f32", "pt_router_e128_f32", "pt_comp_e128_f32"]


def _req(t):
    return t.clone().requires_grad_(True)


@pytest.mark.parametrize("name", MM)
def test_multimodal_oracle_matches_reference(name):
    fx = load_golden(name)
    m = fx["meta"]
    args = SimpleNamespace(**m["args"])
    x, gw = _req(fx["x"]), _req(fx["gate_w"])
    exps = [{k: (_req(v) if torch.is_tensor(v) else v) for k, v in e.items()} for e in fx["experts"]]
    out, aux, _, info, dbg = om.competesmoe_forward(x, gw, exps, m["K"], m["d_out"], args, m["competition"])
    ((out.float() * fx["dy"].float()).sum() + aux.float()).backward()
    f32 = "float32" in m["dtype"]
    tol = dict(rtol=1e-5, atol=1e-6) if f32 else dict(rtol=2e-2, atol=2e-2)
    scores = dbg["affinity"] if m["competition"] else dbg["gate_softmax"]
    margin = om.topk_margin(scores, m["K"])
    agree = (fx["selected"] == dbg["selected"]).all(-1)
    assert bool((margin[~agree] < 1e-3).all())
    assert int((~agree).sum()) == fx["n_exempt"]
    if not m["upcycled"]:
        torch.testing.assert_close(out[agree], fx["out"][agree], **tol)
        torch.testing.assert_close(x.grad[agree], fx["dx"][agree], **tol)
    else:
        # analytic KATs of sparse upcycling: identical experts
        dense = om.expert_forward(fx["experts"][0], fx["x"])
        torch.testing.assert_close(out, dense, rtol=1e-4, atol=1e-5)
        assert abs(float(info["diversity_loss"]) - (1 - 1 / m["K"])) < 1e-5
    if fx["n_exempt"] == 0:
        torch.testing.assert_close(aux.float(), fx["aux"].float(), **tol)
        for k in fx["info"]:
            torch.testing.assert_close(info[k].float(), fx["info"][k].float(), **tol)
        if fx["dgate_w"] is not None:
            torch.testing.assert_close(gw.grad, fx["dgate_w"], **tol)
        names = {"mlp": {"w1": 0, "b1": 1, "w2": 2, "b2": 3}, "glu": {"w1": 0, "w2": 1}}
        for e, (ew, ref) in enumerate(zip(exps, fx["dexperts"])):
            ref_list = list(ref.values())
            for key, pos in names[ew["kind"]].items():
                if ew[key].grad is None:
                    continue
                torch.testing.assert_close(ew[key].grad, ref_list[pos], **tol)


@pytest.mark.parametrize("name", PT)
def test_pretrain_oracle_matches_reference(name):
    fx = load_golden(name)
    m = fx["meta"]
    args = SimpleNamespace(**m["args"])
    x, wg, ks, vs = (_req(fx[n]) for n in ("x", "w_gate", "keys", "values"))
    bs, obs = (_req(fx["bias"]), _req(fx["o_bias"])) if "bias" in fx else (None, None)     # `-moe.bias 1` fixtures
    out, regs, dbg = op.competesmoe_forward(x, wg, ks, vs, m["K"], args, m["competition"], bias=bs, o_bias=obs)
    ((out * fx["dy"]).sum() + sum(regs.values())).backward()
    tol = dict(rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(out, fx["out"], **tol)
    assert set(regs) == set(fx["regs"])
    for k in regs:
        torch.testing.assert_close(regs[k], fx["regs"][k], **tol)
    for got, ref in ((x.grad, "dx"), (wg.grad, "dw_gate"), (ks.grad, "dkeys"), (vs.grad, "dvalues")):
        torch.testing.assert_close(got, fx[ref], **tol)
    if bs is not None:
        torch.testing.assert_close(bs.grad, fx["dbias"], **tol)
        torch.testing.assert_close(obs.grad, fx["do_bias"], **tol)


def test_cvmm_restatement_matches_triton_interpreter_run():
    fx = load_golden("cvmm_triton_interp")
    x, ks, vs, w = (_req(fx[n]) for n in ("x", "keys", "values", "w"))
    out = op.compute_moe_main(x, fx["sel"], w, ks, vs, F.relu, torch.float32)
    (out * fx["dy"]).sum().backward()
    tol = dict(rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(out, fx["out"], **tol)
    for got, ref in ((x.grad, "dx"), (ks.grad, "dkeys"), (vs.grad, "dvalues"), (w.grad, "dw")):
        torch.testing.assert_close(got, fx[ref], **tol)


def test_cvmm_is_the_einsum_identity():
    """cvmm == einsum('td,tkdh->tkh', x, keys[sel])."""
    torch.manual_seed(0)
    T, K, D, H, E = 40, 2, 16, 8, 5
    x = torch.randn(T, D)
    keys = torch.randn(E, D, H)
    sel = torch.stack([torch.randperm(E)[:K] for _ in range(T)]).int()
    s = op.prepare_sel2(sel)
    got = op.cvmm(x, s, keys)
    ref = torch.einsum("td,tkdh->tkh", x, keys[sel.long()])
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5)


def test_permutation_maps_are_a_stable_bijection():
    torch.manual_seed(1)
    sel = torch.randint(0, 7, (100, 3)).int()
    s = op.prepare_sel2(sel)
    flat = sel.flatten()
    assert sorted(s.out_index.tolist()) == list(range(flat.numel()))
    assert torch.equal(flat[s.out_index], s.sel.flatten())
    same = s.sel.flatten()[1:] == s.sel.flatten()[:-1]
    assert bool((s.out_index[1:][same] > s.out_index[:-1][same]).all())  # stable


def test_schedule_respects_cap_and_is_deterministic():
    draws = torch.rand(200, generator=torch.Generator().manual_seed(3))
    prior = [torch.ones(200, dtype=torch.bool) for _ in range(2)] + [torch.zeros(200, dtype=torch.bool)]
    prior[2][::2] = True                      # even steps already have 3 competing layers
    a = om.build_flip_schedule(200, 0.3, 3, prior, draws)
    b = om.build_flip_schedule(200, 0.3, 3, prior, draws)
    assert torch.equal(a, b)
    assert not bool(a[::2].any())             # full steps never get a 4th competing layer
    assert int(a.sum()) <= int((draws < 0.3).sum())


SIB = ["sib_smoe_f32", "sib_smoe_bf16", "sib_sigmoid_f32", "sib_sigmoid_bf16", "sib_xmoe_f32", "sib_perturbed_f32",
       "sib_share_f32", "sib_deepseekv3_f32", "sib_deepseekv3_nograd_f32"]


@pytest.mark.parametrize("name", SIB)
def test_sibling_router_oracle_matches_reference(name):
    """oracle/siblings.py against the outputs of the unmodified reference classes (smoe, smoe_sigmoidgating, xmoe,
    smoe_perturbed, smoe_share, deepseekv3)."""
    from oracle import siblings as osb
    fx = load_golden(name)
    m = fx["meta"]
    args = SimpleNamespace(**m["args"])
    x = fx["x"].clone().requires_grad_(m["requires_grad"])
    gate = {k: _req(v) for k, v in fx["gate"].items()}
    exps = [{k: (_req(v) if torch.is_tensor(v) else v) for k, v in e.items()} for e in fx["experts"]]
    out, aux, _, info, dbg = osb.sibling_forward(m["moe_name"], x, gate, exps, m["K"], m["d_out"], args)
    if m["requires_grad"]:
        ((out.float() * fx["dy"].float()).sum() + aux.float()).backward()
    tol = dict(rtol=1e-5, atol=1e-6) if "float32" in m["dtype"] else dict(rtol=2e-2, atol=2e-2)
    agree = (fx["selected"] == dbg["selected"]).all(-1)
    assert int((~agree).sum()) == fx["n_exempt"]
    torch.testing.assert_close(out[agree], fx["out"][agree], **tol)
    torch.testing.assert_close(dbg["weights"][agree].float(), fx["weights"][agree].float(), **tol)
    assert set(info) == set(fx["info"])
    for k, v in gate.items():                                  # in-place rescaling of the expert embeddings
        torch.testing.assert_close(v.detach(), fx["gate_after"][k], **tol)
    if m["requires_grad"]:
        torch.testing.assert_close(x.grad[agree], fx["dx"][agree], **tol)
    if fx["n_exempt"] == 0:
        torch.testing.assert_close(aux.float(), fx["aux"].float(), **tol)
        for k in fx["info"]:
            torch.testing.assert_close(info[k].float(), fx["info"][k].float(), **tol)
        for k, v in gate.items():
            if fx["dgate"][k] is not None:
                torch.testing.assert_close(v.grad, fx["dgate"][k], **tol)


PTSIB = ["ptsib_smoe_f32", "ptsib_sigmoid_f32", "ptsib_xmoe_f32", "ptsib_perturbed_f32", "ptsib_deepseekv2_f32",
         "ptsib_deepseekv3_f32"]


@pytest.mark.parametrize("name", PTSIB)
def test_pretrain_sibling_oracle_matches_reference(name):
    """oracle/pretrain_siblings.py against the unmodified reference classes of moe_pretrain_model/layers/moe (smoe,
    smoe_sigmoid, xmoe, smoe_perturbed, deepseekv2, deepseekv3)."""
    from oracle import pretrain_siblings as ops_
    fx = load_golden(name)
    m = fx["meta"]
    args = SimpleNamespace(**m["args"])
    x = fx["x"].clone().requires_grad_(True)
    params = {k: _req(v) for k, v in fx["params"].items()}
    out, regs, dbg = ops_.sibling_forward(m["moe_name"], x, params, m["K"], args)
    ((out * fx["dy"]).sum() + sum(regs.values())).backward()
    tol = dict(rtol=1e-4, atol=1e-6)
    assert torch.equal(dbg["selected"], fx["selected"])
    torch.testing.assert_close(out, fx["out"], **tol)
    assert set(regs) == set(fx["regs"])
    for k in regs:
        torch.testing.assert_close(regs[k], fx["regs"][k], **tol)
    torch.testing.assert_close(x.grad, fx["dx"], **tol)
    for k, g in fx["grads"].items():
        if g is not None:
            torch.testing.assert_close(params[k].grad, g, **tol)
    if "expert_embeddings_after" in fx:
        torch.testing.assert_close(params["expert_embeddings"].detach(), fx["expert_embeddings_after"], **tol)


def test_cvmm_oracle_moe_attention_layouts_match_einsum():
    """Known-answer test of oracle cvmm for the selection layouts of MoE attention
    (full_moe_relative_attention.py:453-458): per-head selections with re-divided sel_index / flattened reduction weight."""
    torch.manual_seed(7)
    T, heads, k, E, D, dh = 20, 4, 2, 6, 16, 8
    sel = torch.stack([torch.stack([torch.randperm(E)[:k] for _ in range(heads)]) for _ in range(T)]).int()
    w = torch.rand(T, heads, k)
    rs = op.prepare_sel2(sel)
    x, wq = torch.randn(T, D), torch.randn(E, D, dh)
    a = op.cvmm(x, op.Sel(rs.raw_sel, rs.sel, rs.out_index // (heads * k), rs.out_index, None), wq)
    torch.testing.assert_close(a, torch.einsum("td,thkdn->thkn", x, wq[sel.long()]), rtol=1e-5, atol=1e-5)
    xo, wo = torch.randn(T, heads, dh), torch.randn(E, dh, D)
    b = op.cvmm(xo, op.Sel(rs.raw_sel, rs.sel, rs.out_index // k, rs.out_index, w.flatten(-2)), wo)
    torch.testing.assert_close(b, torch.einsum("thk,thd,thkdn->tn", w, xo, wo[sel.long()]), rtol=1e-5, atol=1e-5)


def test_graph_mode_layers_deepcopy_cleanly():
    """enable_cuda_graphs() installs an instance-level forward; a deep copy (EMA / checkpoint averaging copies whole
    modules) must bind the copy's forward to the copy and must not carry captured graphs over.  CPU only: nothing is
    captured here, the wrapper falls through to the eager forward for non-CUDA inputs."""
    import copy
    import torch.nn as nn
    from types import SimpleNamespace
    import competesmoe_b200.pretrain_siblings  # noqa: F401
    from competesmoe_b200 import siblings  # noqa: F401
    from competesmoe_b200.multimodal import get_moe as get_mm
    from competesmoe_b200.pretrain import get_moe as get_pt
    pt = get_pt("smoe")(32, 4, 8, n_heads=2, args=SimpleNamespace(balance_loss_coef=0.01, test_only=False),
                        activation=F.relu, selection_mode="gate", log_interval=None).enable_cuda_graphs()
    pt._graphs["fake"] = object()
    cp = copy.deepcopy(pt)
    assert cp.forward.__self__ is cp and cp._eager_forward.__self__ is cp and cp._graphs == {}
    assert pt._graphs != {} and pt.forward.__self__ is pt

    def expert():
        m = nn.Module()
        m.fc1, m.fc2, m.activation_fn = nn.Linear(16, 24), nn.Linear(24, 16), nn.GELU(approximate="tanh")
        return m
    mm = get_mm("smoe")(16, 16, 4, 2, nn.ModuleList([expert() for _ in range(4)]), om.default_args()).enable_cuda_graphs()
    mm._graphs["fake"] = object()
    cm = copy.deepcopy(mm)
    assert cm.forward.__self__ is cm and cm._graphs == {} and mm._graphs != {}
    cm.enable_cuda_graphs(False)
    assert "forward" not in cm.__dict__ or cm.forward.__func__ is type(cm).forward


def test_forced_selection_reproduces_own_selection():
    """`forced_selected` (used by the full-size GPU parity tests to compare values under identical routing) with the
    oracle's own decision is the unforced evaluation, bit for bit, in both oracles and both branches."""
    import torch.nn.functional as F
    from oracle import multimodal as om
    from oracle import pretrain as op
    g = torch.Generator().manual_seed(5)
    D, Fh, E, K = 32, 48, 6, 2
    exps = [{"kind": "mlp", "act": "gelu", "w1": torch.randn(Fh, D, generator=g) * 0.2, "b1": torch.randn(Fh, generator=g) * 0.1,
             "w2": torch.randn(D, Fh, generator=g) * 0.2, "b2": torch.randn(D, generator=g) * 0.1} for _ in range(E)]
    gate_w = torch.randn(E, D, generator=g) * 0.3
    x = torch.randn(2, 24, D, generator=g)
    for comp in (False, True):
        a = om.competesmoe_forward(x, gate_w, exps, K, D, om.default_args(), comp)
        b = om.competesmoe_forward(x, gate_w, exps, K, D, om.default_args(), comp, forced_selected=a[4]["selected"])
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[4]["own_selected"], a[4]["selected"])
    wg, ks, vs = torch.randn(E, D, generator=g) * 0.3, torch.randn(E, D, 16, generator=g) * 0.2, torch.randn(E, 16, D, generator=g) * 0.2
    for comp in (False, True):
        a = op.competesmoe_forward(x, wg, ks, vs, K, op.default_args(), comp)
        b = op.competesmoe_forward(x, wg, ks, vs, K, op.default_args(), comp, forced_selected=a[2]["selected"])
        assert torch.equal(a[0], b[0]) and torch.equal(a[2]["own_selected"], a[2]["selected"])
        assert all(torch.equal(a[1][k], b[1][k]) for k in a[1])


def test_moe_attention_projection_oracle_matches_reference_golden():
    """oracle/pretrain_siblings.att_projection against the unmodified reference's smoe_perturbed layer built with
    is_att=True (att_forward + compute_moe; full_moe_relative_attention.py:267-296 builds it this way)."""
    from oracle import pretrain_siblings as ops_
    fx = load_golden("ptatt_perturbed_f32")
    m = fx["meta"]
    x = fx["x"].clone().requires_grad_(True)
    p = {k: v.clone().requires_grad_(True) for k, v in fx["params"].items()}
    out, dbg = ops_.att_projection(x, p, m["E"], m["heads"], m["K"])
    (out * fx["dy"]).sum().backward()
    assert torch.equal(dbg["selected"].sort(-1).values, fx["selected"].sort(-1).values)
    tol = dict(rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(out, fx["out"], **tol)
    torch.testing.assert_close(x.grad, fx["dx"], **tol)
    for k, g in fx["grads"].items():
        if g is None:
            assert p[k].grad is None or float(p[k].grad.abs().max()) == 0.0
        else:
            torch.testing.assert_close(p[k].grad, g, **tol)
