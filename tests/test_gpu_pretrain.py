"""GPU parity of the pretrain drop-in layer and of the `cvmm` op against the golden vectors of the unmodified reference
(fp32 fixtures; this path runs under bf16 autocast like the reference's training recipe, hence the 4e-2 tolerance) and
against the CPU oracle evaluated with op_dtype=bf16 (2e-2)."""
from types import SimpleNamespace

import pytest
import torch
import torch.nn.functional as F

from oracle import multimodal as om
from oracle import pretrain as op

from conftest import load_golden
from helpers import assert_close_rms

pytestmark = pytest.mark.gpu
DEV = "cuda"
PT = ["pt_router_f32", "pt_comp_f32", "pt_comp_hybrid_bal_f32", "pt_comp_intopk_f32", "pt_comp_tribrid_f32",
      "pt_router_cosine_f32", "pt_router_normweight_f32", "pt_router_normsigmoid_f32", "pt_comp_cosine_f32",
      "pt_router_bias_f32", "pt_comp_bias_f32"]


def build_layer(fx):
    from competesmoe_b200.pretrain import CompeteSMoE
    m = fx["meta"]
    args = SimpleNamespace(**m["args"])
    layer = CompeteSMoE(m["D"], m["E"], m["H"], n_heads=m["K"], args=args, activation=F.relu, selection_mode="gate",
                        log_interval=None, bias="bias" in fx)
    with torch.no_grad():
        layer.w_gate.copy_(fx["w_gate"]); layer.keys.copy_(fx["keys"]); layer.values.copy_(fx["values"])
        if "bias" in fx:     # `-moe.bias 1` fixtures (moe.py:129-134)
            layer.bias.copy_(fx["bias"]); layer.o_bias.copy_(fx["o_bias"])
    layer = layer.to(DEV)
    layer.train()
    layer.regularization_present = True
    layer.step_warm = 0
    layer.prob_flips_final = {0: torch.full((8,), bool(m["competition"]), device=DEV)}
    layer.set_current_steps(1)
    return layer, args


@pytest.mark.parametrize("name", PT)
def test_pretrain_layer_matches_reference_golden(name, grad_outliers=0.0):
    fx = load_golden(name)
    m = fx["meta"]
    layer, args = build_layer(fx)
    has_bias = "bias" in fx
    assert set(layer.state_dict().keys()) == {"w_gate", "keys", "values"} | ({"bias", "o_bias"} if has_bias else set())
    x = fx["x"].to(DEV).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = layer(x, id_layer=0)
        regs = layer.get_reg_loss()
    # `res + self.o_bias` (competesmoe.py:613-614) promotes the bf16 result to the fp32 parameter's dtype, here as there
    assert out.dtype == (torch.float32 if has_bias else torch.bfloat16) and out.shape == fx["out"].shape
    assert set(regs) == set(fx["regs"])
    ((out.float() * fx["dy"].to(DEV)).sum() + sum(regs.values())).backward()
    # The oracle (pinned to the reference by these same fp32 fixtures, tests/test_oracle_golden.py) evaluated with the
    # mixed precision this path uses: the apples-to-apples expectation for gradients.  The router gradient is ill
    # conditioned on tokens whose two selected experts have nearly equal dw (differences of O(1) numbers), where fp32
    # and bf16 legitimately differ by tens of percent -- the CPU bf16 oracle shows the same deviation from the fixture.
    xr, wg, ks, vs = (fx[n].clone().requires_grad_(True) for n in ("x", "w_gate", "keys", "values"))
    bs, obs = ((fx[n].clone().requires_grad_(True) for n in ("bias", "o_bias")) if has_bias else (None, None))
    o_out, o_regs, dbg = op.competesmoe_forward(xr, wg, ks, vs, m["K"], args, m["competition"], op_dtype=torch.bfloat16,
                                                bias=bs, o_bias=obs)
    ((o_out.float() * fx["dy"]).sum() + sum(o_regs.values())).backward()
    sel, w = layer.last_routing
    margin = om.topk_margin(dbg["affinity"] if m["competition"] else dbg["gate_softmax"], m["K"])
    agree_ref = (sel.cpu().long() == fx["selected"]).all(-1)
    agree = (sel.cpu().long() == dbg["selected"]).all(-1)
    n_ex = int((~agree).sum())
    assert bool((margin[~agree] < 1e-3).all()), "routing differs from the bf16 oracle on a token with margin >= 1e-3"
    # vs the fp32 reference run: the bf16 resolution of the scores (2^-8 relative) adds to the 1e-3 margin
    assert bool((margin[~agree_ref] < 1e-3 + 4e-3).all()), "routing differs from the reference on a clear-margin token"
    print(f"{name}: {n_ex}/{agree.numel()} low-margin tokens exempt (vs reference fp32 run: {int((~agree_ref).sum())})")
    both = agree & agree_ref
    assert_close_rms(out[both.to(DEV)], fx["out"][both], 4e-2, "output vs reference (fp32)")
    assert_close_rms(out[agree.to(DEV)], o_out.detach()[agree], 2e-2, "output vs oracle (bf16)")
    if n_ex == 0:
        for k in regs:
            got, ref = float(regs[k].detach()), float(o_regs[k].detach())
            assert abs(got - ref) <= 3e-2 * abs(ref) + 2e-5, (k, got, ref)
            assert abs(got - float(fx["regs"][k])) <= 6e-2 * abs(float(fx["regs"][k])) + 5e-5, (k, got)
        assert_close_rms(x.grad, xr.grad, 3e-2, "dx")
        assert_close_rms(layer.keys.grad, ks.grad, 3e-2, "dkeys", outliers=grad_outliers)
        assert_close_rms(layer.values.grad, vs.grad, 3e-2, "dvalues", outliers=grad_outliers)
        assert_close_rms(layer.w_gate.grad, wg.grad, 3e-2, "dw_gate", outliers=grad_outliers)
        if has_bias:
            assert_close_rms(layer.bias.grad, bs.grad, 3e-2, "dbias")
            assert_close_rms(layer.o_bias.grad, obs.grad, 3e-2, "do_bias")
    assert layer.keys.grad.dtype == torch.float32        # fp32 master parameters keep fp32 gradients


PT_WIDE = ["pt_router_e128_f32", "pt_comp_e128_f32"]     # the reference's default -moe.n_experts 128 (transformer_lm_mixin.py:32)


@pytest.mark.first_hw_run
@pytest.mark.parametrize("name", PT_WIDE)
def test_pretrain_layer_with_128_experts_matches_reference_golden(name):
    """More experts than the two-per-lane router / loss kernels hold: the four-per-lane variants.  Written after the GPU
    budget of the round was spent (green on the SIMT emulator, tests/test_simt_layers.py).
    320 (token, expert) pairs over 128 experts: a weight-gradient element is the sum of two or three products, so one
    bf16 rounding of d(scores) (the oracle rounds op by op, this path once) can decide it -- seen on the emulator: 1 of
    65 536 elements of dkeys at 2.3x the 3e-2 band; allowed: 0.1 % of the elements, each within 5x the band."""
    test_pretrain_layer_matches_reference_golden(name, grad_outliers=1e-3)


def test_cvmm_op_both_call_patterns():
    """compute_moe_main's two cvmm calls (competesmoe.py:510-522) through the public op, against the golden run of the
    reference's own Triton kernels on the interpreter."""
    from competesmoe_b200.cvmm import cvmm, cvmm_prepare_sel2
    fx = load_golden("cvmm_triton_interp")
    x = fx["x"].to(DEV).requires_grad_(True)
    keys = fx["keys"].to(DEV).requires_grad_(True)
    values = fx["values"].to(DEV).requires_grad_(True)
    w = fx["w"].to(DEV).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        s = cvmm_prepare_sel2(fx["sel"].to(DEV), n_experts=keys.shape[0])
        scores = cvmm(x, s, keys)
        assert scores.shape == fx["scores"].shape
        s2 = s.clone()
        s2.reduction_weight = w
        s2.sel_index = s2.out_index
        s2.out_index = None
        out = cvmm(F.relu(scores), s2, values)
    assert_close_rms(scores, fx["scores"], 4e-2, "scores")
    assert_close_rms(out, fx["out"], 4e-2, "out")
    (out.float() * fx["dy"].to(DEV)).sum().backward()
    assert_close_rms(x.grad, fx["dx"], 6e-2, "dx")
    assert_close_rms(keys.grad, fx["dkeys"], 6e-2, "dkeys")
    assert_close_rms(values.grad, fx["dvalues"], 6e-2, "dvalues")
    assert_close_rms(w.grad, fx["dw"], 6e-2, "dw")
    # index maps: bit-exact against the (stable) reference maps
    ref = op.prepare_sel2(fx["sel"])
    assert torch.equal(s.sel.cpu(), ref.sel) and torch.equal(s.sel_index.cpu(), ref.sel_index)
    assert torch.equal(s.out_index.cpu(), ref.out_index)


@pytest.mark.parametrize("competition", [False, True])
def test_c1_shape_against_oracle(competition):
    """BASELINE configs[0] shape: d_model=512, 8 experts, top-2, expert size 128, batch 8 x seq 512 (bf16 autocast)."""
    from competesmoe_b200.pretrain import CompeteSMoE
    torch.manual_seed(0)
    args = op.default_args(stop_after=8)
    layer = CompeteSMoE(512, 8, 128, n_heads=2, args=args, activation=F.relu, selection_mode="gate", log_interval=None)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(8, 512, 512, generator=g)
    dy = torch.randn(8, 512, 512, generator=g)
    xr = x.clone().requires_grad_(True)
    wg, ks, vs = (p.detach().clone().requires_grad_(True) for p in (layer.w_gate, layer.keys, layer.values))
    o_out, o_regs, dbg = op.competesmoe_forward(xr, wg, ks, vs, 2, args, competition, op_dtype=torch.bfloat16)
    ((o_out.float() * dy).sum() + sum(o_regs.values())).backward()
    layer = layer.to(DEV)
    layer.train()
    layer.regularization_present = True
    layer.step_warm = 0
    layer.prob_flips_final = {0: torch.full((8,), competition, device=DEV)}
    layer.set_current_steps(0)
    xg = x.to(DEV).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = layer(xg, id_layer=0)
        regs = layer.get_reg_loss()
    ((out.float() * dy.to(DEV)).sum() + sum(regs.values())).backward()
    sel, _ = layer.last_routing
    margin = om.topk_margin(dbg["affinity"] if competition else dbg["gate_softmax"], 2)
    agree = (sel.cpu().long() == dbg["selected"]).all(-1)
    assert bool((margin[~agree] < 1e-3).all()), "routing differs on a token with margin >= 1e-3"
    print(f"C1 competition={competition}: {int((~agree).sum())}/{agree.numel()} low-margin tokens exempt")
    assert_close_rms(out[agree.to(DEV)], o_out.detach()[agree], 2e-2, "output")
    assert set(regs) == set(o_regs)
    if int((~agree).sum()) <= agree.numel() // 200:
        assert_close_rms(layer.keys.grad, ks.grad, 5e-2, "dkeys")
        assert_close_rms(layer.values.grad, vs.grad, 5e-2, "dvalues")


def test_cvmm_moe_attention_call_patterns():
    """The two extra selection layouts SwitchHead-style MoE attention feeds the same op
    (layers/transformer/full_moe_relative_attention.py:453-458): per-head selections [T, heads, k] with
    (a) one input row per token shared by heads x k slots (q/k/v projections) and (b) one input row per (token, head) with
    the reduction weight flattened over (heads, k) (output projection).  Checked against the CPU oracle's cvmm."""
    from competesmoe_b200.cvmm import cvmm, cvmm_prepare_sel2
    torch.manual_seed(7)
    T, heads, k, E, D, dh = 80, 4, 2, 6, 64, 32
    sel = torch.stack([torch.stack([torch.randperm(E)[:k] for _ in range(heads)]) for _ in range(T)]).int()
    w = torch.rand(T, heads, k)
    ref_sel = op.prepare_sel2(sel)
    # (a) q-style: x [T, D] -> [T, heads, k, dh] without reduction; every slot of a token reads the token's row
    x = torch.randn(T, D)
    wq = torch.randn(E, D, dh) / D ** 0.5
    sa = op.Sel(ref_sel.raw_sel, ref_sel.sel, ref_sel.out_index // (heads * k), ref_sel.out_index, None)
    ref_a = op.cvmm(x, sa, wq)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        s = cvmm_prepare_sel2(sel.to(DEV), n_experts=E)
        s.sel_index = s.out_index // (heads * k)
        got_a = cvmm(x.to(DEV), s, wq.to(DEV))
    assert got_a.shape == ref_a.shape == (T, heads, k, dh)
    assert_close_rms(got_a, ref_a, 3e-2, "q-style projection")
    # (b) o-style: x [T, heads, dh] -> [T, D], weighted over heads x k
    xo = torch.randn(T, heads, dh).requires_grad_(True)
    wo = (torch.randn(E, dh, D) / dh ** 0.5).requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    sb = op.Sel(ref_sel.raw_sel, ref_sel.sel, ref_sel.out_index // k, ref_sel.out_index, wr.flatten(-2))
    ref_b = op.cvmm(xo, sb, wo)
    dy = torch.randn(T, D)
    (ref_b * dy).sum().backward()
    xg = xo.detach().to(DEV).requires_grad_(True)
    wg = wo.detach().to(DEV).requires_grad_(True)
    wrg = w.to(DEV).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o_sel = cvmm_prepare_sel2(sel.to(DEV), wrg, n_experts=E).clone()
        o_sel.sel_index = o_sel.out_index // o_sel.reduction_weight.shape[-1]
        o_sel.reduction_weight = o_sel.reduction_weight.flatten(-2)
        got_b = cvmm(xg, o_sel, wg)
    assert got_b.shape == ref_b.shape == (T, D)
    assert_close_rms(got_b, ref_b.detach(), 3e-2, "o-style projection")
    (got_b.float() * dy.to(DEV)).sum().backward()
    assert_close_rms(xg.grad, xo.grad, 4e-2, "dx")
    assert_close_rms(wg.grad, wo.grad, 4e-2, "dW")
    assert_close_rms(wrg.grad, wr.grad, 4e-2, "d reduction_weight")


@pytest.mark.parametrize("name", ["pt_router_f32", "pt_comp_f32", "pt_comp_hybrid_bal_f32"])
def test_pretrain_layer_cuda_graph_mode_matches_eager(name):
    """layer.enable_cuda_graphs(): the unchanged `layer(x, id_layer=0)` call under autocast, replayed from captured
    graphs, gives the eager call's output, regularisers, routing and gradients bit for bit -- on the capture inputs and
    on fresh ones."""
    fx = load_golden(name)
    eager, _ = build_layer(fx)
    graphed, _ = build_layer(fx)
    graphed.enable_cuda_graphs()
    g = torch.Generator().manual_seed(11)
    for trial in range(3):
        x_cpu = fx["x"] if trial == 0 else torch.randn(fx["x"].shape, generator=g)
        dy = (fx["dy"] if trial == 0 else torch.randn(fx["dy"].shape, generator=g)).to(DEV)
        res = []
        for layer in (eager, graphed):
            for p in layer.parameters():
                p.grad = None
            x = x_cpu.to(DEV).requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = layer(x, id_layer=0)
                regs = layer.get_reg_loss()
            ((out.float() * dy).sum() + sum(regs.values())).backward()
            res.append((out.clone(), {k: v.detach().clone() for k, v in regs.items()}, x.grad.clone(),
                        layer.last_routing[0].clone(), {n: p.grad.clone() for n, p in layer.named_parameters()}))
        (o0, r0, dx0, s0, g0), (o1, r1, dx1, s1, g1) = res
        assert torch.equal(o0, o1) and torch.equal(dx0, dx1) and torch.equal(s0, s1)
        assert set(r0) == set(r1) and all(torch.equal(r0[k], r1[k]) for k in r0), (r0, r1)
        assert set(g0) == set(g1) and all(torch.equal(g0[k], g1[k]) for k in g0)
    assert len(graphed._graphs) == 1 and graphed.layer == eager.layer
    graphed.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        assert graphed(fx["x"].to(DEV), id_layer=0).shape == fx["out"].shape


def test_cvmm_triton_library_op_honours_arbitrary_index_tensors():
    """torch.ops.mylib.cvmm_triton (cvmm.py:348-416): out[out_index[i]] = x[sel_index[i]] @ keys[sel[i]] for sorted
    expert ids `sel`, with index tensors that are NOT the stable-sort maps (a shuffled gather and scatter), and with the
    out_index = tensor(-1) marker.  Reference = the same formula in plain torch."""
    import competesmoe_b200.cvmm as C
    g = torch.Generator().manual_seed(3)
    M, R, E, D, N = 300, 120, 6, 64, 96
    sel = torch.sort(torch.randint(0, E, (M,), generator=g)).values.int()
    sel_index = torch.randint(0, R, (M,), generator=g)
    out_index = torch.randperm(M, generator=g)
    x = torch.randn(R, D, generator=g)
    keys = torch.randn(E, D, N, generator=g) / D ** 0.5
    want_sorted = torch.einsum("md,mdn->mn", x.bfloat16().float()[sel_index], keys.bfloat16().float()[sel.long()])
    want = torch.empty(M, N).index_copy_(0, out_index, want_sorted)
    op = torch.ops.mylib.cvmm_triton if C.cvmm_triton_call is torch.ops.mylib.cvmm_triton else C.cvmm_triton_call
    got = op(x.to(DEV), sel_index.to(DEV), sel.to(DEV), keys.to(DEV), torch.float32, out_index.to(DEV))
    assert got.shape == (M, N) and got.dtype == torch.float32
    assert_close_rms(got, want, 1e-2, "cvmm_triton with shuffled indices")
    got2 = op(x.to(DEV), sel_index.to(DEV), sel.to(DEV).view(M // 2, 2), keys.to(DEV), torch.bfloat16, torch.tensor(-1, device=DEV))
    assert got2.shape == (M // 2, 2, N) and got2.dtype == torch.bfloat16
    assert_close_rms(got2.view(M, N), want_sorted, 2e-2, "cvmm_triton, out_index = -1")


@pytest.mark.first_hw_run
def test_cvmm_triton_library_op_is_fp32_accurate_for_fp32_operands(direct=False):
    """out_dtype = fp32 with fp32 operands: the reference's kernel runs tl.dot(..., allow_tf32=False) on fp32 values
    (cvmm.py:389-395), so the result is the fp32 product, not a bf16 one -- rtol 1e-4 against plain fp32 torch.
    direct: call the Python implementation instead of the dispatcher op (the CPU tier has no CUDA dispatch key)."""
    import competesmoe_b200.cvmm as C
    g = torch.Generator().manual_seed(5)
    M, R, E, D, N = 260, 90, 5, 72, 40
    sel = torch.sort(torch.randint(0, E, (M,), generator=g)).values.int()
    sel_index = torch.randint(0, R, (M,), generator=g)
    out_index = torch.randperm(M, generator=g)
    x = torch.randn(R, D, generator=g)
    keys = torch.randn(E, D, N, generator=g) / D ** 0.5
    want = torch.empty(M, N).index_copy_(0, out_index, torch.einsum("md,mdn->mn", x[sel_index], keys[sel.long()]))
    op = C.cvmm_triton if direct else (torch.ops.mylib.cvmm_triton if C.cvmm_triton_call is torch.ops.mylib.cvmm_triton
                                       else C.cvmm_triton_call)
    got = op(x.to(DEV), sel_index.to(DEV), sel.to(DEV), keys.to(DEV), torch.float32, out_index.to(DEV))
    assert got.shape == (M, N) and got.dtype == torch.float32
    assert_close_rms(got, want, 1e-4, "cvmm_triton fp32")


def test_cvmm_rejects_index_tensors_it_cannot_express():
    """cvmm() uses the route built from raw_sel; a CVMMSel whose sel_index was edited to anything but the recognised
    layouts must raise instead of silently computing other rows."""
    from competesmoe_b200.cvmm import cvmm, cvmm_prepare_sel2
    g = torch.Generator().manual_seed(4)
    T, K, E, D = 64, 2, 4, 32
    sel = torch.stack([torch.randperm(E, generator=g)[:K] for _ in range(T)]).int().to(DEV)
    x = torch.randn(T, D, generator=g).to(DEV)
    keys = torch.randn(E, D, 16, generator=g).to(DEV)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        s = cvmm_prepare_sel2(sel, n_experts=E)
        cvmm(x, s, keys)                                   # the maps it built itself are fine
        s.sel_index = torch.flip(s.sel_index, dims=[0])    # not `pos // K` any more
        with pytest.raises(NotImplementedError):
            cvmm(x, s, keys)


def test_moe_attention_projection_layer_is_att():
    """`get_moe(name)(..., is_att=True, inp_expert, out_expert)` -- the expert projections FullMoeRopeAttention builds
    (full_moe_relative_attention.py:267-296) and drives through att_forward / compute_moe (:351-389, moe.py:456-489):
    per-head top-k over sigmoid gates, one [in, out] matrix per (head, expert), weighted sum over the k selections."""
    from competesmoe_b200.pretrain import CompeteSMoE
    torch.manual_seed(0)
    D, heads, E, k, dh, B, N = 128, 4, 5, 2, 32, 2, 96
    layer = CompeteSMoE(D, E * heads, 1, n_heads=heads, topk=k, args=op.default_args(), is_att=True, inp_expert=D,
                        out_expert=dh, std_gate=D ** -0.5, std_expert=D ** -0.5).to(DEV).train()
    assert set(layer.state_dict()) == {"w_gate", "experts"} and layer.experts.shape == (E * heads, D, dh)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, N, D, generator=g).to(DEV).requires_grad_(True)
    dy = torch.randn(B, N, heads, dh, generator=g).to(DEV)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        sel = layer.att_forward(x, n_copies=heads, n_experts=E)
        out = layer.compute_moe(x, sel)
    assert out.shape == (B, N, heads, dh) and sel.raw_sel_index.shape == (B, N, heads, k)
    (out.float() * dy).sum().backward()
    # plain-torch restatement
    xr = x.detach().clone().requires_grad_(True)
    wg = layer.w_gate.detach().clone().requires_grad_(True)
    ex = layer.experts.detach().clone().requires_grad_(True)
    logits = F.linear(xr.bfloat16(), wg.bfloat16()).view(B, N, heads, E)
    idx = sel.raw_sel_index
    val = torch.gather(logits, -1, idx).sigmoid()
    flat = (torch.arange(heads, device=DEV).view(1, 1, heads, 1) * E + idx)
    w_sel = ex.bfloat16().float()[flat]              # [B, N, heads, k, D, dh]; fp32 after the rounding: fp32 grad accumulation
    proj = torch.einsum("bnd,bnhkde->bnhke", xr.bfloat16().float(), w_sel)
    ref = (val.float().unsqueeze(-1) * proj.bfloat16().float()).sum(-2)
    (ref * dy).sum().backward()
    assert_close_rms(out, ref.detach(), 2e-2, "is_att projection")
    assert_close_rms(x.grad, xr.grad, 3e-2, "dx")
    assert_close_rms(layer.experts.grad, ex.grad, 3e-2, "d experts")
    assert_close_rms(layer.w_gate.grad, wg.grad, 4e-2, "d w_gate")


@pytest.mark.first_hw_run
@pytest.mark.parametrize("autocast,variant", [(True, {}), (True, {"norm_sigmoid": True, "scale_weight": 2.0}), (False, {}),
                                              (True, {"is_cosine": True})])
def test_policy_level_methods_match_the_oracle(autocast, variant):
    """The reference's policy-level methods under their own names and signatures -- compute_gate(x), topk_expert(logits),
    router_policy(x), compute_scores(x, sel), compute_moe_main(x, selected, weights), competition_policy_mlp_faster(x),
    zloss, balanceloss (moe.py:273-322,373-416; competesmoe.py:381-414,456-490,510-522) -- against the oracle, values and
    gradients; autocast=False: the fp32-accurate path at rtol 1e-4.  Written after the round's GPU budget was spent."""
    from competesmoe_b200.cvmm import cvmm_prepare_sel2
    from competesmoe_b200.pretrain import CompeteSMoE
    D, E, H, K, B, N = 64, 8, 32, 2, 2, 37
    args = op.default_args(**variant)
    torch.manual_seed(5)
    layer = CompeteSMoE(D, E, H, n_heads=K, args=args, activation=F.relu, selection_mode="gate", log_interval=None)
    g = torch.Generator().manual_seed(6)
    with torch.no_grad():
        layer.w_gate.copy_(torch.randn(E, D, generator=g) * 0.3)
    w_gate, keys, values = (p.detach().clone() for p in (layer.w_gate, layer.keys, layer.values))
    layer = layer.to(DEV).train()
    x = torch.randn(B, N, D, generator=g)
    dy = torch.randn(B, N, D, generator=g)
    odt = torch.bfloat16 if autocast else torch.float32
    rt = 2e-2 if autocast else 1e-4
    fresh = lambda: x.detach().clone()
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        # ---- gate, top-k, router policy
        xr, wg = fresh().requires_grad_(True), w_gate.clone().requires_grad_(True)
        ow, osel, oprobs, ologits = op.router_policy(xr, wg, K, args, odt)
        xg = fresh().to(DEV).requires_grad_(True)
        logits = layer.compute_gate(xg)
        assert logits.shape == (B, N, E)
        assert_close_rms(logits, ologits.detach(), rt, "compute_gate")
        w, sel, probs, logits2 = layer.router_policy(xg, is_normal_mode=True)
        assert sel.dtype == torch.int64 and sel.shape == (B, N, K) and probs.shape == (B, N, E)
        scores = ologits if variant.get("norm_sigmoid") else oprobs
        agree = (sel.cpu() == osel).all(-1)
        assert bool((om.topk_margin(scores.float(), K)[~agree] < 1e-3).all())
        assert_close_rms(probs, oprobs.detach(), rt, "gate softmax")
        assert_close_rms(w.cpu()[agree], ow.detach()[agree], rt, "routing weights")
        tw, tsel, tprobs = layer.topk_expert(logits.detach())
        ref_p = F.softmax(logits.detach().cpu(), dim=-1, dtype=torch.float32)
        rv, ri = om.stable_topk(ref_p, K)
        assert torch.equal(tsel.cpu(), ri)
        torch.testing.assert_close(tw.cpu(), rv, rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(layer.zloss(logits.detach().float()).cpu(),
                                   torch.square(torch.logsumexp(logits.detach().float().cpu(), -1)).mean(), rtol=1e-5, atol=1e-7)
        top1 = F.one_hot(ri[..., 0], E).float().mean(-2)
        torch.testing.assert_close(layer.balanceloss(tsel, tprobs).cpu(), (ref_p.mean(-2) * top1).mean() * E * E, rtol=1e-5, atol=1e-7)
        # ---- compute_scores / compute_moe_main under the oracle's routing
        ks, vs = keys.clone().requires_grad_(True), values.clone().requires_grad_(True)
        o_out = op.compute_moe_main(xr, osel, ow.detach(), ks, vs, F.relu, odt)
        (o_out.float() * dy).sum().backward()
        out = layer.compute_moe_main(xg, osel.to(DEV), ow.detach().to(DEV))
        assert out.shape == (B, N, D)
        assert_close_rms(out, o_out.detach(), rt, "compute_moe_main")
        (out.float() * dy.to(DEV)).sum().backward()
        assert_close_rms(xg.grad, xr.grad, 3e-2 if autocast else rt, "compute_moe_main dx")
        assert_close_rms(layer.keys.grad, ks.grad, 3e-2 if autocast else rt, "dkeys", outliers=1e-3 if autocast else 0.0)
        assert_close_rms(layer.values.grad, vs.grad, 3e-2 if autocast else rt, "dvalues", outliers=1e-3 if autocast else 0.0)
        sel_pp = cvmm_prepare_sel2(osel.to(DEV).int(), n_experts=E)
        sc = layer.compute_scores(fresh().to(DEV), sel_pp)
        o_sc = F.relu(op.cvmm(fresh(), op.prepare_sel2(osel.int()), keys, odt))
        assert sc.shape == (B, N, K, H)
        assert_close_rms(sc, o_sc, rt, "compute_scores")
        # ---- competition policy
        xr2 = fresh().requires_grad_(True)
        cw, csel, csoft, caff, ctop = op.competition_policy(xr2, keys, values, K, F.relu, odt)
        xg2 = fresh().to(DEV).requires_grad_(True)
        w2, sel2, soft2, aff2, top2 = layer.competition_policy_mlp_faster(xg2)
        assert sel2.dtype == torch.int64 and top2.shape == (B, N, K, D) and aff2.shape == (B, N, E)
        agree2 = (sel2.cpu() == csel).all(-1)
        assert bool((om.topk_margin(caff, K)[~agree2] < 1e-3).all())
        assert_close_rms(aff2, caff.detach(), rt, "affinity")
        assert_close_rms(soft2, csoft.detach(), rt, "softmax(affinity)")
        assert_close_rms(w2.cpu()[agree2], cw.detach()[agree2], rt, "competition weights")
        assert_close_rms(top2.cpu()[agree2], ctop.detach()[agree2], rt, "selected outputs")
        if bool(agree2.all()):
            ((ctop.float() * dy.unsqueeze(2)).sum() + (cw.float() * 3).sum()).backward()
            ((top2.float() * dy.to(DEV).unsqueeze(2)).sum() + (w2.float() * 3).sum()).backward()
            assert_close_rms(xg2.grad, xr2.grad, 3e-2 if autocast else rt, "competition_policy dx")
