"""Generate tests/golden/*.pt by running the UNMODIFIED reference modules (imported from /root/reference).

Run in the build container only:   python -m oracle.gen_golden
The GPU box has no /root/reference; tests read only the committed fixtures.  Each fixture stores seeded inputs, weights,
the forced schedule branch, and the reference's outputs / losses / gradients.  While generating, the oracle restatement
is replayed on the same data and must agree (this is the oracle's pin; tests/test_oracle_golden.py repeats it).

Reference entry points exercised:
  moe_model/model/moe/register.py:18 get_moe("competesmoe") -> competesmoe.py:9 CompeteSMoE (forward :337)
  moe_pretrain_model/layers/moe/competesmoe.py:38 CompeteSMoE (forward :524), through the import shim of SURVEY.md B.1
  moe_pretrain_model/layers/cvmm.py:460 CVMM autograd function run on the Triton interpreter
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import types
from pathlib import Path
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


# ------------------------------------------------------------------------------------------------ multimodal
def mm_args(**kw):
    base = dict(rate_flip=0.05, warm_up=0.0, max_compete_in_iter=3, hybrid=False, router_theta=1.0,
                router_loss_coef=0.01, diversity_loss_coef=0.01, bal_comp_loss_coef=0.01, balance_loss_coef=0.01,
                router_z_loss_coef=0.001, norm_sigmoid=False, init_weight=True, moe_name="competesmoe")
    base.update(kw)
    return SimpleNamespace(**base)


class _TinyGLU(nn.Module):
    """Same arithmetic and parameter names as transformers' Phi3MLP (gate_up_proj / down_proj, SiLU-GLU)."""

    def __init__(self, d, f):
        super().__init__()
        self.gate_up_proj = nn.Linear(d, 2 * f, bias=False)
        self.down_proj = nn.Linear(f, d, bias=False)
        self.activation_fn = nn.SiLU()

    def forward(self, x):
        up = self.gate_up_proj(x)
        gate, up = up.chunk(2, dim=-1)
        return self.down_proj(up * self.activation_fn(gate))


def build_expert(kind, d_in, d_out, f, ref_mods):
    if kind == "siglip":
        cfg = SimpleNamespace(hidden_act="gelu_pytorch_tanh", hidden_size=d_in, intermediate_size=f)
        return ref_mods["siglip"].SiglipMLP(cfg)
    if kind == "projector":
        return nn.Sequential(nn.Linear(d_in, d_out), nn.GELU(), nn.Linear(d_out, d_out))
    if kind == "glu":
        try:
            from transformers.models.phi3.modeling_phi3 import Phi3MLP
            from transformers import Phi3Config
            return Phi3MLP(Phi3Config(hidden_size=d_in, intermediate_size=f, hidden_act="silu"))
        except Exception:
            return _TinyGLU(d_in, f)
    raise ValueError(kind)


def expert_weights(kind, mod):
    if kind == "siglip":
        return {"kind": "mlp", "act": "gelu_tanh", "w1": mod.fc1.weight, "b1": mod.fc1.bias, "w2": mod.fc2.weight,
                "b2": mod.fc2.bias}
    if kind == "projector":
        return {"kind": "mlp", "act": "gelu", "w1": mod[0].weight, "b1": mod[0].bias, "w2": mod[2].weight,
                "b2": mod[2].bias}
    return {"kind": "glu", "act": "silu", "w1": mod.gate_up_proj.weight, "w2": mod.down_proj.weight}


def gen_multimodal(ref_mods, name, kind, d_in, d_out, f, E, K, B, N, competition, dtype=torch.float32, upcycled=False,
                   seed=0, **argkw):
    from oracle import multimodal as om

    torch.manual_seed(seed)
    args = mm_args(**argkw)
    with quiet():
        if upcycled:
            expert = build_expert(kind, d_in, d_out, f, ref_mods)
        else:
            expert = nn.ModuleList([build_expert(kind, d_in, d_out, f, ref_mods) for _ in range(E)])
        layer = ref_mods["get_moe"]("competesmoe")(in_embed_dim=d_in, out_embed_dim=d_out, num_of_experts=E,
                                                  num_selected=K, expert=expert, args=args)
        layer = layer.to(dtype)
        layer.set_total_steps(4, id_layer=0, prob_flips_final={})
    layer.prob_flips = torch.full((4,), bool(competition))
    layer.set_current_steps(1)
    g = torch.Generator().manual_seed(1234 + seed)
    x = torch.randn(B, N, d_in, generator=g).to(dtype).requires_grad_(True)
    dy = torch.randn(B, N, d_out, generator=g).to(dtype)
    captured = {}
    orig_compute_moe = layer.compute_moe

    def spy(*a, **kw):  # record the routing decision the reference takes (moe.py:172 arguments)
        captured["selected"] = kw["selected_experts"].detach().clone()
        captured["weights"] = kw["weights"].detach().clone()
        return orig_compute_moe(*a, **kw)

    layer.compute_moe = spy
    out, aux, _, info = layer(x)
    loss = (out.float() * dy.float()).sum() + aux.float()
    loss.backward()
    fx = {
        "meta": dict(name=name, kind=kind, d_in=d_in, d_out=d_out, f=f, E=E, K=K, B=B, N=N, competition=competition,
                     dtype=str(dtype), upcycled=upcycled, args=vars(args)),
        "x": x.detach().clone(), "dy": dy, "gate_w": layer.gate.weight.detach().clone(),
        "experts": [{k: (v.detach().clone() if torch.is_tensor(v) else v)
                     for k, v in expert_weights(kind, m).items()} for m in layer.experts],
        "out": out.detach().clone(), "aux": aux.detach().clone(), "info": {k: v.clone() for k, v in info.items()},
        "dx": x.grad.clone(), "dgate_w": None if layer.gate.weight.grad is None else layer.gate.weight.grad.clone(),
        "dexperts": [{n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None} for m in layer.experts],
    }
    # ---- pin the oracle against the reference on this very case
    x2 = fx["x"].clone().requires_grad_(True)
    gw = fx["gate_w"].clone().requires_grad_(True)
    exps = [{k: (v.clone().requires_grad_(True) if torch.is_tensor(v) else v) for k, v in e.items()}
            for e in fx["experts"]]
    o_out, o_aux, _, o_info, dbg = om.competesmoe_forward(x2, gw, exps, K, d_out, args, competition)
    ((o_out.float() * dy.float()).sum() + o_aux.float()).backward()
    tol = dict(rtol=1e-5, atol=1e-6) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    # Routing: bit-exact except tokens whose k-th/(k+1)-th score margin is < 1e-3 (torch.topk's tie order is
    # unspecified; bf16 affinity scores tie often).  Those tokens are counted and excluded from the value checks.
    scores = dbg["affinity"] if competition else dbg["gate_softmax"]
    margin = om.topk_margin(scores, K)
    agree = (captured["selected"] == dbg["selected"]).all(-1)
    assert bool((margin[~agree] < 1e-3).all()), "routing differs on a token with margin >= 1e-3"
    n_exempt = int((~agree).sum())
    if not upcycled:  # with identical experts every score ties and the reference's order is arbitrary
        torch.testing.assert_close(o_out[agree], fx["out"][agree], **tol)
        torch.testing.assert_close(x2.grad[agree], fx["dx"][agree], **tol)
        torch.testing.assert_close(captured["weights"][agree].float(), dbg["weights"][agree].float(), **tol)
    if n_exempt == 0:
        torch.testing.assert_close(o_aux.float(), fx["aux"].float(), **tol)
        for k_ in info:
            torch.testing.assert_close(o_info[k_].float(), info[k_].float(), **tol)
    fx["selected"] = captured["selected"]          # the reference's own decision
    fx["weights"] = captured["weights"]
    fx["selected_oracle"] = dbg["selected"].clone()
    fx["n_exempt"] = n_exempt
    torch.save(fx, OUT / f"{name}.pt")
    print(f"  wrote {name}.pt  (oracle == reference, {n_exempt} low-margin tokens exempt)  aux={float(aux.detach()):.6f} "
          f"info={ {k: round(float(v), 6) for k, v in info.items()} }")


def sibling_gate_params(name, layer):
    if name in ("xmoe", "smoe_perturbed"):
        return {"inp_reduction_w": layer.inp_reduction.weight, "expert_embeddings": layer.expert_embeddings}
    return {"gate_w": layer.gate.weight}


def gen_sibling(ref_mods, fixture, moe_name, kind, d_in, d_out, f, E, K, B, N, dtype=torch.float32, seed=0,
                requires_grad=True):
    """Fixtures for the sibling routers (smoe.py, smoe_sigmoidgating.py, xmoe.py, smoe_perturbed.py, shard_smoe.py,
    deepseekv3.py): same recipe as gen_multimodal, the oracle is oracle/siblings.py."""
    from oracle import multimodal as om
    from oracle import siblings as osb

    torch.manual_seed(seed)
    args = mm_args(moe_name=moe_name)
    with quiet():
        if moe_name in ("smoe_share", "deepseekv3"):      # these deep-copy ONE expert module (shard_smoe.py:33)
            expert = build_expert(kind, d_in, d_out, f, ref_mods)
        else:
            expert = nn.ModuleList([build_expert(kind, d_in, d_out, f, ref_mods) for _ in range(E)])
        layer = ref_mods["get_moe"](moe_name)(in_embed_dim=d_in, out_embed_dim=d_out, num_of_experts=E, num_selected=K,
                                              expert=expert, args=args)
        if moe_name in ("smoe_share", "deepseekv3"):      # de-correlate the copies so that routing is not all ties
            g0 = torch.Generator().manual_seed(100 + seed)
            with torch.no_grad():
                for m in layer.experts:
                    for p_ in m.parameters():
                        p_.add_(torch.randn(p_.shape, generator=g0) * 0.05)
        layer = layer.to(dtype)
    g = torch.Generator().manual_seed(4321 + seed)
    x = torch.randn(B, N, d_in, generator=g).to(dtype).requires_grad_(requires_grad)
    dy = torch.randn(B, N, d_out, generator=g).to(dtype)
    gate0 = {k: v.detach().clone() for k, v in sibling_gate_params(moe_name, layer).items()}   # before the in-place renorm
    captured = {}
    orig = layer.compute_moe

    def spy(selected_experts, weights, results, x, *a, **kw):
        captured["selected"] = selected_experts.detach().clone()
        captured["weights"] = weights.detach().clone()
        return orig(selected_experts, weights, results, x, *a, **kw)

    layer.compute_moe = spy
    with quiet():
        res = layer(x)
    out, aux, info = res[0], res[1], res[3]
    if requires_grad:
        ((out.float() * dy.float()).sum() + aux.float()).backward()
    gate_after = {k: v.detach().clone() for k, v in sibling_gate_params(moe_name, layer).items()}
    gate_grads = {k: (None if v.grad is None else v.grad.clone()) for k, v in sibling_gate_params(moe_name, layer).items()}
    fx = {
        "meta": dict(name=fixture, moe_name=moe_name, kind=kind, d_in=d_in, d_out=d_out, f=f, E=E, K=K, B=B, N=N,
                     dtype=str(dtype), args=vars(args), requires_grad=requires_grad),
        "x": x.detach().clone(), "dy": dy, "gate": gate0, "gate_after": gate_after, "dgate": gate_grads,
        "experts": [{k: (v.detach().clone() if torch.is_tensor(v) else v)
                     for k, v in expert_weights(kind, m).items()} for m in layer.experts],
        "out": out.detach().clone(), "aux": aux.detach().clone(), "info": {k: v.clone() for k, v in info.items()},
        "dx": x.grad.clone() if requires_grad else None,
        "dexperts": [{n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None} for m in layer.experts],
        "selected": captured["selected"], "weights": captured["weights"],
    }
    # ---- pin the oracle
    x2 = fx["x"].clone().requires_grad_(requires_grad)
    gate = {k: v.clone().requires_grad_(True) for k, v in gate0.items()}
    exps = [{k: (v.clone().requires_grad_(True) if torch.is_tensor(v) else v) for k, v in e.items()} for e in fx["experts"]]
    o_out, o_aux, _, o_info, dbg = osb.sibling_forward(moe_name, x2, gate, exps, K, d_out, args)
    if requires_grad:
        ((o_out.float() * dy.float()).sum() + o_aux.float()).backward()
    tol = dict(rtol=1e-5, atol=1e-6) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    k_eff = K - 1 if moe_name in ("smoe_share", "deepseekv3") else K
    scores = torch.sigmoid(dbg["gate_logits"]) if moe_name == "smoe_sigmoidgating" else dbg["gate_softmax"]
    margin = om.topk_margin(scores, k_eff)
    agree = (captured["selected"] == dbg["selected"]).all(-1)
    assert bool((margin[~agree] < 1e-3).all()), "routing differs on a token with margin >= 1e-3"
    n_exempt = int((~agree).sum())
    torch.testing.assert_close(o_out[agree], fx["out"][agree], **tol)
    torch.testing.assert_close(captured["weights"][agree].float(), dbg["weights"][agree].float(), **tol)
    if requires_grad:
        torch.testing.assert_close(x2.grad[agree], fx["dx"][agree], **tol)
    if n_exempt == 0:
        torch.testing.assert_close(o_aux.float(), fx["aux"].float(), **tol)
        for k_ in info:
            torch.testing.assert_close(o_info[k_].float(), info[k_].float(), **tol)
        for k_, v in gate.items():
            if fx["dgate"][k_] is not None:
                torch.testing.assert_close(v.grad, fx["dgate"][k_], **tol)
    for k_, v in gate.items():      # the in-place rescaling of expert_embeddings is part of the contract
        torch.testing.assert_close(v.detach(), gate_after[k_], **tol)
    fx["selected_oracle"] = dbg["selected"].clone()
    fx["n_exempt"] = n_exempt
    torch.save(fx, OUT / f"{fixture}.pt")
    print(f"  wrote {fixture}.pt  ({moe_name}: oracle == reference, {n_exempt} low-margin tokens exempt)  "
          f"aux={float(aux.detach()):.6f}")


def load_multimodal_reference():
    sys.path.insert(0, str(REF))
    with quiet():
        importlib.import_module("moe_model.model.moe")
        reg = importlib.import_module("moe_model.model.moe.register")
        siglip = importlib.import_module("moe_model.model.multimodal_encoder.siglip_smoe")
        try:   # deepseekv3.py is not imported by the package __init__ (so not registered); it needs `loguru`
            importlib.import_module("loguru")
        except ImportError:
            sys.modules["loguru"] = types.ModuleType("loguru")
        importlib.import_module("moe_model.model.moe.deepseekv3")
    return {"get_moe": reg.get_moe, "siglip": siglip}


# ------------------------------------------------------------------------------------------------ pretrain
def load_pretrain_reference():
    R = str(REF / "moe_pretrain_model")
    sys.path.insert(0, R)

    def stub(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
        return m

    L = stub("layers", R + "/layers")
    LM = stub("layers.moe", R + "/layers/moe")
    FW = stub("framework", R + "/framework")
    with quiet():
        cv = importlib.import_module("layers.cvmm")
        L.cvmm, L.cvmm_prepare_sel = cv.cvmm, cv.cvmm_prepare_sel
        FW.utils = importlib.import_module("framework.utils")
        FW.layers = importlib.import_module("framework.layers")
        base = importlib.import_module("layers.moe.moe")
        LM.MoE = base.MoE
        reg = importlib.import_module("layers.moe.register")
        comp = importlib.import_module("layers.moe.competesmoe")
    return {"cvmm": cv, "base": base, "get_moe": reg.get_moe, "comp": comp}


def pt_args(**kw):
    base = dict(warm_up=0.0, rate_flip=0.07, stop_after=8, max_compete_in_iter=3, is_cosine=False, is_norm_weight=False,
                norm_sigmoid=False, scale_weight=1.0, hybrid=False, tribrid=False, in_topk=False,
                balance_affinity=False, balance_loss_coef=0.01, balance_loss_coef_comp=0.01, router_loss_coef=0.01,
                router_theta=1.0, test_only=False)
    base.update(kw)
    return SimpleNamespace(**base)


def gen_pretrain(pm, name, D, E, H, K, B, N, competition, seed=0, bias=False, **argkw):
    """Reference layer with its `cvmm` name bound to the oracle's per-expert restatement (the Triton op needs a GPU);
    the restatement itself is pinned by gen_cvmm_interpreter below."""
    from oracle import pretrain as op

    def cvmm_standin(x, sel, keys):
        if not isinstance(sel, pm["cvmm"].CVMMSel):
            sel = pm["cvmm"].cvmm_prepare_sel(sel, keys.shape[0])
        s = op.Sel(sel.raw_sel, sel.sel, sel.sel_index, sel.out_index, sel.reduction_weight)
        return op.cvmm(x, s, keys, torch.float32)

    pm["base"].cvmm = cvmm_standin
    pm["comp"].cvmm = cvmm_standin
    args = pt_args(**argkw)
    torch.manual_seed(seed)
    cwd = os.getcwd()
    os.chdir("/tmp")  # set_total_steps appends to ./file_path.txt (competesmoe.py:218-221)
    try:
        with quiet():
            layer = pm["get_moe"]("competesmoe")(D, E, H, n_heads=K, args=args, activation=F.relu, selection_mode="gate",
                                                 log_interval=None, bias=bias)
            if bias:    # `-moe.bias 1` (moe.py:129-134): both biases start at zero in the reference; make them matter
                with torch.no_grad():
                    layer.bias.normal_(0, 0.3)
                    layer.o_bias.normal_(0, 0.3)
            layer.train()
            layer.regularization_present = True
            layer.set_total_steps(id_layer=0)
    finally:
        os.chdir(cwd)
    layer.prob_flips_final[0][:] = bool(competition)
    layer.set_current_steps(1)
    g = torch.Generator().manual_seed(1234 + seed)
    x = torch.randn(B, N, D, generator=g).requires_grad_(True)
    dy = torch.randn(B, N, D, generator=g)
    out = layer(x, id_layer=0)
    regs = layer.get_reg_loss()
    loss = (out * dy).sum() + sum(regs.values())
    loss.backward()
    fx = {
        "meta": dict(name=name, D=D, E=E, H=H, K=K, B=B, N=N, competition=competition, args=vars(args)),
        "x": x.detach().clone(), "dy": dy, "w_gate": layer.w_gate.detach().clone(), "keys": layer.keys.detach().clone(),
        "values": layer.values.detach().clone(), "out": out.detach().clone(),
        "regs": {k: v.detach().clone() for k, v in regs.items()},
        "dx": x.grad.clone(), "dw_gate": layer.w_gate.grad.clone(), "dkeys": layer.keys.grad.clone(),
        "dvalues": layer.values.grad.clone(),
    }
    if bias:
        fx.update({"bias": layer.bias.detach().clone(), "o_bias": layer.o_bias.detach().clone(),
                   "dbias": layer.bias.grad.clone(), "do_bias": layer.o_bias.grad.clone()})
    x2 = fx["x"].clone().requires_grad_(True)
    wg, ks, vs = (fx[n].clone().requires_grad_(True) for n in ("w_gate", "keys", "values"))
    bs, obs = ((fx[n].clone().requires_grad_(True) for n in ("bias", "o_bias")) if bias else (None, None))
    o_out, o_regs, dbg = op.competesmoe_forward(x2, wg, ks, vs, K, args, competition, bias=bs, o_bias=obs)
    ((o_out * dy).sum() + sum(o_regs.values())).backward()
    tol = dict(rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(o_out, fx["out"], **tol)
    assert set(o_regs) == set(regs), (set(o_regs), set(regs))
    for k_ in regs:
        torch.testing.assert_close(o_regs[k_], fx["regs"][k_], **tol)
    torch.testing.assert_close(x2.grad, fx["dx"], **tol)
    torch.testing.assert_close(wg.grad, fx["dw_gate"], **tol)
    torch.testing.assert_close(ks.grad, fx["dkeys"], **tol)
    torch.testing.assert_close(vs.grad, fx["dvalues"], **tol)
    if bias:
        torch.testing.assert_close(bs.grad, fx["dbias"], **tol)
        torch.testing.assert_close(obs.grad, fx["do_bias"], **tol)
    fx["selected"] = dbg["selected"].clone()
    torch.save(fx, OUT / f"{name}.pt")
    print(f"  wrote {name}.pt  (oracle == reference)  regs={ {k: round(float(v), 6) for k, v in regs.items()} }")


def gen_cvmm_interpreter(name="cvmm_triton_interp"):
    """Run the reference's own CVMM autograd function with its Triton kernels on the Triton interpreter (fixed tiles,
    autotuner bypassed) and store inputs/outputs/gradients.  Must run in a fresh process with TRITON_INTERPRET=1."""
    from oracle import pretrain as op

    pm = load_pretrain_reference()
    cv = pm["cvmm"]
    import triton

    def fwd_call(x, sel_index, sel, keys, out_dtype, out_index):
        x = x.flatten(end_dim=-2)
        sel_shape = sel.shape
        sel = sel.flatten()
        M = sel.shape[0]
        O, K, N = keys.shape
        out = torch.empty((M, N), dtype=out_dtype)
        none = out_index.numel() == 1 and out_index == -1
        BM = BN = BK = 32
        grid = (triton.cdiv(M, BM) * triton.cdiv(N, BN),)
        cv.cvmm_kernel.fn[grid](x, keys, out, sel_index, sel, out_index, M, N, K, x.stride(0), x.stride(1),
                                keys.stride(0), keys.stride(1), keys.stride(2), out.stride(0), out.stride(1),
                                sel_index.stride(0), sel.stride(0), 0 if none else out_index.stride(0),
                                out_index_is_none=none, dtype_id=cv.dtype_to_type_id(out.dtype), allow_tf32=False,
                                BLOCK_SIZE_M=BM, BLOCK_SIZE_N=BN, BLOCK_SIZE_K=BK, GROUP_SIZE_M=8)
        return out.view(*sel_shape, N)

    def bwd_call(x, sel_index, sel, grads, n_experts, key_dtype, op_dtype, out_index):
        x = x.flatten(end_dim=-2).transpose(0, 1)
        grads = grads.flatten(end_dim=-2)
        sel = sel.flatten()
        M, _ = x.shape
        K, N = grads.shape
        out = torch.zeros((n_experts, M, N), dtype=key_dtype)
        none = out_index.numel() == 1 and out_index == -1
        BM = BN = 32
        BK, KB = 16, 4
        grid = (triton.cdiv(M, BM) * triton.cdiv(N, BN), triton.cdiv(K, BK * KB))
        cv.cvmm_backward_kernel3.fn[grid](x, grads, out, sel_index, sel, out_index, M, N, K, x.stride(0), x.stride(1),
                                          grads.stride(0), grads.stride(1), out.stride(0), out.stride(1), out.stride(2),
                                          sel_index.stride(0), sel.stride(0), 0 if none else out_index.stride(0),
                                          out_index_is_none=none, out_dtype_id=cv.dtype_to_type_id(out.dtype),
                                          dtype_id=cv.dtype_to_type_id(op_dtype), allow_tf32=False, BLOCK_SIZE_M=BM,
                                          BLOCK_SIZE_N=BN, BLOCK_SIZE_K=BK, GROUP_SIZE_M=8, K_BLOCKS=KB)
        return out

    cv.cvmm_triton_call = fwd_call
    cv.cvmm_triton_backward = bwd_call
    torch.manual_seed(0)
    B, N, D, H, E, K = 2, 12, 32, 32, 4, 2
    x = torch.randn(B, N, D, requires_grad=True)
    keys = (torch.randn(E, D, H) / D ** 0.5).requires_grad_(True)
    values = (torch.randn(E, H, D) / H ** 0.5).requires_grad_(True)
    selx = torch.stack([torch.randperm(E)[:K] for _ in range(B * N)]).view(B, N, K).int()
    w = torch.rand(B, N, K).requires_grad_(True)
    dy = torch.randn(B, N, D)
    # reference: both call patterns of compute_moe_main (competesmoe.py:510-522); stable sort for the index maps
    s = op.prepare_sel2(selx)
    rs = cv.CVMMSel(s.raw_sel, s.sel, s.sel_index, s.out_index, None)
    scores = cv.CVMM.apply(x, rs.sel_index, rs.sel, keys, rs.out_index, None)
    act = F.relu(scores)
    out = cv.CVMM.apply(act, rs.out_index, rs.sel, values, None, w)
    (out * dy).sum().backward()
    fx = {"x": x.detach().clone(), "keys": keys.detach().clone(), "values": values.detach().clone(), "sel": selx,
          "w": w.detach().clone(), "dy": dy, "scores": scores.detach().clone(), "out": out.detach().clone(),
          "dx": x.grad.clone(), "dkeys": keys.grad.clone(), "dvalues": values.grad.clone(), "dw": w.grad.clone()}
    # oracle replay
    x2, k2, v2, w2 = (fx[n].clone().requires_grad_(True) for n in ("x", "keys", "values", "w"))
    o = op.compute_moe_main(x2, selx, w2, k2, v2, F.relu, torch.float32)
    (o * dy).sum().backward()
    tol = dict(rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(o, fx["out"], **tol)
    torch.testing.assert_close(x2.grad, fx["dx"], **tol)
    torch.testing.assert_close(k2.grad, fx["dkeys"], **tol)
    torch.testing.assert_close(v2.grad, fx["dvalues"], **tol)
    torch.testing.assert_close(w2.grad, fx["dw"], **tol)
    torch.save(fx, OUT / f"{name}.pt")
    print(f"  wrote {name}.pt  (oracle cvmm == reference Triton kernels on the interpreter)")


def gen_all_siblings(mm):
    print("multimodal sibling routers:")
    gen_sibling(mm, "sib_smoe_f32", "smoe", "siglip", 64, 64, 128, 4, 2, 2, 24, seed=10)
    gen_sibling(mm, "sib_smoe_bf16", "smoe", "siglip", 64, 64, 128, 4, 2, 2, 24, dtype=torch.bfloat16, seed=11)
    gen_sibling(mm, "sib_sigmoid_f32", "smoe_sigmoidgating", "projector", 48, 64, 64, 4, 2, 2, 16, seed=12)
    gen_sibling(mm, "sib_sigmoid_bf16", "smoe_sigmoidgating", "siglip", 64, 64, 128, 4, 2, 2, 24, dtype=torch.bfloat16, seed=13)
    gen_sibling(mm, "sib_xmoe_f32", "xmoe", "siglip", 64, 64, 128, 8, 2, 2, 24, seed=14)
    gen_sibling(mm, "sib_perturbed_f32", "smoe_perturbed", "glu", 64, 64, 96, 8, 2, 2, 16, seed=15)
    gen_sibling(mm, "sib_share_f32", "smoe_share", "siglip", 64, 64, 128, 5, 3, 2, 24, seed=16)
    gen_sibling(mm, "sib_deepseekv3_f32", "deepseekv3", "siglip", 64, 64, 128, 5, 3, 2, 24, seed=17)
    gen_sibling(mm, "sib_deepseekv3_nograd_f32", "deepseekv3", "siglip", 64, 64, 128, 5, 3, 2, 16, seed=18, requires_grad=False)


PT_SIBLING_MODULES = {"smoe": "smoe", "smoe_sigmoid": "smoeut_norm", "xmoe": "xmoe", "smoe_perturbed": "smoe_perturbed",
                      "deepseekv2": "deepseekv2", "deepseekv3": "deepseekv3"}
PT_SIBLING_EXTRA = {"xmoe": ("expert_embeddings", "expert_sel"), "smoe_perturbed": ("expert_embeddings", "expert_sel"),
                    "deepseekv2": ("keys_shared", "values_shared"),
                    "deepseekv3": ("keys_shared", "values_shared", "e_score_correction_bias")}


def gen_pretrain_sibling(pm, fixture, moe_name, D, E, H, K, B, N, seed=0):
    """Unmodified reference sibling layer (smoe.py, smoeut_norm.py, xmoe.py, smoe_perturbed.py, deepseekv2.py,
    deepseekv3.py) with its `cvmm` name bound to the oracle's per-expert restatement, as in gen_pretrain."""
    from oracle import pretrain as op
    from oracle import pretrain_siblings as ops_

    def cvmm_standin(x, sel, keys):
        if not isinstance(sel, pm["cvmm"].CVMMSel):
            sel = pm["cvmm"].cvmm_prepare_sel(sel, keys.shape[0])
        s = op.Sel(sel.raw_sel, sel.sel, sel.sel_index, sel.out_index, sel.reduction_weight)
        return op.cvmm(x, s, keys, torch.float32)

    with quiet():
        mod = importlib.import_module("layers.moe." + PT_SIBLING_MODULES[moe_name])
    pm["base"].cvmm = cvmm_standin
    mod.cvmm = cvmm_standin
    args = pt_args()
    torch.manual_seed(seed)
    with quiet():
        layer = pm["get_moe"](moe_name)(D, E, H, n_heads=K, args=args, activation=F.relu, selection_mode="gate",
                                        log_interval=None)
    layer.train()
    layer.regularization_present = True
    names = ("w_gate", "keys", "values") + PT_SIBLING_EXTRA.get(moe_name, ())
    before = {n: getattr(layer, n).detach().clone() for n in names}          # xmoe rescales expert_embeddings in forward
    g = torch.Generator().manual_seed(1234 + seed)
    x = torch.randn(B, N, D, generator=g).requires_grad_(True)
    dy = torch.randn(B, N, D, generator=g)
    out = layer(x)
    regs = layer.get_reg_loss()
    loss = (out * dy).sum() + sum(regs.values())
    loss.backward()
    fx = {"meta": dict(name=fixture, moe_name=moe_name, D=D, E=E, H=H, K=K, B=B, N=N, args=vars(args)),
          "x": x.detach().clone(), "dy": dy, "params": before, "out": out.detach().clone(),
          "regs": {k: v.detach().clone() for k, v in regs.items()}, "dx": x.grad.clone(),
          "grads": {n: (getattr(layer, n).grad.clone() if getattr(layer, n).grad is not None else None) for n in names}}
    if "expert_embeddings" in before:
        fx["expert_embeddings_after"] = layer.expert_embeddings.detach().clone()
    # oracle replay
    x2 = fx["x"].clone().requires_grad_(True)
    p2 = {n: v.clone().requires_grad_(True) for n, v in before.items()}
    o_out, o_regs, dbg = ops_.sibling_forward(moe_name, x2, p2, K, args)
    ((o_out * dy).sum() + sum(o_regs.values())).backward()
    tol = dict(rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(o_out, fx["out"], **tol)
    assert set(o_regs) == set(regs), (set(o_regs), set(regs))
    for k_ in regs:
        torch.testing.assert_close(o_regs[k_], fx["regs"][k_], **tol)
    torch.testing.assert_close(x2.grad, fx["dx"], **tol)
    for n in names:
        if fx["grads"][n] is None:
            assert p2[n].grad is None or float(p2[n].grad.abs().max()) == 0.0, n
        else:
            torch.testing.assert_close(p2[n].grad, fx["grads"][n], **tol)
    if "expert_embeddings" in before:
        torch.testing.assert_close(p2["expert_embeddings"].detach(), fx["expert_embeddings_after"], **tol)
    fx["selected"] = dbg["selected"].clone()
    torch.save(fx, OUT / f"{fixture}.pt")
    print(f"  wrote {fixture}.pt  (oracle == reference)  regs={ {k: round(float(v), 6) for k, v in regs.items()} }")


def gen_all_pretrain_siblings(pm):
    print("pretrain sibling routers:")
    gen_pretrain_sibling(pm, "ptsib_smoe_f32", "smoe", 64, 8, 32, 2, 2, 32, seed=20)
    gen_pretrain_sibling(pm, "ptsib_sigmoid_f32", "smoe_sigmoid", 64, 8, 32, 2, 2, 32, seed=21)
    gen_pretrain_sibling(pm, "ptsib_xmoe_f32", "xmoe", 64, 8, 32, 2, 2, 24, seed=22)
    gen_pretrain_sibling(pm, "ptsib_perturbed_f32", "smoe_perturbed", 64, 16, 16, 4, 2, 16, seed=23)
    gen_pretrain_sibling(pm, "ptsib_deepseekv2_f32", "deepseekv2", 64, 8, 32, 2, 2, 24, seed=24)
    gen_pretrain_sibling(pm, "ptsib_deepseekv3_f32", "deepseekv3", 64, 16, 16, 4, 2, 16, seed=25)


def gen_gate_variants(mm, pm):
    """Gate / score variants behind `args` flags (competesmoe.py:456-483 pretrain; competesmoe.py:243-247 multimodal)."""
    print("gate variants:")
    gen_pretrain(pm, "pt_router_cosine_f32", 64, 8, 32, 2, 2, 24, False, seed=30, is_cosine=True)
    gen_pretrain(pm, "pt_router_normweight_f32", 64, 8, 32, 2, 2, 24, False, seed=31, is_norm_weight=True)
    gen_pretrain(pm, "pt_router_normsigmoid_f32", 64, 8, 32, 2, 2, 24, False, seed=32, norm_sigmoid=True, scale_weight=2.0)
    gen_pretrain(pm, "pt_comp_cosine_f32", 64, 8, 32, 2, 2, 16, True, seed=33, is_cosine=True)
    gen_multimodal(mm, "mm_siglip_comp_normsigmoid_f32", "siglip", 64, 64, 128, 4, 2, 2, 16, True, seed=34, norm_sigmoid=True)


def gen_bias(pm):
    """`-moe.bias 1`: hidden bias [E, H] inside the selected experts, o_bias [D] on the layer output (moe.py:129-134,
    :400-401; competesmoe.py:613-614); the competition's dense scoring pass ignores the hidden bias (:381-414)."""
    gen_pretrain(pm, "pt_router_bias_f32", 64, 8, 32, 2, 2, 24, False, seed=40, bias=True)
    gen_pretrain(pm, "pt_comp_bias_f32", 64, 8, 32, 2, 2, 24, True, seed=41, bias=True)


def gen_att_projection(pm, fixture="ptatt_perturbed_f32", D=64, heads=4, E=6, k=2, dh=16, B=2, N=24, seed=60):
    """Expert projection of MoE attention: the one `att_forward` the reference ships live (smoe_perturbed.py:199-226; the
    base class's is commented out), on a layer built the way full_moe_relative_attention.py:267-296 builds it."""
    from oracle import pretrain as op
    from oracle import pretrain_siblings as ops_

    def cvmm_standin(x, sel, keys):
        if not isinstance(sel, pm["cvmm"].CVMMSel):
            sel = pm["cvmm"].cvmm_prepare_sel(sel, keys.shape[0])
        s = op.Sel(sel.raw_sel, sel.sel, sel.sel_index, sel.out_index, sel.reduction_weight)
        return op.cvmm(x, s, keys, torch.float32)

    with quiet():
        mod = importlib.import_module("layers.moe.smoe_perturbed")
    pm["base"].cvmm = cvmm_standin
    mod.cvmm = cvmm_standin
    args = pt_args()
    torch.manual_seed(seed)
    with quiet():
        layer = pm["get_moe"]("smoe_perturbed")(dmodel=D, n_experts=E * heads, expert_size=1, n_heads=heads, topk=k, args=args,
                                                is_att=True, inp_expert=D, out_expert=dh, std_gate=D ** -0.5,
                                                std_expert=D ** -0.5, out_dmodel=heads * dh, log_interval=None)
    with torch.no_grad():      # the reference leaves this parameter as torch.empty (smoe_perturbed.py:99-103)
        layer.expert_embeddings.normal_(0, 0.3)
    layer.train()
    names = ("expert_sel", "expert_embeddings", "experts", "w_gate")
    before = {n: getattr(layer, n).detach().clone() for n in names}
    g = torch.Generator().manual_seed(1234 + seed)
    x = torch.randn(B, N, D, generator=g).requires_grad_(True)
    dy = torch.randn(B, N, heads, dh, generator=g)
    sel = layer.att_forward(x, n_experts=E, n_copies=heads)
    out = layer.compute_moe(x, sel)
    (out * dy).sum().backward()
    fx = {"meta": dict(name=fixture, D=D, heads=heads, E=E, K=k, dh=dh, B=B, N=N, args=vars(args)),
          "x": x.detach().clone(), "dy": dy, "params": before, "out": out.detach().clone(),
          "selected": sel.raw_sel_index.clone(), "weights": sel.sel_val.detach().clone(), "raw_sel": sel.raw_sel.detach().clone(),
          "dx": x.grad.clone(), "expert_embeddings_after": layer.expert_embeddings.detach().clone(),
          "grads": {n: (getattr(layer, n).grad.clone() if getattr(layer, n).grad is not None else None) for n in names}}
    x2 = fx["x"].clone().requires_grad_(True)
    p2 = {n: v.clone().requires_grad_(True) for n, v in before.items()}
    o_out, dbg = ops_.att_projection(x2, p2, E, heads, k)
    (o_out * dy).sum().backward()
    tol = dict(rtol=1e-4, atol=1e-6)
    assert torch.equal(dbg["selected"].sort(-1).values, fx["selected"].sort(-1).values)
    torch.testing.assert_close(o_out, fx["out"], **tol)
    torch.testing.assert_close(x2.grad, fx["dx"], **tol)
    for n in names:
        if fx["grads"][n] is None:
            assert p2[n].grad is None or float(p2[n].grad.abs().max()) == 0.0, n
        else:
            torch.testing.assert_close(p2[n].grad, fx["grads"][n], **tol)
    torch.testing.assert_close(p2["expert_embeddings"].detach(), fx["expert_embeddings_after"], **tol)
    torch.save(fx, OUT / f"{fixture}.pt")
    print(f"  wrote {fixture}.pt  (oracle == reference)")


def gen_wide(pm):
    """The reference's own default expert count, `-moe.n_experts 128` (transformer_lm_mixin.py:32): more experts than the
    two-per-lane router / loss kernels hold, so the 4-per-lane variants run (router.cu / losses.cu "more than 64 experts")."""
    gen_pretrain(pm, "pt_router_e128_f32", 32, 128, 16, 4, 2, 40, False, seed=50)
    gen_pretrain(pm, "pt_comp_e128_f32", 32, 128, 16, 4, 2, 24, True, seed=51, hybrid=True, balance_affinity=True,
                 router_theta=0.5)


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    if "--att-only" in sys.argv:
        gen_att_projection(load_pretrain_reference())
        return
    if "--wide-only" in sys.argv:
        gen_wide(load_pretrain_reference())
        return
    if "--gate-variants-only" in sys.argv:
        gen_gate_variants(load_multimodal_reference(), load_pretrain_reference())
        return
    if "--siblings-only" in sys.argv:
        gen_all_siblings(load_multimodal_reference())
        return
    if "--bias-only" in sys.argv:
        gen_bias(load_pretrain_reference())
        return
    if "--pretrain-siblings-only" in sys.argv:
        gen_all_pretrain_siblings(load_pretrain_reference())
        return
    if os.environ.get("TRITON_INTERPRET") == "1":
        gen_cvmm_interpreter()
        return
    print("multimodal reference (moe_model/model/moe):")
    mm = load_multimodal_reference()
    for comp in (False, True):
        tag = "comp" if comp else "router"
        gen_multimodal(mm, f"mm_siglip_{tag}_f32", "siglip", 64, 64, 128, 4, 2, 2, 24, comp)
        gen_multimodal(mm, f"mm_projector_{tag}_f32", "projector", 48, 64, 64, 4, 2, 2, 16, comp, seed=1)
        gen_multimodal(mm, f"mm_glu_{tag}_f32", "glu", 64, 64, 96, 4, 2, 2, 16, comp, seed=2)
    gen_multimodal(mm, "mm_siglip_comp_hybrid_f32", "siglip", 64, 64, 128, 8, 2, 2, 16, True, seed=3, hybrid=True,
                   router_theta=0.5)
    gen_multimodal(mm, "mm_siglip_router_bf16", "siglip", 64, 64, 128, 4, 2, 2, 24, False, dtype=torch.bfloat16, seed=4)
    gen_multimodal(mm, "mm_siglip_comp_bf16", "siglip", 64, 64, 128, 4, 2, 2, 24, True, dtype=torch.bfloat16, seed=4)
    gen_multimodal(mm, "mm_siglip_comp_upcycled_f32", "siglip", 64, 64, 128, 4, 2, 2, 16, True, upcycled=True, seed=5)
    gen_all_siblings(mm)
    print("pretrain reference (moe_pretrain_model/layers/moe):")
    pm = load_pretrain_reference()
    gen_pretrain(pm, "pt_router_f32", 64, 8, 32, 2, 2, 32, False)
    gen_pretrain(pm, "pt_comp_f32", 64, 8, 32, 2, 2, 32, True, seed=1)
    gen_pretrain(pm, "pt_comp_hybrid_bal_f32", 64, 16, 16, 4, 2, 16, True, seed=2, hybrid=True, balance_affinity=True,
                 router_theta=0.5)
    gen_pretrain(pm, "pt_comp_intopk_f32", 64, 8, 32, 2, 2, 16, True, seed=3, in_topk=True)
    gen_pretrain(pm, "pt_comp_tribrid_f32", 64, 8, 32, 2, 2, 16, True, seed=4, tribrid=True)
    gen_all_pretrain_siblings(pm)
    gen_gate_variants(mm, pm)
    gen_bias(pm)
    gen_wide(pm)
    gen_att_projection(pm)
    print("done; now run:  TRITON_INTERPRET=1 python -m oracle.gen_golden")


if __name__ == "__main__":
    main()
