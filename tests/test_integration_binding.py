"""The reference-side binding of INTEGRATION.md, exercised against the real reference tree when it is present
(the build container).  On the GPU box /root/reference does not exist and these tests skip.  CPU only: construction,
registry lookup, isinstance and checkpoint-key parity -- no kernel runs here."""
import importlib
import io
import contextlib
import sys
import types
from pathlib import Path
from types import SimpleNamespace

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not REF.exists(), reason="reference tree not present")


def _mm_args():
    return SimpleNamespace(rate_flip=0.05, warm_up=0.0, max_compete_in_iter=3, hybrid=False, router_theta=1.0,
                           router_loss_coef=0.01, diversity_loss_coef=0.01, bal_comp_loss_coef=0.01,
                           balance_loss_coef=0.01, router_z_loss_coef=0.001, norm_sigmoid=False, init_weight=True,
                           moe_name="competesmoe_b200")


def test_multimodal_binding_registers_and_matches_checkpoint_layout():
    sys.path.insert(0, str(REF))
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import moe_model.model.moe  # noqa: F401
            reg = importlib.import_module("moe_model.model.moe.register")
            base = importlib.import_module("moe_model.model.moe.moe")
    finally:
        sys.path.remove(str(REF))
    from competesmoe_b200.integrate import bind_multimodal

    bound = bind_multimodal(reg, base, names=("competesmoe_b200",), overwrite=True)
    cls = reg.get_moe("competesmoe_b200")
    assert cls is bound and issubclass(cls, base.MoeLayer)

    def expert():
        m = nn.Module()
        m.fc1, m.fc2, m.activation_fn = nn.Linear(32, 48), nn.Linear(48, 32), nn.GELU(approximate="tanh")
        return m

    torch.manual_seed(0)
    ours = cls(in_embed_dim=32, out_embed_dim=32, num_of_experts=4, num_selected=2,
               expert=nn.ModuleList([expert() for _ in range(4)]), args=_mm_args())
    with contextlib.redirect_stdout(io.StringIO()):
        theirs = reg.get_moe("competesmoe")(in_embed_dim=32, out_embed_dim=32, num_of_experts=4, num_selected=2,
                                            expert=nn.ModuleList([expert() for _ in range(4)]), args=_mm_args())
    assert isinstance(ours, base.MoeLayer)
    sd_o, sd_t = ours.state_dict(), theirs.state_dict()
    assert sorted(sd_o) == sorted(sd_t)
    assert all(sd_o[k].shape == sd_t[k].shape and sd_o[k].dtype == sd_t[k].dtype for k in sd_o)
    # seeded gate init (moe.py:50-70): identical bits
    assert torch.equal(sd_o["gate.weight"], sd_t["gate.weight"])
    # a reference checkpoint loads into the drop-in unchanged
    ours.load_state_dict(sd_t, strict=True)
    # the schedule API the trainer calls (llava_trainer.py:1034-1079)
    flips = ours.set_total_steps(20, id_layer=0, prob_flips_final={})
    assert 0 in flips and flips[0].numel() == 20
    ours.set_current_steps(3)
    assert ours.current_steps == 3 and hasattr(ours, "total_steps")


_CVMM_MOD = None


def _pretrain_shim():
    R = str(REF / "moe_pretrain_model")
    sys.path.insert(0, R)

    def stub(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
        return m

    saved = {k: sys.modules.get(k) for k in ("layers", "layers.moe", "framework")}
    L, LM, Fw = stub("layers", R + "/layers"), stub("layers.moe", R + "/layers/moe"), stub("framework", R + "/framework")
    global _CVMM_MOD
    if _CVMM_MOD is None:          # layers/cvmm.py defines a torch.library op: it can be imported once per process
        _CVMM_MOD = importlib.import_module("layers.cvmm")
    else:
        sys.modules["layers.cvmm"] = _CVMM_MOD
    cv = _CVMM_MOD
    L.cvmm, L.cvmm_prepare_sel = cv.cvmm, cv.cvmm_prepare_sel
    Fw.utils = importlib.import_module("framework.utils")
    Fw.layers = importlib.import_module("framework.layers")
    base = importlib.import_module("layers.moe.moe")
    LM.MoE = base.MoE
    reg = importlib.import_module("layers.moe.register")
    importlib.import_module("layers.moe.competesmoe")
    return L, base, reg, saved, R


def test_pretrain_binding_registers_and_matches_checkpoint_layout(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)   # the reference's set_total_steps appends to ./file_path.txt
    try:
        L, base, reg, saved, R = _pretrain_shim()
    except Exception as e:  # pragma: no cover - the reference tree changed
        pytest.skip(f"reference import shim failed: {e}")
    try:
        from competesmoe_b200.integrate import bind_cvmm, bind_pretrain
        from competesmoe_b200 import cvmm as our_cvmm

        bound = bind_pretrain(reg, base, names=("competesmoe_b200",), overwrite=True)
        ns = SimpleNamespace(warm_up=0.0, rate_flip=0.07, stop_after=10, max_compete_in_iter=3, is_cosine=False,
                             is_norm_weight=False, norm_sigmoid=False, scale_weight=1.0, hybrid=False, tribrid=False,
                             in_topk=False, balance_affinity=False, balance_loss_coef=0.01,
                             balance_loss_coef_comp=0.01, router_loss_coef=0.01, router_theta=1.0, test_only=False)
        kw = dict(n_heads=2, args=ns, activation=F.relu, selection_mode="gate", log_interval=None, bias=True)
        ours = reg.get_moe("competesmoe_b200")(64, 8, 16, **kw)
        theirs = reg.get_moe("competesmoe")(64, 8, 16, **kw)
        assert bound is type(ours) and isinstance(ours, base.MoE)
        sd_o, sd_t = ours.state_dict(), theirs.state_dict()
        assert sorted(sd_o) == sorted(sd_t)
        assert all(sd_o[k].shape == sd_t[k].shape for k in sd_o)
        ours.load_state_dict(sd_t, strict=True)
        assert ours.num_selected == theirs.num_selected == 2
        # regulariser mixin contract (framework/layers/regularized_layer.py:9-62)
        ours.train()
        ours.regularization_present = True
        ours.add_reg(lambda: torch.tensor(2.0), "mlp_ebalance")
        ours.add_reg(lambda: torch.tensor(4.0), "mlp_ebalance")
        assert float(ours.get_reg_loss()["mlp_ebalance"]) == 3.0 and ours.get_reg_loss() == {}
        flips = ours.set_total_steps(id_layer=0)
        assert flips[0].numel() == 10
        bind_cvmm(L)
        assert L.cvmm is our_cvmm.cvmm and L.CVMMSel is our_cvmm.CVMMSel
    finally:
        sys.path.remove(R)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in [k for k in sys.modules if k.startswith(("layers.", "framework."))]:
            sys.modules.pop(k, None)


def test_sibling_routers_bind_and_match_reference_state_dicts():
    """competesmoe_b200.integrate.bind_multimodal_siblings: every sibling router is a subclass of the reference's
    MoeLayer, is found through the reference's get_moe, and has the reference class's state_dict keys and shapes."""
    sys.path.insert(0, str(REF))
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import moe_model.model.moe  # noqa: F401
            reg = importlib.import_module("moe_model.model.moe.register")
            base = importlib.import_module("moe_model.model.moe.moe")
    finally:
        sys.path.remove(str(REF))
    from competesmoe_b200.integrate import bind_multimodal_siblings
    bound = bind_multimodal_siblings(reg, base, suffix="_b200", overwrite=True)

    def expert():
        m = nn.Module()
        m.fc1, m.fc2, m.activation_fn = nn.Linear(32, 48), nn.Linear(48, 32), nn.GELU(approximate="tanh")
        return m

    for name in ("smoe", "smoe_sigmoidgating", "xmoe", "smoe_perturbed", "smoe_share"):   # deepseekv3 is not registered upstream
        cls = reg.get_moe(name + "_b200")
        assert cls is bound[name] and issubclass(cls, base.MoeLayer)
        single = name == "smoe_share"      # the reference deep-copies one module there (shard_smoe.py:33)
        E = 5 if single else 4
        mk = (lambda: expert()) if single else (lambda: nn.ModuleList([expert() for _ in range(E)]))
        torch.manual_seed(0)
        ours = cls(in_embed_dim=32, out_embed_dim=32, num_of_experts=E, num_selected=2, expert=mk(), args=_mm_args())
        with contextlib.redirect_stdout(io.StringIO()):
            theirs = reg.get_moe(name)(in_embed_dim=32, out_embed_dim=32, num_of_experts=E, num_selected=2, expert=mk(),
                                       args=_mm_args())
        sd_o, sd_t = ours.state_dict(), theirs.state_dict()
        assert sorted(sd_o) == sorted(sd_t), (name, sorted(set(sd_o) ^ set(sd_t)))
        assert all(sd_o[k].shape == sd_t[k].shape for k in sd_o), name
        gate_key = "expert_embeddings" if name in ("xmoe", "smoe_perturbed") else "gate.weight"
        assert torch.equal(sd_o[gate_key], sd_t[gate_key]), f"{name}: seeded gate init differs"
        ours.load_state_dict(sd_t)


def test_pretrain_sibling_routers_bind_and_match_reference_state_dicts(tmp_path, monkeypatch):
    """competesmoe_b200.integrate.bind_pretrain_siblings: every pretrain-plugin sibling is a subclass of the reference's
    MoE, is found through the reference's get_moe, and has the reference class's state_dict keys and shapes."""
    monkeypatch.chdir(tmp_path)
    try:
        L, base, reg, saved, R = _pretrain_shim()
        for mod in ("smoe", "smoeut_norm", "xmoe", "smoe_perturbed", "deepseekv2", "deepseekv3"):
            importlib.import_module("layers.moe." + mod)
    except Exception as e:  # pragma: no cover - the reference tree changed
        pytest.skip(f"reference import shim failed: {e}")
    try:
        from competesmoe_b200.integrate import bind_pretrain_siblings
        bound = bind_pretrain_siblings(reg, base, suffix="_b200", overwrite=True)
        ns = SimpleNamespace(balance_loss_coef=0.01, test_only=False)
        kw = dict(n_heads=2, args=ns, activation=F.relu, selection_mode="gate", log_interval=None)
        for name in ("smoe", "smoe_sigmoid", "xmoe", "smoe_perturbed", "deepseekv2", "deepseekv3"):
            cls = reg.get_moe(name + "_b200")
            assert cls is bound[name] and issubclass(cls, base.MoE)
            ours = cls(64, 8, 16, **kw)
            with contextlib.redirect_stdout(io.StringIO()):
                theirs = reg.get_moe(name)(64, 8, 16, **kw)
            sd_o, sd_t = ours.state_dict(), theirs.state_dict()
            assert sorted(sd_o) == sorted(sd_t), (name, sorted(set(sd_o) ^ set(sd_t)))
            assert all(sd_o[k].shape == sd_t[k].shape for k in sd_o), name
            ours.load_state_dict(sd_t, strict=True)
    finally:
        sys.path.remove(R)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in [k for k in sys.modules if k.startswith(("layers.", "framework."))]:
            sys.modules.pop(k, None)


def test_schedule_matches_reference_bit_for_bit_multimodal():
    """set_total_steps of the drop-in against the reference's own (competesmoe.py:35-179): same seed -> the same
    competition flags for a chain of layers, with the per-step cap (`max_compete_in_iter`) biting so that the shift-left /
    shift-right placement is exercised, and the same warm-up offset."""
    sys.path.insert(0, str(REF))
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import moe_model.model.moe  # noqa: F401
            reg = importlib.import_module("moe_model.model.moe.register")
    finally:
        sys.path.remove(str(REF))
    from competesmoe_b200.multimodal import CompeteSMoE

    def expert():
        m = nn.Module()
        m.fc1, m.fc2, m.activation_fn = nn.Linear(16, 24), nn.Linear(24, 16), nn.GELU(approximate="tanh")
        return m

    for rate, cap, warm, total, n_layers in ((0.5, 2, 0.0, 60, 5), (0.3, 1, 0.25, 80, 4), (0.9, 3, 0.1, 50, 6)):
        args = _mm_args()
        args.rate_flip, args.max_compete_in_iter, args.warm_up = rate, cap, warm
        results = []
        for make in (lambda: reg.get_moe("competesmoe"), lambda: CompeteSMoE):
            with contextlib.redirect_stdout(io.StringIO()):
                layers = [make()(in_embed_dim=16, out_embed_dim=16, num_of_experts=4, num_selected=2,
                                 expert=nn.ModuleList([expert() for _ in range(4)]), args=args) for _ in range(n_layers)]
                torch.manual_seed(1234)
                acc = {}
                for i, layer in enumerate(layers):
                    acc = layer.set_total_steps(total, id_layer=i, prob_flips_final=acc)
            results.append(([layer.prob_flips.clone().bool() for layer in layers], [layer.step_warm for layer in layers],
                            {k: v.clone().bool() for k, v in acc.items()}))
        (f_ref, w_ref, a_ref), (f_our, w_our, a_our) = results
        assert w_ref == w_our
        assert sorted(a_ref) == sorted(a_our)
        for a, b in zip(f_ref, f_our):
            assert a.shape == b.shape and torch.equal(a.cpu(), b.cpu())
        per_step = torch.stack([f.cpu() for f in f_our]).sum(0)
        assert int(per_step.max()) <= cap


def test_schedule_matches_reference_bit_for_bit_pretrain(tmp_path, monkeypatch):
    """Same for the pretrain plugin (layers/moe/competesmoe.py:123-273): the LM assigns one shared `prob_flips_final`
    dict to every layer (transformer_lm_mixin.py:265-267) and calls set_total_steps(id_layer) layer by layer."""
    monkeypatch.chdir(tmp_path)   # the reference appends to ./file_path.txt
    try:
        L, base, reg, saved, R = _pretrain_shim()
    except Exception as e:  # pragma: no cover - the reference tree changed
        pytest.skip(f"reference import shim failed: {e}")
    try:
        from competesmoe_b200.pretrain import CompeteSMoE
        for rate, cap, warm, total, n_layers in ((0.5, 2, 0.0, 60, 5), (0.3, 1, 0.25, 80, 4)):
            ns = SimpleNamespace(warm_up=warm, rate_flip=rate, stop_after=total, max_compete_in_iter=cap, is_cosine=False,
                                 is_norm_weight=False, norm_sigmoid=False, scale_weight=1.0, hybrid=False, tribrid=False,
                                 in_topk=False, balance_affinity=False, balance_loss_coef=0.01,
                                 balance_loss_coef_comp=0.01, router_loss_coef=0.01, router_theta=1.0, test_only=False)
            kw = dict(n_heads=2, args=ns, activation=F.relu, selection_mode="gate", log_interval=None)
            results = []
            for cls in (reg.get_moe("competesmoe"), CompeteSMoE):
                with contextlib.redirect_stdout(io.StringIO()):
                    layers = [cls(32, 4, 8, **kw) for _ in range(n_layers)]
                    shared = {}
                    torch.manual_seed(99)
                    for i, layer in enumerate(layers):
                        layer.prob_flips_final = shared
                        shared = layer.set_total_steps(id_layer=i)
                results.append(({k: v.clone().bool().cpu() for k, v in shared.items()}, [layer.step_warm for layer in layers]))
            (f_ref, w_ref), (f_our, w_our) = results
            assert w_ref == w_our and sorted(f_ref) == sorted(f_our) == list(range(n_layers))
            for k in f_ref:
                assert torch.equal(f_ref[k], f_our[k]), k
            assert int(torch.stack(list(f_our.values())).sum(0).max()) <= cap
    finally:
        sys.path.remove(R)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in [k for k in sys.modules if k.startswith(("layers.", "framework."))]:
            sys.modules.pop(k, None)
