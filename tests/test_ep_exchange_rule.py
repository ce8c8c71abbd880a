"""Host logic of the expert-parallel exchange choice and of the parity harness (CPU; no GPU, no process group)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))


def test_prefer_weights_follows_the_byte_counts():
    from competesmoe_b200.ep import WeightExchange as W
    # BASELINE configs[3]: d=1024, H=128, 64 experts top-8, 8192 tokens per GPU: 100 MB of weights+gradients vs 537 MB of rows
    assert W.prefer_weights(64, 2 * 1024 * 128, 8192, 8, 1024, 1024)
    # C1: d=512, H=128, 8 experts top-2, 4096 tokens
    assert W.prefer_weights(8, 2 * 512 * 128, 4096, 2, 512, 512)
    # BASELINE configs[1] (the 5.1B MLP block: 4 GLU experts of 3 x 3072 x 8192 parameters, top-2, 4096 tokens): rows are cheaper
    assert not W.prefer_weights(4, 3 * 3072 * 8192, 4096, 2, 3072, 3072)
    # the rule is the comparison of (2 + 4) bytes per expert parameter with 2 directions x 2 passes of bf16 rows
    E, P, T, K, D = 16, 2 * 256 * 128, 600, 4, 256
    assert W.prefer_weights(E, P, T, K, D, D) == ((2 + 4) * E * P < 2 * T * K * (D + D) * 2)


def test_pretrain_layer_rejects_unknown_exchange_mode():
    import pytest
    import torch.nn.functional as F
    from types import SimpleNamespace
    from competesmoe_b200.pretrain import CompeteSMoE
    args = SimpleNamespace(warm_up=0.0, rate_flip=0.07, stop_after=10, max_compete_in_iter=3, is_cosine=False,
                           is_norm_weight=False, norm_sigmoid=False, scale_weight=1.0, hybrid=False, tribrid=False,
                           in_topk=False, balance_affinity=True, balance_loss_coef=0.01, balance_loss_coef_comp=0.01,
                           router_loss_coef=0.01, router_theta=1.0, test_only=False)
    layer = CompeteSMoE(64, 8, 32, n_heads=2, args=args, activation=F.relu, selection_mode="gate", log_interval=None)
    with pytest.raises(ValueError):
        layer.enable_expert_parallel(SimpleNamespace(rank=0, world=1), max_tokens=16, exchange="rows")
    assert layer._wx is None and layer._ep is None


def test_parity_checks_are_recorded_not_raised():
    """tests/ep_worker.py: a failed comparison must not make one rank leave the call sequence (the others would wait for
    it in a device-side barrier); the verdict is taken collectively in finish_case()."""
    import ep_worker as ew
    ew.FAILS.clear()
    a = torch.zeros(8)
    b = torch.ones(8)
    ew.close(a, b, 1e-2, "demo")                 # records, does not raise
    assert len(ew.FAILS) == 1 and "demo" in ew.FAILS[0]
    try:
        ew.finish_case(torch.device("cpu"), 1, 0, "case")
    except AssertionError as exc:
        assert "1 failed check" in str(exc)
    else:
        raise AssertionError("finish_case must raise when a check failed")
    assert ew.FAILS == []
    # one bf16 ulp of the largest element is tolerated where sums of bf16-rounded rows are compared
    b = torch.cat([torch.tensor([3.0, 0.25]), torch.full((62,), 0.1)])      # rms 0.39: strict bound at 0.25 is 0.0128
    a = b.clone()
    a[1] += 2.0 ** -6                                                         # one ulp of a summand in [2, 4)
    ew.close(a, b, 2e-2, "strict")
    assert len(ew.FAILS) == 1
    ew.FAILS.clear()
    ew.close(a, b, 2e-2, "one ulp of a summand", summed_bf16=True)
    assert ew.FAILS == []
    ew.finish_case(torch.device("cpu"), 1, 0, "case")
