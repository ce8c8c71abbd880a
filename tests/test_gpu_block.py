"""The fused block tail (competesmoe_b200/pretrain_block.py; reference relative_moe_transformer.py:150-159):
`src + dropout(pkm(norm2(src)))` with LayerNorm+cast in one kernel and residual+dropout in the combine epilogue, against the
same three lines written with torch modules around the same layer."""
import pytest
import torch
import torch.nn.functional as F

from oracle import pretrain as op

from helpers import assert_close_rms

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _layer(D, H, E, K, competition=False):
    from competesmoe_b200.pretrain import CompeteSMoE
    torch.manual_seed(0)
    layer = CompeteSMoE(D, E, H, n_heads=K, args=op.default_args(stop_after=8), activation=F.relu, selection_mode="gate",
                        log_interval=None).to(DEV)
    layer.train()
    layer.regularization_present = True
    layer.step_warm = 0
    layer.prob_flips_final = {0: torch.full((8,), competition, device=DEV)}
    layer.set_current_steps(0)
    return layer


@pytest.mark.parametrize("H,competition", [(128, False), (64, False), (128, True)], ids=["fused-combine", "grouped-gemm", "competition"])
def test_block_tail_matches_the_three_reference_lines(H, competition):
    from competesmoe_b200.pretrain_block import FusedPreLNMoEBlock
    D, E, K, B, N = 256, 8, 2, 2, 160
    layer = _layer(D, H, E, K, competition)
    norm = torch.nn.LayerNorm(D).to(DEV)
    with torch.no_grad():
        norm.weight.normal_(1.0, 0.2)
        norm.bias.normal_(0.0, 0.2)
    block = FusedPreLNMoEBlock(norm, layer, 0.0).train()
    g = torch.Generator().manual_seed(1)
    src = torch.randn(B, N, D, generator=g).to(DEV)
    dy = torch.randn(B, N, D, generator=g).to(DEV)
    res = []
    for fused in (False, True):
        for p in list(layer.parameters()) + list(norm.parameters()):
            p.grad = None
        x = src.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if fused:
                out = block(x, id_layer=0)
            else:
                out = x + F.dropout(layer(norm(x), id_layer=0), 0.0, True)     # relative_moe_transformer.py:150-157
            regs = layer.get_reg_loss()
        ((out.float() * dy).sum() + sum(regs.values())).backward()
        res.append((out.detach().float(), x.grad.clone(), norm.weight.grad.clone(), norm.bias.grad.clone(),
                    layer.keys.grad.clone(), layer.w_gate.grad.clone(), layer.last_routing[0].clone()))
    ref, got = res
    assert got[0].dtype == torch.float32
    same = (ref[6] == got[6]).all(-1).reshape(-1)
    assert int((~same).sum()) <= 2, "LayerNorm rounding moved more than a couple of routing decisions"
    m = same.view(B, N)
    assert_close_rms(got[0][m], ref[0][m], 2e-2, "block output")
    if bool(same.all()):
        assert_close_rms(got[1], ref[1], 2e-2, "d src")
        assert_close_rms(got[2], ref[2], 2e-2, "d norm2.weight")
        assert_close_rms(got[3], ref[3], 2e-2, "d norm2.bias")
        assert_close_rms(got[4], ref[4], 3e-2, "d keys")
        assert_close_rms(got[5], ref[5], 3e-2, "d w_gate")


def test_layernorm_cast_kernel_matches_torch():
    from competesmoe_b200.functional import LayerNormCastFn
    g = torch.Generator().manual_seed(2)
    T, D = 300, 1024
    x = (torch.randn(T, D, generator=g) * 3 + 1).to(DEV).requires_grad_(True)
    w = torch.randn(D, generator=g).to(DEV).requires_grad_(True)
    b = torch.randn(D, generator=g).to(DEV).requires_grad_(True)
    dy = torch.randn(T, D, generator=g).to(DEV)
    y = LayerNormCastFn.apply(x, w, b, 1e-5, torch.float32)
    (y * dy).sum().backward()
    xr, wr, br = (t.detach().clone().requires_grad_(True) for t in (x, w, b))
    yr = F.layer_norm(xr, (D,), wr, br, 1e-5)
    (yr * dy).sum().backward()
    torch.testing.assert_close(y, yr, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(x.grad, xr.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(w.grad, wr.grad, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(b.grad, br.grad, rtol=1e-4, atol=1e-3)
    yb = LayerNormCastFn.apply(x.detach(), w.detach(), b.detach(), 1e-5, torch.bfloat16)
    assert yb.dtype == torch.bfloat16
    torch.testing.assert_close(yb.float(), yr.detach(), rtol=2.0 ** -8, atol=1e-5)     # one bf16 rounding of the fp32 result


def test_residual_dropout_mask_is_consistent_between_forward_and_backward():
    from competesmoe_b200.functional import ResidualDropoutFn
    g = torch.Generator().manual_seed(3)
    T, D, p, seed = 512, 256, 0.25, 12345
    v = (torch.randn(T, D, generator=g).abs() + 0.5).bfloat16().to(DEV).requires_grad_(True)   # never zero
    res = torch.randn(T, D, generator=g).to(DEV).requires_grad_(True)
    out = ResidualDropoutFn.apply(v, res, p, seed)
    gout = torch.randn(T, D, generator=g).to(DEV)
    (out * gout).sum().backward()
    kept = (out.detach() - res.detach()) != 0
    frac = float(kept.float().mean())
    assert abs(frac - (1 - p)) < 0.01, frac
    want = res.detach() + torch.where(kept, (v.detach().float() / (1 - p)).bfloat16().float(), torch.zeros_like(res))
    torch.testing.assert_close(out.detach(), want, rtol=0, atol=1e-6)
    torch.testing.assert_close(res.grad, gout)
    torch.testing.assert_close(v.grad.float(), torch.where(kept, gout / (1 - p), torch.zeros_like(gout)).bfloat16().float())
    out2 = ResidualDropoutFn.apply(v.detach(), res.detach(), p, seed)
    assert torch.equal(out2, out.detach())                                     # same seed -> same mask
    out3 = ResidualDropoutFn.apply(v.detach(), res.detach(), p, seed + 1)
    assert not torch.equal(out3, out.detach())


def test_fused_combine_tail_applies_dropout_like_the_stand_alone_kernel():
    """The combine epilogue and the stand-alone kernel draw the same mask for the same (seed, element)."""
    from competesmoe_b200 import ops
    g = torch.Generator().manual_seed(4)
    T, K, E, D, p, seed = 256, 2, 8, 256, 0.1, 777
    sel = torch.stack([torch.randperm(E, generator=g)[:K] for _ in range(T)]).int().to(DEV)
    w = torch.rand(T, K, generator=g).to(DEV)
    route = ops.route_build(sel, E)
    y = torch.randn(route.row_cap, D, generator=g).bfloat16().to(DEV)
    res = torch.randn(T, D, generator=g).to(DEV)
    plain = ops.combine_fwd(y, route.slot_to_row, route.sel, w, T, K, round_w=True)
    a = ops.combine_residual_fwd(y, route.slot_to_row, route.sel, w, T, K, res, p, seed, round_w=True)
    b = ops.residual_dropout_fwd(plain, res, p, seed)
    assert torch.equal(a, b)
