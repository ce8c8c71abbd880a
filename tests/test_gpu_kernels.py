"""GPU parity of the individual libcsmoe kernels (through the C ABI) against the CPU oracle / plain torch."""
import pytest
import torch
import torch.nn.functional as F

from oracle import multimodal as om
from oracle import pretrain as op

from helpers import assert_close_rms

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from competesmoe_b200 import ops as _ops
    return _ops


# ------------------------------------------------------------------------------------------------ routing metadata
@pytest.mark.parametrize("row_tile", [128, 256])
@pytest.mark.parametrize("T,K,E", [(1, 1, 1), (7, 2, 4), (4096, 2, 4), (4096, 2, 8), (3000, 8, 64), (65536, 8, 64),
                                   (513, 3, 1000)])
def test_route_build_bit_exact(ops, T, K, E, row_tile):
    g = torch.Generator().manual_seed(T + K + E)
    sel = torch.stack([torch.randperm(E, generator=g)[:K] for _ in range(min(T, 512))])
    sel = sel.repeat((T + sel.shape[0] - 1) // sel.shape[0], 1)[:T].int()
    if E >= 4:
        sel[sel == 1] = 0            # leave expert 1 empty, make expert 0 hot (duplicates within a token are fine here)
    r = ops.route_build(sel.to(DEV), E, row_tile=row_tile)
    ref = op.prepare_sel2(sel)       # stable argsort = the reference's maps after canonicalisation
    flat = sel.flatten()
    counts = torch.bincount(flat.long(), minlength=E)
    assert torch.equal(r.counts.cpu().long(), counts)
    assert torch.equal(r.offsets.cpu().long(), F.pad(counts.cumsum(0), (1, 0)))
    assert torch.equal(r.sorted_sel.cpu(), ref.sel.flatten())
    assert torch.equal(r.sort_index.cpu(), ref.out_index)
    assert torch.equal(r.sort_index.cpu() // K, ref.sel_index)
    pad = F.pad(((counts + row_tile - 1) // row_tile * row_tile).cumsum(0), (1, 0))
    assert torch.equal(r.pad_offsets.cpu().long(), pad)
    # slot <-> row maps are mutually inverse and expert-major
    s2r, r2s = r.slot_to_row.cpu().long(), r.row_to_slot.cpu().long()
    assert torch.equal(r2s[s2r], torch.arange(T * K))
    assert int((r2s >= 0).sum()) == T * K
    row_expert = torch.bucketize(s2r, pad[1:], right=True)
    assert torch.equal(row_expert, flat.long())
    te = r.tile_expert.cpu().long()
    n_tiles = int(pad[-1]) // 128
    assert torch.equal(te[:n_tiles], torch.bucketize(torch.arange(n_tiles) * 128, pad[1:], right=True))
    assert bool((te[n_tiles:] == -1).all())


def test_route_build_empty_input(ops):
    r = ops.route_build(torch.zeros(0, 2, dtype=torch.int32, device=DEV), 4)
    assert r.counts.tolist() == [0, 0, 0, 0] and r.pad_offsets.tolist() == [0] * 5
    assert bool((r.tile_expert == -1).all())


@pytest.mark.first_hw_run
def test_ops_accept_empty_inputs(ops):
    """Zero tokens (an empty micro-batch, a rank without tokens): the op-level calls return empty / all-padding results
    instead of failing on the NULL data pointer of a tensor without storage."""
    T, D, E, K = 0, 64, 4, 2
    x = torch.zeros(T, D, dtype=torch.bfloat16, device=DEV)
    wg = torch.randn(E, D, device=DEV).bfloat16()
    logits, probs, tw, ti = ops.router_fwd(x, wg, K)
    assert logits.shape == (0, E) and probs.shape == (0, E) and tw.shape == (0, K) and ti.shape == (0, K)
    route = ops.route_build(ti, E)
    assert int(route.counts.sum()) == 0 and bool((route.row_to_slot == -1).all())
    xp = ops.gather_rows(x, route)
    assert xp.shape == (route.row_cap, D) and not bool(xp.any())
    w = torch.zeros(T, K, device=DEV)
    assert ops.combine_fwd(xp, route.slot_to_row, route.sel, w, T, K).shape == (0, D)
    assert ops.scatter_reduce(xp, route.slot_to_row, T, K).shape == (0, D)
    assert ops.combine_bwd_w(xp, x, route.slot_to_row, T, K).shape == (0, K)
    assert ops.cast_bf16(torch.zeros(0, device=DEV)).numel() == 0
    wts, idx = ops.topk_renorm(torch.zeros(0, E, device=DEV), K)
    assert wts.shape == (0, K) and idx.shape == (0, K)


# ------------------------------------------------------------------------------------------------ router
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("T,D,E,K", [(257, 64, 4, 2), (1024, 1152, 4, 2), (512, 1024, 64, 8), (300, 512, 8, 2), (64, 3072, 4, 2)])
def test_router_matches_oracle(ops, dtype, T, D, E, K):
    g = torch.Generator().manual_seed(D + E)
    x = torch.randn(1, T, D, generator=g).to(dtype)
    wg = (torch.randn(E, D, generator=g) * 0.05).to(dtype)
    w_ref, idx_ref, p_ref, l_ref = om.router_policy(x, wg, K)
    logits, probs, tw, ti = ops.router_fwd(x[0].to(DEV), wg.to(DEV), K)
    assert_close_rms(logits, l_ref[0], 2e-2 if dtype == torch.bfloat16 else 1e-4, "logits")
    torch.testing.assert_close(probs.cpu(), p_ref[0], rtol=2e-2 if dtype == torch.bfloat16 else 1e-4, atol=1e-6)
    # routing: bit-exact except tokens with margin < 1e-3 (computed on the oracle's probabilities)
    margin = om.topk_margin(p_ref[0], K)
    agree = (ti.cpu().long() == idx_ref[0]).all(-1)
    assert bool((margin[~agree] < 1e-3).all()), "routing differs on a token with margin >= 1e-3"
    torch.testing.assert_close(tw.cpu()[agree], w_ref[0][agree].float(), rtol=2e-2 if dtype == torch.bfloat16 else 1e-4, atol=1e-6)
    print(f"router {dtype} T={T} E={E}: {int((~agree).sum())} low-margin tokens exempt")


def test_router_tie_break_is_lowest_index(ops):
    x = torch.zeros(8, 64, device=DEV, dtype=torch.bfloat16)      # all logits equal -> all probabilities tie
    wg = torch.randn(8, 64, device=DEV, dtype=torch.bfloat16)
    _, probs, tw, ti = ops.router_fwd(x, wg, 3)
    assert ti.cpu().tolist() == [[0, 1, 2]] * 8
    torch.testing.assert_close(tw.sum(-1).cpu(), torch.ones(8), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("E,K", [(4, 2), (64, 8), (33, 5)])
def test_topk_renorm(ops, E, K, T=500):
    g = torch.Generator().manual_seed(E)
    s = torch.rand(T, E, generator=g)
    w, idx = ops.topk_renorm(s.to(DEV), K)
    wr, ir = om.stable_topk(s, K)
    assert torch.equal(idx.cpu().long(), ir)
    torch.testing.assert_close(w.cpu(), wr / wr.sum(-1, keepdim=True), rtol=1e-5, atol=1e-7)
    w2, idx2 = ops.topk_renorm(s.to(DEV), K, sigmoid=True)
    ws = torch.sigmoid(wr)
    assert torch.equal(idx2.cpu().long(), ir)
    torch.testing.assert_close(w2.cpu(), ws / ws.sum(-1, keepdim=True), rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------------------------------------ permute / combine
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_gather_combine_scatter(ops, dtype):
    T, K, E, D = 777, 2, 4, 256
    g = torch.Generator().manual_seed(5)
    x = torch.randn(T, D, generator=g).to(dtype)
    sel = torch.stack([torch.randperm(E, generator=g)[:K] for _ in range(T)]).int()
    w = torch.rand(T, K, generator=g)
    r = ops.route_build(sel.to(DEV), E)
    xp = ops.gather_rows(x.to(DEV), r)
    r2s = r.row_to_slot.cpu().long()
    ref = torch.zeros(r.row_cap, D, dtype=dtype)
    ref[r2s >= 0] = x[r2s[r2s >= 0] // K]
    assert torch.equal(xp.cpu(), ref)                                   # pure data movement: bit-exact, zero padding
    # combine of the gathered rows with weights == sum_k w * x[t]
    out = ops.combine_fwd(xp, r.slot_to_row, r.sel, w.to(DEV), T, K)
    torch.testing.assert_close(out.cpu().float(), (w.sum(-1, keepdim=True) * x.float()), rtol=2e-2, atol=2e-2)
    # scatter_reduce is the adjoint of gather
    dx = ops.scatter_reduce(xp, r.slot_to_row, T, K)
    torch.testing.assert_close(dx.cpu().float(), K * x.float(), rtol=1e-2, atol=1e-2)
    dw = ops.combine_bwd_w(xp, x.to(DEV), r.slot_to_row, T, K)
    torch.testing.assert_close(dw.cpu(), (x.float() ** 2).sum(-1, keepdim=True).expand(T, K), rtol=1e-2, atol=1e-2)
    # weighted gather (combine backward)
    gw = ops.gather_rows(x.to(DEV), r, slot_w=w.to(DEV))
    refw = torch.zeros(r.row_cap, D)
    refw[r2s >= 0] = x[r2s[r2s >= 0] // K].float() * w.flatten()[r2s[r2s >= 0]].unsqueeze(1)
    torch.testing.assert_close(gw.cpu().float(), refw, rtol=1e-2, atol=1e-2)


def test_combine_order_matches_reference_rounding(ops):
    """moe.py:204: results accumulate in bf16, experts visited in ascending id."""
    T, K, E, D = 64, 2, 4, 128
    g = torch.Generator().manual_seed(9)
    sel = torch.stack([torch.randperm(E, generator=g)[:K] for _ in range(T)]).int()
    y = torch.randn(T * K, D, generator=g).bfloat16()       # row j = output for slot j
    w = torch.rand(T, K, generator=g)
    ref = torch.zeros(T, D, dtype=torch.bfloat16)
    for e in range(E):
        t_idx, k_idx = torch.where(sel == e)
        ref[t_idx] += w[t_idx, k_idx].unsqueeze(0).T * y[t_idx * K + k_idx]
    rows = torch.arange(T * K, dtype=torch.int32, device=DEV)
    out = ops.combine_fwd(y.to(DEV), rows, sel.flatten().to(DEV), w.to(DEV), T, K, round_each=True)
    assert torch.equal(out.cpu(), ref)


# ------------------------------------------------------------------------------------------------ grouped GEMM
def _route_for_counts(ops, counts, row_tile=128):
    sel = torch.cat([torch.full((c,), e, dtype=torch.int32) for e, c in enumerate(counts)])
    return ops.route_build(sel.view(-1, 1).to(DEV), len(counts), row_tile=row_tile)


@pytest.mark.parametrize("row_tile", [128, 256])     # 256: the CTA-pair (cta_group::2) kernel where n >= 256
@pytest.mark.parametrize("kn", [False, True])
@pytest.mark.parametrize("counts,n,k", [([128], 256, 64), ([300, 0, 77, 513], 4304, 1152), ([300, 0, 77, 513], 1152, 4304),
                                        ([100, 200, 300], 128, 512), ([5, 5, 5, 5, 5, 5, 5, 5], 64, 1024)])
def test_gemm_rows(ops, kn, counts, n, k, row_tile):
    E = len(counts)
    r = _route_for_counts(ops, counts, row_tile)
    g = torch.Generator().manual_seed(n + k)
    a = torch.zeros(r.row_cap, k, dtype=torch.bfloat16)
    pad = r.pad_offsets.cpu().tolist()
    for e, c in enumerate(counts):
        a[pad[e]:pad[e] + c] = torch.randn(c, k, generator=g).bfloat16()
    w = (torch.randn(E, k, n, generator=g) / k ** 0.5).bfloat16() if kn else (torch.randn(E, n, k, generator=g) / k ** 0.5).bfloat16()
    bias = torch.randn(E, n, generator=g).bfloat16()
    c_, pre = ops.gemm_rows(a.to(DEV), w.to(DEV), w_is_kn=kn, route=r, bias=bias.to(DEV), act=ops.ACT_GELU_TANH, want_preact=True)
    for e, cnt in enumerate(counts):
        we = w[e].float() if kn else w[e].float().t()
        z = (a[pad[e]:pad[e] + cnt].float() @ we + bias[e].float()).bfloat16()
        assert_close_rms(pre[pad[e]:pad[e] + cnt], z, 2e-2, f"preact e={e}")
        assert_close_rms(c_[pad[e]:pad[e] + cnt], F.gelu(z.float(), approximate="tanh"), 2e-2, f"act e={e}")


@pytest.mark.parametrize("row_tile", [128, 256])
@pytest.mark.parametrize("counts,m,n", [([128], 128, 256), ([300, 0, 77, 513], 4304, 1152), ([300, 0, 77, 513], 1152, 4304),
                                        ([100, 200, 300], 512, 128)])
def test_gemm_reduce(ops, counts, m, n, row_tile):
    E = len(counts)
    r = _route_for_counts(ops, counts, row_tile)
    g = torch.Generator().manual_seed(m + n)
    a = torch.zeros(r.row_cap, m, dtype=torch.bfloat16)
    b = torch.zeros(r.row_cap, n, dtype=torch.bfloat16)
    pad = r.pad_offsets.cpu().tolist()
    for e, c in enumerate(counts):
        a[pad[e]:pad[e] + c] = torch.randn(c, m, generator=g).bfloat16()
        b[pad[e]:pad[e] + c] = torch.randn(c, n, generator=g).bfloat16()
    out = ops.gemm_reduce(a.to(DEV), b.to(DEV), E, route=r, out_dtype=torch.float32)
    for e, c in enumerate(counts):
        ref = a[pad[e]:pad[e] + c].float().t() @ b[pad[e]:pad[e] + c].float()
        assert_close_rms(out[e], ref, 1e-4, f"wgrad e={e}")


@pytest.mark.parametrize("row_tile", [128, 256])
def test_gemm_fused_glu_forward_and_backward_epilogues(ops, row_tile):
    counts = [300, 0, 77, 513]
    E, D, Fh = 4, 512, 384
    r = _route_for_counts(ops, counts, row_tile)
    g = torch.Generator().manual_seed(11)
    pad = r.pad_offsets.cpu().tolist()
    x = torch.zeros(r.row_cap, D, dtype=torch.bfloat16)
    dy = torch.zeros(r.row_cap, D, dtype=torch.bfloat16)
    for e, c in enumerate(counts):
        x[pad[e]:pad[e] + c] = torch.randn(c, D, generator=g).bfloat16()
        dy[pad[e]:pad[e] + c] = torch.randn(c, D, generator=g).bfloat16()
    w1 = (torch.randn(E, 2 * Fh, D, generator=g) / D ** 0.5).bfloat16()
    w2 = (torch.randn(E, D, Fh, generator=g) / Fh ** 0.5).bfloat16()
    h, z = ops.gemm_rows(x.to(DEV), w1.to(DEV), w_is_kn=False, route=r, act=ops.ACT_SILU_GLU)
    assert h.shape == (r.row_cap, Fh) and z.shape == (r.row_cap, 2 * Fh)
    dz = ops.gemm_rows(dy.to(DEV), w2.to(DEV), w_is_kn=True, route=r, act_bwd=ops.ACT_SILU_GLU, aux=z)
    assert dz.shape == (r.row_cap, 2 * Fh)
    for e, c in enumerate(counts):
        if c == 0:
            continue
        sl = slice(pad[e], pad[e] + c)
        zr = (x[sl].float() @ w1[e].float().t()).bfloat16()
        assert_close_rms(z[sl], zr, 2e-2, f"z e={e}")
        zz = z[sl].cpu().float().requires_grad_(True)
        gate, up = zz.chunk(2, dim=-1)
        hr = up * F.silu(gate)
        assert_close_rms(h[sl], hr.detach(), 2e-2, f"h e={e}")
        dh = (dy[sl].float() @ w2[e].float()).bfloat16().float()
        hr.backward(dh)
        assert_close_rms(dz[sl], zz.grad, 3e-2, f"dz e={e}")


@pytest.mark.parametrize("act", ["ACT_RELU", "ACT_GELU_TANH"])
def test_gemm_fused_activation_backward(ops, act):
    counts = [200, 131]
    E, D, Fh = 2, 256, 520
    r = _route_for_counts(ops, counts, 128)
    g = torch.Generator().manual_seed(12)
    pad = r.pad_offsets.cpu().tolist()
    z = torch.randn(r.row_cap, Fh, generator=g).bfloat16()
    dy = torch.randn(r.row_cap, D, generator=g).bfloat16()
    w2 = (torch.randn(E, D, Fh, generator=g) / Fh ** 0.5).bfloat16()
    code = getattr(ops, act)
    dz = ops.gemm_rows(dy.to(DEV), w2.to(DEV), w_is_kn=True, route=r, act_bwd=code, aux=z.to(DEV))
    fn = F.relu if act == "ACT_RELU" else (lambda t: F.gelu(t, approximate="tanh"))
    for e, c in enumerate(counts):
        sl = slice(pad[e], pad[e] + c)
        zz = z[sl].float().requires_grad_(True)
        fn(zz).backward((dy[sl].float() @ w2[e].float()).bfloat16().float())
        assert_close_rms(dz[sl], zz.grad, 3e-2, f"dz e={e}")


def test_router_aux_and_backward_match_autograd(ops, B=3, N=200, D=256, E=8, K=2):
    """csmoe_router_aux_fwd / csmoe_router_bwd against torch autograd of the reference formulas (moe.py:71-132)."""
    g = torch.Generator().manual_seed(13)
    x = torch.randn(B * N, D, generator=g)
    wg = torch.randn(E, D, generator=g) * 0.05
    dtw = torch.randn(B * N, K, generator=g)
    dpx = torch.randn(B * N, E, generator=g) * 0.1
    xr, wr = x.clone().requires_grad_(True), wg.clone().requires_grad_(True)
    w_ref, idx_ref, p_ref, l_ref = om.router_policy(xr.view(B, N, D), wr, K)
    bal = om.balanceloss(idx_ref, p_ref, E)
    zl = om.zloss(l_ref)
    loss = 0.7 * bal + 0.3 * zl + (w_ref * dtw.view(B, N, K)).sum() + (p_ref * dpx.view(B, N, E)).sum()
    loss.backward()
    logits, probs, tw, ti = ops.router_fwd(x.to(DEV), wg.to(DEV), K)
    assert torch.equal(ti.cpu().long().view(B, N, K), idx_ref)
    losses, cnt, lse = ops.router_aux_fwd(logits, probs, ti, B)
    torch.testing.assert_close(losses.cpu(), torch.stack([bal, zl]).detach(), rtol=1e-4, atol=1e-6)
    gl = torch.tensor([0.7, 0.3], device=DEV)
    dx, dwg = ops.router_bwd(x.to(DEV), wg.to(DEV), probs, tw, ti, B, dtw=dtw.to(DEV), dprobs=dpx.to(DEV), lse=lse,
                             cnt=cnt, g_losses=gl)
    assert_close_rms(dx, xr.grad, 1e-3, "dx")
    assert_close_rms(dwg, wr.grad, 1e-3, "dwg")


def test_gemm_linearity_at_full_size(ops):
    """Size-independent property at the bench shape (C2 rows x D x 2F): G(a1 + a2) == G(a1) + G(a2) up to rounding."""
    counts = [2048] * 4
    r = _route_for_counts(ops, counts, 256)
    D, N = 3072, 16384
    g = torch.Generator(device=DEV).manual_seed(0)
    # small integers: a1 + a2 is exact in bf16, so the identity holds up to fp32 accumulation order only
    a1 = torch.randint(-3, 4, (r.row_cap, D), device=DEV, generator=g).bfloat16()
    a2 = torch.randint(-3, 4, (r.row_cap, D), device=DEV, generator=g).bfloat16()
    w = (torch.randn(4, N, D, device=DEV, generator=g) * 0.02).bfloat16()
    s = a1 + a2
    used = int(r.pad_offsets[-1])           # row tiles past the last expert are never written
    y1 = ops.gemm_rows(a1, w, w_is_kn=False, route=r, out_dtype=torch.float32)[:used]
    y2 = ops.gemm_rows(a2, w, w_is_kn=False, route=r, out_dtype=torch.float32)[:used]
    ys = ops.gemm_rows(s, w, w_is_kn=False, route=r, out_dtype=torch.float32)[:used]
    assert_close_rms(ys, y1 + y2, 1e-4, "linearity")
    # and a spot check of one row tile against torch
    ref = a1[:128].float() @ w[0].float().t()
    assert_close_rms(y1[:128], ref, 1e-3, "tile 0")


# ------------------------------------------------------------------------------------------------ elementwise
@pytest.mark.parametrize("act,fn", [("ACT_RELU", F.relu), ("ACT_GELU", F.gelu), ("ACT_GELU_TANH", lambda z: F.gelu(z, approximate="tanh")),
                                    ("ACT_SILU", F.silu)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_activation_fwd_bwd(ops, act, fn, dtype):
    code = getattr(ops, act)
    g = torch.Generator().manual_seed(1)
    z = (torch.randn(300, 264, generator=g) * 2).to(dtype)
    dh = torch.randn(300, 264, generator=g).to(dtype)
    zr = z.float().requires_grad_(True)
    fn(zr).backward(dh.float())
    tol = 2e-2 if dtype == torch.bfloat16 else 1e-4
    assert_close_rms(ops.act_fwd(z.to(DEV), code), fn(z.float()), tol, "fwd")
    assert_close_rms(ops.act_bwd(z.to(DEV), dh.to(DEV), code), zr.grad, tol, "bwd")


def test_glu_fwd_bwd(ops):
    g = torch.Generator().manual_seed(2)
    z = torch.randn(200, 512, generator=g).bfloat16()
    dh = torch.randn(200, 256, generator=g).bfloat16()
    zr = z.float().requires_grad_(True)
    gate, up = zr.chunk(2, dim=-1)
    (up * F.silu(gate)).backward(dh.float())
    assert_close_rms(ops.act_fwd(z.to(DEV), ops.ACT_SILU_GLU), (z.float()[:, 256:] * F.silu(z.float()[:, :256])), 2e-2, "glu fwd")
    assert_close_rms(ops.act_bwd(z.to(DEV), dh.to(DEV), ops.ACT_SILU_GLU), zr.grad, 2e-2, "glu bwd")


def test_bias_grad_and_cast(ops):
    counts = [300, 0, 77, 513]
    r = _route_for_counts(ops, counts)
    g = torch.Generator().manual_seed(3)
    x = torch.zeros(r.row_cap, 4304)
    pad = r.pad_offsets.cpu().tolist()
    for e, c in enumerate(counts):
        x[pad[e]:pad[e] + c] = torch.randn(c, 4304, generator=g)
    db = ops.bias_grad(x.bfloat16().to(DEV), 4, route=r, out_dtype=torch.float32)
    ref = torch.stack([x.bfloat16().float()[pad[e]:pad[e + 1]].sum(0) for e in range(4)])
    assert_close_rms(db, ref, 1e-3, "bias grad")
    src = torch.randn(1000003, generator=g)
    assert torch.equal(ops.cast_bf16(src.to(DEV)).cpu(), src.bfloat16())


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_affinity_fwd_bwd(ops, dtype):
    E, T, D = 4, 200, 1152
    t_pad = 256
    g = torch.Generator().manual_seed(4)
    y = torch.randn(E, t_pad, D, generator=g).to(dtype)
    daff = torch.randn(T, E, generator=g)
    yr = y.float().requires_grad_(True)
    aff_ref = F.softplus(yr[:, :T]).mean(-1).t()        # [T, E]
    aff_ref.backward(daff)
    aff = ops.affinity_fwd(y.view(E * t_pad, D).to(DEV), E, T, t_pad, dtype == torch.bfloat16)
    tol = 2e-2 if dtype == torch.bfloat16 else 1e-4
    assert_close_rms(aff, aff_ref.detach(), tol, "affinity")
    dy = ops.affinity_bwd(y.view(E * t_pad, D).to(DEV), daff.to(DEV), E, T, t_pad)
    assert_close_rms(dy.view(E, t_pad, D), yr.grad, tol, "affinity bwd")


# ------------------------------------------------------------------------------------------------ competition tail
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("E,K,T,t_pad,D", [(4, 2, 300, 512, 1152), (8, 1, 5, 256, 64), (64, 8, 130, 256, 1024),
                                           (6, 3, 257, 512, 264)])
def test_diversity_and_fused_competition_backward(ops, dtype, E, K, T, t_pad, D):
    """csmoe_diversity_fwd / csmoe_compete_bwd against autograd of the reference formulas (competesmoe.py:180-259)."""
    g = torch.Generator().manual_seed(E * 1000 + K)
    y = torch.randn(E, t_pad, D, generator=g).to(dtype).to(DEV)
    scores = torch.rand(T, E, generator=g)
    sel = scores.topk(K, dim=-1).indices.int().to(DEV)
    w = torch.rand(T, K, generator=g).to(DEV)
    daff = torch.randn(T, E, generator=g).to(DEV)
    dout = torch.randn(T, D, generator=g).to(dtype).to(DEV)
    g_div = torch.tensor(0.7, device=DEV)

    loss, inv_norm, sim = ops.diversity_fwd(y.view(E * t_pad, D), sel, T, t_pad)
    dy = ops.compete_bwd(y.view(E * t_pad, D), E, T, t_pad, sel, daff=daff, w=w, dout=dout, inv_norm=inv_norm, sim=sim,
                         g_div=g_div).view(E, t_pad, D)

    yr = y.float().requires_grad_(True)
    tok = torch.arange(T, device=DEV).unsqueeze(1)
    top = yr[sel.long(), tok]                                             # [T, K, D]
    nrm = F.normalize(top, p=2, dim=-1)
    simr = torch.bmm(nrm, nrm.transpose(1, 2))
    lossr = (simr * (1 - torch.eye(K, device=DEV))).mean()
    aff = F.softplus(yr[:, :T]).mean(-1).t()                              # [T, E]
    out = (w.unsqueeze(-1) * top).sum(1)
    total = (aff * daff).sum() + (out * dout.float()).sum() + lossr * g_div
    (dyr,) = torch.autograd.grad(total, yr)

    assert abs(loss.item() - lossr.item()) < 1e-5 + 1e-4 * abs(lossr.item())
    torch.testing.assert_close(sim, simr.detach(), rtol=1e-4, atol=1e-5)
    assert bool((dy[:, T:] == 0).all())
    tol = 2e-2 if dtype == torch.bfloat16 else 1e-4
    assert_close_rms(dy.float(), dyr, rtol=tol, what="d(dense outputs)")
    # each term on its own (NULL pointers switch the others off)
    only_aff = ops.compete_bwd(y.view(E * t_pad, D), E, T, t_pad, sel, daff=daff).view(E, t_pad, D)
    (ref_aff,) = torch.autograd.grad((F.softplus(yr[:, :T]).mean(-1).t() * daff).sum(), yr)
    assert_close_rms(only_aff.float(), ref_aff, rtol=tol, what="score term")


# ------------------------------------------------------------------------------------------------ GEMM variants
@pytest.mark.parametrize("env", [
    {"CSMOE_GEMM_EPI": "staged", "CSMOE_GEMM_WIDE": "15"},   # staged epilogue everywhere, 256x512 tiles wherever legal
    {"CSMOE_GEMM_EPI": "direct", "CSMOE_GEMM_WIDE": "0"},    # register-direct epilogue, 256x256 CTA-pair tiles only
    {"CSMOE_GEMM_PAIR": "0"},                                 # single-CTA kernels only
    {"CSMOE_GEMM_EPI": "tma", "CSMOE_GEMM_WIDE": "0"},       # TMA-store epilogue forced wherever legal, pair tiles only
    {"CSMOE_GEMM_RASTER": "1"},                               # fixed bands of 8 m-blocks instead of expert-aligned bands
    {"CSMOE_GEMM_BAND_CAP": "2"},                             # expert-aligned bands split into several bands per expert
])
def test_gemm_kernel_variants_in_subprocess(env):
    """The kernel / epilogue variant is picked per launch by shape; the switches that force one variant are read once
    per process, so the GEMM tests are re-run in a child process under each setting."""
    import os
    import subprocess
    import sys
    if os.environ.get("CSMOE_VARIANT_CHILD"):
        pytest.skip("already inside a variant run")
    child_env = dict(os.environ, CSMOE_VARIANT_CHILD="1", **env)
    r = subprocess.run([sys.executable, "-m", "pytest", __file__, "-q", "-x", "-m", "gpu", "-k", "gemm and not variants"],
                       env=child_env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("kn", [False, True])
@pytest.mark.parametrize("E,T,n,k", [(4, 256, 512, 128), (3, 512, 1152, 320), (8, 256, 3072, 1024), (64, 256, 1024, 128)])
def test_gemm_dense_sum_over_experts(ops, kn, E, T, n, k):
    """C[t] = sum_e A[e, t] . W[e] in one launch (k loop over experts): the dense pass's gradient w.r.t. its shared input."""
    g = torch.Generator().manual_seed(E + T + n)
    a = (torch.randn(E, T, k, generator=g) * 0.5).bfloat16().to(DEV)
    w = (torch.randn(E, k, n, generator=g) * 0.1).bfloat16().to(DEV)
    wk = w if kn else w.transpose(1, 2).contiguous()          # [E, k, n] or [E, n, k]
    c = ops.gemm_rows(a.view(E * T, k), wk, w_is_kn=kn, dense_rows=T, a_expert_rows=T, sum_experts=True)
    ref = torch.einsum("etk,ekn->tn", a.float(), w.float())
    assert c.shape == (T, n)
    assert_close_rms(c, ref, 2e-2, "sum over experts")


@pytest.mark.parametrize("eager_bf16", [True, False])
@pytest.mark.parametrize("E,T,n,k,kn", [(4, 256, 1152, 320, False), (8, 512, 1024, 128, True), (3, 256, 96, 64, True)])
def test_gemm_epilogue_score_matches_affinity_kernel(ops, eager_bf16, E, T, n, k, kn):
    """The neural-response score reduced in the down projection's epilogue (csmoe_gemm_args.rowsum) against the
    stand-alone kernel that re-reads the dense outputs, and against torch (competesmoe.py:243)."""
    g = torch.Generator().manual_seed(E * n + k)
    h = (torch.randn(E * T, k, generator=g) * 0.5).bfloat16().to(DEV)
    w = (torch.randn(E, k, n, generator=g) * 0.2).bfloat16().to(DEV)
    wk = w if kn else w.transpose(1, 2).contiguous()
    y, rs = ops.gemm_rows(h, wk, w_is_kn=kn, dense_rows=T, a_expert_rows=T, rowsum_softplus=eager_bf16)
    aff = ops.affinity_from_rowsum(rs, E, T - 7, T, n, eager_bf16)
    ref_kernel = ops.affinity_fwd(y, E, T - 7, T, eager_bf16)
    sp = F.softplus(y.float().view(E, T, n)[:, :T - 7])
    ref = (sp.bfloat16().float() if eager_bf16 else sp).mean(-1).t()
    tol = 2 ** -8 if eager_bf16 else 1e-5           # one bf16 ulp when the result is rounded to bf16
    assert float((aff - ref_kernel).abs().max() / ref_kernel.abs().max()) <= tol
    assert float((aff - ref).abs().max() / ref.abs().max()) <= tol


@pytest.mark.parametrize("act", ["relu", "gelu", "gelu_tanh", "silu"])
@pytest.mark.parametrize("dense", [False, True])
def test_act_bwd_bias_matches_separate_kernels(ops, act, dense):
    """csmoe_act_bwd_bias = csmoe_act_bwd followed by csmoe_bias_grad, bit for bit (dz is rounded before it is summed)."""
    code = {"relu": ops.ACT_RELU, "gelu": ops.ACT_GELU, "gelu_tanh": ops.ACT_GELU_TANH, "silu": ops.ACT_SILU}[act]
    torch.manual_seed(5)
    E, n = 4, 520
    if dense:
        rows, route, kw = E * 256, None, dict(dense_rows=256)
    else:
        sel = torch.stack([torch.randperm(E)[:2] for _ in range(333)]).int().to(DEV)
        route = ops.route_build(sel, E)
        rows, kw = route.row_cap, dict(route=route)
    z = torch.randn(rows, n, device=DEV, dtype=torch.bfloat16)
    dh = torch.randn(rows, n, device=DEV, dtype=torch.bfloat16)
    dz_ref = ops.act_bwd(z, dh, code, route)
    db_ref = ops.bias_grad(dz_ref, E, **kw)
    dz, db = ops.act_bwd_bias(z, dh, code, E, **kw)
    if dense:
        assert torch.equal(dz, dz_ref)
    else:
        po = route.pad_offsets.cpu().tolist()
        for e in range(E):
            assert torch.equal(dz[po[e]:po[e + 1]], dz_ref[po[e]:po[e + 1]])
    assert torch.equal(db, db_ref)


# ------------------------------------------------------------------------------------------------ loss kernels (losses.cu)
def _torch_comp_losses(p, aff, aff_idx, gate_idx, B):
    """Plain PyTorch fp32 restatement of the five scalars csmoe_losses_fwd produces (reference lines in include/csmoe.h)."""
    import math
    T, E = p.shape
    N = T // B
    q = F.softmax(aff, dim=-1)
    qd = q.detach()
    li, gi = aff_idx.long(), gate_idx.long()
    l0 = F.mse_loss(p, qd)
    l1 = F.mse_loss(torch.gather(p, -1, li), torch.gather(qd, -1, li))
    l2 = F.mse_loss(torch.gather(p, -1, gi), torch.gather(qd, -1, gi))
    q3 = q.view(B, N, E)
    top1 = F.one_hot(li.view(B, N, -1)[..., 0], E).float()
    l3 = (q3.mean(-2) * top1.mean(-2)).mean() * float(E * E)
    ls = F.log_softmax(q3, dim=-1)
    lm = ls.logsumexp(-2) - math.log(N)
    l4 = (lm * lm.exp()).sum(-1).mean()
    return q, torch.stack([l0, l1, l2, l3, l4])


@pytest.mark.parametrize("B,N,E,K", [(1, 4096, 4, 2), (3, 200, 8, 2), (4, 1024, 64, 8), (2, 77, 33, 5)])
def test_compete_losses_kernels_match_torch(B, N, E, K):
    from competesmoe_b200.functional import CompeteLossesFn
    g = torch.Generator().manual_seed(B * 1000 + E)
    T = B * N
    p = F.softmax(torch.randn(T, E, generator=g), -1).to(DEV).requires_grad_(True)
    aff = (0.7 + 0.1 * torch.randn(T, E, generator=g)).to(DEV).requires_grad_(True)
    aff_idx = torch.stack([torch.randperm(E, generator=g)[:K] for _ in range(T)]).int().to(DEV)
    gate_idx = torch.stack([torch.randperm(E, generator=g)[:K] for _ in range(T)]).int().to(DEV)
    coef = torch.tensor([1.0, 0.7, -0.3, 0.5, 2.0], device=DEV)
    q, losses = CompeteLossesFn.apply(p, aff, aff_idx, gate_idx, B)
    (losses * coef).sum().backward()
    pr, ar = p.detach().clone().requires_grad_(True), aff.detach().clone().requires_grad_(True)
    q_ref, l_ref = _torch_comp_losses(pr, ar, aff_idx, gate_idx, B)
    (l_ref * coef).sum().backward()
    torch.testing.assert_close(q, q_ref, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(losses, l_ref, rtol=2e-5, atol=1e-8)
    torch.testing.assert_close(p.grad, pr.grad, rtol=1e-4, atol=1e-10)
    torch.testing.assert_close(aff.grad, ar.grad, rtol=1e-3, atol=1e-9)
    # deterministic: bit-identical on a second run
    q2, losses2 = CompeteLossesFn.apply(p.detach(), aff.detach(), aff_idx, gate_idx, B)
    assert torch.equal(losses2, losses.detach()) and torch.equal(q2, q)
    # gate_idx = None: term [2] is zero and nothing else changes
    _, l3 = CompeteLossesFn.apply(p.detach(), aff.detach(), aff_idx, None, B)
    assert float(l3[2]) == 0.0 and torch.equal(l3[[0, 1, 3, 4]], losses.detach()[[0, 1, 3, 4]])


@pytest.mark.parametrize("B,N,E", [(1, 512, 4), (8, 512, 8), (4, 1024, 64)])
def test_entropy_balance_kernel_matches_pretrain_formula(B, N, E):
    """moe_pretrain_model/layers/moe/moe.py:323-332 on logits == the kernel on softmax(logits)."""
    import math
    from competesmoe_b200.functional import EntropyBalanceFn
    g = torch.Generator().manual_seed(E)
    logits = torch.randn(B, N, E, generator=g).to(DEV).requires_grad_(True)
    ls = F.log_softmax(logits.float(), dim=-1)
    lm = ls.logsumexp(-2) - math.log(N)
    ref = (lm * lm.exp()).sum(-1).mean()
    ref.backward()
    lg = logits.detach().clone().requires_grad_(True)
    got = EntropyBalanceFn.apply(F.softmax(lg, dim=-1).view(B * N, E), B)
    got.backward()
    torch.testing.assert_close(got, ref, rtol=2e-5, atol=1e-8)
    torch.testing.assert_close(lg.grad, logits.grad, rtol=1e-3, atol=1e-9)


@pytest.mark.parametrize("sigmoid", [False, True])
def test_topk_renorm_bwd_and_dense_rows(sigmoid, ops):
    g = torch.Generator().manual_seed(5)
    T, E, K, t_pad = 300, 16, 4, 512
    scores = (0.7 + 0.2 * torch.randn(T, E, generator=g)).to(DEV)
    w, idx = ops.topk_renorm(scores, K, sigmoid=sigmoid)
    dw = torch.randn(T, K, generator=g).to(DEV)
    s = scores.clone().requires_grad_(True)
    v = torch.gather(torch.sigmoid(s) if sigmoid else s, 1, idx.long())
    (v / v.sum(-1, keepdim=True) * dw).sum().backward()
    got = ops.topk_renorm_bwd(scores, w, idx, dw, sigmoid)
    torch.testing.assert_close(got, s.grad, rtol=1e-4, atol=1e-7)
    base = torch.randn(T, E, generator=g).to(DEV)
    acc = ops.topk_renorm_bwd(scores, w, idx, dw, sigmoid, out=base.clone())
    torch.testing.assert_close(acc, base + s.grad, rtol=1e-4, atol=1e-6)
    rows = ops.dense_rows(idx, t_pad)
    want = (idx.long() * t_pad + torch.arange(T, device=DEV).unsqueeze(1)).reshape(-1).int()
    assert torch.equal(rows, want)


def test_topk_is_total_on_non_finite_scores(ops):
    """NaN / Inf scores must still give K distinct in-range expert ids (NaN sorts as the largest value, like
    torch.topk), and the routing maps built from them must stay in bounds."""
    T, E, K = 64, 8, 2
    scores = torch.randn(T, E, device=DEV)
    scores[3] = float("nan")
    scores[5, 2] = float("nan")
    scores[7, 1] = float("inf")
    scores[9] = float("-inf")
    w, idx = ops.topk_renorm(scores, K)
    assert int(idx.min()) >= 0 and int(idx.max()) < E
    assert bool((idx[:, 0] != idx[:, 1]).all())
    assert idx[3].tolist() == [0, 1] and idx[5, 0].item() == 2 and idx[7, 0].item() == 1
    x = torch.randn(T, 64, device=DEV, dtype=torch.bfloat16)
    x[11] = float("nan")
    wg = torch.randn(E, 64, device=DEV, dtype=torch.bfloat16)
    _, _, _, ti = ops.router_fwd(x, wg, K)
    assert int(ti.min()) >= 0 and int(ti.max()) < E and ti[11].tolist() == [0, 1]
    route = ops.route_build(ti, E)
    assert int(route.counts.sum()) == T * K
    # out-of-range ids handed to the public op are clamped instead of corrupting memory
    bad = torch.tensor([[0, 99], [-5, 1]], dtype=torch.int32, device=DEV)
    r2 = ops.route_build(bad, E)
    assert r2.counts.tolist() == [2, 1, 0, 0, 0, 0, 0, 1]


# ------------------------------------------------------------------------------------------------ more than 64 experts
# 64 < E <= 256 (the pretrain plugin's default is -moe.n_experts 128, transformer_lm_mixin.py:32): the router / loss
# kernels with 4 or 8 experts per lane.  Same assertions as the E <= 64 tests above.  Written after the round's GPU budget
# was spent: green on the SIMT emulator (tests/test_simt_kernels.py), first hardware run pending.
WIDE = pytest.mark.first_hw_run


@WIDE
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("T,D,E,K", [(300, 512, 128, 4), (1024, 1024, 128, 8), (257, 256, 200, 8), (128, 64, 256, 2), (64, 128, 65, 3)])
def test_router_matches_oracle_wide(ops, dtype, T, D, E, K):
    test_router_matches_oracle(ops, dtype, T, D, E, K)


@WIDE
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("T,D,E,K", [(77, 64, 8, 2), (50, 128, 64, 8), (90, 128, 128, 8), (41, 64, 256, 5), (33, 64, 97, 1)])
def test_router_from_logits_equals_fused_router(ops, dtype, T, D, E, K):
    """The softmax / top-k half fed with the fused kernel's own (rounded) logits gives the same bits."""
    g = torch.Generator().manual_seed(T + E)
    x = torch.randn(T, D, generator=g).to(dtype).to(DEV)
    wg = (torch.randn(E, D, generator=g) * 0.05).to(dtype).to(DEV)
    logits, probs, tw, ti = ops.router_fwd(x, wg, K)
    p2, w2, i2 = ops.router_from_logits(logits, K)
    assert torch.equal(p2, probs) and torch.equal(w2, tw) and torch.equal(i2, ti)


@WIDE
@pytest.mark.parametrize("B,N,D,E,K", [(1, 131, 128, 8, 1), (2, 90, 64, 8, 2), (1, 70, 128, 128, 1), (1, 64, 64, 4, 4)])
def test_router_backward_reproduces_autograd_through_the_rounded_denominator(ops, B, N, D, E, K):
    """bf16 model: w = topk(p) / sum(topk(p)).to(bf16) (moe.py:131).  Autograd hands the numerator dtw / r in fp32 and the
    denominator -sum_k dtw_k (w_k / r) rounded to bf16; for top-1 the two cancel to the last bf16 digit and what is left
    is that rounding.  Fed with the same dtw the kernel gives autograd's dx and d gate: all but a few elements bit for bit
    (torch's softmax differs from the kernel's in the last fp32 digit), every element within the bf16 band."""
    g = torch.Generator().manual_seed(3 + K)
    x = torch.randn(B * N, D, generator=g).bfloat16()
    wg = (torch.randn(E, D, generator=g) * 0.3).bfloat16()
    dtw = torch.randn(B * N, K, generator=g) * 8
    xr, wr = x.clone().requires_grad_(True), wg.clone().requires_grad_(True)
    w_ref, idx_ref, p_ref, _ = om.router_policy(xr.view(B, N, D), wr, K)
    (w_ref * dtw.view(B, N, K)).sum().backward()
    logits, probs, tw, ti = ops.router_fwd(x.to(DEV), wg.to(DEV), K)
    margin = om.topk_margin(p_ref[0], K) if K < E else torch.ones(N)
    if not torch.equal(ti.cpu().long().view(B, N, K), idx_ref):
        assert bool((margin.min() < 1e-3)), "routing differs on a clear margin"
        pytest.skip("a low-margin token routes differently: the gradients are not comparable element by element")
    dx, dwg = ops.router_bwd(x.to(DEV), wg.to(DEV), probs, tw, ti, B, dtw=dtw.to(DEV))
    for got, ref, nm in ((dx, xr.grad, "dx"), (dwg, wr.grad, "d gate")):
        same = float((got.cpu() == ref).float().mean())
        assert same >= 0.9, f"{nm}: only {same:.3f} of the elements are bit-identical to autograd's"
        assert_close_rms(got, ref, 2e-2, nm)


@WIDE
def test_router_renorm_dtype_is_the_layer_inputs(ops):
    """fp32 inputs under autocast (pretrain plugin): bf16 activations, but the reference's `.to(x.dtype)` is fp32 -- the
    top-k weights sum to one in fp32, a top-1 weight is exactly 1 and carries (up to fp32 rounding) no gradient."""
    g = torch.Generator().manual_seed(11)
    x = torch.randn(200, 64, generator=g).bfloat16().to(DEV)
    wg = (torch.randn(8, 64, generator=g) * 0.3).bfloat16().to(DEV)
    _, probs, w_bf, idx_bf = ops.router_fwd(x, wg, 2)
    _, probs2, w_f, idx_f = ops.router_fwd(x, wg, 2, renorm_dtype=torch.float32)
    assert torch.equal(idx_bf, idx_f) and torch.equal(probs, probs2)
    top = torch.gather(probs, 1, idx_f.long())
    assert torch.equal(w_f, top / top.sum(-1, keepdim=True))
    assert torch.equal(w_bf, top / top.sum(-1, keepdim=True).bfloat16().float()) and not torch.equal(w_bf, w_f)
    _, probs1, w1, i1 = ops.router_fwd(x, wg, 1, renorm_dtype=torch.float32)
    assert bool((w1 == 1).all())
    dtw = torch.randn(200, 1, generator=g).to(DEV)
    dx, dwg = ops.router_bwd(x, wg, probs1, w1, i1, 1, dtw=dtw, renorm_dtype=torch.float32)
    # dtw / p - dtw * (1 / p): zero up to the last fp32 digit of the two quotients (autograd's own expression)
    assert float(dx.float().abs().max()) < 1e-6 and float(dwg.float().abs().max()) < 1e-5


@WIDE
def test_router_tie_break_is_lowest_index_wide(ops):
    x = torch.zeros(8, 64, device=DEV, dtype=torch.bfloat16)
    wg = torch.randn(160, 64, device=DEV, dtype=torch.bfloat16)
    _, probs, tw, ti = ops.router_fwd(x, wg, 5)
    assert ti.cpu().tolist() == [[0, 1, 2, 3, 4]] * 8
    s = torch.zeros(4, 256, device=DEV)
    s[:, [255, 130, 131, 64, 3]] = 1.0                       # equal scores across four slots of several lanes
    _, idx = ops.topk_renorm(s, 5)
    assert idx.cpu().tolist() == [[3, 64, 130, 131, 255]] * 4


@WIDE
@pytest.mark.parametrize("E,K", [(65, 3), (128, 8), (200, 5), (256, 8)])
def test_topk_renorm_wide(ops, E, K):
    test_topk_renorm(ops, E, K)


@WIDE
@pytest.mark.parametrize("B,N,D,E,K", [(3, 100, 128, 128, 4), (2, 300, 64, 200, 8), (1, 70, 64, 72, 2)])
def test_router_aux_and_backward_match_autograd_wide(ops, B, N, D, E, K):
    test_router_aux_and_backward_match_autograd(ops, B, N, D, E, K)


@WIDE
def test_topk_is_total_on_non_finite_scores_wide(ops):
    T, E, K = 16, 128, 4
    scores = torch.randn(T, E, device=DEV)
    scores[3] = float("nan")
    scores[5, 100] = float("nan")
    scores[7, 127] = float("inf")
    scores[9] = float("-inf")
    w, idx = ops.topk_renorm(scores, K)
    assert int(idx.min()) >= 0 and int(idx.max()) < E
    assert all(len(set(r)) == K for r in idx.cpu().tolist())
    assert idx[3].tolist() == [0, 1, 2, 3] and idx[5, 0].item() == 100 and idx[7, 0].item() == 127 and idx[9].tolist() == [0, 1, 2, 3]


@WIDE
@pytest.mark.parametrize("B,N,E,K", [(2, 130, 128, 8), (3, 77, 200, 5), (1, 300, 65, 2), (2, 64, 256, 8)])
def test_compete_losses_kernels_match_torch_wide(B, N, E, K):
    test_compete_losses_kernels_match_torch(B, N, E, K)


@WIDE
@pytest.mark.parametrize("B,N,E", [(2, 200, 128), (3, 96, 256), (1, 130, 100)])
def test_entropy_balance_kernel_matches_pretrain_formula_wide(B, N, E):
    test_entropy_balance_kernel_matches_pretrain_formula(B, N, E)
