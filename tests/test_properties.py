"""Size-independent properties of the path: permutation maps are a stable bijection, per-expert counts
add up to T*K, combine weights sum to one, the output does not depend on the order of rows inside an expert segment.
hypothesis drives ragged shapes (T = 1, K = E, E > T, empty experts); the CPU half checks the oracle, the GPU half the
CUDA path through the C ABI against the same oracle."""
from types import SimpleNamespace

import pytest
import torch
import torch.nn.functional as F
from hypothesis import given, settings, strategies as st

from oracle import multimodal as om
from oracle import pretrain as op


def _sel(T, K, E, seed, hot):
    g = torch.Generator().manual_seed(seed)
    sel = torch.stack([torch.randperm(E, generator=g)[:K] for _ in range(T)]).int()
    if hot and E > 1:
        sel[sel == E - 1] = 0          # an empty expert and a hot one (duplicates inside a token are legal for the maps)
    return sel


shape = st.tuples(st.integers(1, 97), st.integers(1, 8), st.integers(1, 70), st.integers(0, 10 ** 6), st.booleans()) \
    .filter(lambda s: s[1] <= s[2])


@settings(max_examples=60, deadline=None)
@given(shape)
def test_oracle_maps_are_a_stable_bijection(s):
    T, K, E, seed, hot = s
    sel = _sel(T, K, E, seed, hot)
    r = op.prepare_sel2(sel)
    flat = sel.flatten()
    assert sorted(r.out_index.tolist()) == list(range(T * K))                    # bijection
    ss = r.sel.flatten()
    assert torch.equal(flat[r.out_index], ss) and bool((ss[1:] >= ss[:-1]).all())  # sorted by expert
    same = ss[1:] == ss[:-1]
    assert bool((r.out_index[1:][same] > r.out_index[:-1][same]).all())           # stable inside an expert
    assert torch.equal(r.sel_index, r.out_index // K)
    assert int(torch.bincount(flat.long(), minlength=E).sum()) == T * K


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 64), st.integers(2, 16), st.integers(0, 10 ** 6), st.booleans())
def test_oracle_router_weights_sum_to_one_and_ties_go_to_the_lowest_index(T, E, seed, ties):
    K = min(2, E)
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(T, E, generator=g)
    if ties:
        logits[:, 1] = logits[:, 0]                                               # exact tie between experts 0 and 1
    p = F.softmax(logits, dim=-1)
    w, idx = om.stable_topk(p, K)
    w = w / w.sum(-1, keepdim=True)
    torch.testing.assert_close(w.sum(-1), torch.ones(T), rtol=1e-6, atol=1e-6)
    ref_w, _ = torch.topk(p, K)
    torch.testing.assert_close(w * ref_w.sum(-1, keepdim=True), ref_w)            # same values as torch.topk
    if ties:
        both = (idx == 0).any(-1) & (idx == 1).any(-1)
        pos0 = (idx == 0).float().argmax(-1)
        pos1 = (idx == 1).float().argmax(-1)
        assert bool((pos0[both] < pos1[both]).all())
        only1 = (idx == 1).any(-1) & ~(idx == 0).any(-1)
        assert not bool(only1.any())                                               # never 1 without 0 on an exact tie


@settings(max_examples=25, deadline=None)
@given(st.integers(2, 40), st.integers(1, 4), st.integers(2, 9), st.integers(0, 10 ** 6))
def test_oracle_output_is_invariant_to_row_order_inside_an_expert(T, K, E, seed):
    K = min(K, E)
    g = torch.Generator().manual_seed(seed)
    sel = _sel(T, K, E, seed, False)
    x = torch.randn(T, 16, generator=g)
    keys = torch.randn(E, 16, 8, generator=g)
    values = torch.randn(E, 8, 16, generator=g)
    w = torch.rand(T, K, generator=g)
    ref = op.compute_moe_main(x, sel, w, keys, values, F.relu, torch.float32)
    # an unstable (but valid) sort: reverse the order inside every expert segment
    r = op.prepare_sel2(sel)
    ss = r.sel.flatten()
    perm = torch.arange(T * K)
    for e in range(E):
        seg = (ss == e).nonzero().flatten()
        perm[seg] = seg.flip(0)
    s1 = op.Sel(r.raw_sel, r.sel, r.sel_index[perm], r.out_index[perm], None)
    scores = op.cvmm(x, s1, keys)
    s2 = op.Sel(r.raw_sel, r.sel, r.out_index[perm], None, w)
    out = op.cvmm(F.relu(scores), s2, values)
    torch.testing.assert_close(out, ref, rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------------------------------------ CUDA path
@pytest.mark.gpu
@settings(max_examples=30, deadline=None)
@given(shape, st.sampled_from([128, 256]))
def test_gpu_route_build_properties(s, row_tile):
    from competesmoe_b200 import ops
    T, K, E, seed, hot = s
    sel = _sel(T, K, E, seed, hot)
    r = ops.route_build(sel.cuda(), E, row_tile=row_tile)
    ref = op.prepare_sel2(sel)
    assert int(r.counts.sum()) == T * K
    assert torch.equal(r.sorted_sel.cpu(), ref.sel.flatten()) and torch.equal(r.sort_index.cpu(), ref.out_index)
    s2r, r2s = r.slot_to_row.cpu().long(), r.row_to_slot.cpu().long()
    assert torch.equal(r2s[s2r], torch.arange(T * K)) and int((r2s >= 0).sum()) == T * K
    po = r.pad_offsets.cpu().long()
    assert bool((po % row_tile == 0).all()) and int(po[-1]) <= r.row_cap
    assert torch.equal(torch.bucketize(s2r, po[1:], right=True), sel.flatten().long())   # every slot sits in its expert's segment


@pytest.mark.gpu
@settings(max_examples=12, deadline=None)
@given(st.integers(1, 300), st.sampled_from([(4, 2), (8, 2), (64, 8), (5, 3)]), st.integers(0, 10 ** 6))
def test_gpu_sigma_moe_matches_oracle_on_ragged_shapes(T, ek, seed):
    """compute_moe_main through the public cvmm op on ragged token counts: same selection maps, outputs within the bf16
    tolerance of the oracle, combine weights untouched."""
    from competesmoe_b200.cvmm import cvmm, cvmm_prepare_sel2
    from helpers import assert_close_rms
    E, K = ek
    g = torch.Generator().manual_seed(seed)
    sel = _sel(T, K, E, seed, True)
    x = torch.randn(T, 64, generator=g)
    keys = torch.randn(E, 64, 32, generator=g) / 8
    values = torch.randn(E, 32, 64, generator=g) / 6
    w = torch.rand(T, K, generator=g)
    ref = op.compute_moe_main(x, sel, w, keys, values, F.relu, torch.bfloat16)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        s = cvmm_prepare_sel2(sel.cuda(), n_experts=E)
        scores = cvmm(x.cuda(), s, keys.cuda())
        s2 = s.clone()
        s2.reduction_weight, s2.sel_index, s2.out_index = w.cuda(), s2.out_index, None
        out = cvmm(F.relu(scores), s2, values.cuda())
    assert out.shape == (T, 64)
    assert_close_rms(out, ref, 3e-2, "sigma-MoE output on a ragged shape")
