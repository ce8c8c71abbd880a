"""Achieved HBM bandwidth of the bandwidth-bound kernels of the MoE step (permutation, combine, activation backward,
router, competition tail) at the bench shape C2 and at the sigma-MoE shape C4, each timed alone with CUDA events over
inputs larger than would stay hot in L2 across iterations (a 256 MiB buffer is overwritten between iterations).

    python scripts/hbm_kernels_bench.py            -> markdown table (algorithmic bytes per SURVEY.md 8(d) / time)
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from competesmoe_b200 import ops  # noqa: E402

dev = torch.device("cuda")
PEAK = 6554.9   # MEASURED_PEAKS.json hbm_gbs (copy bandwidth, read + write)
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush_buf.fill_(1)                      # evict the previous iteration's lines from the 126 MB L2
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters * 1e3   # us


def shape(name, T, K, E, D, Dh, glu):
    g = torch.Generator().manual_seed(3)
    bf = dict(device=dev, dtype=torch.bfloat16)
    sel = torch.stack([torch.randperm(E, generator=g)[:K] for _ in range(T)]).int().to(dev)
    w = torch.rand(T, K, generator=g).to(dev)
    x = torch.randn(T, D, generator=g).to(**bf)
    wg = (torch.randn(E, D, generator=g) * 0.02).to(**bf)
    route = ops.route_build(sel, E)
    rows, s = route.row_cap, 2
    y = torch.randn(rows, D, **bf)
    dout = torch.randn(T, D, **bf)
    zc = 2 * Dh if glu else Dh
    z = torch.randn(rows, zc, **bf)
    dh = torch.randn(rows, Dh, **bf)
    n_routed = T * K
    t_pad = (T + 255) // 256 * 256
    y_all = torch.randn(E * t_pad, D, **bf)
    aff_idx = sel
    daff = torch.randn(T, E, device=dev)
    logits, probs, tw, ti = ops.router_fwd(x, wg, K)
    div, inv_norm, sim = ops.diversity_fwd(y_all, aff_idx, T, t_pad)
    gd = torch.tensor(1.0, device=dev)
    act = ops.ACT_SILU_GLU if glu else ops.ACT_RELU
    cases = [
        ("route_build (hist + scan + scatter)", lambda: ops.route_build(sel, E), n_routed * 4 + 3 * n_routed * 4 + n_routed * 8),
        ("router_fwd (gate GEMM + softmax + top-k)", lambda: ops.router_fwd(x, wg, K), T * D * s + T * E * (s + 4) + T * K * 8),
        ("gather_rows (permute)", lambda: ops.gather_rows(x, route), T * D * s + n_routed * D * s + rows * 4),
        ("combine_fwd (gate-weighted, ascending expert)", lambda: ops.combine_fwd(y, route.slot_to_row, route.sel, w, T, K, round_each=True),
         n_routed * D * s + T * D * s + n_routed * 12),
        ("combine_bwd_w", lambda: ops.combine_bwd_w(y, dout, route.slot_to_row, T, K), n_routed * D * s + T * D * s),
        ("gather_rows with slot weights (w * dY)", lambda: ops.gather_rows(dout, route, slot_w=w), T * D * s + n_routed * D * s),
        ("scatter_reduce (dX over k)", lambda: ops.scatter_reduce(y, route.slot_to_row, T, K), n_routed * D * s + T * D * s),
        (f"act_bwd ({'SiLU-GLU' if glu else 'ReLU'})", lambda: ops.act_bwd(z, dh, act, route), n_routed * (Dh + 2 * zc) * s),
        ("affinity_fwd (mean softplus, all experts)", lambda: ops.affinity_fwd(y_all, E, T, t_pad, True), E * T * D * s),
        ("diversity_fwd (K x K cosine per token)", lambda: ops.diversity_fwd(y_all, aff_idx, T, t_pad), n_routed * D * s),
        ("compete_bwd (d dense outputs, 3 sources)", lambda: ops.compete_bwd(y_all, E, T, t_pad, aff_idx, daff=daff, w=w, dout=dout,
                                                                             inv_norm=inv_norm, sim=sim, g_div=gd),
         2 * E * t_pad * D * s + T * D * s),
    ]
    print(f"\n### {name}: T={T} K={K} E={E} D={D} hidden={Dh} ({rows} padded rows)\n")
    print("| kernel | algorithmic MB | us | GB/s | % of measured HBM peak (6555 GB/s) |")
    print("|---|---:|---:|---:|---:|")
    for nm, fn, nbytes in cases:
        us = timeit(fn)
        gbs = nbytes / us / 1e3
        print(f"| {nm} | {nbytes / 1e6:.1f} | {us:.1f} | {gbs:.0f} | {100 * gbs / PEAK:.1f} |", flush=True)


if __name__ == "__main__":
    print(f"# HBM-bound kernels, {torch.cuda.get_device_name(0)}: L2 flushed between iterations, CUDA events, 10 iterations")
    shape("C2 (CompeteSMoE-5.1B MoE MLP block)", 4096, 2, 4, 3072, 8192, True)
    shape("C4 (sigma-MoE pretrain layer, per-GPU share)", 8192, 8, 64, 1024, 128, False)
