"""Per-launch timing of the six grouped GEMMs of one router step at the bench shape (C2: d=3072, ffn=8192, E=4, top-2,
4096 tokens -> 8192 routed rows), each timed alone with CUDA events, next to cuBLAS (torch.matmul) on the equivalent
dense problem in the same process (same clocks / power state).  Run on a B200:  python scripts/gemm_bench.py [iters]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from competesmoe_b200 import ops  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
ITERS = int(sys.argv[1]) if len(sys.argv) > 1 else 20
SHAPE = sys.argv[2] if len(sys.argv) > 2 else "c2"
T, K, E, D, F = 4096, 2, 4, 3072, 8192


def timeit(fn, iters=ITERS):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def siglip():
    """C2' / C5: SigLIP-so400m MoE MLP, fc1 [4304, 1152] + bias, GELU-tanh, fc2 [1152, 4304] + bias, 12800 tokens top-2."""
    T, K, E, D, F = 12800, 2, 4, 1152, 4304
    g = torch.Generator().manual_seed(1)
    sel = torch.stack([torch.randperm(E, generator=g)[:K] for _ in range(T)]).int().to(dev)
    route = ops.route_build(sel, E)
    rows, n_rows = route.row_cap, T * K
    print(f"siglip: row_tile={route.row_tile} row_cap={rows} counts={route.counts.tolist()}")
    bf = dict(device=dev, dtype=torch.bfloat16)
    xp, h, dy, z = torch.randn(rows, D, **bf), torch.randn(rows, F, **bf), torch.randn(rows, D, **bf), torch.randn(rows, F, **bf)
    w1, w2 = torch.randn(E, F, D, **bf) * 0.02, torch.randn(E, D, F, **bf) * 0.02
    b1, b2 = torch.randn(E, F, **bf), torch.randn(E, D, **bf)
    fl = 2 * n_rows * D * F
    cases = [
        ("fc1 + bias + GELU-tanh, z and h stored ", lambda: ops.gemm_rows(xp, w1, w_is_kn=False, bias=b1, act=ops.ACT_GELU_TANH, want_preact=True, route=route)),
        ("fc1 + bias + GELU-tanh, h only         ", lambda: ops.gemm_rows(xp, w1, w_is_kn=False, bias=b1, act=ops.ACT_GELU_TANH, route=route)),
        ("fc1 + bias + ReLU, z and h stored      ", lambda: ops.gemm_rows(xp, w1, w_is_kn=False, bias=b1, act=ops.ACT_RELU, want_preact=True, route=route)),
        ("fc1 + bias + ReLU, h only              ", lambda: ops.gemm_rows(xp, w1, w_is_kn=False, bias=b1, act=ops.ACT_RELU, route=route)),
        ("fc1 + bias + SiLU, z and h stored      ", lambda: ops.gemm_rows(xp, w1, w_is_kn=False, bias=b1, act=ops.ACT_SILU, want_preact=True, route=route)),
        ("fc1 + bias only                        ", lambda: ops.gemm_rows(xp, w1, w_is_kn=False, bias=b1, route=route)),
        ("fc1 plain                              ", lambda: ops.gemm_rows(xp, w1, w_is_kn=False, route=route)),
        ("fc2 + bias                             ", lambda: ops.gemm_rows(h, w2, w_is_kn=False, bias=b2, route=route)),
        ("dgrad2 dy.W2                           ", lambda: ops.gemm_rows(dy, w2, w_is_kn=True, route=route)),
        ("dgrad1 dz.W1                           ", lambda: ops.gemm_rows(z, w1, w_is_kn=True, route=route)),
        ("wgrad2 dy^T.h                          ", lambda: ops.gemm_reduce(dy, h, E, route=route, out_dtype=torch.bfloat16)),
        ("wgrad1 dz^T.x                          ", lambda: ops.gemm_reduce(z, xp, E, route=route, out_dtype=torch.bfloat16)),
    ]
    for name, fn in cases:
        ms = timeit(fn)
        print(f"{name}: {ms:.4f} ms  {fl / ms / 1e9:7.1f} TFLOP/s", flush=True)
    for name, fn in [("cuBLAS [25600x1152].[1152x4304]", lambda: torch.matmul(xp[:n_rows], w1[0].t())),
                     ("cuBLAS [25600x4304].[4304x1152]", lambda: torch.matmul(h[:n_rows], w2[0].t()))]:
        ms = timeit(fn)
        print(f"{name}: {ms:.4f} ms  {fl / ms / 1e9:7.1f} TFLOP/s", flush=True)


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    if SHAPE == "siglip":
        return siglip()
    g = torch.Generator().manual_seed(1)
    sel = torch.stack([torch.randperm(E, generator=g)[:K] for _ in range(T)]).int().to(dev)
    route = ops.route_build(sel, E)
    rows = route.row_cap
    print(f"row_tile={route.row_tile} row_cap={rows} counts={route.counts.tolist()}")
    bf = dict(device=dev, dtype=torch.bfloat16)
    xp = torch.randn(rows, D, **bf)
    w1 = torch.randn(E, 2 * F, D, **bf) * 0.02
    w2 = torch.randn(E, D, F, **bf) * 0.02
    h = torch.randn(rows, F, **bf)
    z = torch.randn(rows, 2 * F, **bf)
    dy = torch.randn(rows, D, **bf)
    n_rows = T * K
    cases = [
        ("fwd1 GLU-fused  x.W1^T  k=3072 n=2x8192", lambda: ops.gemm_rows(xp, w1, w_is_kn=False, act=ops.ACT_SILU_GLU, route=route), 2 * n_rows * D * 2 * F),
        ("fwd1 plain      x.W1^T  k=3072 n=16384 ", lambda: ops.gemm_rows(xp, w1, w_is_kn=False, route=route), 2 * n_rows * D * 2 * F),
        ("fwd2            h.W2^T  k=8192 n=3072  ", lambda: ops.gemm_rows(h, w2, w_is_kn=False, route=route), 2 * n_rows * F * D),
        ("wgrad2          dy^T.h  [3072x8192]/e  ", lambda: ops.gemm_reduce(dy, h, E, route=route, out_dtype=torch.bfloat16), 2 * n_rows * F * D),
        ("dgrad2          dy.W2   k=3072 n=8192  ", lambda: ops.gemm_rows(dy, w2, w_is_kn=True, route=route), 2 * n_rows * F * D),
        ("dgrad2 GLU-bwd  dy.W2 * act'(z)        ", lambda: ops.gemm_rows(dy, w2, w_is_kn=True, route=route, act_bwd=ops.ACT_SILU_GLU, aux=z), 2 * n_rows * F * D),
        ("wgrad1          dz^T.x  [16384x3072]/e ", lambda: ops.gemm_reduce(z, xp, E, route=route, out_dtype=torch.bfloat16), 2 * n_rows * D * 2 * F),
        ("dgrad1          dz.W1   k=16384 n=3072 ", lambda: ops.gemm_rows(z, w1, w_is_kn=True, route=route), 2 * n_rows * D * 2 * F),
    ]
    for name, fn, flops in cases:
        ms = timeit(fn)
        print(f"{name}: {ms:.4f} ms  {flops / ms / 1e9:7.1f} TFLOP/s", flush=True)
    # cuBLAS on the equivalent dense problems (one expert's weights, all routed rows)
    xa = xp[:n_rows]
    for name, fn, flops in [
        ("cuBLAS [8192x3072].[3072x16384]", lambda: torch.matmul(xa, w1[0].t()), 2 * n_rows * D * 2 * F),
        ("cuBLAS [8192x8192].[8192x3072] ", lambda: torch.matmul(h[:n_rows], w2[0].t()), 2 * n_rows * F * D),
        ("cuBLAS [8192x3072].[3072x8192] ", lambda: torch.matmul(dy[:n_rows], w2[0]), 2 * n_rows * F * D),
        ("cuBLAS [8192x8192]^3           ", lambda: torch.matmul(h[:8192, :8192], z[:8192, :8192]), 2 * 8192 ** 3),
    ]:
        ms = timeit(fn)
        print(f"{name}: {ms:.4f} ms  {flops / ms / 1e9:7.1f} TFLOP/s", flush=True)
    act_ms = timeit(lambda: ops.act_bwd(z, h, ops.ACT_SILU_GLU))
    print(f"act_bwd GLU standalone: {act_ms:.4f} ms  {rows * F * 2 * 5 / act_ms / 1e6:.0f} GB/s")


if __name__ == "__main__":
    main()
