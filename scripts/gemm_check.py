"""GPU bring-up check for csmoe_grouped_gemm: every operand-layout variant against torch.matmul, then a timing pass.

Run on a B200:  python scripts/gemm_check.py [--perf]
"""
import ctypes as C
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from competesmoe_b200._lib import BF16, F32, GEMM_REDUCE, GEMM_ROWS, GemmArgs, LIB_PATH  # noqa: E402

lib = C.CDLL(str(LIB_PATH))
lib.csmoe_grouped_gemm.restype = C.c_int32
lib.csmoe_grouped_gemm.argtypes = [C.POINTER(GemmArgs), C.c_void_p]
lib.csmoe_last_error.restype = C.c_char_p

dev = torch.device("cuda")
torch.manual_seed(0)


def run(args):
    rc = lib.csmoe_grouped_gemm(C.byref(args), torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError(f"rc={rc}: {lib.csmoe_last_error().decode()}")


def report(name, got, ref):
    got = got.float()
    ref = ref.float()
    if ref.numel() == 0:
        print(f"[OK] {name}: empty", flush=True)
        return True
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-9
    rel = err.max().item() / denom
    ok = rel < 2e-2
    print(f"[{'OK' if ok else 'FAIL'}] {name}: max_abs_err={err.max().item():.4e} ref_max={denom:.3e} rel={rel:.3e}", flush=True)
    if not ok:
        bad = (err > 2e-2 * denom)
        idx = bad.nonzero()
        print("   n_bad", idx.shape[0], "of", bad.numel(), " first bad:", idx[:8].tolist())
        # coarse map of bad 8x8 blocks (first 128x128 region)
        r = min(128, bad.shape[-2]); c = min(128, bad.shape[-1])
        sub = bad.reshape(-1, bad.shape[-2], bad.shape[-1])[0, :r, :c]
        blk = sub.reshape(r // 8, 8, c // 8, 8).any(1).any(-1)
        for row in blk.int().tolist():
            print("   ", "".join(str(v) for v in row))
    return ok


def make_groups(counts):
    """padded offsets + tile_expert for given per-expert row counts."""
    pad = [0]
    for c in counts:
        pad.append(pad[-1] + (c + 127) // 128 * 128)
    tiles = []
    for e, c in enumerate(counts):
        tiles += [e] * ((c + 127) // 128)
    return pad, tiles


def test_rows(b_layout, counts, n, k, c_dtype=BF16, bias=False, extra_tiles=1):
    E = len(counts)
    pad, tiles = make_groups(counts)
    m = pad[-1] + 128 * extra_tiles
    tiles = tiles + [-1] * extra_tiles
    A = torch.zeros(m, k, device=dev, dtype=torch.bfloat16)
    for e, c in enumerate(counts):
        A[pad[e]:pad[e] + c] = torch.randn(c, k, device=dev).to(torch.bfloat16)
    if b_layout == 0:
        B = (torch.randn(E, n, k, device=dev) / k ** 0.5).to(torch.bfloat16)
    else:
        B = (torch.randn(E, k, n, device=dev) / k ** 0.5).to(torch.bfloat16)
    bias_t = torch.randn(E, n, device=dev).to(torch.bfloat16) if bias else None
    out_dtype = torch.bfloat16 if c_dtype == BF16 else torch.float32
    Cc = torch.full((m, n), 777.0, device=dev, dtype=out_dtype)
    te = torch.tensor(tiles, device=dev, dtype=torch.int32)
    a = GemmArgs()
    a.mode, a.b_layout, a.num_experts, a.dense = GEMM_ROWS, b_layout, E, 0
    a.m, a.n, a.k = m, n, k
    a.a, a.lda = A.data_ptr(), k
    a.b, a.ldb = B.data_ptr(), (k if b_layout == 0 else n)
    a.b_expert_stride = n * k
    a.c, a.ldc, a.c_dtype = Cc.data_ptr(), n, c_dtype
    a.tile_expert = te.data_ptr()
    if bias:
        a.bias, a.bias_dtype = bias_t.data_ptr(), BF16
    run(a)
    torch.cuda.synchronize()
    ok = True
    for e, c in enumerate(counts):
        Ae = A[pad[e]:pad[e] + c].float()
        Be = B[e].float()
        ref = Ae @ (Be.t() if b_layout == 0 else Be)
        if bias:
            ref = ref + bias_t[e].float()
        ok &= report(f"rows b_layout={b_layout} E={E} n={n} k={k} e={e} cnt={c}", Cc[pad[e]:pad[e] + c], ref)
    if extra_tiles:
        untouched = (Cc[pad[-1]:] == 777.0).all().item()
        print("   unused tiles untouched:", untouched)
        ok &= untouched
    return ok


def test_reduce(counts, m_out, n_out, c_dtype=F32):
    E = len(counts)
    pad, _ = make_groups(counts)
    rows = pad[-1]
    A = torch.zeros(rows, m_out, device=dev, dtype=torch.bfloat16)
    B = torch.zeros(rows, n_out, device=dev, dtype=torch.bfloat16)
    for e, c in enumerate(counts):
        A[pad[e]:pad[e] + c] = torch.randn(c, m_out, device=dev).to(torch.bfloat16)
        B[pad[e]:pad[e] + c] = torch.randn(c, n_out, device=dev).to(torch.bfloat16)
    out_dtype = torch.bfloat16 if c_dtype == BF16 else torch.float32
    Cc = torch.full((E, m_out, n_out), 777.0, device=dev, dtype=out_dtype)
    po = torch.tensor(pad, device=dev, dtype=torch.int32)
    a = GemmArgs()
    a.mode, a.num_experts, a.dense = GEMM_REDUCE, E, 0
    a.m, a.n, a.k = m_out, n_out, max(rows, 128)
    a.a, a.lda = A.data_ptr(), m_out
    a.b, a.ldb = B.data_ptr(), n_out
    a.c, a.ldc, a.c_expert_stride, a.c_dtype = Cc.data_ptr(), n_out, m_out * n_out, c_dtype
    a.pad_offsets = po.data_ptr()
    run(a)
    torch.cuda.synchronize()
    ok = True
    for e, c in enumerate(counts):
        ref = A[pad[e]:pad[e] + c].float().t() @ B[pad[e]:pad[e] + c].float()
        ok &= report(f"reduce E={E} m={m_out} n={n_out} e={e} cnt={c}", Cc[e], ref)
    return ok


def test_dense(E, T, n, k):
    """competition layout: shared A [T,k], C [E*T, n]."""
    A = torch.randn(T, k, device=dev).to(torch.bfloat16)
    B = (torch.randn(E, n, k, device=dev) / k ** 0.5).to(torch.bfloat16)
    Cc = torch.zeros(E * T, n, device=dev, dtype=torch.bfloat16)
    a = GemmArgs()
    a.mode, a.b_layout, a.num_experts, a.dense = GEMM_ROWS, 0, E, 1
    a.dense_rows = T
    a.m, a.n, a.k = E * T, n, k
    a.a, a.lda, a.a_expert_rows = A.data_ptr(), k, 0
    a.b, a.ldb, a.b_expert_stride = B.data_ptr(), k, n * k
    a.c, a.ldc, a.c_dtype = Cc.data_ptr(), n, BF16
    run(a)
    torch.cuda.synchronize()
    ref = torch.einsum("tk,enk->etn", A.float(), B.float()).reshape(E * T, n)
    return report(f"dense rows E={E} T={T} n={n} k={k}", Cc, ref)


def perf():
    print("---- perf (C2 shapes: T*K=8192 rows, D=3072, F=8192, E=4)")
    counts = [2048] * 4
    pad, tiles = make_groups(counts)
    m = pad[-1]
    te = torch.tensor(tiles, device=dev, dtype=torch.int32)
    po = torch.tensor(pad, device=dev, dtype=torch.int32)
    D, F2, F = 3072, 16384, 8192
    x = torch.randn(m, D, device=dev).to(torch.bfloat16)
    w1 = (torch.randn(4, F2, D, device=dev) * 0.02).to(torch.bfloat16)
    z = torch.empty(m, F2, device=dev, dtype=torch.bfloat16)
    h = torch.randn(m, F, device=dev).to(torch.bfloat16)
    w2 = (torch.randn(4, D, F, device=dev) * 0.02).to(torch.bfloat16)
    y = torch.empty(m, D, device=dev, dtype=torch.bfloat16)
    dw1 = torch.empty(4, F2, D, device=dev, dtype=torch.bfloat16)

    def gemm_rows(A, B, Cc, n, k, b_layout):
        a = GemmArgs()
        a.mode, a.b_layout, a.num_experts = GEMM_ROWS, b_layout, 4
        a.m, a.n, a.k = m, n, k
        a.a, a.lda = A.data_ptr(), A.shape[1]
        a.b, a.ldb, a.b_expert_stride = B.data_ptr(), B.shape[2], B.shape[1] * B.shape[2]
        a.c, a.ldc, a.c_dtype = Cc.data_ptr(), Cc.shape[1], BF16
        a.tile_expert = te.data_ptr()
        return a

    def gemm_reduce(A, B, Cc):
        a = GemmArgs()
        a.mode, a.num_experts = GEMM_REDUCE, 4
        a.m, a.n, a.k = A.shape[1], B.shape[1], m
        a.a, a.lda = A.data_ptr(), A.shape[1]
        a.b, a.ldb = B.data_ptr(), B.shape[1]
        a.c, a.ldc, a.c_expert_stride, a.c_dtype = Cc.data_ptr(), Cc.shape[2], Cc.shape[1] * Cc.shape[2], BF16
        a.pad_offsets = po.data_ptr()
        return a

    cases = [
        ("fwd1  x@W1^T   [8192x3072]x[3072x16384]", gemm_rows(x, w1, z, F2, D, 0), 2 * m * D * F2),
        ("fwd2  h@W2^T   [8192x8192]x[8192x3072]", gemm_rows(h, w2, y, D, F, 0), 2 * m * F * D),
        ("dgrad2 dy@W2   [8192x3072]x[3072x8192]", gemm_rows(y, w2, h, F, D, 1), 2 * m * F * D),
        ("dgrad1 dz@W1   [8192x16384]x[16384x3072]", gemm_rows(z, w1, y, D, F2, 1), 2 * m * D * F2),
        ("wgrad1 dz^T@x  E x [16384x2048]x[2048x3072]", gemm_reduce(z, x, dw1), 2 * m * D * F2),
    ]
    for name, a, flops in cases:
        for _ in range(3):
            run(a)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        n_it = 10
        for _ in range(n_it):
            run(a)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / n_it
        print(f"{name}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)
    # cuBLAS reference point
    xa = x[:2048]
    wb = w1[0]
    for _ in range(3):
        torch.matmul(xa, wb.t())
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        torch.matmul(xa, wb.t())
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print(f"cuBLAS [2048x3072]x[3072x16384]: {ms:.3f} ms {2 * 2048 * 3072 * 16384 / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    t0 = time.time()
    print(torch.cuda.get_device_name(0))
    ok = True
    ok &= test_rows(0, [128], 256, 64, extra_tiles=0)
    ok &= test_rows(0, [128], 256, 256, extra_tiles=0)
    ok &= test_rows(1, [128], 256, 64, extra_tiles=0)
    ok &= test_rows(1, [128], 256, 256, extra_tiles=0)
    ok &= test_reduce([128], 128, 256)
    ok &= test_reduce([256], 128, 256)
    ok &= test_rows(0, [128], 128, 128, extra_tiles=0)
    ok &= test_rows(1, [128], 128, 128, extra_tiles=0)
    ok &= test_reduce([128], 128, 128)
    ok &= test_rows(0, [300, 0, 77, 513], 4304, 1152, bias=True)
    ok &= test_rows(0, [300, 0, 77, 513], 1152, 4304, c_dtype=F32)
    ok &= test_rows(1, [300, 0, 77, 513], 4304, 1152)
    ok &= test_rows(1, [100, 200, 300], 128, 512)
    ok &= test_reduce([300, 0, 77, 513], 4304, 1152, c_dtype=BF16)
    ok &= test_reduce([300, 0, 77, 513], 1152, 4304)
    ok &= test_reduce([100, 200, 300], 512, 128)
    ok &= test_dense(4, 384, 1152, 512)
    print("ALL OK" if ok else "SOME FAILED", f"({time.time() - t0:.1f}s)")
    if "--perf" in sys.argv:
        perf()
    sys.exit(0 if ok else 1)
