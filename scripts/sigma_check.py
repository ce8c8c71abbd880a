"""Correctness + timing of the fused sigma-MoE kernels (csrc/sigma_ffn.cu) against plain torch, kernel by kernel.

    python scripts/sigma_check.py [--mode gather|tiled] [--shape small|c4|c1] [--time]

--mode tiled feeds pre-gathered rows through ordinary TMA tile loads (no gather4): isolates the MMA / epilogue logic
from the gather4 addressing.  CSMOE_GATHER4_BOX_ROWS=1|4 selects the tensor-map box height used for gather4.
"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from competesmoe_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="gather")
ap.add_argument("--shape", default="small")
ap.add_argument("--time", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda")
T, D, Dout, E, K = {"small": (1000, 256, 256, 8, 2), "c4": (8192, 1024, 1024, 64, 8), "c1": (4096, 512, 512, 8, 2)}[a.shape]
H = 128
g = torch.Generator().manual_seed(0)
x = torch.randn(T, D, generator=g).bfloat16().to(dev)
dout = torch.randn(T, Dout, generator=g).bfloat16().to(dev)
keys = (torch.randn(E, D, H, generator=g) * D ** -0.5).bfloat16().to(dev)
values = (torch.randn(E, H, Dout, generator=g) * H ** -0.5).bfloat16().to(dev)
bias = (torch.randn(E, H, generator=g) * 0.2).to(dev)
sel = torch.stack([torch.randperm(E, generator=g)[:K] for _ in range(T)]).int().to(dev)
if a.shape == "small":
    sel[sel == 5] = 4          # one expert without tokens
w = torch.rand(T, K, generator=g).to(dev) + 0.1
route = ops.route_build(sel, E, row_tile=128)
r2s = route.row_to_slot.long()
valid = r2s >= 0
tok = torch.where(valid, r2s // K, torch.zeros_like(r2s))
row_e = torch.full((route.row_cap,), -1, device=dev, dtype=torch.long)
row_e[valid] = sel.reshape(-1).long()[r2s[valid]]
po = route.pad_offsets.tolist()
cnt = route.counts.tolist()


def err(got, ref, what):
    got, ref = got.float(), ref.float()
    rms = ref.pow(2).mean().sqrt()
    e = ((got - ref).abs() / (ref.abs() + rms + 1e-30)).max().item()
    print(f"  {what:28s} band err {e:.3e}   (rms {rms.item():.3e})", flush=True)
    return e


def reference():
    xp = torch.zeros(route.row_cap, D, device=dev)
    xp[valid] = x.float()[tok[valid]]
    dyp = torch.zeros(route.row_cap, Dout, device=dev)
    dyp[valid] = dout.float()[tok[valid]]
    wrow = torch.zeros(route.row_cap, device=dev)
    wrow[valid] = w.reshape(-1)[r2s[valid]]
    h = torch.zeros(route.row_cap, H, device=dev)
    y = torch.zeros(route.row_cap, Dout, device=dev)
    dz = torch.zeros(route.row_cap, H, device=dev)
    hw = torch.zeros(route.row_cap, H, device=dev)
    dxr = torch.zeros(route.row_cap, D, device=dev)
    dwrow = torch.zeros(route.row_cap, device=dev)
    dkeys = torch.zeros(E, D, H, device=dev)
    dvalues = torch.zeros(E, H, Dout, device=dev)
    for e in range(E):
        r0, r1 = po[e], po[e] + cnt[e]
        if r1 == r0:
            continue
        he = torch.relu((xp[r0:r1] @ keys[e].float()).bfloat16().float() + bias[e]).bfloat16().float()
        h[r0:r1] = he
        y[r0:r1] = he @ values[e].float()
        dh = dyp[r0:r1] @ values[e].float().T
        dwrow[r0:r1] = (he * dh).sum(-1)
        dze = (wrow[r0:r1, None] * dh * (he > 0)).bfloat16().float()
        dz[r0:r1] = dze
        hwe = (wrow[r0:r1, None] * he).bfloat16().float()
        hw[r0:r1] = hwe
        dxr[r0:r1] = dze @ keys[e].float().T
        dvalues[e] = hwe.T @ dyp[r0:r1]
        dkeys[e] = xp[r0:r1].T @ dze
    return xp, dyp, h, y, dz, hw, dxr, dwrow, dkeys, dvalues


xp_r, dyp_r, h_r, y_r, dz_r, hw_r, dxr_r, dwrow_r, dkeys_r, dvalues_r = reference()
tiled = a.mode == "tiled"
print(f"shape {a.shape}: T={T} D={D} Dout={Dout} E={E} K={K} rows={route.row_cap} mode={a.mode}", flush=True)
worst = 0.0
y, h = ops.sigma_ffn_fwd(x, keys, values, bias, route, xp=xp_r.bfloat16() if tiled else None)
torch.cuda.synchronize()
worst = max(worst, err(h[valid], h_r[valid], "fwd h"), err(y[valid], y_r[valid], "fwd y"))
inside = torch.arange(route.row_cap, device=dev) < po[-1]
pad_rows = h[~valid & inside]
assert pad_rows.numel() == 0 or float(pad_rows.abs().max()) == 0.0, "h padding rows must be zero"
dz, hw, dxr, dwp = ops.sigma_ffn_bwd(dout, keys, values, route, w, h_r.bfloat16(), dyp=dyp_r.bfloat16() if tiled else None)
torch.cuda.synchronize()
dw = dwp.sum(0)
dw_ref = torch.zeros(T * K, device=dev)
dw_ref[r2s[valid]] = dwrow_r[valid]
worst = max(worst, err(dz[valid], dz_r[valid], "bwd dz"), err(hw[valid], hw_r[valid], "bwd hw"),
            err(dxr[valid], dxr_r[valid], "bwd dx rows"), err(dw, dw_ref, "bwd dw"))
if not tiled:
    dv = ops.sigma_wgrad(hw_r.bfloat16(), dout, E, route, transpose=False)
    dk = ops.sigma_wgrad(dz_r.bfloat16(), x, E, route, transpose=True)
    torch.cuda.synchronize()
    worst = max(worst, err(dv, dvalues_r, "wgrad dvalues"), err(dk, dkeys_r, "wgrad dkeys"))
print(f"WORST {worst:.3e} -> {'OK' if worst < 2e-2 else 'MISMATCH'}", flush=True)

if a.time:
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    hb = h_r.bfloat16()
    hwb, dzb = hw_r.bfloat16(), dz_r.bfloat16()

    def timeit(fn, n=10):
        for _ in range(2):
            fn()
        tot = 0.0
        for _ in range(n):
            flush.fill_(1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record()
            torch.cuda.synchronize()
            tot += s.elapsed_time(e)
        return tot / n * 1e3
    rows = T * K
    print(f"  fwd fused        {timeit(lambda: ops.sigma_ffn_fwd(x, keys, values, bias, route)):8.1f} us   "
          f"(writes {(rows * (Dout + H) * 2) / 1e6:.0f} MB)")
    print(f"  bwd fused        {timeit(lambda: ops.sigma_ffn_bwd(dout, keys, values, route, w, hb)):8.1f} us")
    xpt, dypt = xp_r.bfloat16(), dyp_r.bfloat16()
    print(f"  fwd fused, pre-gathered rows (tiled TMA)  {timeit(lambda: ops.sigma_ffn_fwd(x, keys, values, bias, route, xp=xpt)):8.1f} us")
    print(f"  bwd fused, pre-gathered rows (tiled TMA)  {timeit(lambda: ops.sigma_ffn_bwd(dout, keys, values, route, w, hb, dyp=dypt)):8.1f} us")
    print(f"  wgrad dvalues    {timeit(lambda: ops.sigma_wgrad(hwb, dout, E, route, False)):8.1f} us")
    print(f"  wgrad dkeys      {timeit(lambda: ops.sigma_wgrad(dzb, x, E, route, True)):8.1f} us")
    print(f"  combine_fwd      {timeit(lambda: ops.combine_fwd(y, route.slot_to_row, route.sel, w, T, K, round_w=True)):8.1f} us")
    print(f"  scatter_reduce   {timeit(lambda: ops.scatter_reduce(dxr, route.slot_to_row, T, K)):8.1f} us")
    # the unfused pieces they replace
    xpb = ops.gather_rows(x, route)
    print(f"  [unfused] gather_rows      {timeit(lambda: ops.gather_rows(x, route)):8.1f} us")
    print(f"  [unfused] gemm1+relu       {timeit(lambda: ops.gemm_rows(xpb, keys, w_is_kn=True, route=route, act=ops.ACT_RELU, want_preact=True)):8.1f} us")
    print(f"  [unfused] gemm2            {timeit(lambda: ops.gemm_rows(hb, values, w_is_kn=True, route=route)):8.1f} us")
    print(f"  [unfused] wgrad2 (reduce)  {timeit(lambda: ops.gemm_reduce(hb, y, E, route=route)):8.1f} us")
