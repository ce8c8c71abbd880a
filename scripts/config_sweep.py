"""Throughput of the MoE layer at every BASELINE.json config shape (C1, C2', C3, C4, C5), router step and
competition step, fwd+bwd, CUDA-event timed, synthetic inputs resident in HBM.  bench.py covers C2 (the headline).

    python scripts/config_sweep.py [--steps 10] [--only C3]                 # one GPU
    torchrun --nproc-per-node P scripts/config_sweep.py --only C4            # expert parallel over P ranks (C4 / C5)

Prints a markdown table (rank 0).  FLOPs per token follow SURVEY.md 8(d): router 3*(K*F_e + R), competition 3*(E*F_e + R).
"""
from __future__ import annotations

import argparse
import os
import sys
from pathlib import Path
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import MLPExpert  # noqa: E402

PEAK_TF, PEAK_GBS = 1607.8, 6554.9   # MEASURED_PEAKS.json (burst bf16, HBM copy)
EXCHANGE = "auto"                     # pretrain layers under expert parallelism: "auto" / "tokens" / "weights"
GRAPHS = False                        # --graphs: the layers' opt-in CUDA-graph mode (single GPU)


def pretrain_args():
    return SimpleNamespace(warm_up=0.0, rate_flip=0.07, stop_after=100, max_compete_in_iter=16, is_cosine=False,
                           is_norm_weight=False, norm_sigmoid=False, scale_weight=1.0, hybrid=False, tribrid=False,
                           in_topk=False, balance_affinity=False, balance_loss_coef=0.01, balance_loss_coef_comp=0.01,
                           router_loss_coef=0.01, router_theta=1.0, test_only=False)


def mm_args():
    return SimpleNamespace(rate_flip=0.05, warm_up=0.0, max_compete_in_iter=3, hybrid=False, router_theta=1.0,
                           router_loss_coef=0.01, diversity_loss_coef=0.01, bal_comp_loss_coef=0.01,
                           balance_loss_coef=0.01, router_z_loss_coef=0.001, norm_sigmoid=False, init_weight=True,
                           moe_name="competesmoe")


class Case:
    def __init__(self, name, kind, T, D, hidden, E, K, d_out=None, autocast=True, note="", moe_name="competesmoe", key=None):
        self.name, self.kind, self.T, self.D, self.hidden, self.E, self.K = name, kind, T, D, hidden, E, K
        self.key = key or name.split()[0]
        self.moe_name = moe_name
        self.d_out = d_out or D
        self.autocast, self.note = autocast, note

    def f_e(self):
        if self.kind == "pretrain":
            return 4.0 * self.D * self.hidden
        return 2.0 * self.D * self.hidden + 2.0 * self.hidden * self.d_out

    def flops_per_token(self, competition):
        r = 2.0 * self.D * self.E
        if self.moe_name in ("smoe_share", "deepseekv3"):      # k-1 routed choices + the shared expert = k expert passes
            return 3.0 * (self.K * self.f_e() + 2.0 * self.D * (self.E - 1))
        return 3.0 * ((self.E if competition else self.K) * self.f_e() + r)

    def bytes_per_token_router(self, s=2):
        """Unfused algorithmic HBM bytes per token of the router step's expert path: x read + K rows
        written/read around each GEMM + combine; used to classify the sigma-MoE shapes (H = 128) that are HBM bound."""
        K, D, H, Do = self.K, self.D, self.hidden, self.d_out
        fwd = s * (D + K * D + K * D + K * H + K * H + K * Do + K * Do + Do)
        return 3.0 * fwd


def build(case: Case, dev, ep):
    if case.kind == "pretrain":
        from competesmoe_b200.pretrain import CompeteSMoE
        layer = CompeteSMoE(case.D, case.E, case.hidden, n_heads=case.K, args=pretrain_args(), activation=F.relu,
                            selection_mode="gate", log_interval=None).to(dev)
        layer.train()
        layer.regularization_present = True
        layer.step_warm = 0
        if ep is not None:
            layer.enable_expert_parallel(ep, max_tokens=case.T, exchange=EXCHANGE)
        if GRAPHS:
            layer.enable_cuda_graphs()     # under expert parallelism the router step only (the layer decides)

        def set_branch(comp):
            layer.prob_flips_final = {0: torch.full((8,), bool(comp), device=dev)}
            layer.set_current_steps(1)

        one = torch.ones((), device=dev)
        cast = {}

        def step(x, dy):
            # upstream gradient handed straight to autograd (a `(out.float() * dy).sum()` harness adds five [T, D]
            # elementwise / reduction kernels of its own to every step, 7 % of the C4 step)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=case.autocast):
                out = layer(x, id_layer=0)
                regs = list(layer.get_reg_loss().values())
            if cast.get("key") != (id(dy), out.dtype):
                cast["key"], cast["dy"] = (id(dy), out.dtype), dy.to(out.dtype)
            torch.autograd.backward([out] + regs, [cast["dy"]] + [one.to(r.dtype) if r.dtype != one.dtype else one for r in regs])
        x_dtype = torch.float32
    else:
        from competesmoe_b200 import siblings  # noqa: F401  (registers smoe, xmoe, ...)
        from competesmoe_b200.multimodal import get_moe
        CompeteSMoE = get_moe(case.moe_name)
        if case.kind == "siglip":
            experts = nn.ModuleList([MLPExpert(case.D, case.hidden, case.d_out, "gelu_tanh") for _ in range(case.E)])
        else:   # projector: Sequential(Linear, GELU, Linear)
            experts = nn.ModuleList([nn.Sequential(nn.Linear(case.D, case.hidden), nn.GELU(), nn.Linear(case.hidden, case.d_out))
                                     for _ in range(case.E)])
        layer = CompeteSMoE(case.D, case.d_out, case.E, case.K, experts, mm_args()).to(device=dev, dtype=torch.bfloat16)
        layer.total_steps, layer.step_warm = 2, 0
        layer.train()
        if ep is not None:
            layer.enable_expert_parallel(ep, max_tokens=case.T)
        if GRAPHS and hasattr(layer, "enable_cuda_graphs"):
            layer.enable_cuda_graphs()

        def set_branch(comp):
            if case.moe_name == "competesmoe":
                layer.prob_flips = torch.full((2,), bool(comp), device=dev)
                layer.set_current_steps(0)

        def step(x, dy):
            out, aux, _, _ = layer(x)
            torch.autograd.backward((out, aux), (dy.to(out.dtype), torch.ones_like(aux)))
        x_dtype = torch.bfloat16
    return layer, set_branch, step, x_dtype


def time_case(case: Case, dev, ep, steps, warmup, dist_on, graphs=None, exchange=None):
    import torch.distributed as dist
    global GRAPHS, EXCHANGE
    if graphs is not None:
        GRAPHS = graphs
    if exchange is not None:
        EXCHANGE = exchange
    layer, set_branch, step, x_dtype = build(case, dev, ep)
    params = list(layer.parameters())
    rank = ep.rank if ep is not None else 0
    g = torch.Generator().manual_seed(1234 + rank)
    x = torch.randn(1, case.T, case.D, generator=g).to(x_dtype).to(dev).requires_grad_(True)
    dy = torch.randn(1, case.T, case.d_out, generator=g).to(dev)
    res = {}
    for comp in ((False, True) if case.moe_name == "competesmoe" else (False,)):
        set_branch(comp)

        def one():
            for p in params:
                p.grad = None
            x.grad = None
            step(x, dy)
        for _ in range(warmup):
            one()
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize()
        from competesmoe_b200 import ops as _ops
        n0 = _ops.launch_count
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            one()
        e.record()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e) / steps], device=dev)
        if dist_on:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        res[comp] = float(ms)
        # libcsmoe launches issued from Python per step (0 when the step is replayed from captured CUDA graphs)
        if not hasattr(case, "launches"):
            case.launches = {}
        case.launches[comp] = (_ops.launch_count - n0) // max(steps, 1)
    del layer, params, x, dy
    torch.cuda.empty_cache()
    return res


def profile_case(case: Case, dev, ep):
    from torch.profiler import ProfilerActivity, profile
    layer, set_branch, step, x_dtype = build(case, dev, ep)
    params = list(layer.parameters())
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(1, case.T, case.D, generator=g).to(x_dtype).to(dev).requires_grad_(True)
    dy = torch.randn(1, case.T, case.d_out, generator=g).to(dev)
    for comp in (False, True):
        set_branch(comp)

        def one():
            for p in params:
                p.grad = None
            x.grad = None
            step(x, dy)
        for _ in range(3):
            one()
        torch.cuda.synchronize()
        n = 3
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for _ in range(n):
                one()
            torch.cuda.synchronize()
        rows = sorted(((e.key, e.device_time_total / n, e.count / n) for e in prof.key_averages()
                       if getattr(e, "device_time_total", 0) > 0 and "cuda" in str(getattr(e, "device_type", "")).lower()),
                      key=lambda r: -r[1])
        if not ep or ep.rank == 0:
            print(f"## {case.name} -- {'competition' if comp else 'router'} step: {sum(r[2] for r in rows):.0f} launches, "
                  f"{sum(r[1] for r in rows):.1f} us of kernel time per step")
            for k, t, c in rows[:22]:
                print(f"{t:9.1f} us  x{c:5.1f}  {k[:120]}")
    del layer, params, x, dy
    torch.cuda.empty_cache()


def cpu_baseline_rows():
    """the reference's algorithm (the oracle port, fp32) timed on this box's host cores beside the GPU
    numbers -- C1 at its full size (BASELINE configs[0] is this CPU-runnable case) and the SigLIP MoE MLP of C2'/C5 on a
    bounded sample of its tokens.  A reported baseline, not the target."""
    import time
    from oracle import multimodal as om
    from oracle import pretrain as opr
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    rows = []

    def timeit(step, n=3):
        step()
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            step()
            ts.append(time.perf_counter() - t0)
        return sorted(ts)[len(ts) // 2]

    g = torch.Generator().manual_seed(1234)
    D, E, H, K, T = 512, 8, 128, 2, 4096
    wg = (torch.randn(E, D, generator=g) * D ** -0.5).requires_grad_(True)
    keys = (torch.randn(E, D, H, generator=g) * D ** -0.5).requires_grad_(True)
    values = (torch.randn(E, H, D, generator=g) * (E * H) ** -0.5).requires_grad_(True)
    x = torch.randn(8, 512, D, generator=g).requires_grad_(True)
    dy = torch.randn(8, 512, D, generator=g)
    for comp in (False, True):
        def step():
            for t in (wg, keys, values, x):
                t.grad = None
            out, regs, _ = opr.competesmoe_forward(x, wg, keys, values, K, pretrain_args(), comp)
            ((out * dy).sum() + sum(regs.values())).backward()
        dt = timeit(step)
        rows.append(f"| C1 pretrain layer d=512 E=8 K=2 H=128 T=8x512, fp32, CPU oracle port ({cores} threads) | host | "
                    f"{'competition' if comp else 'router'} | {dt * 1e3:.1f} | {T / dt:,.0f} | - | - | - | - |")
    D, F_, E, K, T = 1152, 4304, 4, 2, 640
    exps = []
    for _ in range(E):
        exps.append({"kind": "mlp", "act": "gelu_tanh",
                     "w1": (torch.randn(F_, D, generator=g) * 0.02).requires_grad_(True), "b1": torch.zeros(F_, requires_grad=True),
                     "w2": (torch.randn(D, F_, generator=g) * 0.02).requires_grad_(True), "b2": torch.zeros(D, requires_grad=True)})
    gw = (torch.randn(E, D, generator=g) * 0.02).requires_grad_(True)
    x = torch.randn(1, T, D, generator=g).requires_grad_(True)
    dy = torch.randn(1, T, D, generator=g)

    def step2():
        out, aux, _, _, _ = om.competesmoe_forward(x, gw, exps, K, D, om.default_args(), False)
        ((out * dy).sum() + aux).backward()
    dt = timeit(step2, 2)
    rows.append(f"| C5/C2' SigLIP MoE MLP d=1152 F=4304 E=4 K=2, fp32, CPU oracle port ({cores} threads), {T} of 12800 tokens | host | "
                f"router | {dt * 1e3:.1f} | {T / dt:,.0f} | - | - | - | - |")
    return rows


def cases(world):
    cs = [Case("C1 pretrain layer d=512 E=8 K=2 H=128 T=8x512 (fp32 in, bf16 autocast)", "pretrain", 4096, 512, 128, 8, 2,
               key="C1")]
    for E, K in ((8, 2), (16, 2), (32, 2), (64, 2), (64, 8)):
        cs.append(Case(f"C3 competition sweep d=1024 H=128 E={E} K={K} T=16384", "pretrain", 16384, 1024, 128, E, K,
                       key=f"C3_E{E}_K{K}"))
    cs.append(Case(f"C4 pretrain LM layer d=1024 E=64 K=8 H=128, {65536 // max(world, 1) if world > 1 else 8192} tokens/GPU",
                   "pretrain", 65536 // world if world > 1 else 8192, 1024, 128, 64, 8, key="C4"))
    cs.append(Case("C5/C2' SigLIP MoE MLP d=1152 F=4304 E=4 K=2 gelu-tanh+bias, 12800 tokens/GPU", "siglip", 12800, 1152, 4304, 4, 2,
                   key="C5_siglip"))
    cs.append(Case("C5/C2' projector MoE 2304->3072->3072 E=4 K=2 gelu+bias, 1280 tokens/GPU", "projector", 1280, 2304, 3072, 4, 2,
                   d_out=3072, key="C5_projector"))
    for nm, E, K in (("smoe", 4, 2), ("smoe_sigmoidgating", 4, 2), ("xmoe", 4, 2), ("smoe_perturbed", 4, 2),
                     ("smoe_share", 5, 3), ("deepseekv3", 5, 3)):
        cs.append(Case(f"S sibling router {nm}: SigLIP MoE MLP d=1152 F=4304 E={E} K={K}, 12800 tokens", "siglip", 12800, 1152,
                       4304, E, K, moe_name=nm, key=f"S_{nm}"))
    return cs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--only", default="")
    ap.add_argument("--profile", action="store_true", help="print the per-kernel timeline (torch.profiler) of each selected case")
    ap.add_argument("--graphs", action="store_true", help="enable the layers' CUDA-graph mode (single-GPU cases)")
    ap.add_argument("--cpu-baseline", action="store_true", help="append the CPU oracle port timed on this box (C1, C2' sample)")
    a = ap.parse_args()
    global GRAPHS
    GRAPHS = a.graphs
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ep = None
    if world > 1:
        import torch.distributed as dist
        from competesmoe_b200.ep import EPGroup
        dist.init_process_group("nccl", device_id=dev)
    rows = []
    for case in cases(world):
        if a.only and a.only not in case.name:
            continue
        if (world > 1 and not case.name.startswith(("C4", "C5"))) or (case.name.startswith("S ") and "sibling" not in a.only and not a.only.startswith("S ")):
            continue
        ep = None
        if world > 1:
            import torch.distributed as dist
            from competesmoe_b200.ep import EPGroup
            p = max(q for q in (1, 2, 4, 8, 16) if q <= world and world % q == 0 and case.E % q == 0)
            my = None
            for g0 in range(0, world, p):
                pg = dist.new_group(list(range(g0, g0 + p)))
                if g0 <= rank < g0 + p:
                    my = pg
            ep = EPGroup(my, dev)
        if a.profile:
            profile_case(case, dev, ep)
        res = time_case(case, dev, ep, a.steps, a.warmup, world > 1)
        if ep is not None:
            ep.close()
        rows.append((case, res, ep.world if ep is not None else 1))
    if rank == 0:
        print(f"| config | GPUs (EP) | step | ms/step | tokens/s (all GPUs) | model TFLOP/s per GPU | % of {PEAK_TF:.0f} TF | "
              f"unfused-algorithm GB/s per GPU | % of {PEAK_GBS:.0f} GB/s |")
        print("|---|---|---|---:|---:|---:|---:|---:|---:|")
        for case, res, p in rows:
            for comp in sorted(res):
                ms = res[comp]
                tf = case.flops_per_token(comp) * case.T / (ms * 1e-3) / 1e12
                gbs = case.bytes_per_token_router() * case.T / (ms * 1e-3) / 1e9 if not comp else float("nan")
                print(f"| {case.name} | {world} (EP{p}) | {'competition' if comp else 'router'} | {ms:.3f} | "
                      f"{case.T * world / (ms * 1e-3):,.0f} | {tf:.1f} | {100 * tf / PEAK_TF:.1f} | "
                      f"{gbs:.0f} | {100 * gbs / PEAK_GBS:.1f} |")
        if a.cpu_baseline and world == 1:
            for line in cpu_baseline_rows():
                print(line)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
