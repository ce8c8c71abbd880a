#!/usr/bin/env python
"""bench.py -- CompeteSMoE MoE-layer forward+backward throughput (tokens/s) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): the CompeteSMoE-5.1B-shaped MoE MLP block -- Phi-3.5-mini sized gate_up/down SiLU-GLU
experts, d=3072, ffn=8192, 4 experts, top-2, bf16, 4096 tokens per GPU, synthetic N(0,1) tokens, random-init weights.
One "step" = one forward + backward of the layer over one batch (router step: gate -> top-2 -> sparse experts ->
combine, all gradients).  The competition step (all 4 experts dense + affinity scoring + distillation losses) is timed
in a second region and reported under "competition"; "mix" is the schedule-weighted blend at rate_flip = 0.05.

Prints ONE JSON line (rank 0).  `value` = tokens/s with inputs resident in HBM, CUDA-event timed, max over ranks;
`e2e` = the same through the public nn.Module call with HOST buffers (pinned H2D of tokens and upstream gradient and a
D2H read of the loss inside the timed region); `roofline` = the grouped GEMM's achieved TFLOP/s from per-launch CUDA
events inside the timed region; `cpu_baseline` = the CPU oracle (port of the reference's PyTorch path) timed on this
box's host cores on a bounded sample.  `--impl reference` times only that CPU path.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch
import torch.nn as nn

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

D_MODEL, FFN, N_EXPERTS, TOP_K, TOKENS = 3072, 8192, 4, 2, 4096
RATE_FLIP = 0.05
WORKLOAD = "competesmoe-5.1b-moe-mlp-block d=3072 ffn=8192 E=4 top2 glu-silu bf16 4096 tokens/gpu fwd+bwd"
CPU_SAMPLE_TOKENS = 256


class GLUExpert(nn.Module):
    """Phi3MLP-shaped expert (same parameter names: gate_up_proj.weight [2F, D], down_proj.weight [D, F])."""

    def __init__(self, d, f):
        super().__init__()
        self.gate_up_proj = nn.Linear(d, 2 * f, bias=False)
        self.down_proj = nn.Linear(f, d, bias=False)
        self.activation_fn = nn.SiLU()


def layer_args():
    from types import SimpleNamespace
    return SimpleNamespace(rate_flip=RATE_FLIP, warm_up=0.0, max_compete_in_iter=3, hybrid=False, router_theta=1.0,
                           router_loss_coef=0.01, diversity_loss_coef=0.01, bal_comp_loss_coef=0.01,
                           balance_loss_coef=0.01, router_z_loss_coef=0.001, norm_sigmoid=False, init_weight=True,
                           moe_name="competesmoe")


def flops_per_token(competition: bool) -> float:
    """F_e = 6*D*F (GLU), router R = 2*D*E; step = 3*(K*F_e + R), competition 3*(E*F_e + R)."""
    f_e = 6.0 * D_MODEL * FFN
    r = 2.0 * D_MODEL * N_EXPERTS
    return 3.0 * ((N_EXPERTS if competition else TOP_K) * f_e + r)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock, power and throttle reasons sampled every ~5 ms through NVML while the timed regions run (falls back
    to `nvidia-smi -lms` if the NVML binding is missing).  Reported: median SM clock over samples taken under load."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.rows, self.proc, self.nvml, self._stop = [], None, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = [(getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_slowdown"),
                (getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40), "hw_thermal_slowdown"),
                (getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_thermal_slowdown"),
                (getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4), "sw_power_cap")]
        while not self._stop:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                watts = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                util = n.nvmlDeviceGetUtilizationRates(self.handle).gpu
                try:
                    mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.rows.append((mhz, watts, util, [name for bit, name in bits if mask & bit]))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self._stop = True
            self.thread.join(timeout=1)
            rows = self.rows
            load = [r for r in rows if r[1] >= 0.5 * max(x[1] for x in rows)] if rows else []   # samples under load
            reasons = sorted({name for r in load for name in r[3]})
            return {"sm_mhz": statistics.median(r[0] for r in load) if load else None, "sm_max_mhz": self.max_mhz,
                    "reasons": reasons, "samples": len(rows), "samples_under_load": len(load),
                    "power_w_max": max((r[1] for r in rows), default=None), "source": "nvml, 5 ms period"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(self.NAMES, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 50"}


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_reference_step_time(steps: int, warmup: int, tokens: int = CPU_SAMPLE_TOKENS):
    """The oracle (CPU port of the reference's PyTorch path) on this box's host cores; fp32 like BASELINE configs[0]."""
    from oracle import multimodal as om
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(1235)
    exps = [{"kind": "glu", "act": "silu", "w1": (torch.randn(2 * FFN, D_MODEL, generator=g) * 0.02).requires_grad_(True),
             "w2": (torch.randn(D_MODEL, FFN, generator=g) * 0.02).requires_grad_(True)} for _ in range(N_EXPERTS)]
    gate_w = (torch.randn(N_EXPERTS, D_MODEL, generator=g) * 0.02).requires_grad_(True)
    x = torch.randn(1, tokens, D_MODEL, generator=g).requires_grad_(True)
    dy = torch.randn(1, tokens, D_MODEL, generator=g)
    args = om.default_args()
    times = []
    for i in range(warmup + steps):
        for t in [x, gate_w] + [e[k] for e in exps for k in ("w1", "w2")]:
            t.grad = None
        t0 = time.perf_counter()
        out, aux, _, _, _ = om.competesmoe_forward(x, gate_w, exps, TOP_K, D_MODEL, args, competition=False)
        torch.autograd.backward((out, aux), (dy, torch.ones_like(aux)))
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return statistics.median(times), cores, tokens


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(a.steps, 5)), max(1, min(a.warmup, 2))
    dt, cores, tokens = cpu_reference_step_time(steps, warmup)
    v = tokens / dt
    line = {"impl": "reference", "metric": "moe_layer_fwd_bwd_tokens_per_s", "value": v, "unit": "tokens/s",
            "n_gpus": a.gpus, "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "step": "router", "sample": f"{tokens} of {TOKENS} tokens per step"},
            "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": cores, "kind": "port",
                             "sample": f"{tokens} of {TOKENS} tokens per step, fp32, oracle/multimodal.py"},
            "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def build_layer(device, ep_group=None):
    from competesmoe_b200.multimodal import CompeteSMoE
    torch.manual_seed(0)
    experts = nn.ModuleList([GLUExpert(D_MODEL, FFN) for _ in range(N_EXPERTS)])
    layer = CompeteSMoE(D_MODEL, D_MODEL, N_EXPERTS, TOP_K, experts, layer_args())
    layer = layer.to(device=device, dtype=torch.bfloat16)
    if ep_group is not None:
        layer.enable_expert_parallel(ep_group, max_tokens=TOKENS)   # this rank keeps E / P experts
        torch.cuda.empty_cache()
    layer.total_steps, layer.step_warm = 2, 0
    layer.train()
    return layer


def set_branch(layer, competition: bool):
    layer.prob_flips = torch.full((2,), bool(competition), device=layer.gate.weight.device)
    layer.set_current_steps(0)


def one_step(layer, x, dy, params):
    for p in params:
        p.grad = None
    x.grad = None
    out, aux, _, _ = layer(x)
    torch.autograd.backward((out, aux), (dy, torch.ones_like(aux)))
    return aux


def timed_region(layer, x, dy, params, steps, warmup, dist_on):
    import torch.distributed as dist
    for _ in range(warmup):
        one_step(layer, x, dy, params)
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        one_step(layer, x, dy, params)
    e.record()
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    ms = torch.tensor([s.elapsed_time(e)], device=x.device)
    if dist_on:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms) / steps


def e2e_region(layer, x_host, dy_host, params, steps, warmup, dist_on, device):
    """Public-API call with host buffers: every step's tokens and upstream gradient are copied from pinned host memory
    (H2D) and the loss is read back (D2H) inside the timed region.  The copies of step i+1 are issued on a second
    stream into the other half of a double buffer while step i computes (what a prefetching data loader does); the
    compute stream waits on the copy's event before it touches a buffer, and the copy stream waits until the previous
    user of that buffer is done."""
    import torch.distributed as dist
    main = torch.cuda.current_stream(device)
    copy = torch.cuda.Stream(device)
    bufs = [(torch.empty_like(x_host, device=device).requires_grad_(True), torch.empty_like(dy_host, device=device))
            for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def prefetch(i):
        x_dev, dy_dev = bufs[i % 2]
        with torch.cuda.stream(copy), torch.no_grad():
            copy.wait_event(consumed[i % 2])
            x_dev.copy_(x_host, non_blocking=True)
            dy_dev.copy_(dy_host, non_blocking=True)
            ready[i % 2].record(copy)

    def run(n):
        for e in consumed:
            e.record(main)
        prefetch(0)
        for i in range(n):
            if i + 1 < n:
                prefetch(i + 1)
            x_dev, dy_dev = bufs[i % 2]
            main.wait_event(ready[i % 2])
            aux = one_step(layer, x_dev, dy_dev, params)
            consumed[i % 2].record(main)
            loss_host.copy_(aux.detach().float(), non_blocking=True)

    run(warmup)
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    run(steps)
    e.record()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e)], device=device)
    if dist_on:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    h2d = x_host.numel() * x_host.element_size() + dy_host.numel() * dy_host.element_size()
    return float(ms) / steps, h2d, loss_host.numel() * loss_host.element_size()


def run_ours(a):
    import torch.distributed as dist
    from competesmoe_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist_on = world > 1
    if dist_on:
        dist.init_process_group("nccl", device_id=device)
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak_tf, peak_src = (peaks["bf16_tflops"], "measured (MEASURED_PEAKS.json, burst)") if "bf16_tflops" in peaks else \
        (1590.0, "fallback (B200_PROFILING.md)")

    # N > 1: expert parallelism (north_star stage 6).  The 4 experts are sharded over EP groups of P = min(N, 4) ranks
    # (N = 8: two replicas of an EP4 group); tokens stay data-parallel, 4096 per GPU (weak scaling).
    ep_group, ep_p = None, 1
    if dist_on and a.parallel != "replicas":
        from competesmoe_b200.ep import EPGroup
        ep_p = max(p for p in (1, 2, 4) if p <= world and world % p == 0 and N_EXPERTS % p == 0)
        my_pg = None
        for g0 in range(0, world, ep_p):
            pg = dist.new_group(list(range(g0, g0 + ep_p)))
            if g0 <= rank < g0 + ep_p:
                my_pg = pg
        ep_group = EPGroup(my_pg, device)
    layer = build_layer(device, ep_group)
    params = [p for p in layer.parameters()]
    g = torch.Generator().manual_seed(1235 + rank)
    x_host = torch.randn(1, TOKENS, D_MODEL, generator=g).bfloat16().pin_memory()
    dy_host = torch.randn(1, TOKENS, D_MODEL, generator=g).bfloat16().pin_memory()
    x = x_host.to(device).requires_grad_(True)
    dy = dy_host.to(device)

    # ---- timed region 1: router step, inputs resident in HBM (the headline `value`)
    set_branch(layer, False)
    for _ in range(2):
        one_step(layer, x, dy, params)          # first-call costs (module load, storage fusing) outside any timing
    sampler = ClockSampler(local) if rank == 0 else None
    ops.gemm_timing = []
    launches0 = ops.launch_count
    ms_router = timed_region(layer, x, dy, params, a.steps, a.warmup, dist_on)
    launches = (ops.launch_count - launches0) * a.steps // (a.steps + a.warmup)
    torch.cuda.synchronize()
    timed = ops.gemm_timing[-(len(ops.gemm_timing) * a.steps // (a.steps + a.warmup)):]
    ops.gemm_timing = None
    gemm_ms = [s.elapsed_time(e) for s, e, _, _ in timed]
    gemm_flops = [f for _, _, f, _ in timed]
    gemm_tflops = sum(gemm_flops) / (sum(gemm_ms) * 1e-3) / 1e12 if gemm_ms else 0.0
    gemm_share = sum(gemm_ms) / (ms_router * a.steps) if gemm_ms else 0.0   # of the eager pass the events were taken in
    per_step = len(timed) // max(a.steps, 1)
    gemm_detail = []
    for i in range(per_step):
        ms_i = [gemm_ms[j] for j in range(i, len(timed), per_step)]
        gemm_detail.append({"kind": timed[i][3], "ms": round(statistics.median(ms_i), 4),
                            "tflops": round(timed[i][2] / (statistics.median(ms_i) * 1e-3) / 1e12, 1)})

    # ---- timed region 1b: the same call with the layer's CUDA-graph mode on (layer.enable_cuda_graphs(): forward and
    # backward replayed from captured graphs behind the unchanged nn.Module call).  Not available under expert
    # parallelism.  When it works it is the headline `value`; the eager number stays in the line as "eager".
    ms_graph = None
    if ep_group is None and a.graphs:
        try:
            layer.enable_cuda_graphs()
            one_step(layer, x, dy, params)          # capture (router branch)
            ms_graph = timed_region(layer, x, dy, params, a.steps, a.warmup, dist_on)
        except Exception as exc:   # a capture failure must not cost the eager numbers
            print(f"bench: CUDA-graph mode disabled: {exc}", file=sys.stderr)
            layer.enable_cuda_graphs(False)
            torch.cuda.synchronize()
            ms_graph = None
    ms_eager = ms_router
    used_graphs = ms_graph is not None and ms_graph < ms_eager    # the headline is the faster of the two modes of the same call
    if used_graphs:
        ms_router = ms_graph

    # ---- timed region 2: competition step
    set_branch(layer, True)
    ms_comp = timed_region(layer, x, dy, params, max(2, a.steps // 2), a.warmup, dist_on)
    # ---- timed region 3: end to end through the module with host buffers (router step)
    set_branch(layer, False)
    ms_e2e, h2d, d2h = e2e_region(layer, x_host, dy_host, params, a.steps, a.warmup, dist_on, device)
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        tok = TOKENS * world
        value = tok / (ms_router * 1e-3)
        comp = tok / (ms_comp * 1e-3)
        mix_ms = (1 - RATE_FLIP) * ms_router + RATE_FLIP * ms_comp
        cpu = None
        if world == 1:
            dt, cores, tokens = cpu_reference_step_time(2, 1)
            cpu = {"value": tokens / dt, "unit": "tokens/s", "cores": cores, "kind": "port",
                   "sample": f"{tokens} of {TOKENS} tokens per step, fp32, router step, oracle/multimodal.py"}
        traffic = None
        tf = ROOT / "profiles" / "gemm_traffic.json"
        if tf.exists():
            traffic = json.loads(tf.read_text()).get("dram_bytes_per_launch")
        line = {
            "metric": "moe_layer_fwd_bwd_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_router, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "step": "router", "tokens_per_gpu": TOKENS,
                       "cuda_graphs": used_graphs,
                       "parallelism": "single GPU" if world == 1 else (
                           f"EP{ep_p} x DP{world // ep_p}: experts sharded over NVLink peer memory, tokens data-parallel"
                           if ep_group is not None else f"{world} independent data-parallel replicas"),
                       "l2": "per-step working set (0.6 GB of expert weights + 0.5 GB activations) exceeds the 126 MB L2; no flush"},
            "model_tflops": flops_per_token(False) * tok / (ms_router * 1e-3) / 1e12,
            "model_frac_of_peak": flops_per_token(False) * TOKENS / (ms_router * 1e-3) / 1e12 / peak_tf,
            "competition": {"ms_per_step": ms_comp, "tokens_per_s": comp,
                            "model_tflops": flops_per_token(True) * tok / (ms_comp * 1e-3) / 1e12,
                            "model_frac_of_peak": flops_per_token(True) * TOKENS / (ms_comp * 1e-3) / 1e12 / peak_tf},
            "eager": {"ms_per_step": ms_eager, "tokens_per_s": tok / (ms_eager * 1e-3),
                      "note": "same call without CUDA graphs; the roofline's per-launch events were taken in this pass"},
            "graphed": None if ms_graph is None else {"ms_per_step": ms_graph, "tokens_per_s": tok / (ms_graph * 1e-3),
                                                      "note": "layer.enable_cuda_graphs(): same call replayed from captured graphs"},
            "mix": {"rate_flip": RATE_FLIP, "ms_per_step": mix_ms, "tokens_per_s": tok / (mix_ms * 1e-3)},
            "e2e": {"value": tok / (ms_e2e * 1e-3), "unit": "tokens/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
            "roofline": {"bound": "tensor", "achieved": gemm_tflops, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": gemm_tflops / peak_tf, "traffic": traffic, "kernel": "grouped_gemm_kernel (tcgen05)",
                         "peak_source": peak_src, "launches_per_step": len(timed) // max(a.steps, 1),
                         "share_of_step": gemm_share, "per_launch": gemm_detail},
            "cpu_baseline": cpu, "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if ep_group is not None:
        ep_group.close()
    if dist_on:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--graphs", type=int, default=1, help="1: use the layer's CUDA-graph mode where available (single GPU / replicas)")
    ap.add_argument("--parallel", default="ep", choices=["ep", "replicas"],
                    help="N > 1: expert-parallel groups (default) or N independent replicas of the layer")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
