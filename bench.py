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
D2H read of the loss, awaited by the host, inside the timed region); `roofline` = the grouped GEMM's achieved TFLOP/s from per-launch CUDA
events inside the timed region; `cpu_baseline` = the reference's CPU path timed on this box's host cores on a bounded sample
(the reference's own modules when /root/reference is importable, else the oracle port -- `kind` says which).
`configs` = the other BASELINE.json shapes (C1, the C3 sweep, C4, C5) at one GPU, router and competition step each;
`hbm_stage` = achieved HBM GB/s of the permute / combine kernels; at N > 1 `ep_parity` (tests/ep_worker.py run on the
job's ranks + a bitwise check of the bench layer against its unsharded copy) and `c4_ep` (configs[3] expert-parallel
over all N ranks next to the unsharded layer on every GPU).  `--impl reference` times only the CPU path, at the
requested --steps / --warmup, on all 4096 tokens of the workload; both arms also time the CPU path of the
pretrain-layer shapes (C1 = BASELINE configs[0], the reference's own CPU-runnable case, and C4) -- `configs.C1/C4.cpu_baseline`.
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
REFERENCE_ROOTS = [os.environ.get("CSMOE_REFERENCE_ROOT", ""), "/root/reference"]


class _GLUForward(GLUExpert):
    """GLUExpert with the Phi3MLP forward (the reference layer deep-copies / calls whatever module it is handed)."""

    def forward(self, x):
        gate, up = self.gate_up_proj(x).chunk(2, dim=-1)
        return self.down_proj(up * self.activation_fn(gate))


def _reference_layer():
    """The UNMODIFIED reference CompeteSMoE (moe_model/model/moe/competesmoe.py:9) when its tree is importable on this
    machine (the build container; it does not travel to the GPU box), else None."""
    import contextlib
    import importlib
    import io
    for root in REFERENCE_ROOTS:
        if root and (Path(root) / "moe_model" / "model" / "moe" / "competesmoe.py").exists():
            try:
                sys.path.insert(0, root)
                with contextlib.redirect_stdout(io.StringIO()):
                    importlib.import_module("moe_model.model.moe")
                    reg = importlib.import_module("moe_model.model.moe.register")
                return reg.get_moe("competesmoe"), root
            except Exception as exc:       # missing optional dependency etc.: fall back to the port, say why
                print(f"bench: reference at {root} not importable ({exc!r}); timing the oracle port", file=sys.stderr)
            finally:
                if sys.path and sys.path[0] == root:
                    sys.path.pop(0)
    return None, None


def cpu_reference_step_time(steps: int, warmup: int, tokens: int = TOKENS, device: str = "cpu"):
    """One router step (fwd + bwd) of the workload on the host cores, fp32 like BASELINE configs[0]: the reference's own
    module when importable, else the oracle (CPU port of the same PyTorch path).  Returns (median s/step, cores, tokens,
    kind, source).  device="cuda" runs the same eager code on the GPU (informational: the reference's GPU path)."""
    from oracle import multimodal as om
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dt = torch.float32 if device == "cpu" else torch.bfloat16
    g = torch.Generator().manual_seed(1235)
    x = torch.randn(1, tokens, D_MODEL, generator=g).to(device, dt).requires_grad_(True)
    dy = torch.randn(1, tokens, D_MODEL, generator=g).to(device, dt)
    ref_cls, root = _reference_layer() if device == "cpu" else (None, None)
    if ref_cls is not None:
        import contextlib
        import io
        torch.manual_seed(0)
        experts = nn.ModuleList([_GLUForward(D_MODEL, FFN) for _ in range(N_EXPERTS)])
        with contextlib.redirect_stdout(io.StringIO()):     # the reference's constructor prints; stdout is the JSON line's
            layer = ref_cls(in_embed_dim=D_MODEL, out_embed_dim=D_MODEL, num_of_experts=N_EXPERTS, num_selected=TOP_K,
                            expert=experts, args=layer_args())
        layer.total_steps, layer.step_warm, layer.current_steps = 2, 0, 0
        layer.prob_flips = torch.zeros(2)
        layer.train()
        leaves = list(layer.parameters())

        def step():
            out, aux, _, _ = layer(x)
            torch.autograd.backward((out, aux), (dy, torch.ones_like(aux)))
        kind, source = "reference", f"{root}/moe_model/model/moe/competesmoe.py (unmodified)"
    else:
        exps = [{"kind": "glu", "act": "silu",
                 "w1": (torch.randn(2 * FFN, D_MODEL, generator=g) * 0.02).to(device, dt).requires_grad_(True),
                 "w2": (torch.randn(D_MODEL, FFN, generator=g) * 0.02).to(device, dt).requires_grad_(True)} for _ in range(N_EXPERTS)]
        gate_w = (torch.randn(N_EXPERTS, D_MODEL, generator=g) * 0.02).to(device, dt).requires_grad_(True)
        leaves = [gate_w] + [e[k] for e in exps for k in ("w1", "w2")]
        args = om.default_args()

        def step():
            out, aux, _, _, _ = om.competesmoe_forward(x, gate_w, exps, TOP_K, D_MODEL, args, competition=False)
            torch.autograd.backward((out, aux), (dy, torch.ones_like(aux)))
        kind, source = "port", "oracle/multimodal.py"
    times = []
    for i in range(warmup + steps):
        for t in [x] + leaves:
            t.grad = None
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()
        if device != "cpu":
            torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return statistics.median(times), cores, tokens, kind, source


# the language-pretraining layer shapes among BASELINE.json's configs
PRETRAIN_SHAPES = {
    "C1": dict(what="configs[0]: pretrain layer d=512, 8 experts top-2, expert size 128, batch 8 x seq 512", B=8, N=512, D=512, E=8, K=2, H=128),
    "C4": dict(what="configs[3]: pretrain LM layer d=1024, 64 experts top-8, expert size 128, 8 x 1024 tokens per GPU", B=8, N=1024, D=1024, E=64, K=8,
               H=128),
}


def pretrain_port_step_time(key: str, steps: int, warmup: int, device: str = "cpu"):
    """One router step (fwd + bwd, regularisers included) of the language-pretraining CompeteSMoE layer
    (moe_pretrain_model/layers/moe/competesmoe.py:524) at a named shape.  The reference computes this layer through its
    Triton CVMM kernels, which have no CPU path; its CPU-runnable statement is the per-expert-loop form of the same
    algebra, i.e. oracle/pretrain.py -- `kind` is always "port" here.  fp32 on the CPU; device="cuda" runs
    the same eager code on the GPU under the reference's bf16 autocast convention (informational).
    Returns (median s/step, tokens, cores)."""
    from oracle import pretrain as opr
    sh = PRETRAIN_SHAPES[key]
    B, N, D, E, K, H = (sh[k] for k in ("B", "N", "D", "E", "K", "H"))
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(1234)
    mk = lambda *shape, std=1.0: (torch.randn(*shape, generator=g) * std).to(device).requires_grad_(True)  # noqa: E731
    w_gate, keys, values = mk(E, D, std=D ** -0.5), mk(E, D, H, std=D ** -0.5), mk(E, H, D, std=(E * H) ** -0.5)
    x = mk(B, N, D)
    dy = torch.randn(B, N, D, generator=g).to(device)
    args = opr.default_args()
    op_dtype = torch.float32 if device == "cpu" else torch.bfloat16
    times = []
    for i in range(warmup + steps):
        for t in (x, w_gate, keys, values):
            t.grad = None
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        out, regs, _ = opr.competesmoe_forward(x, w_gate, keys, values, K, args, False, op_dtype=op_dtype)
        ((out.float() * dy).sum() + sum(r.float() for r in regs.values())).backward()
        if device != "cpu":
            torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return statistics.median(times), B * N, cores


def pretrain_port_entries(steps: int = 3, warmup: int = 1, gpu: bool = False) -> dict:
    """`cpu_baseline` (and, with gpu=True, `gpu_eager_loop`) of the pretrain-layer shapes, keyed like `configs`."""
    out = {}
    for key, sh in PRETRAIN_SHAPES.items():
        entry = {"what": sh["what"]}
        try:
            dt, tokens, cores = pretrain_port_step_time(key, steps, warmup)
            entry["cpu_baseline"] = {"value": tokens / dt, "unit": "tokens/s", "ms_per_step": dt * 1e3, "cores": cores, "kind": "port",
                                     "sample": f"{steps} router steps (after {warmup} warm-up) over all {tokens} tokens, fp32, "
                                               "oracle/pretrain.py (per-expert-loop CVMM: the reference's Triton CVMM has no CPU path)"}
        except Exception as exc:
            entry["cpu_baseline"] = {"unavailable": repr(exc)[:300]}
        if gpu:
            try:
                gdt, tokens, _ = pretrain_port_step_time(key, max(steps, 5), 2, device="cuda")
                entry["gpu_eager_loop"] = {"value": tokens / gdt, "unit": "tokens/s", "ms_per_step": gdt * 1e3, "dtype": "bf16",
                                           "what": "oracle/pretrain.py run on cuda:0 (eager PyTorch, one matmul per expert, none of "
                                                   "this repo's kernels)"}
            except Exception as exc:
                entry["gpu_eager_loop"] = {"unavailable": repr(exc)[:300]}
        out[key] = entry
    return out


def run_reference_arm(a):
    """The reference's CPU implementation of the path on this box's host cores: every one of the requested --steps
    (after --warmup) is one fwd + bwd router step over ALL 4096 tokens of the workload (about 2-3 s each on 16 cores)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, a.steps), max(1, a.warmup)
    dt, cores, tokens, kind, source = cpu_reference_step_time(steps, warmup)
    v = tokens / dt
    sample = f"all {tokens} tokens of the workload per step, fp32, router step, {source}"
    line = {"impl": "reference", "metric": "moe_layer_fwd_bwd_tokens_per_s", "value": v, "unit": "tokens/s",
            "n_gpus": a.gpus, "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "step": "router", "tokens_per_gpu": TOKENS,
                       "note": "CPU arm: one process on the host cores whatever --gpus says"},
            "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if torch.cuda.is_available():
        # informational: the same eager per-expert loop (torch.where + cuBLAS per expert, one host
        # sync per expert) on the B200 in bf16 -- what the reference's GPU path does at this shape
        try:
            gdt, _, _, _, gsrc = cpu_reference_step_time(max(3, min(steps, 10)), 2, device="cuda")
            line["gpu_eager_loop"] = {"value": tokens / gdt, "unit": "tokens/s", "ms_per_step": gdt * 1e3, "dtype": "bf16",
                                      "what": f"{gsrc} run on cuda:0 (eager PyTorch / cuBLAS, none of this repo's kernels)"}
        except Exception as exc:
            line["gpu_eager_loop"] = {"unavailable": repr(exc)}
    # the pretrain-layer shapes of BASELINE.json (configs[0] is the reference's own CPU-runnable case) beside the headline
    line["configs"] = pretrain_port_entries(steps=3, warmup=1, gpu=torch.cuda.is_available())
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
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    landed = [torch.cuda.Event() for _ in range(2)]
    seen = []

    def prefetch(i):
        x_dev, dy_dev = bufs[i % 2]
        with torch.cuda.stream(copy), torch.no_grad():
            copy.wait_event(consumed[i % 2])
            x_dev.copy_(x_host, non_blocking=True)
            dy_dev.copy_(dy_host, non_blocking=True)
            ready[i % 2].record(copy)

    def read_loss(i):
        # the host WAITS for step i's loss to land in pinned memory and reads it, one step behind the step it has just
        # issued (a training loop's logger): every step's result reaches the host inside the timed region
        landed[i % 2].synchronize()
        seen.append(float(loss_host[i % 2]))

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
            loss_host[i % 2].copy_(aux.detach().float(), non_blocking=True)
            landed[i % 2].record(main)
            if i >= 1:
                read_loss(i - 1)
        if n >= 1:
            read_loss(n - 1)

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
    finite = [v == v and abs(v) != float("inf") for v in seen]
    reads = {"losses_read_on_host": len(seen) - warmup, "last_loss": seen[-1] if seen and finite[-1] else None,
             "all_finite": all(finite)}
    return float(ms) / steps, h2d, loss_host[0].numel() * loss_host[0].element_size(), reads


# ------------------------------------------------------------------------------------------------ extra sections
def hbm_stage_gbs(device, peak_gbs):
    """Achieved HBM bandwidth of the permutation / combine kernels (north_star stages 3 and 5) at the bench shape, each
    timed alone with CUDA events, a 256 MiB buffer overwritten between iterations (L2 flush).  Algorithmic bytes per
    SURVEY.md 8(d): gather reads T*D*s + the map, writes T*K*D*s; combine reads T*K*D*s + maps, writes T*D*s."""
    from competesmoe_b200 import ops
    T, K, E, D, s = TOKENS, TOP_K, N_EXPERTS, D_MODEL, 2
    g = torch.Generator().manual_seed(3)
    sel = torch.stack([torch.randperm(E, generator=g)[:K] for _ in range(T)]).int().to(device)
    w = torch.rand(T, K, generator=g).to(device)
    x = torch.randn(T, D, generator=g).to(device, torch.bfloat16)
    route = ops.route_build(sel, E)
    y = torch.randn(route.row_cap, D, device=device, dtype=torch.bfloat16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    cases = {
        "gather_rows": (lambda: ops.gather_rows(x, route), T * D * s + T * K * D * s + route.row_cap * 4),
        "combine_fwd": (lambda: ops.combine_fwd(y, route.slot_to_row, route.sel, w, T, K, round_each=True),
                        T * K * D * s + T * D * s + T * K * 12),
        "scatter_reduce": (lambda: ops.scatter_reduce(y, route.slot_to_row, T, K), T * K * D * s + T * D * s),
    }
    out = {}
    for name, (fn, nbytes) in cases.items():
        for _ in range(2):
            fn()
        tot = 0.0
        for _ in range(8):
            flush.fill_(1)
            s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            fn()
            e0.record()
            torch.cuda.synchronize()
            tot += s0.elapsed_time(e0)
        us = tot / 8 * 1e3
        out[name] = {"us": round(us, 2), "algorithmic_bytes": nbytes, "gbs": round(nbytes / us / 1e3, 1),
                     "frac_of_hbm_peak": round(nbytes / us / 1e3 / peak_gbs, 4)}
    del flush
    torch.cuda.empty_cache()
    return out


def named_configs(device, steps, peak_tf, peak_gbs):
    """BASELINE.json configs[0], [2], [3], [4] at one GPU (shapes: SURVEY.md section 8 table), router and competition
    step, fwd + bwd, inputs resident in HBM, CUDA events; the layers' CUDA-graph mode (`graphed_ms`) and the plain call
    (`eager_ms`, with the number of libcsmoe launches per step).  Fractions: model FLOPs per SURVEY.md 8(d) against the
    bf16 peak; for the sigma-MoE shapes (H = 128, HBM-bound) the unfused algorithmic bytes of the expert path against the
    HBM peak."""
    sys.path.insert(0, str(ROOT / "scripts"))
    import config_sweep as cs
    out = {}
    for case in cs.cases(1):
        if case.key.startswith("S_"):
            continue
        try:
            case.launches = {}
            eager = cs.time_case(case, device, None, max(3, steps // 4), 2, False, graphs=False)
            launches = dict(case.launches)
            graphed = cs.time_case(case, device, None, steps, 3, False, graphs=True)
        except Exception as exc:   # one config failing must not cost the headline line
            out[case.key] = {"error": repr(exc)}
            torch.cuda.synchronize()
            continue
        entry = {"what": case.name, "tokens": case.T}
        for comp, nm in ((False, "router"), (True, "competition")):
            ms = min(graphed[comp], eager[comp])
            tf = case.flops_per_token(comp) * case.T / (ms * 1e-3) / 1e12
            e = {"ms_per_step": round(ms, 4), "graphed_ms": round(graphed[comp], 4), "eager_ms": round(eager[comp], 4),
                 "launches_per_step": launches.get(comp), "tokens_per_s": round(case.T / (ms * 1e-3), 1),
                 "model_tflops": round(tf, 2), "frac_of_bf16_peak": round(tf / peak_tf, 4)}
            if not comp and case.kind == "pretrain":
                gbs = case.bytes_per_token_router() * case.T / (ms * 1e-3) / 1e9
                e.update(unfused_algorithmic_gbs=round(gbs, 1), frac_of_hbm_peak=round(gbs / peak_gbs, 4))
            entry[nm] = e
        # schedule-weighted blend: rate_flip 0.07 in the pretraining sweeps' yaml, 0.05 the multimodal default
        flip = 0.07 if case.kind == "pretrain" else 0.05
        mix = (1 - flip) * entry["router"]["ms_per_step"] + flip * entry["competition"]["ms_per_step"]
        entry["mix"] = {"rate_flip": flip, "ms_per_step": round(mix, 4), "tokens_per_s": round(case.T / (mix * 1e-3), 1)}
        out[case.key] = entry
    return out


def ep_parity_section(device, world, rank, bench_layer, x, dy, params):
    """Expert-parallel parity on the ranks of THIS job: (1) tests/ep_worker.py (outputs, dx, every expert / gate gradient
    of sharded vs unsharded layers, both plugins, both steps, ragged case) on a world-wide EP group; (2) the bench layer
    itself against an unsharded copy of the same full layer on this rank's tokens -- the router step's output and dx do
    not depend on where a row is computed, so they must agree bit for bit.  Any failure is reported, not swallowed."""
    import torch.distributed as dist
    res = {"world": world, "ok": True, "cases": []}
    try:
        sys.path.insert(0, str(ROOT / "tests"))
        import ep_worker as ew
        from competesmoe_b200.ep import EPGroup
        group = EPGroup(None, device)
        try:
            for comp in (False, True):
                ew.run_multimodal(group, device, kind="mlp", E=4 if world <= 4 else 8, K=2, D=256, Fh=520, B=2, N=200, competition=comp)
                ew.run_multimodal(group, device, kind="glu", E=8, K=2, D=512, Fh=1024, B=1, N=1000, competition=comp)
                ew.run_pretrain(group, device, E=16, K=4, D=256, H=128, B=2, N=300, competition=comp, exchange="weights")
                ew.run_pretrain(group, device, E=16, K=2, D=256, H=128, B=1, N=500, competition=comp, exchange="weights", bias=True)
                ew.run_pretrain(group, device, E=16, K=4, D=256, H=128, B=2, N=300, competition=comp)
                res["cases"] += [f"multimodal mlp {'comp' if comp else 'router'}", f"multimodal glu {'comp' if comp else 'router'}",
                                 f"pretrain E=16 K=4 {'comp' if comp else 'router'} (weights exchanged)",
                                 f"pretrain E=16 K=2 bias {'comp' if comp else 'router'} (weights exchanged)",
                                 f"pretrain E=16 K=4 {'comp' if comp else 'router'} (tokens exchanged)"]
            ew.run_multimodal(group, device, kind="mlp", E=8, K=1, D=128, Fh=256, B=1, N=3 + 5 * group.rank, competition=False,
                              max_tokens=3 + 5 * (group.world - 1))
            res["cases"].append("ragged top-1")
        finally:
            group.close()
        # (2) the bench layer against its unsharded twin (same seed -> same full weights)
        local = build_layer(device, None)
        set_branch(local, False)
        set_branch(bench_layer, False)
        xl = x.detach().clone().requires_grad_(True)
        for p in local.parameters():
            p.grad = None
        out_l, aux_l, _, _ = local(xl)
        torch.autograd.backward((out_l, aux_l), (dy, torch.ones_like(aux_l)))
        for p in params:
            p.grad = None
        x.grad = None
        out_e, aux_e, _, _ = bench_layer(x)
        torch.autograd.backward((out_e, aux_e), (dy, torch.ones_like(aux_e)))
        same = bool(torch.equal(out_e, out_l)) and bool(torch.equal(x.grad, xl.grad)) and \
            bool(torch.equal(bench_layer.last_routing[0], local.last_routing[0]))
        flag = torch.tensor([0 if same else 1], device=device)
        dist.all_reduce(flag)
        res["bench_layer_bitwise"] = int(flag) == 0
        res["ok"] = res["ok"] and res["bench_layer_bitwise"]
        del local, out_l, aux_l, xl
        torch.cuda.empty_cache()
    except Exception as exc:
        res["ok"] = False
        res["error"] = repr(exc)[:500]
        print(f"bench: rank {rank}: EP parity section failed: {res['error']}", file=sys.stderr, flush=True)
    if rank == 0:
        print(f"EP parity {'ok' if res['ok'] else 'FAILED'} world={world}: {len(res['cases'])} ep_worker cases"
              f"{', bench layer bitwise equal to its unsharded copy' if res.get('bench_layer_bitwise') else ''}"
              f"{' -- ' + res['error'] if 'error' in res else ''}", file=sys.stderr, flush=True)
    return res


def c4_ep_section(device, world, rank, steps, out):
    """BASELINE.json configs[3] (d=1024, H=128, 64 experts, top-8, bf16 autocast): 8192 tokens per GPU (N = 8 gives the
    yaml's global batch of 64 x 1024), expert-parallel over all N ranks, next to the UNSHARDED layer running the same
    per-GPU batch on every GPU at the same time (the 1-GPU program under the same power conditions).
    efficiency = local_ms / ep_ms = tokens/s(N) / (N * tokens/s(1)).  Both exchange modes of the pretrain layer are
    timed: "weights" (owners publish bf16 expert copies, gradients reduced onto the owners; what "auto" picks at this
    shape) and then "tokens" (every (token, expert) row travels to the expert's owner and back).  `out` is filled as the
    parts complete, so that a hang in a later part leaves the earlier numbers in the line."""
    sys.path.insert(0, str(ROOT / "scripts"))
    import config_sweep as cs
    from competesmoe_b200.ep import EPGroup, WeightExchange
    case = cs.Case("C4 pretrain LM layer d=1024 E=64 K=8 H=128, 8192 tokens/GPU", "pretrain", 8192, 1024, 128, 64, 8, key="C4")
    auto = "weights" if WeightExchange.prefer_weights(case.E, 2 * case.D * case.hidden, case.T, case.K, case.D, case.D) else "tokens"
    out.update({"what": case.name, "tokens_per_gpu": case.T, "world": world, "exchange_auto": auto})
    names = ((False, "router"), (True, "competition"))
    try:
        local_eager = cs.time_case(case, device, None, steps, 3, True, graphs=False)
        local_graph = cs.time_case(case, device, None, steps, 3, True, graphs=True)
        for comp, nm in names:
            out[nm] = {"local_eager_ms": round(local_eager[comp], 4), "local_graphed_ms": round(local_graph[comp], 4)}
        group = EPGroup(None, device)
        try:
            for mode in ((auto,) + tuple(m for m in ("weights", "tokens") if m != auto)):
                ep = cs.time_case(case, device, group, steps, 3, True, graphs=False, exchange=mode)
                try:
                    ep_graph = cs.time_case(case, device, group, steps, 3, True, graphs=True, exchange=mode)
                except Exception as exc:
                    ep_graph = {False: float("nan"), True: float("nan")}
                    out[f"ep_graph_error_{mode}"] = repr(exc)[:300]
                for comp, nm in names:
                    best_local = min(local_eager[comp], local_graph[comp])
                    best_ep = min(v for v in (ep[comp], ep_graph[comp]) if v == v)
                    row = {"ep_ms": round(best_ep, 4), "ep_eager_ms": round(ep[comp], 4),
                           "ep_graphed_ms": None if ep_graph[comp] != ep_graph[comp] else round(ep_graph[comp], 4),
                           "tokens_per_s": round(case.T * world / (best_ep * 1e-3), 1),
                           "efficiency_vs_local_best": round(best_local / best_ep, 4),
                           "efficiency_vs_local_eager": round(local_eager[comp] / ep[comp], 4)}
                    out[nm][mode] = row
                    if mode == auto:
                        out[nm].update({"ep_ms": row["ep_ms"], "tokens_per_s": row["tokens_per_s"],
                                        "efficiency_vs_local_best": row["efficiency_vs_local_best"]})
        finally:
            cs.EXCHANGE = "auto"
            group.close()
    except Exception as exc:
        out["error"] = repr(exc)[:500]
    return out


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
    peak_gbs = float(peaks.get("hbm_gbs", 6550.0))

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
                            "ms_mean": round(statistics.fmean(ms_i), 4), "ms_max": round(max(ms_i), 4),
                            "tflops": round(timed[i][2] / (statistics.median(ms_i) * 1e-3) / 1e12, 1)})
    # the same ratio on the per-launch MEDIANS: the mean (the contract's `achieved`) carries every power-cap dip and
    # host hiccup of the eager pass, the median does not; both are in the line
    med_ms = sum(d["ms"] for d in gemm_detail)
    gemm_tflops_median = sum(timed[i][2] for i in range(per_step)) / (med_ms * 1e-3) / 1e12 if med_ms > 0 else 0.0

    # ---- timed region 1b: the same call with the layer's CUDA-graph mode on (layer.enable_cuda_graphs(): forward and
    # backward replayed from captured graphs behind the unchanged nn.Module call).  Under expert parallelism the router
    # step is captured too (the exchange kernels and the device-side barriers capture like any launch); the competition
    # step of the multimodal layer stays eager there (NCCL all-gather of the expert weights).  When the graphed call is
    # faster it is the headline `value`; the eager number stays in the line as "eager".
    ms_graph = None
    if a.graphs:
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
    ms_e2e, h2d, d2h, e2e_reads = e2e_region(layer, x_host, dy_host, params, a.steps, a.warmup, dist_on, device)
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        tok = TOKENS * world
        value = tok / (ms_router * 1e-3)
        comp = tok / (ms_comp * 1e-3)
        mix_ms = (1 - RATE_FLIP) * ms_router + RATE_FLIP * ms_comp
        cpu = None
        if world == 1:
            dt, cores, tokens, kind, source = cpu_reference_step_time(3, 1)
            cpu = {"value": tokens / dt, "unit": "tokens/s", "cores": cores, "kind": kind,
                   "sample": f"3 steps (after 1 warm-up) over all {tokens} tokens of the workload, fp32, router step, {source}"}
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
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e, **e2e_reads,
                    "note": "H2D of the tokens and the upstream gradient from pinned memory every step (next step's copy "
                            "overlapped on a second stream); the host waits for and reads every step's loss, one step behind "
                            "the step it has just issued"},
            "roofline": {"bound": "tensor", "achieved": gemm_tflops, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": gemm_tflops / peak_tf, "achieved_median": gemm_tflops_median,
                         "frac_median": gemm_tflops_median / peak_tf, "traffic": traffic, "kernel": "grouped_gemm_kernel (tcgen05)",
                         "peak_source": peak_src, "launches_per_step": len(timed) // max(a.steps, 1),
                         "share_of_step": gemm_share, "per_launch": gemm_detail},
            "cpu_baseline": cpu, "gpu_launches": launches, "clocks": clocks,
            "hbm_stage": None, "configs": None, "ep_parity": None, "c4_ep": None,
        }
        if traffic is not None:
            line["roofline"]["traffic_source"] = ("profiles/gemm_traffic.json (ncu --set full of the same six launches, mean per launch, committed; not re-measured "
                                                  "in this run; captured before the expert-aligned tile raster, which lowers the fc1 / fc2 / dgrad reads)")
    else:
        line = {}

    # ---- outside the headline regions: the stage-3/5 kernels alone, the other named shapes, expert-parallel parity and
    # the C4 expert-parallel step.  The headline numbers above are complete at this point; a watchdog makes sure that a
    # hang in one of these sections (a peer that left the barrier sequence) costs that section, not the JSON line.
    done = threading.Event()

    def emit():
        if rank == 0:
            print(json.dumps(line), flush=True)

    def watchdog():
        if not done.wait(a.sections_timeout):
            line["sections_timeout"] = f"optional sections did not finish within {a.sections_timeout} s; the line carries what was complete"
            print(f"bench: rank {rank}: optional sections timed out", file=sys.stderr, flush=True)
            emit()
            os._exit(0)

    if a.sections:
        threading.Thread(target=watchdog, daemon=True).start()

        def guarded(name, fn):       # an exception in an optional section is recorded in the line, it does not cost the line
            try:
                line[name] = fn()
            except Exception as exc:
                line[name] = {"error": repr(exc)[:500]}
                print(f"bench: rank {rank}: section {name} failed: {exc!r}"[:600], file=sys.stderr, flush=True)

        guarded("hbm_stage", lambda: hbm_stage_gbs(device, peak_gbs))
        if ep_group is not None:
            line["ep_parity"] = ep_parity_section(device, world, rank, layer, x, dy, params)
        # a failed comparison is raised by every rank together (finish_case) and leaves the ranks aligned: the timing
        # section still runs; any other exception (a CUDA error after a trapped barrier) means the context is gone
        par = line.get("ep_parity") or {}
        if dist_on and 64 % world == 0 and (par.get("ok", True) or par.get("error", "").startswith("AssertionError")):
            line["c4_ep"] = {}
            c4_ep_section(device, world, rank, max(5, a.steps // 2), line["c4_ep"])
        if world == 1:
            guarded("configs", lambda: named_configs(device, max(6, a.steps // 2), peak_tf, peak_gbs))

            try:       # the CPU path timed beside the GPU numbers of the pretrain shapes (C1 is required)
                for key, entry in pretrain_port_entries(steps=3, warmup=1).items():
                    if isinstance(line.get("configs"), dict) and isinstance(line["configs"].get(key), dict):
                        line["configs"][key]["cpu_baseline"] = entry["cpu_baseline"]
            except Exception as exc:
                line["configs_cpu_error"] = repr(exc)[:300]
    done.set()
    emit()

    def teardown_watchdog():       # the line is out: a peer that died in a section must not keep this rank in a collective
        time.sleep(60.0)
        print(f"bench: rank {rank}: teardown did not return within 60 s, leaving", file=sys.stderr, flush=True)
        os._exit(0)

    if dist_on:
        threading.Thread(target=teardown_watchdog, daemon=True).start()
    try:
        if ep_group is not None:
            ep_group.close()
        if dist_on:
            dist.destroy_process_group()
    except Exception as exc:      # a section left the CUDA context in an error state: the line is out, leave quietly
        print(f"bench: rank {rank}: teardown failed: {exc!r}"[:400], file=sys.stderr, flush=True)
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--graphs", type=int, default=1, help="1: use the layer's CUDA-graph mode where available")
    ap.add_argument("--sections", type=int, default=1,
                    help="1: also report hbm_stage, the other named configs (N=1), EP parity and C4 expert-parallel (N>1)")
    ap.add_argument("--sections-timeout", type=float, default=240.0,
                    help="seconds the optional sections may take before the line is printed without them")
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
