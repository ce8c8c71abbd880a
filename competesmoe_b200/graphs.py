"""Whole-step CUDA-graph capture of a MoE layer's forward + backward.

Every libcsmoe entry point is capturable (no host sync, no allocation, caller-owned buffers), the layer's schedule test
runs on the host, and all shapes are static, so one training step of the layer can be replayed as a single graph launch.
That removes the ~40 per-step launch gaps (torch glue + ctypes calls), which is what "capture launch-bound inner loops in
CUDA graphs" buys on a 3.5 ms step.  Two graphs are kept per layer -- router step and competition step -- and the host
picks one per call exactly like the reference's `prob_flips[...] == 1` test.
"""
from __future__ import annotations

import contextlib
import gc
from typing import Dict, Optional, Tuple

import torch


@contextlib.contextmanager
def capture_guard():
    """Around every graph capture: collect garbage first and keep the cyclic collector off until the capture has ended.
    A layer in graph mode is part of a reference cycle (layer -> captured callable -> closure -> layer), so an *old*
    layer's CUDA graphs and their private memory pool are destroyed whenever the collector happens to run -- and a
    cudaGraphExecDestroy / cudaFree in the middle of another capture invalidates that capture
    (cudaErrorStreamCaptureInvalidated; seen with several graph-mode layers created one after the other)."""
    gc.collect()
    torch.cuda.synchronize()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was_enabled:
            gc.enable()


class GraphedStep:
    """Replays `out, aux = layer(x); backward((out, aux), (dy, 1))` from a captured graph.

    After `run(x, dy)`: `self.out`, `self.aux` hold the forward results, `self.dx` the input gradient and every
    parameter's `.grad` the parameter gradient (static buffers owned by the graph: consume them before the next run).
    """

    def __init__(self, layer: torch.nn.Module, x_example: torch.Tensor, warmup: int = 3, **fwd_kwargs):
        assert x_example.is_cuda
        self.layer = layer
        self.kwargs = fwd_kwargs
        self.x = x_example.detach().clone().requires_grad_(True)
        self.dy: Optional[torch.Tensor] = None
        self.graphs: Dict[bool, Tuple[torch.cuda.CUDAGraph, torch.Tensor, torch.Tensor]] = {}
        self.warmup = warmup
        self.pool = None
        self.out = self.aux = self.dx = None

    def _step(self):
        res = self.layer(self.x, **self.kwargs)
        out, aux = (res[0], res[1]) if isinstance(res, tuple) else (res, None)
        if self.dy is None:
            self.dy = torch.zeros_like(out)
        if aux is not None and aux.requires_grad:
            torch.autograd.backward((out, aux), (self.dy, torch.ones_like(aux)))
        else:
            out.backward(self.dy)
        return out, aux

    def _capture(self, branch: bool):
        params = [p for p in self.layer.parameters() if p.requires_grad]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                for p in params:
                    p.grad = None
                self.x.grad = None
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for p in params:
            p.grad = None
        self.x.grad = None
        g = torch.cuda.CUDAGraph()
        with capture_guard(), torch.cuda.graph(g, pool=self.pool):
            out, aux = self._step()
        if self.pool is None:
            self.pool = g.pool()
        self.graphs[branch] = (g, out, aux, {id(p): p.grad for p in params}, self.x.grad)

    def run(self, x: torch.Tensor, dy: torch.Tensor, branch: Optional[bool] = None):
        """branch: True = competition step, False = router step, None = ask the layer's schedule."""
        if branch is None:
            probe = getattr(self.layer, "_is_competition_step", None)
            branch = bool(probe(self.x)) if probe is not None else False
        if self.dy is None:
            self.dy = torch.zeros_like(dy)
        if branch not in self.graphs:
            self._capture(branch)
        g, out, aux, grads, dx = self.graphs[branch]
        self.x.detach().copy_(x, non_blocking=True)
        self.dy.copy_(dy, non_blocking=True)
        g.replay()
        for p in self.layer.parameters():
            if id(p) in grads:
                p.grad = grads[id(p)]
        self.out, self.aux, self.dx = out, aux, dx
        return out, aux
