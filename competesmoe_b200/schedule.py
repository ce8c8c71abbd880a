"""Competition schedule shared by both plugins.

Host-side integer logic, kept in Python on purpose (init-time only): for a given RNG stream it must produce the same
flags as the reference's set_total_steps (moe_model/model/moe/competesmoe.py:35-179;
moe_pretrain_model/layers/moe/competesmoe.py:123-273), so it consumes the generator exactly the same way: one
`torch.rand(1, device=...)` per step, on the CUDA device when one is available.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import torch
import torch.distributed as dist


def _first_free(cur, freq, cap: int, positions: Iterable[int]) -> Optional[int]:
    for j in positions:
        if freq[j] < cap and not cur[j]:
            return j
    return None


def draw_flags(flip_steps: int, rate_flip: float, cap: int, prior: Optional[Dict[int, torch.Tensor]],
               device: torch.device) -> torch.Tensor:
    """Bernoulli(rate_flip) flag per step; a step that already has `cap` competing layers (summed over `prior`) hands
    its flag to the nearest earlier free step, else the nearest later one."""
    freq = [0] * flip_steps
    for v in (prior or {}).values():
        for i, f in enumerate(v.int().tolist()):
            freq[i] += f
    cur = [False] * flip_steps
    for i in range(flip_steps):
        if torch.rand(1, device=device).item() >= rate_flip:
            continue
        j: Optional[int] = i
        if freq[i] >= cap:
            j = _first_free(cur, freq, cap, range(i - 1, -1, -1))
            if j is None:
                j = _first_free(cur, freq, cap, range(i + 1, flip_steps))
        if j is not None:
            cur[j] = True
            freq[j] += 1
    return torch.tensor(cur, dtype=torch.bool, device=device)


def make_layer_schedule(total_steps: int, warm_up: float, rate_flip: float, cap: int,
                        prior: Optional[Dict[int, torch.Tensor]]):
    """Returns (step_warm, flags[flip_steps]) with rank 0 drawing and every other rank receiving a broadcast."""
    step_warm = int(warm_up * total_steps)
    flip_steps = total_steps - step_warm
    if flip_steps <= 0:
        raise ValueError("self.total_steps - self.step_warm must be greater than 0.")
    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if distributed else 0
    world = dist.get_world_size() if distributed else 1
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    if rank == 0:
        flags = draw_flags(flip_steps, rate_flip, cap, prior, device)
    else:
        flags = torch.empty(flip_steps, dtype=torch.bool, device=device)
    if world > 1:
        dist.broadcast(flags, src=0)
    return step_warm, flags
