"""Expert-parallel path (stage 6) on real GPUs: tests/ep_worker.py run as world size 1 in a fresh process, and under
torchrun with 2 ranks when the box has at least 2 GPUs."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _run(cmd, timeout=600):
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0 and "EP_WORKER_OK" in r.stdout, f"{' '.join(cmd)}\n--- stdout\n{r.stdout[-3000:]}\n--- stderr\n{r.stderr[-3000:]}"
    return r.stdout


def test_ep_world_size_1_matches_local_layer():
    out = _run([sys.executable, "tests/ep_worker.py"])
    assert out.count(": ok") >= 7


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_ep_world_size_2_matches_local_layer():
    out = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                "127.0.0.1", "--master-port", "29533", "tests/ep_worker.py"])
    assert out.count(": ok") >= 7


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_grad_sync_nccl_world_size_2():
    """EP-aware bucketed gradient reduction over NCCL against the unsharded layer run on all ranks' tokens
    (tests/grad_sync_worker.py; reference loop: framework/task/simple_task.py:403-413)."""
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29534", "tests/grad_sync_worker.py"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "GRAD_SYNC_OK" in r.stdout, f"--- stdout\n{r.stdout[-3000:]}\n--- stderr\n{r.stderr[-3000:]}"
