"""CPU checks of bench.py's host logic: the reference arm's JSON contract (the CPU path of the workload on a small token
sample), the silent non-zero ranks of that arm, the FLOP model of SURVEY.md 8(d), and the e2e loop's buffer / event
sequence run on stand-ins for the CUDA stream objects (the loop itself needs a GPU; a Python-level slip in it would cost
the whole bench line, so its control flow is exercised here)."""
import contextlib
import json
import types

import pytest
import torch
import torch.nn as nn

import bench


def test_flop_model_follows_survey_8d():
    # C2: F_e = 6 D F = 150 994 944; router step 3 (K F_e + R) = 906 MFLOP / token, competition 1 812 MFLOP / token
    assert bench.flops_per_token(False) == 3 * (2 * 150_994_944 + 2 * 3072 * 4)
    assert bench.flops_per_token(True) == 3 * (4 * 150_994_944 + 2 * 3072 * 4)
    assert abs(bench.flops_per_token(False) * 4096 / 1e12 - 3.711) < 2e-3


@pytest.mark.parametrize("force_port", [False, True])
def test_cpu_step_time_runs_reference_or_port(monkeypatch, force_port):
    if force_port:
        monkeypatch.setattr(bench, "REFERENCE_ROOTS", [""])
    dt, cores, tokens, kind, source = bench.cpu_reference_step_time(1, 1, tokens=32)
    assert dt > 0 and cores >= 1 and tokens == 32
    if force_port:
        assert kind == "port" and source == "oracle/multimodal.py"
    else:
        assert kind in ("reference", "port")
        assert (kind == "reference") == source.endswith("competesmoe.py (unmodified)")


def test_reference_arm_line_contract(monkeypatch, capsys):
    monkeypatch.setattr(bench, "TOKENS", 32)
    monkeypatch.setattr(bench, "REFERENCE_ROOTS", [""])
    monkeypatch.setenv("RANK", "0")
    real = bench.cpu_reference_step_time
    monkeypatch.setattr(bench, "cpu_reference_step_time", lambda s, w, **kw: real(s, w, tokens=32, **kw))
    bench.run_reference_arm(types.SimpleNamespace(steps=2, warmup=1, gpus=4))
    lines = [l for l in capsys.readouterr().out.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "moe_layer_fwd_bwd_tokens_per_s"
    assert line["unit"] == "tokens/s" and line["higher_is_better"] is True
    assert line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 4
    assert line["config"]["workload"] == bench.WORKLOAD
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_reference_arm_is_silent_off_rank_zero(monkeypatch, capsys):
    monkeypatch.setenv("RANK", "3")
    monkeypatch.setattr(bench, "cpu_reference_step_time", lambda *a, **k: pytest.fail("rank 3 must not work"))
    bench.run_reference_arm(types.SimpleNamespace(steps=2, warmup=1, gpus=4))
    assert capsys.readouterr().out == ""


class _Event:
    log = []

    def __init__(self, enable_timing=False):
        self.recorded = 0

    def record(self, stream=None):
        self.recorded += 1

    def synchronize(self):
        assert self.recorded > 0, "the host waited on an event nobody recorded"

    def elapsed_time(self, other):
        return 7.0


class _Stream:
    def __init__(self, device=None):
        pass

    def wait_event(self, e):
        pass


class _Layer(nn.Module):
    """Stands in for the plugin: (out, aux, None, info) with aux depending on the batch."""

    def __init__(self):
        super().__init__()
        self.w = nn.Parameter(torch.ones(8))
        self.calls = 0

    def forward(self, x):
        self.calls += 1
        return x * self.w, (x * self.w).mean() + self.calls, None, None


def test_e2e_loop_reads_every_loss_on_the_host(monkeypatch):
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "Stream", _Stream)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: _Stream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    layer = _Layer()
    x = torch.randn(1, 4, 8)
    dy = torch.randn(1, 4, 8)
    steps, warmup = 5, 3
    ms, h2d, d2h, reads = bench.e2e_region(layer, x, dy, list(layer.parameters()), steps, warmup, False, torch.device("cpu"))
    assert ms == 7.0 / steps
    assert h2d == 2 * x.numel() * 4 and d2h == 4
    assert layer.calls == steps + warmup
    assert reads["losses_read_on_host"] == steps and reads["all_finite"]
    want = float((x * layer.w).detach().mean()) + steps + warmup      # the LAST step's loss was the last one read
    assert abs(reads["last_loss"] - want) < 1e-5
    json.dumps(reads)


def test_pretrain_shapes_have_a_cpu_baseline(monkeypatch):
    # the shapes are BASELINE.json configs[0] / configs[3]; shrink the token count so that the check stays cheap
    assert bench.PRETRAIN_SHAPES["C1"] | {"what": ""} == dict(what="", B=8, N=512, D=512, E=8, K=2, H=128)
    assert {k: bench.PRETRAIN_SHAPES["C4"][k] for k in ("D", "E", "K", "H")} == dict(D=1024, E=64, K=8, H=128)
    small = {k: dict(v, B=2, N=64) for k, v in bench.PRETRAIN_SHAPES.items()}
    monkeypatch.setattr(bench, "PRETRAIN_SHAPES", small)
    out = bench.pretrain_port_entries(steps=1, warmup=1)
    assert set(out) == {"C1", "C4"}
    for entry in out.values():
        cb = entry["cpu_baseline"]
        assert cb["kind"] == "port" and cb["unit"] == "tokens/s" and cb["value"] > 0 and cb["cores"] >= 1
        assert "gpu_eager_loop" not in entry
    json.dumps(out)
