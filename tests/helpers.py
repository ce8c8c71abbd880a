"""Shared helpers for the GPU parity tests: build the B200 modules from golden fixtures / oracle weight dicts."""
from types import SimpleNamespace

import torch
import torch.nn as nn


class MLPExpert(nn.Module):
    """Same attribute names as the reference's SiglipMLP (siglip_smoe.py:85-97)."""

    def __init__(self, d_in, f, d_out, act):
        super().__init__()
        self.activation_fn = {"gelu_tanh": nn.GELU(approximate="tanh"), "gelu": nn.GELU(), "relu": nn.ReLU(),
                              "silu": nn.SiLU()}[act]
        self.fc1 = nn.Linear(d_in, f)
        self.fc2 = nn.Linear(f, d_out)

    def forward(self, x):
        return self.fc2(self.activation_fn(self.fc1(x)))


class GLUExpert(nn.Module):
    """Same attribute names as transformers' Phi3MLP."""

    def __init__(self, d, f):
        super().__init__()
        self.gate_up_proj = nn.Linear(d, 2 * f, bias=False)
        self.down_proj = nn.Linear(f, d, bias=False)
        self.activation_fn = nn.SiLU()

    def forward(self, x):
        g, u = self.gate_up_proj(x).chunk(2, dim=-1)
        return self.down_proj(u * self.activation_fn(g))


def expert_from_weights(ew):
    if ew["kind"] == "mlp":
        f, d_in = ew["w1"].shape
        d_out = ew["w2"].shape[0]
        if ew["act"] == "gelu":   # projector-style Sequential, exercises the second recognised layout
            m = nn.Sequential(nn.Linear(d_in, f), nn.GELU(), nn.Linear(f, d_out))
            lin1, lin2 = m[0], m[2]
        else:
            m = MLPExpert(d_in, f, d_out, ew["act"])
            lin1, lin2 = m.fc1, m.fc2
        with torch.no_grad():
            lin1.weight.copy_(ew["w1"]); lin1.bias.copy_(ew["b1"])
            lin2.weight.copy_(ew["w2"]); lin2.bias.copy_(ew["b2"])
        return m
    f2, d = ew["w1"].shape
    m = GLUExpert(d, f2 // 2)
    with torch.no_grad():
        m.gate_up_proj.weight.copy_(ew["w1"]); m.down_proj.weight.copy_(ew["w2"])
    return m


def expert_linears(m):
    if isinstance(m, nn.Sequential):
        return m[0], m[2]
    if hasattr(m, "fc1"):
        return m.fc1, m.fc2
    return m.gate_up_proj, m.down_proj


def build_multimodal_layer(fx, device, dtype=None):
    from competesmoe_b200.multimodal import CompeteSMoE
    m = fx["meta"]
    args = SimpleNamespace(**m["args"])
    experts = nn.ModuleList([expert_from_weights(ew) for ew in fx["experts"]])
    layer = CompeteSMoE(m["d_in"], m["d_out"], m["E"], m["K"], experts, args)
    with torch.no_grad():
        layer.gate.weight.copy_(fx["gate_w"])
    dtype = dtype or fx["x"].dtype
    layer = layer.to(device=device, dtype=dtype)
    layer.total_steps, layer.step_warm = 4, 0
    layer.prob_flips = torch.full((4,), bool(m["competition"]), device=device)
    layer.set_current_steps(1)
    return layer


def rel_err(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return float((got - ref).abs().max() / (ref.abs().max() + 1e-12))


def assert_close_rms(got, ref, rtol, what="", outliers=0.0):
    """|got - ref| <= rtol * (|ref| + rms(ref)): the north-star's rtol with an atol scaled by the tensor's RMS.
    `outliers`: fraction of elements allowed outside that band (ill-conditioned gate gradients in bf16), each still
    within 5x the band, and the whole tensor within rtol in the Frobenius norm."""
    got, ref = got.float().cpu(), ref.float().cpu()
    rms = ref.pow(2).mean().sqrt()
    bad = (got - ref).abs() > rtol * (ref.abs() + rms) + 1e-12
    if outliers > 0.0 and 0 < int(bad.sum()) <= outliers * bad.numel():
        far = (got - ref).abs() > 5 * rtol * (ref.abs() + rms) + 1e-12
        fro = float((got - ref).norm() / (ref.norm() + 1e-30))
        assert not bool(far.any()) and fro <= rtol, f"{what}: outliers too far ({int(far.sum())}) or Frobenius error {fro:.3e} > {rtol}"
        return
    assert not bool(bad.any()), (f"{what}: {int(bad.sum())}/{bad.numel()} elements off; max abs err "
                                 f"{float((got - ref).abs().max()):.4e}, ref rms {float(rms):.4e}")
