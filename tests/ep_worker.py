"""Expert-parallel parity worker.  Run directly (world size 1) or under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/ep_worker.py

Every rank builds the same full layer (same seed) twice: one stays unsharded and processes the rank's own tokens
locally (the single-GPU path, itself checked against the oracle elsewhere), one is sharded over the group with
enable_expert_parallel().  Outputs / dx must match bit-for-bit (row results do not depend on where a row is computed);
expert-weight gradients are compared with the all-reduced single-GPU gradients (different summation order -> rtol).
Prints one line per case and "EP_WORKER_OK" at the end.
"""
import os
import sys
from pathlib import Path
from types import SimpleNamespace

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


# A failed check must not make ONE rank leave the call sequence: the other ranks would wait for it in the next
# device-side barrier.  Checks are recorded and the verdict is taken collectively at the end of a case (finish_case).
FAILS = []


def expect(cond, what):
    if not bool(cond):
        FAILS.append(str(what))


def close(a, b, rtol, what, summed_bf16=False):
    """|a - b| <= rtol * (|b| + rms(b)).  summed_bf16: a and b are sums of bf16-rounded rows computed along two different
    (equally valid) rounding sequences; a one-ulp difference in a summand as large as the largest element survives a
    cancelling sum unchanged, so one bf16 ulp of max|b| (2^-7 * max|b|) is added to the absolute term.  Seen at world 2:
    1 element of 153 600 off by exactly 2^-6 where the sum is ~0.3 and the summands are in [2, 4)."""
    a, b = a.float(), b.float()
    rms = b.pow(2).mean().sqrt()
    bad = (a - b).abs() > rtol * (b.abs() + rms) + 1e-6 + (2.0 ** -7 * b.abs().max() if summed_bf16 else 0.0)
    expect(not bool(bad.any()), f"{what}: {int(bad.sum())} / {bad.numel()} elements differ, max abs {float((a - b).abs().max()):.4g}")


def finish_case(dev, world, rank, name):
    """Collective: every rank learns whether any rank recorded a failure in this case; all raise together."""
    flag = torch.tensor([len(FAILS)], device=dev, dtype=torch.int32)
    if world > 1:
        dist.all_reduce(flag)
    mine = list(FAILS)
    FAILS.clear()
    if mine:
        print(f"[rank {rank}] {name}: " + " | ".join(mine), file=sys.stderr, flush=True)
    if int(flag) > 0:
        raise AssertionError(f"{name}: {int(flag)} failed check(s) over the ranks" + (": " + " | ".join(mine) if mine else ""))


class MLP(nn.Module):
    def __init__(self, d, f, dout):
        super().__init__()
        self.fc1, self.fc2, self.activation_fn = nn.Linear(d, f), nn.Linear(f, dout), nn.GELU(approximate="tanh")


class GLU(nn.Module):
    def __init__(self, d, f):
        super().__init__()
        self.gate_up_proj, self.down_proj = nn.Linear(d, 2 * f, bias=False), nn.Linear(f, d, bias=False)
        self.activation_fn = nn.SiLU()


def mm_args():
    return SimpleNamespace(rate_flip=0.05, warm_up=0.0, max_compete_in_iter=3, hybrid=False, router_theta=1.0,
                           router_loss_coef=0.01, diversity_loss_coef=0.01, bal_comp_loss_coef=0.01,
                           balance_loss_coef=0.01, router_z_loss_coef=0.001, norm_sigmoid=False, init_weight=True,
                           moe_name="competesmoe")


def pt_args():
    return SimpleNamespace(warm_up=0.0, rate_flip=0.07, stop_after=10, max_compete_in_iter=3, is_cosine=False,
                           is_norm_weight=False, norm_sigmoid=False, scale_weight=1.0, hybrid=False, tribrid=False,
                           in_topk=False, balance_affinity=True, balance_loss_coef=0.01, balance_loss_coef_comp=0.01,
                           router_loss_coef=0.01, router_theta=1.0, test_only=False)


def all_reduce_(t, world):
    if world > 1:
        dist.all_reduce(t)
    return t


def run_multimodal(group, dev, kind, E, K, D, Fh, B, N, competition, max_tokens=None):
    from competesmoe_b200.multimodal import CompeteSMoE
    rank, world = group.rank, group.world

    def build():
        torch.manual_seed(7)
        ex = nn.ModuleList([MLP(D, Fh, D) if kind == "mlp" else GLU(D, Fh) for _ in range(E)])
        layer = CompeteSMoE(D, D, E, K, ex, mm_args()).to(dev, torch.bfloat16)
        layer.total_steps, layer.step_warm = 2, 0
        layer.prob_flips = torch.full((2,), competition, device=dev)
        layer.set_current_steps(0)
        return layer.train()

    ref, epl = build(), build()
    epl.enable_expert_parallel(group, max_tokens=max_tokens or B * N)
    g = torch.Generator().manual_seed(100 + rank)
    x0 = torch.randn(B, N, D, generator=g).bfloat16().to(dev)
    dy = torch.randn(B, N, D, generator=g).bfloat16().to(dev)
    res = []
    for layer in (ref, epl):
        for rep in range(2):     # twice: the exchange buffers are reused across steps
            for p in layer.parameters():
                p.grad = None
            x = x0.clone().requires_grad_(True)
            out, aux, _, info = layer(x)
            torch.autograd.backward((out, aux), (dy, torch.ones_like(aux)))
        res.append((out.detach(), aux.detach(), x.grad.clone(), layer))
    (o_r, a_r, dx_r, lr), (o_e, a_e, dx_e, le) = res
    expect(torch.equal(lr.last_routing[0], le.last_routing[0]), "routing differs between the EP and local layers")
    if competition:
        close(o_e, o_r, 2e-2, "EP output"); close(dx_e, dx_r, 3e-2, "EP dx")
    else:
        expect(torch.equal(o_e, o_r), f"EP output differs bitwise: max {float((o_e.float() - o_r.float()).abs().max())}")
        expect(torch.equal(dx_e, dx_r), f"EP dx differs bitwise: max {float((dx_e.float() - dx_r.float()).abs().max())}")
    close(a_e, a_r, 1e-3, "EP aux")
    lo, El = le.ep_expert_offset, len(le.experts)
    for e in range(E):   # every rank joins every all-reduce; only the owner compares
        for n, q in lr.experts[e].named_parameters():
            want = all_reduce_(q.grad.float().clone(), world)
            if lo <= e < lo + El:
                close(dict(le.experts[e - lo].named_parameters())[n].grad, want, 3e-2, f"EP d experts.{e}.{n}")
    close(le.gate.weight.grad, lr.gate.weight.grad, 3e-2, "EP d gate")
    finish_case(dev, world, rank, f"ep multimodal {kind} E={E} K={K} T={B * N} {'competition' if competition else 'router'}")
    if rank == 0:
        print(f"ep multimodal {kind} E={E} K={K} D={D} F={Fh} T={B * N} world={world} "
              f"{'competition' if competition else 'router'}: ok", flush=True)


def run_pretrain(group, dev, E, K, D, H, B, N, competition, exchange="tokens", bias=False, check_graphs=True):
    from competesmoe_b200.pretrain import CompeteSMoE
    rank, world = group.rank, group.world

    def build():
        torch.manual_seed(11)
        layer = CompeteSMoE(D, E, H, n_heads=K, args=pt_args(), activation=F.relu, selection_mode="gate",
                            log_interval=None, bias=bias).to(dev)
        if bias:
            with torch.no_grad():
                layer.bias.normal_(0, 0.1)
                layer.o_bias.normal_(0, 0.1)
        layer.train()
        layer.regularization_present = True
        layer.set_total_steps(id_layer=0)
        layer.prob_flips_final[0] = torch.full_like(layer.prob_flips_final[0], competition)
        layer.set_current_steps(1)
        return layer

    ref, epl = build(), build()
    epl.enable_expert_parallel(group, max_tokens=B * N, exchange=exchange)
    expect((epl._wx is not None) == (exchange == "weights"), "exchange mode not honoured")
    g = torch.Generator().manual_seed(200 + rank)
    x0 = torch.randn(B, N, D, generator=g).to(dev)
    dy = torch.randn(B, N, D, generator=g).to(dev)
    res = []
    for layer in (ref, epl):
        for rep in range(2):
            for p in layer.parameters():
                p.grad = None
            x = x0.clone().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = layer(x, id_layer=0)
            regs = layer.get_reg_loss()
            loss = (out.float() * dy).sum() + sum(regs.values())
            loss.backward()
        res.append((out.detach(), x.grad.clone(), layer))
    (o_r, dx_r, lr), (o_e, dx_e, le) = res
    expect(torch.equal(lr.last_routing[0], le.last_routing[0]), "routing differs between the EP and local layers")
    if exchange == "weights":
        # the same kernels on the same operands (the gathered bf16 copies are the casts of the same fp32 parameters)
        expect(torch.equal(o_e, o_r), f"EP(weights) output differs bitwise: max {float((o_e.float() - o_r.float()).abs().max())}")
        expect(torch.equal(dx_e, dx_r), f"EP(weights) dx differs bitwise: max {float((dx_e.float() - dx_r.float()).abs().max())}")
    elif competition:
        close(o_e, o_r, 2e-2, "EP pretrain output"); close(dx_e, dx_r, 3e-2, "EP pretrain dx")
    else:
        # the local layer runs the fused sigma-MoE kernels (expert size 128), the expert-parallel one the grouped-GEMM
        # path on the received rows: same rounding points forward (bit-equal outputs were observed), but the fused
        # backward rounds dh once instead of twice -> bf16 tolerance on dx
        close(o_e, o_r, 2e-2, "EP pretrain output", summed_bf16=True)
        close(dx_e, dx_r, 2e-2, "EP pretrain dx", summed_bf16=True)
    lo, El = le.ep_expert_offset, E // world
    for n in ("keys", "values") + (("bias",) if bias else ()):
        want = all_reduce_(getattr(lr, n).grad.float().clone(), world)[lo:lo + El]
        # weights exchanged: the owner adds the ranks' fp32 gradients in rank order, NCCL in its own order -> 1e-5
        close(getattr(le, n).grad, want, 1e-5 if exchange == "weights" else 3e-2, f"EP d {n}")
    close(le.w_gate.grad, lr.w_gate.grad, 3e-2, "EP d w_gate")
    if check_graphs and (not competition or exchange == "weights"):
        # the same expert-parallel call replayed from CUDA graphs (device-side barriers are captured like any launch):
        # bit-identical to the eager expert-parallel step, on fresh inputs too.  A fresh layer: gradient accumulators
        # created by earlier eager backward passes live on the default stream and would be waited on during capture.
        epg = build()
        epg.enable_expert_parallel(group, max_tokens=B * N, exchange=exchange)
        epg.enable_cuda_graphs()
        for rep in range(3):
            gx = torch.Generator().manual_seed(300 + rank + rep)
            xn = torch.randn(B, N, D, generator=gx).to(dev)
            outs = []
            for layer in (le, epg):
                for p in layer.parameters():
                    p.grad = None
                x = xn.clone().requires_grad_(True)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    out = layer(x, id_layer=0)
                regs = layer.get_reg_loss()
                ((out.float() * dy).sum() + sum(regs.values())).backward()
                outs.append((out.detach().clone(), x.grad.clone(), layer.keys.grad.clone(), layer.w_gate.grad.clone()))
            expect(all(torch.equal(a, b) for a, b in zip(*outs)), "EP graph replay differs from the eager EP step")
        expect(len(epg._graphs) == 1, f"{len(epg._graphs)} captured graphs instead of 1")
    finish_case(dev, world, rank, f"ep pretrain E={E} K={K} H={H} T={B * N} exchange={exchange} {'competition' if competition else 'router'}")
    if rank == 0:
        print(f"ep pretrain E={E} K={K} D={D} H={H} T={B * N} world={world} exchange={exchange}{' bias' if bias else ''} "
              f"{'competition' if competition else 'router'}: ok", flush=True)


def main():
    import faulthandler
    faulthandler.enable()
    if os.environ.get("EP_WORKER_DUMP_AFTER"):      # debugging aid: every rank prints its Python stack if still running
        faulthandler.dump_traceback_later(float(os.environ["EP_WORKER_DUMP_AFTER"]), exit=False)
    only = os.environ.get("EP_WORKER_ONLY", "")     # "pretrain": the pretrain-plugin cases only
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from competesmoe_b200.ep import EPGroup
    group = EPGroup(None, dev)
    try:
        for comp in (False, True):
            # the expert count must be a multiple of the group size: 4 experts up to 4 ranks, 8 on an 8-GPU box
            if only != "pretrain":
                run_multimodal(group, dev, kind="mlp", E=4 if world <= 4 else 8, K=2, D=256, Fh=520, B=2, N=200, competition=comp)
                run_multimodal(group, dev, kind="glu", E=8, K=2, D=512, Fh=1024, B=1, N=1000, competition=comp)
            run_pretrain(group, dev, E=16, K=4, D=256, H=128, B=2, N=300, competition=comp, exchange="weights")
            run_pretrain(group, dev, E=16, K=4, D=256, H=128, B=2, N=300, competition=comp)
            run_pretrain(group, dev, E=2 * world, K=2, D=256, H=128, B=2, N=300, competition=comp)   # two experts per rank
            run_pretrain(group, dev, E=16, K=2, D=256, H=128, B=1, N=500, competition=comp, exchange="weights", bias=True)
            run_pretrain(group, dev, E=8, K=2, D=128, H=64, B=1, N=200, competition=comp, exchange="weights")   # grouped-GEMM path
        # ragged: a rank with very few tokens, top-1, more experts than tokens
        if only != "pretrain":
            run_multimodal(group, dev, kind="mlp", E=8, K=1, D=128, Fh=256, B=1, N=3 + 5 * group.rank, competition=False,
                           max_tokens=3 + 5 * (group.world - 1))
    finally:
        group.close()
        if world > 1:
            dist.destroy_process_group()
    if int(os.environ.get("RANK", "0")) == 0:
        print("EP_WORKER_OK", flush=True)


if __name__ == "__main__":
    main()
