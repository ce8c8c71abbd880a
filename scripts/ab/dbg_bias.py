"""Debug aid: which gradient of the pretrain layer deviates from the oracle for (bias, H) combinations."""
import sys
from pathlib import Path
import torch
import torch.nn.functional as F
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from oracle import pretrain as op
import test_gpu_full_configs as tf

for bias in (False, True):
    for H in (64, 128):
        for comp in (False,):
            res = tf._pt_case(256, H, 8, 2, 2, 192, comp, seed=1240, bias=bias)
            layer, names, ref_p, xg, out, regs, sel, w, xr, o_out, o_regs, dbg = res
            def be(a, b):
                a, b = a.detach().float().cpu(), b.detach().float().cpu()
                rms = b.pow(2).mean().sqrt()
                return ((a - b).abs() / (b.abs() + rms)), rms
            e, _ = be(xg.grad, xr.grad)
            bad_tok = (e > 0.02).any(-1).nonzero()
            print(f"bias={bias} H={H}: out {float(be(out, o_out)[0].max()):.3e} dx max {float(e.max()):.3e} bad elems {int((e>0.02).sum())} bad tokens {bad_tok.shape[0]} e.g. {bad_tok[:6].tolist()}")
            for n in names:
                en, _ = be(getattr(layer, n).grad, ref_p[n].grad)
                print(f"     d{n}: {float(en.max()):.3e} bad {int((en>0.02).sum())}/{en.numel()}")
            if bad_tok.shape[0]:
                b, t = bad_tok[0].tolist()
                print("     token", b, t, "sel", sel[b, t].tolist(), "w", w[b, t].tolist(), "gpu", xg.grad[b, t, :6].tolist(), "ref", xr.grad[b, t, :6].tolist())
