"""Debug aid: which expert-parallel call breaks CUDA-graph capture (world size 1)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from competesmoe_b200 import ops
from competesmoe_b200.ep import EPGroup, EPLayerState

dev = torch.device("cuda")
torch.cuda.set_device(0)
g = EPGroup(None, dev)
T, K, E, D, H = 600, 4, 16, 256, 128
st = EPLayerState(g, E, K, D, D, T, 128)
gen = torch.Generator().manual_seed(0)
sel = torch.stack([torch.randperm(E, generator=gen)[:K] for _ in range(T)]).int().to(dev)
x = torch.randn(T, D, generator=gen).bfloat16().to(dev)
w = torch.rand(T, K, generator=gen).to(dev)
keys = torch.randn(E, D, H, generator=gen).bfloat16().to(dev)
values = torch.randn(E, H, D, generator=gen).bfloat16().to(dev)
state = {}


def s_route(): state["route"] = ops.route_build(sel, E, row_tile=128)
def s_plan(): state["plan"] = g.exchange_plan(state["route"].counts, E, st.row_tile, st.row_cap)
def s_dispatch(): g.dispatch(x, K, state["route"], state["plan"], st.recv_x.peers(), st.tags.peers())
def s_barrier(): g.barrier()
def s_rowptrs():
    xp = st.recv_x.tensor(0, (st.row_cap, D), torch.bfloat16)
    tags = st.tags.tensor(0, (st.row_cap,), torch.int64)
    state["xp"] = xp
    state["c_rows"] = g.row_ptrs(tags, state["plan"], st.ret_y.peers(), D, torch.bfloat16, xp)
def s_froute(): state["fr"] = state["plan"].fused_route(state["c_rows"])
def s_fwd(): state["y"], state["h"] = ops.sigma_ffn_fwd(state["xp"], keys, values, None, state["fr"], slots_per_row=1, xp=state["xp"])
def s_push():
    y = state["y"]
    ops._call("csmoe_ep_push_rows", y.data_ptr(), ops.BF16, y.shape[1], y.shape[1], y.shape[0], state["c_rows"].data_ptr(), ops._stream())
def s_combine():
    y = st.ret_y.tensor(0, (T * K, D), torch.bfloat16)
    state["out"] = ops.combine_fwd(y, g.identity(T * K), state["route"].sel, w, T, K, round_w=True)
def s_scratch():
    state["recv_dy"] = g.scratch("recv_dy", st.row_cap * D * 2)
    state["ret_dx"] = g.scratch("ret_dx", st.max_slots * D * 2)
def s_dispatch_bwd(): g.dispatch(x, K, state["route"], state["plan"], state["recv_dy"].peers(), None, slot_w=w)
def s_bwd():
    dyp = state["recv_dy"].tensor(0, (st.row_cap, D), torch.bfloat16)
    ones = torch.ones(st.row_cap, dtype=torch.float32, device=dev)
    state["bw"] = ops.sigma_ffn_bwd(dyp, keys, values, state["fr"], ones, state["h"], slots_per_row=1, dyp=dyp)
def s_wgrad():
    dyp = state["recv_dy"].tensor(0, (st.row_cap, D), torch.bfloat16)
    state["dv"] = ops.sigma_wgrad(state["h"], dyp, E, state["fr"], transpose=False, slots_per_row=1)


steps = [s_route, s_plan, s_dispatch, s_barrier, s_rowptrs, s_froute, s_fwd, s_push, s_barrier, s_combine, s_scratch, s_dispatch_bwd,
         s_barrier, s_bwd, s_wgrad]
for f in steps:       # eager pass first
    f()
torch.cuda.synchronize()
side = torch.cuda.Stream()
for f in steps:
    try:
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=side):
            f()
        print(f"{f.__name__:16s} capture ok", flush=True)
    except Exception as exc:
        print(f"{f.__name__:16s} CAPTURE FAILED: {str(exc).splitlines()[0]}", flush=True)
        torch.cuda.synchronize()
g.close()
