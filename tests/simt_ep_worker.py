"""Test infrastructure: one emulated expert-parallel rank (a spawned process) running the sharded-vs-unsharded parity
cases of tests/ep_worker.py -- the same functions `bench.py --gpus N` and tests/test_gpu_ep.py run on GPUs -- on the CPU:
kernels from their own source on the SIMT emulator (tests/simt/), csmoe_grouped_gemm stood in by tests/simt/ref_gemm.cpp,
`torch.distributed` over gloo, "CUDA IPC" over POSIX shared memory (tests/simt/simt_rt.cpp).  See tests/test_simt_ep.py."""
import contextlib
import ctypes as C
import os
import sys
import traceback
from pathlib import Path

HERE = Path(__file__).resolve().parent
for p in (str(HERE), str(HERE.parent)):
    if p not in sys.path:
        sys.path.insert(0, p)


def _patch(lib):
    import torch
    from competesmoe_b200 import _lib, ep, functional, ops

    _lib.load = lambda: lib
    ops._cuda = lambda *ts: None
    ops._stream = lambda: None
    ops._ROUTER_GEMM = False
    functional._SIGMA_FUSED = False
    ep._WX_OVERLAP = False                                   # no side streams on the CPU
    state = {"on": False, "dtype": torch.bfloat16}

    @contextlib.contextmanager
    def autocast(device_type, dtype=torch.bfloat16, enabled=True, cache_enabled=None):
        old = dict(state)
        state.update(on=bool(enabled), dtype=dtype)
        try:
            yield
        finally:
            state.update(old)

    torch.autocast = autocast
    torch.is_autocast_enabled = lambda *a: state["on"]
    torch.get_autocast_dtype = lambda dev: state["dtype"]
    torch.cuda.is_current_stream_capturing = lambda: False
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.device = lambda d: contextlib.nullcontext()

    class _Stream:                                            # ep.py asks for the current stream even without overlap
        def wait_stream(self, other):
            pass

    torch.cuda.current_stream = lambda *a, **k: _Stream()

    def sb_init(self, local_ptr, peer_ptrs, nbytes, device):     # alias host memory instead of __cuda_array_interface__
        self.local_ptr, self.peer_ptrs, self.nbytes, self.device = local_ptr, peer_ptrs, nbytes, device
        self._keep = (C.c_uint8 * nbytes).from_address(local_ptr)
        self._bytes = torch.frombuffer(self._keep, dtype=torch.uint8)

    ep.SymmetricBuffer.__init__ = sb_init


def rank_main(rank: int, world: int, port: int, so_path: str, queue):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          OMP_NUM_THREADS="1")
        os.environ.setdefault("SIMT_SMS", "1")               # fewer blocks per launch: grid-stride loops, same results
        import torch
        import torch.distributed as dist
        torch.set_num_threads(1)
        from competesmoe_b200 import _lib
        lib = C.CDLL(so_path)
        for name, (res, args) in _lib._SIGNATURES.items():
            fn = getattr(lib, name, None)
            if fn is not None:
                fn.restype, fn.argtypes = res, args
        _patch(lib)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import ep_worker as ew
        from competesmoe_b200.ep import EPGroup
        dev = torch.device("cpu")
        group = EPGroup(None, dev)
        done = []
        try:
            E = max(8, world)
            full = os.environ.get("CSMOE_SIMT_FULL", "0") == "1"
            for comp in (False, True):
                ew.run_multimodal(group, dev, kind="mlp", E=E, K=2, D=64, Fh=72, B=2, N=12, competition=comp)
                done.append(f"multimodal mlp {'competition' if comp else 'router'}")
                if comp and not full:
                    continue         # the pretrain competition step under EP moves weights like the router step does
                ew.run_pretrain(group, dev, E=2 * world, K=2, D=128, H=64, B=1, N=24, competition=comp, exchange="weights",
                                check_graphs=False)
                done.append(f"pretrain weights {'competition' if comp else 'router'}")
                ew.run_pretrain(group, dev, E=2 * world, K=2, D=128, H=64, B=1, N=24, competition=comp, exchange="tokens",
                                bias=True, check_graphs=False)
                done.append(f"pretrain tokens bias {'competition' if comp else 'router'}")
            ew.run_multimodal(group, dev, kind="glu", E=E, K=1, D=64, Fh=64, B=1, N=3 + 2 * rank, competition=False,
                              max_tokens=3 + 2 * (world - 1))
            done.append("ragged glu top-1")
        finally:
            group.close()
            dist.destroy_process_group()
        queue.put((rank, True, done))
    except BaseException:
        queue.put((rank, False, traceback.format_exc()))
