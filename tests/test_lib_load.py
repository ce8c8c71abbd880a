"""CPU-side checks of the C-ABI boundary: the library builds, loads and exports every symbol include/csmoe.h declares
(no compute calls without a GPU), and argument errors are reported through status codes, not crashes."""
import ctypes as C
import re
from pathlib import Path

import pytest

from competesmoe_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "csmoe.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(csmoe_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == _lib.exported_symbols()


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.csmoe_abi_version() == _lib.ABI_VERSION == 2


def test_row_cap_and_workspace_queries():
    lib = _lib.load()
    assert lib.csmoe_route_row_cap(8192, 4, 128) == 8704  # 8192 + 4*127 rounded up to 128
    assert lib.csmoe_route_row_cap(8192, 4, 256) == 9216  # 8192 + 4*255 rounded up to 256
    assert lib.csmoe_route_row_cap(0, 1, 128) == 128
    assert lib.csmoe_route_row_cap(10, 0, 128) == -1
    assert lib.csmoe_route_row_cap(10, 4, 100) == -1
    assert lib.csmoe_route_workspace_bytes(4096, 8) == 2 * 8 * 4


def test_argument_errors_are_status_codes():
    lib = _lib.load()
    # NULL args struct -> CSMOE_ERR_ARG, with a message; must not crash even without a GPU
    assert lib.csmoe_grouped_gemm(None, None) == -1
    assert b"NULL" in lib.csmoe_last_error()
    assert lib.csmoe_router_fwd(None, None, 1, 4, 64, 4, 2, 1, None, None, None, None, None) == -1
    g = _lib.GemmArgs()
    g.a = g.b = g.c = 16
    g.mode, g.num_experts, g.m, g.n, g.k = 0, 1, 128, 12, 64   # n not a multiple of 8
    g.lda = g.ldb = g.ldc = 64
    assert lib.csmoe_grouped_gemm(C.byref(g), None) == -1
    assert b"multiple of 8" in lib.csmoe_last_error()


def test_ops_refuse_cpu_tensors():
    import torch
    from competesmoe_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.router_fwd(torch.zeros(4, 64), torch.zeros(4, 64), 2)


def test_header_is_plain_c():
    """include/csmoe.h is the drop-in boundary: it must compile as C (cgo / JNI / ctypes-style bindings include it as is)
    and as C++."""
    import shutil
    import subprocess
    from pathlib import Path
    hdr = Path(__file__).resolve().parent.parent / "include" / "csmoe.h"
    for cc, args in (("gcc", ["-x", "c", "-std=c11"]), ("g++", ["-x", "c++", "-std=c++17"])):
        if shutil.which(cc) is None:
            continue
        r = subprocess.run([cc, *args, "-fsyntax-only", "-Wall", "-Werror", str(hdr)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
