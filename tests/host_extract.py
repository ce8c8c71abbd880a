"""Test infrastructure: cut pure integer / scalar device functions out of the shipped .cu / .h sources AS THEY ARE and
compile them for the host with g++, so that `-m "not gpu"` tests can check the kernels' own index algebra and scalar math
(tile rasters, the Philox stream of the dropout mask, activation derivatives) without a GPU.  Nothing here is used by the
product."""
import ctypes as C
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "competesmoe_b200" / "csrc"


def block(text: str, start_pat: str, what: str = "") -> str:
    """The source from the match of start_pat to the brace that closes the first '{' after it (plus a trailing ';')."""
    m = re.search(start_pat, text)
    assert m, f"{start_pat!r} not found{' in ' + what if what else ''}: the host-extraction test needs updating"
    i = text.index("{", m.start())
    depth = 0
    for j in range(i, len(text)):
        if text[j] == "{":
            depth += 1
        elif text[j] == "}":
            depth -= 1
            if depth == 0:
                end = j + 1
                if text[end:end + 1] == ";":
                    end += 1
                return text[m.start():end]
    raise AssertionError("unbalanced braces")


def compile_host(source: str, workdir: Path, name: str) -> C.CDLL:
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    src = workdir / f"{name}.cpp"
    src.write_text(source)
    so = workdir / f"{name}.so"
    r = subprocess.run([gxx, "-O1", "-std=c++17", "-shared", "-fPIC", "-I", str(ROOT / "include"), str(src), "-o", str(so)],
                       capture_output=True, text=True)
    assert r.returncode == 0, f"the extracted device source no longer compiles for the host:\n{r.stderr[-3000:]}"
    return C.CDLL(str(so))
