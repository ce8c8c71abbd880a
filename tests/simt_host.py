"""Test infrastructure: build the shipped .cu sources of the bandwidth-bound kernels for the HOST against the SIMT emulator
(tests/simt/simt.h) and bind the result with the product's own ctypes signature table.

The sources are used as they are, with three mechanical rewrites (no kernel code is duplicated here):
  * `kernel<<<grid, block, smem, stream>>>(args);`  ->  `simt::launch(grid, block, smem, [&]() { kernel(args); });`
  * `extern __shared__ T name[];`                   ->  `T* name = reinterpret_cast<T*>(simt::dyn_smem());`
  * the three inline-PTX one-liners these files contain (`%lanemask_lt`, `tanh.approx.f32`) -> their C++ meaning.
ep.cu (peer-memory kernels) is built on request: its release / acquire flag accesses become __atomic builtins, and the
"ranks" of a test are forked processes that share anonymous mappings.  The tensor-core / TMA kernels (gemm_tcgen05.cu,
sigma_ffn.cu) are not built here.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "competesmoe_b200" / "csrc"
SIMT = Path(__file__).resolve().parent / "simt"
SOURCES = ["common.cu", "routing.cu", "router.cu", "permute.cu", "losses.cu", "compete.cu", "act.cu", "block.cu"]


def _match_paren(text: str, i: int) -> int:
    """Index just past the parenthesis that closes text[i] == '('."""
    assert text[i] == "("
    depth = 0
    for j in range(i, len(text)):
        if text[j] == "(":
            depth += 1
        elif text[j] == ")":
            depth -= 1
            if depth == 0:
                return j + 1
    raise AssertionError("unbalanced parentheses after a kernel launch")


def _kernel_name_start(text: str, end: int) -> int:
    """Start of the kernel expression that ends at text[end] (exclusive): identifier, `ns::`, balanced template args."""
    i = end
    while i > 0 and text[i - 1].isspace():
        i -= 1
    while i > 0:
        ch = text[i - 1]
        if ch == ">":
            depth, i = 1, i - 1
            while i > 0 and depth:
                i -= 1
                depth += {">": 1, "<": -1}.get(text[i], 0)
        elif ch.isalnum() or ch in "_:":
            i -= 1
        else:
            break
    return i


def _split_top(s: str) -> list[str]:
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "(<[{":
            depth += 1
        elif ch in ")>]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def rewrite_launches(text: str) -> tuple[str, int]:
    out, pos, n = "", 0, 0
    while True:
        k = text.find("<<<", pos)
        if k < 0:
            return out + text[pos:], n
        start = _kernel_name_start(text, k)
        close = text.index(">>>", k)
        cfg = _split_top(text[k + 3:close])
        assert 2 <= len(cfg) <= 4, f"launch configuration not understood: {text[k:close + 3]!r}"
        a = close + 3
        while text[a].isspace():
            a += 1
        b = _match_paren(text, a)
        name, args = text[start:k].strip(), text[a + 1:b - 1]
        smem = cfg[2] if len(cfg) > 2 else "0"
        out += text[pos:start] + f"simt::launch({cfg[0]}, {cfg[1]}, {smem}, [&]() {{ {name}({args}); }})"
        pos, n = b, n + 1


def rewrite(text: str) -> tuple[str, dict]:
    text, n_launch = rewrite_launches(text)
    text, n_dyn = re.subn(r"extern\s+__shared__\s+([\w:]+)\s+(\w+)\s*\[\s*\]\s*;",
                          r"\1* \2 = reinterpret_cast<\1*>(simt::dyn_smem());", text)
    text, n_lm = re.subn(r'asm\("mov\.u32 %0, %%lanemask_lt;"\s*:\s*"=r"\((\w+)\)\);',
                         r"\1 = (1u << simt::ctx().lane) - 1u;", text)
    text, n_th = re.subn(r'asm\("tanh\.approx\.f32 %0, %1;"\s*:\s*"=f"\((\w+)\)\s*:\s*"f"\((.+?)\)\);', r"\1 = tanhf(\2);", text)
    # ep.cu: system-scope release / acquire on the barrier flags, the wall clock of the bounded spin
    text, n_st = re.subn(r'asm volatile\("st\.release\.sys\.global\.s32 \[%0\], %1;"\s*::\s*"l"\((\w+)\),\s*"r"\((\w+)\)\s*:\s*"memory"\);',
                         r"__atomic_store_n(\1, \2, __ATOMIC_RELEASE);", text)
    text, n_ld = re.subn(r'asm volatile\("ld\.acquire\.sys\.global\.s32 %0, \[%1\];"\s*:\s*"=r"\((\w+)\)\s*:\s*"l"\((\w+)\)\s*:\s*"memory"\);',
                         r"\1 = __atomic_load_n(\2, __ATOMIC_ACQUIRE);", text)
    text, n_gt = re.subn(r'asm volatile\("mov\.u64 %0, %globaltimer;"\s*:\s*"=l"\((\w+)\)\);', r"\1 = simt::wall_ns();", text)
    assert "asm(" not in text and "asm volatile" not in text, "inline PTX the emulator has no rewrite for"
    return text, dict(launches=n_launch, dynamic_smem=n_dyn, lanemask=n_lm, tanh_approx=n_th, sys_flags=n_st + n_ld + n_gt)


def build(workdir: Path, ref_gemm: bool = False, ep: bool = False, tsan: bool = False, load: bool = True):
    """ref_gemm: also link tests/simt/ref_gemm.cpp, a plain-loop statement of csmoe_grouped_gemm's contract (NOT the
    tensor-core kernel), so that whole layers can run on the emulator."""
    import os
    prebuilt = os.environ.get("CSMOE_SIMT_PREBUILT")       # scripts/simt_racecheck.py: a ThreadSanitizer build of everything
    if prebuilt and load:
        return _bind(C.CDLL(prebuilt), {"prebuilt": prebuilt, "ep.cu": dict(sys_flags=3, launches=12)})
    gxx = shutil.which("g++")
    if gxx is None or not Path("/usr/local/cuda/include/cuda_runtime.h").exists():
        pytest.skip("g++ or the CUDA headers are not available")
    workdir.mkdir(parents=True, exist_ok=True)
    stats, objs = {}, []
    # common.h is a header of the emulated sources as well (activations, vector loads): it goes through the same rewrite,
    # and the rewritten copy is the only one on the include path
    inc = workdir / "inc"
    inc.mkdir(exist_ok=True)
    for h in sorted(CSRC.glob("*.h")):
        t, stats[h.name] = rewrite(h.read_text())
        (inc / h.name).write_text(t)
    # tsan: ThreadSanitizer build.  The barriers are std::mutex / std::atomic / sem_t, which TSAN models, so two emulated
    # CUDA threads touching the same shared or global word without a __syncthreads / __syncwarp / collective in between
    # are reported as a data race (scripts/simt_racecheck.py; needs LD_PRELOAD=libtsan.so, hence load=False)
    flags = (["-fsanitize=thread", "-g"] if tsan else []) + ["-O1", "-std=c++17", "-fPIC", "-pthread", "-w", "-I", str(inc), "-I", str(SIMT), "-I", str(ROOT / "include"),
             "-I", "/usr/local/cuda/include", "-include", "simt.h"]
    jobs = []
    for name in SOURCES + (["ep.cu"] if ep else []):     # ep: the peer-memory kernels (ranks = forked processes, shared mappings)
        t, stats[name] = rewrite((CSRC / name).read_text())
        src = workdir / (Path(name).stem + "_simt.cpp")
        src.write_text(t)
        obj = src.with_suffix(".o")
        jobs.append((obj, subprocess.Popen([gxx, *flags, "-c", str(src), "-o", str(obj)], stdout=subprocess.PIPE,
                                           stderr=subprocess.STDOUT, text=True)))
    for extra in ["simt_rt.cpp"] + (["ref_gemm.cpp"] if ref_gemm else []):
        obj = workdir / (Path(extra).stem + ".o")
        std = ["-std=c++20"] if extra == "simt_rt.cpp" else []      # C++20 atomic wait / notify in the barrier
        jobs.append((obj, subprocess.Popen([gxx, *flags, *std, "-c", str(SIMT / extra), "-o", str(obj)], stdout=subprocess.PIPE,
                                           stderr=subprocess.STDOUT, text=True)))
    for obj, p in jobs:
        out, _ = p.communicate()
        assert p.returncode == 0, f"{obj.name}: the source no longer builds against the SIMT emulator:\n{out[-4000:]}"
        objs.append(str(obj))
    so = workdir / "libcsmoe_simt.so"
    r = subprocess.run([gxx, "-shared", "-pthread", *(["-fsanitize=thread"] if tsan else []), "-Wl,-Bsymbolic", "-o", str(so), *objs, "-lrt"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    if not load:
        return so, stats
    return _bind(C.CDLL(str(so)), stats)


def _bind(lib: C.CDLL, stats: dict):
    from competesmoe_b200 import _lib
    bound = 0
    for name, (res, args) in _lib._SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is not None:
            fn.restype, fn.argtypes = res, args
            bound += 1
    stats["bound_symbols"] = bound
    return lib, stats


def source_digest() -> str:
    h = hashlib.sha256()
    for p in [CSRC / s for s in SOURCES] + sorted(CSRC.glob("*.h")) + sorted(SIMT.glob("*")):
        h.update(p.read_bytes())
    return h.hexdigest()[:16]
