"""Build libcsmoe.so (the C-ABI CUDA library) in-tree for sm_100a.

`python -m competesmoe_b200.build` cross-compiles with nvcc (no GPU needed).  The .so lands in
`competesmoe_b200/lib/` so that it travels with the source tree to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
OBJDIR = PKG / "build"
LIB = LIBDIR / "libcsmoe.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-I", str(ROOT / "include"), "-I", str(CSRC),
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libcsmoe.so cannot be built")
    return exe


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest(paths: list[Path]) -> str:
    h = hashlib.sha256()
    for p in paths:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    # the flags without the checkout's absolute path: the tree is copied to another directory on the GPU box, and a
    # digest that changed there would rebuild the library on every run (and race between the ranks of one job)
    h.update(" ".join(f.replace(str(ROOT), "<root>") for f in NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = _sources()
    deps = srcs + sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "csmoe.h"]
    stamp = LIBDIR / "libcsmoe.stamp"
    digest = _digest(deps)
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB
    OBJDIR.mkdir(exist_ok=True)
    LIBDIR.mkdir(exist_ok=True)
    # one builder at a time (the ranks of a multi-GPU job import the package concurrently); whoever waited re-checks
    import fcntl
    with open(LIBDIR / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
            return LIB
        return _build_locked(srcs, deps, stamp, digest, force, verbose)


def _build_locked(srcs, deps, stamp, digest, force, verbose) -> Path:
    nvcc = _nvcc()

    headers = [p for p in deps if p not in srcs]

    def compile_one(src: Path) -> Path:
        obj = OBJDIR / (src.stem + ".o")
        # per-object stamp: a source is recompiled only when it, a header or the flags changed
        ostamp, odigest = OBJDIR / (src.stem + ".stamp"), _digest([src] + headers)
        if not force and not verbose and obj.exists() and ostamp.exists() and ostamp.read_text() == odigest:
            return obj
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        ostamp.write_text(odigest)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    tmp = LIB.with_suffix(f".so.tmp{os.getpid()}")
    cmd = [nvcc, "-shared", "-o", str(tmp), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)     # a process that is loading the library never sees a half-written file
    stamp.write_text(digest)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
