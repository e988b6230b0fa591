"""Build ``libtabmat_b200.so`` (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Usage: ``python -m tabmat_b200.build [--force] [--verbose]``.
The shared object is written next to this file so that it travels with the repo snapshot.
"""

from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libtabmat_b200.so"
OBJ = PKG / "_build"
SOURCES = ["common.cu", "dense.cu", "dense_tc.cu", "sparse.cu", "categorical.cu", "split_fused.cu", "split_index.cu", "split.cu"]
HEADERS = [CSRC / "tm_common.cuh", PKG.parent / "include" / "tabmat_b200.h"]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """(Re)build the library when a source is newer than it.  Safe under one-process-per-GPU
    launches: an exclusive file lock serialises the ranks (the first one builds, the others find
    everything up to date), objects and the library are written to temporary names and moved
    into place atomically, so no rank can ever load a half-written file."""
    import fcntl

    OBJ.mkdir(exist_ok=True)
    with open(OBJ / ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> Path:
    jobs = []
    for src in SOURCES:
        s = CSRC / src
        o = OBJ / (src + ".o")
        if force or _stale(o, [s, *HEADERS]):
            cmd = [NVCC, *FLAGS, "-c", str(s), "-o", str(o) + ".tmp"]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose:
            print(r.stdout + r.stderr)
        out = cmd[-1]
        if out.endswith(".tmp"):
            os.replace(out, out[:-4])

    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            list(ex.map(run, jobs))
    objs = [str(OBJ / (s + ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", *objs, "-lcudart", "-o", str(LIB) + ".tmp"]
        run(cmd)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print("built", p)
