"""Compile ``oracle/tabmat_oracle.c`` (the CPU restatement) into ``oracle/_build/`` with plain gcc.

TEST INFRASTRUCTURE ONLY — see the header of tabmat_oracle.c.  ``-O2 -ffp-contract=off``: no
fused multiply-add contraction, so the restatement's arithmetic is exactly what the C text says.
"""

from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "tabmat_oracle.c"
OUT_DIR = HERE / "_build"
LIB = OUT_DIR / "libtabmat_oracle.so"
CC = os.environ.get("TABMAT_ORACLE_CC", "/usr/bin/gcc")


def build(force: bool = False) -> Path:
    OUT_DIR.mkdir(exist_ok=True)
    if LIB.exists() and not force and LIB.stat().st_mtime >= SRC.stat().st_mtime:
        return LIB
    cmd = [CC, "-O2", "-ffp-contract=off", "-march=x86-64-v2", "-shared", "-fPIC", "-o", str(LIB),
           str(SRC)]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print("built", build(force="--force" in sys.argv))
