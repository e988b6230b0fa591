"""Build the reference's own CPU kernels into ``oracle/_ref/`` (TEST INFRASTRUCTURE ONLY).

Nothing shipped by ``tabmat_b200`` imports this.  Only ``tests/``,
``__graft_entry__`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg
use the result, and there only as the checker / the timed baseline.

What it does (reference ``setup.py:19-47,100-171`` is the recipe it replaces;
we do NOT run the reference's build system):

1. renders the three Mako templates ``src/tabmat/ext/*-tmpl.cpp`` straight from
   ``/root/reference`` with ``oracle/mini_mako.py`` into a scratch dir in /tmp;
2. cythonizes the four ``.pyx`` files from where they lie (copies of the .pyx
   are made in the scratch dir only, never in the repo);
3. compiles with ``/usr/bin/g++ -O3 -fopenmp -std=c++17 -march=x86-64-v3`` against
   the stand-in ``xsimd`` / ``jemalloc`` headers in ``oracle/shims`` (the real ones
   are not installed and there is no network);
4. writes ONLY the four resulting ``.so`` files to ``oracle/_ref/``.

``oracle/_ref`` is git-ignored but travels to the GPU box.  ``-march=x86-64-v3``
(AVX2+FMA) instead of ``-march=native`` so the objects run on the GPU box's host CPU.

Caveat recorded in DESIGN.md: kernels are the reference's source; the SIMD and
allocator layer underneath is a stand-in, so CPU timings are labelled
"reference source, stand-in xsimd/jemalloc".
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_ROOT = Path(os.environ.get("TABMAT_REFERENCE", "/root/reference"))
REF_EXT = REF_ROOT / "src" / "tabmat" / "ext"
OUT = HERE / "_ref"
MODULES = ["dense", "sparse", "categorical", "split"]
TEMPLATES = ["dense_helpers", "sparse_helpers", "cat_split_helpers"]

CXX = os.environ.get("TABMAT_ORACLE_CXX", "/usr/bin/g++")
MARCH = os.environ.get("TABMAT_ORACLE_MARCH", "x86-64-v3")

# reference setup.py:162-171 (debug_build False)
DIRECTIVES = {
    "language_level": "3",
    "boundscheck": False,
    "wraparound": False,
    "initializedcheck": False,
    "nonecheck": False,
    "cdivision": True,
    "cpow": True,
    "legacy_implicit_noexcept": True,
}


def so_name(mod: str) -> str:
    return mod + sysconfig.get_config_var("EXT_SUFFIX")


def is_built() -> bool:
    return all((OUT / so_name(m)).exists() for m in MODULES)


def build(force: bool = False, verbose: bool = True) -> bool:
    """Return True if oracle/_ref is usable afterwards."""
    if is_built() and not force:
        return True
    if not REF_EXT.exists():
        if verbose:
            print(f"[oracle] {REF_EXT} absent; cannot build oracle/_ref here")
        return False
    import numpy as np
    from Cython.Compiler.Main import CompilationOptions, compile_single
    from Cython.Compiler import Options  # noqa: F401

    sys.path.insert(0, str(HERE))
    import mini_mako

    OUT.mkdir(exist_ok=True)
    work = Path(tempfile.mkdtemp(prefix="tabmat_ref_build_"))
    try:
        for t in TEMPLATES:
            mini_mako.render_file(str(REF_EXT / f"{t}-tmpl.cpp"), str(work / f"{t}.cpp"))
        shutil.copy(REF_EXT / "alloc.h", work / "alloc.h")
        for m in MODULES:
            shutil.copy(REF_EXT / f"{m}.pyx", work / f"{m}.pyx")
        py_inc = sysconfig.get_paths()["include"]
        for m in MODULES:
            opts = CompilationOptions(
                cplus=True,
                compiler_directives=dict(DIRECTIVES),
                output_file=str(work / f"{m}_cy.cpp"),
            )
            res = compile_single(str(work / f"{m}.pyx"), opts, full_module_name=f"tabmat.ext.{m}")
            if res.num_errors:
                raise RuntimeError(f"cython failed on {m}.pyx")
            cmd = [
                CXX, "-shared", "-fPIC", "-O3", "-fopenmp", "-std=c++17", f"-march={MARCH}",
                "-DJEMALLOC_INSTALL_SUFFIX=", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
                "-w",
                f"-I{np.get_include()}", f"-I{py_inc}", f"-I{HERE / 'shims'}", f"-I{work}",
                str(work / f"{m}_cy.cpp"), "-o", str(OUT / so_name(m)),
            ]
            if verbose:
                print("[oracle]", " ".join(cmd[:8]), "...", m, flush=True)
            subprocess.run(cmd, check=True)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    return is_built()


PKG = OUT / "tabmat"


def package_installed() -> bool:
    return (PKG / "split_matrix.py").exists() and all(
        (PKG / "ext" / so_name(m)).exists() for m in MODULES)


def install_package(verbose: bool = True) -> bool:
    """Install the reference's stock Python package into ``oracle/_ref/tabmat`` (git-ignored,
    like ``pip install --target``; never committed): the unmodified ``src/tabmat/*.py`` and
    ``benchmark/`` files plus the extension modules built above, so that
    ``bench.py --impl reference`` runs the reference's own ``tabmat.SplitMatrix.sandwich``
    (split_matrix.py:324-356) on the GPU box, where ``/root/reference`` does not exist."""
    src = REF_ROOT / "src" / "tabmat"
    if not src.exists() or not is_built():
        return package_installed()
    (PKG / "ext").mkdir(parents=True, exist_ok=True)
    for f in src.iterdir():
        if f.suffix == ".py":
            shutil.copy(f, PKG / f.name)
    if (src / "benchmark").exists():
        shutil.copytree(src / "benchmark", PKG / "benchmark", dirs_exist_ok=True,
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    (PKG / "ext" / "__init__.py").write_text("")
    for m in MODULES:
        shutil.copy(OUT / so_name(m), PKG / "ext" / so_name(m))
    if verbose:
        print(f"[oracle] reference package installed into {PKG}")
    return package_installed()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    ok = ok and install_package()
    print("oracle/_ref built:", ok)
    sys.exit(0 if ok else 1)
