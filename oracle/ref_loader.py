"""Load the reference's compiled CPU kernels from ``oracle/_ref`` (TEST INFRASTRUCTURE ONLY).

Two entry points:

* ``load_ext()`` — imports the four compiled extension modules (``dense``, ``sparse``,
  ``categorical``, ``split``) straight from ``oracle/_ref/*.so``.  Works on the GPU box
  (``/root/reference`` is not needed at run time).
* ``import_reference_package()`` — THIS CONTAINER ONLY: makes the reference's whole Python
  package importable (``import tabmat``) by overlaying symlinks to the read-only sources
  in ``/root/reference/src/tabmat`` with the ``.so`` files of ``oracle/_ref`` in a scratch
  directory under /tmp, and stubbing the absent ``formulaic`` / ``interface_meta``
  dependencies.  Used by ``tests/golden/make_golden.py`` to generate fixtures and to run
  the reference's own test-suite against the oracle build.

Nothing under ``tabmat_b200/`` may import this module.
"""

from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import sysconfig
import types
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
MODULES = ("dense", "sparse", "categorical", "split")
_EXT_CACHE: types.SimpleNamespace | None = None


def available() -> bool:
    suf = sysconfig.get_config_var("EXT_SUFFIX")
    return all((REF_DIR / f"{m}{suf}").exists() for m in MODULES)


def load_ext() -> types.SimpleNamespace:
    """Return a namespace with the reference's compiled ext modules."""
    global _EXT_CACHE
    if _EXT_CACHE is not None:
        return _EXT_CACHE
    if not available():
        raise ImportError(
            "oracle/_ref is not built; run `python oracle/build_ref.py` in the build container"
        )
    suf = sysconfig.get_config_var("EXT_SUFFIX")
    ns = types.SimpleNamespace()
    for m in MODULES:
        name = f"tabmat_ref_ext.{m}"
        path = str(REF_DIR / f"{m}{suf}")
        loader = importlib.machinery.ExtensionFileLoader(name, path)
        spec = importlib.util.spec_from_file_location(name, path, loader=loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        setattr(ns, m, mod)
    _EXT_CACHE = ns
    return ns


class _AnyMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __getitem__(cls, item):
        return cls

    def __or__(cls, other):
        return cls

    def __ror__(cls, other):
        return cls


class _Anything(metaclass=_AnyMeta):
    """Permissive stand-in class for symbols of absent optional dependencies."""

    def __init__(self, *a, **k):
        pass

    def __class_getitem__(cls, item):
        return cls

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return _Anything()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name in ("override", "stateful_transform"):
            return lambda f=None, *a, **k: f if callable(f) else (lambda g: g)
        return type(name, (_Anything,), {})


_STUBS = (
    "formulaic",
    "formulaic.errors",
    "formulaic.materializers",
    "formulaic.materializers.types",
    "formulaic.materializers.base",
    "formulaic.parser",
    "formulaic.parser.types",
    "formulaic.transforms",
    "formulaic.utils",
    "formulaic.utils.layered_mapping",
    "formulaic.utils.null_handling",
    "interface_meta",
)


def package_available() -> bool:
    """Is the stock reference package installed under ``oracle/_ref/tabmat``
    (``oracle/build_ref.py:install_package``)?  True on the GPU box too."""
    suf = sysconfig.get_config_var("EXT_SUFFIX")
    pkg = REF_DIR / "tabmat"
    return (pkg / "split_matrix.py").exists() and all(
        (pkg / "ext" / f"{m}{suf}").exists() for m in MODULES)


def set_omp_threads(n: int) -> int:
    """Make the OpenMP runtime the reference kernels link against use ``n`` threads, whatever
    OMP_NUM_THREADS says (torch.distributed.run exports OMP_NUM_THREADS=1 to every rank when
    nproc > 1).  Returns the thread count in force afterwards."""
    import ctypes

    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(int(n))
        return int(gomp.omp_get_max_threads())
    except OSError:
        return n


def import_installed_package():
    """``import tabmat`` = the unmodified reference package from ``oracle/_ref/tabmat`` (works
    wherever ``oracle/_ref`` travelled to; the absent formulaic / interface_meta dependencies
    are stubbed, which only disables ``from_formula``)."""
    if "tabmat" in sys.modules:
        return sys.modules["tabmat"]
    if not package_available():
        raise ImportError("oracle/_ref/tabmat is not installed (python oracle/build_ref.py)")
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = _StubModule(name)
    sys.path.insert(0, str(REF_DIR))
    import tabmat  # noqa: E402

    return tabmat


def import_reference_package(reference_root: str = "/root/reference"):
    """Import the reference ``tabmat`` package against the oracle build (container only)."""
    if "tabmat" in sys.modules:
        return sys.modules["tabmat"]
    src = Path(reference_root) / "src" / "tabmat"
    if not src.exists():
        raise ImportError(f"{src} not present (only available in the build container)")
    if not available():
        raise ImportError("oracle/_ref is not built")
    overlay = Path("/tmp/tabmat_ref_overlay")
    pkg = overlay / "tabmat"
    (pkg / "ext").mkdir(parents=True, exist_ok=True)
    for f in src.iterdir():
        if f.suffix == ".py":
            dst = pkg / f.name
            if not dst.exists():
                os.symlink(f, dst)
    bench = pkg / "benchmark"
    if not bench.exists() and (src / "benchmark").exists():
        os.symlink(src / "benchmark", bench)
    (pkg / "ext" / "__init__.py").touch()
    suf = sysconfig.get_config_var("EXT_SUFFIX")
    for m in MODULES:
        dst = pkg / "ext" / f"{m}{suf}"
        if dst.is_symlink() or dst.exists():
            dst.unlink()
        os.symlink(REF_DIR / f"{m}{suf}", dst)
    for name in (
        "formulaic",
        "formulaic.errors",
        "formulaic.materializers",
        "formulaic.materializers.types",
        "formulaic.materializers.base",
        "formulaic.parser",
        "formulaic.parser.types",
        "formulaic.transforms",
        "formulaic.utils",
        "formulaic.utils.layered_mapping",
        "formulaic.utils.null_handling",
        "interface_meta",
    ):
        if name not in sys.modules:
            sys.modules[name] = _StubModule(name)
    sys.path.insert(0, str(overlay))
    import tabmat  # noqa: E402

    return tabmat
