"""The C-ABI shared library loads on a CPU-only box and exports every symbol that
include/tabmat_b200.h declares (no compute calls here)."""

import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "tabmat_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from tabmat_b200 import _lib

    names = _declared()
    assert len(names) >= 40
    dll = ctypes.CDLL(str(_lib.LIB_PATH))
    missing = [n for n in names if not hasattr(dll, n)]
    assert not missing, missing
    assert set(_lib.EXPORTED) == set(names)


def test_version_and_error_string():
    from tabmat_b200 import _lib

    assert _lib.lib.tm_version() >= 100
    assert isinstance(_lib.lib.tm_last_error(), bytes)


def test_block_descriptor_layout_matches_the_library():
    from tabmat_b200 import _lib

    assert _lib.lib.tm_sizeof_block_desc() == ctypes.sizeof(_lib.BlockDesc) == 168
    descs = (_lib.BlockDesc * 2)()
    descs[0].kind, descs[0].ncols = 0, 5          # dense 5 columns
    descs[1].kind, descs[1].ncols = 2, 3          # categorical 3 columns
    # self blocks 25 + 3 (diagonal), cross block 15, each padded to a multiple of 4 elements
    assert _lib.lib.tm_split_workspace_elems(descs, 2) == 28 + 16 + 4
    assert _lib.lib.tm_split_workspace_head_elems(descs, 2) == 28 + 16


def test_product_path_does_not_touch_the_oracle():
    """Nothing under tabmat_b200/ may import or reference oracle/ (test infrastructure)."""
    pat = re.compile(r"(import|from)\s+oracle|c_oracle|ref_loader|libtabmat_oracle|oracle/_ref")
    for f in list((ROOT / "tabmat_b200").rglob("*.py")) + list((ROOT / "tabmat_b200" / "csrc").glob("*")):
        assert not pat.search(f.read_text()), f
