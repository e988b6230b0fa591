"""ctypes binding of ``libtabmat_b200.so`` (the C-ABI declared in ``include/tabmat_b200.h``).

There is deliberately NO fallback: if the shared library is missing, importing this module
raises; if no CUDA device is present, every compute entry point raises.  The product path
never routes through ``oracle/`` or any CPU implementation.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libtabmat_b200.so"

c_f32p = C.c_void_p
c_i32p = C.c_void_p
c_i64 = C.c_int64
c_int = C.c_int
c_ptr = C.c_void_p


def _load() -> C.CDLL:
    if os.environ.get("TABMAT_B200_AUTOBUILD", "1") == "1":
        # (re)build when the library is missing or older than its sources; a no-op otherwise
        import importlib.util
        import shutil

        spec = importlib.util.spec_from_file_location("_tabmat_b200_build", _PKG / "build.py")
        _build = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_build)
        if shutil.which(_build.NVCC) or os.path.exists(_build.NVCC):
            _build.build()
    if not LIB_PATH.exists():
        if True:
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m tabmat_b200.build` "
                "(tabmat_b200 has no CPU fallback)"
            )
    return C.CDLL(str(LIB_PATH))


lib = _load()

# name -> argtypes (restype is int unless listed in _RESTYPES)
P = c_ptr
I = c_i64  # noqa: E741
N = c_int
_SIGS = {
    "tm_dense_sandwich": [P, I, I, N, P, P, I, P, I, P, P],
    "tm_dense_matvec": [P, I, I, N, P, P, I, P, I, P, N, P],
    "tm_dense_rmatvec": [P, I, I, N, P, P, I, P, I, P, P],
    "tm_dense_sq_dot_weights": [P, I, I, N, P, P, P, P],
    "tm_sparse_sandwich": [P, P, P, P, I, I, I, P, P, I, P, I, P, P],
    "tm_csr_dense_sandwich": [P, P, P, I, I, P, I, N, P, P, I, P, I, P, I, P, P],
    "tm_csc_dense_gather_sandwich": [P, P, P, I, I, P, I, P, P, P],
    "tm_csr_matvec": [P, P, P, I, I, P, P, I, P, I, P, N, P],
    "tm_csc_rmatvec": [P, P, P, I, I, P, P, I, P, I, P, N, P],
    "tm_csc_sq_dot_weights": [P, P, P, I, I, P, P, P],
    "tm_cat_sandwich": [P, I, P, P, I, I, N, P, P],
    "tm_cat_transpose_matvec": [P, I, P, P, I, P, I, I, N, P, P],
    "tm_cat_matvec": [P, I, P, P, I, I, N, P, P],
    "tm_cat_to_csr": [P, I, N, P, P, P, P, P],
    "tm_cat_segment_sum": [P, P, P, P, I, P, N, P],
    "tm_cat_dense_sandwich": [P, I, I, N, P, P, I, N, P, I, P, I, P, P],
    "tm_cat_cat_sandwich": [P, P, I, I, I, N, N, P, P, I, P, P],
    "tm_cat_sparse_sandwich": [P, I, I, N, P, P, P, P, P, I, I, P, I, P, I, P, P],
    "tm_dense_cross_sandwich": [P, I, I, P, P, I, N, P, P, P, P, P, P, P, I, P, P],
    "tm_split_sandwich_blocks": [P, N, I, P, P, I, P, P],
    "tm_split_sandwich_rmatvec_blocks": [P, N, I, P, P, P, I, P, P, P],
    "tm_split_sandwich_assemble": [P, N, P, P, I, P],
    "tm_split_sandwich_assemble_band": [P, N, P, P, I, I, I, P],
    "tm_split_sandwich_assemble_part_band": [P, N, P, P, I, N, I, I, P],
    "tm_split_sandwich_blocks_part": [P, N, I, P, P, I, P, N, P],
    "tm_split_sandwich_assemble_part": [P, N, P, P, I, N, P],
    "tm_scatter_block": [P, I, I, P, P, P, I, N, P],
    "tm_scatter_diag": [P, I, P, P, I, P],
    "tm_std_sandwich_combine": [P, N, N, P, P, P, P, I, P, P],
    "tm_permute_gather": [P, P, I, P, N, P],
    "tm_permute_scatter": [P, P, I, P, N, P],
}

#: every symbol include/tabmat_b200.h declares (checked by tests/test_capi_symbols.py)
EXPORTED = ["tm_version", "tm_last_error", "tm_launch_count", "tm_reset_launch_count",
            "tm_has_tcgen05", "tm_set_dense_f32_mode", "tm_set_cross_runs_mode",
            "tm_set_tc_scatter_warps", "tm_set_tc_round_mode", "tm_set_tc_flush_steps", "tm_set_sm_reserve",
            "tm_set_tc_sm_reserve",
            "tm_split_workspace_elems", "tm_split_workspace_head_elems", "tm_memcpy2d_to_host",
            "tm_dense_onehot_sandwich_f32", "tm_split_profile_enable", "tm_split_profile_read", "tm_sizeof_block_desc",
            "tm_split_last_plan"]
for _name, _args in _SIGS.items():
    for _suf in ("f32", "f64"):
        _fn = getattr(lib, f"{_name}_{_suf}")
        _fn.argtypes = _args
        _fn.restype = c_int
        EXPORTED.append(f"{_name}_{_suf}")



class BlockDesc(C.Structure):
    """ctypes mirror of ``tm_block_desc`` (include/tabmat_b200.h)."""

    _fields_ = [("kind", C.c_int32), ("c_order", C.c_int32), ("drop_first", C.c_int32),
                ("flags", C.c_int32), ("ncols", C.c_int64), ("data", C.c_void_p),
                ("csr_indices", C.c_void_p), ("csr_indptr", C.c_void_p), ("csr_row", C.c_void_p),
                ("nnz", C.c_int64), ("col_index", C.c_void_p), ("cat_perm", C.c_void_p),
                ("cat_segptr", C.c_void_p), ("cat_nvalid", C.c_int64), ("csc_data", C.c_void_p),
                ("csc_indices", C.c_void_p), ("csc_indptr", C.c_void_p), ("csc_row_blocks", C.c_int64),
                ("csc_cat_codes", C.c_void_p), ("gcsc_data", C.c_void_p),
                ("gcsc_indices", C.c_void_p), ("gcsc_indptr", C.c_void_p),
                ("gcsc_row_blocks", C.c_int64)]

#: rows per block of the row-blocked CSC copy (TM_CSC_ROW_BLOCK in include/tabmat_b200.h)
CSC_ROW_BLOCK = 1 << 20


lib.tm_dense_onehot_sandwich_f32.argtypes = [P, I, I, P, P, I, N, P, P, P, P, P, P]
lib.tm_dense_onehot_sandwich_f32.restype = c_int
lib.tm_sizeof_block_desc.restype = c_i64
if lib.tm_sizeof_block_desc() != C.sizeof(BlockDesc):  # pragma: no cover
    raise ImportError("tm_block_desc layout mismatch between libtabmat_b200.so and _lib.BlockDesc")
lib.tm_split_last_plan.restype = c_int
lib.tm_split_profile_enable.argtypes = [c_int]
lib.tm_split_profile_enable.restype = None
lib.tm_split_profile_read.argtypes = [C.c_void_p]
lib.tm_split_profile_read.restype = c_int
lib.tm_split_workspace_elems.argtypes = [C.c_void_p, c_int]
lib.tm_split_workspace_elems.restype = c_i64
lib.tm_split_workspace_head_elems.argtypes = [C.c_void_p, c_int]
lib.tm_split_workspace_head_elems.restype = c_i64
lib.tm_version.restype = c_int
lib.tm_last_error.restype = C.c_char_p
lib.tm_launch_count.restype = c_i64
lib.tm_reset_launch_count.restype = None
lib.tm_has_tcgen05.restype = c_int
lib.tm_set_dense_f32_mode.argtypes = [c_int]
lib.tm_set_dense_f32_mode.restype = None
lib.tm_memcpy2d_to_host.argtypes = [P, I, P, I, I, I, P]
lib.tm_memcpy2d_to_host.restype = c_int
lib.tm_set_sm_reserve.argtypes = [c_int]
lib.tm_set_sm_reserve.restype = None
lib.tm_set_tc_sm_reserve.argtypes = [c_int]
lib.tm_set_tc_sm_reserve.restype = None
lib.tm_set_tc_scatter_warps.argtypes = [c_int]
lib.tm_set_tc_scatter_warps.restype = None
lib.tm_set_tc_round_mode.argtypes = [c_int]
lib.tm_set_tc_round_mode.restype = None
lib.tm_set_tc_flush_steps.argtypes = [c_int]
lib.tm_set_tc_flush_steps.restype = None
lib.tm_set_cross_runs_mode.argtypes = [c_int]
lib.tm_set_cross_runs_mode.restype = None
# TABMAT_B200_CROSS_RUNS: 0 auto (run-aggregating cross kernel for row-sorted matrices) |
# 1 always | 2 never
if os.environ.get("TABMAT_B200_CROSS_RUNS"):
    lib.tm_set_cross_runs_mode(int(os.environ["TABMAT_B200_CROSS_RUNS"]))
# TABMAT_B200_DENSE_F32_MODE: 0 auto (tcgen05 TF32 when eligible) | 1 CUDA-core only (exact fp32)
# | 2 force tcgen05 | 3 fp32-accurate 3xTF32 on the tensor cores
if os.environ.get("TABMAT_B200_DENSE_F32_MODE"):
    lib.tm_set_dense_f32_mode(int(os.environ["TABMAT_B200_DENSE_F32_MODE"]))


#: opt-in deterministic mode (set_deterministic): categorical transpose_matvec / sandwich add in
#: a fixed order (tm_cat_segment_sum) instead of with atomics; TABMAT_B200_DETERMINISTIC=1
_DETERMINISTIC = os.environ.get("TABMAT_B200_DETERMINISTIC") == "1"


def set_deterministic(on: bool) -> None:
    """Bit-reproducible categorical ``transpose_matvec`` / ``sandwich`` (the operations the
    reference made deterministic, CHANGELOG.rst:134).  Every other kernel adds partial sums with
    float atomics / REDs in arrival order and stays run-to-run reproducible only to rounding
    (DESIGN.md "Numerics")."""
    global _DETERMINISTIC
    _DETERMINISTIC = bool(on)


def deterministic() -> bool:
    return _DETERMINISTIC


class TabmatB200Error(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        raise TabmatB200Error(lib.tm_last_error().decode("utf-8", "replace"))


def fn(name: str, dtype_suffix: str):
    return getattr(lib, f"{name}_{dtype_suffix}")


def launch_count() -> int:
    return int(lib.tm_launch_count())


def reset_launch_count() -> None:
    lib.tm_reset_launch_count()
