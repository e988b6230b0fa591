"""Mirror of ``tabmat/ext/split.pyx`` (reference: split.pyx:32-217).

``sandwich_cat_dense`` / ``sandwich_cat_cat`` / ``sandwich_cat_sparse`` launch device kernels;
``split_col_subsets`` and ``is_sorted`` are host-side bookkeeping exactly as in the reference
(they work on the small int64 column-index lists of a SplitMatrix)."""

from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import _dev
from .._lib import check, fn
from .dense import dense_layout
from .sparse import DeviceCSR


def sandwich_cat_dense(i_indices: torch.Tensor, i_ncol: int, d: torch.Tensor, mat_j: torch.Tensor,
                       rows: Optional[torch.Tensor], j_cols: Optional[torch.Tensor],
                       drop_first: bool = False) -> torch.Tensor:
    """res[i_indices[k]-drop_first, b] += d[k] * mat_j[k, j_cols[b]] for k in rows
    (sandwich_cat_dense, split.pyx:32-80)."""
    n, q, c_order = dense_layout(mat_j)
    nJ = q if j_cols is None else _dev.length(j_cols)
    out = torch.empty((i_ncol, nJ), dtype=mat_j.dtype, device=mat_j.device)
    if i_ncol == 0 or nJ == 0:
        return out
    if n == 0 or (rows is not None and _dev.length(rows) == 0):
        return out.zero_()
    check(fn("tm_cat_dense_sandwich", _dev.suffix(mat_j.dtype))(
        _dev.ptr(i_indices), n, i_ncol, int(drop_first), _dev.ptr(d), _dev.ptr(mat_j), q, c_order,
        _dev.ptr(rows), _dev.length(rows), _dev.ptr(j_cols), _dev.length(j_cols), _dev.ptr(out),
        _dev.stream_ptr()))
    return out


def sandwich_cat_cat(i_indices: torch.Tensor, j_indices: torch.Tensor, i_ncol: int, j_ncol: int,
                     d: torch.Tensor, rows: Optional[torch.Tensor], i_drop_first: bool = False,
                     j_drop_first: bool = False) -> torch.Tensor:
    """res[i_indices[k]-dfi, j_indices[k]-dfj] += d[k] for k in rows
    (sandwich_cat_cat, split.pyx:83-111)."""
    n = int(i_indices.numel())
    out = torch.empty((i_ncol, j_ncol), dtype=d.dtype, device=d.device)
    if i_ncol == 0 or j_ncol == 0:
        return out
    if n == 0 or (rows is not None and _dev.length(rows) == 0):
        return out.zero_()
    check(fn("tm_cat_cat_sandwich", _dev.suffix(d.dtype))(
        _dev.ptr(i_indices), _dev.ptr(j_indices), n, i_ncol, j_ncol, int(i_drop_first),
        int(j_drop_first), _dev.ptr(d), _dev.ptr(rows), _dev.length(rows), _dev.ptr(out),
        _dev.stream_ptr()))
    return out


def sandwich_cat_sparse(i_indices: torch.Tensor, i_ncol: int, d: torch.Tensor, A: DeviceCSR,
                        rows: Optional[torch.Tensor], s_cols: Optional[torch.Tensor],
                        drop_first: bool = False) -> torch.Tensor:
    """res[i_indices[k]-drop_first, s] += d[k] * A[k, s_cols[s]] for k in rows.

    The reference has no native function here: CategoricalMatrix._cross_sparse goes through
    scipy's csr_matmat (categorical_matrix.py:825-838)."""
    n, p = A.shape
    nS = p if s_cols is None else _dev.length(s_cols)
    out = torch.empty((i_ncol, nS), dtype=d.dtype, device=d.device)
    if i_ncol == 0 or nS == 0:
        return out
    if A.nnz == 0 or (rows is not None and _dev.length(rows) == 0):
        return out.zero_()
    check(fn("tm_cat_sparse_sandwich", _dev.suffix(d.dtype))(
        _dev.ptr(i_indices), n, i_ncol, int(drop_first), _dev.ptr(d), _dev.ptr(A.data),
        _dev.ptr(A.indices), _dev.ptr(A.indptr), _dev.ptr(A.row), p, A.nnz, _dev.ptr(rows),
        _dev.length(rows), _dev.ptr(s_cols), _dev.length(s_cols), _dev.ptr(out),
        _dev.stream_ptr()))
    return out


def dense_cross_sandwich(X: torch.Tensor, d: torch.Tensor, rows: Optional[torch.Tensor], cats,
                         A: Optional[DeviceCSR]):
    """All cross blocks that share the dense operand in ONE pass over X (tm_dense_cross_sandwich).

    ``cats`` is a list of ``(codes, n_cols, drop_first)``; returns ``(list of (n_cols_i, p)
    tensors, (p_sparse, p) tensor or None)``.  Fused form of the reference's per-pair loop
    (split_matrix.py:346-354 -> split.pyx:32-80 / sparse.pyx:211-260)."""
    import ctypes as C

    n, p, c_order = dense_layout(X)
    if not c_order:
        raise ValueError("dense_cross_sandwich needs a C-contiguous dense block")
    nc = len(cats)
    outs = [torch.empty((int(k), p), dtype=X.dtype, device=X.device) for (_, k, _) in cats]
    out_s = None
    if A is not None and A.shape[1] > 0:
        out_s = torch.empty((A.shape[1], p), dtype=X.dtype, device=X.device)
    codes_arr = (C.c_void_p * max(nc, 1))(*[c.data_ptr() for (c, _, _) in cats])
    out_arr = (C.c_void_p * max(nc, 1))(*[o.data_ptr() for o in outs])
    K_arr = (C.c_int64 * max(nc, 1))(*[int(k) for (_, k, _) in cats])
    df_arr = (C.c_int32 * max(nc, 1))(*[int(bool(f)) for (_, _, f) in cats])
    check(fn("tm_dense_cross_sandwich", _dev.suffix(X.dtype))(
        _dev.ptr(X), n, p, _dev.ptr(d), _dev.ptr(rows), _dev.length(rows), nc,
        C.cast(codes_arr, C.c_void_p), C.cast(K_arr, C.c_void_p), C.cast(df_arr, C.c_void_p),
        C.cast(out_arr, C.c_void_p),
        _dev.ptr(A.data) if out_s is not None else None,
        _dev.ptr(A.indices) if out_s is not None else None,
        _dev.ptr(A.indptr) if out_s is not None else None,
        A.shape[1] if out_s is not None else 0, _dev.ptr(out_s), _dev.stream_ptr()))
    return outs, out_s


def split_col_subsets(indices: list, cols: np.ndarray):
    """For sorted ``cols``: per block, the positions inside ``cols`` that fall into the block
    and the block-local column ids (split_col_subsets, split.pyx:157-209).  Host side."""
    cols = np.asarray(cols, dtype=np.int64)
    subset_cols_indices = []
    subset_cols = []
    for idx in indices:
        idx = np.asarray(idx, dtype=np.int64)
        # both idx and cols are sorted: positions of cols inside idx
        pos = np.searchsorted(idx, cols)
        pos_c = np.minimum(pos, max(len(idx) - 1, 0))
        hit = (pos < len(idx)) & (idx[pos_c] == cols) if len(idx) else np.zeros(len(cols), bool)
        subset_cols_indices.append(np.flatnonzero(hit).astype(np.int32))
        subset_cols.append(pos[hit].astype(np.int32))
    return subset_cols_indices, subset_cols, len(cols)


def is_sorted(a) -> bool:
    """Monotone non-decreasing check (is_sorted, split.pyx:211-217).  Host side."""
    a = np.asarray(a)
    return bool(np.all(a[1:] >= a[:-1])) if a.size > 1 else True
