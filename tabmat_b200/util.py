"""Argument set-up and checks shared by the matrix classes (reference: util.py:6-115).

Error types and messages follow the reference so that its tests read the same here."""

from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _dev


def set_up_rows_or_cols(arr, length: int) -> Optional[torch.Tensor]:
    """None stays None (= all; the reference materialises np.arange, util.py:6-14)."""
    return _dev.idx32(arr)


def setup_restrictions(shape, rows, cols):
    return _dev.idx32(rows), _dev.idx32(cols)


def is_unrestricted(idx, full_length: int) -> bool:
    """The reference treats a full-length index as 'all' without looking at it
    (dense_matrix.py:208-210, sparse_matrix.py:242-243)."""
    return idx is None or len(idx) == full_length


def _shape0(x) -> int:
    return int(x.shape[0])


def _check_out_shape(out, expected_first_dim: int) -> None:
    if out is not None and out.shape[0] != expected_first_dim:
        raise ValueError(
            f"""The first dimension of 'out' must be {expected_first_dim}, but it is
            {out.shape[0]}."""
        )


def check_transpose_matvec_out_shape(mat, out) -> None:
    _check_out_shape(out, mat.shape[1])


def check_matvec_out_shape(mat, out) -> None:
    _check_out_shape(out, mat.shape[0])


def check_matvec_dimensions(mat, vec, transpose: bool) -> None:
    match_dim = 0 if transpose else 1
    if mat.shape[match_dim] != vec.shape[0]:
        raise ValueError(
            f"shapes {tuple(mat.shape)} and {tuple(vec.shape)} not aligned: "
            f"{mat.shape[match_dim]} (dim {match_dim}) != {vec.shape[0]} (dim 0)"
        )


def _np_dtype_of(x) -> np.dtype:
    if isinstance(x, torch.Tensor):
        return _dev.np_dtype(x.dtype)
    return np.dtype(x.dtype)


def check_sandwich_compatible(mat, d) -> None:
    if mat.shape[0] != d.shape[0]:
        raise ValueError(
            f"shapes {tuple(mat.shape)} and {tuple(d.shape)} not aligned: "
            f"{mat.shape[0]} (dim 0) != {d.shape[0]} (dim 0)"
        )
    if not np.dtype(mat.dtype) == _np_dtype_of(d):
        raise TypeError(
            f"""self and d need to be of same dtype, either np.float64
            or np.float32. self is of type {mat.dtype}, while d is of type
            {_np_dtype_of(d)}."""
        )


def _check_indexer(indexer):
    """Canonicalise an indexer into (rows, cols) (reference util.py:70-115)."""
    if not isinstance(indexer, tuple):
        indexer = (indexer, slice(None, None, None))
    if len(indexer) > 2:
        raise ValueError("More than two indexers are not supported.")
    row_indexer, col_indexer = indexer
    if isinstance(row_indexer, slice):
        if isinstance(col_indexer, slice):
            return row_indexer, col_indexer
        col_indexer = np.asarray(col_indexer)
        if col_indexer.ndim > 1:
            raise ValueError("Indexing would result in a matrix with more than 2 dimensions.")
        return row_indexer, col_indexer.reshape(-1)
    if isinstance(col_indexer, slice):
        row_indexer = np.asarray(row_indexer)
        if row_indexer.ndim > 1:
            raise ValueError("Indexing would result in a matrix with more than 2 dimensions.")
        return row_indexer.reshape(-1), col_indexer
    row_indexer = np.asarray(row_indexer)
    col_indexer = np.asarray(col_indexer)
    if row_indexer.ndim <= 1 and col_indexer.ndim <= 1:
        return np.ix_(row_indexer.reshape(-1), col_indexer.reshape(-1))
    if (row_indexer.ndim == 2 and row_indexer.shape[1] == 1 and col_indexer.ndim == 2
            and col_indexer.shape[0] == 1):
        return row_indexer, col_indexer
    raise ValueError("This type of indexing is not supported.")
