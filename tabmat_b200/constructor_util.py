"""Dense / sparse column partition of a CSC matrix (reference: constructor_util.py:11-49)."""

from __future__ import annotations

from collections.abc import Sequence
from typing import Optional

import numpy as np
import scipy.sparse as sps

from .dense_matrix import DenseMatrix
from .sparse_matrix import SparseMatrix


def _split_sparse_and_dense_parts(
    arg1: sps.csc_matrix,
    threshold: float = 0.1,
    column_names: Optional[Sequence[Optional[str]]] = None,
    term_names: Optional[Sequence[Optional[str]]] = None,
):
    """(DenseMatrix of the columns denser than ``threshold``, SparseMatrix of the rest, and the
    two column-index vectors)."""
    if not isinstance(arg1, sps.csc_matrix):
        raise TypeError(
            f"X must be of type scipy.sparse.csc_matrix or matrix.SparseMatrix,not {type(arg1)}"
        )
    if not 0 <= threshold <= 1:
        raise ValueError("Threshold must be between 0 and 1.")
    n_rows, n_cols = arg1.shape
    density = np.diff(arg1.indptr) / n_rows
    dense_cols = np.flatnonzero(density > threshold)
    sparse_cols = np.setdiff1d(np.arange(n_cols), dense_cols)
    names = [None] * n_cols if column_names is None else list(column_names)
    terms = names if term_names is None else list(term_names)
    dense = DenseMatrix(
        arg1[:, dense_cols].toarray(),   # row-major: the layout the device kernels stream
        column_names=[names[i] for i in dense_cols],
        term_names=[terms[i] for i in dense_cols],
    )
    sparse = SparseMatrix(
        arg1[:, sparse_cols],
        column_names=[names[i] for i in sparse_cols],
        term_names=[terms[i] for i in sparse_cols],
    )
    return dense, sparse, dense_cols, sparse_cols
