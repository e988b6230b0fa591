"""Constructors that decide the block layout of a SplitMatrix (reference: constructor.py:29-302).

``from_csc`` and ``from_pandas`` / ``from_df`` keep the reference's layout rules — numeric
columns denser than ``sparse_threshold`` go to the dense block, the rest to the sparse block,
categorical columns with at least ``cat_threshold`` levels become CategoricalMatrix blocks,
smaller ones are one-hot expanded — because the layout decides which kernels run.  Only pandas
data frames are ingested here (the reference goes through narwhals to also accept polars etc.;
that front end is out of scope); formulas are not supported."""

from __future__ import annotations

import warnings

import numpy as np
import scipy.sparse as sps

from .categorical_matrix import CategoricalMatrix
from .constructor_util import _split_sparse_and_dense_parts
from .dense_matrix import DenseMatrix
from .sparse_matrix import SparseMatrix
from .split_matrix import SplitMatrix

try:
    import pandas as pd
except Exception:  # pragma: no cover
    pd = None  # type: ignore


def from_csc(mat: sps.csc_matrix, threshold=0.1, column_names=None, term_names=None):
    """CSC matrix -> SplitMatrix of a dense and a sparse block (constructor.py:294-302)."""
    dense, sparse, dense_idx, sparse_idx = _split_sparse_and_dense_parts(
        mat, threshold, column_names, term_names)
    return SplitMatrix([dense, sparse], [dense_idx, sparse_idx])


def _shift_categorical_positions(indices, is_cat, next_col):
    """cat_position='end': categorical blocks are numbered after all other columns
    (constructor.py:283-291)."""
    out = []
    for idx, cat in zip(indices, is_cat):
        if cat:
            out.append(np.asarray(idx) + next_col)
            next_col += len(idx)
        else:
            out.append(idx)
    return out


def from_pandas(
    df,
    dtype=np.float64,
    sparse_threshold: float = 0.1,
    cat_threshold: int = 4,
    object_as_cat: bool = False,
    cat_position: str = "expand",
    drop_first: bool = False,
    categorical_format: str = "{name}[{category}]",
    cat_missing_method: str = "fail",
    cat_missing_name: str = "(MISSING)",
):
    """pandas DataFrame -> SplitMatrix (or a single block); same arguments and layout rules as
    ``tabmat.from_pandas`` / ``tabmat.from_df`` (constructor.py:30-280)."""
    if pd is None:
        raise ModuleNotFoundError("from_pandas requires pandas")
    if cat_position not in ("expand", "end"):
        raise ValueError("cat_position must be 'expand' or 'end'")
    matrices, indices, is_cat = [], [], []
    dense_src, dense_pos, sparse_src, sparse_pos, ignored = [], [], [], [], []
    next_col = 0
    for j, name in enumerate(df.columns):
        col = df.iloc[:, j]
        if object_as_cat and (col.dtype == object or isinstance(col.dtype, pd.StringDtype)):
            col = col.astype("category")
        if isinstance(col.dtype, pd.SparseDtype):
            sparse_src.append(j)
            sparse_pos.append(next_col)
            next_col += 1
        elif isinstance(col.dtype, pd.CategoricalDtype):
            cat = CategoricalMatrix(
                col, drop_first=drop_first, dtype=dtype, column_name=name, term_name=name,
                column_name_format=categorical_format, cat_missing_method=cat_missing_method,
                cat_missing_name=cat_missing_name)
            if len(cat.categories) < cat_threshold:
                dense, sparse, d_idx, s_idx = _split_sparse_and_dense_parts(
                    sps.csc_matrix(cat.tocsr(), dtype=dtype), threshold=sparse_threshold,
                    column_names=cat.get_names("column"), term_names=cat.get_names("term"))
                matrices += [dense, sparse]
                is_cat += [True, True]
                if cat_position == "expand":
                    indices += [next_col + d_idx, next_col + s_idx]
                    next_col += len(d_idx) + len(s_idx)
                else:
                    indices += [d_idx, s_idx]
            else:
                matrices.append(cat)
                is_cat.append(True)
                if cat_position == "expand":
                    indices.append(next_col + np.arange(cat.shape[1]))
                    next_col += cat.shape[1]
                else:
                    indices.append(np.arange(cat.shape[1]))
        elif pd.api.types.is_bool_dtype(col.dtype) or pd.api.types.is_numeric_dtype(col.dtype):
            values = col.to_numpy()
            if (values != 0).mean() <= sparse_threshold:
                sparse_src.append(j)
                sparse_pos.append(next_col)
            else:
                dense_src.append(j)
                dense_pos.append(next_col)
            next_col += 1
        else:
            ignored.append(name)
    if ignored:
        warnings.warn(f"Columns {ignored} were ignored. Make sure they have a valid dtype.")
    names = np.asarray(df.columns)
    if dense_src:
        matrices.append(DenseMatrix(
            df.iloc[:, dense_src].to_numpy().astype(dtype, copy=False),
            column_names=names[dense_src], term_names=names[dense_src]))
        indices.append(np.asarray(dense_pos))
        is_cat.append(False)
    if sparse_src:
        block = df.iloc[:, sparse_src]
        coo = (block.sparse.to_coo() if all(isinstance(t, pd.SparseDtype) for t in block.dtypes)
               else sps.coo_matrix(block.to_numpy(dtype=dtype)))
        matrices.append(SparseMatrix(sps.csc_matrix(coo, dtype=dtype), dtype=dtype,
                                     column_names=names[sparse_src], term_names=names[sparse_src]))
        indices.append(np.asarray(sparse_pos))
        is_cat.append(False)
    if cat_position == "end":
        indices = _shift_categorical_positions(indices, is_cat, next_col)
    if len(matrices) > 1:
        return SplitMatrix(matrices, indices)
    if not matrices:
        raise ValueError("DataFrame contained no valid column")
    return matrices[0]


from_df = from_pandas
