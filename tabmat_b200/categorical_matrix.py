"""CategoricalMatrix on the device (reference: categorical_matrix.py:293-980).

A one-hot encoded column stored as its int32 category codes in HBM (missing = -1).  All
numeric methods are segmented reductions / gathers over the code vector:

* ``sandwich``            diag[c_k] += d_k                       (tm_cat_sandwich)
* ``transpose_matvec``    out[c_k]  += v_k                       (tm_cat_transpose_matvec)
* ``matvec``              out[k]    += v[c_k]                    (tm_cat_matvec)
* cross with dense        res[c_k, :] += d_k * Y[k, :]           (tm_cat_dense_sandwich)
* cross with categorical  res[ci_k, cj_k] += d_k                 (tm_cat_cat_sandwich)
* cross with sparse       res[c_k, j] += d_k * A[k, j]           (tm_cat_sparse_sandwich; the
  reference goes through scipy here, categorical_matrix.py:825-838)

Ingestion from data frames is out of scope: the constructor accepts
array-likes / pandas categoricals, or integer codes plus ``categories``."""

from __future__ import annotations

import re
from typing import Optional, Union

import numpy as np
import torch
from scipy import sparse as sps

from . import _dev, _lib
from ._lib import check, fn
from .dense_matrix import DenseMatrix, _accumulate_out
from .ext import categorical as ext_cat
from .ext import split as ext_split
from .matrix_base import MatrixBase, _vec_in
from .sparse_matrix import SparseMatrix
from .util import (
    _check_indexer,
    check_matvec_dimensions,
    check_matvec_out_shape,
    check_sandwich_compatible,
    check_transpose_matvec_out_shape,
    is_unrestricted,
)

try:  # pandas is optional, exactly as in the reference
    import pandas as pd
except Exception:  # pragma: no cover
    pd = None  # type: ignore


def _is_indexer_full_length(full_length: int, indexer):
    if isinstance(indexer, np.ndarray):
        if (indexer > full_length - 1).any():
            raise IndexError("Index out-of-range.")
        return np.array_equal(indexer.ravel(), np.arange(full_length))
    elif isinstance(indexer, slice):
        return len(range(*indexer.indices(full_length))) == full_length


def _factorize(x: np.ndarray):
    """Sorted-unique factorisation with -1 for missing (categorical_matrix.py:221-228)."""
    x = np.asarray(x)
    if x.dtype.kind == "f":
        na_mask = x != x
    elif x.dtype == object:
        na_mask = np.array([(v is None) or (v != v) for v in x], dtype=bool)
    else:
        na_mask = np.zeros(x.shape, dtype=bool)
    categories, indices_nona = np.unique(x[~na_mask], return_inverse=True)
    indices = np.full(x.shape, -1, dtype=np.int32)
    indices[~na_mask] = indices_nona
    return indices, categories


def _extract_codes_and_categories(cat_vec):
    if pd is not None:
        if isinstance(cat_vec, pd.Categorical):
            return np.asarray(cat_vec.codes), cat_vec.categories.to_numpy()
        if isinstance(cat_vec, pd.Series):
            if isinstance(cat_vec.dtype, pd.CategoricalDtype):
                return cat_vec.cat.codes.to_numpy(), cat_vec.cat.categories.to_numpy()
            indices, categories = pd.factorize(cat_vec, sort=True)
            return indices, np.asarray(categories)
        indices, categories = pd.factorize(np.asarray(cat_vec), sort=True)
        return indices, np.asarray(categories)
    return _factorize(np.asarray(cat_vec))


def _row_col_indexing(arr, rows, cols):
    """Subset a result block by rows/cols lists (categorical_matrix.py:296-317)."""
    if isinstance(rows, slice) and rows == slice(None, None, None):
        rows = None
    if isinstance(cols, slice) and cols == slice(None, None, None):
        cols = None
    is_row_indexed = not (rows is None or len(rows) == arr.shape[0])
    is_col_indexed = not (cols is None or len(cols) == arr.shape[1])
    if isinstance(arr, torch.Tensor):
        if is_row_indexed:
            arr = arr.index_select(0, _dev.idx32(rows).to(torch.int64))
        if is_col_indexed:
            arr = arr.index_select(1, _dev.idx32(cols).to(torch.int64))
        return arr
    if is_row_indexed and is_col_indexed:
        return arr[np.ix_(rows, cols)]
    elif is_row_indexed:
        return arr[rows]
    elif is_col_indexed:
        return arr[:, cols]
    return arr


class CategoricalMatrix(MatrixBase):
    """One-hot encoded categorical column; same API as ``tabmat.CategoricalMatrix``."""

    def __init__(
        self,
        cat_vec,
        categories: Optional[np.ndarray] = None,
        drop_first: bool = False,
        dtype=np.float64,
        column_name: Optional[str] = None,
        term_name: Optional[str] = None,
        column_name_format: str = "{name}[{category}]",
        cat_missing_method: str = "fail",
        cat_missing_name: str = "(MISSING)",
    ):
        if cat_missing_method not in {"fail", "zero", "convert"}:
            raise ValueError(
                "cat_missing_method must be one of 'fail' 'zero' or 'convert'; "
                f" got {cat_missing_method}."
            )
        self._missing_method = cat_missing_method
        self._missing_category = cat_missing_name

        dev_codes = None
        if _dev.is_dev(cat_vec):
            # integer codes already in HBM (used by the device-side generators / row slicing)
            if categories is None:
                raise ValueError("categories are required when passing device codes")
            dev_codes = cat_vec.to(torch.int32).contiguous()
            self.categories = np.asarray(categories)
            has_neg = bool((dev_codes < 0).any().item()) if dev_codes.numel() else False
            indices = None
        else:
            if not hasattr(cat_vec, "dtype"):
                cat_vec = np.asarray(cat_vec)
            if categories is not None:
                self.categories = np.asarray(categories)
                indices = np.nan_to_num(np.asarray(cat_vec, dtype=np.float64), nan=-1)
                if len(indices) and max(indices) >= len(categories):
                    raise ValueError("Indices exceed length of categories.")
                if len(indices) and min(indices) < -1:
                    raise ValueError("Indices must be non-negative (or -1 for missing).")
            else:
                indices, self.categories = _extract_codes_and_categories(cat_vec)
            indices = np.asarray(indices)
            has_neg = bool(np.any(indices == -1))

        if has_neg:
            if self._missing_method == "fail":
                raise ValueError(
                    "Categorical data can't have missing values if cat_missing_method='fail'."
                )
            elif self._missing_method == "convert":
                if self._missing_category in self.categories:
                    raise ValueError(f"Missing category {self._missing_category} already exists.")
                self.categories = np.hstack(
                    [np.asarray(self.categories, dtype=object), self._missing_category])
                if dev_codes is not None:
                    dev_codes = torch.where(dev_codes < 0, len(self.categories) - 1, dev_codes)
                    dev_codes = dev_codes.to(torch.int32)
                else:
                    indices = np.where(indices < 0, len(self.categories) - 1, indices)
                self._has_missings = False
            else:
                self._has_missings = True
        else:
            self._has_missings = False

        self.drop_first = bool(drop_first)
        if dev_codes is None:
            try:
                host_codes = np.ascontiguousarray(indices).astype(np.int32, copy=False)
            except ValueError:
                raise ValueError(
                    "When creating a CategoricalMatrix with indices and categories, "
                    "indices must be castable to a numpy int32 dtype."
                )
            dev_codes = _dev.to_dev(host_codes)
        self._codes = dev_codes  # int32 CUDA tensor, read-only for all kernels
        self.shape = (int(dev_codes.numel()), max(len(self.categories) - int(drop_first), 0))
        self.dtype = np.dtype(dtype)
        self._colname = column_name
        self._colname_format = column_name_format
        self._term = self._colname if term_name is None else term_name

    # ---- host mirrors ----------------------------------------------------------------
    @property
    def indices(self) -> np.ndarray:
        """Host copy of the int32 codes (the reference's ``indices`` attribute)."""
        return _dev.to_host(self._codes)

    @property
    def cat(self):
        if pd is None:
            raise ModuleNotFoundError("the `cat` property requires pandas")
        return pd.Categorical.from_codes(self.indices, categories=self.categories)

    def recover_orig(self) -> np.ndarray:
        """1-D array with the data originally fed to __init__ (categorical_matrix.py:452-470)."""
        idx = self.indices
        orig = self.categories[idx]
        if self._has_missings:
            orig = orig.view(np.ma.MaskedArray)
            orig.mask = idx == -1
        elif self._missing_method == "convert" and self._missing_category in self.categories:
            orig = orig.view(np.ma.MaskedArray)
            orig.mask = idx == len(self.categories) - 1
        return orig

    def _sorted_perm(self):
        """(perm, segptr, n_valid): the rows with a valid column ordered by column (stable) and
        the shape[1]+1 segment offsets — built once and cached, the analogue of the reference's
        cached CSR (sparse_matrix.py:133-143); feeds the sorted-gather categorical x dense
        kernel."""
        cached = self.__dict__.get("_perm_cache")
        if cached is None:
            col = self._codes.to(torch.int64) - int(self.drop_first)
            order = torch.argsort(col, stable=True)
            n_neg = int((col < 0).sum().item())
            perm = order[n_neg:].to(torch.int32).contiguous()
            counts = torch.bincount(col[col >= 0], minlength=self.shape[1])
            segptr = torch.zeros(self.shape[1] + 1, dtype=torch.int64, device=col.device)
            segptr[1:] = torch.cumsum(counts, 0)
            cached = (perm, segptr.to(torch.int32).contiguous(), int(perm.numel()))
            self.__dict__["_perm_cache"] = cached
        return cached

    @property
    def _tdtype(self) -> torch.dtype:
        return _dev.torch_dtype(self.dtype)

    def _restrict_cols(self, cols):
        if cols is None or len(cols) == self.shape[1]:
            return None
        return _dev.idx32(cols)

    # ---- hot path --------------------------------------------------------------------
    def matvec(self, other, cols=None, out=None):
        """out[i] += other[codes[i]] (categorical_matrix.py:495-541); 1-D ``other`` only."""
        check_matvec_out_shape(self, out)
        if not _dev.is_dev(other):
            other = np.asarray(other)
        if other.ndim > 1:
            raise NotImplementedError(
                """CategoricalMatrix.matvec is only implemented for 1d arrays."""
            )
        check_matvec_dimensions(self, other, transpose=False)
        is_int = (not _dev.is_dev(other)) and np.issubdtype(other.dtype, np.signedinteger)
        if is_int:
            other = other.astype(float)
        other_t, host = _vec_in(other)
        if other_t.dtype not in (torch.float32, torch.float64):
            other_t = other_t.to(torch.float64)
        cols_t = self._restrict_cols(cols)
        if out is not None and _dev.is_dev(out) and out.dtype == other_t.dtype:
            ext_cat.matvec(self._codes, other_t, self.shape[0], cols_t, self.shape[1], out,
                           self.drop_first)
            return out
        res = torch.zeros(self.shape[0], dtype=other_t.dtype, device=other_t.device)
        ext_cat.matvec(self._codes, other_t, self.shape[0], cols_t, self.shape[1], res,
                       self.drop_first)
        if out is not None:
            return _accumulate_out(out, res, None)
        res_out = _dev.ret(res, host)
        if is_int:
            return res_out.astype(int)
        return res_out

    def transpose_matvec(self, vec, rows=None, cols=None, out=None):
        """out[c] += sum_{k in rows, codes[k]==c} vec[k] (categorical_matrix.py:543-616)."""
        if not _dev.is_dev(vec):
            vec = np.asarray(vec)
        check_matvec_dimensions(self, vec, transpose=True)
        if vec.ndim > 1:
            raise NotImplementedError(
                "CategoricalMatrix.transpose_matvec is only implemented for 1d arrays."
            )
        out_is_none = out is None
        if not out_is_none:
            check_transpose_matvec_out_shape(self, out)
        vec_t, host = _vec_in(vec)
        if vec_t.dtype not in (torch.float32, torch.float64):
            vec_t = vec_t.to(self._tdtype)
        rows_t = None if is_unrestricted(rows, self.shape[0]) else _dev.idx32(rows)
        cols_t = self._restrict_cols(cols)
        if _lib.deterministic() and self.shape[0] < 2**31 - 1:
            res = self._segment_sum(vec_t, rows_t)
            if cols_t is not None:   # columns outside the restriction contribute nothing
                keep = torch.zeros_like(res)
                keep[cols_t.to(torch.int64)] = 1
                res = res * keep
            if not out_is_none:
                if _dev.is_dev(out) and out.dtype == res.dtype:
                    out += res
                    return out
                return _accumulate_out(out, res, None)
            if res.dtype != self._tdtype:
                res = res.to(self._tdtype)
            if cols is not None:
                res = res.index_select(0, _dev.idx32(cols).to(torch.int64))
            return _dev.ret(res, host)
        if not out_is_none and _dev.is_dev(out) and out.dtype == vec_t.dtype:
            ext_cat.transpose_matvec(self._codes, vec_t, self.shape[1], rows_t, cols_t, out,
                                     self.drop_first)
            return out
        # the reference allocates `out` in the MATRIX dtype (categorical_matrix.py:587-588)
        res = torch.zeros(self.shape[1], dtype=vec_t.dtype, device=vec_t.device)
        ext_cat.transpose_matvec(self._codes, vec_t, self.shape[1], rows_t, cols_t, res,
                                 self.drop_first)
        if not out_is_none:
            return _accumulate_out(out, res, None)
        if res.dtype != self._tdtype:
            res = res.to(self._tdtype)
        if cols is not None:
            res = res.index_select(0, _dev.idx32(cols).to(torch.int64))
        return _dev.ret(res, host)

    def _segment_sum(self, w_t: torch.Tensor, rows_t) -> torch.Tensor:
        """Fixed-order weighted histogram of the codes (``tm_cat_segment_sum``): the
        deterministic form of transpose_matvec / sandwich."""
        perm, segptr, _ = self._sorted_perm()
        row_w = None
        if rows_t is not None:
            row_w = torch.zeros(self.shape[0], dtype=w_t.dtype, device=w_t.device)
            row_w[rows_t.to(torch.int64)] = 1
        out = torch.empty(self.shape[1], dtype=w_t.dtype, device=w_t.device)
        check(fn("tm_cat_segment_sum", _dev.suffix(w_t.dtype))(
            _dev.ptr(w_t.contiguous()), _dev.ptr(row_w), _dev.ptr(perm), _dev.ptr(segptr),
            self.shape[1], _dev.ptr(out), 0, _dev.stream_ptr()))
        return out

    def sandwich(self, d, rows=None, cols=None):
        """Diagonal sandwich, returned as ``scipy.sparse.dia_matrix`` for host input
        (categorical_matrix.py:618-653) or as the 1-D diagonal CUDA tensor for device input."""
        res_diag, host = self._sandwich_diag(d, rows, cols)
        if host:
            return sps.diags(_dev.to_host(res_diag))
        return res_diag

    def _sandwich_diag(self, d, rows=None, cols=None):
        if not _dev.is_dev(d):
            d = np.asarray(d)
        check_sandwich_compatible(self, d)
        d_t, host = _vec_in(d)
        rows_t = _dev.idx32(rows)
        if _lib.deterministic() and self.shape[0] < 2**31 - 1:
            res_diag = self._segment_sum(d_t, rows_t)
        else:
            res_diag = ext_cat.sandwich_categorical(self._codes, d_t, rows_t, self.shape[1],
                                                    self.drop_first)
        if cols is not None and len(cols) < self.shape[1]:
            res_diag = res_diag.index_select(0, _dev.idx32(cols).to(torch.int64))
        return res_diag, host

    def _cross_sandwich(self, other, d, rows=None, L_cols=None, R_cols=None):
        """self[rows, L_cols].T @ diag(d[rows]) @ other[rows, R_cols]."""
        if isinstance(other, DenseMatrix):
            return other._trim_cols(self._cross_dense(other._native(), d, rows, L_cols, R_cols),
                                    R_cols)
        if isinstance(other, SparseMatrix):
            return self._cross_sparse(other, d, rows, L_cols, R_cols)
        if isinstance(other, CategoricalMatrix):
            return self._cross_categorical(other, d, rows, L_cols, R_cols)
        raise TypeError

    def _cross_dense(self, other, d, rows, L_cols, R_cols):
        """(categorical_matrix.py:759-791)"""
        from .dense_matrix import _dense_to_dev

        other_t = other if _dev.is_dev(other) else _dense_to_dev(other)
        d_t, host = _vec_in(d)
        res = ext_split.sandwich_cat_dense(self._codes, self.shape[1], d_t, other_t,
                                           _dev.idx32(rows), _dev.idx32(R_cols), self.drop_first)
        return _dev.ret(_row_col_indexing(res, L_cols, None), host)

    def _cross_categorical(self, other, d, rows, L_cols, R_cols):
        """(categorical_matrix.py:793-823)"""
        if not isinstance(other, CategoricalMatrix):
            raise TypeError
        d_t, host = _vec_in(d)
        res = ext_split.sandwich_cat_cat(self._codes, other._codes, self.shape[1], other.shape[1],
                                         d_t, _dev.idx32(rows), self.drop_first, other.drop_first)
        return _dev.ret(_row_col_indexing(res, L_cols, R_cols), host)

    def _cross_sparse(self, other, d, rows, L_cols, R_cols):
        """(categorical_matrix.py:825-838; scipy csr_matmat in the reference)"""
        d_t, host = _vec_in(d)
        res = ext_split.sandwich_cat_sparse(self._codes, self.shape[1], d_t, other._csr,
                                            _dev.idx32(rows), _dev.idx32(R_cols), self.drop_first)
        return _dev.ret(_row_col_indexing(res, L_cols, None), host)

    def _get_col_stds(self, weights, col_means):
        """sqrt(max(0, mean - mean^2)) for 0/1 columns (categorical_matrix.py:728-737)."""
        host = not _dev.is_dev(weights)
        mean = self.transpose_matvec(weights)
        if host:
            return np.sqrt(np.maximum(mean - np.asarray(col_means) ** 2, 0))
        cm, _ = _vec_in(col_means, mean.dtype)
        return torch.sqrt(torch.clamp_min(mean - cm * cm, 0))

    # ---- conversions / indexing ------------------------------------------------------
    def getcol(self, i: int) -> SparseMatrix:
        """Column i as an (n x 1) SparseMatrix of ones, built on the device
        (categorical_matrix.py:674-688)."""
        i %= self.shape[1]  # wrap-around indexing
        i_corr = i + 1 if self.drop_first else i
        # codes == i_corr -> 0 (kept, column 0), everything else -> -1 (dropped)
        hit = torch.where(self._codes == i_corr, 0, -1).to(torch.int32)
        data, indices, indptr = self._device_csr_of(hit, 0, None, _dev.torch_dtype(self.dtype))
        return SparseMatrix.from_device_csr(data, indices, indptr, (self.shape[0], 1),
                                            column_names=[self.column_names[i]],
                                            term_names=[self.term_names[i]])

    @staticmethod
    def _device_csr_of(codes: torch.Tensor, drop_first: int, d: Optional[torch.Tensor],
                       tdt: torch.dtype):
        """(data, indices, indptr) of diag(d) @ onehot(codes) in HBM via ``tm_cat_to_csr``
        (multiply_complex / subset_categorical_complex, categorical.pyx:221-315)."""
        n = int(codes.numel())
        dev = codes.device
        indptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        indices = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        data = torch.empty(max(n, 1), dtype=tdt, device=dev)
        check(fn("tm_cat_to_csr", _dev.suffix(tdt))(
            _dev.ptr(codes), n, int(drop_first), None if d is None else _dev.ptr(d),
            _dev.ptr(data), _dev.ptr(indices), _dev.ptr(indptr), _dev.stream_ptr()))
        nnz = int(indptr[-1].item())
        return data[:nnz], indices[:nnz], indptr

    def _to_sparse_dev(self, d: Optional[torch.Tensor] = None, tdt=None) -> SparseMatrix:
        tdt = _dev.torch_dtype(self.dtype) if tdt is None else tdt
        data, indices, indptr = self._device_csr_of(self._codes, int(self.drop_first), d, tdt)
        return SparseMatrix.from_device_csr(data, indices, indptr, self.shape,
                                            column_names=self.column_names,
                                            term_names=self.term_names)

    def tocsr(self) -> sps.csr_matrix:
        """Host scipy CSR of ones (categorical_matrix.py:690-721); the structure is built on the
        device and copied out once."""
        _, indices, indptr = self._device_csr_of(self._codes, int(self.drop_first), None,
                                                 torch.float32)
        ind = _dev.to_host(indices)
        return sps.csr_matrix((np.ones(len(ind), dtype=int), ind, _dev.to_host(indptr)),
                              shape=self.shape)

    def to_sparse_matrix(self):
        return self._to_sparse_dev()

    def toarray(self) -> np.ndarray:
        return self.tocsr().toarray()

    def unpack(self):
        return self.cat

    def astype(self, dtype, order="K", casting="unsafe", copy=True):
        self.dtype = np.dtype(dtype)
        return self

    def _take_rows_dev(self, index) -> "CategoricalMatrix":
        """X[index, :] (torch index: slice, bool / int64 CUDA tensor); stays categorical."""
        sub = self._codes[index]
        new = CategoricalMatrix.__new__(CategoricalMatrix)
        new.__dict__.update(self.__dict__)
        new.__dict__.pop("_perm_cache", None)
        new.__dict__.pop("_run_sorted", None)
        new._codes = sub.contiguous()
        new.shape = (int(sub.numel()), self.shape[1])
        return new

    def __getitem__(self, item):
        if _dev.is_dev(item):
            return self._take_rows_dev(item if item.dtype == torch.bool else item.to(torch.int64))
        if isinstance(item, tuple) and len(item) == 2 and _dev.is_dev(item[0]) \
                and isinstance(item[1], slice) and item[1] == slice(None, None, None):
            return self._take_rows_dev(item[0] if item[0].dtype == torch.bool
                                       else item[0].to(torch.int64))
        row, col = _check_indexer(item)
        if _is_indexer_full_length(self.shape[1], col):
            if isinstance(row, np.ndarray):
                row = row.ravel()
            from .dense_matrix import _torch_index

            return self._take_rows_dev(_torch_index(row, self._codes.device))
        # column subset -> SparseMatrix, like the reference (issue #101 there)
        return self.to_sparse_matrix()[row, col]

    def multiply(self, other) -> SparseMatrix:
        """diag(other) @ X as a SparseMatrix (categorical_matrix.py:840-876)."""
        if not _dev.is_dev(other):
            other = np.asarray(other)
        if self.shape[0] != other.shape[0]:
            raise ValueError(
                f"Shapes do not match. Expected length of {self.shape[0]}. Got {len(other)}."
            )
        o_t, _ = _vec_in(other if _dev.is_dev(other) else np.squeeze(other))
        o_t = o_t.reshape(-1)
        if o_t.dtype not in (torch.float32, torch.float64):
            o_t = o_t.to(torch.float64)
        # device-side multiply_complex (categorical.pyx:221-272): no host round trip of the codes
        return self._to_sparse_dev(o_t.contiguous(), o_t.dtype)

    def __repr__(self):
        return f"{self.__class__.__name__}\nCategories: {self.categories}"

    # ---- names -----------------------------------------------------------------------
    def get_names(self, type: str = "column", missing_prefix: Optional[str] = None,
                  indices: Optional[list] = None) -> list:
        if type == "column":
            name = self._colname
        elif type == "term":
            name = self._term
        else:
            raise ValueError(f"Type must be 'column' or 'term', got {type}")
        if indices is None:
            indices = list(range(len(self.categories) - self.drop_first))
        if name is None and missing_prefix is None:
            return [None] * (len(self.categories) - self.drop_first)
        elif name is None:
            name = f"{missing_prefix}{indices[0]}-{indices[-1]}"
        if type == "column":
            return [self._colname_format.format(name=name, category=cat)
                    for cat in self.categories[self.drop_first:]]
        return [name] * (len(self.categories) - self.drop_first)

    def set_names(self, names: Union[str, list], type: str = "column"):
        if isinstance(names, str):
            names = [names]
        if len(names) != 1:
            if type == "column":
                base_names = []
                for name, cat in zip(names, self.categories[self.drop_first:]):
                    partial_name = self._colname_format.format(name="__CAPTURE__", category=cat)
                    pattern = re.escape(partial_name).replace("__CAPTURE__", "(.*)")
                    match = re.search(pattern, name) if name is not None else None
                    base_names.append(match.group(1) if match is not None else name)
                names = base_names
            if len(names) == self.shape[1] and all(name == names[0] for name in names):
                names = [names[0]]
        if len(names) != 1:
            raise ValueError("A categorical matrix has only one name")
        if type == "column":
            self._colname = names[0]
        elif type == "term":
            self._term = names[0]
        else:
            raise ValueError(f"Type must be 'column' or 'term', got {type}")
