"""SparseMatrix on the device (reference: sparse_matrix.py:27-407).

HBM layout: CSC (data, indices, indptr) AND CSR (data, indices, indptr, row) with int32
indices, both built once at construction — the analogue of the reference's CSC + lazily
cached CSR (sparse_matrix.py:133-143).  int64-indexed input is accepted (``idx_dtype`` is
kept for API parity) but stored as int32 on the device."""

from __future__ import annotations

from typing import Optional

import numpy as np
import torch
from scipy import sparse as sps

from . import _dev
from .dense_matrix import _accumulate_out, _np_dtype
from .ext.sparse import (
    DeviceCSC,
    DeviceCSR,
    csc_rmatvec,
    csr_dense_sandwich,
    csr_matvec,
    sparse_sandwich,
    transpose_square_dot_weights,
)
from .matrix_base import MatrixBase, _names_with_default, _vec_in
from .util import (
    _check_indexer,
    check_matvec_dimensions,
    check_matvec_out_shape,
    check_sandwich_compatible,
    check_transpose_matvec_out_shape,
    is_unrestricted,
    setup_restrictions,
)

_I32_MAX = np.iinfo(np.int32).max


def _csr_rows(indptr: torch.Tensor, nnz: int) -> torch.Tensor:
    n = indptr.numel() - 1
    counts = (indptr[1:] - indptr[:-1]).to(torch.int64)
    return torch.repeat_interleave(
        torch.arange(n, device=indptr.device, dtype=torch.int32), counts, output_size=nnz)


class SparseMatrix(MatrixBase):
    """Sparse block (CSC + CSR in HBM); same API as ``tabmat.SparseMatrix``."""

    def __init__(self, input_array, shape=None, dtype=None, copy=False, column_names=None,
                 term_names=None):
        if isinstance(input_array, np.ndarray):
            if input_array.ndim == 1:
                input_array = input_array.reshape(-1, 1)
            elif input_array.ndim > 2:
                raise ValueError("Input array must be 1- or 2-dimensional")
        host = sps.csc_matrix(input_array, shape, dtype, copy)
        self.idx_dtype = max(host.indices.dtype, host.indptr.dtype)
        if not host.has_sorted_indices:
            host.sort_indices()
        if host.dtype not in (np.float32, np.float64):
            host = host.astype(np.float64)
        self._init_from_host(host)
        self._init_names(column_names, term_names)

    # ---- construction helpers --------------------------------------------------------
    def _init_from_host(self, host: sps.csc_matrix):
        n, p = host.shape
        if n > _I32_MAX or p > _I32_MAX or host.nnz > _I32_MAX:
            raise ValueError("tabmat_b200 sparse blocks are int32-indexed per GPU shard")
        self._shape = (int(n), int(p))
        self._host_csc = host
        csr = host.tocsr()
        csr.sort_indices()
        self._csc = DeviceCSC(
            data=_dev.to_dev(host.data),
            indices=_dev.to_dev(host.indices.astype(np.int32, copy=False)),
            indptr=_dev.to_dev(host.indptr.astype(np.int32, copy=False)),
            shape=self._shape,
        )
        indptr = _dev.to_dev(csr.indptr.astype(np.int32, copy=False))
        self._csr = DeviceCSR(
            data=_dev.to_dev(csr.data),
            indices=_dev.to_dev(csr.indices.astype(np.int32, copy=False)),
            indptr=indptr,
            row=_csr_rows(indptr, int(csr.nnz)),
            shape=self._shape,
        )

    @classmethod
    def from_device_csr(cls, data: torch.Tensor, indices: torch.Tensor, indptr: torch.Tensor,
                        shape, column_names=None, term_names=None):
        """Build from CSR arrays already in HBM (column ids sorted inside each row)."""
        self = cls.__new__(cls)
        n, p = int(shape[0]), int(shape[1])
        self._shape = (n, p)
        self.idx_dtype = np.dtype(np.int32)
        self._host_csc = None
        indices = indices.to(torch.int32).contiguous()
        indptr = indptr.to(torch.int32).contiguous()
        nnz = int(data.numel())
        row = _csr_rows(indptr, nnz)
        self._csr = DeviceCSR(data.contiguous(), indices, indptr, row, self._shape)
        # CSC = stable sort of the CSR triplets by column (rows stay sorted inside a column)
        order = torch.argsort(indices.to(torch.int64), stable=True)
        counts = torch.bincount(indices.to(torch.int64), minlength=p)
        cptr = torch.zeros(p + 1, dtype=torch.int64, device=data.device)
        cptr[1:] = torch.cumsum(counts, 0)
        self._csc = DeviceCSC(data[order].contiguous(), row[order].contiguous(),
                              cptr.to(torch.int32), self._shape)
        self._init_names(column_names, term_names)
        return self

    def _row_blocked_csc(self, block_rows: int):
        """(data, row ids, offsets, n_blocks): the non-zeros ordered by (row block, column, row)
        with ``n_blocks * ncols + 1`` offsets — the row-blocked CSC copy that keeps the per-row
        gathers of the categorical x sparse kernel inside the L2 (built once and cached, like
        the reference's cached CSR, sparse_matrix.py:133-143)."""
        cache = self.__dict__.setdefault("_bcsc", {})
        cached = cache.get(block_rows)
        if cached is None:
            c = self._csr
            n, p = self._shape
            n_blocks = max(1, -(-n // block_rows))
            key = (c.row.to(torch.int64) // block_rows) * p + c.indices.to(torch.int64)
            order = torch.argsort(key, stable=True)  # CSR order is row-major: rows stay sorted
            counts = torch.bincount(key, minlength=n_blocks * p)
            del key
            ptr = torch.zeros(n_blocks * p + 1, dtype=torch.int64, device=c.data.device)
            ptr[1:] = torch.cumsum(counts, 0)
            cached = (c.data[order].contiguous(), c.row[order].contiguous(),
                      ptr.to(torch.int32).contiguous(), n_blocks, block_rows)
            cache[block_rows] = cached
        return cached[:4]

    def _init_names(self, column_names, term_names):
        if column_names is not None:
            if len(column_names) != self.shape[1]:
                raise ValueError(f"Expected {self.shape[1]} column names, got {len(column_names)}")
            self._colnames = column_names
        else:
            self._colnames = [None] * self.shape[1]
        if term_names is not None:
            if len(term_names) != self.shape[1]:
                raise ValueError(f"Expected {self.shape[1]} term names, got {len(term_names)}")
            self._terms = term_names
        else:
            self._terms = self._colnames

    # ---- array-like surface ----------------------------------------------------------
    @property
    def _array(self) -> sps.csc_matrix:
        """Host scipy mirror (built on demand when constructed from device arrays)."""
        if self._host_csc is None:
            c = self._csc
            self._host_csc = sps.csc_matrix(
                (_dev.to_host(c.data), _dev.to_host(c.indices), _dev.to_host(c.indptr)),
                shape=self._shape)
        return self._host_csc

    def __getitem__(self, key):
        if isinstance(key, tuple) and len(key) == 2 and _dev.is_dev(key[0]) \
                and isinstance(key[1], slice) and key[1] == slice(None, None, None):
            return self._take_rows_dev(key[0])
        row, col = _check_indexer(key)
        if isinstance(col, slice) and col == slice(None, None, None):
            # row subset (CV folds, row re-ordering): stays in HBM, no host round trip
            return self._take_rows_dev(row)
        colnames = np.array(self.column_names, dtype=object)[col].ravel().tolist()
        terms = np.array(self.term_names, dtype=object)[col].ravel().tolist()
        return type(self)(self._array.__getitem__((row, col)), column_names=colnames,
                          term_names=terms)

    def _take_rows_dev(self, row) -> "SparseMatrix":
        """X[row, :] built on the device from the CSR arrays (``row``: slice, int / bool array
        or CUDA index tensor; duplicates and any order allowed, like scipy's fancy indexing)."""
        c = self._csr
        n = self._shape[0]
        dev = c.data.device
        if isinstance(row, slice):
            rows_t = torch.arange(n, device=dev)[row]
        elif _dev.is_dev(row):
            rows_t = row.nonzero().reshape(-1) if row.dtype == torch.bool else row.to(torch.int64)
        else:
            a = np.asarray(row).reshape(-1)
            a = np.flatnonzero(a) if a.dtype == bool else a.astype(np.int64)
            rows_t = torch.from_numpy(a).to(dev)
        rows_t = torch.where(rows_t < 0, rows_t + n, rows_t)
        m = int(rows_t.numel())
        indptr = c.indptr.to(torch.int64)
        counts = (indptr[1:] - indptr[:-1])[rows_t]
        new_indptr = torch.zeros(m + 1, dtype=torch.int64, device=dev)
        new_indptr[1:] = torch.cumsum(counts, 0)
        nnz = int(new_indptr[-1].item()) if m else 0
        new_row = torch.repeat_interleave(torch.arange(m, device=dev), counts, output_size=nnz)
        src = indptr[rows_t][new_row] + (torch.arange(nnz, device=dev) - new_indptr[:-1][new_row])
        del new_row, counts
        return SparseMatrix.from_device_csr(
            c.data[src], c.indices[src], new_indptr.to(torch.int32), (m, self._shape[1]),
            column_names=self.column_names, term_names=self.term_names)

    @property
    def shape(self):  # type: ignore
        return self._shape

    @property
    def ndim(self):  # type: ignore
        return 2

    @property
    def dtype(self):  # type: ignore
        return _dev.np_dtype(self._csc.data.dtype)

    @property
    def indices(self):
        return self._array.indices

    @property
    def indptr(self):
        return self._array.indptr

    @property
    def data(self):
        return self._array.data

    @property
    def array_csc(self):
        return self._array

    @property
    def array_csr(self):
        return self._array.tocsr()

    def tocsc(self, copy=False):
        return self._array.tocsc(copy=copy)

    def transpose(self):
        return type(self)(self._array.T)

    T = property(transpose)

    def getcol(self, i):
        return type(self)(self._array[:, [i]], column_names=[self.column_names[i]],
                          term_names=[self.term_names[i]])

    def unpack(self):
        return self._array

    def toarray(self):
        return self._array.toarray()

    def dot(self, other):
        return self._array.dot(other)

    def astype(self, dtype, order="K", casting="unsafe", copy=True):
        return type(self)(self._array.astype(dtype, casting, copy))

    def multiply(self, other):
        other = np.asarray(other) if not _dev.is_dev(other) else _dev.to_host(other)
        if other.ndim == 1:
            other = other[:, np.newaxis]
        return type(self)(sps.csc_matrix(self._array.multiply(other)),
                          column_names=self.column_names, term_names=self.term_names)

    # ---- hot path --------------------------------------------------------------------
    def sandwich(self, d, rows=None, cols=None):
        """A[rows, cols].T @ diag(d[rows]) @ A[rows, cols] (sparse_matrix.py:175-185)."""
        if not _dev.is_dev(d):
            d = np.asarray(d)
        check_sandwich_compatible(self, d)
        d_t, host = _vec_in(d)
        rows_t, cols_t = setup_restrictions(self.shape, rows, cols)
        return _dev.ret(sparse_sandwich(self._csr, d_t, rows_t, cols_t), host)

    def _cross_sandwich(self, other, d, rows, L_cols=None, R_cols=None):
        from .categorical_matrix import CategoricalMatrix
        from .dense_matrix import DenseMatrix

        if isinstance(other, DenseMatrix):
            return other._trim_cols(self.sandwich_dense(other._native(), d, rows, L_cols, R_cols),
                                    R_cols)
        if isinstance(other, CategoricalMatrix):
            res = other._cross_sandwich(self, d, rows, R_cols, L_cols)
            return res.T if isinstance(res, np.ndarray) else res.t()
        raise TypeError

    def sandwich_dense(self, B, d, rows, L_cols, R_cols):
        """self[rows, L_cols].T @ diag(d[rows]) @ B[rows, R_cols] (sparse_matrix.py:206-229)."""
        if not hasattr(d, "dtype"):
            d = np.asarray(d)
        b_dtype = _np_dtype(B)
        if self.dtype != _np_dtype(d) or b_dtype != _np_dtype(d):
            raise TypeError(
                f"""self, B and d all need to be of same dtype, either
                np.float64 or np.float32. This matrix is of type {self.dtype},
                B is of type {b_dtype}, while d is of type {_np_dtype(d)}."""
            )
        from .dense_matrix import _dense_to_dev

        B_t = B if _dev.is_dev(B) else _dense_to_dev(B)
        d_t, host = _vec_in(d)
        rows_t, L_t = setup_restrictions(self.shape, rows, L_cols)
        R_t = _dev.idx32(R_cols)
        res = self._sandwich_dense_gather(B_t, d_t, rows_t, L_t, R_t)
        if res is None:
            res = csr_dense_sandwich(self._csr, B_t, d_t, rows_t, L_t, R_t)
        return _dev.ret(res, host)

    def _sandwich_dense_gather(self, B_t, d_t, rows_t, L_t, R_t):
        """Gather form of the sparse x dense cross sandwich (``tm_csc_dense_gather_sandwich``):
        large matrices, row-major B, no column restriction.  One vector RED per (row block,
        column) run of a row-blocked CSC copy (built once and cached, like the reference's cached
        CSR) instead of one per non-zero.  None = not applicable."""
        import os

        from ._lib import check, fn

        if os.environ.get("TABMAT_B200_DXS") == "red" or L_t is not None or R_t is not None:
            return None
        n, p = self._shape
        q = int(B_t.shape[1]) if B_t.dim() == 2 else 0
        fsize = B_t.element_size()
        width = 16 // fsize
        if (not B_t.is_contiguous() or q <= 0 or q % width or q > 64 * width
                or B_t.data_ptr() % 16 or not self._csr.nnz):
            return None
        cap = int(os.environ.get("TABMAT_B200_GATHER_MB", "64")) * (1 << 20)
        rows_blk = 1 << (max(1024, cap // (q * fsize)).bit_length() - 1)
        if n <= rows_blk + rows_blk // 2:
            return None
        bd, br, bp, nblk = self._row_blocked_csc(rows_blk)
        dd = d_t
        if rows_t is not None:   # the restriction becomes a zero weight outside `rows`
            dd = torch.zeros_like(d_t)
            idx = rows_t.to(torch.int64)
            dd[idx] = d_t[idx]
        out = torch.empty((p, q), dtype=B_t.dtype, device=B_t.device)
        check(fn("tm_csc_dense_gather_sandwich", _dev.suffix(B_t.dtype))(
            _dev.ptr(bd), _dev.ptr(br), _dev.ptr(bp), p, nblk, _dev.ptr(B_t), q,
            _dev.ptr(dd.contiguous()), _dev.ptr(out), _dev.stream_ptr()))
        return out

    def _matvec_helper(self, vec, rows, cols, out, transpose: bool):
        if not _dev.is_dev(vec):
            vec = np.asarray(vec)
        check_matvec_dimensions(self, vec, transpose)
        tdt = self._csc.data.dtype
        vec_t, host = _vec_in(vec, tdt)
        if is_unrestricted(rows, self.shape[0]):
            rows = None
        if is_unrestricted(cols, self.shape[1]):
            cols = None
        rows_t, cols_t = setup_restrictions(self.shape, rows, cols)

        def fast(v, acc=None):
            if transpose:
                return csc_rmatvec(self._csc, v, rows_t, cols_t, acc)
            return csr_matvec(self._csr, v, rows_t, cols_t, acc)

        # unrestricted 1-D with a device `out`: accumulate in the kernel, like the reference's
        # *_unrestricted functions (sparse.pyx:79-103, 142-166)
        if (vec_t.dim() == 1 and out is not None and _dev.is_dev(out) and rows_t is None
                and cols_t is None and out.dtype == tdt and out.is_contiguous()):
            fast(vec_t, out)
            return out
        if vec_t.dim() == 1:
            res = fast(vec_t)
        else:
            flat = vec_t.reshape(vec_t.shape[0], -1)
            res = torch.stack([fast(flat[:, j].contiguous()) for j in range(flat.shape[1])],
                              dim=1).reshape((-1,) + tuple(vec_t.shape[1:]))
        if out is None:
            return _dev.ret(res, host)
        return _accumulate_out(out, res, cols_t if transpose else rows_t)

    def matvec(self, vec, cols=None, out=None):
        """self[:, cols] @ vec[cols] (sparse_matrix.py:277-282)."""
        check_matvec_out_shape(self, out)
        return self._matvec_helper(vec, None, cols, out, False)

    def transpose_matvec(self, vec, rows=None, cols=None, out=None):
        """self[rows, cols].T @ vec[rows] (sparse_matrix.py:284-293)."""
        check_transpose_matvec_out_shape(self, out)
        return self._matvec_helper(vec, rows, cols, out, True)

    def _get_col_stds(self, weights, col_means):
        """sqrt(max(0, sum_i w_i x_ij^2 - mean_j^2)) (sparse_matrix.py:295-311)."""
        tdt = self._csc.data.dtype
        w_t, host = _vec_in(weights, tdt)
        m_t, _ = _vec_in(col_means, tdt)
        sqrt_arg = transpose_square_dot_weights(self._csc, w_t) - m_t * m_t
        return _dev.ret(torch.sqrt(torch.clamp_min(sqrt_arg, 0)), host)

    # ---- names -----------------------------------------------------------------------
    def get_names(self, type: str = "column", missing_prefix: Optional[str] = None,
                  indices: Optional[list] = None) -> list:
        if type == "column":
            names = self._colnames
        elif type == "term":
            names = self._terms
        else:
            raise ValueError(f"Type must be 'column' or 'term', got {type}")
        return _names_with_default(names, missing_prefix, indices)

    def set_names(self, names, type: str = "column"):
        if isinstance(names, str):
            names = [names]
        if len(names) != self.shape[1]:
            raise ValueError(f"Length of names must be {self.shape[1]}")
        if type == "column":
            self._colnames = names
        elif type == "term":
            self._terms = names
        else:
            raise ValueError(f"Type must be 'column' or 'term', got {type}")
