"""StandardizedMatrix (reference: standardized_mat.py:18-378).

``self[i, j] = mult[j] * mat[i, j] + shift[j]`` without ever materialising it: every
operation is the inner matrix' device operation plus rank-1 corrections.  The O(p^2)
corrections of ``sandwich`` run on the device too (the reference does them with numpy
outer products, standardized_mat.py:154-171)."""

from __future__ import annotations

from typing import Optional

import numpy as np
import torch
from scipy import sparse as sps

from . import _dev
from ._lib import check, fn
from .dense_matrix import DenseMatrix, _accumulate_out
from .matrix_base import MatrixBase, _vec_in
from .sparse_matrix import SparseMatrix
from .util import (
    check_matvec_dimensions,
    check_matvec_out_shape,
    check_sandwich_compatible,
    check_transpose_matvec_out_shape,
)


class StandardizedMatrix:
    """Shifted / scaled view of a MatrixBase; same API as ``tabmat.StandardizedMatrix``."""

    __array_priority__ = 11
    __array_ufunc__ = None

    def __init__(self, mat: MatrixBase, shift, mult=None):
        shift_arr = np.atleast_1d(np.squeeze(_host(shift)))
        expected_shape = (mat.shape[1],)
        if not isinstance(mat, MatrixBase):
            raise TypeError("mat should be an instance of a MatrixBase subclass.")
        if not shift_arr.shape == expected_shape:
            raise ValueError(
                f"""Expected shift to be able to conform to shape {expected_shape},
            but it has shape {np.asarray(shift).shape}"""
            )
        if mult is not None:
            mult_arr = np.atleast_1d(np.squeeze(_host(mult)))
            if not mult_arr.shape == expected_shape:
                raise ValueError(
                    f"""Expected mult to be able to conform to shape {expected_shape},
                but it has shape {np.asarray(mult).shape}"""
                )
        else:
            mult_arr = None
        self.shift = shift_arr
        self.mult = mult_arr
        self.mat = mat
        self.shape = mat.shape
        self.ndim = mat.ndim
        self.dtype = mat.dtype
        self._dev_cache: dict = {}

    # device copies of shift / mult in a given dtype
    def _shift_t(self, dtype: torch.dtype) -> torch.Tensor:
        key = ("shift", dtype)
        if key not in self._dev_cache:
            self._dev_cache[key] = _dev.to_dev(self.shift, dtype)
        return self._dev_cache[key]

    def _mult_t(self, dtype: torch.dtype) -> Optional[torch.Tensor]:
        if self.mult is None:
            return None
        key = ("mult", dtype)
        if key not in self._dev_cache:
            self._dev_cache[key] = _dev.to_dev(self.mult, dtype)
        return self._dev_cache[key]

    # ---- hot path --------------------------------------------------------------------
    def matvec(self, other_mat, cols=None, out=None):
        """mat.matvec(mult * v, cols) + shift[cols] . v[cols] (standardized_mat.py:69-97)."""
        if not _dev.is_dev(other_mat):
            other_mat = np.asarray(other_mat)
        check_matvec_dimensions(self, other_mat, transpose=False)
        check_matvec_out_shape(self, out)
        v_t, host = _vec_in(other_mat)
        if v_t.dtype not in (torch.float32, torch.float64):
            v_t = v_t.to(torch.float64)
        cols_l = None if cols is None else _dev.idx32(cols).to(torch.int64)
        mult = self._mult_t(v_t.dtype)
        mult_other = v_t
        if mult is not None:
            mult_other = mult.reshape((-1,) + (1,) * (v_t.dim() - 1)) * v_t
        out_dev = out if (out is not None and _dev.is_dev(out)) else None
        mat_part = self.mat.matvec(mult_other, cols, out=out_dev)
        shift = self._shift_t(mat_part.dtype)
        vv = v_t.to(mat_part.dtype)
        if cols_l is None:
            shift_part = torch.tensordot(shift, vv, dims=([0], [0]))
        else:
            shift_part = torch.tensordot(shift[cols_l], vv[cols_l], dims=([0], [0]))
        mat_part += shift_part
        if out is not None and out_dev is None:
            return _accumulate_out(out, mat_part, None)
        if out_dev is not None:
            return out_dev
        return _dev.ret(mat_part, host)

    def getcol(self, i: int):
        mult = None
        if self.mult is not None:
            mult = [self.mult[i]]
        col = self.mat.getcol(i)
        if isinstance(col, sps.csc_matrix) and not isinstance(col, MatrixBase):
            col = SparseMatrix(col)
        return StandardizedMatrix(col, [self.shift[i]], mult)

    def sandwich(self, d, rows=None, cols=None):
        """inner sandwich * outer(mult, mult) + three rank-1 terms
        (standardized_mat.py:123-172)."""
        if not hasattr(d, "dtype"):
            d = np.asarray(d)
        check_sandwich_compatible(self, d)
        d_t, host = _vec_in(d)
        tdt = d_t.dtype
        rows_t = _dev.idx32(rows)
        cols_t = _dev.idx32(cols)
        cols_l = None if cols_t is None else cols_t.to(torch.int64)

        from .categorical_matrix import CategoricalMatrix

        if isinstance(self.mat, CategoricalMatrix):
            term1, _ = self.mat._sandwich_diag(d_t, rows_t, cols_t)
            term1_is_diag = True
            d_mat = self.mat.transpose_matvec(d_t, rows_t, cols_t).to(tdt)
        else:
            # the inner sandwich and inner.T @ d come from ONE pass where the inner matrix can
            # do that (SplitMatrix: tm_split_sandwich_rmatvec_blocks); two calls otherwise, like
            # the reference (standardized_mat.py:141-147)
            term1, d_mat = self.mat.sandwich_and_transpose_matvec(d_t, d_t, rows_t, cols_t)
            d_mat = d_mat.to(tdt)
            term1_is_diag = False
        shift = self._shift_t(tdt)
        mult = self._mult_t(tdt)
        limited_shift = (shift if cols_l is None else shift[cols_l]).contiguous()
        limited_mult = None
        if mult is not None:
            limited_mult = (mult if cols_l is None else mult[cols_l]).contiguous()
        sum_d = (d_t.sum() if rows_t is None else d_t[rows_t.to(torch.int64)].sum()).reshape(1)
        if term1.dtype not in (tdt, torch.float64):
            term1 = term1.to(tdt)
        term1 = term1.contiguous()
        m = int(limited_shift.numel())
        # the reference adds in place into the rank-1 sum, so the result keeps ITS dtype (f32
        # even over a SplitMatrix); one kernel instead of five eager ones
        res = torch.empty((m, m), dtype=tdt, device=d_t.device)
        check(fn("tm_std_sandwich_combine", _dev.suffix(tdt))(
            _dev.ptr(term1), int(term1.dtype == torch.float64), int(term1_is_diag),
            _dev.ptr(d_mat.contiguous()), _dev.ptr(limited_shift), _dev.ptr(limited_mult),
            _dev.ptr(sum_d), m, _dev.ptr(res), _dev.stream_ptr()))
        return _dev.ret(res, host)

    def unstandardize(self) -> MatrixBase:
        return self.mat

    def transpose_matvec(self, other, rows=None, cols=None, out=None):
        """mult[cols] * mat.transpose_matvec(v, rows, cols) + outer(shift[cols], sum v[rows])
        (standardized_mat.py:178-230)."""
        check_transpose_matvec_out_shape(self, out)
        if not _dev.is_dev(other):
            other = np.asarray(other)
        check_matvec_dimensions(self, other, transpose=True)
        v_t, host = _vec_in(other)
        if v_t.dtype not in (torch.float32, torch.float64):
            v_t = v_t.to(torch.float64)
        rows_t = _dev.idx32(rows)
        cols_t = _dev.idx32(cols)
        cols_l = None if cols_t is None else cols_t.to(torch.int64)
        res = self.mat.transpose_matvec(v_t, rows_t, cols_t)
        rdt = res.dtype
        vv = v_t.to(rdt)
        other_sum = vv.sum(0) if rows_t is None else vv[rows_t.to(torch.int64)].sum(0)
        shift = self._shift_t(rdt)
        lshift = shift if cols_l is None else shift[cols_l]
        shift_part = lshift.reshape((-1,) + (1,) * (res.dim() - 1)) * other_sum
        mult = self._mult_t(rdt)
        if mult is not None:
            lmult = mult if cols_l is None else mult[cols_l]
            res = res * lmult.reshape((-1,) + (1,) * (res.dim() - 1))
        res = res + shift_part
        if out is None:
            return _dev.ret(res, host)
        return _accumulate_out(out, res, cols_t)

    def __rmatmul__(self, other):
        if not hasattr(other, "T"):
            other = np.asarray(other)
        return self.transpose_matvec(other.T).T

    def __matmul__(self, other):
        return self.matvec(other)

    def multiply(self, other) -> DenseMatrix:
        return DenseMatrix(self.toarray()).multiply(other)

    def toarray(self) -> np.ndarray:
        mat_part = self.mat.toarray()
        if self.mult is not None:
            mat_part = self.mult[None, :] * mat_part
        return mat_part + self.shift[None, :]

    @property
    def A(self) -> np.ndarray:
        return self.toarray()

    def astype(self, dtype, order="K", casting="unsafe", copy=True):
        return type(self)(
            self.mat.astype(dtype, casting=casting, copy=copy),
            self.shift.astype(dtype, order=order, casting=casting, copy=copy),
        )

    def __getitem__(self, item):
        if isinstance(item, tuple):
            row, col = item
        else:
            row = item
            col = slice(None, None, None)
        mat_part = self.mat.__getitem__(item)
        shift_part = self.shift[col]
        mult_part = self.mult
        if mult_part is not None:
            mult_part = np.atleast_1d(mult_part[col])
        if isinstance(row, int):
            out = mat_part.toarray()
            if mult_part is not None:
                out = out * mult_part
            return out + shift_part
        return StandardizedMatrix(mat_part, np.atleast_1d(shift_part), mult_part)

    def __repr__(self):
        return f"""StandardizedMat. Mat: {type(self.mat)} of shape {self.mat.shape}.
        Shift: {self.shift}
        Mult: {self.mult}
        """

    def get_names(self, type: str = "column", missing_prefix: Optional[str] = None,
                  indices: Optional[list] = None) -> list:
        return self.mat.get_names(type, missing_prefix, indices)

    def set_names(self, names, type: str = "column"):
        self.mat.set_names(names, type)

    @property
    def column_names(self):
        return self.get_names(type="column")

    @column_names.setter
    def column_names(self, names):
        self.set_names(names, type="column")

    @property
    def term_names(self):
        return self.get_names(type="term")

    @term_names.setter
    def term_names(self, names):
        self.set_names(names, type="term")


def _host(x):
    return _dev.to_host(x) if _dev.is_dev(x) else np.asarray(x)
