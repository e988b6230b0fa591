"""DenseMatrix on the device (reference: dense_matrix.py:24-347).

The array lives in HBM as a C- or F-contiguous 2-D CUDA tensor.  ``sandwich`` runs the
tcgen05 weighted-SYRK kernel (fp32, C order, p <= 256) or the CUDA-core kernel; ``matvec`` /
``transpose_matvec`` run GEMV kernels with optional row/column restrictions."""

from __future__ import annotations

import os
import warnings
from typing import Optional

import numpy as np
import torch

from . import _dev
from .ext.dense import (
    dense_matvec,
    dense_rmatvec,
    dense_sandwich,
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


def _dense_to_dev(input_array) -> torch.Tensor:
    if _dev.is_dev(input_array):
        t = input_array
        if t.dim() == 1:
            t = t.reshape(-1, 1)
        elif t.dim() > 2:
            raise ValueError("Input array must be 1- or 2-dimensional")
        if not (t.is_contiguous() or t.t().is_contiguous()):
            warnings.warn("Input array is not contiguous; making a copy.", UserWarning,
                          stacklevel=3)
            t = t.t().contiguous().t()  # F order, like the reference (dense_matrix.py:49-58)
        return t
    a = np.asarray(input_array)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    elif a.ndim > 2:
        raise ValueError("Input array must be 1- or 2-dimensional")
    if not a.flags["C_CONTIGUOUS"] and not a.flags["F_CONTIGUOUS"]:
        warnings.warn("Input array is not contiguous; making a copy.", UserWarning, stacklevel=3)
        a = np.asfortranarray(a)
    dev = _dev.require_cuda()
    if a.flags["C_CONTIGUOUS"]:
        if not a.flags.writeable:
            a = a.copy()
        return torch.from_numpy(a).to(dev)
    # F order on the host (what pandas' to_numpy and the reference's constructors produce,
    # constructor_util.py:36): upload the transposed (C-contiguous) buffer and store the block
    # ROW-MAJOR in HBM.  The storage order is this library's choice, like the cached CSR: the
    # TMA / tcgen05 SYRK, the one-hot MMAs and the fused scatter warps all stream row tiles, and
    # a host-built matrix would otherwise never reach them (one transposing copy at upload).
    at = a.T
    if not at.flags.writeable:
        at = at.copy()
    return torch.from_numpy(at).to(dev).t().contiguous()


class DenseMatrix(MatrixBase):
    """Dense block backed by a CUDA tensor; same API as ``tabmat.DenseMatrix``."""

    def __init__(self, input_array, column_names=None, term_names=None):
        arr = _dense_to_dev(input_array)
        # float32, row-major, width not a multiple of 4: the TMA / tcgen05 SYRK, the one-hot MMAs
        # and the 16-byte vector kernels all need a 16-byte row pitch.  The block is STORED with
        # its width padded to the next multiple of 4 by zero columns (<= 3 extra columns of HBM);
        # `_array` is the logical (n x p) view of it, `_store` what the kernels see.  Zero columns
        # change nothing in X^T D X, X v or X^T v apart from zero rows / entries that are cut off
        # again.  TABMAT_B200_DENSE_PAD=0 keeps the width (CUDA-core kernels then).
        self._store = None
        if (os.environ.get("TABMAT_B200_DENSE_PAD", "1") != "0" and arr.dtype == torch.float32
                and arr.dim() == 2 and arr.is_contiguous() and arr.shape[0] > 0
                and arr.shape[1] > 0 and arr.shape[1] % 4):
            n, p = arr.shape
            store = torch.zeros((n, (p + 3) // 4 * 4), dtype=arr.dtype, device=arr.device)
            store[:, :p] = arr
            self._store = store
            arr = store[:, :p]
        self._array = arr
        width = self._array.shape[1]
        if column_names is not None:
            if len(column_names) != width:
                raise ValueError(f"Expected {width} column names, got {len(column_names)}")
            self._colnames = column_names
        else:
            self._colnames = [None] * width
        if term_names is not None:
            if len(term_names) != width:
                raise ValueError(f"Expected {width} term names, got {len(term_names)}")
            self._terms = term_names
        else:
            self._terms = self._colnames

    # ---- padded storage ----------------------------------------------------------------
    def _native(self) -> torch.Tensor:
        """The tensor handed to the native kernels: the zero-padded storage when there is one."""
        return self._array if self._store is None else self._store

    def _pad_vec(self, v_t: torch.Tensor) -> torch.Tensor:
        """A length-p vector (or p x k matrix) over the columns, extended by zeros to the stored
        width."""
        if self._store is None or v_t.shape[0] == self._store.shape[1]:
            return v_t
        extra = self._store.shape[1] - v_t.shape[0]
        return torch.cat([v_t, v_t.new_zeros((extra,) + tuple(v_t.shape[1:]))], dim=0)

    def _trim_cols(self, res, cols):
        """Cut the padding columns off a result whose LAST axis runs over all stored columns
        (`cols` is None); results over a `cols` selection never contain them."""
        if self._store is None or cols is not None:
            return res
        p = self._array.shape[1]
        out = res[..., :p]
        return out.contiguous() if isinstance(out, torch.Tensor) else np.ascontiguousarray(out)

    # ---- array-like surface ----------------------------------------------------------
    def _take_rows_dev(self, rows_t: torch.Tensor) -> "DenseMatrix":
        """X[rows_t, :] for an int64 CUDA index tensor; keeps the C / F storage order."""
        A = self._array
        if self._store is not None:
            # gather the padded rows (one contiguous copy) and adopt them as the new storage
            new = type(self).__new__(type(self))
            new._store = self._store.index_select(0, rows_t)
            new._array = new._store[:, :A.shape[1]]
            new._colnames, new._terms = self._colnames, self._terms
            return new
        if A.dim() == 2 and not A.is_contiguous() and A.t().is_contiguous():
            sub = A.t().index_select(1, rows_t).t()
        else:
            sub = A.index_select(0, rows_t)
        return type(self)(sub, column_names=self.column_names, term_names=self.term_names)

    def __getitem__(self, key):
        if isinstance(key, tuple) and len(key) == 2 and _dev.is_dev(key[0]) \
                and isinstance(key[1], slice) and key[1] == slice(None, None, None):
            return self._take_rows_dev(key[0].to(torch.int64))
        row, col = _check_indexer(key)
        colnames = np.array(self.column_names, dtype=object)[col].ravel().tolist()
        terms = np.array(self.term_names, dtype=object)[col].ravel().tolist()
        sub = self._array[_torch_index(row, self._array.device),
                          _torch_index(col, self._array.device)]
        if sub.dim() == 2 and not (sub.is_contiguous() or sub.t().is_contiguous()):
            sub = sub.contiguous()
        return type(self)(sub, column_names=colnames, term_names=terms)

    def __str__(self):
        return "{}x{} DenseMatrix:\n\n".format(*self.shape) + np.array_str(self.toarray())

    def __repr__(self):
        return f"{type(self).__name__}({np.array2string(self.toarray(), separator=', ')})"

    @property
    def shape(self):  # type: ignore
        return tuple(self._array.shape)

    @property
    def ndim(self):  # type: ignore
        return 2

    @property
    def dtype(self):  # type: ignore
        return _dev.np_dtype(self._array.dtype)

    def transpose(self):
        a = self._array if self._store is None else self._array.contiguous()
        return type(self)(a.t())

    T = property(transpose)

    def astype(self, dtype, order="K", casting="unsafe", copy=True):
        a = self._array if self._store is None else self._array.contiguous()
        return type(self)(a.to(_dev.torch_dtype(dtype)), column_names=self.column_names,
                          term_names=self.term_names)

    def getcol(self, i):
        return type(self)(self._array[:, [i]].contiguous(), column_names=[self.column_names[i]],
                          term_names=[self.term_names[i]])

    def toarray(self):
        a = self._array
        if self._store is not None:
            return _dev.to_host(a.contiguous())
        if a.is_contiguous():
            return _dev.to_host(a)
        return _dev.to_host(a.t()).T  # keeps F order on the host

    def unpack(self):
        return self.toarray()

    # ---- hot path --------------------------------------------------------------------
    def sandwich(self, d, rows=None, cols=None):
        """X[rows, cols].T @ diag(d[rows]) @ X[rows, cols] (dense_matrix.py:153-163)."""
        if not _dev.is_dev(d):
            d = np.asarray(d)
        check_sandwich_compatible(self, d)
        d_t, host = _vec_in(d)
        rows_t, cols_t = setup_restrictions(self.shape, rows, cols)
        res = dense_sandwich(self._native(), d_t, rows_t, cols_t)
        if self._store is not None and cols_t is None:
            p = self._array.shape[1]
            res = res[:p, :p].contiguous()
        return _dev.ret(res, host)

    def _cross_sandwich(self, other, d, rows=None, L_cols=None, R_cols=None):
        from .categorical_matrix import CategoricalMatrix
        from .sparse_matrix import SparseMatrix

        if isinstance(other, (SparseMatrix, CategoricalMatrix)):
            res = other._cross_sandwich(self, d, rows, R_cols, L_cols)
            return res.T if isinstance(res, np.ndarray) else res.t()
        raise TypeError

    def _get_col_stds(self, weights, col_means):
        w_t, host = _vec_in(weights, self._array.dtype)
        m_t, _ = _vec_in(col_means, self._array.dtype)
        sqrt_arg = transpose_square_dot_weights(self._native(), w_t, self._pad_vec(m_t))
        sqrt_arg = self._trim_cols(sqrt_arg, None)
        return _dev.ret(torch.sqrt(torch.clamp_min(sqrt_arg, 0)), host)

    def _matvec_helper(self, vec, rows, cols, out, transpose: bool):
        if not _dev.is_dev(vec):
            vec = np.asarray(vec)
        check_matvec_dimensions(self, vec, transpose=transpose)
        res_dtype = np.result_type(self.dtype, _np_dtype(vec))
        vec_t, host = _vec_in(vec, self._array.dtype)
        if is_unrestricted(rows, self.shape[0]):
            rows = None
        if is_unrestricted(cols, self.shape[1]):
            cols = None
        rows_t, cols_t = setup_restrictions(self.shape, rows, cols)
        fast = dense_rmatvec if transpose else dense_matvec
        X = self._native()
        if not transpose:
            vec_t = self._pad_vec(vec_t)   # X v: the padding columns meet zeros
        if vec_t.dim() == 1:
            res = fast(X, vec_t, rows_t, cols_t)
        else:
            flat = vec_t.reshape(vec_t.shape[0], -1)
            cols_out = [fast(X, flat[:, j].contiguous(), rows_t, cols_t)
                        for j in range(flat.shape[1])]
            res = torch.stack(cols_out, dim=1).reshape((-1,) + tuple(vec_t.shape[1:]))
        if transpose and self._store is not None and cols_t is None:
            res = res[:self._array.shape[1]].contiguous()   # X^T v: cut the padding entries off
        if res_dtype != self.dtype and np.issubdtype(res_dtype, np.floating):
            res = res.to(_dev.torch_dtype(res_dtype))
        if out is None:
            return _dev.ret(res, host)
        sel = cols_t if transpose else rows_t
        return _accumulate_out(out, res, sel)

    def transpose_matvec(self, vec, rows=None, cols=None, out=None):
        """self[rows, cols].T @ vec[rows] (dense_matrix.py:238-247)."""
        check_transpose_matvec_out_shape(self, out)
        return self._matvec_helper(vec, rows, cols, out, True)

    def matvec(self, vec, cols=None, out=None):
        """self[:, cols] @ vec[cols] (dense_matrix.py:249-257)."""
        check_matvec_out_shape(self, out)
        return self._matvec_helper(vec, None, cols, out, False)

    def multiply(self, other):
        """Row-wise scaling by a vector (or element-wise by a matrix)."""
        o, _ = _vec_in(other, self._array.dtype)
        if o.dim() == 1:
            o = o[:, None]
        return type(self)((self._array * o), column_names=self.column_names,
                          term_names=self.term_names)

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


def _np_dtype(x) -> np.dtype:
    if isinstance(x, torch.Tensor):
        return _dev.np_dtype(x.dtype)
    return np.asarray(x).dtype


def _torch_index(ix, device):
    if isinstance(ix, slice):
        return ix
    a = np.asarray(ix)
    if a.dtype == bool:
        return torch.from_numpy(a).to(device)
    return torch.from_numpy(a.astype(np.int64)).to(device)


def _accumulate_out(out, res: torch.Tensor, sel: Optional[torch.Tensor]):
    """out[sel] += res, in place, for a numpy or CUDA ``out``; returns ``out`` itself."""
    if _dev.is_dev(out):
        if sel is None:
            out += res.to(out.dtype)
        else:
            out.index_add_(0, sel.to(torch.int64), res.to(out.dtype))
        return out
    res_h = _dev.to_host(res)
    if sel is None:
        out += res_h
    else:
        out[_dev.to_host(sel)] += res_h
    return out
