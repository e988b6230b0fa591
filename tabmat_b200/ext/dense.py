"""Mirror of ``tabmat/ext/dense.pyx`` on the device (reference: dense.pyx:19-122)."""

from __future__ import annotations

from typing import Optional

import torch

from .. import _dev
from .._lib import check, fn


def dense_layout(X: torch.Tensor):
    """(n, p, c_order) of a 2-D CUDA tensor that is C- or F-contiguous."""
    if X.dim() != 2:
        raise ValueError("X must be 2-dimensional")
    n, p = X.shape
    if X.is_contiguous():
        return n, p, 1
    if X.t().is_contiguous():
        return n, p, 0
    # dense.pyx:43
    raise Exception("The matrix X is not contiguous.")


def dense_sandwich(X: torch.Tensor, d: torch.Tensor, rows: Optional[torch.Tensor],
                   cols: Optional[torch.Tensor]) -> torch.Tensor:
    """out = X[rows, cols].T @ diag(d[rows]) @ X[rows, cols]   (dense.pyx:19-44)."""
    n, p, c_order = dense_layout(X)
    m = p if cols is None else _dev.length(cols)
    out = torch.empty((m, m), dtype=X.dtype, device=X.device)
    if m == 0:
        return out
    if (rows is not None and _dev.length(rows) == 0) or n == 0:
        return out.zero_()
    check(fn("tm_dense_sandwich", _dev.suffix(X.dtype))(
        _dev.ptr(X), n, p, c_order, _dev.ptr(d), _dev.ptr(rows), _dev.length(rows),
        _dev.ptr(cols), _dev.length(cols), _dev.ptr(out), _dev.stream_ptr()))
    return out


def dense_rmatvec(X: torch.Tensor, v: torch.Tensor, rows: Optional[torch.Tensor],
                  cols: Optional[torch.Tensor]) -> torch.Tensor:
    """out[c] = sum_{i in rows} X[i, cols[c]] * v[i]   (dense.pyx:48-73)."""
    n, p, c_order = dense_layout(X)
    m = p if cols is None else _dev.length(cols)
    out = torch.empty((m,), dtype=X.dtype, device=X.device)
    if m == 0:
        return out
    if (rows is not None and _dev.length(rows) == 0) or n == 0:
        return out.zero_()
    check(fn("tm_dense_rmatvec", _dev.suffix(X.dtype))(
        _dev.ptr(X), n, p, c_order, _dev.ptr(v), _dev.ptr(rows), _dev.length(rows),
        _dev.ptr(cols), _dev.length(cols), _dev.ptr(out), _dev.stream_ptr()))
    return out


def dense_matvec(X: torch.Tensor, v: torch.Tensor, rows: Optional[torch.Tensor],
                 cols: Optional[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """res[r] = sum_{c in cols} X[rows[r], c] * v[c]   (dense.pyx:76-101).

    With ``out`` the result is accumulated into it in place (the reference adds in Python,
    dense_matrix.py:229-236)."""
    n, p, c_order = dense_layout(X)
    nr = n if rows is None else _dev.length(rows)
    accumulate = out is not None
    if out is None:
        out = torch.empty((nr,), dtype=X.dtype, device=X.device)
    if nr == 0:
        return out
    if (cols is not None and _dev.length(cols) == 0) or p == 0:
        return out if accumulate else out.zero_()
    check(fn("tm_dense_matvec", _dev.suffix(X.dtype))(
        _dev.ptr(X), n, p, c_order, _dev.ptr(v), _dev.ptr(rows), _dev.length(rows),
        _dev.ptr(cols), _dev.length(cols), _dev.ptr(out), 1 if accumulate else 0,
        _dev.stream_ptr()))
    return out


def transpose_square_dot_weights(X: torch.Tensor, weights: torch.Tensor,
                                 shift: torch.Tensor) -> torch.Tensor:
    """out[j] = sum_i weights[i] * (X[i, j] - shift[j])**2   (dense.pyx:103-122)."""
    n, p, c_order = dense_layout(X)
    out = torch.empty((p,), dtype=X.dtype, device=X.device)
    if p == 0:
        return out
    check(fn("tm_dense_sq_dot_weights", _dev.suffix(X.dtype))(
        _dev.ptr(X), n, p, c_order, _dev.ptr(weights), _dev.ptr(shift), _dev.ptr(out),
        _dev.stream_ptr()))
    return out
