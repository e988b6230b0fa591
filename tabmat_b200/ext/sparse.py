"""Mirror of ``tabmat/ext/sparse.pyx`` on the device (reference: sparse.pyx:17-282).

A device sparse block is a :class:`DeviceCSR` / :class:`DeviceCSC` pair of int32-indexed
arrays; the CSR side also carries ``row`` (the row id of every non-zero)."""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from .. import _dev
from .._lib import check, fn
from .dense import dense_layout


@dataclass
class DeviceCSR:
    data: torch.Tensor      # (nnz,) float
    indices: torch.Tensor   # (nnz,) int32 column ids, sorted inside each row
    indptr: torch.Tensor    # (n+1,) int32
    row: torch.Tensor       # (nnz,) int32 row id per non-zero
    shape: tuple

    @property
    def nnz(self) -> int:
        return int(self.data.numel())


@dataclass
class DeviceCSC:
    data: torch.Tensor      # (nnz,) float
    indices: torch.Tensor   # (nnz,) int32 row ids, sorted inside each column
    indptr: torch.Tensor    # (p+1,) int32
    shape: tuple

    @property
    def nnz(self) -> int:
        return int(self.data.numel())


def sparse_sandwich(A: DeviceCSR, d: torch.Tensor, rows: Optional[torch.Tensor],
                    cols: Optional[torch.Tensor]) -> torch.Tensor:
    """A[rows, cols].T @ diag(d[rows]) @ A[rows, cols]   (sparse.pyx:17-77)."""
    n, p = A.shape
    m = p if cols is None else _dev.length(cols)
    out = torch.empty((m, m), dtype=A.data.dtype, device=A.data.device)
    if m == 0:
        return out
    check(fn("tm_sparse_sandwich", _dev.suffix(A.data.dtype))(
        _dev.ptr(A.data), _dev.ptr(A.indices), _dev.ptr(A.indptr), _dev.ptr(A.row), n, p, A.nnz,
        _dev.ptr(d), _dev.ptr(rows), _dev.length(rows), _dev.ptr(cols), _dev.length(cols),
        _dev.ptr(out), _dev.stream_ptr()))
    return out


def csr_dense_sandwich(A: DeviceCSR, B: torch.Tensor, d: torch.Tensor,
                       rows: Optional[torch.Tensor], A_cols: Optional[torch.Tensor],
                       B_cols: Optional[torch.Tensor]) -> torch.Tensor:
    """(A[rows, A_cols].T * d[rows]) @ B[rows, B_cols]   (sparse.pyx:211-260)."""
    n, p = A.shape
    nb_rows, q, b_c_order = dense_layout(B)
    nA = p if A_cols is None else _dev.length(A_cols)
    nB = q if B_cols is None else _dev.length(B_cols)
    out = torch.empty((nA, nB), dtype=A.data.dtype, device=A.data.device)
    if nA == 0 or nB == 0:
        return out
    if (rows is not None and _dev.length(rows) == 0) or A.nnz == 0:
        return out.zero_()
    check(fn("tm_csr_dense_sandwich", _dev.suffix(A.data.dtype))(
        _dev.ptr(A.data), _dev.ptr(A.indices), _dev.ptr(A.indptr), n, p, _dev.ptr(B), q, b_c_order,
        _dev.ptr(d), _dev.ptr(rows), _dev.length(rows), _dev.ptr(A_cols), _dev.length(A_cols),
        _dev.ptr(B_cols), _dev.length(B_cols), _dev.ptr(out), _dev.stream_ptr()))
    return out


def csr_matvec(A: DeviceCSR, v: torch.Tensor, rows: Optional[torch.Tensor],
               cols: Optional[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """res[t] = sum_{j in cols} A[rows[t], j] * v[j]; accumulated into ``out`` when given
    (csr_matvec_unrestricted / csr_matvec, sparse.pyx:79-140)."""
    n, p = A.shape
    nr = n if rows is None else _dev.length(rows)
    accumulate = out is not None
    if out is None:
        out = torch.empty((nr,), dtype=A.data.dtype, device=A.data.device)
    if nr == 0:
        return out
    check(fn("tm_csr_matvec", _dev.suffix(A.data.dtype))(
        _dev.ptr(A.data), _dev.ptr(A.indices), _dev.ptr(A.indptr), n, p, _dev.ptr(v),
        _dev.ptr(rows), _dev.length(rows), _dev.ptr(cols), _dev.length(cols), _dev.ptr(out),
        1 if accumulate else 0, _dev.stream_ptr()))
    return out


def csc_rmatvec(A: DeviceCSC, v: torch.Tensor, rows: Optional[torch.Tensor],
                cols: Optional[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """res[c] = sum_{i in rows} A[i, cols[c]] * v[i]; accumulated into ``out`` when given
    (csc_rmatvec_unrestricted / csc_rmatvec, sparse.pyx:142-199)."""
    n, p = A.shape
    nc = p if cols is None else _dev.length(cols)
    accumulate = out is not None
    if out is None:
        out = torch.empty((nc,), dtype=A.data.dtype, device=A.data.device)
    if nc == 0:
        return out
    check(fn("tm_csc_rmatvec", _dev.suffix(A.data.dtype))(
        _dev.ptr(A.data), _dev.ptr(A.indices), _dev.ptr(A.indptr), n, p, _dev.ptr(v),
        _dev.ptr(rows), _dev.length(rows), _dev.ptr(cols), _dev.length(cols), _dev.ptr(out),
        1 if accumulate else 0, _dev.stream_ptr()))
    return out


def transpose_square_dot_weights(A: DeviceCSC, weights: torch.Tensor) -> torch.Tensor:
    """out[j] = sum_{nz (i,j)} weights[i] * A[i,j]**2   (sparse.pyx:262-282)."""
    n, p = A.shape
    out = torch.empty((p,), dtype=A.data.dtype, device=A.data.device)
    if p == 0:
        return out
    check(fn("tm_csc_sq_dot_weights", _dev.suffix(A.data.dtype))(
        _dev.ptr(A.data), _dev.ptr(A.indices), _dev.ptr(A.indptr), n, p, _dev.ptr(weights),
        _dev.ptr(out), _dev.stream_ptr()))
    return out
