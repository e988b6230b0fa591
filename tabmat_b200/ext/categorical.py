"""Mirror of ``tabmat/ext/categorical.pyx`` on the device (reference: categorical.pyx:23-218).

The reference splits every function into ``_fast`` (no drop_first, no missing) and ``_complex``
variants; on the GPU the branch costs nothing, so one kernel takes ``drop_first`` and skips
negative columns."""

from __future__ import annotations

from typing import Optional

import torch

from .. import _dev
from .._lib import check, fn


def transpose_matvec(indices: torch.Tensor, other: torch.Tensor, n_cols: int,
                     rows: Optional[torch.Tensor], cols: Optional[torch.Tensor],
                     out: torch.Tensor, drop_first: bool = False) -> None:
    """out[indices[k]-drop_first] += other[k] for k in rows, restricted to ``cols``; writes at
    the ABSOLUTE column index (transpose_matvec_{fast,complex}, categorical.pyx:23-117)."""
    n = int(indices.numel())
    if n == 0 or n_cols == 0:
        return
    if rows is not None and _dev.length(rows) == 0:
        return
    check(fn("tm_cat_transpose_matvec", _dev.suffix(other.dtype))(
        _dev.ptr(indices), n, _dev.ptr(other), _dev.ptr(rows), _dev.length(rows), _dev.ptr(cols),
        _dev.length(cols), n_cols, int(drop_first), _dev.ptr(out), _dev.stream_ptr()))


def matvec(indices: torch.Tensor, other: torch.Tensor, n_rows: int,
           cols: Optional[torch.Tensor], n_cols: int, out_vec: torch.Tensor,
           drop_first: bool = False) -> None:
    """out_vec[i] += other[indices[i]-drop_first] (matvec_{fast,complex}, categorical.pyx:128-180)."""
    if n_rows == 0 or n_cols == 0:
        return
    if cols is not None and _dev.length(cols) == 0:
        return
    check(fn("tm_cat_matvec", _dev.suffix(other.dtype))(
        _dev.ptr(indices), n_rows, _dev.ptr(other), _dev.ptr(cols), _dev.length(cols), n_cols,
        int(drop_first), _dev.ptr(out_vec), _dev.stream_ptr()))


def sandwich_categorical(indices: torch.Tensor, d: torch.Tensor, rows: Optional[torch.Tensor],
                         n_cols: int, drop_first: bool = False) -> torch.Tensor:
    """diag[indices[k]-drop_first] += d[k] for k in rows
    (sandwich_categorical_{fast,complex}, categorical.pyx:183-218)."""
    n = int(indices.numel())
    out = torch.empty((n_cols,), dtype=d.dtype, device=d.device)
    if n_cols == 0:
        return out
    if n == 0 or (rows is not None and _dev.length(rows) == 0):
        return out.zero_()
    check(fn("tm_cat_sandwich", _dev.suffix(d.dtype))(
        _dev.ptr(indices), n, _dev.ptr(d), _dev.ptr(rows), _dev.length(rows), n_cols,
        int(drop_first), _dev.ptr(out), _dev.stream_ptr()))
    return out
