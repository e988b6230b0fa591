"""Row-sorted storage for a SplitMatrix — a B200-side layout choice with no counterpart in the
reference, which keeps the caller's row order (``split_matrix.py:171-267``).

X^T diag(d) X, X^T v and the column moments do not depend on the order of the rows, and
X v only needs its result put back.  On B200 the SplitMatrix sandwich is bound by the L2
atomic units: every row sends one 512-byte vector RED per many-level categorical block
(dense x categorical, ``cat_split_helpers-tmpl.cpp:97-151``) and per sparse non-zero.  If
the rows are STORED sorted by the codes of the many-level categorical blocks, equal codes sit
in consecutive rows, a warp can add a whole run in registers and issue one RED per run
(``k_dense_cross_runs``, ``k_cat_hist2``, ``k_cat_cat2``), and the categorical share of that
traffic all but disappears.

:class:`RowSortedMatrix` wraps the re-ordered SplitMatrix and keeps the MatrixBase surface in
the CALLER's row order: length-n inputs (``d``, ``v`` of transpose_matvec, weights) are
gathered through the permutation on the way in (``tm_permute_gather``), matvec results are
scattered back (``tm_permute_scatter``), ``rows`` restrictions are mapped through the inverse
permutation.  The sort happens once at construction, like the reference's cached CSR
(``sparse_matrix.py:133-143``).
"""

from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import _dev
from ._lib import check, fn
from .categorical_matrix import CategoricalMatrix
from .matrix_base import MatrixBase, _vec_in
from .split_matrix import SplitMatrix
from .util import (
    check_matvec_dimensions,
    check_matvec_out_shape,
    check_sandwich_compatible,
    check_transpose_matvec_out_shape,
)

#: categorical blocks at most this wide ride along the tcgen05 pass as one-hot MMAs
#: (csrc/split.cu) and gain nothing from sorted rows
_ONEHOT_MAX_LEVELS = 256


def choose_sort_blocks(matrices, max_keys: int = 2) -> List[int]:
    """Indices (into ``matrices``) of the categorical blocks to sort the rows by, primary key
    first: the widest blocks that are too wide for the one-hot tensor path, else the widest
    categorical block, else nothing."""
    cats = [(m.shape[1], i) for i, m in enumerate(matrices) if isinstance(m, CategoricalMatrix)]
    wide = sorted([c for c in cats if c[0] > _ONEHOT_MAX_LEVELS], reverse=True)
    if wide:
        return [i for _, i in wide[:max_keys]]
    return [max(cats)[1]] if cats else []


def sort_permutation(matrices, blocks: List[int]) -> Optional[torch.Tensor]:
    """Stable permutation (int64 CUDA tensor, stored row i = original row perm[i]) that orders
    the rows lexicographically by the codes of ``blocks``; missing codes (-1) sort last."""
    if not blocks:
        return None
    key = None
    for b in blocks:
        m = matrices[b]
        K = len(m.categories)
        c = m._codes.to(torch.int64)
        c = torch.where(c < 0, torch.full_like(c, K), c)
        key = c if key is None else key * (K + 1) + c
    return torch.argsort(key, stable=True)


class RowSortedMatrix(MatrixBase):
    """A SplitMatrix stored with its rows sorted for the kernels; behaves like the original.

    ``RowSortedMatrix.from_split(X)`` picks the sort keys, re-orders every block on the device
    and wraps the result.  All MatrixBase methods take and return vectors in the ORIGINAL row
    order.
    """

    def __init__(self, sorted_mat: SplitMatrix, perm: torch.Tensor, sort_blocks=()):
        n = sorted_mat.shape[0]
        if perm.numel() != n:
            raise ValueError(f"perm has {perm.numel()} entries, the matrix {n} rows")
        if n >= 2**31:
            raise ValueError("RowSortedMatrix is int32-indexed per GPU shard")
        self.mat = sorted_mat
        self.shape = sorted_mat.shape
        self.dtype = sorted_mat.dtype
        self.indices = sorted_mat.indices
        self.sort_blocks = tuple(sort_blocks)
        self._perm = perm.to(torch.int32).contiguous()
        inv = torch.empty(n, dtype=torch.int32, device=perm.device)
        inv[perm.to(torch.int64)] = torch.arange(n, dtype=torch.int32, device=perm.device)
        self._inv = inv

    # ---- construction ------------------------------------------------------------------
    @classmethod
    def from_split(cls, X: SplitMatrix, max_keys: int = 2, blocks: Optional[List[int]] = None):
        if not isinstance(X, SplitMatrix):
            raise TypeError("RowSortedMatrix.from_split expects a SplitMatrix")
        blocks = choose_sort_blocks(X.matrices, max_keys) if blocks is None else list(blocks)
        perm = sort_permutation(X.matrices, blocks)
        if perm is None:
            perm = torch.arange(X.shape[0], device=_dev.require_cuda())
            sorted_mat = X
        else:
            sorted_mat = SplitMatrix([m[perm, :] if not isinstance(m, CategoricalMatrix)
                                      else m[perm] for m in X.matrices], X.indices)
        for q, b in enumerate(blocks):
            # block order is preserved by the SplitMatrix constructor for already-combined input;
            # tm_block_desc.flags: bit 0 = sort key, bit 1 = primary sort key
            sorted_mat.matrices[b]._run_sorted = 3 if q == 0 else 1
        return cls(sorted_mat, perm, blocks)

    @property
    def matrices(self):
        return self.mat.matrices

    def unsorted(self) -> SplitMatrix:
        """The same matrix back in the caller's row order (a plain SplitMatrix)."""
        return self.mat[self._inv.to(torch.int64)]

    # ---- vector plumbing ---------------------------------------------------------------
    def _gather(self, v_t: torch.Tensor) -> torch.Tensor:
        """v in stored order: out[i] = v[perm[i]] (1-d through the C-ABI kernel)."""
        if v_t.dim() != 1 or v_t.dtype not in (torch.float32, torch.float64):
            return v_t.index_select(0, self._perm.to(torch.int64))
        out = torch.empty_like(v_t)
        check(fn("tm_permute_gather", _dev.suffix(v_t.dtype))(
            _dev.ptr(v_t), _dev.ptr(self._perm), v_t.numel(), _dev.ptr(out), 0,
            _dev.stream_ptr()))
        return out

    def _scatter(self, y_t: torch.Tensor) -> torch.Tensor:
        """y from stored order back to the original one: out[perm[i]] = y[i]."""
        if y_t.dim() != 1 or y_t.dtype not in (torch.float32, torch.float64):
            return y_t.index_select(0, self._inv.to(torch.int64))
        out = torch.empty_like(y_t)
        check(fn("tm_permute_scatter", _dev.suffix(y_t.dtype))(
            _dev.ptr(y_t), _dev.ptr(self._perm), y_t.numel(), _dev.ptr(out), 0,
            _dev.stream_ptr()))
        return out

    def to_stored_order(self, v) -> torch.Tensor:
        """A length-n vector from the caller's row order into the stored one (CUDA tensor).  For
        per-row data that an iteration keeps on the device (the response, offsets, prior
        weights): permuted ONCE, so that ``irls_step(..., stored_order=True)`` needs no
        permutation per iteration."""
        v_t, _ = _vec_in(v if _dev.is_dev(v) else np.asarray(v))
        return self._gather(v_t)

    def from_stored_order(self, y) -> torch.Tensor:
        """The inverse of :meth:`to_stored_order`."""
        y_t, _ = _vec_in(y if _dev.is_dev(y) else np.asarray(y))
        return self._scatter(y_t)

    def _rows_in(self, rows) -> Optional[torch.Tensor]:
        """Original row ids -> stored positions, ascending (so that runs stay runs)."""
        r = _dev.idx32(rows)
        if r is None:
            return None
        pos = self._inv[r.to(torch.int64)]
        return torch.sort(pos).values.contiguous()

    # ---- hot path ----------------------------------------------------------------------
    def sandwich(self, d, rows=None, cols=None):
        """X[rows, cols].T @ diag(d[rows]) @ X[rows, cols] (split_matrix.py:324-356)."""
        if not _dev.is_dev(d):
            d = np.asarray(d)
        check_sandwich_compatible(self, d)
        d_t, host = _vec_in(d)
        out = self.mat._sandwich_dev(self._gather(d_t), self._rows_in(rows), cols)
        return _dev.ret(out, host)

    def sandwich_and_transpose_matvec(self, d, v, rows=None, cols=None):
        """Fused IRLS pass (see ``SplitMatrix.sandwich_and_transpose_matvec``); ``d`` and ``v``
        in the caller's row order."""
        if not _dev.is_dev(d):
            d = np.asarray(d)
        if not _dev.is_dev(v):
            v = np.asarray(v)
        check_sandwich_compatible(self, d)
        check_matvec_dimensions(self, v, transpose=True)
        d_t, host = _vec_in(d)
        v_t, _ = _vec_in(v)
        H, g = self.mat._sandwich_rmatvec_dev(self._gather(d_t), self._gather(v_t),
                                              self._rows_in(rows), cols)
        return _dev.ret(H, host), _dev.ret(g, host)

    def _sandwich_rmatvec_blocks_dev(self, d_t, v_t, rows_t):
        return self.mat._sandwich_rmatvec_blocks_dev(self._gather(d_t), self._gather(v_t),
                                                     self._rows_in(rows_t))

    def _rmatvec_assemble_dev(self, vec, cols=None):
        return self.mat._rmatvec_assemble_dev(vec, cols)

    def sandwich_into(self, d, out, rows=None, reduce=None, band=None):
        """Host-buffer form of :meth:`sandwich` (see ``SplitMatrix.sandwich_into``): ``d`` from
        host or device memory in the caller's row order, the result into the host array ``out``."""
        if _dev.is_dev(d):
            d_t = d.contiguous()
        else:
            src = d if isinstance(d, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(d))
            d_t = torch.empty(src.shape, dtype=src.dtype, device=_dev.require_cuda())
            d_t.copy_(src, non_blocking=True)
        check_sandwich_compatible(self, d_t)
        return self.mat._sandwich_into_dev(self._gather(d_t), self._rows_in(rows), out, reduce,
                                           band)

    def _sandwich_blocks_dev(self, d_t: torch.Tensor, rows_t):
        """Flat block workspace (for the row-sharded allreduce, distributed.py)."""
        return self.mat._sandwich_blocks_dev(self._gather(d_t), self._rows_in(rows_t))

    def _sandwich_blocks_overlapped_dev(self, d_t, rows_t, reduce_async, reduce):
        return self.mat._sandwich_blocks_overlapped_dev(self._gather(d_t), self._rows_in(rows_t),
                                                        reduce_async, reduce)

    def _assemble_dev(self, ws: torch.Tensor, cols=None) -> torch.Tensor:
        return self.mat._assemble_dev(ws, cols)

    def _assemble_band_dev(self, ws: torch.Tensor, row0: int, row1: int) -> torch.Tensor:
        return self.mat._assemble_band_dev(ws, row0, row1)

    def transpose_matvec(self, v, rows=None, cols=None, out=None):
        """X[rows, cols].T @ v[rows] (split_matrix.py:419-460)."""
        if not _dev.is_dev(v):
            v = np.asarray(v)
        check_matvec_dimensions(self, v, transpose=True)
        check_transpose_matvec_out_shape(self, out)
        v_t, host = _vec_in(v)
        # `out` lives in column space: the row order does not touch it
        res = self.mat.transpose_matvec(self._gather(v_t), self._rows_in(rows), cols, out=out)
        if out is not None:
            return res
        return _dev.ret(res, host) if _dev.is_dev(res) else res

    def matvec(self, v, cols=None, out=None):
        """X[:, cols] @ v[cols], rows in the caller's order (split_matrix.py:373-417)."""
        if not _dev.is_dev(v):
            v = np.asarray(v)
        check_matvec_dimensions(self, v, transpose=False)
        check_matvec_out_shape(self, out)
        host = not _dev.is_dev(v)
        v_t, _ = _vec_in(v)
        res = self._scatter(self.mat.matvec(v_t, cols, out=None))
        if out is None:
            return _dev.ret(res, host)
        from .dense_matrix import _accumulate_out

        return _accumulate_out(out, res, None)

    # ---- the rest of the MatrixBase surface ----------------------------------------------
    def _get_col_means(self, weights):
        w_t, host = _vec_in(weights)
        return _dev.ret(self.mat._get_col_means(self._gather(w_t)), host)

    def _get_col_stds(self, weights, col_means):
        w_t, host = _vec_in(weights)
        cm_t, _ = _vec_in(col_means)
        return _dev.ret(self.mat._get_col_stds(self._gather(w_t), cm_t), host)

    def getcol(self, i: int):
        return self.mat.getcol(i)[self._inv.to(torch.int64), :]

    def toarray(self) -> np.ndarray:
        return self.mat.toarray()[_dev.to_host(self._inv).astype(np.int64)]

    def astype(self, dtype, order="K", casting="unsafe", copy=True):
        new = RowSortedMatrix(self.mat.astype(dtype, order, casting, copy), self._perm,
                              self.sort_blocks)
        for q, b in enumerate(self.sort_blocks):
            new.mat.matrices[b]._run_sorted = 3 if q == 0 else 1
        return new

    def multiply(self, other):
        o_t, _ = _vec_in(np.asarray(other) if not _dev.is_dev(other) else other)
        return RowSortedMatrix(self.mat.multiply(self._gather(o_t.reshape(-1))), self._perm,
                               self.sort_blocks)

    def __getitem__(self, key):
        """Row subsets come back as a plain SplitMatrix in the requested order."""
        if isinstance(key, tuple):
            row, col = key
        else:
            row, col = key, slice(None, None, None)
        if not (isinstance(col, slice) and col == slice(None, None, None)):
            raise NotImplementedError(f"Only row indexing is supported. Index passed was {key}.")
        if isinstance(row, int):
            row = [row]
        if isinstance(row, slice):
            pos = self._inv[row]
        elif _dev.is_dev(row):
            pos = self._inv[row if row.dtype == torch.bool else row.to(torch.int64)]
        else:
            a = np.asarray(row).reshape(-1)
            a = np.flatnonzero(a) if a.dtype == bool else a.astype(np.int64)
            pos = self._inv[torch.from_numpy(a).to(self._inv.device)]
        return self.mat[pos.to(torch.int64)]

    def get_names(self, type: str = "column", missing_prefix: Optional[str] = None,
                  indices=None):
        return self.mat.get_names(type, missing_prefix, indices)

    def set_names(self, names, type: str = "column"):
        self.mat.set_names(names, type)

    def __repr__(self):
        return (f"RowSortedMatrix(rows sorted by blocks {list(self.sort_blocks)}):\n"
                + repr(self.mat))

    __array_priority__ = 13
