"""Row-sharded matrices across GPUs: one process per GPU, ``torch.distributed`` (NCCL over
NVLink / NVSwitch) for the one exchange step the path has.

X^T diag(d) X is additive over disjoint row blocks, so every rank holds the
contiguous row shard ``[lo, hi)`` of every column block and computes its local p x p; a single
sum-allreduce produces the replicated result.  ``transpose_matvec`` / column means reduce a
length-p vector the same way; ``matvec`` needs no collective (the output rows are sharded like
the input rows).  The reference has no multi-process code at all; this layer is new.

For a SplitMatrix the payload of the sandwich allreduce is the flat block workspace of
``tm_split_sandwich_blocks`` (every cross block once, categorical self blocks as diagonals,
block dtype): 98 MB f32 at the 4e7-row benchmark (p = 6388) instead of the 326 MB float64
square; the p x p float64 is assembled after the collective.  Other matrices reduce the packed
lower triangle (p(p+1)/2 elements).
"""

from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous ceil(n / world_size)-row shard of ``rank``: rows ``[lo, hi)``."""
    per = -(-n // world_size)
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def shard_rows(rows, lo: int, hi: int):
    """The part of a sorted global ``rows`` restriction that falls into ``[lo, hi)``, rebased
    to shard-local row ids.  ``None`` (all rows) stays ``None``."""
    if rows is None:
        return None
    rows = np.asarray(rows)
    a, b = np.searchsorted(rows, lo), np.searchsorted(rows, hi)
    return (rows[a:b] - lo).astype(np.int32)


_TRIL_CACHE: dict = {}


def _tril_index(p: int, device) -> torch.Tensor:
    key = (p, str(device))
    if key not in _TRIL_CACHE:
        _TRIL_CACHE[key] = torch.tril_indices(p, p, device=device)
    return _TRIL_CACHE[key]


def pack_lower(sq: torch.Tensor, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Lower triangle (diagonal included) of a symmetric p x p tensor as a flat vector."""
    idx = _tril_index(sq.shape[0], sq.device)
    v = sq[idx[0], idx[1]]
    return v if dtype is None else v.to(dtype)


def unpack_lower(v: torch.Tensor, p: int, dtype: torch.dtype) -> torch.Tensor:
    """Inverse of :func:`pack_lower`: the full symmetric matrix in ``dtype``."""
    idx = _tril_index(p, v.device)
    out = torch.empty((p, p), dtype=dtype, device=v.device)
    vv = v.to(dtype)
    out[idx[0], idx[1]] = vv
    out[idx[1], idx[0]] = vv
    return out


class SharedHostResult:
    """One p x p float64 result buffer in host memory shared by the ranks of a node: a file in
    /dev/shm mapped by every rank and page-locked for CUDA (``cudaHostRegister``), so that every
    rank can copy ITS row band of a row-sharded sandwich straight to the host over its own PCIe
    link (``RowShardedMatrix.sandwich_into_shared``) and the consumer — any rank, or another
    process on the node — reads the whole matrix from ``.array``."""

    def __init__(self, p: int, name: str, create: bool):
        import os

        self.p = int(p)
        self.path = os.path.join("/dev/shm", name)
        nbytes = self.p * self.p * 8
        if create:
            with open(self.path, "wb") as f:
                f.truncate(nbytes)
        self.array = np.memmap(self.path, dtype=np.float64, mode="r+", shape=(self.p, self.p))
        self._registered = False
        if torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self.array.ctypes.data, nbytes, 0)
            self._registered = int(rc) == 0

    @classmethod
    def for_group(cls, p: int, group=None, tag: str = "tabmat_b200_result"):
        """Collective: rank 0 creates the buffer, the others map it."""
        import os

        rank = dist.get_rank(group) if dist.is_initialized() else 0
        name = f"{tag}_{os.environ.get('MASTER_PORT', '0')}_{p}"
        if rank == 0:
            buf = cls(p, name, create=True)
        if dist.is_initialized():
            dist.barrier(group)
        if rank != 0:
            buf = cls(p, name, create=False)
        return buf

    def close(self, unlink: bool = False):
        import os

        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self.array.ctypes.data)
            self._registered = False
        if unlink:
            try:
                os.unlink(self.path)
            except OSError:
                pass


class RowShardedMatrix:
    """A matrix whose rows are sharded over the ranks of a process group.

    ``local`` is this rank's shard (any object with the MatrixBase ``sandwich`` / ``matvec`` /
    ``transpose_matvec`` methods working on ``torch`` tensors that live where ``local``
    computes — CUDA for the tabmat_b200 classes).  ``n_global`` is the total row count.
    Vectors of length n (``d``, ``v`` of transpose_matvec, result of matvec) are sharded the
    same way and are passed / returned as the local slice.
    """

    def __init__(self, local, n_global: int, group=None, pack: bool = True,
                 reduce_dtype: Optional[torch.dtype] = None):
        self.local = local
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lo, self.hi = shard_bounds(n_global, self.world_size, self.rank)
        if local.shape[0] != self.hi - self.lo:
            raise ValueError(
                f"rank {self.rank}: local shard has {local.shape[0]} rows, expected "
                f"{self.hi - self.lo} (rows [{self.lo}, {self.hi}) of {n_global})")
        self.shape = (n_global, local.shape[1])
        self.dtype = local.dtype
        self.pack = pack
        self.reduce_dtype = reduce_dtype
        # TABMAT_B200_DIST_OVERLAP=1: start the allreduce of the index blocks while the
        # dense-operand passes still run (sandwich of a SplitMatrix shard).  Off by default:
        # measured slower (12.75 vs 11.55 ms at 2 GPUs, profiles/bench_r2r_*): the collective
        # shares L2 and SMs with the L2-bound gather kernel, which goes from 3.8 to 5.0 ms
        import os

        # =2: the same with the SMs taken from the tcgen05 kernel (which runs first) instead of
        # the gather kernel
        self.overlap_mode = int(os.environ.get("TABMAT_B200_DIST_OVERLAP", "0") or 0)
        self.overlap = self.overlap_mode in (1, 2)
        self.sm_reserve = int(os.environ.get("TABMAT_B200_DIST_SM_RESERVE", "8"))

    # -- collectives ---------------------------------------------------------------------
    def _allreduce(self, t: torch.Tensor, dst: Optional[int] = None) -> torch.Tensor:
        """Sum over ranks, in place: replicated (allreduce) or only on rank ``dst`` (reduce)."""
        if self.world_size > 1:
            if dst is None:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            else:
                dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM, group=self.group)
        return t

    # -- hot path ------------------------------------------------------------------------
    def sandwich(self, d_local, rows=None, cols=None, dst: Optional[int] = None):
        """(X[rows, cols].T * d[rows]) @ X[rows, cols]; ``d_local`` is this rank's slice of d,
        ``rows`` a sorted GLOBAL restriction (or None).  ``dst=None``: the result is replicated
        on every rank (allreduce); ``dst=r``: it is reduced to rank ``r`` only, the other ranks
        get ``None`` (MPI-style; saves the redundant copies when one process consumes it)."""
        local_rows = shard_rows(rows, self.lo, self.hi)
        if hasattr(self.local, "_sandwich_blocks_dev"):
            # SplitMatrix: allreduce the flat block workspace (every structurally distinct
            # entry once, block dtype) and assemble the p x p float64 after the collective
            from . import _dev

            ws = None
            if (self.world_size > 1 and dst is None and self.overlap
                    and hasattr(self.local, "_sandwich_blocks_overlapped_dev")):
                # the allreduce of the index blocks (95 % of the payload) runs while the
                # dense-operand passes compute; a few SMs are left to the collective
                from ._lib import lib

                if self.overlap_mode == 2:
                    lib.tm_set_tc_sm_reserve(self.sm_reserve)
                else:
                    lib.tm_set_sm_reserve(self.sm_reserve)
                try:
                    ws = self.local._sandwich_blocks_overlapped_dev(
                        d_local, _dev.idx32(local_rows),
                        lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group,
                                                  async_op=True),
                        self._allreduce)
                finally:
                    lib.tm_set_sm_reserve(0)
                    lib.tm_set_tc_sm_reserve(0)
                if ws is not None:
                    return self.local._assemble_dev(ws, cols)
            ws = self.local._sandwich_blocks_dev(d_local, _dev.idx32(local_rows))
            if ws is not None:
                self._allreduce(ws, dst)
                if dst is not None and self.rank != dst:
                    return None
                # a `cols` selection is applied by the assembly (negative destinations dropped)
                return self.local._assemble_dev(ws, cols)
        part = self.local.sandwich(d_local, local_rows, cols)
        if self.world_size == 1:
            return part
        if dst is not None:
            self._allreduce(part, dst)
            return part if self.rank == dst else None
        if part.dim() == 1 or part.shape[0] != part.shape[1]:
            # a categorical block's sandwich is its diagonal; rectangular = a cross block
            return self._allreduce(part)
        p = part.shape[0]
        if self.pack:
            v = pack_lower(part, self.reduce_dtype)
            self._allreduce(v)
            return unpack_lower(v, p, part.dtype)
        if self.reduce_dtype is not None and self.reduce_dtype != part.dtype:
            return self._allreduce(part.to(self.reduce_dtype)).to(part.dtype)
        return self._allreduce(part)

    def sandwich_and_transpose_matvec(self, d_local, v_local, rows=None, cols=None):
        """Replicated ``(X.T diag(d) X, X.T v)`` of one IRLS step: the local fused pass
        (``SplitMatrix.sandwich_and_transpose_matvec``) and ONE allreduce of
        [flat block workspace | X.T v] (the length-p vector rides in the same payload)."""
        local_rows = shard_rows(rows, self.lo, self.hi)
        if hasattr(self.local, "_sandwich_rmatvec_blocks_dev"):
            from . import _dev

            res = self.local._sandwich_rmatvec_blocks_dev(d_local, v_local, _dev.idx32(local_rows))
            if res is not None:
                buf, elems = res
                self._allreduce(buf)
                return (self.local._assemble_dev(buf[:elems], cols),
                        self.local._rmatvec_assemble_dev(buf[elems:], cols))
        return self.sandwich(d_local, rows, cols), self.transpose_matvec(v_local, rows, cols)

    def sandwich_into(self, d_local, out, rows=None, dst: Optional[int] = 0):
        """``sandwich`` with host buffers: ``d_local`` (this rank's slice, host or device) in,
        the p x p float64 result into the host array ``out`` on rank ``dst`` (every rank when
        ``dst`` is None).  One rank: the two-phase path of ``SplitMatrix.sandwich_into`` (host
        copy overlapped with compute).  Asynchronous on the current stream."""
        if hasattr(self.local, "sandwich_into"):
            if self.world_size == 1:
                return self.local.sandwich_into(d_local, out, shard_rows(rows, self.lo, self.hi))

            def reduce(ws):  # sum a slice of the flat block workspace over the ranks
                if ws.numel():
                    self._allreduce(ws, dst)
                return dst is None or self.rank == dst

            return self.local.sandwich_into(d_local, out, shard_rows(rows, self.lo, self.hi),
                                            reduce=reduce)
        from . import _dev
        from ._lib import check, lib

        if not _dev.is_dev(d_local):
            src = d_local if isinstance(d_local, torch.Tensor) else torch.from_numpy(d_local)
            d_dev = torch.empty(src.shape, dtype=src.dtype, device=_dev.require_cuda())
            d_dev.copy_(src, non_blocking=True)
            d_local = d_dev
        res = self.sandwich(d_local, rows, dst=dst)
        if res is None:
            return None
        p = res.shape[0]
        ptr = out.data_ptr() if isinstance(out, torch.Tensor) else out.ctypes.data
        check(lib.tm_memcpy2d_to_host(ptr, p * 8, res.data_ptr(), p * 8, p * 8, p,
                                      _dev.stream_ptr()))
        return out

    def sandwich_into_shared(self, d_local, shared: SharedHostResult, rows=None):
        """Row-sharded ``sandwich`` with the result delivered to a host buffer shared by the
        ranks: local blocks, one allreduce of the flat block workspace, then every rank places
        rows ``[r * p / N, (r + 1) * p / N)`` of the p x p float64 result and copies that band to
        ``shared`` over its own PCIe link — N links instead of the single 8 p^2-byte copy of
        ``sandwich_into(dst=0)``.  A one-element allreduce enqueued after the copy is the
        completion fence: once it has completed on a rank's stream, every band is in place.
        Asynchronous on the current stream.  SplitMatrix-like locals only."""
        from . import _dev
        from ._lib import check, lib

        if not hasattr(self.local, "_assemble_band_dev"):
            raise TypeError("sandwich_into_shared needs a SplitMatrix / RowSortedMatrix shard")
        p = self.shape[1]
        r0, r1 = shard_bounds(p, self.world_size, self.rank)

        def reduce(ws):   # every rank keeps the sum: each one places its own band
            if ws.numel():
                self._allreduce(ws)
            return True

        # the two-phase path of SplitMatrix.sandwich_into, restricted to this rank's band: the
        # blocks without the dense operand are reduced, placed and copied while the dense passes
        # still run
        lib.tm_set_sm_reserve(self.sm_reserve if self.world_size > 1 and self.overlap else 0)
        try:
            self.local.sandwich_into(d_local, shared.array, shard_rows(rows, self.lo, self.hi),
                                     reduce=reduce, band=(r0, r1))
        finally:
            lib.tm_set_sm_reserve(0)
        fence = getattr(self, "_fence", None)
        if fence is None:
            fence = self._fence = torch.zeros(1, dtype=torch.float32, device=_dev.require_cuda())
        self._allreduce(fence)
        return shared.array

    def transpose_matvec(self, v_local, rows=None, cols=None) -> torch.Tensor:
        """Replicated X[rows, cols].T @ v[rows] (length-p allreduce)."""
        part = self.local.transpose_matvec(v_local, shard_rows(rows, self.lo, self.hi), cols)
        return self._allreduce(part)

    def matvec(self, v, cols=None) -> torch.Tensor:
        """This rank's rows of X[:, cols] @ v[cols]; no collective."""
        return self.local.matvec(v, cols)

    def _get_col_means(self, weights_local) -> torch.Tensor:
        """Weighted column means (matrix_base.py:118-120); ``weights`` sums to 1 over ALL rows."""
        return self.transpose_matvec(weights_local)

    def _get_col_stds(self, weights_local, col_means) -> torch.Tensor:
        """sqrt(sum_k w_k (x_kj - mean_j)^2) over all shards: the local sums of squares are
        additive, so the local stds are squared, summed over the ranks and rooted again
        (dense_matrix.py:180-187 / sparse_matrix.py:305-315 per shard)."""
        # Dense blocks use the centred form sum w (x - mean)^2 (dense.pyx:103-122), which IS
        # additive over shards.  Sparse and categorical blocks use sum w x^2 - mean^2
        # (sparse_matrix.py:305-315, categorical_matrix.py:655-672) with the GLOBAL mean: only
        # the raw second moment is additive there (and a shard's share of it may well be below
        # mean^2), so those columns are taken with a zero mean and the mean is subtracted once,
        # after the allreduce.  One payload of 2p values.
        def sq(x):
            x = x if isinstance(x, torch.Tensor) else torch.as_tensor(x)
            return x * x

        centred = sq(self.local._get_col_stds(weights_local, col_means))
        dense = self._dense_column_mask(centred.device)
        if bool(dense.all()):
            return torch.sqrt(self._allreduce(centred))
        raw = sq(self.local._get_col_stds(weights_local, torch.zeros_like(col_means)))
        both = self._allreduce(torch.stack([centred, raw]))
        var = torch.where(dense, both[0], (both[1] - col_means * col_means).clamp_min(0))
        return torch.sqrt(var)

    def _dense_column_mask(self, device) -> torch.Tensor:
        """True for the columns whose local ``_get_col_stds`` is the centred (additive) form."""
        p = self.shape[1]
        mat = getattr(self.local, "mat", self.local)   # RowSortedMatrix wraps a SplitMatrix
        name = type(mat).__name__
        if name in ("SparseMatrix", "CategoricalMatrix"):
            return torch.zeros(p, dtype=torch.bool, device=device)
        mask = torch.ones(p, dtype=torch.bool, device=device)
        if name == "SplitMatrix":
            for m, idx in zip(mat.matrices, mat.indices):
                if type(m).__name__ != "DenseMatrix":
                    mask[torch.as_tensor(np.asarray(idx), dtype=torch.int64, device=device)] = False
        return mask

    def _global_sum(self, x_local: torch.Tensor, rows=None) -> torch.Tensor:
        """sum of x over the (restricted) rows of every shard, as a 1-element tensor."""
        local_rows = shard_rows(rows, self.lo, self.hi)
        if local_rows is None:
            part = x_local.sum().reshape(1)
        else:
            idx = torch.as_tensor(local_rows, dtype=torch.int64, device=x_local.device)
            part = x_local.index_select(0, idx).sum().reshape(1)
        return self._allreduce(part)

    def standardize(self, weights_local, center_predictors: bool, scale_predictors: bool):
        """(RowShardedStandardizedMatrix, column means, column stds): matrix_base.py:128-167 over
        the row shards.  The moments are length-p allreduces; the shift / scale vectors are then
        identical on every rank and the rank-1 corrections of the standardized sandwich are
        applied AFTER the collective (standardized_mat.py:123-172)."""
        col_means = self._get_col_means(weights_local)
        if scale_predictors:
            col_stds = self._get_col_stds(weights_local, col_means)
            # matrix_base.py:248-258: 1 / std, 1 where |std| < 1e-7
            mult = torch.where(col_stds.abs() < 1e-7, torch.ones_like(col_stds), 1.0 / col_stds)
            shifter = -col_means * mult if center_predictors else torch.zeros_like(col_means)
        else:
            col_stds, mult = None, None
            shifter = -col_means if center_predictors else torch.zeros_like(col_means)
        out_means = col_means if center_predictors else torch.zeros_like(col_means)
        return RowShardedStandardizedMatrix(self, shifter, mult), out_means, col_stds


class RowShardedStandardizedMatrix:
    """``StandardizedMatrix`` (standardized_mat.py:17-230) over a :class:`RowShardedMatrix`:
    ``(X + 1 shift^T) diag(mult)`` with X sharded by rows.  Vectors of length n are local
    slices, ``shift`` / ``mult`` are replicated CUDA tensors of length p."""

    def __init__(self, mat: RowShardedMatrix, shift: torch.Tensor, mult: Optional[torch.Tensor]):
        self.mat = mat
        self.shift = shift
        self.mult = mult
        self.shape = mat.shape
        self.dtype = mat.dtype

    def _sel(self, t, cols):
        if t is None or cols is None:
            return t
        return t.index_select(0, torch.as_tensor(np.asarray(cols), dtype=torch.int64,
                                                 device=t.device))

    def sandwich(self, d_local, rows=None, cols=None) -> torch.Tensor:
        """Replicated standardized sandwich: ONE allreduce of [block workspace | X.T d] for a
        SplitMatrix (the inner sandwich and inner.T @ d come from the same pass and the same
        payload), a 1-element allreduce for sum(d), then the rank-1 epilogue kernel."""
        from . import _dev
        from ._lib import check, fn

        term1, d_mat = self.mat.sandwich_and_transpose_matvec(d_local, d_local, rows, cols)
        tdt = d_local.dtype
        sum_d = self.mat._global_sum(d_local, rows).to(tdt)
        shift = self._sel(self.shift, cols).to(tdt).contiguous()
        mult = None if self.mult is None else self._sel(self.mult, cols).to(tdt).contiguous()
        diag = term1.dim() == 1
        if term1.dtype not in (tdt, torch.float64):
            term1 = term1.to(tdt)
        m = int(shift.numel())
        res = torch.empty((m, m), dtype=tdt, device=d_local.device)
        check(fn("tm_std_sandwich_combine", _dev.suffix(tdt))(
            _dev.ptr(term1.contiguous()), int(term1.dtype == torch.float64), int(diag),
            _dev.ptr(d_mat.to(tdt).contiguous()), _dev.ptr(shift), _dev.ptr(mult),
            _dev.ptr(sum_d), m, _dev.ptr(res), _dev.stream_ptr()))
        return res

    def matvec(self, v, cols=None) -> torch.Tensor:
        """This rank's rows of (X[:, cols] * mult[cols]) @ v[cols] + shift[cols] . v[cols]
        (standardized_mat.py:69-109); no collective."""
        vv = v if self.mult is None else v * self.mult.to(v.dtype)
        res = self.mat.matvec(vv, cols)
        sv = self._sel(self.shift.to(v.dtype), cols)
        return res + (sv * self._sel(v, cols)).sum()

    def transpose_matvec(self, v_local, rows=None, cols=None) -> torch.Tensor:
        """Replicated mult[cols] * X[rows, cols].T v[rows] + shift[cols] * sum(v[rows])
        (standardized_mat.py:178-230)."""
        res = self.mat.transpose_matvec(v_local, rows, cols)
        total = self.mat._global_sum(v_local, rows).to(res.dtype)
        if self.mult is not None:
            res = res * self._sel(self.mult, cols).to(res.dtype)
        return res + self._sel(self.shift, cols).to(res.dtype) * total
