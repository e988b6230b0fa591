"""SplitMatrix: column blocks of dense / sparse / categorical parts (reference:
split_matrix.py:22-554).

``sandwich`` computes every self block and every cross block on the device and places them
straight into the p x p float64 output with scatter kernels (the reference assembles the
output with numpy fancy indexing on the host, split_matrix.py:336-354).  Like the reference,
all dense blocks are merged into one DenseMatrix and all sparse blocks into one SparseMatrix
at construction (split_matrix.py:85-141)."""

from __future__ import annotations

import ctypes as C
import os
import warnings
from collections.abc import Sequence
from typing import Optional, Union

import numpy as np
import torch
from scipy import sparse as sps

from . import _dev
from ._lib import CSC_ROW_BLOCK, BlockDesc, check, fn, lib
from .categorical_matrix import CategoricalMatrix
from .dense_matrix import DenseMatrix, _accumulate_out
from .ext.split import dense_cross_sandwich, is_sorted, split_col_subsets
from .matrix_base import MatrixBase, _vec_in
from .sparse_matrix import SparseMatrix
from .standardized_mat import StandardizedMatrix
from .util import (
    check_matvec_dimensions,
    check_matvec_out_shape,
    check_sandwich_compatible,
    check_transpose_matvec_out_shape,
)


def as_tabmat(a):
    """Convert an array to the corresponding MatrixBase type (split_matrix.py:22-38)."""
    if isinstance(a, (MatrixBase, StandardizedMatrix)):
        return a
    elif sps.issparse(a):
        return SparseMatrix(a.tocsc(copy=False))
    elif isinstance(a, np.ndarray) or _dev.is_dev(a):
        return DenseMatrix(a)
    else:
        raise ValueError(f"Cannot convert type {type(a)} to Matrix.")


def hstack(tup: Sequence) -> MatrixBase:
    """Stack matrices horizontally (split_matrix.py:41-63)."""
    matrices = [as_tabmat(a) for a in tup]
    if len(matrices) == 0:
        raise ValueError("Need at least one array to concatenate.")
    if all(isinstance(mat, SparseMatrix) for mat in matrices):
        return _hstack_sparse(matrices)
    elif all(isinstance(mat, DenseMatrix) for mat in matrices):
        return _hstack_dense(matrices)
    else:
        return SplitMatrix(matrices)


def _hstack_dense(mats) -> DenseMatrix:
    arrs = [m._array for m in mats]
    if all(a.t().is_contiguous() and not a.is_contiguous() for a in arrs):
        # all F order: stack the transposes' rows, keep F order
        return DenseMatrix(torch.cat([a.t() for a in arrs], dim=0).t())
    return DenseMatrix(torch.cat(arrs, dim=1).contiguous())


def _hstack_sparse(mats) -> SparseMatrix:
    """Column-concatenate sparse blocks on the device (CSR entries interleaved per row)."""
    n = mats[0].shape[0]
    dev = mats[0]._csr.data.device
    counts = [(m._csr.indptr[1:] - m._csr.indptr[:-1]).to(torch.int64) for m in mats]
    total = torch.stack(counts, 0).sum(0)
    new_indptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    new_indptr[1:] = torch.cumsum(total, 0)
    nnz = int(new_indptr[-1].item())
    if nnz > np.iinfo(np.int32).max:
        raise ValueError("merged sparse block exceeds int32 indexing")
    data = torch.empty(nnz, dtype=mats[0]._csr.data.dtype, device=dev)
    indices = torch.empty(nnz, dtype=torch.int32, device=dev)
    prefix = torch.zeros(n, dtype=torch.int64, device=dev)
    col_off = 0
    for m, cnt in zip(mats, counts):
        c = m._csr
        row = c.row.to(torch.int64)
        pos = torch.arange(c.nnz, device=dev, dtype=torch.int64) - c.indptr.to(torch.int64)[row]
        dest = new_indptr[row] + prefix[row] + pos
        data[dest] = c.data.to(data.dtype)
        indices[dest] = c.indices + col_off
        prefix += cnt
        col_off += m.shape[1]
    return SparseMatrix.from_device_csr(data, indices, new_indptr.to(torch.int32), (n, col_off))


def _filter_out_empty(matrices, indices):
    keep_idxs = [i for i, m in enumerate(matrices) if m.shape[1] > 0]
    return [matrices[i] for i in keep_idxs], [indices[i] for i in keep_idxs]


def _combine_matrices(matrices, indices):
    """Merge all DenseMatrix blocks into one and all SparseMatrix blocks into one; categorical
    blocks stay as they are (split_matrix.py:85-141)."""
    n_row = matrices[0].shape[0]
    for mat_type_, stack_fn in [(DenseMatrix, _hstack_dense), (SparseMatrix, _hstack_sparse)]:
        this_type = [i for i, mat in enumerate(matrices) if isinstance(mat, mat_type_)]
        if len(this_type) > 1:
            new_matrix = stack_fn([matrices[i] for i in this_type])
            new_indices = np.concatenate([indices[i] for i in this_type])
            new_colnames = np.concatenate(
                [np.array(matrices[i]._colnames, dtype=object) for i in this_type])
            new_terms = np.concatenate(
                [np.array(matrices[i]._terms, dtype=object) for i in this_type])
            sorter = np.argsort(new_indices)
            if np.array_equal(sorter, np.arange(len(sorter))):
                sorted_matrix = new_matrix
            else:
                sorted_matrix = new_matrix[:, sorter]
            sorted_matrix._colnames = new_colnames[sorter].tolist()
            sorted_matrix._terms = new_terms[sorter].tolist()
            assert sorted_matrix.shape[0] == n_row
            matrices[this_type[0]] = sorted_matrix
            indices[this_type[0]] = new_indices[sorter]
            indices = [idx for i, idx in enumerate(indices) if i not in this_type[1:]]
            matrices = [mat for i, mat in enumerate(matrices) if i not in this_type[1:]]
    return matrices, indices


class SplitMatrix(MatrixBase):
    """Matrix with dense, sparse and categorical column blocks; same API as
    ``tabmat.SplitMatrix``."""

    def __init__(self, matrices: Sequence[MatrixBase], indices: Optional[list] = None):
        flatten_matrices = []
        index_corrections = []
        for mat in matrices:
            if not isinstance(mat, MatrixBase):
                raise ValueError("Expected all elements of matrices to be subclasses of MatrixBase.")
            if isinstance(mat, SplitMatrix):
                current_idx = 0
                for iind, imat in zip(mat.indices, mat.matrices):
                    flatten_matrices.append(imat)
                    index_corrections.append(
                        iind - np.arange(len(iind), dtype=np.int64) - current_idx)
                    current_idx += len(iind)
            else:
                flatten_matrices.append(mat)
                index_corrections.append(np.zeros(mat.shape[1], dtype=np.int64))

        self.dtype = flatten_matrices[0].dtype
        n_row = flatten_matrices[0].shape[0]
        for i, mat in enumerate(flatten_matrices):
            if mat.dtype != self.dtype:
                warnings.warn(
                    "Matrices do not all have the same dtype. Dtypes are "
                    f"{[elt.dtype for elt in flatten_matrices]}."
                )
            if not mat.shape[0] == n_row:
                raise ValueError(
                    "All matrices should have the same first dimension, "
                    f"but the first matrix has first dimension {n_row} and matrix {i} "
                    f"has first dimension {mat.shape[0]}."
                )

        if indices is None:
            indices = []
            current_idx = 0
            for mat, ind_corr in zip(flatten_matrices, index_corrections):
                indices.append(
                    np.arange(current_idx, current_idx + mat.shape[1], dtype=np.int64) + ind_corr)
                current_idx += mat.shape[1]
            n_col = current_idx
        else:
            indices = list(indices)
            all_indices = np.concatenate(indices)
            n_col = len(all_indices)
            if (np.arange(n_col, dtype=np.int64) != np.sort(all_indices)).any():
                raise ValueError(
                    "Indices should contain all integers from 0 to one less than the "
                    "number of columns."
                )
            for i in range(len(indices)):
                indices[i] = np.asarray(indices[i])
                if not is_sorted(indices[i]):
                    raise ValueError(
                        f"Each index block should be sorted, but indices[{i}] was not sorted"
                    )

        for i, (mat, idx) in enumerate(zip(flatten_matrices, indices)):
            if not mat.shape[1] == len(idx):
                raise ValueError(
                    f"Element {i} of indices should should have length {mat.shape[1]}, "
                    f"but it has shape {idx.shape}"
                )

        filtered_mats, filtered_idxs = _filter_out_empty(flatten_matrices, indices)
        combined_matrices, combined_indices = _combine_matrices(filtered_mats, filtered_idxs)

        self.matrices = combined_matrices
        self.indices = [np.asarray(elt, dtype=np.int64) for elt in combined_indices]
        self.shape = (n_row, n_col)
        self._indices_dev: Optional[list] = None
        assert self.shape[1] > 0

    # ---- helpers ---------------------------------------------------------------------
    def _dev_indices(self):
        if self._indices_dev is None:
            self._indices_dev = [_dev.to_dev(idx) for idx in self.indices]
        return self._indices_dev

    def _native_indices(self):
        """Per block: destination column of every STORED column (int64 numpy), -1 for the zero
        padding columns of a dense block (DenseMatrix._store)."""
        out = []
        for mat, idx in zip(self.matrices, self.indices):
            extra = mat._native().shape[1] - mat.shape[1] if isinstance(mat, DenseMatrix) else 0
            out.append(np.concatenate([idx, np.full(extra, -1, dtype=np.int64)]) if extra else idx)
        return out

    def _block_order(self):
        """(int32 CUDA tensor: column position of every block column, blocks concatenated;
        block offsets).  One ``tm_permute_gather`` / ``tm_permute_scatter`` with it moves a
        length-p vector between column order and block order for ALL blocks at once (the
        reference indexes per block with numpy, split_matrix.py:391-417, :448-460)."""
        cached = self.__dict__.get("_block_order_cache")
        if cached is None:
            perm = _dev.to_dev(np.concatenate(self.indices).astype(np.int32))
            offs = np.concatenate([[0], np.cumsum([len(i) for i in self.indices])]).astype(int)
            cached = self.__dict__["_block_order_cache"] = (perm, offs)
        return cached

    def _to_block_order(self, v_t: torch.Tensor) -> torch.Tensor:
        perm, _ = self._block_order()
        out = torch.empty_like(v_t)
        check(fn("tm_permute_gather", _dev.suffix(v_t.dtype))(
            _dev.ptr(v_t), _dev.ptr(perm), v_t.numel(), _dev.ptr(out), 0, _dev.stream_ptr()))
        return out

    def _from_block_order(self, vb_t: torch.Tensor) -> torch.Tensor:
        perm, _ = self._block_order()
        out = torch.empty_like(vb_t)
        check(fn("tm_permute_scatter", _dev.suffix(vb_t.dtype))(
            _dev.ptr(vb_t), _dev.ptr(perm), vb_t.numel(), _dev.ptr(out), 0, _dev.stream_ptr()))
        return out

    def _split_col_subsets(self, cols):
        """(positions of each block's columns in the output, block-local column ids, n_cols);
        see split_matrix.py:269-291."""
        if cols is None:
            return self.indices, [None for _ in range(len(self.indices))], self.shape[1]
        if _dev.is_dev(cols):
            cols = _dev.to_host(cols)
        cols = np.asarray(cols).astype(np.int32)
        return split_col_subsets(self.indices, cols)

    def astype(self, dtype, order="K", casting="unsafe", copy=True):
        new_matrices = [mat.astype(dtype=dtype, order=order, casting=casting, copy=True)
                        for mat in self.matrices]
        return SplitMatrix(new_matrices, self.indices)

    def toarray(self) -> np.ndarray:
        out = np.empty(self.shape)
        for mat, idx in zip(self.matrices, self.indices):
            out[:, idx] = mat.toarray()
        return out

    def getcol(self, i: int):
        i %= self.shape[1]
        for mat, idx in zip(self.matrices, self.indices):
            if i in idx:
                loc = np.where(idx == i)[0][0]
                return mat.getcol(loc)
        raise RuntimeError(f"Column {i} was not found.")

    # ---- hot path --------------------------------------------------------------------
    def sandwich(self, d, rows=None, cols=None):
        """X[rows, cols].T @ diag(d[rows]) @ X[rows, cols] as float64 (split_matrix.py:324-356)."""
        if not _dev.is_dev(d):
            d = np.asarray(d)
        check_sandwich_compatible(self, d)
        d_t, host = _vec_in(d)
        out = self._sandwich_dev(d_t, _dev.idx32(rows), cols)
        return _dev.ret(out, host)

    # ---- whole-matrix native path (cols=None): two C calls instead of a Python block loop --
    def _native_plan(self, tdtype: torch.dtype):
        """(ctypes array of tm_block_desc, workspace elements) or None when the blocks cannot go
        through tm_split_sandwich_* (mixed dtypes)."""
        key = ("plan", tdtype)
        cache = self.__dict__.setdefault("_plan_cache", {})
        if key in cache:
            return cache[key]
        plan = None
        ok = tdtype in (torch.float32, torch.float64)
        descs = (BlockDesc * len(self.matrices))()
        for b, (mat, idx_t) in enumerate(zip(self.matrices, self._dev_indices())):
            if not ok:
                break
            dsc = descs[b]
            dsc.ncols = mat.shape[1]
            dsc.col_index = idx_t.data_ptr()
            if isinstance(mat, DenseMatrix):
                X = mat._native()
                ok = X.dtype == tdtype
                dsc.kind, dsc.c_order, dsc.data = 0, int(X.is_contiguous()), X.data_ptr()
                if X.shape[1] != mat.shape[1]:
                    # zero-padded storage (DenseMatrix.__init__): the block has the stored width,
                    # its padding columns have no destination (dropped by the assembly)
                    pad_idx = _dev.to_dev(self._native_indices()[b])
                    cache[("pad_idx", b)] = pad_idx
                    dsc.ncols = X.shape[1]
                    dsc.col_index = pad_idx.data_ptr()
            elif isinstance(mat, SparseMatrix):
                c = mat._csr
                ok = c.data.dtype == tdtype
                dsc.kind, dsc.data, dsc.nnz = 1, c.data.data_ptr(), c.nnz
                dsc.csr_indices, dsc.csr_indptr = c.indices.data_ptr(), c.indptr.data_ptr()
                dsc.csr_row = c.row.data_ptr()
                block_rows = int(os.environ.get("TABMAT_B200_CSC_ROW_BLOCK", CSC_ROW_BLOCK))
                has_cat = any(isinstance(m, CategoricalMatrix) for m in self.matrices)
                if (os.environ.get("TABMAT_B200_CSC_ROW_BLOCKS", "1") != "0" and c.nnz and has_cat
                        and mat.shape[0] > block_rows + block_rows // 2):
                    # row-blocked CSC + the column-owner kernel (k_cat_sparse_cols): the per-row
                    # gathers of the categorical x sparse kernel stay inside a moving window of
                    # block_rows rows that sits in L2.  One more copy of the non-zeros, built once
                    # (TABMAT_B200_CSC_ROW_BLOCKS=0 keeps the plain CSC order).
                    bd, br, bp, nblk = mat._row_blocked_csc(block_rows)
                    dsc.csc_data, dsc.csc_indices = bd.data_ptr(), br.data_ptr()
                    dsc.csc_indptr, dsc.csc_row_blocks = bp.data_ptr(), nblk
                    csc_rows = br
                else:
                    cc = mat._csc
                    dsc.csc_data, dsc.csc_indices = cc.data.data_ptr(), cc.indices.data_ptr()
                    dsc.csc_indptr, dsc.csc_row_blocks = cc.indptr.data_ptr(), 0
                    csc_rows = cc.indices
                pk = self._packed_cat_codes(csc_rows, tdtype) if has_cat and c.nnz else None
                if pk is not None:
                    dsc.csc_cat_codes = pk.data_ptr()
                # second row-blocked copy for the gather form of dense x sparse: blocks small
                # enough that a block of the dense operand (rows x ncols x sizeof) stays in L2
                gb = self._gather_block_rows(tdtype)
                if gb and c.nnz and mat.shape[0] > gb + gb // 2:
                    gd, gr, gp, gnblk = mat._row_blocked_csc(gb)
                    dsc.gcsc_data, dsc.gcsc_indices = gd.data_ptr(), gr.data_ptr()
                    dsc.gcsc_indptr, dsc.gcsc_row_blocks = gp.data_ptr(), gnblk
            elif isinstance(mat, CategoricalMatrix):
                ok = _dev.torch_dtype(mat.dtype) == tdtype
                dsc.kind, dsc.data, dsc.drop_first = 2, mat._codes.data_ptr(), int(mat.drop_first)
                # rows stored sorted by this block's codes (row_order.py): run-aggregating kernels
                dsc.flags = int(getattr(mat, "_run_sorted", 0))  # 1 sort key, 3 primary key
                # many levels: too wide for the one-hot tensor path.  Opt-in sorted-gather kernel
                # (TABMAT_B200_GATHER=1): measured slower than the RED scatter pass on B200
                # (5.5 ms vs 3.4 ms per block at n = 4e7), kept for machines / shapes where the
                # L2 atomic units are the scarcer resource.
                if (os.environ.get("TABMAT_B200_GATHER") == "1" and mat.shape[1] > 256
                        and tdtype == torch.float32 and mat.shape[0] < 2**31):
                    perm, segptr, n_valid = mat._sorted_perm()
                    dsc.cat_perm, dsc.cat_segptr = perm.data_ptr(), segptr.data_ptr()
                    dsc.cat_nvalid = n_valid
            else:
                ok = False
        if ok:
            plan = (descs, int(lib.tm_split_workspace_elems(descs, len(self.matrices))))
        cache[key] = plan
        return plan

    def _gather_block_rows(self, tdtype) -> int:
        """Rows per block of the row-blocked CSC copy that feeds the gather form of the
        dense x sparse block (csrc/split_fused.cu: k_csc_dense_gather): the largest power of two
        with rows * (dense width) * sizeof <= 64 MB, so that a block of the dense operand stays in
        the 126 MB L2 (128 MB blocks measured 3 % faster still: the window is not critical).  0 = do not build it (no single row-major
        dense block of a supported width, or TABMAT_B200_DXS=red)."""
        # Measured at the benchmark shape (n = 4e7, 128 fp32 dense columns, ~3 non-zeros per row,
        # profiles/bench_r2m_*): gather 7.4-7.6 ms + the categorical run sums in 4 scatter warps
        # of the tcgen05 kernel (8.1 ms instead of 6.7) = 22.7 ms per step against 25.1 ms for
        # the RED form (scatter pass 11.5 ms).  Default for that case (fp32, dense width <= 128,
        # tcgen05 available); TABMAT_B200_DXS=red keeps the REDs, =gather forces the gather form
        # elsewhere (fp64 / wider dense blocks: the categorical blocks then need their own pass).
        mode = os.environ.get("TABMAT_B200_DXS", "auto")
        if mode == "red":
            return 0
        dense = [m for m in self.matrices if isinstance(m, DenseMatrix)]
        if len(dense) != 1 or dense[0]._native().dtype != tdtype or not dense[0]._native().is_contiguous():
            return 0
        fsize = 4 if tdtype == torch.float32 else 8
        q = dense[0]._native().shape[1]
        width = 16 // fsize
        if q <= 0 or q % width or q > 64 * width:
            return 0
        if mode != "gather" and not (tdtype == torch.float32 and 8 <= q <= 128
                                     and lib.tm_has_tcgen05()):
            return 0
        cap = int(os.environ.get("TABMAT_B200_GATHER_MB", "64")) * (1 << 20)
        rows = max(1024, cap // (q * fsize))
        return 1 << (rows.bit_length() - 1)

    def _packed_cat_codes(self, csc_rows: torch.Tensor, tdtype) -> Optional[torch.Tensor]:
        """For every non-zero of the sparse block's CSC copy (in ITS order) the categorical
        codes of the non-zero's row, bit-packed into one int64 (tm_block_desc.csc_cat_codes):
        built once per matrix and cached, like the reference's cached CSR
        (sparse_matrix.py:133-143).  None when the codes need more than 64 bits."""
        if os.environ.get("TABMAT_B200_CSC_PACKED") == "0":
            return None
        cats = [m for m in self.matrices if isinstance(m, CategoricalMatrix)]
        widths = [max(1, int(m.shape[1]).bit_length()) for m in cats]
        if not cats or len(cats) > 7 or sum(widths) > 64 or any(m.shape[1] <= 0 for m in cats):
            return None
        cache = self.__dict__.setdefault("_packed_codes_cache", {})
        key = csc_rows.data_ptr()
        if key not in cache:
            rows64 = csc_rows.to(torch.int64)
            pk = torch.zeros(rows64.numel(), dtype=torch.int64, device=rows64.device)
            shift = 0
            for m, w in zip(cats, widths):
                c = m._codes.index_select(0, rows64).to(torch.int64) - int(m.drop_first)
                marker = (1 << w) - 1
                c = torch.where(c < 0, torch.full_like(c, marker), c)
                if shift + w == 64:   # the top field reaches the sign bit of the int64 carrier
                    c = torch.where(c >= (1 << (w - 1)), c - (1 << w), c)
                pk |= c << shift
                shift += w
                del c
            cache[key] = (pk, csc_rows)   # keep the row array alive: its address is the key
        return cache[key][0]

    def _sandwich_blocks_dev(self, d_t: torch.Tensor, rows_t) -> Optional[torch.Tensor]:
        """Flat workspace with every self / cross block (layout: csrc/split.cu), or None."""
        plan = self._native_plan(d_t.dtype)
        if plan is None:
            return None
        descs, elems = plan
        ws = torch.empty(elems, dtype=d_t.dtype, device=d_t.device)
        check(fn("tm_split_sandwich_blocks", _dev.suffix(d_t.dtype))(
            descs, len(self.matrices), self.shape[0], _dev.ptr(d_t), _dev.ptr(rows_t),
            _dev.length(rows_t), _dev.ptr(ws), _dev.stream_ptr()))
        return ws

    def _sandwich_blocks_overlapped_dev(self, d_t: torch.Tensor, rows_t, reduce_async, reduce):
        """The flat block workspace of a ROW-SHARDED sandwich with the collective overlapped:
        the blocks without the dense operand (95 % of the payload) are computed first and
        ``reduce_async(ws_tail)`` starts their allreduce (it returns an object with ``wait()``);
        the dense-operand passes run meanwhile (the gather kernel leaves SMs free for the
        collective, ``tm_set_sm_reserve``); ``reduce(ws_head)`` then sums the small rest.  None
        when the layout does not allow the two phases (the dense block must come first)."""
        plan = self._native_plan(d_t.dtype)
        if plan is None or not isinstance(self.matrices[0], DenseMatrix) or len(self.matrices) < 2:
            return None
        descs, elems = plan
        nb = len(self.matrices)
        head = int(lib.tm_split_workspace_head_elems(descs, nb))
        ws = torch.empty(elems, dtype=d_t.dtype, device=d_t.device)
        args = (descs, nb, self.shape[0], _dev.ptr(d_t), _dev.ptr(rows_t), _dev.length(rows_t),
                _dev.ptr(ws))
        suf = _dev.suffix(d_t.dtype)
        st = _dev.stream_ptr()
        check(fn("tm_split_sandwich_blocks_part", suf)(*args, 1, st))
        work = reduce_async(ws[head:]) if elems > head else None
        check(fn("tm_split_sandwich_blocks_part", suf)(*args, 2, st))
        if work is not None:
            work.wait()
        reduce(ws[:head])
        return ws

    # ---- fused IRLS pass: Hessian and score from one pass over the dense block -------------
    def sandwich_and_transpose_matvec(self, d, v, rows=None, cols=None):
        """``(X.T diag(d) X, X.T v)`` restricted to ``rows`` / ``cols`` — what a glum-style IRLS
        step needs (Hessian with ``d`` = the working weights, score / right-hand side with
        ``v``).  The reference's callers make two calls (``sandwich``, split_matrix.py:324-356,
        and ``transpose_matvec``, :422-460), i.e. two passes over X; here the dense block — the
        bulk of the bytes — is read ONCE: its share of ``X.T v`` is accumulated by the tcgen05
        kernel's scale warps (fp32 FMAs, not TF32) while they stage the tile for the MMAs
        (``tm_split_sandwich_rmatvec_blocks``).  The sparse and categorical shares are their own
        small HBM-bound kernels.  Host in -> host out, device in -> device out."""
        if not _dev.is_dev(d):
            d = np.asarray(d)
        if not _dev.is_dev(v):
            v = np.asarray(v)
        check_sandwich_compatible(self, d)
        check_matvec_dimensions(self, v, transpose=True)
        d_t, host = _vec_in(d)
        v_t, _ = _vec_in(v)
        H, g = self._sandwich_rmatvec_dev(d_t, v_t, _dev.idx32(rows), cols)
        return _dev.ret(H, host), _dev.ret(g, host)

    def _sandwich_rmatvec_blocks_dev(self, d_t, v_t, rows_t):
        """(flat block workspace, X.T v in block order as one flat vector) or None."""
        plan = self._native_plan(d_t.dtype)
        if plan is None or v_t.dtype != d_t.dtype or v_t.dim() != 1:
            return None
        descs, elems = plan
        p = self.shape[1]
        # one buffer = [workspace | X.T v block after block]: a row-sharded caller reduces it
        # with ONE collective
        buf = torch.empty(elems + p, dtype=d_t.dtype, device=d_t.device)
        ws, vec = buf[:elems], buf[elems:]
        offs = np.concatenate([[0], np.cumsum([m.shape[1] for m in self.matrices])]).astype(int)
        dense = [b for b, m in enumerate(self.matrices) if isinstance(m, DenseMatrix)]
        dvec = vec[offs[dense[0]]:offs[dense[0] + 1]] if len(dense) == 1 else vec[:0]
        if len(dense) == 1:
            stored = self.matrices[dense[0]]._native().shape[1]
            # keep the RED target vector aligned; zero-padded storage: the kernel writes one
            # entry per STORED column
            if (dvec.data_ptr() % 16) or stored != dvec.numel():
                dvec = torch.empty(stored, dtype=d_t.dtype, device=d_t.device)
        check(fn("tm_split_sandwich_rmatvec_blocks", _dev.suffix(d_t.dtype))(
            descs, len(self.matrices), self.shape[0], _dev.ptr(d_t), _dev.ptr(v_t),
            _dev.ptr(rows_t), _dev.length(rows_t), _dev.ptr(ws),
            _dev.ptr(dvec) if len(dense) == 1 else None, _dev.stream_ptr()))
        for b, m in enumerate(self.matrices):
            part = vec[offs[b]:offs[b + 1]]
            if len(dense) == 1 and b == dense[0]:
                if dvec.data_ptr() != part.data_ptr():
                    part.copy_(dvec[:part.numel()])
                continue
            part.copy_(m.transpose_matvec(v_t, rows=rows_t))
        return buf, elems

    def _rmatvec_assemble_dev(self, vec: torch.Tensor, cols=None) -> torch.Tensor:
        """X.T v from block order into column order (optionally only ``cols``)."""
        out = self._from_block_order(vec.contiguous())
        if cols is not None:
            out = out.index_select(0, _dev.idx32(cols).to(torch.int64))
        return out

    def _sandwich_rmatvec_dev(self, d_t, v_t, rows_t, cols):
        res = None
        if cols is None or self._cols_on_native_path(cols):
            res = self._sandwich_rmatvec_blocks_dev(d_t, v_t, rows_t)
        if res is None:
            tdt = _dev.torch_dtype(np.result_type(self.dtype, _dev.np_dtype(v_t.dtype)))
            g = self.transpose_matvec(v_t.to(tdt), rows=rows_t, cols=cols)
            return self._sandwich_dev(d_t, rows_t, cols), g
        buf, elems = res
        return self._assemble_dev(buf[:elems], cols), self._rmatvec_assemble_dev(buf[elems:], cols)

    def _assemble_dev(self, ws: torch.Tensor, cols=None) -> torch.Tensor:
        """Place the flat block workspace into the float64 result.  With ``cols`` (sorted,
        unique column ids, split.pyx:157-209) only the selected rows / columns are placed: the
        blocks are computed whole by the fused passes and the selection costs nothing extra."""
        descs, _ = self._native_plan(ws.dtype)
        p = self.shape[1]
        keep = None
        if cols is not None:
            cols = np.asarray(_dev.to_host(cols) if _dev.is_dev(cols) else cols).astype(np.int64)
            dest = np.full(p, -1, dtype=np.int64)
            dest[cols] = np.arange(len(cols), dtype=np.int64)
            p = len(cols)
            sel = (BlockDesc * len(self.matrices))()
            keep = []
            dest = np.concatenate([dest, [-1]])   # index -1 (padding column) -> no destination
            for b, idx in enumerate(self._native_indices()):
                C.memmove(C.byref(sel[b]), C.byref(descs[b]), C.sizeof(BlockDesc))
                t = _dev.to_dev(dest[idx])
                keep.append(t)
                sel[b].col_index = t.data_ptr()
            descs = sel
        out = torch.empty((p, p), dtype=torch.float64, device=ws.device)
        if p:
            check(fn("tm_split_sandwich_assemble", _dev.suffix(ws.dtype))(
                descs, len(self.matrices), _dev.ptr(ws), _dev.ptr(out), p, _dev.stream_ptr()))
        del keep
        return out

    def _assemble_band_dev(self, ws: torch.Tensor, row0: int, row1: int) -> torch.Tensor:
        """Rows [row0, row1) of the float64 result from the flat block workspace."""
        descs, _ = self._native_plan(ws.dtype)
        p = self.shape[1]
        out = torch.empty((max(row1 - row0, 0), p), dtype=torch.float64, device=ws.device)
        if out.numel():
            check(fn("tm_split_sandwich_assemble_band", _dev.suffix(ws.dtype))(
                descs, len(self.matrices), _dev.ptr(ws), _dev.ptr(out), p, row0, row1,
                _dev.stream_ptr()))
        return out

    # ---- result straight into host memory, copy overlapped with the dense-operand passes ----
    def _column_runs(self):
        """[(start, stop, is_dense)]: maximal runs of consecutive result columns that belong to
        the dense block / to the other blocks."""
        runs = self.__dict__.get("_col_runs")
        if runs is None:
            p = self.shape[1]
            is_dense = np.zeros(p, dtype=bool)
            for mat, idx in zip(self.matrices, self.indices):
                if isinstance(mat, DenseMatrix):
                    is_dense[idx] = True
            cuts = np.flatnonzero(np.diff(is_dense.astype(np.int8))) + 1
            edges = [0, *cuts.tolist(), p]
            runs = [(a, b, bool(is_dense[a])) for a, b in zip(edges[:-1], edges[1:])]
            self.__dict__["_col_runs"] = runs
        return runs

    def sandwich_into(self, d, out, rows=None, reduce=None, band=None):
        """``out[:] = X[rows].T @ diag(d[rows]) @ X[rows]`` for a HOST float64 ``out`` (p x p,
        C-contiguous numpy array or CPU tensor; pinned memory gives full PCIe speed).

        The host-buffer form of :meth:`sandwich` (split_matrix.py:324-356 returns a fresh
        ndarray): ``d`` may live on the host (pinned or not) or on the device.  The blocks
        without a dense operand are computed, placed and copied to the host first, while the
        tensor-core and scatter passes of the dense block still run (two-phase C-ABI calls
        ``tm_split_sandwich_blocks_part`` / ``_assemble_part`` / ``tm_memcpy2d_to_host``).
        Asynchronous on the current CUDA stream: ``out`` is complete after
        ``torch.cuda.current_stream().synchronize()``.
        """
        if _dev.is_dev(d):
            d_t = d.contiguous()
        else:
            src = d if isinstance(d, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(d))
            d_t = torch.empty(src.shape, dtype=src.dtype, device=_dev.require_cuda())
            d_t.copy_(src, non_blocking=True)
        check_sandwich_compatible(self, d_t)
        return self._sandwich_into_dev(d_t, _dev.idx32(rows), out, reduce, band)

    def _sandwich_into_dev(self, d_t: torch.Tensor, rows_t, out, reduce=None, band=None):
        """``reduce(ws_slice) -> bool`` (row-sharded callers): sum the slice of the flat block
        workspace over the ranks in place and tell whether this rank holds the result (and so
        places and copies it).  ``band = (r0, r1)``: this rank places and copies only rows
        ``[r0, r1)`` of the result (into the same rows of ``out``, e.g. a host buffer shared by
        the ranks); needs the two-phase path."""
        p = self.shape[1]
        if isinstance(out, torch.Tensor):
            ok = (not out.is_cuda) and out.dtype == torch.float64 and out.is_contiguous()
            out_ptr = out.data_ptr()
        else:
            ok = (isinstance(out, np.ndarray) and out.dtype == np.float64
                  and out.flags["C_CONTIGUOUS"] and out.flags["WRITEABLE"])
            out_ptr = out.ctypes.data if ok else 0
        if not ok or tuple(out.shape) != (p, p):
            raise ValueError(f"out must be a C-contiguous host float64 array of shape {(p, p)}")
        st = _dev.stream_ptr()
        row_bytes = p * 8
        suf = _dev.suffix(d_t.dtype)

        b0, b1 = (0, p) if band is None else (int(band[0]), int(band[1]))

        def copy2d(buf, r0, r1, c0, c1, stream):
            """rows [r0, r1) x columns [c0, c1) of the result, clipped to this rank's band;
            ``buf`` holds the band (its row 0 = result row b0)."""
            r0, r1 = max(r0, b0), min(r1, b1)
            if r1 <= r0:
                return
            check(lib.tm_memcpy2d_to_host(out_ptr + (r0 * p + c0) * 8, row_bytes,
                                          buf.data_ptr() + ((r0 - b0) * p + c0) * 8, row_bytes,
                                          (c1 - c0) * 8, r1 - r0, stream))

        def assemble(ws_t, buf, part, stream):
            if band is None:
                check(fn("tm_split_sandwich_assemble_part", suf)(descs, nb, _dev.ptr(ws_t),
                                                                _dev.ptr(buf), p, part, stream))
            elif b1 > b0:
                check(fn("tm_split_sandwich_assemble_part_band", suf)(
                    descs, nb, _dev.ptr(ws_t), _dev.ptr(buf), p, part, b0, b1, stream))

        plan = self._native_plan(d_t.dtype)
        runs = self._column_runs()
        dense_runs = [r for r in runs if r[2]]
        other_runs = [r for r in runs if not r[2]]
        # with a collective in between, the two phases need the dense block's part of the
        # workspace to be one contiguous piece: true when the dense block comes first
        two_phase = (plan is not None and dense_runs and other_runs and len(runs) <= 6
                     and (reduce is None or isinstance(self.matrices[0], DenseMatrix)))
        if not two_phase and band is not None:
            # one phase, this rank's band only
            ws = self._sandwich_blocks_dev(d_t, rows_t)
            if ws is None:
                raise TypeError("band output needs a native block plan (one dtype for all blocks)")
            if reduce is not None and not reduce(ws):
                return None
            res = self._assemble_band_dev(ws, b0, b1)
            copy2d(res, b0, b1, 0, p, st)
            return out
        if not two_phase:
            if reduce is not None and plan is not None:
                ws = self._sandwich_blocks_dev(d_t, rows_t)
                if not reduce(ws):
                    return None
                res = self._assemble_dev(ws)
            elif reduce is not None:
                res = self._sandwich_dev(d_t, rows_t, None)
                if not reduce(res):
                    return None
            else:
                res = self._sandwich_dev(d_t, rows_t, None)
            copy2d(res, 0, p, 0, p, st)
            return out
        descs, elems = plan
        nb = len(self.matrices)
        ws = torch.empty(elems, dtype=d_t.dtype, device=d_t.device)
        args = (descs, nb, self.shape[0], _dev.ptr(d_t), _dev.ptr(rows_t), _dev.length(rows_t),
                _dev.ptr(ws))
        # elements of the dense block's self + cross blocks at the head of the workspace
        head = (int(lib.tm_split_workspace_head_elems(descs, nb))
                if isinstance(self.matrices[0], DenseMatrix) else 0)
        # phase 1: everything without the dense operand, then its host copy on a second stream
        check(fn("tm_split_sandwich_blocks_part", suf)(*args, 1, st))
        mine = True if reduce is None else reduce(ws[head:])
        buf = None
        cs = None
        if mine:
            buf = torch.empty((b1 - b0, p), dtype=torch.float64, device=d_t.device)
            assemble(ws, buf, 1, st)
            cs = self.__dict__.get("_copy_stream")
            if cs is None:
                cs = self.__dict__["_copy_stream"] = torch.cuda.Stream()
            cs.wait_stream(torch.cuda.current_stream())
            for (r0, r1, _) in other_runs:
                for (c0, c1, _) in other_runs:
                    copy2d(buf, r0, r1, c0, c1, cs.cuda_stream)
        # phase 2: the dense block's own and cross blocks
        check(fn("tm_split_sandwich_blocks_part", suf)(*args, 2, st))
        if reduce is not None:
            reduce(ws[:head])
        if not mine:
            return None
        assemble(ws, buf, 2, st)
        for (r0, r1, _) in dense_runs:
            copy2d(buf, r0, r1, 0, p, st)
        for (r0, r1, _) in other_runs:
            for (c0, c1, _) in dense_runs:
                copy2d(buf, r0, r1, c0, c1, st)
        torch.cuda.current_stream().wait_stream(cs)
        buf.record_stream(cs)
        return out

    def _sandwich_dev(self, d_t: torch.Tensor, rows_t, cols) -> torch.Tensor:
        if cols is None or self._cols_on_native_path(cols):
            ws = self._sandwich_blocks_dev(d_t, rows_t)
            if ws is not None:
                return self._assemble_dev(ws, cols)
        subset_cols_indices, subset_cols, n_cols = self._split_col_subsets(cols)
        if cols is None:
            pos_t = self._dev_indices()
            sub_t = [None] * len(self.matrices)
        else:
            pos_t = [_dev.to_dev(np.asarray(p, dtype=np.int64)) for p in subset_cols_indices]
            sub_t = [_dev.idx32(s) for s in subset_cols]
        st = _dev.stream_ptr()
        # every (i, j) entry is written by exactly one block (or its mirror): no zero-fill
        out = torch.empty((n_cols, n_cols), dtype=torch.float64, device=d_t.device)
        k = len(self.matrices)
        fused = self._fused_dense_cross(d_t, rows_t) if cols is None else {}
        for i in range(k):
            mat_i = self.matrices[i]
            mi = int(pos_t[i].numel())
            if mi == 0:
                continue
            if isinstance(mat_i, CategoricalMatrix):
                diag, _ = mat_i._sandwich_diag(d_t, rows_t, sub_t[i])
                check(fn("tm_scatter_diag", _dev.suffix(diag.dtype))(
                    _dev.ptr(diag), mi, _dev.ptr(pos_t[i]), _dev.ptr(out), n_cols, st))
            else:
                res = mat_i.sandwich(d_t, rows_t, sub_t[i])
                check(fn("tm_scatter_block", _dev.suffix(res.dtype))(
                    _dev.ptr(res), mi, mi, _dev.ptr(pos_t[i]), _dev.ptr(pos_t[i]), _dev.ptr(out),
                    n_cols, 0, st))
            for j in range(i + 1, k):
                mj = int(pos_t[j].numel())
                if mj == 0:
                    continue
                if (j, i) in fused:  # computed as (block j rows) x (block i cols)
                    res = fused[(j, i)]
                    check(fn("tm_scatter_block", _dev.suffix(res.dtype))(
                        _dev.ptr(res), mj, mi, _dev.ptr(pos_t[j]), _dev.ptr(pos_t[i]),
                        _dev.ptr(out), n_cols, 1, st))
                    continue
                if (i, j) in fused:
                    res = fused[(i, j)]
                else:
                    res = self.matrices[i]._cross_sandwich(self.matrices[j], d_t, rows_t,
                                                           sub_t[i], sub_t[j])
                    res = res.contiguous()
                check(fn("tm_scatter_block", _dev.suffix(res.dtype))(
                    _dev.ptr(res), mi, mj, _dev.ptr(pos_t[i]), _dev.ptr(pos_t[j]), _dev.ptr(out),
                    n_cols, 1, st))
        return out

    def _cols_on_native_path(self, cols) -> bool:
        """An active-set ``cols`` (glum's normal call) goes through the fused whole-matrix
        passes + a selecting assembly when it keeps a fair share of the columns; a narrow
        selection is cheaper block by block with the restricted per-pair kernels."""
        if os.environ.get("TABMAT_B200_COLS_NATIVE") == "0":
            return False
        m = len(cols)
        return m * 8 >= self.shape[1] or self.shape[0] >= 1_000_000

    def _fused_dense_cross(self, d_t: torch.Tensor, rows_t) -> dict:
        """{(a, b): block a rows x dense-block b cols} for every categorical / sparse block a,
        from ONE pass over the dense block (ext.split.dense_cross_sandwich), when the layout
        allows it; {} otherwise (the per-pair kernels are used then)."""
        dense = [i for i, m in enumerate(self.matrices) if isinstance(m, DenseMatrix)]
        if len(dense) != 1:
            return {}
        b = dense[0]
        X = self.matrices[b]._native()
        width = 4 if X.dtype == torch.float32 else 2
        p = X.shape[1]
        p_log = self.matrices[b].shape[1]   # < p with zero-padded storage
        if (X.dtype != d_t.dtype or not X.is_contiguous() or p % width or p > 64 * width
                or X.data_ptr() % 16):
            return {}
        cats = [(i, m) for i, m in enumerate(self.matrices)
                if isinstance(m, CategoricalMatrix) and m.shape[1] > 0
                and _dev.torch_dtype(m.dtype) == X.dtype]
        sparse = [(i, m) for i, m in enumerate(self.matrices)
                  if isinstance(m, SparseMatrix) and m._csr.data.dtype == X.dtype and m._csr.nnz]
        if len(cats) > 8 or len(sparse) > 1 or not (cats or sparse):
            return {}
        outs, out_s = dense_cross_sandwich(
            X, d_t, rows_t, [(m._codes, m.shape[1], m.drop_first) for _, m in cats],
            sparse[0][1]._csr if sparse else None)
        if p_log != p:
            outs = [o[:, :p_log].contiguous() for o in outs]
            out_s = out_s[:, :p_log].contiguous() if out_s is not None else None
        fused = {(i, b): o for (i, _), o in zip(cats, outs)}
        if sparse and out_s is not None:
            fused[(sparse[0][0], b)] = out_s
        return fused

    def _get_col_means(self, weights):
        host = not _dev.is_dev(weights)
        parts = [mat._get_col_means(weights) for mat in self.matrices]
        return self._gather_cols(parts, host)

    def _get_col_stds(self, weights, col_means):
        host = not _dev.is_dev(weights)
        parts = []
        for idx, mat in zip(self.indices, self.matrices):
            cm = col_means[idx] if host else col_means[_dev.to_dev(idx)]
            parts.append(mat._get_col_stds(weights, cm))
        return self._gather_cols(parts, host)

    def _gather_cols(self, parts, host: bool):
        if host:
            out = np.empty(self.shape[1], dtype=self.dtype)
            for idx, part in zip(self.indices, parts):
                out[idx] = part
            return out
        out = torch.empty(self.shape[1], dtype=_dev.torch_dtype(self.dtype), device=parts[0].device)
        for idx_t, part in zip(self._dev_indices(), parts):
            out[idx_t] = part.to(out.dtype)
        return out

    def matvec(self, v, cols=None, out=None):
        """self[:, cols] @ v[cols] (split_matrix.py:373-420)."""
        assert not isinstance(v, sps.spmatrix)
        if not _dev.is_dev(v):
            v = np.asarray(v)
        check_matvec_dimensions(self, v, transpose=False)
        check_matvec_out_shape(self, out)
        v_t, host = _vec_in(v)
        if v_t.dim() > 1 and any(isinstance(m, CategoricalMatrix) for m in self.matrices):
            # the reference calls every block's matvec, and the categorical one is 1-d only
            # (categorical_matrix.py:478-481)
            raise NotImplementedError(
                """CategoricalMatrix.matvec is only implemented for 1d arrays."""
            )
        _, subset_cols, n_cols = self._split_col_subsets(cols)
        out_dtype = np.result_type(self.dtype, _dev.np_dtype(v_t.dtype)
                                   if v_t.dtype in (torch.float32, torch.float64) else np.float64)
        tdt = _dev.torch_dtype(out_dtype)
        if v_t.dtype != tdt:
            v_t = v_t.to(tdt)
        out_shape = [self.shape[0]] + list(v_t.shape[1:])
        if out is not None and _dev.is_dev(out):
            if out.dtype != tdt:
                raise ValueError(
                    f"out array is required to have dtype {out_dtype} but hasdtype {out.dtype}")
            acc = out
        else:
            acc = torch.zeros(out_shape, dtype=tdt, device=v_t.device)
        sub_t = [None if s is None else _dev.idx32(s) for s in subset_cols]
        # v in block order for all blocks with one gather kernel (1-d float vectors)
        one_gather = v_t.dim() == 1 and v_t.dtype in (torch.float32, torch.float64)
        if one_gather:
            vb, offs = self._to_block_order(v_t.contiguous()), self._block_order()[1]
        for b, (sub, idx_t, mat) in enumerate(zip(sub_t, self._dev_indices(), self.matrices)):
            if sub is not None and sub.numel() == 0:
                continue
            in_vec = vb[offs[b]:offs[b + 1]] if one_gather else v_t.index_select(0, idx_t)
            if in_vec.dtype != _dev.torch_dtype(mat.dtype) and not isinstance(mat, CategoricalMatrix):
                # mixed-dtype split: compute the block in its own dtype, add into acc
                acc += mat.matvec(in_vec, sub).to(tdt)
            elif acc.dim() == 1:
                mat.matvec(in_vec, sub, out=acc)
            else:
                acc += mat.matvec(in_vec, sub).to(tdt)
        if out is not None and not _dev.is_dev(out):
            return _accumulate_out(out, acc, None)
        if out is not None:
            return out
        return _dev.ret(acc, host)

    def transpose_matvec(self, v, rows=None, cols=None, out=None):
        """self[rows, cols].T @ v[rows] (split_matrix.py:422-460)."""
        if not _dev.is_dev(v):
            v = np.asarray(v)
        check_matvec_dimensions(self, v, transpose=True)
        check_transpose_matvec_out_shape(self, out)
        v_t, host = _vec_in(v)
        subset_cols_indices, subset_cols, n_cols = self._split_col_subsets(cols)
        out_dtype = np.result_type(self.dtype, _dev.np_dtype(v_t.dtype)
                                   if v_t.dtype in (torch.float32, torch.float64) else np.float64)
        tdt = _dev.torch_dtype(out_dtype)
        if v_t.dtype != tdt:
            v_t = v_t.to(tdt)
        rows_t = _dev.idx32(rows)
        if cols is None and v_t.dim() == 1:
            # every block writes its slice of a block-ordered vector; ONE scatter kernel puts it
            # into column order
            offs = self._block_order()[1]
            parts = [mat.transpose_matvec(v_t, rows=rows_t).to(tdt) if mat.shape[1] else
                     torch.empty(0, dtype=tdt, device=v_t.device) for mat in self.matrices]
            res = self._from_block_order(torch.cat(parts)) if len(parts) > 1 else \
                self._from_block_order(parts[0].contiguous())
            assert res.numel() == offs[-1]
            if out is None:
                return _dev.ret(res, host)
            return _accumulate_out(out, res, None)
        res = torch.zeros([n_cols] + list(v_t.shape[1:]), dtype=tdt, device=v_t.device)
        if cols is None:
            pos_t = self._dev_indices()
            sub_t = [None] * len(self.matrices)
        else:
            pos_t = [_dev.to_dev(np.asarray(p, dtype=np.int64)) for p in subset_cols_indices]
            sub_t = [_dev.idx32(s) for s in subset_cols]
        for pos, sub, mat in zip(pos_t, sub_t, self.matrices):
            if pos.numel() == 0:
                continue
            part = mat.transpose_matvec(v_t, rows=rows_t, cols=sub)
            res.index_add_(0, pos, part.to(tdt))
        if out is None:
            return _dev.ret(res, host)
        sel = None if cols is None else _dev.idx32(cols)
        return _accumulate_out(out, res, sel)

    def __getitem__(self, key):
        if isinstance(key, tuple):
            row, col = key
        else:
            row = key
            col = slice(None, None, None)
        if col == slice(None, None, None):
            if isinstance(row, int):
                row = [row]
            if not _dev.is_dev(row) and not isinstance(row, slice):
                # one host -> device copy of the index for all blocks
                a = np.asarray(row).reshape(-1)
                a = np.flatnonzero(a) if a.dtype == bool else a.astype(np.int64)
                row = torch.from_numpy(a).to(_dev.require_cuda())
                row = torch.where(row < 0, row + self.shape[0], row)
            return SplitMatrix([mat[row, :] for mat in self.matrices], self.indices)
        raise NotImplementedError(f"Only row indexing is supported. Index passed was {key}.")

    def multiply(self, other):
        return SplitMatrix([mat.multiply(other) for mat in self.matrices], indices=self.indices)

    def __repr__(self):
        out = "SplitMatrix:"
        for i, mat in enumerate(self.matrices):
            out += f"\n\nComponent {i} with type {mat.__class__.__name__}\n" + mat.__repr__()
        return out

    __array_priority__ = 13

    def get_names(self, type: str = "column", missing_prefix: Optional[str] = None,
                  indices: Optional[list] = None) -> list:
        names = np.empty(self.shape[1], dtype=object)
        for idx, mat in zip(self.indices, self.matrices):
            names[idx] = mat.get_names(type, missing_prefix, idx)
        return names.tolist()

    def set_names(self, names: Union[str, list], type: str = "column"):
        names_array = np.array(names, dtype=object)
        if len(names) != self.shape[1]:
            raise ValueError(f"Length of names must be {self.shape[1]}")
        for idx, mat in zip(self.indices, self.matrices):
            mat.set_names(names_array[idx].tolist(), type)
