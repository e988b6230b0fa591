"""Boundary entries that round 1 left on the host: the device-side CSR form of a categorical
block (multiply_complex / subset_categorical_complex, categorical.pyx:221-315) behind
``CategoricalMatrix.multiply / getcol / tocsr / to_sparse_matrix``, int64-indexed sparse input
(sparse.pyx:13-15 ``win_integral``), and the one-kernel block gather / scatter of
SplitMatrix.matvec / transpose_matvec."""

import numpy as np
import pytest
import scipy.sparse as sps

from tests import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("drop_first,missing", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_categorical_csr_forms_on_device(drop_first, missing, dt):
    import torch

    import tabmat_b200 as tm

    n, K = 5003, 37
    rng = np.random.default_rng(int(drop_first) + 2 * int(missing))
    codes = rng.integers(0, K, size=n).astype(np.int32)
    if missing:
        codes[rng.random(n) < 0.1] = -1
    C = tm.CategoricalMatrix(codes, categories=np.arange(K), dtype=dt, drop_first=drop_first,
                             cat_missing_method="zero" if missing else "fail")
    col = codes - int(drop_first)
    dense = np.where((col >= 0)[:, None], np.eye(K - int(drop_first))[np.maximum(col, 0)], 0.0)
    # tocsr / toarray / to_sparse_matrix
    R = C.tocsr()
    assert R.shape == dense.shape and R.has_canonical_format
    np.testing.assert_array_equal(R.toarray(), dense)
    np.testing.assert_array_equal(C.to_sparse_matrix().toarray(), dense)
    # multiply: host and device vectors, result is a SparseMatrix with the scaled rows
    w = rng.standard_normal(n).astype(dt)
    for other in (w, w[:, None], torch.from_numpy(w).cuda()):
        M = C.multiply(other)
        assert isinstance(M, tm.SparseMatrix) and M.shape == dense.shape
        np.testing.assert_allclose(M.toarray(), dense * w[:, None], rtol=1e-6)
    with pytest.raises(ValueError, match="Shapes do not match"):
        C.multiply(w[:-1])
    # getcol: wrap-around, (n x 1) SparseMatrix of ones
    for i in (0, 5, -1):
        g = C.getcol(i)
        assert isinstance(g, tm.SparseMatrix) and g.shape == (n, 1)
        np.testing.assert_array_equal(g.toarray()[:, 0], dense[:, i % dense.shape[1]])
    # the raw C-ABI entry: structure only (data = NULL) and with weights
    from tabmat_b200._lib import check, lib

    cd = torch.from_numpy(codes).cuda()
    indptr = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    indices = torch.empty(n, dtype=torch.int32, device="cuda")
    check(lib.tm_cat_to_csr_f64(cd.data_ptr(), n, int(drop_first), None, None, indices.data_ptr(),
                                indptr.data_ptr(), torch.cuda.current_stream().cuda_stream))
    keep = col >= 0
    np.testing.assert_array_equal(indptr.cpu().numpy(), np.concatenate([[0], np.cumsum(keep)]))
    np.testing.assert_array_equal(indices.cpu().numpy()[:keep.sum()], col[keep])


def test_int64_indexed_sparse_input_is_narrowed():
    """scipy matrices with int64 index arrays (the reference dispatches on `win_integral`) are
    accepted: the per-shard device arrays are int32 (scipy itself narrows index arrays that fit);
    a shard with >= 2^31 rows / columns / non-zeros raises ValueError at construction."""
    import tabmat_b200 as tm

    rng = np.random.default_rng(0)
    A = sps.random(300, 17, density=0.1, format="csc", random_state=rng)
    A64 = sps.csc_matrix((A.data, A.indices.astype(np.int64), A.indptr.astype(np.int64)), shape=A.shape)
    S32, S64 = tm.SparseMatrix(A), tm.SparseMatrix(A64)
    assert S64._csr.indices.dtype == S64._csc.indices.dtype  # int32 on the device either way
    d = rng.random(300)
    rows = np.arange(0, 300, 2, dtype=np.int64)
    cols = np.arange(0, 17, 3, dtype=np.int64)
    np.testing.assert_allclose(S64.sandwich(d, rows, cols), S32.sandwich(d, rows, cols), rtol=1e-12)
    np.testing.assert_allclose(S64.matvec(np.ones(17)), A @ np.ones(17), rtol=1e-12)
    np.testing.assert_allclose(S64.transpose_matvec(d, rows, cols), (A.T @ np.where(np.isin(np.arange(300), rows), d, 0))[cols],
                               rtol=1e-12)


def test_split_matvec_uses_block_order_gather(monkeypatch):
    """matvec / transpose_matvec of a SplitMatrix with interleaved column indices: one gather /
    scatter kernel moves the vector between column order and block order."""
    import tabmat_b200 as tm

    rng = np.random.default_rng(4)
    n = 2000
    X = rng.standard_normal((n, 6))
    A = sps.random(n, 5, density=0.2, format="csc", random_state=rng)
    c = rng.integers(0, 4, size=n).astype(np.int32)
    idx = [np.array([0, 3, 6, 9, 12, 14]), np.array([1, 4, 7, 10, 13]), np.array([2, 5, 8, 11])]
    S = tm.SplitMatrix([tm.DenseMatrix(X), tm.SparseMatrix(A),
                        tm.CategoricalMatrix(c, categories=np.arange(4))], idx)
    full = np.empty((n, 15))
    full[:, idx[0]], full[:, idx[1]], full[:, idx[2]] = X, A.toarray(), np.eye(4)[c]
    v, w = rng.standard_normal(15), rng.standard_normal(n)
    tm.reset_launch_count()
    cases.assert_close(S.matvec(v), full @ v, np.float64, "interleaved matvec")
    cases.assert_close(S.transpose_matvec(w), full.T @ w, np.float64, "interleaved transpose_matvec")
    out = np.ones(15)
    S.transpose_matvec(w, out=out)
    cases.assert_close(out, 1 + full.T @ w, np.float64, "transpose_matvec accumulates into out")
    rows = np.arange(0, n, 3)
    cases.assert_close(S.transpose_matvec(w, rows=rows), full[rows].T @ w[rows], np.float64, "rows")


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("m", [7, 220])
def test_sparse_self_sandwich_shared_memory_tables(dt, m, monkeypatch):
    """The narrow-block form of the sparse self sandwich (k_sparse_sandwich_smem: packed lower
    triangle in shared memory, one table per SM) == the L2 RED form == float64 recomputation
    (sparse.pyx:17-78), with `rows` / `cols` restrictions, empty rows, zero weights."""
    import scipy.sparse as sps

    import tabmat_b200 as tm
    from tests import cases

    rng = np.random.default_rng(m)
    n = 30_011
    A = sps.random(n, m, density=min(0.9, 6.0 / m), random_state=rng, format="csc").astype(dt)
    d = rng.standard_normal(n).astype(dt)
    d[rng.random(n) < 0.1] = 0
    rows = np.sort(rng.choice(n, size=n // 3, replace=False)).astype(np.int32)
    cols = np.sort(rng.choice(m, size=max(2, m // 2), replace=False)).astype(np.int32)
    S = tm.SparseMatrix(A)
    Ad = A.toarray().astype(np.float64)
    for r, c in ((None, None), (rows, None), (None, cols), (rows, cols)):
        Ar = Ad if r is None else Ad[r]
        dr = d.astype(np.float64) if r is None else d.astype(np.float64)[r]
        Ar = Ar if c is None else Ar[:, c]
        ref = Ar.T @ (dr[:, None] * Ar)
        monkeypatch.setenv("TABMAT_B200_SPARSE_SMEM", "2")
        got = S.sandwich(d, r, c)
        monkeypatch.setenv("TABMAT_B200_SPARSE_SMEM", "0")
        red = S.sandwich(d, r, c)
        # packed-triangle accumulation (the form for squares larger than the L2, C4)
        monkeypatch.setenv("TABMAT_B200_SPARSE_TRI", "2")
        tri = S.sandwich(d, r, c)
        monkeypatch.delenv("TABMAT_B200_SPARSE_TRI")
        cases.assert_close(got, ref, dt, "shared-memory tables")
        cases.assert_close(red, ref, dt, "L2 RED form")
        cases.assert_close(tri, ref, dt, "packed triangle")
        assert np.array_equal(got, got.T) and np.array_equal(tri, tri.T)
