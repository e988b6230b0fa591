"""float32 dense blocks whose width is not a multiple of 4 are STORED zero-padded to the next
multiple (DenseMatrix._store), so that matrices built the ordinary way (pandas frames with 7
numeric columns ...) reach the TMA / tcgen05 kernels and the fused SplitMatrix passes instead of
falling back to the CUDA-core kernels.  Everything observable must be unchanged: shapes, every
MatrixBase method with and without restrictions (dense_matrix.py:153-257, split_matrix.py:324-460),
against float64 recomputation and against the unpadded storage (TABMAT_B200_DENSE_PAD=0)."""

import numpy as np
import pytest
import scipy.sparse as sps

from tests import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("p", [1, 5, 130, 255])
def test_dense_matrix_surface_with_padded_storage(p, order, monkeypatch):
    import tabmat_b200 as tm

    rng = np.random.default_rng(p)
    n = 3001
    Xh = rng.standard_normal((n, p)).astype(np.float32)
    if order == "F":
        Xh = np.asfortranarray(Xh)
    X = tm.DenseMatrix(Xh)
    assert X._store is not None and X._store.shape == (n, (p + 3) // 4 * 4)
    assert X.shape == (n, p) and X.toarray().shape == (n, p)
    np.testing.assert_array_equal(X.toarray(), Xh)
    monkeypatch.setenv("TABMAT_B200_DENSE_PAD", "0")
    Y = tm.DenseMatrix(Xh)
    monkeypatch.delenv("TABMAT_B200_DENSE_PAD")
    assert Y._store is None
    d = rng.standard_normal(n).astype(np.float32)
    v = rng.standard_normal(p).astype(np.float32)
    w = rng.standard_normal(n).astype(np.float32)
    rows = np.sort(rng.choice(n, size=n // 2, replace=False)).astype(np.int32)
    cols = np.sort(rng.choice(p, size=max(1, p // 2), replace=False)).astype(np.int32)
    X64 = Xh.astype(np.float64)
    for r in (None, rows):
        for c in (None, cols):
            Xs = X64 if r is None else X64[r]
            Xs = Xs if c is None else Xs[:, c]
            dr = d.astype(np.float64) if r is None else d.astype(np.float64)[r]
            ref = Xs.T @ (dr[:, None] * Xs)
            got = X.sandwich(d, r, c)
            assert got.shape == ref.shape
            cases.assert_close(got, ref, np.float32, f"sandwich p={p}")
            cases.assert_close(Y.sandwich(d, r, c), ref, np.float32, "unpadded")
            wr = w.astype(np.float64) if r is None else w.astype(np.float64)[r]
            cases.assert_close(X.transpose_matvec(w, r, c), Xs.T @ wr, np.float32, "rmatvec")
        for c in (None, cols):
            vv = v.astype(np.float64)
            ref = X64 @ vv if c is None else X64[:, c] @ vv[c]
            cases.assert_close(X.matvec(v, c), ref, np.float32, "matvec")
    V2 = rng.standard_normal((p, 3)).astype(np.float32)
    cases.assert_close(X.matvec(V2), X64 @ V2.astype(np.float64), np.float32, "matvec 2-d")
    W2 = rng.standard_normal((n, 2)).astype(np.float32)
    cases.assert_close(X.transpose_matvec(W2), X64.T @ W2.astype(np.float64), np.float32, "rmatvec 2-d")
    out = np.zeros(n, dtype=np.float32)
    X.matvec(v, out=out)
    cases.assert_close(out, X64 @ v.astype(np.float64), np.float32, "matvec out=")
    wt = np.abs(w) / np.abs(w).sum()
    mu = X._get_col_means(wt)
    cases.assert_close(mu, wt.astype(np.float64) @ X64, np.float32, "col means")
    cases.assert_close(X._get_col_stds(wt, mu),
                       np.sqrt(wt.astype(np.float64) @ (X64 - mu.astype(np.float64)) ** 2),
                       np.float32, "col stds")
    # array-like surface
    np.testing.assert_array_equal(X.getcol(p - 1).toarray(), Xh[:, [p - 1]])
    np.testing.assert_array_equal(X[rows, :].toarray(), Xh[rows])
    np.testing.assert_array_equal(X[:, cols].toarray(), Xh[:, cols])
    np.testing.assert_array_equal(X.T.toarray(), Xh.T)
    np.testing.assert_array_equal(X.astype(np.float64).toarray(), X64)
    np.testing.assert_allclose(X.multiply(w).toarray(), Xh * w[:, None], rtol=1e-6)
    Z, means, stds = X.standardize(wt, True, True)
    Xz = (X64 - means.astype(np.float64)) / stds.astype(np.float64)
    cases.assert_close(Z.sandwich(d), Xz.T @ (d.astype(np.float64)[:, None] * Xz), np.float32,
                       "standardized sandwich")


@pytest.mark.parametrize("pd", [7, 126])
def test_split_matrix_with_padded_dense_block(pd, monkeypatch):
    import tabmat_b200 as tm

    lib = tm._lib.lib
    rng = np.random.default_rng(pd)
    n = 6007
    Xd = rng.standard_normal((n, pd)).astype(np.float32)
    A = sps.random(n, 40, density=0.05, random_state=rng, format="csc").astype(np.float32)
    levels = (10, 300, 1200)
    codes = [rng.integers(0, K, size=n).astype(np.int32) for K in levels]

    def build():
        return tm.SplitMatrix([tm.DenseMatrix(Xd), tm.SparseMatrix(A)] +
                              [tm.CategoricalMatrix(c, categories=np.arange(K), dtype=np.float32)
                               for c, K in zip(codes, levels)])

    S = build()
    assert S.matrices[0]._store is not None and S.shape[1] == pd + 40 + sum(levels)
    full = np.hstack([Xd.astype(np.float64), A.toarray().astype(np.float64)] +
                     [np.eye(K)[c] for c, K in zip(codes, levels)])
    d = rng.standard_normal(n).astype(np.float32)
    v = rng.standard_normal(n).astype(np.float32)
    beta = rng.standard_normal(S.shape[1]).astype(np.float32)
    rows = np.sort(rng.choice(n, size=n // 3, replace=False)).astype(np.int32)
    cols = np.sort(rng.choice(S.shape[1], size=S.shape[1] // 2, replace=False)).astype(np.int32)
    narrow = np.sort(rng.choice(S.shape[1], size=20, replace=False)).astype(np.int32)
    R = tm.RowSortedMatrix.from_split(S)
    for M, what in ((S, "split"), (R, "row-sorted")):
        for r in (None, rows):
            for c in (None, cols, narrow):
                F = full if r is None else full[r]
                F = F if c is None else F[:, c]
                dd = d.astype(np.float64) if r is None else d.astype(np.float64)[r]
                ref = (F * dd[:, None]).T @ F
                got = M.sandwich(d, r, c)
                assert got.shape == ref.shape
                cases.assert_close(got, ref, np.float32, f"{what} sandwich")
            F = full if r is None else full[r]
            vv = v.astype(np.float64) if r is None else v.astype(np.float64)[r]
            cases.assert_close(M.transpose_matvec(v, r), F.T @ vv, np.float32, f"{what} rmatvec")
            H, g = M.sandwich_and_transpose_matvec(d, v, r)
            dd = d.astype(np.float64) if r is None else d.astype(np.float64)[r]
            cases.assert_close(H, (F * dd[:, None]).T @ F, np.float32, f"{what} fused H")
            cases.assert_close(g, F.T @ vv, np.float32, f"{what} fused g")
        cases.assert_close(M.matvec(beta), full @ beta.astype(np.float64), np.float32, f"{what} matvec")
        cases.assert_close(M.matvec(beta, cols), full[:, cols] @ beta.astype(np.float64)[cols],
                           np.float32, f"{what} matvec cols")
        out = np.zeros((S.shape[1], S.shape[1]))
        M.sandwich_into(d, out)
        import torch

        torch.cuda.synchronize()
        cases.assert_close(out, (full * d.astype(np.float64)[:, None]).T @ full, np.float32,
                           f"{what} sandwich_into")
    # the padded block takes the fused native passes (tcgen05 when the hardware has it)
    S.sandwich(d)
    assert S._native_plan(tm._dev.torch_dtype(np.float32)) is not None
    if lib.tm_has_tcgen05():
        assert lib.tm_split_last_plan() & 1
    monkeypatch.setenv("TABMAT_B200_DENSE_PAD", "0")
    U = build()
    assert U.matrices[0]._store is None
    cases.assert_close(U.sandwich(d), (full * d.astype(np.float64)[:, None]).T @ full, np.float32,
                       "unpadded split")
