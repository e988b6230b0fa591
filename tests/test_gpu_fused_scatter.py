"""The fused form of the SplitMatrix sandwich (csrc/dense_tc.cu, scatter warps): the vector REDs
of dense x sparse and dense x many-level categoricals are issued from the TMA-staged tile of the
tcgen05 kernel, so the dense block is read once.  Checked against dense float64 recomputation
(the reference's test strategy, tests/test_split_matrix.py:170-288) and against the separate
scatter pass (tm_set_tc_scatter_warps(0)) for 4 and 8 scatter warps, ragged row counts, rows
with > 32 non-zeros per warp slice, few-level tables that are replicated, row restrictions,
drop_first / missing codes, both row orders."""

import numpy as np
import pytest
import scipy.sparse as sps

from tests import cases

pytestmark = pytest.mark.gpu


def _build(n, p_dense, p_sparse, density, levels, seed):
    import tabmat_b200 as tm

    dt = np.float32
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, p_dense)).astype(dt)
    mats, cols = [tm.DenseMatrix(X)], [X.astype(np.float64)]
    if p_sparse:
        A = sps.random(n, p_sparse, density=density, random_state=rng, format="csc").astype(dt)
        mats.append(tm.SparseMatrix(A))
        cols.append(A.toarray().astype(np.float64))
    for q, K in enumerate(levels):
        c = rng.integers(0, K, size=n).astype(np.int32)
        df = q % 3 == 1
        miss = q % 3 == 2
        if miss:
            c[rng.random(n) < 0.1] = -1
        mats.append(tm.CategoricalMatrix(c, categories=np.arange(K), dtype=dt, drop_first=df,
                                         cat_missing_method="zero" if miss else "fail"))
        oh = np.where((c >= 0)[:, None], np.eye(K)[np.maximum(c, 0)], 0.0)
        cols.append(oh[:, 1:] if df else oh)
    d = rng.standard_normal(n).astype(dt)
    d[rng.random(n) < 0.05] = 0
    return tm.SplitMatrix(mats), np.hstack(cols), d, rng


CASES = [
    # n, p_dense, p_sparse, density, levels
    (6007, 128, 300, 0.01, (10, 50, 200, 1000, 2000)),   # the benchmark's block structure
    (4099, 64, 50, 0.5, (12, 300)),                       # ~25 nnz per row: > 32 per warp slice
    (33, 8, 7, 0.3, (400,)),                              # two row tiles, one of them ragged
    (5000, 128, 0, 0.0, (300, 700, 260, 1500)),           # four scatter categoricals, no sparse
    (5000, 32, 40, 0.05, ()),                             # sparse only
    (5000, 128, 40, 0.05, (200, 200)),                    # 2nd block overflows the one-hot slots:
                                                          # few-level table, replicated
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"n{c[0]}p{c[1]}s{c[2]}c{len(c[4])}")
@pytest.mark.parametrize("scw", [4, 8])
def test_fused_scatter_matches_dense_recomputation(case, scw):
    import tabmat_b200 as tm

    lib = tm._lib.lib
    if not lib.tm_has_tcgen05():
        pytest.skip("needs sm_100")
    n, pd, ps, dens, levels = case
    X, full, d, rng = _build(n, pd, ps, dens, levels, seed=n + pd + scw)
    rows = np.sort(rng.choice(n, size=max(1, n // 3), replace=False)).astype(np.int32)
    S = tm.RowSortedMatrix.from_split(X)
    try:
        for r in (None, rows):
            F = full if r is None else full[r]
            dd = d.astype(np.float64) if r is None else d.astype(np.float64)[r]
            ref = (F * dd[:, None]).T @ F
            lib.tm_set_tc_scatter_warps(0)
            sep = X.sandwich(d, r)
            lib.tm_set_tc_scatter_warps(scw)
            got = X.sandwich(d, r)
            got_sorted = S.sandwich(d, r)
            cases.assert_close(sep, ref, np.float32, "separate scatter pass")
            cases.assert_close(got, ref, np.float32, f"fused scatter, {scw} warps")
            cases.assert_close(got_sorted, ref, np.float32, f"fused scatter, sorted rows, {scw} warps")
            # the two forms add the same terms (in a different order): they agree far inside
            # the TF32 tolerance on the blocks the scatter work produces
            err = np.abs(got - sep).max() / np.abs(ref).max()
            assert err < 1e-4, err
    finally:
        lib.tm_set_tc_scatter_warps(-1)


def test_fused_scatter_counts_are_exact():
    """d == 1 and X == 1: the dense x categorical / dense x sparse blocks are exact counts."""
    import tabmat_b200 as tm

    lib = tm._lib.lib
    if not lib.tm_has_tcgen05():
        pytest.skip("needs sm_100")
    n = 40_000
    rng = np.random.default_rng(5)
    X = np.ones((n, 16), dtype=np.float32)
    c = rng.integers(0, 700, size=n).astype(np.int32)
    A = sps.random(n, 90, density=0.03, random_state=rng, format="csc", dtype=np.float32)
    A.data[:] = 1.0
    S = tm.SplitMatrix([tm.DenseMatrix(X), tm.SparseMatrix(A),
                        tm.CategoricalMatrix(c, categories=np.arange(700), dtype=np.float32)])
    out = S.sandwich(np.ones(n, dtype=np.float32))
    counts = np.bincount(c, minlength=700).astype(np.float64)
    nnz_col = np.asarray((A != 0).sum(axis=0)).ravel().astype(np.float64)
    assert np.array_equal(out[16 + 90:, :16], np.repeat(counts[:, None], 16, axis=1))
    assert np.array_equal(out[16:16 + 90, :16], np.repeat(nnz_col[:, None], 16, axis=1))


@pytest.mark.parametrize("frac", [0.5, 0.03])
def test_cols_selection_on_the_native_path(frac, monkeypatch):
    """``cols`` (glum's active set): the fused whole-matrix passes + a selecting assembly give
    the same X[:, cols]^T D X[:, cols] as the restricted per-pair kernels and as float64
    recomputation (reference: split.pyx:157-209 + split_matrix.py:334-354)."""
    import tabmat_b200 as tm

    n, pd, ps, dens, levels = CASES[0]
    X, full, d, rng = _build(n, pd, ps, dens, levels, seed=77)
    p = X.shape[1]
    cols = np.sort(rng.choice(p, size=max(2, int(p * frac)), replace=False)).astype(np.int32)
    rows = np.sort(rng.choice(n, size=n // 2, replace=False)).astype(np.int32)
    S = tm.RowSortedMatrix.from_split(X)
    assert X._cols_on_native_path(cols) == (frac >= 0.125)
    for r in (None, rows):
        F = (full if r is None else full[r])[:, cols]
        dd = d.astype(np.float64) if r is None else d.astype(np.float64)[r]
        ref = (F * dd[:, None]).T @ F
        monkeypatch.setattr(type(X), "_cols_on_native_path", lambda self, c: True)
        a = X.sandwich(d, r, cols)
        b = S.sandwich(d, r, cols)
        monkeypatch.setattr(type(X), "_cols_on_native_path", lambda self, c: False)
        c = X.sandwich(d, r, cols)
        for got, what in ((a, "native"), (b, "native, sorted rows"), (c, "per-pair")):
            assert got.shape == ref.shape
            cases.assert_close(got, ref, np.float32, f"cols {what}")


@pytest.mark.parametrize("suf", ["f32", "f64"])
@pytest.mark.parametrize("case", [(6007, 128, 300, 0.01, (10, 50, 200, 1000, 2000)),
                                  (5000, 32, 40, 0.05, ()), (4099, 64, 50, 0.5, (12, 300))],
                         ids=lambda c: f"n{c[0]}p{c[1]}s{c[2]}")
def test_dense_x_sparse_by_row_blocked_gather(suf, case, monkeypatch):
    """The gather form of the dense x sparse block (k_csc_dense_gather: one RED per (row block,
    column) run from a second row-blocked CSC copy): shrink the block so that a small matrix
    spans several row blocks; against float64 recomputation and against the RED form, plain and
    row-sorted, with a row restriction; and the stand-alone entry behind sandwich_dense."""
    import tabmat_b200 as tm

    n, pd, ps, dens, levels = case
    dt = cases.DTYPES[suf]
    X, full, d, rng = _build(n, pd, ps, dens, levels, seed=n + 1)
    if dt == np.float64:
        X = X.astype(np.float64)
        d = d.astype(np.float64)
    monkeypatch.setattr(type(X), "_gather_block_rows", lambda self, tdt: 512)
    plan = X._native_plan(tm._dev.torch_dtype(dt))
    assert plan is not None and plan[0][1].gcsc_row_blocks == -(-n // 512)
    rows = np.sort(rng.choice(n, size=n // 3, replace=False)).astype(np.int32)
    S = tm.RowSortedMatrix.from_split(X)
    Y, _, _, _ = _build(n, pd, ps, dens, levels, seed=n + 1)   # same data, RED form
    if dt == np.float64:
        Y = Y.astype(np.float64)
    monkeypatch.setattr(type(Y), "_gather_block_rows", lambda self, tdt: 0)
    for r in (None, rows):
        F = full if r is None else full[r]
        dd = d.astype(np.float64) if r is None else d.astype(np.float64)[r]
        ref = (F * dd[:, None]).T @ F
        cases.assert_close(X.sandwich(d, r), ref, dt, "gather form")
        cases.assert_close(S.sandwich(d, r), ref, dt, "gather form, sorted rows")
    assert not Y._native_plan(tm._dev.torch_dtype(dt))[0][1].gcsc_row_blocks
    cases.assert_close(Y.sandwich(d), (full * d.astype(np.float64)[:, None]).T @ full, dt, "RED form")


@pytest.mark.parametrize("suf", ["f32", "f64"])
def test_sparse_sandwich_dense_gather_entry(suf, monkeypatch):
    """SparseMatrix.sandwich_dense (sparse.pyx:211-260) through the gather entry point
    (TABMAT_B200_GATHER_MB shrunk so that 40k rows span many blocks) == the RED kernel == f64."""
    import tabmat_b200 as tm

    dt = cases.DTYPES[suf]
    rng = np.random.default_rng(12)
    n, p, q = 40_000, 300, 32
    A = sps.random(n, p, density=0.01, random_state=rng, format="csc").astype(dt)
    B = rng.standard_normal((n, q)).astype(dt)
    d = rng.random(n).astype(dt)
    rows = np.sort(rng.choice(n, size=n // 2, replace=False)).astype(np.int32)
    S = tm.SparseMatrix(A)
    monkeypatch.setenv("TABMAT_B200_GATHER_MB", "0")   # -> blocks of 1024 rows
    tm.reset_launch_count()
    for r in (None, rows):
        Ad = A.toarray().astype(np.float64)
        Ar, Br, dr = (Ad, B, d) if r is None else (Ad[r], B[r], d[r])
        ref = Ar.T @ (dr.astype(np.float64)[:, None] * Br.astype(np.float64))
        got = S.sandwich_dense(B, d, r, None, None)
        cases.assert_close(got, ref, dt, "gather sandwich_dense")
    assert 1024 in S.__dict__["_bcsc"]
    monkeypatch.setenv("TABMAT_B200_DXS", "red")
    cases.assert_close(S.sandwich_dense(B, d, None, None, None),
                       A.toarray().astype(np.float64).T @ (d.astype(np.float64)[:, None] * B.astype(np.float64)),
                       dt, "RED sandwich_dense")
