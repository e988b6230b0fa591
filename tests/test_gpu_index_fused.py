"""The fused index pass of tm_split_sandwich_blocks (csrc/split_index.cu): categorical self /
pair blocks from 32-byte row records and the atomics-free CSC categorical x sparse kernel,
against dense float64 recomputation (the reference's own test strategy,
tests/test_split_matrix.py:170-288) for 1..8 categorical blocks, tables inside and outside
shared memory, row restrictions, drop_first and missing codes, both row orders."""

import numpy as np
import pytest
import scipy.sparse as sps

from tests import cases

pytestmark = pytest.mark.gpu


def _build(dt, n, levels, seed, with_sparse=True, with_dense=True, p_sparse=37):
    import tabmat_b200 as tm

    rng = np.random.default_rng(seed)
    mats, cols = [], []
    if with_dense:
        X = rng.standard_normal((n, 8)).astype(dt)
        mats.append(tm.DenseMatrix(X))
        cols.append(X.astype(np.float64))
    if with_sparse:
        A = sps.random(n, p_sparse, density=0.08, random_state=rng, format="csc").astype(dt)
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
    d[rng.random(n) < 0.1] = 0
    return tm.SplitMatrix(mats), np.hstack(cols), d, rng


LEVELS = [
    (5,),
    (3, 700),
    (10, 50, 200, 1000, 2000),       # the benchmark's block structure
    (4, 9, 300, 7, 2, 40),           # 6: the most a float64 record holds
    (4, 9, 300, 7, 2, 40, 11),       # 7: the most a float32 record holds
    (4, 9, 30, 7, 2, 40, 11, 3),     # 8: falls back to the per-pair kernels
    (3000, 400),                     # pair table far beyond shared memory
]


@pytest.mark.parametrize("suf", ["f32", "f64"])
@pytest.mark.parametrize("levels", LEVELS, ids=lambda lv: "x".join(map(str, lv)))
@pytest.mark.parametrize("sparse", [True, False])
def test_fused_index_blocks(suf, levels, sparse):
    import tabmat_b200 as tm

    dt = cases.DTYPES[suf]
    n = 6007
    X, full, d, rng = _build(dt, n, levels, seed=len(levels) * 7 + sparse, with_sparse=sparse)
    rows = np.sort(rng.choice(n, size=n // 3, replace=False)).astype(np.int32)
    S = tm.RowSortedMatrix.from_split(X)
    for r in (None, rows):
        F = full if r is None else full[r]
        dd = d.astype(np.float64) if r is None else d.astype(np.float64)[r]
        ref = (F * dd[:, None]).T @ F
        cases.assert_close(X.sandwich(d, r), ref, dt, f"fused index {levels} rows={r is not None}")
        cases.assert_close(S.sandwich(d, r), ref, dt, f"fused index, sorted rows {levels}")


def test_fused_index_counts_are_exact():
    """d == 1: every categorical block is an exact (co-)occurrence count."""
    import tabmat_b200 as tm

    n = 50_000
    rng = np.random.default_rng(0)
    c1 = rng.integers(0, 12, size=n).astype(np.int32)
    c2 = rng.integers(0, 900, size=n).astype(np.int32)
    A = sps.random(n, 20, density=0.05, random_state=rng, format="csc")
    A.data[:] = 1.0
    A = A.astype(np.float32)
    X = tm.SplitMatrix([tm.SparseMatrix(A),
                        tm.CategoricalMatrix(c1, categories=np.arange(12), dtype=np.float32),
                        tm.CategoricalMatrix(c2, categories=np.arange(900), dtype=np.float32)])
    for M in (X, tm.RowSortedMatrix.from_split(X)):
        got = M.sandwich(np.ones(n, dtype=np.float32))
        o = [0, 20, 32, 932]
        np.testing.assert_array_equal(np.diag(got[o[1]:o[2], o[1]:o[2]]), np.bincount(c1, minlength=12))
        np.testing.assert_array_equal(np.diag(got[o[2]:o[3], o[2]:o[3]]), np.bincount(c2, minlength=900))
        pair = np.zeros((12, 900))
        np.add.at(pair, (c1, c2), 1.0)
        np.testing.assert_array_equal(got[o[1]:o[2], o[2]:o[3]], pair)
        cs = np.zeros((12, 20))
        Ac = A.tocoo()
        np.add.at(cs, (c1[Ac.row], Ac.col), 1.0)
        np.testing.assert_array_equal(got[o[1]:o[2], o[0]:o[1]], cs)


@pytest.mark.parametrize("suf", ["f32", "f64"])
@pytest.mark.parametrize("levels", [(5,), (10, 50, 200, 1000, 2000), (300, 40), (4, 9, 300, 7, 2, 40)],
                         ids=lambda lv: "x".join(map(str, lv)))
def test_row_blocked_csc_path(suf, levels, monkeypatch):
    """Row-blocked CSC copy of the sparse block (many-row matrices): shrink the block so that a
    6007-row matrix spans 7 row blocks and compare with the dense recomputation."""
    import tabmat_b200 as tm
    from tabmat_b200 import split_matrix

    monkeypatch.setattr(split_matrix, "CSC_ROW_BLOCK", 1000)
    monkeypatch.setenv("TABMAT_B200_CSC_ROW_BLOCKS", "1")  # opt-in layout
    dt = cases.DTYPES[suf]
    n = 6007
    X, full, d, rng = _build(dt, n, levels, seed=99)
    plan = X._native_plan(tm._dev.torch_dtype(dt))
    assert plan is not None and plan[0][1].csc_row_blocks == 7
    rows = np.sort(rng.choice(n, size=n // 3, replace=False)).astype(np.int32)
    S = tm.RowSortedMatrix.from_split(X)
    for r in (None, rows):
        F = full if r is None else full[r]
        dd = d.astype(np.float64) if r is None else d.astype(np.float64)[r]
        ref = (F * dd[:, None]).T @ F
        cases.assert_close(X.sandwich(d, r), ref, dt, f"row-blocked csc {levels}")
        cases.assert_close(S.sandwich(d, r), ref, dt, f"row-blocked csc, sorted rows {levels}")


@pytest.mark.parametrize("suf", ["f32", "f64"])
@pytest.mark.parametrize("blocked", [False, True])
def test_row_record_path_without_packed_codes(suf, blocked, monkeypatch):
    """TABMAT_B200_CSC_PACKED=0: no bit-packed codes in the plan, so categorical x sparse goes
    through the per-call 32-byte row records (the layout for code sets wider than 64 bits)."""
    import tabmat_b200 as tm
    from tabmat_b200 import split_matrix

    monkeypatch.setenv("TABMAT_B200_CSC_PACKED", "0")
    if blocked:
        monkeypatch.setattr(split_matrix, "CSC_ROW_BLOCK", 1000)
    dt = cases.DTYPES[suf]
    n = 6007
    X, full, d, rng = _build(dt, n, (10, 50, 200, 1000, 2000), seed=5)
    plan = X._native_plan(tm._dev.torch_dtype(dt))
    assert plan is not None and not plan[0][1].csc_cat_codes
    assert (plan[0][1].csc_row_blocks == 7) == blocked
    ref = (full * d.astype(np.float64)[:, None]).T @ full
    cases.assert_close(X.sandwich(d), ref, dt, "row records")
    monkeypatch.delenv("TABMAT_B200_CSC_PACKED")
    Y, full2, d2, _ = _build(dt, n, (10, 50, 200, 1000, 2000), seed=5)
    assert Y._native_plan(tm._dev.torch_dtype(dt))[0][1].csc_cat_codes
    cases.assert_close(Y.sandwich(d2), ref, dt, "packed codes")
