"""Pin the oracle (oracle/tabmat_oracle.c) before trusting it: against the golden vectors
generated from the reference itself (tests/golden/make_golden.py), against a dense float64
recomputation (the reference's own test strategy) and — when oracle/_ref is
present — against the reference's compiled kernels run right here on random shapes."""

import numpy as np
import pytest
import scipy.sparse as sps

from tests import cases

ALL_CASES = list(cases.boundary_cases())


@pytest.mark.parametrize("name,kind,args", ALL_CASES, ids=[c[0] for c in ALL_CASES])
def test_oracle_matches_reference_golden(name, kind, args, golden_boundary):
    got = cases.run_oracle(kind, args)
    ref = golden_boundary[name]
    assert got.dtype == ref.dtype
    cases.assert_close(got, ref, got.dtype, name)


@pytest.mark.parametrize("name,kind,args", ALL_CASES[::5], ids=[c[0] for c in ALL_CASES[::5]])
def test_oracle_matches_dense_recomputation(name, kind, args):
    got = cases.run_oracle(kind, args)
    cases.assert_close(got, cases.run_numpy(kind, args), got.dtype, name)


def test_golden_covers_every_case(golden_boundary):
    assert set(golden_boundary.files) == {c[0] for c in ALL_CASES}


def test_cat_sandwich_unit_weights_is_exact_count():
    """With d == 1 the categorical sandwich is an exact count (bit-exact vs np.bincount)."""
    from oracle import c_oracle as orc

    rng = np.random.default_rng(0)
    codes = rng.integers(0, 37, size=5000).astype(np.int32)
    for dt in (np.float32, np.float64):
        got = orc.cat_sandwich(codes, np.ones(5000, dt), None, 37, False)
        assert np.array_equal(got, np.bincount(codes, minlength=37).astype(dt))


def _ref_ext():
    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("oracle/_ref not built")
    try:
        return ref_loader.load_ext()
    except ImportError as e:  # e.g. a different CPU / python on the box
        pytest.skip(f"oracle/_ref not loadable: {e}")


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("seed", range(4))
def test_oracle_vs_compiled_reference_kernels(dt, seed):
    """The reference's own test_fast_sandwich.py shapes (random, < 200) through both."""
    from oracle import c_oracle as orc

    ext = _ref_ext()
    rng = np.random.default_rng(seed)
    n, p = int(rng.integers(1, 200)), int(rng.integers(1, 60))
    X = np.asfortranarray(rng.standard_normal((n, p)).astype(dt))
    d = rng.random(n).astype(dt)
    d[rng.choice(n, size=min(10, n), replace=False)] = 0
    rows = np.flatnonzero(np.abs(d) > 1e-14).astype(np.int32)
    cols = rng.permutation(p)[: max(1, p // 2)].astype(np.int32)
    cols.sort()
    ref = ext.dense.dense_sandwich(X, d, rows, cols)
    cases.assert_close(orc.dense_sandwich(X, d, rows, cols), ref, dt, "dense")
    A = sps.random(n, p, density=0.2, random_state=rng, format="csc").astype(dt)
    A.sort_indices()
    ar_n, ar_p = np.arange(n, dtype=np.int32), np.arange(p, dtype=np.int32)
    ref = ext.sparse.sparse_sandwich(A, A.tocsr(), d, ar_n, ar_p)
    cases.assert_close(orc.sparse_sandwich(A, d), ref, dt, "sparse")
    ref = ext.sparse.csr_dense_sandwich(A.tocsr(), X, d, rows, cols, cols)
    cases.assert_close(orc.csr_dense_sandwich(A, X, d, rows, cols, cols), ref, dt, "csr_dense")
    K = 11
    codes = rng.integers(0, K, size=n).astype(np.int32)
    ref = ext.split.sandwich_cat_dense(codes, K, d, X, rows, cols, False, False, False)
    cases.assert_close(orc.cat_dense_sandwich(codes, K, d, X, rows, cols), ref, dt, "cat_dense")


def test_ref_split_driver_matches_dense_recomputation():
    """oracle/ref_split.py (the CPU baseline bench.py times) == dense recomputation."""
    from oracle.ref_split import RefSplit

    ext = _ref_ext()
    rng = np.random.default_rng(9)
    n = 500
    for dt in (np.float32, np.float64):
        X = rng.standard_normal((n, 6)).astype(dt)
        A = sps.random(n, 12, density=0.1, random_state=rng, format="csc").astype(dt)
        c1 = rng.integers(0, 4, size=n).astype(np.int32)
        c2 = rng.integers(0, 7, size=n).astype(np.int32)
        d = rng.random(n).astype(dt)
        S = RefSplit([("dense", X), ("sparse", A), ("cat", c1, 4), ("cat", c2, 7)], ext)
        got = S.sandwich(d)
        full = np.hstack([X.astype(np.float64), A.toarray().astype(np.float64),
                          np.eye(4)[c1], np.eye(7)[c2]])
        ref = full.T @ (d.astype(np.float64)[:, None] * full)
        cases.assert_close(got, ref, dt, "ref split")
