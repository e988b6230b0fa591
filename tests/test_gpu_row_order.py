"""Row-sorted storage (tabmat_b200.RowSortedMatrix), the run-aggregating kernels behind it and
device-side row indexing.

The reference keeps the caller's row order (split_matrix.py:171-267); the wrapper must be
indistinguishable from the plain SplitMatrix through the MatrixBase API, so every check here
is "sorted storage == original storage == dense float64 recomputation / oracle".
"""

import numpy as np
import pytest
import scipy.sparse as sps

from tests import cases

pytestmark = pytest.mark.gpu


def _mats(dt, n, seed, K_big=300, K_mid=270, p_dense=8, missing=True):
    import tabmat_b200 as tm

    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, p_dense)).astype(dt)
    A = sps.random(n, 11, density=0.2, random_state=rng, format="csc").astype(dt)
    c_small = rng.integers(0, 5, size=n).astype(np.int32)
    c_big = rng.integers(0, K_big, size=n).astype(np.int32)
    c_mid = rng.integers(0, K_mid, size=n).astype(np.int32)
    if missing:
        c_mid[rng.random(n) < 0.1] = -1
    mats = [
        tm.DenseMatrix(X),
        tm.SparseMatrix(A),
        tm.CategoricalMatrix(c_small, categories=np.arange(5), dtype=dt),
        tm.CategoricalMatrix(c_big, categories=np.arange(K_big), dtype=dt, drop_first=True),
        tm.CategoricalMatrix(c_mid, categories=np.arange(K_mid), dtype=dt,
                             cat_missing_method="zero"),
    ]
    full = np.hstack([
        X.astype(np.float64), A.toarray().astype(np.float64), np.eye(5)[c_small],
        np.eye(K_big)[c_big][:, 1:],
        np.where((c_mid >= 0)[:, None], np.eye(K_mid)[np.maximum(c_mid, 0)], 0.0),
    ])
    d = rng.standard_normal(n).astype(dt)
    d[rng.random(n) < 0.1] = 0
    return mats, full, d, rng


@pytest.mark.parametrize("suf", ["f32", "f64"])
@pytest.mark.parametrize("n", [1, 257, 5000])
def test_row_sorted_matches_original_and_dense(suf, n):
    import tabmat_b200 as tm

    dt = cases.DTYPES[suf]
    mats, full, d, rng = _mats(dt, n, seed=11 + n)
    X = tm.SplitMatrix(mats)
    S = tm.RowSortedMatrix.from_split(X)
    assert S.shape == X.shape and S.dtype == X.dtype
    # the two many-level blocks are the sort keys, widest first
    assert list(S.sort_blocks) == [3, 4]
    codes = S.matrices[3].indices
    assert np.all(np.diff(codes) >= 0)  # stored rows are sorted by the primary key
    p = X.shape[1]
    rows = np.sort(rng.choice(n, size=max(1, n // 3), replace=False)).astype(np.int32)
    cols = np.sort(rng.choice(p, size=p // 2, replace=False)).astype(np.int32)
    for r in (None, rows):
        for c in (None, cols):
            F = full if r is None else full[r]
            dd = d.astype(np.float64) if r is None else d.astype(np.float64)[r]
            F = F if c is None else F[:, c]
            ref = (F * dd[:, None]).T @ F
            got = S.sandwich(d, r, c)
            assert got.dtype == np.float64
            cases.assert_close(got, ref, dt, f"sorted sandwich rows={r is not None} cols={c is not None}")
            cases.assert_close(got, X.sandwich(d, r, c), dt, "sorted vs original sandwich")
            v = rng.standard_normal(n).astype(dt)
            ref = F.T @ (v.astype(np.float64) if r is None else v.astype(np.float64)[r])
            cases.assert_close(S.transpose_matvec(v, r, c), ref, dt, "sorted transpose_matvec")
    v = rng.standard_normal(p).astype(dt)
    cases.assert_close(S.matvec(v), full @ v, dt, "sorted matvec")
    vc = v.copy()
    mask = np.zeros(p, bool)
    mask[cols] = True
    cases.assert_close(S.matvec(v, cols), full[:, cols] @ vc[cols], dt, "sorted matvec cols")
    # out= accumulates in place and returns the same object (test_matrices.py:129-171)
    out = np.ones(n, dtype=dt)
    res = S.matvec(v, out=out)
    assert res is out
    cases.assert_close(out, 1.0 + full @ v, dt, "sorted matvec out=")
    out = np.ones(p, dtype=dt)
    vn = rng.standard_normal(n).astype(dt)
    res = S.transpose_matvec(vn, out=out)
    assert res is out
    cases.assert_close(out, 1.0 + full.T @ vn, dt, "sorted transpose_matvec out=")
    # array surface in the caller's order
    np.testing.assert_array_equal(S.toarray(), X.toarray())
    np.testing.assert_array_equal(S.unsorted().toarray(), X.toarray())
    np.testing.assert_array_equal(S[rows].toarray(), X.toarray()[rows])
    for i in (0, 9, p - 1):
        np.testing.assert_array_equal(np.asarray(S.getcol(i).toarray()).ravel(),
                                      X.toarray()[:, i])


@pytest.mark.parametrize("suf", ["f32", "f64"])
def test_row_sorted_standardize(suf):
    import tabmat_b200 as tm

    dt = cases.DTYPES[suf]
    n = 3000
    mats, full, d, rng = _mats(dt, n, seed=5)
    X = tm.SplitMatrix(mats)
    S = tm.RowSortedMatrix.from_split(X)
    w = rng.random(n).astype(dt)
    w /= w.sum()
    Zs, ms, ss = S.standardize(w, True, True)
    Zx, mx, sx = X.standardize(w, True, True)
    cases.assert_close(ms, mx, dt, "means")
    cases.assert_close(ss, sx, dt, "stds")
    cases.assert_close(Zs.sandwich(np.abs(d)), Zx.sandwich(np.abs(d)), dt, "standardized sandwich")
    v = rng.standard_normal(X.shape[1]).astype(dt)
    cases.assert_close(Zs.matvec(v), Zx.matvec(v), dt, "standardized matvec")
    vn = rng.standard_normal(n).astype(dt)
    cases.assert_close(Zs.transpose_matvec(vn), Zx.transpose_matvec(vn), dt,
                       "standardized transpose_matvec")


def test_row_sorted_device_tensors_stay_on_device():
    import torch

    import tabmat_b200 as tm

    dt = np.float32
    n = 4096
    mats, full, d, rng = _mats(dt, n, seed=3)
    S = tm.RowSortedMatrix.from_split(tm.SplitMatrix(mats))
    d_t = torch.from_numpy(d).cuda()
    out = S.sandwich(d_t)
    assert out.is_cuda and out.dtype == torch.float64
    ref = (full * d.astype(np.float64)[:, None]).T @ full
    cases.assert_close(out.cpu().numpy(), ref, dt, "device sandwich")
    y = S.matvec(torch.from_numpy(rng.standard_normal(S.shape[1]).astype(dt)).cuda())
    assert y.is_cuda and y.shape == (n,)


@pytest.mark.parametrize("suf", ["f32", "f64"])
@pytest.mark.parametrize("mode", [1, 2])
def test_cross_kernel_variants_agree_with_the_oracle(suf, mode):
    """k_dense_cross_runs (mode 1, forced on UNSORTED rows: every run has length ~1) and
    k_dense_cross_fused (mode 2) both against the oracle's per-pair blocks."""
    import tabmat_b200 as tm
    from oracle import c_oracle as orc
    from tabmat_b200._lib import lib

    dt = cases.DTYPES[suf]
    n = 7001
    rng = np.random.default_rng(8)
    X = rng.standard_normal((n, 16)).astype(dt)
    A = sps.random(n, 40, density=0.05, random_state=rng, format="csc").astype(dt)
    c1 = rng.integers(0, 400, size=n).astype(np.int32)
    c2 = np.sort(rng.integers(0, 300, size=n)).astype(np.int32)  # long runs
    c2[rng.random(n) < 0.05] = -1
    d = rng.standard_normal(n).astype(dt)
    rows = np.sort(rng.choice(n, size=n // 2, replace=False)).astype(np.int32)
    S = tm.SplitMatrix([
        tm.DenseMatrix(X), tm.SparseMatrix(A),
        tm.CategoricalMatrix(c1, categories=np.arange(400), dtype=dt),
        tm.CategoricalMatrix(c2, categories=np.arange(300), dtype=dt, cat_missing_method="zero"),
    ])
    lib.tm_set_cross_runs_mode(mode)
    try:
        for r in (None, rows):
            got = S.sandwich(d, r)
            o = [0, 16, 56, 456, 756]
            cases.assert_close(got[o[1]:o[2], o[0]:o[1]], orc.csr_dense_sandwich(A, X, d, r), dt,
                               "sparse x dense")
            cases.assert_close(got[o[2]:o[3], o[0]:o[1]], orc.cat_dense_sandwich(c1, 400, d, X, r),
                               dt, "cat400 x dense")
            cases.assert_close(got[o[3]:o[4], o[0]:o[1]], orc.cat_dense_sandwich(c2, 300, d, X, r),
                               dt, "cat300 (sorted, missing) x dense")
    finally:
        lib.tm_set_cross_runs_mode(0)


@pytest.mark.parametrize("suf", ["f32", "f64"])
@pytest.mark.parametrize("layout", ["random", "sorted", "short-runs"])
@pytest.mark.parametrize("K", [7, 2000, 20000])
def test_histogram_kernels_on_runs(suf, layout, K):
    """k_cat_hist2 / k_cat_cat2: segmented pre-aggregation over runs of equal codes."""
    import torch

    from oracle import c_oracle as orc
    from tabmat_b200.ext import categorical as ecat
    from tabmat_b200.ext import split as esplit

    dt = cases.DTYPES[suf]
    n = 100_003
    rng = np.random.default_rng(K)
    codes = rng.integers(0, K, size=n).astype(np.int32)
    if layout == "sorted":
        codes = np.sort(codes)
    elif layout == "short-runs":
        codes = np.repeat(rng.integers(0, K, size=n // 5 + 1), 5)[:n].astype(np.int32)
    codes[rng.random(n) < 0.02] = -1
    other = rng.integers(0, 6, size=n).astype(np.int32)
    d = rng.standard_normal(n).astype(dt)
    rows = np.sort(rng.choice(n, size=n // 3, replace=False)).astype(np.int32)
    t = lambda a: None if a is None else torch.from_numpy(a).cuda()  # noqa: E731
    for r in (None, rows):
        got = ecat.sandwich_categorical(t(codes), t(d), t(r), K, False).cpu().numpy()
        cases.assert_close(got, orc.cat_sandwich(codes, d, r, K), dt, f"hist {layout} K={K}")
        got = esplit.sandwich_cat_cat(t(codes), t(other), K, 6, t(d), t(r), False, False)
        cases.assert_close(got.cpu().numpy(), orc.cat_cat_sandwich(codes, other, K, 6, d, r), dt,
                           f"cat x cat {layout} K={K}")
    # d == 1: exact counts (bit-exact for categorical index counts)
    ones = np.ones(n, dtype=dt)
    got = ecat.sandwich_categorical(t(codes), t(ones), None, K, False).cpu().numpy()
    np.testing.assert_array_equal(got, np.bincount(codes[codes >= 0], minlength=K).astype(dt))


@pytest.mark.parametrize("suf", ["f32", "f64"])
def test_sparse_row_indexing_on_device(suf):
    import torch

    import tabmat_b200 as tm

    dt = cases.DTYPES[suf]
    rng = np.random.default_rng(2)
    A = sps.random(500, 23, density=0.1, random_state=rng, format="csc").astype(dt)
    M = tm.SparseMatrix(A)
    idx = rng.integers(-500, 500, size=777)  # duplicates, negatives, any order
    for key in (idx, slice(10, 400, 3), rng.random(500) < 0.3, torch.from_numpy(idx).cuda()):
        sub = M[key, :] if not isinstance(key, torch.Tensor) else M[key, :]
        k = key.cpu().numpy() if isinstance(key, torch.Tensor) else key
        ref = A.tocsr()[k].toarray()
        assert isinstance(sub, tm.SparseMatrix) and sub.shape == ref.shape
        np.testing.assert_array_equal(sub.toarray(), ref)
        v = rng.standard_normal(23).astype(dt)
        cases.assert_close(sub.matvec(v), ref.astype(np.float64) @ v, dt, "row-subset matvec")
        w = rng.standard_normal(ref.shape[0]).astype(dt)
        cases.assert_close(sub.transpose_matvec(w), ref.astype(np.float64).T @ w, dt,
                           "row-subset transpose_matvec")


def test_row_sharded_wrapper_accepts_row_sorted_shards():
    import torch

    import tabmat_b200 as tm
    from tabmat_b200.distributed import RowShardedMatrix

    dt = np.float32
    n = 6000
    mats, full, d, rng = _mats(dt, n, seed=21)
    X = tm.SplitMatrix(mats)
    R = RowShardedMatrix(tm.RowSortedMatrix.from_split(X), n, pack=True,
                         reduce_dtype=torch.float32)
    got = R.sandwich(torch.from_numpy(d).cuda())
    ref = (full * d.astype(np.float64)[:, None]).T @ full
    cases.assert_close(got.cpu().numpy(), ref, dt, "row-sharded(row-sorted) sandwich")
    rows = np.sort(rng.choice(n, size=n // 4, replace=False))
    got = R.sandwich(torch.from_numpy(d).cuda(), rows=rows)
    ref = (full[rows] * d.astype(np.float64)[rows, None]).T @ full[rows]
    cases.assert_close(got.cpu().numpy(), ref, dt, "row-sharded(row-sorted) sandwich rows")


@pytest.mark.parametrize("suf", ["f32", "f64"])
@pytest.mark.parametrize("layout", ["dense-first", "dense-middle", "no-dense", "dense-only"])
def test_sandwich_into_host_buffer(suf, layout):
    """Two-phase host-buffer path (blocks without the dense operand are copied to the host
    while the dense passes run) == the plain sandwich, for every column layout."""
    import torch

    import tabmat_b200 as tm

    dt = cases.DTYPES[suf]
    n = 4099
    mats, full, d, rng = _mats(dt, n, seed=31)
    if layout == "no-dense":
        mats = mats[1:]
    elif layout == "dense-only":
        mats = mats[:1]
    p = sum(m.shape[1] for m in mats)
    indices = None
    if layout == "dense-middle":
        # dense columns at result positions 40..47, everything else around them
        order = list(range(8, 48)) + list(range(0, 8)) + list(range(48, p))
        pos = np.empty(p, dtype=np.int64)
        pos[order] = np.arange(p)
        indices, o = [], 0
        for m in mats:
            indices.append(np.sort(pos[o:o + m.shape[1]]))
            o += m.shape[1]
    X = tm.SplitMatrix(mats, indices)
    rows = np.sort(rng.choice(n, size=n // 2, replace=False)).astype(np.int32)
    variants = [X] if layout in ("no-dense", "dense-only") else [X, tm.RowSortedMatrix.from_split(X)]
    for M in variants:
        for r in (None, rows):
            ref = X.sandwich(d, r)
            out_np = np.full((p, p), np.nan)
            assert M.sandwich_into(d, out_np, r) is out_np
            torch.cuda.synchronize()
            cases.assert_close(out_np, ref, dt, f"sandwich_into numpy {layout}")
            out_t = torch.full((p, p), float("nan"), dtype=torch.float64).pin_memory()
            d_pinned = torch.from_numpy(d).pin_memory()
            M.sandwich_into(d_pinned, out_t, r)
            torch.cuda.synchronize()
            cases.assert_close(out_t.numpy(), ref, dt, f"sandwich_into pinned {layout}")
    with pytest.raises(ValueError):
        X.sandwich_into(d, np.zeros((p, p), dtype=np.float32))


@pytest.mark.parametrize("dense_first", [True, False])
def test_sandwich_into_with_a_collective_between_the_phases(dense_first):
    """The row-sharded caller's hook: `reduce` sees the index part of the flat workspace first,
    then the dense block's part (both contiguous when the dense block comes first), and decides
    whether this rank holds the result."""
    import torch

    import tabmat_b200 as tm

    dt = np.float32
    n = 3001
    mats, full, d, rng = _mats(dt, n, seed=41)
    if not dense_first:
        mats = mats[1:] + mats[:1]
        full = np.hstack([full[:, 8:], full[:, :8]])
    X = tm.SplitMatrix(mats)
    p = X.shape[1]
    ref = (full * d.astype(np.float64)[:, None]).T @ full
    seen = []

    def twice(ws):  # a two-rank sum where both ranks hold the same shard
        seen.append(int(ws.numel()))
        ws *= 2
        return True

    for M in (X, tm.RowSortedMatrix.from_split(X)):
        seen.clear()
        out = np.full((p, p), np.nan)
        assert M.sandwich_into(d, out, reduce=twice) is out
        torch.cuda.synchronize()
        cases.assert_close(out, 2 * ref, dt, "sandwich_into with reduce hook")
        if dense_first:
            assert seen[1] == 8 * p and len(seen) == 2  # index part, then the dense block's part
        else:
            assert len(seen) == 1  # one collective over the whole workspace
        assert M.sandwich_into(d, out, reduce=lambda ws: False) is None
