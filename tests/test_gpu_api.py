"""MatrixBase API contract on the device classes, checked the way the reference checks its own
classes (dense numpy recomputation from ``toarray()``; reference: tests/test_matrices.py,
tests/test_standardized_mat.py, tests/test_categorical_matrix.py — restated, not copied).

Covers: out-parameter semantics and shape errors, alignment / dtype errors, empty and list
restrictions, 2-d right-hand sides, every cross-sandwich pair, getcol / toarray / astype,
column means and standard deviations, standardize, indexing, multiply, hstack."""

import numpy as np
import pytest
import scipy.sparse as sps

pytestmark = pytest.mark.gpu

N = 6
BASE = np.array([[0.0, 1.5, 0.0], [2.0, 0.0, -1.0], [0.0, 0.0, 0.0], [1.0, -3.0, 0.5],
                 [0.0, 0.25, 0.0], [4.0, 0.0, 2.0]])
CODES = np.array([2, 0, 1, 2, 2, 0])


def _zoo():
    import tabmat_b200 as tm

    def make():
        return {
            "dense_C": tm.DenseMatrix(np.ascontiguousarray(BASE)),
            "dense_F": tm.DenseMatrix(np.asfortranarray(BASE)),
            "sparse": tm.SparseMatrix(sps.csc_matrix(BASE)),
            "sparse_i64": tm.SparseMatrix(
                (sps.csc_matrix(BASE).data, sps.csc_matrix(BASE).indices.astype(np.int64),
                 sps.csc_matrix(BASE).indptr.astype(np.int64)), shape=BASE.shape),
            "cat": tm.CategoricalMatrix(CODES),
            "cat_drop": tm.CategoricalMatrix(CODES, drop_first=True),
            "cat_missing": tm.CategoricalMatrix(np.array([1, -1, 0, 1, -1, 2]),
                                                categories=np.arange(3),
                                                cat_missing_method="zero"),
        }

    z = make()
    z["split"] = tm.SplitMatrix(list(make().values()))
    rng = np.random.default_rng(0)
    out = dict(z)
    for name, m in z.items():
        p = m.shape[1]
        out["std_shift:" + name] = tm.StandardizedMatrix(m, rng.random(p))
        out["std_both:" + name] = tm.StandardizedMatrix(m, rng.random(p), rng.random(p) + 0.5)
    return out


NAMES = None


def _names():
    global NAMES
    if NAMES is None:
        NAMES = ["dense_C", "dense_F", "sparse", "sparse_i64", "cat", "cat_drop", "cat_missing",
                 "split"]
        NAMES = NAMES + ["std_shift:" + n for n in NAMES] + ["std_both:" + n for n in NAMES]
    return NAMES


@pytest.fixture(scope="module")
def zoo():
    return _zoo()


RESTR = [None, [], [1], np.array([0, 2])]


def _fit(cols, p):
    """Drop restriction entries beyond the matrix width (the zoo has 2- and 3-column blocks)."""
    if cols is None:
        return None
    kept = [int(c) for c in np.asarray(cols, dtype=int) if c < p]
    return kept if isinstance(cols, list) else np.asarray(kept, dtype=int)


def _sel(A, rows, cols):
    if rows is not None:
        A = A[np.asarray(rows, dtype=int), :]
    if cols is not None:
        A = A[:, np.asarray(cols, dtype=int)]
    return A


@pytest.mark.parametrize("name", _names())
@pytest.mark.parametrize("cols", RESTR)
def test_matvec(zoo, name, cols):
    mat = zoo[name]
    cols = _fit(cols, mat.shape[1])
    A = mat.toarray()
    v = np.linspace(-1.0, 2.0, mat.shape[1])
    expected = _sel(A, None, cols) @ (v if cols is None else v[np.asarray(cols, dtype=int)])
    np.testing.assert_allclose(mat.matvec(v, cols), expected, atol=1e-12)
    np.testing.assert_allclose(mat.matvec(list(v), cols), expected, atol=1e-12)
    # out: wrong shape raises, right shape accumulates in place and is returned
    with pytest.raises(ValueError, match="first dimension of 'out' must be"):
        mat.matvec(v, cols, np.zeros(mat.shape[0] + 1))
    out = np.arange(mat.shape[0], dtype=np.float64)
    keep = out.copy()
    res = mat.matvec(v, cols, out)
    assert res is out
    np.testing.assert_allclose(out, keep + expected, atol=1e-12)


@pytest.mark.parametrize("name", _names())
def test_matvec_2d_and_misaligned(zoo, name):
    import tabmat_b200 as tm

    mat = zoo[name]
    inner = mat.mat if isinstance(mat, tm.StandardizedMatrix) else mat
    has_cat = isinstance(inner, tm.CategoricalMatrix) or (
        isinstance(inner, tm.SplitMatrix)
        and any(isinstance(m, tm.CategoricalMatrix) for m in inner.matrices))
    V = np.arange(2.0 * mat.shape[1]).reshape(mat.shape[1], 2)
    if has_cat:
        with pytest.raises(NotImplementedError, match="only implemented for 1d"):
            mat.matvec(V)
    else:
        np.testing.assert_allclose(mat.matvec(V), mat.toarray() @ V, atol=1e-12)
    with pytest.raises(ValueError):
        mat.matvec(np.ones(mat.shape[1] + 1))
    with pytest.raises(ValueError):
        mat.transpose_matvec(np.ones(mat.shape[0] + 1))


@pytest.mark.parametrize("name", _names())
@pytest.mark.parametrize("rows", RESTR)
@pytest.mark.parametrize("cols", RESTR)
def test_transpose_matvec(zoo, name, rows, cols):
    mat = zoo[name]
    cols = _fit(cols, mat.shape[1])
    A = mat.toarray()
    v = np.array([3.0, -0.1, 0.0, 1.0, 2.0, -2.0])
    expected = _sel(A, rows, cols).T @ (v if rows is None else v[np.asarray(rows, dtype=int)])
    np.testing.assert_allclose(mat.transpose_matvec(v, rows, cols), expected, atol=1e-12)
    with pytest.raises(ValueError, match="dimension of 'out' must be"):
        mat.transpose_matvec(v, rows, cols, np.zeros(mat.shape[1] + 1))
    out = np.ones(mat.shape[1])
    res = mat.transpose_matvec(v, rows, cols, out)
    assert res is out
    full = np.ones(mat.shape[1])
    if cols is None:
        full += expected
    else:
        full[np.asarray(cols, dtype=int)] += expected
    np.testing.assert_allclose(out, full, atol=1e-12)


@pytest.mark.parametrize("name", _names())
@pytest.mark.parametrize("rows", RESTR)
@pytest.mark.parametrize("cols", RESTR)
def test_sandwich(zoo, name, rows, cols):
    mat = zoo[name]
    cols = _fit(cols, mat.shape[1])
    A = _sel(mat.toarray(), rows, cols)
    d = np.array([3.0, 0.1, 1.0, -2.0, 0.0, 0.5])
    dd = d if rows is None else d[np.asarray(rows, dtype=int)]
    res = mat.sandwich(d, rows, cols)
    if not isinstance(res, np.ndarray):
        res = res.toarray()  # categorical: scipy dia_matrix
    np.testing.assert_allclose(res, A.T @ (dd[:, None] * A), atol=1e-12)


@pytest.mark.parametrize("name", _names())
def test_sandwich_errors(zoo, name):
    mat = zoo[name]
    with pytest.raises(ValueError, match="not aligned"):
        mat.sandwich(np.ones(mat.shape[0] + 1))
    with pytest.raises(TypeError, match="same dtype"):
        mat.sandwich(np.ones(mat.shape[0], dtype=np.float32))


PLAIN = ["dense_C", "dense_F", "sparse", "cat", "cat_drop", "cat_missing"]


@pytest.mark.parametrize("left", PLAIN)
@pytest.mark.parametrize("right", PLAIN)
@pytest.mark.parametrize("rows", [None, [0, 3, 5]])
@pytest.mark.parametrize("lcols", [None, [0], [0, 1]])
@pytest.mark.parametrize("rcols", [None, [1]])
def test_cross_sandwich(zoo, left, right, rows, lcols, rcols):
    L, R = zoo[left], zoo[right]
    if (left.startswith("dense") and right.startswith("dense")) or left == right == "sparse":
        with pytest.raises(TypeError):  # same-type blocks are merged, never crossed
            L._cross_sandwich(R, np.ones(N), rows, lcols, rcols)
        return
    d = np.array([3.0, 0.1, 1.0, -2.0, 0.0, 0.5])
    dd = d if rows is None else d[rows]
    res = L._cross_sandwich(R, d, rows, lcols, rcols)
    expected = _sel(L.toarray(), rows, lcols).T @ (dd[:, None] * _sel(R.toarray(), rows, rcols))
    np.testing.assert_allclose(np.asarray(res), expected, atol=1e-12)


@pytest.mark.parametrize("name", _names())
def test_getcol_toarray_rmatmul(zoo, name):
    mat = zoo[name]
    A = mat.toarray()
    assert A.shape == mat.shape
    for i in (0, -1):
        np.testing.assert_allclose(mat.getcol(i).toarray()[:, 0], A[:, i], atol=1e-12)
    v = np.arange(1.0, mat.shape[0] + 1)
    np.testing.assert_allclose(v @ mat, v @ A, atol=1e-12)
    np.testing.assert_allclose(mat @ np.ones(mat.shape[1]), A.sum(1), atol=1e-12)


@pytest.mark.parametrize("name", PLAIN + ["split"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_astype_and_result_dtype(zoo, name, dtype):
    mat = _zoo()[name].astype(dtype)
    assert mat.dtype == dtype
    v = np.ones(mat.shape[1], dtype=dtype)
    assert mat.matvec(v).dtype == dtype
    res = mat.sandwich(np.ones(mat.shape[0], dtype=dtype))
    if name == "split":
        assert res.dtype == np.float64  # split_matrix.py:336
    elif not name.startswith("cat"):
        assert res.dtype == dtype


@pytest.mark.parametrize("name", PLAIN + ["split"])
@pytest.mark.parametrize("center", [False, True])
@pytest.mark.parametrize("scale", [False, True])
def test_col_stats_and_standardize(zoo, name, center, scale):
    import tabmat_b200 as tm

    mat = zoo[name]
    A = mat.toarray()
    w = np.array([0.1, 0.3, 0.05, 0.25, 0.1, 0.2])
    means = A.T @ w
    stds = np.sqrt(np.maximum(((A - means) ** 2).T @ w, 0))
    np.testing.assert_allclose(mat._get_col_means(w), means, atol=1e-12)
    np.testing.assert_allclose(mat._get_col_stds(w, means), stds, atol=1e-7)
    S, m_out, s_out = mat.standardize(w, center, scale)
    assert isinstance(S, tm.StandardizedMatrix)
    np.testing.assert_allclose(m_out, means if center else 0 * means, atol=1e-12)
    if scale:
        np.testing.assert_allclose(s_out, stds, atol=1e-7)
    else:
        assert s_out is None
    expected = A.copy()
    if center:
        expected = expected - means
    if scale:
        mult = np.where(np.abs(stds) < 1e-7, 1.0, 1 / np.where(stds == 0, 1, stds))
        expected = expected * mult
    np.testing.assert_allclose(S.toarray(), expected, atol=1e-7)
    d = np.array([3.0, 0.1, 1.0, 2.0, 0.0, 0.5])
    np.testing.assert_allclose(S.sandwich(d), expected.T @ (d[:, None] * expected), atol=1e-6)


def test_zero_variance_column_keeps_mult_one():
    import tabmat_b200 as tm

    X = tm.DenseMatrix(np.column_stack([np.ones(5), np.arange(5.0)]))
    S, _, stds = X.standardize(np.full(5, 0.2), True, True)
    assert stds[0] == 0 and S.mult[0] == 1.0


def test_col_std_accuracy_large_mean():
    """Two-pass shifted form (dense_matrix.py:180-187): large mean, tiny spread."""
    import tabmat_b200 as tm

    for dt, tol in ((np.float64, 1e-9), (np.float32, 1e-2)):
        col = (1e4 + np.array([0.0, 0.1, 0.2, 0.3, 0.4])).astype(dt)
        X = tm.DenseMatrix(col[:, None])
        w = np.full(5, 0.2, dtype=dt)
        _, _, stds = X.standardize(w, True, True)
        assert abs(stds[0] - np.std(col.astype(np.float64))) < tol * np.std(col)


@pytest.mark.parametrize("name", PLAIN + ["split"])
def test_row_indexing_and_multiply(zoo, name):
    mat = zoo[name]
    A = mat.toarray()
    rows = [0, 2, 5]
    np.testing.assert_allclose(mat[rows, :].toarray(), A[rows, :], atol=1e-12)
    scale = np.arange(1.0, 7.0)
    np.testing.assert_allclose(mat.multiply(scale).toarray(), A * scale[:, None], atol=1e-12)


def test_column_indexing_of_leaf_matrices(zoo):
    for name in ("dense_C", "sparse", "cat"):
        mat = zoo[name]
        A = mat.toarray()
        np.testing.assert_allclose(mat[:, [0, 1]].toarray(), A[:, [0, 1]], atol=1e-12)
        np.testing.assert_allclose(mat[[1, 3], [0]].toarray(), A[np.ix_([1, 3], [0])], atol=1e-12)


def test_hstack_type_rules(zoo):
    import tabmat_b200 as tm

    assert isinstance(tm.hstack([zoo["dense_C"], zoo["dense_F"]]), tm.DenseMatrix)
    assert isinstance(tm.hstack([zoo["sparse"], zoo["sparse_i64"]]), tm.SparseMatrix)
    mixed = tm.hstack([zoo["dense_C"], zoo["sparse"], zoo["cat"]])
    assert isinstance(mixed, tm.SplitMatrix)
    np.testing.assert_allclose(
        mixed.toarray(), np.hstack([zoo[k].toarray() for k in ("dense_C", "sparse", "cat")]))
    with pytest.raises(ValueError):
        tm.hstack([])


def test_split_merges_blocks_and_rejects_unsorted_indices(zoo):
    import tabmat_b200 as tm

    S = zoo["split"]
    kinds = [type(m).__name__ for m in S.matrices]
    assert kinds.count("DenseMatrix") == 1 and kinds.count("SparseMatrix") == 1
    assert kinds.count("CategoricalMatrix") == 3
    with pytest.raises(ValueError, match="sorted"):
        tm.SplitMatrix([zoo["dense_C"], zoo["cat"]], [np.array([2, 0, 1]), np.array([3, 4, 5])])


def test_categorical_int_vector_roundtrip_and_names():
    import tabmat_b200 as tm

    cat = tm.CategoricalMatrix(np.array(["b", "a", "b", "c"]), column_name="x")
    assert cat.get_names() == ["x[a]", "x[b]", "x[c]"]
    res = cat.matvec(np.array([1, 2, 3]))
    assert res.dtype.kind == "i" and res.tolist() == [2, 1, 2, 3]
    np.testing.assert_array_equal(cat.recover_orig(), np.array(["b", "a", "b", "c"]))
    with pytest.raises(ValueError, match="missing"):
        tm.CategoricalMatrix(np.array([0, -1]), categories=np.arange(2))


def test_device_tensors_stay_on_device(zoo):
    import torch

    S = zoo["split"]
    d = torch.arange(1.0, 7.0, device="cuda", dtype=torch.float64)
    H = S.sandwich(d)
    assert isinstance(H, torch.Tensor) and H.is_cuda and H.dtype == torch.float64
    A = S.toarray()
    np.testing.assert_allclose(H.cpu().numpy(), A.T @ (d.cpu().numpy()[:, None] * A), atol=1e-12)
    v = torch.ones(S.shape[1], device="cuda", dtype=torch.float64)
    assert S.matvec(v).is_cuda and S.transpose_matvec(d).is_cuda
