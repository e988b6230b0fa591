"""Host-side logic of the row-sorted storage and of device-side sparse row indexing, on CPU
tensors (no kernels run here: the permutation kernels are replaced by their torch definition and
the wrapped matrix by a dense stand-in; the GPU versions are in tests/test_gpu_row_order.py)."""

import numpy as np
import pytest
import scipy.sparse as sps
import torch

import tabmat_b200 as tm
from tabmat_b200 import _dev, row_order
from tabmat_b200.matrix_base import MatrixBase


class _DenseStandIn(MatrixBase):
    """What RowSortedMatrix needs from the SplitMatrix it wraps, over a dense CPU array."""

    def __init__(self, F):
        self.F = torch.as_tensor(F, dtype=torch.float64)
        self.shape = tuple(self.F.shape)
        self.dtype = np.dtype(np.float64)
        self.indices = [np.arange(self.shape[1])]
        self.matrices = []

    def _sub(self, rows, cols):
        F = self.F if rows is None else self.F[rows.to(torch.int64)]
        return F if cols is None else F[:, torch.as_tensor(np.asarray(cols), dtype=torch.int64)]

    def _sandwich_dev(self, d_t, rows_t, cols):
        F = self._sub(rows_t, cols)
        dd = d_t if rows_t is None else d_t[rows_t.to(torch.int64)]
        return (F * dd[:, None]).T @ F

    def sandwich(self, d, rows=None, cols=None):
        return self._sandwich_dev(d, rows, cols)

    def matvec(self, v, cols=None, out=None):
        return self._sub(None, cols) @ (v if cols is None else v[np.asarray(cols)])

    def transpose_matvec(self, v, rows=None, cols=None, out=None):
        vv = v if rows is None else v[rows.to(torch.int64)]
        res = self._sub(rows, cols).T @ vv
        if out is not None:
            out += res
            return out
        return res

    def _get_col_stds(self, weights, col_means):
        return torch.sqrt(((self.F - col_means) ** 2 * weights[:, None]).sum(0))

    def getcol(self, i):
        raise NotImplementedError

    def astype(self, dtype, order="K", casting="unsafe", copy=True):
        return self

    def __getitem__(self, item):
        return _DenseStandIn(self.F[item])

    def toarray(self):
        return self.F.numpy()

    def get_names(self, *a, **k):
        return []

    def set_names(self, *a, **k):
        pass


@pytest.fixture
def cpu_plumbing(monkeypatch):
    """CPU tensors through the host classes: the torch definitions of the two permutation
    kernels, no device requirement."""
    monkeypatch.setattr(_dev, "is_dev", lambda x: isinstance(x, torch.Tensor))
    monkeypatch.setattr(_dev, "require_cuda", lambda: torch.device("cpu"))
    monkeypatch.setattr(row_order.RowSortedMatrix, "_gather",
                        lambda self, v: v.index_select(0, self._perm.to(torch.int64)))

    def scatter(self, y):
        out = torch.empty_like(y)
        out[self._perm.to(torch.int64)] = y
        return out

    monkeypatch.setattr(row_order.RowSortedMatrix, "_scatter", scatter)


def test_row_sorted_wrapper_maps_vectors_and_rows(cpu_plumbing):
    rng = np.random.default_rng(0)
    n, p = 50, 6
    F = rng.standard_normal((n, p))
    perm = torch.from_numpy(rng.permutation(n))
    S = tm.RowSortedMatrix(_DenseStandIn(F[perm.numpy()]), perm)
    d = torch.from_numpy(rng.standard_normal(n))
    rows = np.sort(rng.choice(n, size=17, replace=False))
    cols = np.array([0, 2, 5])
    np.testing.assert_allclose(S.sandwich(d), (F * d.numpy()[:, None]).T @ F)
    Fr = F[rows][:, cols]
    np.testing.assert_allclose(S.sandwich(d, rows, cols), (Fr * d.numpy()[rows, None]).T @ Fr)
    v = torch.from_numpy(rng.standard_normal(p))
    np.testing.assert_allclose(S.matvec(v), F @ v.numpy())
    out = torch.ones(n, dtype=torch.float64)
    assert S.matvec(v, out=out) is out
    np.testing.assert_allclose(out, 1 + F @ v.numpy())
    w = torch.from_numpy(rng.standard_normal(n))
    np.testing.assert_allclose(S.transpose_matvec(w), F.T @ w.numpy())
    np.testing.assert_allclose(S.transpose_matvec(w, rows, cols), Fr.T @ w.numpy()[rows])
    np.testing.assert_allclose(S.toarray(), F)
    np.testing.assert_allclose(S[rows].toarray(), F[rows])
    np.testing.assert_allclose(S[3:20:2].toarray(), F[3:20:2])
    np.testing.assert_allclose(S.unsorted().toarray(), F)
    ww = torch.from_numpy(rng.random(n))
    means = S._get_col_means(ww)
    np.testing.assert_allclose(means, F.T @ ww.numpy())
    np.testing.assert_allclose(S._get_col_stds(ww, means),
                               np.sqrt((((F - means.numpy()) ** 2) * ww.numpy()[:, None]).sum(0)))


def test_sort_keys_and_permutation():
    class Cat:  # the two attributes sort_permutation reads
        def __init__(self, codes, K):
            self._codes = torch.as_tensor(codes, dtype=torch.int32)
            self.categories = np.arange(K)
            self.shape = (len(codes), K)

    a = np.array([2, 0, -1, 2, 1, 0, 2], dtype=np.int32)
    b = np.array([1, 1, 0, 0, 2, 0, 0], dtype=np.int32)
    perm = row_order.sort_permutation([Cat(a, 3), Cat(b, 3)], [0, 1]).numpy()
    # lexicographic by (a, b), missing last, stable
    np.testing.assert_array_equal(perm, [5, 1, 4, 3, 6, 0, 2])
    assert row_order.sort_permutation([Cat(a, 3)], []) is None


def test_sparse_row_take_matches_scipy(cpu_plumbing):
    rng = np.random.default_rng(1)
    A = sps.random(60, 9, density=0.2, random_state=rng, format="csr").astype(np.float64)
    A.sort_indices()
    M = tm.SparseMatrix.from_device_csr(torch.from_numpy(A.data), torch.from_numpy(A.indices),
                                        torch.from_numpy(A.indptr), A.shape)
    idx = rng.integers(-60, 60, size=100)
    for key in (idx, slice(5, 50, 4), rng.random(60) < 0.4, torch.from_numpy(idx)):
        sub = M._take_rows_dev(key)
        k = key.numpy() if isinstance(key, torch.Tensor) else key
        ref = A[k]
        got = sps.csr_matrix((sub._csr.data.numpy(), sub._csr.indices.numpy(),
                              sub._csr.indptr.numpy()), shape=sub.shape)
        assert sub.shape == ref.shape
        np.testing.assert_array_equal(got.toarray(), ref.toarray())
        # the CSC copy built alongside describes the same matrix
        csc = sps.csc_matrix((sub._csc.data.numpy(), sub._csc.indices.numpy(),
                              sub._csc.indptr.numpy()), shape=sub.shape)
        np.testing.assert_array_equal(csc.toarray(), ref.toarray())


def test_choose_sort_blocks_prefers_wide_categoricals():
    class Cat(tm.CategoricalMatrix):
        def __init__(self, K):  # shape only
            self.shape = (10, K)

    class Other:
        shape = (10, 5)

    mats = [Other(), Cat(10), Cat(2000), Cat(300), Cat(1000)]
    assert row_order.choose_sort_blocks(mats) == [2, 4]          # two widest blocks > 256 levels
    assert row_order.choose_sort_blocks(mats, max_keys=1) == [2]
    assert row_order.choose_sort_blocks([Other(), Cat(10), Cat(50)]) == [2]  # else the widest
    assert row_order.choose_sort_blocks([Other()]) == []


def test_column_runs_of_a_split_matrix():
    """_column_runs: maximal runs of result columns owned by the dense block / by the others
    (drives the two-phase host copy of sandwich_into)."""
    from tabmat_b200.split_matrix import SplitMatrix

    class D(tm.DenseMatrix):
        def __init__(self):
            pass

    class C(tm.CategoricalMatrix):
        def __init__(self):
            pass

    S = SplitMatrix.__new__(SplitMatrix)
    S.matrices = [D(), C(), C()]
    S.shape = (3, 10)
    S.indices = [np.array([3, 4, 5]), np.array([0, 1, 2, 9]), np.array([6, 7, 8])]
    assert S._column_runs() == [(0, 3, False), (3, 6, True), (6, 10, False)]
    S2 = SplitMatrix.__new__(SplitMatrix)
    S2.matrices = [D(), C()]
    S2.shape = (3, 5)
    S2.indices = [np.array([0, 1]), np.array([2, 3, 4])]
    assert S2._column_runs() == [(0, 2, True), (2, 5, False)]
