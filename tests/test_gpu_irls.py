"""The fused IRLS pass: ``X.sandwich_and_transpose_matvec(d, v)`` = ``(X.sandwich(d),
X.transpose_matvec(v))`` of the reference's two separate calls (matrix_base.py:15-77,
split_matrix.py:324-356 and :422-460) from one pass over the dense block, the one-pass
``StandardizedMatrix.sandwich`` built on it (standardized_mat.py:123-172) and
``tabmat_b200.irls_step``.  The score X.T v is accumulated in fp32 (never TF32): it is checked
element-wise against float64."""

import numpy as np
import pytest
import scipy.sparse as sps

from tests import cases

pytestmark = pytest.mark.gpu


def _split(dt, n, p_dense, order="C", seed=0):
    import tabmat_b200 as tm

    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, p_dense)).astype(dt)
    A = sps.random(n, 30, density=0.05, random_state=rng, format="csc").astype(dt)
    c1 = rng.integers(0, 12, size=n).astype(np.int32)
    c2 = rng.integers(0, 400, size=n).astype(np.int32)
    Xd = tm.DenseMatrix(X)
    if order == "F":
        import torch

        Xd = tm.DenseMatrix(torch.from_numpy(np.ascontiguousarray(X.T)).cuda().t())
    S = tm.SplitMatrix([Xd, tm.SparseMatrix(A),
                        tm.CategoricalMatrix(c1, categories=np.arange(12), dtype=dt),
                        tm.CategoricalMatrix(c2, categories=np.arange(400), dtype=dt, drop_first=True)])
    full = np.hstack([X, A.toarray(), np.eye(12)[c1], np.eye(400)[c2][:, 1:]]).astype(np.float64)
    return S, full, rng


@pytest.mark.parametrize("suf", ["f32", "f64"])
@pytest.mark.parametrize("n,p_dense,order", [(6007, 64, "C"), (33, 8, "C"), (5000, 128, "C"),
                                              (4000, 10, "C"), (4000, 16, "F")])
def test_fused_hessian_and_score(suf, n, p_dense, order):
    import tabmat_b200 as tm

    dt = cases.DTYPES[suf]
    S, full, rng = _split(dt, n, p_dense, order, seed=n + p_dense)
    d = rng.random(n).astype(dt)
    v = rng.standard_normal(n).astype(dt)
    p = S.shape[1]
    rows = np.sort(rng.choice(n, size=n // 2, replace=False)).astype(np.int32)
    cols = np.sort(rng.choice(p, size=p // 2, replace=False)).astype(np.int32)
    for M in (S, tm.RowSortedMatrix.from_split(S)):
        for r in (None, rows):
            for c in (None, cols):
                F = full if r is None else full[r]
                F = F if c is None else F[:, c]
                dd, vv = (d, v) if r is None else (d[r], v[r])
                H, g = M.sandwich_and_transpose_matvec(d, v, r, c)
                cases.assert_close(H, (F * dd.astype(np.float64)[:, None]).T @ F, dt, "fused Hessian")
                ref_g = F.T @ vv.astype(np.float64)
                tol = 2e-5 if dt == np.float32 else 1e-12
                bound = (np.abs(F).T @ np.abs(vv.astype(np.float64)))   # sum of |terms| per column
                assert (np.abs(g - ref_g) <= tol * np.maximum(bound, 1e-30)).all(), \
                    f"score not at {dt.__name__} accuracy"
                # identical to the two separate calls (up to summation order)
                cases.assert_close(g, M.transpose_matvec(v, r, c), dt, "fused score vs separate")


def test_default_implementation_for_leaf_matrices():
    import tabmat_b200 as tm

    rng = np.random.default_rng(0)
    X = rng.standard_normal((500, 12))
    d, v = rng.random(500), rng.standard_normal(500)
    for M in (tm.DenseMatrix(X), tm.SparseMatrix(sps.csc_matrix(np.where(X > 1, X, 0)))):
        H, g = M.sandwich_and_transpose_matvec(d, v)
        cases.assert_close(H, M.sandwich(d), np.float64, "leaf fused Hessian")
        cases.assert_close(g, M.transpose_matvec(v), np.float64, "leaf fused score")


@pytest.mark.parametrize("suf", ["f32", "f64"])
@pytest.mark.parametrize("center,scale", [(True, True), (True, False), (False, True)])
def test_standardized_sandwich_one_pass(suf, center, scale):
    """StandardizedMatrix.sandwich over a SplitMatrix: inner sandwich + inner.T d from one pass,
    rank-1 epilogue in one kernel; against the dense standardized recomputation."""
    dt = cases.DTYPES[suf]
    n = 5003
    S, full, rng = _split(dt, n, 32, seed=3)
    w = rng.random(n)
    w /= w.sum()
    Z, means, stds = S.standardize(w.astype(dt), center, scale)
    mu = w @ full
    sd = np.sqrt(w @ (full - mu) ** 2)
    mult = np.where(np.abs(sd) < 1e-7, 1.0, 1 / np.where(sd == 0, 1, sd)) if scale else np.ones(full.shape[1])
    Xs = (full - (mu if center else 0)) * mult
    d = rng.random(n).astype(dt)
    p = S.shape[1]
    rows = np.sort(rng.choice(n, size=n // 2, replace=False)).astype(np.int32)
    cols = np.sort(rng.choice(p, size=p // 3, replace=False)).astype(np.int32)
    for r in (None, rows):
        for c in (None, cols):
            F = Xs if r is None else Xs[r]
            F = F if c is None else F[:, c]
            dd = d if r is None else d[r]
            got = Z.sandwich(d, r, c)
            assert got.dtype == dt
            ref = (F * dd.astype(np.float64)[:, None]).T @ F
            scale_ = np.abs(ref).max()
            # f32: the rank-1 corrections cancel large terms (the reference has the same loss)
            assert np.abs(got - ref).max() / scale_ <= (5e-3 if dt == np.float32 else 1e-8)


def test_irls_step_logistic_matches_float64_newton():
    """Three Newton / IRLS steps of a logistic regression driven by irls_step stay on the
    float64 numpy iteration."""
    import torch

    import tabmat_b200 as tm

    n = 20_000
    S, full, rng = _split(np.float64, n, 16, seed=11)
    p = S.shape[1]
    beta_true = rng.standard_normal(p) * 0.3
    y = (rng.random(n) < 1 / (1 + np.exp(-(full @ beta_true)))).astype(np.float64)
    y_t = torch.from_numpy(y).cuda()

    def weights_fn(eta):
        mu = torch.sigmoid(eta)
        return mu * (1 - mu), y_t - mu

    beta = np.zeros(p)
    beta_ref = np.zeros(p)
    lam = 1e-3 * np.eye(p)
    for _ in range(3):
        H, g, eta = tm.irls_step(S, beta, weights_fn)
        assert H.is_cuda and g.is_cuda and eta.is_cuda
        beta = beta + np.linalg.solve(H.cpu().numpy() + lam, g.cpu().numpy())
        mu = 1 / (1 + np.exp(-(full @ beta_ref)))
        Href = (full * (mu * (1 - mu))[:, None]).T @ full
        beta_ref = beta_ref + np.linalg.solve(Href + lam, full.T @ (y - mu))
        np.testing.assert_allclose(beta, beta_ref, rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_irls_step_in_stored_order(dt):
    """irls_step(..., stored_order=True) on a RowSortedMatrix: the response is permuted once
    (X.to_stored_order), eta / d / v stay in the stored row order, and the Hessian and score are
    those of the ordinary call (which permutes eta, d and v on every step); a `rows` restriction
    stays in the caller's numbering."""
    import torch

    import tabmat_b200 as tm

    n = 30_011
    S, full, rng = _split(dt, n, 16, seed=5)
    R = tm.RowSortedMatrix.from_split(S)
    p = S.shape[1]
    y = (rng.random(n) < 0.4).astype(dt)
    off = (0.1 * rng.standard_normal(n)).astype(dt)
    beta = (0.2 * rng.standard_normal(p)).astype(dt)
    rows = np.sort(rng.choice(n, size=n // 2, replace=False)).astype(np.int32)
    y_t, off_t = torch.from_numpy(y).cuda(), torch.from_numpy(off).cuda()
    y_s, off_s = R.to_stored_order(y), R.to_stored_order(off)
    assert torch.equal(R.from_stored_order(y_s), y_t)

    def fn_for(resp):
        def weights_fn(eta):
            mu = torch.sigmoid(eta)
            return mu * (1 - mu), resp - mu
        return weights_fn

    tol = 2e-4 if dt == np.float32 else 1e-10
    for r in (None, rows):
        H0, g0, eta0 = tm.irls_step(R, beta, fn_for(y_t), offset=off_t, rows=r)
        H1, g1, eta1 = tm.irls_step(R, beta, fn_for(y_s), offset=off_s, rows=r, stored_order=True)
        assert float((H1 - H0).abs().max() / H0.abs().max()) <= tol
        assert float((g1 - g0).abs().max() / g0.abs().max()) <= tol
        assert float((R.from_stored_order(eta1) - eta0).abs().max()) <= tol * 10
    # a plain SplitMatrix has one row order: the flag changes nothing
    H2, g2, _ = tm.irls_step(S, beta, fn_for(y_t), offset=off_t, stored_order=True)
    H0, g0, _ = tm.irls_step(S, beta, fn_for(y_t), offset=off_t)
    assert float((H2 - H0).abs().max() / H0.abs().max()) <= tol
