"""Full-size property checks (no oracle finishes at these sizes): identities that tie the
sandwich kernels to the independently implemented matvec kernels, symmetry, linearity and
exact counts, at the BASELINE.json shapes (C3, C4) and at 1/10 of C5."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _quadratic_form_check(X, d, p, tol):
    """v^T (X^T D X) v == sum_k d_k (X v)_k^2, with X v from the matvec kernels."""
    g = torch.Generator(device="cuda").manual_seed(3)
    H = X.sandwich(d)
    H = H if isinstance(H, torch.Tensor) else torch.as_tensor(H.toarray(), device="cuda")
    assert torch.equal(H, H.t()), "sandwich must be exactly symmetric"
    for _ in range(3):
        v = torch.randn(p, device="cuda", dtype=d.dtype, generator=g)
        Xv = X.matvec(v).double()
        rhs = (d.double() * Xv * Xv).sum().item()
        lhs = (v.double() @ (H.double() @ v.double())).item()
        assert abs(lhs - rhs) <= tol * max(abs(rhs), 1.0), (lhs, rhs)


def test_c5_split_sandwich_identity_at_4e6_rows():
    """C5 layout (128 dense + 3x1000 CSC + cat{10,50,200,1000,2000}) at n = 4e6, f32: the
    tcgen05 / scatter / index passes against the matvec kernels, plus the transpose_matvec
    identity  X^T d == column sums of diag(d) X."""
    import bench

    X, d, _, _ = bench.C5(None).device_matrix(4_000_000, seed=7, device=torch.device("cuda", 0))
    p = X.shape[1]
    assert p == 6388
    _quadratic_form_check(X, d, p, 2e-3)
    # linearity in d
    d2 = torch.rand_like(d)
    H1, H2, H12 = X.sandwich(d), X.sandwich(d2), X.sandwich(d + d2)
    scale = H12.abs().max().item()
    assert (H1 + H2 - H12).abs().max().item() / scale < 1e-3
    # diagonal of the categorical self blocks == transpose_matvec(d) on those columns
    tmv = X.transpose_matvec(d).double()
    cat_cols = torch.cat([torch.as_tensor(idx, device="cuda") for idx, m in
                          zip(X.indices, X.matrices) if type(m).__name__ == "CategoricalMatrix"])
    diag = H1.diagonal()[cat_cols]
    assert ((diag - tmv[cat_cols]).abs().max() / tmv[cat_cols].abs().max()).item() < 1e-4


def test_c3_categorical_counts_exact_at_1e7():
    import tabmat_b200 as tm

    g = torch.Generator(device="cuda").manual_seed(2)
    n, K = 10_000_000, 2000
    codes = torch.randint(0, K, (n,), device="cuda", dtype=torch.int32, generator=g)
    C = tm.CategoricalMatrix(codes, categories=np.arange(K), dtype=np.float32)
    ones = torch.ones(n, device="cuda", dtype=torch.float32)
    diag = C.sandwich(ones)
    assert torch.equal(diag, torch.bincount(codes.long(), minlength=K).float())  # bit-exact
    assert torch.equal(C.transpose_matvec(ones), diag)
    v = torch.arange(K, device="cuda", dtype=torch.float32)
    assert torch.equal(C.matvec(v), codes.float())


def test_c4_sparse_sandwich_identity_at_1e7():
    import tabmat_b200 as tm

    g = torch.Generator(device="cuda").manual_seed(3)
    n, p, nnz_t = 10_000_000, 5000, 50_000_000
    key = torch.unique(torch.randint(0, n, (nnz_t,), device="cuda", generator=g) * p
                       + torch.randint(0, p, (nnz_t,), device="cuda", generator=g))
    rows = torch.div(key, p, rounding_mode="floor")
    cols = (key - rows * p).to(torch.int32)
    indptr = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    indptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
    vals = torch.randn(cols.numel(), device="cuda", dtype=torch.float64, generator=g)
    A = tm.SparseMatrix.from_device_csr(vals, cols, indptr.to(torch.int32), (n, p))
    d = torch.rand(n, device="cuda", dtype=torch.float64, generator=g)
    _quadratic_form_check(A, d, p, 1e-9)
    # cross term against a dense block: v^T (A^T D B) w == sum d (A v)(B w)
    B = tm.DenseMatrix(torch.randn((n, 128), device="cuda", dtype=torch.float64, generator=g))
    Hc = A._cross_sandwich(B, d, None, None, None)
    v = torch.randn(p, device="cuda", dtype=torch.float64, generator=g)
    w = torch.randn(128, device="cuda", dtype=torch.float64, generator=g)
    lhs = (v @ (Hc @ w)).item()
    rhs = (d * A.matvec(v) * B.matvec(w)).sum().item()
    assert abs(lhs - rhs) <= 1e-9 * abs(rhs)
