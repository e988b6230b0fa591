"""The tcgen05 / TMEM / TMA weighted-SYRK kernel (dense_tc.cu) against the oracle and against
the CUDA-core kernel, forced through ``tm_set_dense_f32_mode``."""

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


@pytest.fixture
def force_mode():
    from tabmat_b200._lib import lib

    def set_mode(m):
        lib.tm_set_dense_f32_mode(m)

    yield set_mode
    lib.tm_set_dense_f32_mode(0)


@pytest.mark.parametrize("n,p", [(1, 8), (31, 32), (32, 64), (33, 128), (1000, 132), (4097, 256),
                                 (300_001, 256), (70_000, 96), (5000, 200), (257, 12)])
def test_tcgen05_syrk_matches_oracle(n, p, force_mode):
    from tabmat_b200._lib import lib
    from tests.gpu_runner import run_cuda

    if not lib.tm_has_tcgen05():
        pytest.fail("tcgen05 path unavailable on this device (expected sm_100)")
    rng = np.random.default_rng(n + p)
    X = rng.standard_normal((n, p)).astype(np.float32)
    d = rng.standard_normal(n).astype(np.float32)
    args = dict(X=X, d=d, rows=None, cols=None)
    force_mode(2)
    got = run_cuda("dense_sandwich", args)
    force_mode(1)
    core = run_cuda("dense_sandwich", args)
    Xd = X.astype(np.float64)
    ref = Xd.T @ (d.astype(np.float64)[:, None] * Xd)
    cases.assert_close(got, ref, np.float32, f"tcgen05 n={n} p={p}")
    cases.assert_close(core, ref, np.float32, f"cuda-core n={n} p={p}")
    assert np.array_equal(got, got.T), "output must be exactly symmetric"
    if n <= 5000:
        cases.assert_close(got, cases.run_oracle("dense_sandwich", args), np.float32, "vs oracle")


def test_tcgen05_syrk_row_restriction(force_mode):
    from tests.gpu_runner import run_cuda

    rng = np.random.default_rng(3)
    n, p = 10_000, 64
    X = rng.standard_normal((n, p)).astype(np.float32)
    d = rng.random(n).astype(np.float32)
    rows = np.sort(rng.choice(n, size=n // 3, replace=False)).astype(np.int32)
    force_mode(2)
    got = run_cuda("dense_sandwich", dict(X=X, d=d, rows=rows, cols=None))
    Xd = X[rows].astype(np.float64)
    cases.assert_close(got, Xd.T @ (d[rows].astype(np.float64)[:, None] * Xd), np.float32, "rows")


def test_tcgen05_linearity_at_scale(force_mode):
    """Full-size property check (no oracle at this size): sandwich is linear in d and equals
    the column sums of d*X^2 on the diagonal."""
    import torch

    from tabmat_b200.ext.dense import dense_sandwich

    n, p = 2_000_000, 256
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn((n, p), device="cuda", dtype=torch.float32, generator=g)
    d1 = torch.rand(n, device="cuda", generator=g)
    d2 = torch.rand(n, device="cuda", generator=g)
    force_mode(2)
    s1, s2, s12 = dense_sandwich(X, d1, None, None), dense_sandwich(X, d2, None, None), \
        dense_sandwich(X, d1 + d2, None, None)
    scale = s12.abs().max().item()
    assert (s1 + s2 - s12).abs().max().item() / scale < 1e-3
    diag = (d1[:, None].double() * X.double() ** 2).sum(0)
    assert ((s1.diagonal().double() - diag).abs().max() / diag.abs().max()).item() < 1e-3


@pytest.mark.parametrize("n,p", [(5003, 128), (70_001, 64), (999, 8), (40_000, 100)])
@pytest.mark.parametrize("with_rows", [False, True])
def test_onehot_fused_dense_and_small_cats(n, p, with_rows):
    """tm_dense_onehot_sandwich_f32: SYRK + one-hot MMAs for categoricals with few levels."""
    import ctypes as C

    import torch

    from oracle import c_oracle as orc
    from tabmat_b200._lib import check, lib

    rng = np.random.default_rng(n + p)
    X = rng.standard_normal((n, p)).astype(np.float32)
    d = rng.standard_normal(n).astype(np.float32)
    d[rng.random(n) < 0.05] = 0
    Ks, dfs = [10, 50, 200, 1, 33], [False, True, False, False, True]
    codes = [rng.integers(-1, K + int(df), size=n).astype(np.int32) for K, df in zip(Ks, dfs)]
    rows = np.sort(rng.choice(n, size=n // 2, replace=False)).astype(np.int32) if with_rows else None
    Xd, dd = torch.from_numpy(X).cuda(), torch.from_numpy(d).cuda()
    cd = [torch.from_numpy(c).cuda() for c in codes]
    rd = None if rows is None else torch.from_numpy(rows).cuda()
    out_dense = torch.empty((p, p), dtype=torch.float32, device="cuda")
    out_cat = torch.empty((sum(Ks), p), dtype=torch.float32, device="cuda")
    nc = len(Ks)
    codes_arr = (C.c_void_p * nc)(*[c.data_ptr() for c in cd])
    K_arr = (C.c_int64 * nc)(*Ks)
    df_arr = (C.c_int32 * nc)(*[int(f) for f in dfs])
    check(lib.tm_dense_onehot_sandwich_f32(
        Xd.data_ptr(), n, p, dd.data_ptr(), None if rd is None else rd.data_ptr(),
        0 if rd is None else len(rows), nc, C.cast(codes_arr, C.c_void_p),
        C.cast(K_arr, C.c_void_p), C.cast(df_arr, C.c_void_p), out_dense.data_ptr(),
        out_cat.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    cases.assert_close(out_dense.cpu().numpy(), orc.dense_sandwich(X, d, rows, None), np.float32,
                       "dense self")
    got = out_cat.cpu().numpy()
    o = 0
    for c, K, df in zip(codes, Ks, dfs):
        ref = orc.cat_dense_sandwich(c, K, d, X, rows, None, df)
        cases.assert_close(got[o:o + K], ref, np.float32, f"one-hot cat K={K}")
        o += K


@pytest.mark.parametrize("n,p", [(4, 8), (32, 64), (36, 128), (1000, 132), (4100, 256),
                                 (300_004, 256), (70_000, 96), (5000, 37), (260, 12)])
def test_tcgen05_syrk_column_major(n, p, force_mode):
    """F-ordered X (dense_helpers-tmpl.cpp:266-308 has a C and an F variant): the TMA box lands
    K-major in shared memory, no transpose; needs n % 4 == 0 (TMA pitch), any 8 <= p <= 256."""
    from tests.gpu_runner import run_cuda

    rng = np.random.default_rng(n + p)
    X = np.asfortranarray(rng.standard_normal((n, p)).astype(np.float32))
    d = rng.standard_normal(n).astype(np.float32)
    rows = np.sort(rng.choice(n, size=max(1, n // 2), replace=False)).astype(np.int32)
    Xd = X.astype(np.float64)
    for r in (None, rows):
        force_mode(2)
        got = run_cuda("dense_sandwich", dict(X=X, d=d, rows=r, cols=None))
        Xr, dr = (Xd, d) if r is None else (Xd[r], d[r])
        ref = Xr.T @ (dr.astype(np.float64)[:, None] * Xr)
        cases.assert_close(got, ref, np.float32, f"tcgen05 F-order n={n} p={p} rows={r is not None}")
        assert np.array_equal(got, got.T)


@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("frac", [0.9, 0.5, 0.25, 0.05])
def test_tcgen05_syrk_column_selection(order, frac, force_mode):
    """`cols` (dense.pyx:19-44): full SYRK on the tensor cores + selection while the selection
    keeps >= 1/5 of the columns, the CUDA-core kernel below that; both against float64."""
    from tests.gpu_runner import run_cuda

    rng = np.random.default_rng(int(frac * 100))
    n, p = 20_000, 128
    X = rng.standard_normal((n, p)).astype(np.float32)
    if order == "F":
        X = np.asfortranarray(X)
    d = rng.random(n).astype(np.float32)
    cols = np.sort(rng.choice(p, size=max(1, int(p * frac)), replace=False)).astype(np.int32)
    rows = np.sort(rng.choice(n, size=n // 3, replace=False)).astype(np.int32)
    force_mode(0)
    for r in (None, rows):
        got = run_cuda("dense_sandwich", dict(X=X, d=d, rows=r, cols=cols))
        Xs = X.astype(np.float64)[:, cols]
        Xr, dr = (Xs, d) if r is None else (Xs[r], d[r])
        cases.assert_close(got, Xr.T @ (dr.astype(np.float64)[:, None] * Xr), np.float32,
                           f"cols {order} frac={frac}")
    if frac >= 0.2:
        force_mode(2)   # must be accepted: the selection is eligible for the tensor path
        run_cuda("dense_sandwich", dict(X=X, d=d, rows=None, cols=cols))


@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("n,p", [(50_000, 64), (20_000, 256), (4099 * 4, 128)])
def test_tf32x3_restores_fp32_accuracy(n, p, order, force_mode):
    """tm_set_dense_f32_mode(3): hi/lo operand split, three MMAs per product.  On an
    ill-conditioned X (columns with a large common offset, so X^T D X has big entries whose small
    differences matter) single-pass TF32 is ~10x less accurate than the fp32 CUDA-core kernel (the
    reference's f32 kernels use full-precision FMAs); 3xTF32 must come within a small factor of
    it.  The three errors are printed: this is the record of the TF32 margin."""
    from tests.gpu_runner import run_cuda

    rng = np.random.default_rng(n + p)
    X = (30.0 + rng.standard_normal((n, p))).astype(np.float32)
    if order == "F":
        X = np.asfortranarray(X)
    d = rng.random(n).astype(np.float32)
    Xd = X.astype(np.float64)
    ref = Xd.T @ (d.astype(np.float64)[:, None] * Xd)
    # what matters downstream: the centred second moments = small differences of big entries
    mu = (d.astype(np.float64) @ Xd) / d.sum()
    cen = ref - d.sum() * np.outer(mu, mu)
    err = {}
    for mode, name in ((2, "tf32"), (3, "tf32x3"), (1, "fp32 cuda cores")):
        force_mode(mode)
        got = run_cuda("dense_sandwich", dict(X=X, d=d, rows=None, cols=None)).astype(np.float64)
        err[name] = (np.abs(got - ref).max() / np.abs(ref).max(),
                     np.abs((got - d.sum() * np.outer(mu, mu)) - cen).max() / np.abs(cen).max())
    print(f"\nn={n} p={p} {order}: normwise / centred-moment error  " +
          "  ".join(f"{k}: {a:.2e} / {b:.2e}" for k, (a, b) in err.items()))
    # measured on B200 (profiles/pytest_gpu_r2b_tf32x3.txt): tf32 5.6e-6..8.0e-6 normwise and
    # 5e-3..7e-3 on the centred moments, tf32x3 2.2e-6..2.6e-6 / 1.9e-3..2.3e-3, fp32 CUDA cores
    # 5e-7..9e-7 / 5e-4..8e-4: the split operands remove the operand-rounding error, what is left
    # is the tensor core's own accumulation (not a full-precision fp32 adder)
    assert err["tf32"][0] <= 1e-3
    assert err["tf32x3"][0] <= 5e-6 and err["tf32x3"][0] <= 8 * max(err["fp32 cuda cores"][0], 1e-7)
    assert err["tf32x3"][1] <= 8 * max(err["fp32 cuda cores"][1], 1e-6)
    assert err["tf32x3"][0] < err["tf32"][0] and err["tf32x3"][1] < err["tf32"][1]


def test_tf32x3_with_onehot_blocks_in_a_split_matrix(force_mode):
    """Mode 3 through the fused SplitMatrix pass: the one-hot columns are exact 0/1 and must be
    added in two of the three sub-passes only."""
    import scipy.sparse as sps

    import tabmat_b200 as tm

    rng = np.random.default_rng(8)
    n = 20_000
    X = (10.0 + rng.standard_normal((n, 32))).astype(np.float32)
    c1 = rng.integers(0, 12, size=n).astype(np.int32)
    c2 = rng.integers(0, 500, size=n).astype(np.int32)
    A = sps.random(n, 40, density=0.05, random_state=rng, format="csc").astype(np.float32)
    S = tm.SplitMatrix([tm.DenseMatrix(X), tm.SparseMatrix(A),
                        tm.CategoricalMatrix(c1, categories=np.arange(12), dtype=np.float32),
                        tm.CategoricalMatrix(c2, categories=np.arange(500), dtype=np.float32)])
    full = np.hstack([X, A.toarray(), np.eye(12)[c1], np.eye(500)[c2]]).astype(np.float64)
    d = rng.random(n).astype(np.float32)
    ref = (full * d.astype(np.float64)[:, None]).T @ full
    errs = {}
    for mode in (0, 3):
        force_mode(mode)
        got = S.sandwich(d)
        errs[mode] = np.abs(got - ref).max() / np.abs(ref).max()
        blk = np.abs(got[72:84, :32] - ref[72:84, :32]).max() / np.abs(ref[72:84, :32]).max()
        errs[(mode, "onehot x dense")] = blk
    print("\nsplit f32, tf32 vs tf32x3:", {str(k): f"{v:.2e}" for k, v in errs.items()})
    assert errs[0] <= 1e-3 and errs[3] <= 5e-6 and errs[(3, "onehot x dense")] <= 5e-6


@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("n,p", [(5000, 260), (4100, 384), (3000, 512), (2500, 700), (36, 1028)])
def test_tcgen05_syrk_wide_matrices_by_panels(n, p, order, force_mode):
    """p > 256: passes over pairs of 128-column panels (diagonal pairs give three tiles, every
    other cross tile is its own pass), ragged last panel, odd panel counts; C and F order
    (dense_helpers-tmpl.cpp:266-308 handles any p in both orders)."""
    from tests.gpu_runner import run_cuda

    rng = np.random.default_rng(n + p)
    X = rng.standard_normal((n, p)).astype(np.float32)
    if order == "F":
        X = np.asfortranarray(X)
    d = rng.standard_normal(n).astype(np.float32)
    rows = np.sort(rng.choice(n, size=max(1, n // 2), replace=False)).astype(np.int32)
    cols = np.sort(rng.choice(p, size=p // 2, replace=False)).astype(np.int32)
    Xd = X.astype(np.float64)
    force_mode(2)
    for r, c in ((None, None), (rows, None), (None, cols)):
        got = run_cuda("dense_sandwich", dict(X=X, d=d, rows=r, cols=c))
        Xs = Xd if c is None else Xd[:, c]
        Xr, dr = (Xs, d) if r is None else (Xs[r], d[r])
        ref = Xr.T @ (dr.astype(np.float64)[:, None] * Xr)
        cases.assert_close(got, ref, np.float32, f"panels n={n} p={p} {order}")
        assert np.array_equal(got, got.T)
    force_mode(1)
    core = run_cuda("dense_sandwich", dict(X=X, d=d, rows=None, cols=None))
    cases.assert_close(core, Xd.T @ (d.astype(np.float64)[:, None] * Xd), np.float32, "cuda-core")


def test_tf32_rounding_modes_agree():
    """fp32 -> tf32 of the MMA operands: the cvt.rna.tf32.f32 instruction (mode 0), integer
    add + mask (1) and integer add alone (2, the tensor core ignores the low 13 bits) give the
    same sandwich up to the order of the atomic adds; X is offset so that a truncating
    conversion (no rounding at all) would show as a bias of ~2e-4."""
    import tabmat_b200 as tm

    lib = tm._lib.lib
    if not lib.tm_has_tcgen05():
        pytest.skip("needs sm_100")
    rng = np.random.default_rng(3)
    n, p = 200_000, 128
    X = (1.0 + rng.random((n, p))).astype(np.float32)
    d = rng.random(n).astype(np.float32)
    D = tm.DenseMatrix(X)
    ref = (X.astype(np.float64) * d[:, None].astype(np.float64)).T @ X.astype(np.float64)
    res = {}
    try:
        for mode in (0, 1, 2):
            lib.tm_set_tc_round_mode(mode)
            res[mode] = D.sandwich(d)
    finally:
        lib.tm_set_tc_round_mode(-1)
    scale = np.abs(ref).max()
    for mode in (1, 2):
        assert np.abs(res[mode] - res[0]).max() / scale < 2e-6, mode
    for mode in (0, 1, 2):
        # unbiased rounding: far below the 2^-11 relative bias of a truncating conversion
        assert abs((res[mode] - ref).mean()) / scale < 2e-5, mode


@pytest.mark.parametrize("p,order", [(64, "C"), (256, "C"), (128, "F")])
def test_accumulator_chains_are_cut(p, order):
    """The tensor core's fp32 accumulator truncates at every K = 8 step: without a limit on the
    chain length the sandwich of 2e6 rows of all-positive data is ~5e-5 too small (and 1e-3 at
    4e7 rows).  tm_set_tc_flush_steps bounds the chain: the accumulators are drained into the
    result every so many steps.  Many drains per CTA must give the same result, more accurately."""
    import torch

    import tabmat_b200 as tm

    lib = tm._lib.lib
    if not lib.tm_has_tcgen05():
        pytest.skip("needs sm_100")
    n = 2_000_000
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(p)
    X = 1.0 + torch.rand((n, p), device=dev, dtype=torch.float32, generator=g)
    if order == "F":
        X = X.t().contiguous().t()
    d = torch.rand(n, device=dev, dtype=torch.float32, generator=g)
    ref = torch.zeros((p, p), device=dev, dtype=torch.float64)
    for lo in range(0, n, 250_000):
        Xc = X[lo:lo + 250_000].double()
        ref += Xc.t() @ (Xc * d[lo:lo + 250_000].double()[:, None])
    D = tm.DenseMatrix(X)
    err = {}
    try:
        for steps in (0, 128, 2048):
            lib.tm_set_tc_flush_steps(steps)
            got = D.sandwich(d).double()
            err[steps] = float(((got - ref).abs().max() / ref.abs().max()).item())
    finally:
        lib.tm_set_tc_flush_steps(-1)
    print(f"\np={p} {order}: normwise error by chain length", {k: f"{v:.2e}" for k, v in err.items()})
    assert err[0] < 1e-3 and err[128] < 1e-3 and err[2048] < 1e-3
    assert err[128] < err[0] / 3, err      # ~13 drains per CTA: the truncation loss is gone
    assert err[2048] <= err[0] * 1.05, err
