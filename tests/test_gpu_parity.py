"""GPU parity, boundary level: every C-ABI entry point against the oracle (C restatement)
and against the golden vectors produced by the reference itself, on identical seeded inputs.

Tolerance (north_star): normwise 1e-5 fp64 / 1e-3 fp32; categorical counts bit-exact."""

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu

ALL_CASES = list(cases.boundary_cases())


@pytest.mark.parametrize("name,kind,args", ALL_CASES, ids=[c[0] for c in ALL_CASES])
def test_cuda_matches_oracle_and_golden(name, kind, args, golden_boundary):
    from tests.gpu_runner import run_cuda

    got = run_cuda(kind, args)
    orc = cases.run_oracle(kind, args)
    assert got.dtype == orc.dtype
    cases.assert_close(got, orc, got.dtype, name + " vs oracle")
    cases.assert_close(got, golden_boundary[name], got.dtype, name + " vs reference golden")


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_cat_sandwich_counts_bit_exact(dt):
    """d == 1: the segmented reduction is an exact count (north_star: bit-exact)."""
    from tests.gpu_runner import run_cuda

    rng = np.random.default_rng(5)
    n, K = 1_000_003, 2000
    codes = rng.integers(0, K, size=n).astype(np.int32)
    got = run_cuda("cat_sandwich", dict(codes=codes, d=np.ones(n, dt), rows=None, K=K,
                                        drop_first=False))
    assert np.array_equal(got, np.bincount(codes, minlength=K).astype(dt))
    ci = rng.integers(0, 50, size=n).astype(np.int32)
    cj = rng.integers(0, 30, size=n).astype(np.int32)
    got = run_cuda("cat_cat_sandwich", dict(ic=ci, jc=cj, Ki=50, Kj=30, d=np.ones(n, dt),
                                            rows=None, i_drop_first=False, j_drop_first=False))
    ref = np.zeros((50, 30))
    np.add.at(ref, (ci, cj), 1)
    assert np.array_equal(got, ref.astype(dt))


@pytest.mark.parametrize("seed", range(3))
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_random_medium_shapes(seed, dt):
    """Medium shapes (oracle still finishes in seconds): ragged n, many tiles."""
    from tests.gpu_runner import run_cuda

    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(3000, 9000))
    I = cases.make_inputs(200 + seed, n, dt, p_dense=int(rng.integers(65, 200)),  # noqa: E741
                          p_sparse=int(rng.integers(100, 400)), Ki=int(rng.integers(20, 300)),
                          Kj=17, density=0.02)
    for kind, args in [
        ("dense_sandwich", dict(X=I["X"], d=I["d"], rows=None, cols=None)),
        ("dense_sandwich", dict(X=np.asfortranarray(I["X"]), d=I["d"], rows=I["rows"],
                                cols=I["cols_dense"])),
        ("sparse_sandwich", dict(A=I["A"], d=I["d"], rows=I["rows"], cols=None)),
        ("csr_dense_sandwich", dict(A=I["A"], B=I["X"], d=I["d"], rows=None, A_cols=None,
                                    B_cols=None)),
        ("cat_dense_sandwich", dict(codes=I["ci_missing"], K=I["Ki"], d=I["d"], Y=I["X"],
                                    rows=None, j_cols=None, drop_first=False)),
        ("cat_cat_sandwich", dict(ic=I["ci_missing"], jc=I["cj"], Ki=I["Ki"], Kj=I["Kj"],
                                  d=I["d"], rows=I["rows"], i_drop_first=False,
                                  j_drop_first=False)),
        ("cat_sparse_sandwich", dict(codes=I["ci"], K=I["Ki"], d=I["d"], A=I["A"], rows=None,
                                     s_cols=None, drop_first=False)),
        ("csr_matvec", dict(A=I["A"], v=I["v_sparse"], rows=None, cols=None)),
        ("csc_rmatvec", dict(A=I["A"], v=I["v_n"], rows=None, cols=None)),
        ("dense_rmatvec", dict(X=I["X"], v=I["v_n"], rows=None, cols=None)),
        ("dense_matvec", dict(X=I["X"], v=I["v_dense"], rows=None, cols=None)),
    ]:
        got = run_cuda(kind, args)
        cases.assert_close(got, cases.run_oracle(kind, args), dt, f"{kind} seed={seed}")


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("p", [8, 64, 128, 200])
@pytest.mark.parametrize("with_rows", [False, True])
def test_fused_dense_cross(dt, p, with_rows):
    """tm_dense_cross_sandwich (one pass over the dense block) == the per-pair oracle results,
    incl. replicated small tables, drop_first, missing codes and a row restriction."""
    import torch

    from oracle import c_oracle as orc
    from tabmat_b200.ext.split import dense_cross_sandwich
    from tests.gpu_runner import _csr, _dev, _i32

    if dt == np.float64 and p > 128:
        pytest.skip("f64 fused path is limited to 128 columns")
    rng = np.random.default_rng(p)
    n = 5003
    X = rng.standard_normal((n, p)).astype(dt)
    import scipy.sparse as sps

    A = sps.random(n, 300, density=0.01, random_state=rng, format="csr").astype(dt)
    d = rng.standard_normal(n).astype(dt)
    d[rng.random(n) < 0.1] = 0
    Ks = [3, 40, 1500]
    dfs = [False, True, False]
    codes = [rng.integers(-1, K + int(df), size=n).astype(np.int32) for K, df in zip(Ks, dfs)]
    rows = np.sort(rng.choice(n, size=n // 2, replace=False)).astype(np.int32) if with_rows else None
    outs, out_s = dense_cross_sandwich(
        _dev(X), _dev(d), _i32(rows), [(_i32(c), K, df) for c, K, df in zip(codes, Ks, dfs)],
        _csr(A))
    torch.cuda.synchronize()
    for o, c, K, df in zip(outs, codes, Ks, dfs):
        ref = orc.cat_dense_sandwich(c, K, d, X, rows, None, df)
        cases.assert_close(o.cpu().numpy(), ref, dt, f"fused cat K={K}")
    cases.assert_close(out_s.cpu().numpy(), orc.csr_dense_sandwich(A, X, d, rows), dt, "fused sparse")


@pytest.mark.parametrize("with_rows", [False, True])
@pytest.mark.parametrize("p_dense", [128, 64, 200])
@pytest.mark.parametrize("gather", ["0", "1"])
def test_split_native_path_all_kernel_families(with_rows, p_dense, gather, monkeypatch):
    """SplitMatrix.sandwich through tm_split_sandwich_* with every kernel family engaged: the
    tcgen05 SYRK + one-hot MMAs (few levels), the sorted-gather kernel (many levels), the RED
    scatter pass (sparse x dense), and the index-only blocks — against the oracle, block by
    block, on a matrix with missing codes and drop_first."""
    import scipy.sparse as sps

    import tabmat_b200 as tm
    from oracle import c_oracle as orc

    monkeypatch.setenv("TABMAT_B200_GATHER", gather)  # read when the native plan is built
    rng = np.random.default_rng(11 + p_dense)
    n = 30_011
    X = rng.standard_normal((n, p_dense)).astype(np.float32)
    A = sps.random(n, 70, density=0.03, random_state=rng, format="csc").astype(np.float32)
    specs = [(7, False, False), (300, True, True), (40, False, True), (1200, False, False)]
    cats, codes = [], []
    for K, df, missing in specs:
        c = rng.integers(0, K, size=n).astype(np.int32)
        if missing:
            c[rng.random(n) < 0.1] = -1
        codes.append(c)
        cats.append(tm.CategoricalMatrix(c, categories=np.arange(K), drop_first=df,
                                         dtype=np.float32, cat_missing_method="zero"))
    S = tm.SplitMatrix([tm.DenseMatrix(X), tm.SparseMatrix(A)] + cats)
    d = rng.standard_normal(n).astype(np.float32)
    rows = np.sort(rng.choice(n, size=n // 3, replace=False)).astype(np.int32) if with_rows else None
    got = S.sandwich(d, rows)
    full = np.hstack([X.astype(np.float64), A.toarray().astype(np.float64)] + [
        np.eye(K + 1)[c + 1][:, 1 + int(df):] for (K, df, _), c in zip(specs, codes)])
    sel = full if rows is None else full[rows]
    dd = d.astype(np.float64) if rows is None else d[rows].astype(np.float64)
    ref = sel.T @ (dd[:, None] * sel)
    cases.assert_close(got, ref, np.float32, "split native sandwich")
    # block-level check of the two dense cross kernels against the oracle
    o_dense = p_dense
    o_cat3 = p_dense + 70 + 7 + 299 + 40
    blk = got[o_cat3:o_cat3 + 1200, :o_dense]
    cases.assert_close(blk, orc.cat_dense_sandwich(codes[3], 1200, d, X, rows), np.float32,
                       "gather block vs oracle")
