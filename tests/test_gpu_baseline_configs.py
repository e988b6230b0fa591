"""The BASELINE.json configs at 1e6 rows against the REFERENCE ITSELF: the stock ``tabmat``
package + its Cython/C++ kernels built from source (``oracle/_ref/tabmat``, OpenMP, seconds per
case) and tabmat_b200 are fed the identical host arrays (the sample generators of bench.py).

  C2  DenseMatrix f32, p = 256          -> the tcgen05 SYRK (three 128x128 tiles)
  C3  CategoricalMatrix, 2000 levels    -> weighted histogram; bit-exact counts with d == 1
  C4  [dense 128 | CSC 5000] f64        -> sparse self + dense x sparse cross (DMMA dense self)
  C5  128 dense + 3000 CSC + 5 cats f32 -> the whole fused SplitMatrix path, both row orders

The normwise error (max|ours - ref| / max|ref|, tolerance 1e-3 f32 / 1e-5 f64 as BASELINE.json
states) is asserted AND printed, together with the element-wise figures of the TF32 dense
block, so that the TF32 margin is on record (run with -s / see gpurun_out/parity_configs.json).
Reference lines matched: tests/test_split_matrix.py:229-288, tests/test_matrices.py:255-432."""

import json
import os
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]
ROWS = 1_000_000
_RECORD = {}


def _reference(wl, sample):
    from oracle import ref_loader

    if not ref_loader.package_available():
        pytest.skip("oracle/_ref/tabmat (the stock reference package) is not installed")
    ref_loader.set_omp_threads(len(os.sched_getaffinity(0)))
    tabmat = ref_loader.import_installed_package()
    return wl.dense_result(wl.ref_from_sample(tabmat, sample).sandwich(sample["d"]))


def _record(key, **vals):
    _RECORD[key] = vals
    print(f"\nparity {key}: " + ", ".join(f"{k}={v:.3e}" if isinstance(v, float) else f"{k}={v}"
                                           for k, v in vals.items()))
    out = ROOT / "gpurun_out"
    if out.is_dir():
        (out / "parity_configs.json").write_text(json.dumps(_RECORD, indent=1))


def _errs(got, ref):
    d, r = np.abs(got - ref), np.abs(ref)
    return float(d.max() / r.max()), float((d / np.maximum(r, 1e-30)).max())


def test_c2_dense_f32_tcgen05_vs_reference():
    import bench
    import tabmat_b200 as tm

    wl = bench.C2(None)
    s = wl.host_sample(ROWS, seed=21)
    ref = _reference(wl, s)
    X = wl.ours_from_sample(s, 0, ROWS)
    assert X._array.is_contiguous()
    tm.reset_launch_count()
    got = X.sandwich(s["d"])
    assert tm.launch_count() > 0
    ref64 = (s["X"].astype(np.float64) * s["d"].astype(np.float64)[:, None]).T @ s["X"].astype(np.float64)
    norm, elem = _errs(got, ref)
    offd = ~np.eye(256, dtype=bool)
    offn = float(np.abs(got - ref)[offd].max() / np.abs(ref)[offd].max())
    n64, _ = _errs(got, ref64)
    r64, _ = _errs(ref, ref64)
    _record("c2", normwise=norm, offdiag_normwise=offn, max_elementwise_rel=elem,
            ours_vs_f64=n64, reference_vs_f64=r64, tol=1e-3, rows=ROWS)
    assert norm <= 1e-3 and offn <= 1e-3 * 50  # off-diagonal entries are ~sqrt(n) times smaller
    assert np.array_equal(got, got.T)


def test_c3_categorical_vs_reference_and_exact_counts():
    import bench

    wl = bench.C3(None)
    s = wl.host_sample(ROWS, seed=22)
    ref = _reference(wl, s)
    X = wl.ours_from_sample(s, 0, ROWS)
    got = wl.dense_result(X.sandwich(s["d"]))
    norm, _ = _errs(got, ref)
    ones = wl.dense_result(X.sandwich(np.ones(ROWS, dtype=np.float32)))
    exact = bool(np.array_equal(np.diag(ones), np.bincount(s["codes"], minlength=wl.K)))
    _record("c3", normwise=norm, counts_bit_exact=exact, tol=1e-3, rows=ROWS)
    assert norm <= 1e-3 and exact


def test_c4_sparse_f64_self_and_cross_vs_reference():
    import bench

    wl = bench.C4(None)
    s = wl.host_sample(ROWS, seed=23)
    ref = _reference(wl, s)
    X = wl.ours_from_sample(s, 0, ROWS)
    got = X.sandwich(s["d"])          # [A^T D A | A^T D B], the two calls configs[3] names
    P = wl.P
    n_all, _ = _errs(got, ref)
    n_self, _ = _errs(got[:, :P], ref[:, :P])
    n_cross, _ = _errs(got[:, P:], ref[:, P:])
    _record("c4", normwise=n_all, sparse_self=n_self, dense_x_sparse=n_cross, tol=1e-5, rows=ROWS)
    assert max(n_all, n_self, n_cross) <= 1e-5
    # the same two blocks inside a SplitMatrix [dense | sparse] (fused split path, f64)
    import tabmat_b200 as tm

    S = tm.SplitMatrix([tm.DenseMatrix(s["X"]), tm.SparseMatrix(s["A"])])
    H = S.sandwich(s["d"])
    q = wl.Q
    n2, _ = _errs(H[q:, q:], ref[:, :P])
    n3, _ = _errs(H[q:, :q], ref[:, P:])
    assert max(n2, n3) <= 1e-5


@pytest.mark.parametrize("order", ["sorted", "original"])
def test_c5_split_f32_vs_reference(order):
    import bench
    import tabmat_b200 as tm

    wl = bench.C5(None)
    s = wl.host_sample(ROWS, seed=24)
    ref = _reference(wl, s)
    X = wl.ours_from_sample(s, 0, ROWS)
    if order == "sorted":
        X = tm.RowSortedMatrix.from_split(X)
    got = X.sandwich(s["d"])
    P = bench.P_DENSE
    norm, _ = _errs(got, ref)
    dn, de = _errs(got[:P, :P], ref[:P, :P])
    rest = np.abs(got - ref)
    rest[:P, :P] = 0
    _record(f"c5_{order}", normwise=norm, dense_block_normwise=dn,
            dense_block_max_elementwise_rel=de, other_blocks_normwise=float(rest.max() / np.abs(ref).max()),
            tol=1e-3, rows=ROWS)
    assert norm <= 1e-3
    # rows= restriction at this size (sorted unique, every 3rd row) against the reference
    rows = np.arange(0, ROWS, 3, dtype=np.int32)
    from oracle import ref_loader

    tabmat = ref_loader.import_installed_package()
    ref_r = wl.ref_from_sample(tabmat, s).sandwich(s["d"], rows=rows)
    nr, _ = _errs(X.sandwich(s["d"], rows=rows), ref_r)
    assert nr <= 1e-3, nr
