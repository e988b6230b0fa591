"""The 16-byte-load forms of the categorical histogram (sandwich / transpose_matvec) and gather
(matvec) kernels, csrc/categorical.cu: taken for unrestricted calls with n >= 4096.  Against
numpy (np.bincount / fancy indexing = what categorical.pyx:23-218 computes) for ragged n,
both dtypes, drop_first, missing codes, random and sorted codes (the one-atomic-per-warp path),
and a misaligned view (falls back to the scalar kernels)."""

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("suf", ["f32", "f64"])
@pytest.mark.parametrize("n", [4096, 100_003, 262_145])
@pytest.mark.parametrize("K,drop_first,missing,sort", [
    (2000, False, False, False), (7, True, False, False), (300, False, True, False),
    (50, True, True, True), (2000, False, False, True), (1, False, False, False)])
def test_cat_vec_kernels(suf, n, K, drop_first, missing, sort):
    import tabmat_b200 as tm

    dt = cases.DTYPES[suf]
    rng = np.random.default_rng(n + K)
    codes = rng.integers(0, K, size=n).astype(np.int32)
    if missing:
        codes[rng.random(n) < 0.07] = -1
    if sort:
        codes = np.sort(codes)
    C = tm.CategoricalMatrix(codes, categories=np.arange(K), dtype=dt, drop_first=drop_first,
                             cat_missing_method="zero" if missing else "fail")
    d = rng.standard_normal(n).astype(dt)
    col = codes - int(drop_first)
    ok = col >= 0
    width = K - int(drop_first)
    ref = np.bincount(col[ok], weights=d[ok].astype(np.float64), minlength=width)
    if width == 0:
        return
    got = np.asarray(C.sandwich(d).diagonal())
    cases.assert_close(got, ref, dt, "cat sandwich (vec)")
    cases.assert_close(C.transpose_matvec(d), ref, dt, "cat transpose_matvec (vec)")
    ones = np.asarray(C.sandwich(np.ones(n, dtype=dt)).diagonal())
    assert np.array_equal(ones, np.bincount(col[ok], minlength=width)), "counts must be exact"
    v = rng.standard_normal(width).astype(dt)
    mv_ref = np.where(ok, v.astype(np.float64)[np.maximum(col, 0)], 0.0)
    cases.assert_close(C.matvec(v), mv_ref, dt, "cat matvec (vec)")
    out = rng.standard_normal(n).astype(dt)
    base = out.astype(np.float64).copy()
    C.matvec(v, out=out)
    cases.assert_close(out, base + mv_ref, dt, "cat matvec accumulates into out")


def test_cat_vec_misaligned_views_fall_back():
    import torch

    from tabmat_b200.ext import categorical as ecat

    n, K = 50_001, 40
    rng = np.random.default_rng(1)
    codes = torch.from_numpy(rng.integers(0, K, size=n + 1).astype(np.int32)).cuda()[1:]
    d = torch.from_numpy(rng.random(n + 1).astype(np.float32)).cuda()[1:]
    got = ecat.sandwich_categorical(codes, d, None, K, False).cpu().numpy()
    ref = np.bincount(codes.cpu().numpy(), weights=d.cpu().numpy().astype(np.float64), minlength=K)
    cases.assert_close(got, ref, np.float32, "misaligned cat sandwich")


@pytest.mark.parametrize("suf", ["f32", "f64"])
def test_deterministic_mode_is_bit_reproducible(suf):
    """tabmat_b200.set_deterministic(True): categorical transpose_matvec / sandwich add in a fixed
    order (tm_cat_segment_sum) — bit-identical across runs (the reference made this operation
    deterministic, CHANGELOG.rst:134) and equal to numpy within rounding."""
    import tabmat_b200 as tm

    dt = cases.DTYPES[suf]
    n, K = 300_001, 500
    rng = np.random.default_rng(9)
    codes = rng.integers(-1, K, size=n).astype(np.int32)
    C = tm.CategoricalMatrix(codes, categories=np.arange(K), dtype=dt, drop_first=True,
                             cat_missing_method="zero")
    v = (rng.standard_normal(n) * 10.0 ** rng.integers(-3, 4, size=n)).astype(dt)
    rows = np.sort(rng.choice(n, size=n // 2, replace=False)).astype(np.int32)
    cols = np.arange(0, K - 1, 3, dtype=np.int32)
    col = codes - 1
    try:
        tm.set_deterministic(True)
        a = [C.transpose_matvec(v) for _ in range(4)]
        assert all(np.array_equal(a[0], x) for x in a[1:]), "not bit-reproducible"
        ok = col >= 0
        ref = np.bincount(col[ok], weights=v[ok].astype(np.float64), minlength=K - 1)
        cases.assert_close(a[0], ref, dt, "deterministic transpose_matvec")
        m = np.zeros(n, dtype=bool)
        m[rows] = True
        ref_r = np.bincount(col[ok & m], weights=v[ok & m].astype(np.float64), minlength=K - 1)
        cases.assert_close(C.transpose_matvec(v, rows=rows), ref_r, dt, "deterministic, rows")
        cases.assert_close(C.transpose_matvec(v, rows=rows, cols=cols), ref_r[cols], dt,
                           "deterministic, rows + cols")
        d = np.abs(v)
        s = [np.asarray(C.sandwich(d).diagonal()) for _ in range(3)]
        assert np.array_equal(s[0], s[1]) and np.array_equal(s[0], s[2])
        cases.assert_close(s[0], np.bincount(col[ok], weights=d[ok].astype(np.float64), minlength=K - 1),
                           dt, "deterministic sandwich")
    finally:
        tm.set_deterministic(False)
