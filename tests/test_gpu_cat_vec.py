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
