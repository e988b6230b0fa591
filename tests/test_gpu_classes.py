"""Class-level parity: SplitMatrix / StandardizedMatrix through the MatrixBase API against
golden results produced by the reference package itself (tests/golden/make_golden.py)."""

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


def _build(dt, seed=77, n=83):
    import tabmat_b200 as tm

    I = cases.make_inputs(seed, n, dt, p_dense=5, p_sparse=6, Ki=4, Kj=3)  # noqa: E741
    mats = [
        tm.DenseMatrix(I["X"]),
        tm.SparseMatrix(I["A"]),
        tm.CategoricalMatrix(I["ci_missing"], categories=np.arange(I["Ki"]), dtype=dt,
                             cat_missing_method="zero"),
        tm.CategoricalMatrix(I["cj"], categories=np.arange(I["Kj"]), dtype=dt, drop_first=True),
    ]
    return tm.SplitMatrix(mats), I


@pytest.mark.parametrize("suf", ["f32", "f64"])
def test_split_matrix_against_reference(suf, golden_classes):
    dt = cases.DTYPES[suf]
    G = {k[len(suf) + 1:]: golden_classes[k] for k in golden_classes.files if k.startswith(suf)}
    X, I = _build(dt)  # noqa: E741
    assert X.shape[1] == int(G["p"])
    np.testing.assert_array_equal(X.toarray(), G["toarray"])
    cols, v_p = G["cols"], G["v_p"]
    for rname, rows in (("all", None), ("rows", I["rows"])):
        for cname, c in (("all", None), ("cols", cols)):
            t = f"{rname}-{cname}"
            got = X.sandwich(I["d"], rows, c)
            assert got.dtype == np.float64  # split_matrix.py:336
            cases.assert_close(got, G[f"sandwich-{t}"], dt, f"sandwich {t}")
            got = X.transpose_matvec(I["v_n"], rows, c)
            assert got.dtype == G[f"transpose_matvec-{t}"].dtype
            cases.assert_close(got, G[f"transpose_matvec-{t}"], dt, f"transpose_matvec {t}")
        got = X.matvec(v_p, None if rname == "all" else cols)
        cases.assert_close(got, G[f"matvec-{rname}"], dt, f"matvec {rname}")


@pytest.mark.parametrize("suf", ["f32", "f64"])
@pytest.mark.parametrize("center", [False, True])
@pytest.mark.parametrize("scale", [False, True])
def test_standardize_against_reference(suf, center, scale, golden_classes):
    dt = cases.DTYPES[suf]
    G = {k[len(suf) + 1:]: golden_classes[k] for k in golden_classes.files if k.startswith(suf)}
    X, I = _build(dt)  # noqa: E741
    t = f"c{int(center)}s{int(scale)}"
    S, means, stds = X.standardize(I["w"] / I["w"].sum(), center, scale)
    cases.assert_close(means, G[f"std-means-{t}"], dt, "means")
    if scale:
        cases.assert_close(stds, G[f"std-stds-{t}"], dt, "stds")
    else:
        assert stds is None
    cases.assert_close(S.sandwich(I["d"]), G[f"std-sandwich-{t}"], dt, "std sandwich")
    cases.assert_close(S.sandwich(I["d"], I["rows"], G["cols"]), G[f"std-sandwich-rc-{t}"], dt,
                       "std sandwich rows/cols")
    cases.assert_close(S.matvec(G["v_p"]), G[f"std-matvec-{t}"], dt, "std matvec")
    cases.assert_close(S.transpose_matvec(I["v_n"]), G[f"std-transpose_matvec-{t}"], dt,
                       "std transpose_matvec")
