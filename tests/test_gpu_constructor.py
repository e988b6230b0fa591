"""from_pandas / from_csc: the block layout (which columns go dense / sparse / categorical, and
where) and the resulting matrix against golden results produced by the reference's own
constructors (tests/golden/make_golden.py: constructor_level)."""

import numpy as np
import pytest
import scipy.sparse as sps

from tests import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    from pathlib import Path

    return np.load(Path(__file__).resolve().parent / "golden" / "constructors.npz")


@pytest.mark.parametrize("pos", ["expand", "end"])
@pytest.mark.parametrize("drop_first", [False, True])
def test_from_pandas_layout_and_values(pos, drop_first, golden):
    import tabmat_b200 as tm

    X = tm.from_pandas(cases.constructor_frame(), cat_position=pos, drop_first=drop_first)
    key = f"{pos}-df{int(drop_first)}/"
    assert [type(m).__name__ for m in X.matrices] == list(golden[key + "kinds"])
    np.testing.assert_array_equal(np.concatenate(X.indices), golden[key + "index_concat"])
    np.testing.assert_array_equal([len(i) for i in X.indices], golden[key + "index_sizes"])
    np.testing.assert_allclose(X.toarray(), golden[key + "toarray"])
    d = np.linspace(0.5, 1.5, X.shape[0])
    cases.assert_close(X.sandwich(d), golden[key + "sandwich"], np.float64, "from_pandas sandwich")


def test_from_csc_threshold_split(golden):
    import tabmat_b200 as tm

    A = sps.csc_matrix((golden["csc/A_data"], golden["csc/A_indices"], golden["csc/A_indptr"]),
                       shape=(50, 12))
    C = tm.from_csc(A, threshold=0.3)
    assert [type(m).__name__ for m in C.matrices] == ["DenseMatrix", "SparseMatrix"]
    np.testing.assert_array_equal(np.concatenate(C.indices), golden["csc/index_concat"])
    np.testing.assert_array_equal([len(i) for i in C.indices], golden["csc/index_sizes"])
    np.testing.assert_allclose(C.toarray(), golden["csc/toarray"])
    with pytest.raises(TypeError):
        tm.from_csc(A.tocsr())
    with pytest.raises(ValueError, match="between 0 and 1"):
        tm.from_csc(A, threshold=1.5)


def test_from_pandas_rejects_and_warns():
    import pandas as pd

    import tabmat_b200 as tm

    with pytest.warns(UserWarning, match="ignored"):
        X = tm.from_pandas(pd.DataFrame({"x": [1.0, 2.0, 3.0], "s": ["a", "b", "c"]}))
    assert isinstance(X, tm.DenseMatrix) and X.shape == (3, 1)
    with pytest.raises(ValueError, match="no valid column"):
        with pytest.warns(UserWarning):
            tm.from_pandas(pd.DataFrame({"s": ["a", "b"]}))
    C = tm.from_pandas(pd.DataFrame({"s": ["a", "b", "a", "c", "d"]}), object_as_cat=True)
    assert isinstance(C, tm.CategoricalMatrix) and C.shape == (5, 4)


@pytest.mark.parametrize("n_num", [8, 7])
def test_from_pandas_matrix_takes_the_tensor_path(n_num):
    """A SplitMatrix built the normal tabmat way (pandas -> F-ordered numpy columns) is stored
    row-major in HBM - zero-padded to a multiple of 4 columns when it has 7 numeric columns - and
    its f32 sandwich runs the tcgen05 pass of tm_split_sandwich_blocks (pass timer of the native
    call > 0), not the CUDA-core fallback."""
    import ctypes

    import pandas as pd

    import tabmat_b200 as tm

    lib = tm._lib.lib
    if not lib.tm_has_tcgen05():
        pytest.skip("needs sm_100")
    rng = np.random.default_rng(3)
    n = 5000
    cols = {f"x{i}": rng.standard_normal(n).astype(np.float32) for i in range(n_num)}
    cols["big"] = pd.Categorical(rng.integers(0, 300, size=n))
    cols["small"] = pd.Categorical(rng.integers(0, 6, size=n))
    df = pd.DataFrame(cols)
    assert not df[[f"x{i}" for i in range(n_num)]].to_numpy().flags["C_CONTIGUOUS"]
    X = tm.from_pandas(df, dtype=np.float32)
    dense = [m for m in X.matrices if isinstance(m, tm.DenseMatrix)]
    assert len(dense) == 1 and dense[0]._native().is_contiguous()
    assert dense[0].shape[1] == n_num and dense[0]._native().shape[1] % 4 == 0
    d = rng.random(n).astype(np.float32)
    lib.tm_split_profile_enable(1)
    got = X.sandwich(d)
    ms = (ctypes.c_float * 3)()
    lib.tm_split_profile_read(ms)
    lib.tm_split_profile_enable(0)
    assert ms[0] > 0, "the tensor pass did not run"
    full = X.toarray()
    cases.assert_close(got, (full * d[:, None].astype(np.float64)).T @ full, np.float32,
                       "from_pandas f32 sandwich")
    C = tm.from_csc(sps.random(400, 16, density=0.5, format="csc", random_state=rng), threshold=0.1)
    assert all(m._native().is_contiguous() for m in C.matrices if isinstance(m, tm.DenseMatrix))
