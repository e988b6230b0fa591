"""Generate ``tests/golden/*.npz`` from the REFERENCE itself (build container only).

Runs every case of ``tests/cases.py`` through the reference's own compiled kernels
(``oracle/_ref``: the reference's Cython/C++ sources built by ``oracle/build_ref.py``) and its
own Python classes (imported from ``/root/reference`` by ``oracle/ref_loader.py``), and stores
inputs + outputs.  The fixtures travel to the GPU box; the reference does not.

    python tests/golden/make_golden.py

Files written:
  boundary.npz   one entry per boundary-level case (ext-function level)
  classes.npz    class-level results: SplitMatrix / StandardizedMatrix sandwich, matvec,
                 transpose_matvec, standardize() on a mixed dense+sparse+categorical matrix
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import scipy.sparse as sps

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))

import ref_loader  # noqa: E402

from tests import cases  # noqa: E402

tm = ref_loader.import_reference_package()
from tabmat.ext import categorical as rcat  # noqa: E402
from tabmat.ext import dense as rdense  # noqa: E402
from tabmat.ext import sparse as rsparse  # noqa: E402
from tabmat.ext import split as rsplit  # noqa: E402


def _ar(x, n):
    return np.arange(n, dtype=np.int32) if x is None else np.asarray(x, dtype=np.int32)


def run_reference(kind, a):
    if kind == "dense_sandwich":
        return tm.DenseMatrix(a["X"]).sandwich(a["d"], a["rows"], a["cols"])
    if kind == "dense_matvec":
        n, p = a["X"].shape
        return rdense.dense_matvec(a["X"], a["v"], _ar(a["rows"], n), _ar(a["cols"], p))
    if kind == "dense_rmatvec":
        n, p = a["X"].shape
        return rdense.dense_rmatvec(a["X"], a["v"], _ar(a["rows"], n), _ar(a["cols"], p))
    if kind == "dense_sq_dot_weights":
        return rdense.transpose_square_dot_weights(a["X"], a["w"], a["shift"])
    if kind == "sparse_sandwich":
        return tm.SparseMatrix(a["A"]).sandwich(a["d"], a["rows"], a["cols"])
    if kind == "csr_dense_sandwich":
        return tm.SparseMatrix(a["A"]).sandwich_dense(a["B"], a["d"], a["rows"], a["A_cols"],
                                                      a["B_cols"])
    if kind == "csr_matvec":
        n, p = a["A"].shape
        return rsparse.csr_matvec(a["A"].tocsr(), a["v"], _ar(a["rows"], n), _ar(a["cols"], p))
    if kind == "csc_rmatvec":
        n, p = a["A"].shape
        return rsparse.csc_rmatvec(a["A"], a["v"], _ar(a["rows"], n), _ar(a["cols"], p))
    if kind == "csc_sq_dot_weights":
        A = a["A"]
        return rsparse.transpose_square_dot_weights(A.data, A.indices, A.indptr, a["w"], A.dtype)
    if kind == "cat_sandwich":
        n = len(a["codes"])
        return np.asarray(rcat.sandwich_categorical_complex(
            a["codes"], a["d"], _ar(a["rows"], n), a["d"].dtype, a["K"], a["drop_first"]))
    if kind == "cat_transpose_matvec":
        out = np.zeros(a["K"], a["v"].dtype)
        rcat.transpose_matvec_complex(a["codes"], a["v"], a["K"], a["v"].dtype, a["rows"],
                                      a["cols"], out, a["drop_first"])
        return out
    if kind == "cat_matvec":
        n = len(a["codes"])
        out = np.zeros(n, a["v"].dtype)
        rcat.matvec_complex(a["codes"], a["v"], n, a["cols"], a["K"], out, a["drop_first"])
        return out
    if kind == "cat_dense_sandwich":
        n, q = a["Y"].shape
        return rsplit.sandwich_cat_dense(
            a["codes"], a["K"], a["d"], a["Y"], _ar(a["rows"], n), _ar(a["j_cols"], q),
            bool(a["Y"].flags["C_CONTIGUOUS"]), True, a["drop_first"])
    if kind == "cat_cat_sandwich":
        n = len(a["ic"])
        return rsplit.sandwich_cat_cat(
            a["ic"], a["jc"], a["Ki"], a["Kj"], a["d"], _ar(a["rows"], n), a["d"].dtype,
            a["i_drop_first"], a["j_drop_first"], True, True)
    if kind == "cat_sparse_sandwich":
        # no native function in the reference: CategoricalMatrix._cross_sparse (scipy)
        Kfull = a["K"] + int(a["drop_first"])
        cm = tm.CategoricalMatrix(a["codes"], categories=np.arange(Kfull),
                                  drop_first=a["drop_first"], cat_missing_method="zero",
                                  dtype=a["d"].dtype)
        return np.asarray(cm._cross_sandwich(tm.SparseMatrix(a["A"]), a["d"], a["rows"], None,
                                             a["s_cols"]))
    raise KeyError(kind)


def class_level(seed=77, n=83):
    """Class-level golden: a mixed SplitMatrix through the reference's public API."""
    out = {}
    for suf, dt in cases.DTYPES.items():
        I = cases.make_inputs(seed, n, dt, p_dense=5, p_sparse=6, Ki=4, Kj=3)  # noqa: E741
        mats = [
            tm.DenseMatrix(I["X"]),
            tm.SparseMatrix(I["A"]),
            tm.CategoricalMatrix(I["ci_missing"], categories=np.arange(I["Ki"]), dtype=dt,
                                 cat_missing_method="zero"),
            tm.CategoricalMatrix(I["cj"], categories=np.arange(I["Kj"]), dtype=dt,
                                 drop_first=True),
        ]
        X = tm.SplitMatrix(mats)
        p = X.shape[1]
        rng = np.random.default_rng(seed + 1)
        v_p = rng.standard_normal(p).astype(dt)
        cols = np.sort(rng.choice(p, size=p // 2, replace=False)).astype(np.int32)
        pre = f"{suf}/"
        out[pre + "p"] = np.array(p)
        out[pre + "v_p"] = v_p
        out[pre + "cols"] = cols
        out[pre + "toarray"] = X.toarray()
        for rname, rows in (("all", None), ("rows", I["rows"])):
            for cname, c in (("all", None), ("cols", cols)):
                t = f"{rname}-{cname}"
                out[pre + f"sandwich-{t}"] = X.sandwich(I["d"], rows, c)
                out[pre + f"transpose_matvec-{t}"] = X.transpose_matvec(I["v_n"], rows, c)
            out[pre + f"matvec-{rname}"] = X.matvec(v_p, None if rname == "all" else cols)
        for center in (False, True):
            for scale in (False, True):
                S, means, stds = X.standardize(I["w"] / I["w"].sum(), center, scale)
                t = f"c{int(center)}s{int(scale)}"
                out[pre + f"std-means-{t}"] = means
                if stds is not None:
                    out[pre + f"std-stds-{t}"] = stds
                out[pre + f"std-sandwich-{t}"] = S.sandwich(I["d"])
                out[pre + f"std-sandwich-rc-{t}"] = S.sandwich(I["d"], I["rows"], cols)
                out[pre + f"std-matvec-{t}"] = S.matvec(v_p)
                out[pre + f"std-transpose_matvec-{t}"] = S.transpose_matvec(I["v_n"])
    return out


def constructor_level():
    out = {}
    df = cases.constructor_frame()
    for pos in ("expand", "end"):
        for df_first in (False, True):
            X = tm.from_pandas(df, cat_position=pos, drop_first=df_first)
            key = f"{pos}-df{int(df_first)}/"
            out[key + "toarray"] = X.toarray()
            out[key + "kinds"] = np.array([type(m).__name__ for m in X.matrices])
            out[key + "index_concat"] = np.concatenate(X.indices)
            out[key + "index_sizes"] = np.array([len(i) for i in X.indices])
            d = np.linspace(0.5, 1.5, X.shape[0])
            out[key + "sandwich"] = X.sandwich(d)
    rng = np.random.default_rng(8)
    A = sps.random(50, 12, density=0.2, random_state=rng, format="csc")
    A[:, 3] = rng.standard_normal((50, 1))
    A = sps.csc_matrix(A)
    C = tm.from_csc(A, threshold=0.3)
    out["csc/A_data"], out["csc/A_indices"], out["csc/A_indptr"] = A.data, A.indices, A.indptr
    out["csc/toarray"] = C.toarray()
    out["csc/index_concat"] = np.concatenate(C.indices)
    out["csc/index_sizes"] = np.array([len(i) for i in C.indices])
    return out


def main():
    here = Path(__file__).resolve().parent
    boundary = {}
    count = 0
    for name, kind, args in cases.boundary_cases():
        ref = np.asarray(run_reference(kind, args))
        boundary[name] = ref
        # cross-check right here: the reference agrees with a dense float64 recomputation
        cases.assert_close(ref, cases.run_numpy(kind, args), ref.dtype, name)
        count += 1
    np.savez_compressed(here / "boundary.npz", **boundary)
    cls = class_level()
    np.savez_compressed(here / "classes.npz", **cls)
    ctor = constructor_level()
    np.savez_compressed(here / "constructors.npz", **ctor)
    print(f"wrote {count} boundary cases, {len(cls)} class-level and {len(ctor)} constructor arrays")


if __name__ == "__main__":
    main()
