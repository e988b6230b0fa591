"""Seeded boundary-level parity cases shared by the golden generator, the CPU oracle tests and
the GPU parity tests.

A case is ``(name, kind, args)`` where ``kind`` names one entry point of
``include/tabmat_b200.h`` (without the ``tm_`` prefix and dtype suffix) and ``args`` is a dict
of host numpy / scipy inputs.  ``run_oracle`` evaluates it with the C restatement under
``oracle/``; ``tests/golden/make_golden.py`` evaluates it with the reference package itself;
``tests/test_gpu_parity.py`` evaluates it through the C-ABI on the device.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sps

DTYPES = {"f32": np.float32, "f64": np.float64}


def _subset(rng, n, frac=0.6):
    k = max(1, int(n * frac))
    return np.sort(rng.choice(n, size=k, replace=False)).astype(np.int32)


def make_inputs(seed: int, n: int, dtype, p_dense=7, p_sparse=9, Ki=5, Kj=4, density=0.25):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, p_dense)).astype(dtype)
    A = sps.random(n, p_sparse, density=density, random_state=rng, format="csc", dtype=np.float64)
    A = sps.csc_matrix(A.astype(dtype))
    A.sort_indices()
    ci = rng.integers(0, Ki, size=n).astype(np.int32)
    cj = rng.integers(0, Kj, size=n).astype(np.int32)
    ci_missing = ci.copy()
    ci_missing[rng.random(n) < 0.15] = -1
    d = rng.standard_normal(n).astype(dtype)  # d may be negative
    d[rng.random(n) < 0.1] = 0
    return dict(
        n=n, X=X, A=A, ci=ci, cj=cj, ci_missing=ci_missing, Ki=Ki, Kj=Kj, d=d,
        v_n=rng.standard_normal(n).astype(dtype),
        v_dense=rng.standard_normal(p_dense).astype(dtype),
        v_sparse=rng.standard_normal(p_sparse).astype(dtype),
        v_Ki=rng.standard_normal(Ki).astype(dtype),
        w=rng.random(n).astype(dtype),
        rows=_subset(rng, n),
        cols_dense=_subset(rng, p_dense),
        cols_sparse=_subset(rng, p_sparse),
        cols_Ki=_subset(rng, Ki),
    )


def boundary_cases(seed=1234, n=67):
    """Yield (name, kind, args) for every entry point x dtype x layout x restriction."""
    for suf, dt in DTYPES.items():
        I = make_inputs(seed, n, dt)  # noqa: E741
        for order in ("C", "F"):
            X = np.asarray(I["X"], order=order)
            for rname, rows in (("all", None), ("rows", I["rows"])):
                for cname, cols in (("all", None), ("cols", I["cols_dense"])):
                    tag = f"{suf}-{order}-{rname}-{cname}"
                    yield (f"dense_sandwich-{tag}", "dense_sandwich",
                           dict(X=X, d=I["d"], rows=rows, cols=cols))
                    yield (f"dense_matvec-{tag}", "dense_matvec",
                           dict(X=X, v=I["v_dense"], rows=rows, cols=cols))
                    yield (f"dense_rmatvec-{tag}", "dense_rmatvec",
                           dict(X=X, v=I["v_n"], rows=rows, cols=cols))
                    yield (f"csr_dense_sandwich-{tag}", "csr_dense_sandwich",
                           dict(A=I["A"], B=X, d=I["d"], rows=rows,
                                A_cols=None if cols is None else I["cols_sparse"], B_cols=cols))
                    for df in (False, True):
                        yield (f"cat_dense_sandwich-{tag}-df{int(df)}", "cat_dense_sandwich",
                               dict(codes=I["ci_missing"], K=I["Ki"] - int(df), d=I["d"], Y=X,
                                    rows=rows, j_cols=cols, drop_first=df))
            shift = X.mean(axis=0).astype(dt)
            yield (f"dense_sq_dot_weights-{suf}-{order}", "dense_sq_dot_weights",
                   dict(X=X, w=I["w"], shift=shift))
        for rname, rows in (("all", None), ("rows", I["rows"])):
            for cname, cols in (("all", None), ("cols", I["cols_sparse"])):
                tag = f"{suf}-{rname}-{cname}"
                yield (f"sparse_sandwich-{tag}", "sparse_sandwich",
                       dict(A=I["A"], d=I["d"], rows=rows, cols=cols))
                yield (f"csr_matvec-{tag}", "csr_matvec",
                       dict(A=I["A"], v=I["v_sparse"], rows=rows, cols=cols))
                yield (f"csc_rmatvec-{tag}", "csc_rmatvec",
                       dict(A=I["A"], v=I["v_n"], rows=rows, cols=cols))
                for df in (False, True):
                    yield (f"cat_sparse_sandwich-{tag}-df{int(df)}", "cat_sparse_sandwich",
                           dict(codes=I["ci_missing"], K=I["Ki"] - int(df), d=I["d"], A=I["A"],
                                rows=rows, s_cols=cols, drop_first=df))
            for codes_name in ("ci", "ci_missing"):
                for df in (False, True):
                    K = I["Ki"] - int(df)
                    tag = f"{suf}-{rname}-{codes_name}-df{int(df)}"
                    yield (f"cat_sandwich-{tag}", "cat_sandwich",
                           dict(codes=I[codes_name], d=I["d"], rows=rows, K=K, drop_first=df))
                    for cname, cols in (("all", None), ("cols", I["cols_Ki"][I["cols_Ki"] < K])):
                        yield (f"cat_transpose_matvec-{tag}-{cname}", "cat_transpose_matvec",
                               dict(codes=I[codes_name], v=I["v_n"], rows=rows, cols=cols, K=K,
                                    drop_first=df))
                    yield (f"cat_cat_sandwich-{tag}", "cat_cat_sandwich",
                           dict(ic=I[codes_name], jc=I["cj"], Ki=K, Kj=I["Kj"], d=I["d"],
                                rows=rows, i_drop_first=df, j_drop_first=False))
        for codes_name in ("ci", "ci_missing"):
            for df in (False, True):
                K = I["Ki"] - int(df)
                for cname, cols in (("all", None), ("cols", I["cols_Ki"][I["cols_Ki"] < K])):
                    yield (f"cat_matvec-{suf}-{codes_name}-df{int(df)}-{cname}", "cat_matvec",
                           dict(codes=I[codes_name], v=I["v_Ki"][:K].copy(), cols=cols, K=K,
                                drop_first=df))
        yield (f"csc_sq_dot_weights-{suf}", "csc_sq_dot_weights", dict(A=I["A"], w=I["w"]))


def run_oracle(kind: str, args: dict):
    """Evaluate a case with the C restatement (oracle/c_oracle.py)."""
    from oracle import c_oracle as orc

    a = args
    if kind == "dense_sandwich":
        return orc.dense_sandwich(a["X"], a["d"], a["rows"], a["cols"])
    if kind == "dense_matvec":
        return orc.dense_matvec(a["X"], a["v"], a["rows"], a["cols"])
    if kind == "dense_rmatvec":
        return orc.dense_rmatvec(a["X"], a["v"], a["rows"], a["cols"])
    if kind == "dense_sq_dot_weights":
        return orc.dense_sq_dot_weights(a["X"], a["w"], a["shift"])
    if kind == "sparse_sandwich":
        return orc.sparse_sandwich(a["A"], a["d"], a["rows"], a["cols"])
    if kind == "csr_dense_sandwich":
        return orc.csr_dense_sandwich(a["A"], a["B"], a["d"], a["rows"], a["A_cols"], a["B_cols"])
    if kind == "csr_matvec":
        return orc.csr_matvec(a["A"], a["v"], a["rows"], a["cols"])
    if kind == "csc_rmatvec":
        return orc.csc_rmatvec(a["A"], a["v"], a["rows"], a["cols"])
    if kind == "csc_sq_dot_weights":
        return orc.csc_sq_dot_weights(a["A"], a["w"])
    if kind == "cat_sandwich":
        return orc.cat_sandwich(a["codes"], a["d"], a["rows"], a["K"], a["drop_first"])
    if kind == "cat_transpose_matvec":
        return orc.cat_transpose_matvec(a["codes"], a["v"], a["rows"], a["cols"], a["K"],
                                        a["drop_first"])
    if kind == "cat_matvec":
        return orc.cat_matvec(a["codes"], a["v"], a["cols"], a["K"], a["drop_first"])
    if kind == "cat_dense_sandwich":
        return orc.cat_dense_sandwich(a["codes"], a["K"], a["d"], a["Y"], a["rows"], a["j_cols"],
                                      a["drop_first"])
    if kind == "cat_cat_sandwich":
        return orc.cat_cat_sandwich(a["ic"], a["jc"], a["Ki"], a["Kj"], a["d"], a["rows"],
                                    a["i_drop_first"], a["j_drop_first"])
    if kind == "cat_sparse_sandwich":
        return orc.cat_sparse_sandwich(a["codes"], a["K"], a["d"], a["A"], a["rows"], a["s_cols"],
                                       a["drop_first"])
    raise KeyError(kind)


def run_numpy(kind: str, args: dict):
    """Dense float64 recomputation (the reference's own test strategy)."""
    a = args

    def sel(M, rows, cols):
        M = np.asarray(M.todense() if sps.issparse(M) else M, dtype=np.float64)
        if rows is not None:
            M = M[rows, :]
        if cols is not None:
            M = M[:, cols]
        return M

    def onehot(codes, K, df):
        c = codes.astype(np.int64) - int(df)
        M = np.zeros((len(codes), K))
        ok = c >= 0
        M[np.flatnonzero(ok), c[ok]] = 1
        return M

    def rsel(v, rows):
        v = np.asarray(v, dtype=np.float64)
        return v if rows is None else v[rows]

    if kind == "dense_sandwich":
        M = sel(a["X"], a["rows"], a["cols"])
        return M.T @ (rsel(a["d"], a["rows"])[:, None] * M)
    if kind == "dense_matvec":
        v = np.asarray(a["v"], dtype=np.float64)
        M = sel(a["X"], a["rows"], a["cols"])
        return M @ (v if a["cols"] is None else v[a["cols"]])
    if kind == "dense_rmatvec":
        return sel(a["X"], a["rows"], a["cols"]).T @ rsel(a["v"], a["rows"])
    if kind == "dense_sq_dot_weights":
        X = np.asarray(a["X"], dtype=np.float64)
        return (a["w"].astype(np.float64)[:, None] * (X - a["shift"].astype(np.float64)) ** 2).sum(0)
    if kind == "sparse_sandwich":
        M = sel(a["A"], a["rows"], a["cols"])
        return M.T @ (rsel(a["d"], a["rows"])[:, None] * M)
    if kind == "csr_dense_sandwich":
        MA = sel(a["A"], a["rows"], a["A_cols"])
        MB = sel(a["B"], a["rows"], a["B_cols"])
        return MA.T @ (rsel(a["d"], a["rows"])[:, None] * MB)
    if kind == "csr_matvec":
        v = np.asarray(a["v"], dtype=np.float64)
        M = sel(a["A"], a["rows"], a["cols"])
        return M @ (v if a["cols"] is None else v[a["cols"]])
    if kind == "csc_rmatvec":
        return sel(a["A"], a["rows"], a["cols"]).T @ rsel(a["v"], a["rows"])
    if kind == "csc_sq_dot_weights":
        M = sel(a["A"], None, None)
        return (a["w"].astype(np.float64)[:, None] * M ** 2).sum(0)
    if kind == "cat_sandwich":
        M = sel(onehot(a["codes"], a["K"], a["drop_first"]), a["rows"], None)
        return M.T @ rsel(a["d"], a["rows"])
    if kind == "cat_transpose_matvec":
        # absolute-index convention: length K, zero outside cols
        M = sel(onehot(a["codes"], a["K"], a["drop_first"]), a["rows"], None)
        out = M.T @ rsel(a["v"], a["rows"])
        if a["cols"] is not None:
            keep = np.zeros(a["K"], bool)
            keep[a["cols"]] = True
            out[~keep] = 0
        return out
    if kind == "cat_matvec":
        M = onehot(a["codes"], a["K"], a["drop_first"])
        v = np.asarray(a["v"], dtype=np.float64).copy()
        if a["cols"] is not None:
            keep = np.zeros(a["K"], bool)
            keep[a["cols"]] = True
            v[~keep] = 0
        return M @ v
    if kind == "cat_dense_sandwich":
        M = sel(onehot(a["codes"], a["K"], a["drop_first"]), a["rows"], None)
        Y = sel(a["Y"], a["rows"], a["j_cols"])
        return M.T @ (rsel(a["d"], a["rows"])[:, None] * Y)
    if kind == "cat_cat_sandwich":
        Mi = sel(onehot(a["ic"], a["Ki"], a["i_drop_first"]), a["rows"], None)
        Mj = sel(onehot(a["jc"], a["Kj"], a["j_drop_first"]), a["rows"], None)
        return Mi.T @ (rsel(a["d"], a["rows"])[:, None] * Mj)
    if kind == "cat_sparse_sandwich":
        M = sel(onehot(a["codes"], a["K"], a["drop_first"]), a["rows"], None)
        S = sel(a["A"], a["rows"], a["s_cols"])
        return M.T @ (rsel(a["d"], a["rows"])[:, None] * S)
    raise KeyError(kind)


def constructor_frame(seed=5, n=60):
    """The data frame both sides ingest (golden generator and tests/test_gpu_constructor.py)."""
    import pandas as pd

    rng = np.random.default_rng(seed)
    return pd.DataFrame({
        "dense_a": rng.standard_normal(n),
        "sparse_a": np.where(rng.random(n) < 0.05, rng.standard_normal(n), 0.0),
        "cat_big": pd.Categorical(rng.choice(list("abcdefg"), size=n)),
        "flag": rng.random(n) < 0.5,
        "cat_small": pd.Categorical(rng.choice(["u", "v"], size=n)),
        "dense_b": rng.integers(0, 5, size=n),
        "sparse_b": np.where(rng.random(n) < 0.08, 1.0, 0.0),
    })


def tolerance(dtype) -> float:
    """north_star: 1e-5 relative fp64 / 1e-3 fp32, normwise (max|delta| / max|ref|)."""
    return 1e-3 if np.dtype(dtype) == np.float32 else 1e-5


def assert_close(got, ref, dtype, what=""):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} != {ref.shape}"
    if ref.size == 0:
        return
    scale = max(np.abs(ref).max(), 1e-30)
    err = np.abs(got - ref).max() / scale
    assert err <= tolerance(dtype), f"{what}: normwise error {err:.3e} > {tolerance(dtype):.0e}"
