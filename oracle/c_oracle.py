"""numpy-facing wrapper of the C restatement ``oracle/tabmat_oracle.c`` (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module; nothing under ``tabmat_b200/`` does.

Every function takes host numpy arrays, mirrors one ``tm_*`` entry point of
``include/tabmat_b200.h`` (same argument meaning; ``rows`` / ``cols`` None = all) and returns a
fresh numpy array.  The reference file:line each restates is cited in tabmat_oracle.c.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_LIB = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = HERE / "_build" / "libtabmat_oracle.so"
        if not path.exists():
            import importlib.util

            spec = importlib.util.spec_from_file_location("_orc_build", HERE / "build_oracle.py")
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()
        _LIB = C.CDLL(str(path))
    return _LIB


def _suf(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(dtype)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a), dtype=np.int32)


def _n(a, full):
    return int(full) if a is None else int(len(a))


def _layout(X):
    X = np.asarray(X)
    if X.flags["C_CONTIGUOUS"]:
        return X, 1
    if X.flags["F_CONTIGUOUS"]:
        return X, 0
    return np.ascontiguousarray(X), 1


def _call(name, dtype, *args):
    fn = getattr(lib(), f"orc_{name}_{_suf(dtype)}")
    fn.restype = None
    conv = []
    for a in args:
        if isinstance(a, np.ndarray) or a is None:
            conv.append(_p(a))
        elif isinstance(a, (int, np.integer, bool)):
            conv.append(C.c_int64(int(a)))
        else:
            raise TypeError(type(a))
    fn(*conv)


def _ci(x):  # C `int` arguments
    return C.c_int(int(x))


def _raw(name, dtype, args):
    fn = getattr(lib(), f"orc_{name}_{_suf(dtype)}")
    fn.restype = None
    fn(*args)


def dense_sandwich(X, d, rows=None, cols=None):
    X, c = _layout(X)
    n, p = X.shape
    rows, cols = _i32(rows), _i32(cols)
    m = _n(cols, p)
    out = np.zeros((m, m), X.dtype)
    _raw("dense_sandwich", X.dtype, [_p(X), C.c_int64(n), C.c_int64(p), _ci(c), _p(d), _p(rows),
                                     C.c_int64(_n(rows, n)), _p(cols), C.c_int64(m), _p(out)])
    return out


def dense_matvec(X, v, rows=None, cols=None):
    X, c = _layout(X)
    n, p = X.shape
    rows, cols = _i32(rows), _i32(cols)
    out = np.zeros(_n(rows, n), X.dtype)
    _raw("dense_matvec", X.dtype, [_p(X), C.c_int64(n), C.c_int64(p), _ci(c), _p(v), _p(rows),
                                   C.c_int64(_n(rows, n)), _p(cols), C.c_int64(_n(cols, p)),
                                   _p(out)])
    return out


def dense_rmatvec(X, v, rows=None, cols=None):
    X, c = _layout(X)
    n, p = X.shape
    rows, cols = _i32(rows), _i32(cols)
    out = np.zeros(_n(cols, p), X.dtype)
    _raw("dense_rmatvec", X.dtype, [_p(X), C.c_int64(n), C.c_int64(p), _ci(c), _p(v), _p(rows),
                                    C.c_int64(_n(rows, n)), _p(cols), C.c_int64(_n(cols, p)),
                                    _p(out)])
    return out


def dense_sq_dot_weights(X, w, shift):
    X, c = _layout(X)
    n, p = X.shape
    out = np.zeros(p, X.dtype)
    _raw("dense_sq_dot_weights", X.dtype, [_p(X), C.c_int64(n), C.c_int64(p), _ci(c), _p(w),
                                           _p(shift), _p(out)])
    return out


def _csx(A):
    """(data, indices int32, indptr int32) of a scipy csr/csc matrix with sorted indices."""
    A = A.copy()
    A.sort_indices()
    return (np.ascontiguousarray(A.data), A.indices.astype(np.int32), A.indptr.astype(np.int32))


def sparse_sandwich(A_csc, d, rows=None, cols=None):
    n, p = A_csc.shape
    cd, ci, cp = _csx(A_csc.tocsc())
    rd, ri, rp = _csx(A_csc.tocsr())
    rows, cols = _i32(rows), _i32(cols)
    m = _n(cols, p)
    out = np.zeros((m, m), cd.dtype)
    _raw("sparse_sandwich", cd.dtype, [_p(cd), _p(ci), _p(cp), _p(rd), _p(ri), _p(rp),
                                       C.c_int64(n), C.c_int64(p), _p(d), _p(rows),
                                       C.c_int64(_n(rows, n)), _p(cols), C.c_int64(m), _p(out)])
    return out


def csr_dense_sandwich(A, B, d, rows=None, A_cols=None, B_cols=None):
    n, p = A.shape
    rd, ri, rp = _csx(A.tocsr())
    B, c = _layout(B)
    q = B.shape[1]
    rows, A_cols, B_cols = _i32(rows), _i32(A_cols), _i32(B_cols)
    nA, nB = _n(A_cols, p), _n(B_cols, q)
    out = np.zeros((nA, nB), rd.dtype)
    _raw("csr_dense_sandwich", rd.dtype, [_p(rd), _p(ri), _p(rp), C.c_int64(n), C.c_int64(p),
                                          _p(B), C.c_int64(q), _ci(c), _p(d), _p(rows),
                                          C.c_int64(_n(rows, n)), _p(A_cols), C.c_int64(nA),
                                          _p(B_cols), C.c_int64(nB), _p(out)])
    return out


def csr_matvec(A, v, rows=None, cols=None):
    n, p = A.shape
    rd, ri, rp = _csx(A.tocsr())
    rows, cols = _i32(rows), _i32(cols)
    out = np.zeros(_n(rows, n), rd.dtype)
    _raw("csr_matvec", rd.dtype, [_p(rd), _p(ri), _p(rp), C.c_int64(n), C.c_int64(p), _p(v),
                                  _p(rows), C.c_int64(_n(rows, n)), _p(cols),
                                  C.c_int64(_n(cols, p)), _p(out)])
    return out


def csc_rmatvec(A, v, rows=None, cols=None):
    n, p = A.shape
    cd, ci, cp = _csx(A.tocsc())
    rows, cols = _i32(rows), _i32(cols)
    out = np.zeros(_n(cols, p), cd.dtype)
    _raw("csc_rmatvec", cd.dtype, [_p(cd), _p(ci), _p(cp), C.c_int64(n), C.c_int64(p), _p(v),
                                   _p(rows), C.c_int64(_n(rows, n)), _p(cols),
                                   C.c_int64(_n(cols, p)), _p(out)])
    return out


def csc_sq_dot_weights(A, w):
    n, p = A.shape
    cd, ci, cp = _csx(A.tocsc())
    out = np.zeros(p, cd.dtype)
    _raw("csc_sq_dot_weights", cd.dtype, [_p(cd), _p(ci), _p(cp), C.c_int64(p), _p(w), _p(out)])
    return out


def cat_sandwich(codes, d, rows=None, K=None, drop_first=False):
    codes = _i32(codes)
    rows = _i32(rows)
    out = np.zeros(K, d.dtype)
    _raw("cat_sandwich", d.dtype, [_p(codes), _p(d), _p(rows), C.c_int64(_n(rows, len(codes))),
                                   C.c_int64(K), _ci(drop_first), _p(out)])
    return out


def cat_transpose_matvec(codes, v, rows=None, cols=None, K=None, drop_first=False, out=None):
    codes = _i32(codes)
    rows, cols = _i32(rows), _i32(cols)
    if out is None:
        out = np.zeros(K, v.dtype)
    _raw("cat_transpose_matvec", v.dtype, [_p(codes), _p(v), _p(rows),
                                           C.c_int64(_n(rows, len(codes))), _p(cols),
                                           C.c_int64(_n(cols, K)), C.c_int64(K), _ci(drop_first),
                                           _p(out)])
    return out


def cat_matvec(codes, v, cols=None, K=None, drop_first=False, out=None):
    codes = _i32(codes)
    cols = _i32(cols)
    if out is None:
        out = np.zeros(len(codes), v.dtype)
    _raw("cat_matvec", v.dtype, [_p(codes), C.c_int64(len(codes)), _p(v), _p(cols),
                                 C.c_int64(_n(cols, K)), C.c_int64(K), _ci(drop_first), _p(out)])
    return out


def cat_dense_sandwich(codes, K, d, Y, rows=None, j_cols=None, drop_first=False):
    codes = _i32(codes)
    Y, c = _layout(Y)
    n, q = Y.shape
    rows, j_cols = _i32(rows), _i32(j_cols)
    nJ = _n(j_cols, q)
    out = np.zeros((K, nJ), Y.dtype)
    _raw("cat_dense_sandwich", Y.dtype, [_p(codes), C.c_int64(n), C.c_int64(K), _ci(drop_first),
                                         _p(d), _p(Y), C.c_int64(q), _ci(c), _p(rows),
                                         C.c_int64(_n(rows, n)), _p(j_cols), C.c_int64(nJ),
                                         _p(out)])
    return out


def cat_cat_sandwich(ic, jc, Ki, Kj, d, rows=None, i_drop_first=False, j_drop_first=False):
    ic, jc = _i32(ic), _i32(jc)
    rows = _i32(rows)
    out = np.zeros((Ki, Kj), d.dtype)
    _raw("cat_cat_sandwich", d.dtype, [_p(ic), _p(jc), C.c_int64(Ki), C.c_int64(Kj),
                                       _ci(i_drop_first), _ci(j_drop_first), _p(d), _p(rows),
                                       C.c_int64(_n(rows, len(ic))), _p(out)])
    return out


def cat_sparse_sandwich(codes, K, d, A, rows=None, s_cols=None, drop_first=False):
    codes = _i32(codes)
    n, p = A.shape
    rd, ri, rp = _csx(A.tocsr())
    rows, s_cols = _i32(rows), _i32(s_cols)
    nS = _n(s_cols, p)
    out = np.zeros((K, nS), d.dtype)
    _raw("cat_sparse_sandwich", d.dtype, [_p(codes), C.c_int64(n), C.c_int64(K), _ci(drop_first),
                                          _p(d), _p(rd), _p(ri), _p(rp), C.c_int64(p), _p(rows),
                                          C.c_int64(_n(rows, n)), _p(s_cols), C.c_int64(nS),
                                          _p(out)])
    return out
