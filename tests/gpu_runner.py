"""Evaluate a boundary-level case of tests/cases.py on the device, through the C-ABI
(``tabmat_b200.ext.*`` are thin ctypes callers of ``libtabmat_b200.so``)."""

from __future__ import annotations

import numpy as np
import scipy.sparse as sps
import torch


def _dev(x):
    if x is None:
        return None
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _dense(X):
    X = np.asarray(X)
    if X.flags["C_CONTIGUOUS"]:
        return _dev(X)
    return _dev(X.T).t()  # F order kept on the device


def _csr(A):
    from tabmat_b200.ext.sparse import DeviceCSR

    R = sps.csr_matrix(A)
    R.sort_indices()
    row = np.repeat(np.arange(R.shape[0], dtype=np.int32), np.diff(R.indptr))
    return DeviceCSR(_dev(R.data), _dev(R.indices.astype(np.int32)),
                     _dev(R.indptr.astype(np.int32)), _dev(row), R.shape)


def _csc(A):
    from tabmat_b200.ext.sparse import DeviceCSC

    Cm = sps.csc_matrix(A)
    Cm.sort_indices()
    return DeviceCSC(_dev(Cm.data), _dev(Cm.indices.astype(np.int32)),
                     _dev(Cm.indptr.astype(np.int32)), Cm.shape)


def _i32(x):
    return None if x is None else _dev(np.asarray(x, dtype=np.int32))


def run_cuda(kind: str, a: dict) -> np.ndarray:
    from tabmat_b200.ext import categorical as ecat
    from tabmat_b200.ext import dense as edense
    from tabmat_b200.ext import sparse as esparse
    from tabmat_b200.ext import split as esplit

    if kind == "dense_sandwich":
        r = edense.dense_sandwich(_dense(a["X"]), _dev(a["d"]), _i32(a["rows"]), _i32(a["cols"]))
    elif kind == "dense_matvec":
        r = edense.dense_matvec(_dense(a["X"]), _dev(a["v"]), _i32(a["rows"]), _i32(a["cols"]))
    elif kind == "dense_rmatvec":
        r = edense.dense_rmatvec(_dense(a["X"]), _dev(a["v"]), _i32(a["rows"]), _i32(a["cols"]))
    elif kind == "dense_sq_dot_weights":
        r = edense.transpose_square_dot_weights(_dense(a["X"]), _dev(a["w"]), _dev(a["shift"]))
    elif kind == "sparse_sandwich":
        r = esparse.sparse_sandwich(_csr(a["A"]), _dev(a["d"]), _i32(a["rows"]), _i32(a["cols"]))
    elif kind == "csr_dense_sandwich":
        r = esparse.csr_dense_sandwich(_csr(a["A"]), _dense(a["B"]), _dev(a["d"]),
                                       _i32(a["rows"]), _i32(a["A_cols"]), _i32(a["B_cols"]))
    elif kind == "csr_matvec":
        r = esparse.csr_matvec(_csr(a["A"]), _dev(a["v"]), _i32(a["rows"]), _i32(a["cols"]))
    elif kind == "csc_rmatvec":
        r = esparse.csc_rmatvec(_csc(a["A"]), _dev(a["v"]), _i32(a["rows"]), _i32(a["cols"]))
    elif kind == "csc_sq_dot_weights":
        r = esparse.transpose_square_dot_weights(_csc(a["A"]), _dev(a["w"]))
    elif kind == "cat_sandwich":
        r = ecat.sandwich_categorical(_i32(a["codes"]), _dev(a["d"]), _i32(a["rows"]), a["K"],
                                      a["drop_first"])
    elif kind == "cat_transpose_matvec":
        r = torch.zeros(a["K"], dtype=_dev(a["v"]).dtype, device="cuda")
        ecat.transpose_matvec(_i32(a["codes"]), _dev(a["v"]), a["K"], _i32(a["rows"]),
                              _i32(a["cols"]), r, a["drop_first"])
    elif kind == "cat_matvec":
        n = len(a["codes"])
        r = torch.zeros(n, dtype=_dev(a["v"]).dtype, device="cuda")
        ecat.matvec(_i32(a["codes"]), _dev(a["v"]), n, _i32(a["cols"]), a["K"], r,
                    a["drop_first"])
    elif kind == "cat_dense_sandwich":
        r = esplit.sandwich_cat_dense(_i32(a["codes"]), a["K"], _dev(a["d"]), _dense(a["Y"]),
                                      _i32(a["rows"]), _i32(a["j_cols"]), a["drop_first"])
    elif kind == "cat_cat_sandwich":
        r = esplit.sandwich_cat_cat(_i32(a["ic"]), _i32(a["jc"]), a["Ki"], a["Kj"], _dev(a["d"]),
                                    _i32(a["rows"]), a["i_drop_first"], a["j_drop_first"])
    elif kind == "cat_sparse_sandwich":
        r = esplit.sandwich_cat_sparse(_i32(a["codes"]), a["K"], _dev(a["d"]), _csr(a["A"]),
                                       _i32(a["rows"]), _i32(a["s_cols"]), a["drop_first"])
    else:
        raise KeyError(kind)
    torch.cuda.synchronize()
    return r.cpu().numpy()
