"""Reference CPU path of the SplitMatrix sandwich on host arrays (TEST / BASELINE INFRASTRUCTURE).

``/root/reference`` does not exist on the GPU box, only the reference's compiled kernels in
``oracle/_ref`` do.  This module drives those kernels exactly the way the reference's Python
classes do — one native call per self block and per cross pair, scipy ``csr_matmat`` for the
categorical x sparse pair, numpy ``np.ix_`` placement into a float64 p x p — following
split_matrix.py:324-356, dense_matrix.py:153-178, sparse_matrix.py:175-229 and
categorical_matrix.py:618-838.  It is what ``bench.py --impl reference`` and the
``cpu_baseline`` leg time ("reference source kernels, stand-in xsimd/jemalloc layer").

Blocks are given as a list of ("dense", ndarray) / ("sparse", csc_matrix) / ("cat", codes, K)
with unrestricted rows / cols (the benchmark case).  Nothing under ``tabmat_b200/`` imports
this module.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sps


class RefSplit:
    def __init__(self, blocks, ext):
        self.ext = ext
        self.blocks = []
        off = 0
        n = None
        for b in blocks:
            kind = b[0]
            if kind == "dense":
                X = b[1]
                width = X.shape[1]
                item = dict(kind=kind, X=X, c=bool(X.flags["C_CONTIGUOUS"]))
                n = X.shape[0]
            elif kind == "sparse":
                A = sps.csc_matrix(b[1])
                A.sort_indices()
                R = A.tocsr()  # the reference caches this lazily (sparse_matrix.py:133-143)
                R.sort_indices()
                width = A.shape[1]
                item = dict(kind=kind, csc=A, csr=R)
                n = A.shape[0]
            elif kind == "cat":
                codes, K = np.ascontiguousarray(b[1], dtype=np.int32), int(b[2])
                width = K
                item = dict(kind=kind, codes=codes, K=K)
                n = len(codes)
            else:
                raise ValueError(kind)
            item["idx"] = np.arange(off, off + width)
            item["cols"] = np.arange(width, dtype=np.int32)
            off += width
            self.blocks.append(item)
        self.n = n
        self.p = off
        self.rows = np.arange(n, dtype=np.int32)

    # one native call per block, as the reference's classes make them
    def _self(self, b, d):
        e = self.ext
        if b["kind"] == "dense":
            return e.dense.dense_sandwich(b["X"], d, self.rows, b["cols"])
        if b["kind"] == "sparse":
            return e.sparse.sparse_sandwich(b["csc"], b["csr"], d, self.rows, b["cols"])
        return np.asarray(e.categorical.sandwich_categorical_fast(b["codes"], d, self.rows,
                                                                  d.dtype, b["K"]))

    def _cross(self, bi, bj, d):
        e = self.ext
        ki, kj = bi["kind"], bj["kind"]
        if ki == "dense" and kj == "sparse":
            return e.sparse.csr_dense_sandwich(bj["csr"], bi["X"], d, self.rows, bj["cols"],
                                               bi["cols"]).T
        if ki == "dense" and kj == "cat":
            return e.split.sandwich_cat_dense(bj["codes"], bj["K"], d, bi["X"], self.rows,
                                              bi["cols"], bi["c"], False, False).T
        if ki == "sparse" and kj == "cat":
            # categorical_matrix.py:825-838: scipy csr_matmat
            term_1 = sps.csr_matrix((d, bj["codes"], np.arange(self.n + 1, dtype=int)),
                                    shape=(self.n, bj["K"]))
            return term_1.T.dot(bi["csc"]).toarray().T
        if ki == "cat" and kj == "cat":
            return e.split.sandwich_cat_cat(bi["codes"], bj["codes"], bi["K"], bj["K"], d,
                                            self.rows, d.dtype, False, False, False, False)
        raise TypeError((ki, kj))

    def sandwich(self, d):
        out = np.zeros((self.p, self.p))
        nb = len(self.blocks)
        for i in range(nb):
            bi = self.blocks[i]
            res = self._self(bi, d)
            if bi["kind"] == "cat":
                out[(bi["idx"], bi["idx"])] += res
            else:
                out[np.ix_(bi["idx"], bi["idx"])] = res
            for j in range(i + 1, nb):
                bj = self.blocks[j]
                res = self._cross(bi, bj, d)
                out[np.ix_(bi["idx"], bj["idx"])] = res
                out[np.ix_(bj["idx"], bi["idx"])] = res.T
        return out
