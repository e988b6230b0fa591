#!/usr/bin/env python
"""bench.py — sandwich (X^T diag(d) X) throughput on B200, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config c5|c2|c3|c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload (BASELINE.json configs[4], SURVEY.md §8d "C5"): SplitMatrix with 128 dense
columns, 3 CSC blocks of 1000 columns at density 1e-3 (merged by the constructor into one
3000-column sparse block) and 5 categorical blocks with 10/50/200/1000/2000 levels, p = 6388,
float32, n = 4e7 rows in total, row-sharded contiguously over the N ranks (strong scaling: the
total row count is fixed), one NCCL allreduce of the flat block workspace per step.
``--config c2|c3|c4`` selects the other GPU configs of BASELINE.json (dense f32 SYRK,
categorical histogram, CSC sparse f64 + its dense cross term), one line each.

A "step" is one ``X.sandwich(d)``.  ``value`` = algorithmic GFLOP/s with everything resident
in HBM; ``e2e`` = the same through the public API with ``d`` in pinned host memory and the
result copied back to pinned host memory inside the timed region.  The matrix X is the
resident operator (built once, like the reference's cached CSR); ``d`` is the per-step input.

``parity``: outside the timed region a bounded row sample of the same workload is built ONCE on
the host; the identical arrays go to the reference's CPU implementation (the ``cpu_baseline``
leg) and, uploaded (row-sharded over the ranks when N > 1), through the same tabmat_b200 code
path as the timed run; the two results are compared normwise.

``--impl reference`` times the reference's own stock ``tabmat.<Class>.sandwich`` (the
unmodified Python package + its Cython/C++ kernels built from source under oracle/_ref) on the
same bounded sample on all host cores.
"""

from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_TOTAL = 40_000_000
P_DENSE = 128
SPARSE_BLOCKS = 3
SPARSE_COLS = 1000
SPARSE_DENSITY = 1e-3
CAT_LEVELS = (10, 50, 200, 1000, 2000)
P_TOTAL = P_DENSE + SPARSE_BLOCKS * SPARSE_COLS + sum(CAT_LEVELS)
ROW_ORDER_DEFAULT = "sorted"
PARITY_TOL = {"f32": 1e-3, "f64": 1e-5}   # BASELINE.json north_star, normwise


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c5", choices=["c5", "c2", "c3", "c4"],
                    help="BASELINE.json configs[4] (default) / [1] / [2] / [3]")
    ap.add_argument("--n", type=int, default=None, help="total rows (default: the config's)")
    ap.add_argument("--cpu-rows", type=int, default=1_000_000,
                    help="rows of the bounded CPU-baseline / parity sample")
    ap.add_argument("--no-cpu-baseline", action="store_true",
                    help="skip the reference CPU leg (then parity is skipped too)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="print per-block times to stderr")
    ap.add_argument("--row-order", default=os.environ.get("TABMAT_B200_BENCH_ROW_ORDER", ROW_ORDER_DEFAULT),
                    choices=["original", "sorted"],
                    help="c5 only. 'sorted': the resident matrix is stored with its rows sorted "
                         "by the many-level categorical codes (tabmat_b200.RowSortedMatrix, built "
                         "once like the reference's cached CSR); d arrives in the caller's order "
                         "and is permuted inside the timed region")
    return ap.parse_args()


def host_cores() -> int:
    return len(os.sched_getaffinity(0))


# =========================================================================================
# workloads
# =========================================================================================
class Workload:
    """One BASELINE.json config: device-side generation for the timed run, a host-side sample
    for the reference / parity leg, algorithmic work."""

    key = ""
    metric = ""
    dtype = "f32"
    n_default = N_TOTAL
    bound = "hbm"

    def __init__(self, n):
        self.n = int(n or self.n_default)

    @property
    def np_dtype(self):
        return np.float32 if self.dtype == "f32" else np.float64

    # -- to be provided ---------------------------------------------------------------
    def describe(self):
        raise NotImplementedError

    def device_matrix(self, n_local, seed, device):
        """-> (matrix, d, flops, info) generated in HBM."""
        raise NotImplementedError

    def host_sample(self, rows, seed):
        """-> dict(parts=..., d=..., flops=...) of numpy / scipy objects."""
        raise NotImplementedError

    def ours_from_sample(self, sample, lo, hi):
        raise NotImplementedError

    def ref_from_sample(self, tabmat, sample):
        raise NotImplementedError

    def dense_result(self, res):
        """result of sandwich -> dense float64 ndarray (for the parity comparison)"""
        import scipy.sparse as sps

        if sps.issparse(res):
            return np.asarray(res.todense(), dtype=np.float64)
        if hasattr(res, "detach"):
            res = res.detach().cpu().numpy()
        res = np.asarray(res, dtype=np.float64)
        return np.diag(res) if res.ndim == 1 else res


def split_flops(sparse_row_counts_sum, sparse_row_counts_sq_sum, n):
    """FLOPs = sum_k r_k (r_k + 1), r_k = non-zeros of row k."""
    base = P_DENSE + len(CAT_LEVELS)
    # sum (base + s)(base + s + 1) = n*base*(base+1) + (2*base+1)*sum s + sum s^2
    return n * base * (base + 1) + (2 * base + 1) * sparse_row_counts_sum + sparse_row_counts_sq_sum


def split_bytes(n, nnz, fsize=4):
    """every block array + d read once + the p x p float64 result written once"""
    return (n * (P_DENSE * fsize + len(CAT_LEVELS) * 4 + fsize) + nnz * (fsize + 4) + 4 * (n + 1)
            + P_TOTAL * P_TOTAL * 8)


class C5(Workload):
    key = "c5"
    metric = "SplitMatrix sandwich GFLOP/s"
    dtype = "f32"
    n_default = N_TOTAL

    def describe(self):
        return ("SplitMatrix 128 dense + 3x1000 CSC @1e-3 + cat{10,50,200,1000,2000}, p=6388, "
                "f32, n=%d total rows" % self.n)

    def device_matrix(self, n, seed, device):
        import torch

        import tabmat_b200 as tm

        g = torch.Generator(device=device).manual_seed(seed)
        X = torch.randn((n, P_DENSE), device=device, dtype=torch.float32, generator=g)
        mats = [tm.DenseMatrix(X)]
        nnz_total = 0
        counts_total = torch.zeros(n, dtype=torch.int64, device=device)
        for _ in range(SPARSE_BLOCKS):
            nnz = int(n * SPARSE_COLS * SPARSE_DENSITY)
            r = torch.randint(0, n, (nnz,), device=device, generator=g, dtype=torch.int64)
            c = torch.randint(0, SPARSE_COLS, (nnz,), device=device, generator=g, dtype=torch.int64)
            key = torch.unique(r * SPARSE_COLS + c)  # sorted by (row, col), duplicates dropped
            del r, c
            rows = torch.div(key, SPARSE_COLS, rounding_mode="floor")
            cols = (key - rows * SPARSE_COLS).to(torch.int32)
            del key
            cnt = torch.bincount(rows, minlength=n)
            counts_total += cnt
            indptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
            indptr[1:] = torch.cumsum(cnt, 0)
            vals = torch.randn(cols.numel(), device=device, dtype=torch.float32, generator=g)
            nnz_total += int(cols.numel())
            mats.append(tm.SparseMatrix.from_device_csr(vals, cols, indptr.to(torch.int32),
                                                        (n, SPARSE_COLS)))
            del rows, cnt, indptr
        for K in CAT_LEVELS:
            codes = torch.randint(0, K, (n,), device=device, generator=g, dtype=torch.int32)
            mats.append(tm.CategoricalMatrix(codes, categories=np.arange(K), dtype=np.float32))
        cf = counts_total.to(torch.float64)
        s_sum, s_sq = float(cf.sum().item()), float((cf * cf).sum().item())
        del counts_total, cf
        Xs = tm.SplitMatrix(mats)
        torch.cuda.empty_cache()
        d = torch.rand(n, device=device, dtype=torch.float32, generator=g)
        return Xs, d, split_flops(s_sum, s_sq, n), {"nnz": nnz_total}

    def host_sample(self, rows, seed):
        import scipy.sparse as sps

        n = rows
        rng = np.random.default_rng(seed)
        X = rng.standard_normal((n, P_DENSE), dtype=np.float32)
        mats = []
        for _ in range(SPARSE_BLOCKS):
            nnz = int(n * SPARSE_COLS * SPARSE_DENSITY)
            r = rng.integers(0, n, size=nnz)
            c = rng.integers(0, SPARSE_COLS, size=nnz)
            v = rng.standard_normal(nnz, dtype=np.float32)
            mats.append(sps.csc_matrix((v, (r, c)), shape=(n, SPARSE_COLS)))
        cats = [rng.integers(0, K, size=n).astype(np.int32) for K in CAT_LEVELS]
        d = rng.random(n, dtype=np.float32)
        counts = sum(np.diff(m.tocsr().indptr) for m in mats).astype(np.float64)
        return {"X": X, "sparse": mats, "cats": cats, "d": d,
                "flops": split_flops(counts.sum(), (counts ** 2).sum(), n)}

    def ours_from_sample(self, s, lo, hi):
        import tabmat_b200 as tm

        mats = [tm.DenseMatrix(np.ascontiguousarray(s["X"][lo:hi]))]
        for A in s["sparse"]:
            mats.append(tm.SparseMatrix(A.tocsr()[lo:hi].tocsc()))
        for codes, K in zip(s["cats"], CAT_LEVELS):
            mats.append(tm.CategoricalMatrix(codes[lo:hi], categories=np.arange(K), dtype=np.float32))
        return tm.SplitMatrix(mats)

    def ref_from_sample(self, tabmat, s):
        mats = [tabmat.DenseMatrix(s["X"])]
        mats += [tabmat.SparseMatrix(A) for A in s["sparse"]]
        mats += [tabmat.CategoricalMatrix(codes, categories=np.arange(K), dtype=np.float32)
                 for codes, K in zip(s["cats"], CAT_LEVELS)]
        return tabmat.SplitMatrix(mats)   # merges the sparse blocks (split_matrix.py:85-141)


class C2(Workload):
    key = "c2"
    metric = "DenseMatrix sandwich GFLOP/s"
    dtype = "f32"
    n_default = 10_000_000
    P = 256

    @staticmethod
    def f_order():
        # TABMAT_B200_BENCH_C2_ORDER=F: a column-major device tensor (served by the F-order TMA
        # box of the tcgen05 kernel, no transpose anywhere)
        return os.environ.get("TABMAT_B200_BENCH_C2_ORDER", "C").upper() == "F"

    def describe(self):
        order = "F" if self.f_order() else "C"
        return (f"DenseMatrix.sandwich f32, n={self.n}, p={self.P}, {order}-order "
                "(BASELINE.json configs[1])")

    def device_matrix(self, n, seed, device):
        import torch

        import tabmat_b200 as tm

        g = torch.Generator(device=device).manual_seed(seed)
        if self.f_order():
            X = torch.randn((self.P, n), device=device, dtype=torch.float32, generator=g).t()
        else:
            X = torch.randn((n, self.P), device=device, dtype=torch.float32, generator=g)
        d = torch.rand(n, device=device, dtype=torch.float32, generator=g)
        return tm.DenseMatrix(X), d, float(n) * self.P * (self.P + 1), {}

    def host_sample(self, rows, seed):
        rng = np.random.default_rng(seed)
        X = rng.standard_normal((rows, self.P), dtype=np.float32)
        return {"X": X, "d": rng.random(rows, dtype=np.float32),
                "flops": float(rows) * self.P * (self.P + 1)}

    def ours_from_sample(self, s, lo, hi):
        import tabmat_b200 as tm

        return tm.DenseMatrix(np.ascontiguousarray(s["X"][lo:hi]))

    def ref_from_sample(self, tabmat, s):
        return tabmat.DenseMatrix(s["X"])

    def bytes(self, n, info):
        return n * self.P * 4 + n * 4 + self.P * self.P * 4


class C3(Workload):
    key = "c3"
    metric = "CategoricalMatrix sandwich GB/s"
    dtype = "f32"
    n_default = 10_000_000
    K = 2000

    def describe(self):
        return f"CategoricalMatrix.sandwich f32 d, n={self.n}, {self.K} levels (BASELINE.json configs[2])"

    def device_matrix(self, n, seed, device):
        import torch

        import tabmat_b200 as tm

        g = torch.Generator(device=device).manual_seed(seed)
        codes = torch.randint(0, self.K, (n,), device=device, generator=g, dtype=torch.int32)
        d = torch.rand(n, device=device, dtype=torch.float32, generator=g)
        return (tm.CategoricalMatrix(codes, categories=np.arange(self.K), dtype=np.float32), d,
                float(n), {})

    def host_sample(self, rows, seed):
        rng = np.random.default_rng(seed)
        return {"codes": rng.integers(0, self.K, size=rows).astype(np.int32),
                "d": rng.random(rows, dtype=np.float32), "flops": float(rows)}

    def ours_from_sample(self, s, lo, hi):
        import tabmat_b200 as tm

        return tm.CategoricalMatrix(s["codes"][lo:hi], categories=np.arange(self.K), dtype=np.float32)

    def ref_from_sample(self, tabmat, s):
        return tabmat.CategoricalMatrix(s["codes"], categories=np.arange(self.K), dtype=np.float32)

    def bytes(self, n, info):
        return n * 8 + self.K * 4


class _SparsePlusCross:
    """The two calls BASELINE.json configs[3] names, as one object with a ``sandwich``:
    ``A.sandwich(d)`` (sparse_matrix.py:175-185) and the dense x sparse cross term
    ``A._cross_sandwich(B, d)`` (sparse_matrix.py:206-229); the result is the (p_s, p_s + q)
    matrix [A^T D A | A^T D B].  Works for tabmat_b200 and for the reference package alike."""

    def __init__(self, A, B):
        self.A, self.B = A, B
        self.shape = (A.shape[0], A.shape[1] + B.shape[1])
        self.dtype = A.dtype

    def sandwich(self, d, rows=None, cols=None):
        s = self.A.sandwich(d, rows, None)
        c = self.A._cross_sandwich(self.B, d, rows, None, None)
        if hasattr(s, "is_cuda"):
            import torch

            return torch.cat([s, c], dim=1)
        return np.hstack([np.asarray(s), np.asarray(c)])


class C4(Workload):
    """CSC sparse f64 self sandwich + the dense x sparse cross term (BASELINE.json configs[3])."""

    key = "c4"
    metric = "SparseMatrix sandwich + dense cross GFLOP/s"
    dtype = "f64"
    n_default = 10_000_000
    P = 5000
    Q = 128
    NNZ_PER_ROW = 5

    def describe(self):
        return (f"SparseMatrix.sandwich (CSC {self.P} cols, ~{self.NNZ_PER_ROW} nnz/row) + its "
                f"cross term with a dense n x {self.Q} block, f64, n={self.n} "
                "(BASELINE.json configs[3])")

    def _flops(self, counts_sum, counts_sq_sum, n):
        # sparse self: sum s_k (s_k + 1); cross: 2 nnz q
        return counts_sum + counts_sq_sum + 2 * self.Q * counts_sum

    def device_matrix(self, n, seed, device):
        import torch

        import tabmat_b200 as tm

        g = torch.Generator(device=device).manual_seed(seed)
        X = torch.randn((n, self.Q), device=device, dtype=torch.float64, generator=g)
        nnz = n * self.NNZ_PER_ROW
        r = torch.randint(0, n, (nnz,), device=device, generator=g, dtype=torch.int64)
        c = torch.randint(0, self.P, (nnz,), device=device, generator=g, dtype=torch.int64)
        key = torch.unique(r * self.P + c)
        del r, c
        rows = torch.div(key, self.P, rounding_mode="floor")
        cols = (key - rows * self.P).to(torch.int32)
        del key
        cnt = torch.bincount(rows, minlength=n)
        indptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
        indptr[1:] = torch.cumsum(cnt, 0)
        vals = torch.randn(cols.numel(), device=device, dtype=torch.float64, generator=g)
        A = tm.SparseMatrix.from_device_csr(vals, cols, indptr.to(torch.int32), (n, self.P))
        cf = cnt.to(torch.float64)
        fl = self._flops(float(cf.sum().item()), float((cf * cf).sum().item()), n)
        nnz = int(cols.numel())
        del rows, cnt, indptr, cf
        d = torch.rand(n, device=device, dtype=torch.float64, generator=g)
        return _SparsePlusCross(A, tm.DenseMatrix(X)), d, fl, {"nnz": nnz}

    def host_sample(self, rows, seed):
        import scipy.sparse as sps

        n = rows
        rng = np.random.default_rng(seed)
        X = rng.standard_normal((n, self.Q))
        nnz = n * self.NNZ_PER_ROW
        A = sps.csc_matrix((rng.standard_normal(nnz), (rng.integers(0, n, size=nnz),
                                                       rng.integers(0, self.P, size=nnz))),
                           shape=(n, self.P))
        counts = np.diff(A.tocsr().indptr).astype(np.float64)
        return {"X": X, "A": A, "d": rng.random(n),
                "flops": self._flops(counts.sum(), (counts ** 2).sum(), n)}

    def ours_from_sample(self, s, lo, hi):
        import tabmat_b200 as tm

        return _SparsePlusCross(tm.SparseMatrix(s["A"].tocsr()[lo:hi].tocsc()),
                                tm.DenseMatrix(np.ascontiguousarray(s["X"][lo:hi])))

    def ref_from_sample(self, tabmat, s):
        return _SparsePlusCross(tabmat.SparseMatrix(s["A"]), tabmat.DenseMatrix(s["X"]))

    def bytes(self, n, info):
        # sparse self 9.2e8 B + cross 1.0965e10 B at n = 1e7
        self_b = info["nnz"] * 12 + 4 * (n + 1) + n * 8 + self.P * self.P * 8
        cross_b = n * self.Q * 8 + info["nnz"] * 12 + 4 * (n + 1) + n * 8 + self.P * self.Q * 8
        return self_b + cross_b


WORKLOADS = {"c5": C5, "c2": C2, "c3": C3, "c4": C4}


# =========================================================================================
# reference arm / cpu baseline (host cores)
# =========================================================================================
def cpu_reference_run(wl: Workload, n_rows, steps, warmup, keep_result=False, sample=None):
    """Time the reference's own CPU implementation of the path on an n_rows sample of the
    workload.  Returns a dict (with the p x p result under "result" when keep_result)."""
    from oracle import ref_loader

    cores = host_cores()
    threads = ref_loader.set_omp_threads(cores)   # torchrun exports OMP_NUM_THREADS=1
    if sample is None:
        sample = wl.host_sample(n_rows, seed=4)
    kind, how, why = "reference", None, None
    mat = None
    try:
        tabmat = ref_loader.import_installed_package()
        mat = wl.ref_from_sample(tabmat, sample)
        how = ("stock tabmat.%s.sandwich (unmodified reference Python package + its Cython/C++ "
               "kernels built from source, stand-in xsimd/jemalloc layer; oracle/_ref/tabmat)"
               % type(mat).__name__)
    except Exception as e:  # package not installed / not importable
        why = repr(e)
    if mat is None and wl.key == "c5":
        try:
            from oracle.ref_split import RefSplit
            import scipy.sparse as sps

            ext = ref_loader.load_ext()
            blocks = [("dense", sample["X"]), ("sparse", sps.hstack(sample["sparse"], format="csc"))]
            blocks += [("cat", c, K) for c, K in zip(sample["cats"], CAT_LEVELS)]
            mat = RefSplit(blocks, ext)
            how = ("reference Cython/C++ kernels (oracle/_ref) driven by the block loop of "
                   "split_matrix.py:324-356 restated in oracle/ref_split.py")
        except Exception as e:
            why = f"{why}; {e!r}"
    if mat is None:
        if wl.key != "c5":
            raise RuntimeError(f"reference implementation unavailable: {why}")
        return _cpu_port_run(wl, sample, steps, why)
    d = sample["d"]
    res = None
    for _ in range(max(1, warmup)):
        res = mat.sandwich(d)
    times = []
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        res = mat.sandwich(d)
        times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    flops = sample["flops"]
    out = dict(
        value=flops / t / 1e9, unit="GFLOP/s", cores=cores, omp_threads=threads, kind=kind,
        sample=(f"{n_rows} rows of the same workload ({wl.dtype}), mean of {len(times)} after "
                f"{max(1, warmup)} warm-up, {threads} OpenMP threads on {cores} host cores; {how}"),
        ms_per_step=t * 1e3, ms_min=float(np.min(times)) * 1e3, rows=n_rows,
        ms_extrapolated_full=t * 1e3 * (wl.n / n_rows),
    )
    if keep_result:
        out["result"] = wl.dense_result(res)
    return out


def _cpu_port_run(wl, sample, steps, why):
    """Fallback (c5 only): the single-threaded C restatement (oracle/tabmat_oracle.c)."""
    import scipy.sparse as sps

    from oracle import c_oracle as orc

    n_rows = min(len(sample["d"]), 100_000)
    X = sample["X"][:n_rows]
    A = sps.hstack(sample["sparse"], format="csr")[:n_rows].tocsc()
    cats = [(c[:n_rows], K) for c, K in zip(sample["cats"], CAT_LEVELS)]
    d = sample["d"][:n_rows]

    def once():
        orc.dense_sandwich(X, d)
        orc.sparse_sandwich(A, d)
        orc.csr_dense_sandwich(A, X, d)
        for codes, K in cats:
            orc.cat_sandwich(codes, d, None, K)
            orc.cat_dense_sandwich(codes, K, d, X)
            orc.cat_sparse_sandwich(codes, K, d, A)
        for i in range(len(cats)):
            for j in range(i + 1, len(cats)):
                orc.cat_cat_sandwich(cats[i][0], cats[j][0], cats[i][1], cats[j][1], d)

    once()
    times = []
    for _ in range(max(1, min(steps, 3))):
        t0 = time.perf_counter()
        once()
        times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    flops = sample["flops"] * n_rows / len(sample["d"])
    return dict(value=flops / t / 1e9, unit="GFLOP/s", cores=1, omp_threads=1, kind="port",
                sample=f"{n_rows} rows, single-threaded C restatement (oracle/_ref unusable: {why})",
                ms_per_step=t * 1e3, ms_min=float(np.min(times)) * 1e3, rows=n_rows,
                ms_extrapolated_full=t * 1e3 * (wl.n / n_rows))


CPU_KEYS = ("value", "unit", "cores", "omp_threads", "kind", "sample")


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(wl, args.cpu_rows, args.steps, args.warmup)
    unit = "GFLOP/s"
    value = r["value"]
    if wl.key == "c3":   # the categorical config is quoted in GB/s of algorithmic bytes
        unit = "GB/s"
        value = wl.bytes(r["rows"], {}) / (r["ms_per_step"] * 1e-3) / 1e9
    cb = {k: r[k] for k in CPU_KEYS}
    cb["value"], cb["unit"] = value, unit
    line = {
        "impl": "reference", "metric": wl.metric, "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
        "config": {"workload": wl.describe(), "sample_rows": r["rows"],
                   "ms_extrapolated_to_full_n": r["ms_extrapolated_full"],
                   "omp_threads": r["omp_threads"], "host_cores": r["cores"]},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# =========================================================================================
# our arm
# =========================================================================================
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def block_breakdown(Xs, d, reps=3):
    """CUDA-event time of every block computation of one sandwich (current stream)."""
    import torch

    from tabmat_b200.categorical_matrix import CategoricalMatrix

    mats = Xs.matrices
    names = []
    for m in mats:
        names.append({"DenseMatrix": "dense", "SparseMatrix": "sparse"}.get(type(m).__name__,
                                                                           f"cat{m.shape[1]}"))
    out = {}

    def timed(label, fn):
        fn()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        out[label] = float(np.mean(ts))

    ws = Xs._sandwich_blocks_dev(d, None)
    if ws is not None:
        timed("native.blocks(all)", lambda: Xs._sandwich_blocks_dev(d, None))
        timed("native.assemble", lambda: Xs._assemble_dev(ws))
    fused_pairs = set(Xs._fused_dense_cross(d, None).keys())
    if fused_pairs:
        timed("dense.cross_fused", lambda: Xs._fused_dense_cross(d, None))
    for i, mi in enumerate(mats):
        if isinstance(mi, CategoricalMatrix):
            timed(f"{names[i]}.self", lambda mi=mi: mi._sandwich_diag(d))
        else:
            timed(f"{names[i]}.self", lambda mi=mi: mi.sandwich(d))
        for j in range(i + 1, len(mats)):
            if (i, j) in fused_pairs or (j, i) in fused_pairs:
                continue
            timed(f"{names[i]}x{names[j]}", lambda mi=mi, mj=mats[j]: mi._cross_sandwich(mj, d, None, None, None))
    return out


def load_traffic(kernel_key, rows_local):
    """ncu dram__bytes (read + write) of the dominant kernel from a SEPARATE ``ncu --set full``
    capture of this command (profiles/traffic.json: bytes at the row count it was captured
    at), scaled linearly to this run's rows per GPU.  None when there is no capture."""
    try:
        t = json.loads((ROOT / "profiles" / "traffic.json").read_text()).get(kernel_key)
        if not t:
            return None, None
        return (float(t["dram_bytes"]) * rows_local / float(t["rows"]),
                f"ncu --set full, separate run ({t['source']}), {t['dram_bytes']:.4g} B at "
                f"{t['rows']} rows scaled to {rows_local} rows per GPU")
    except Exception:
        return None, None


def parity_check(wl, args, world, rank, device, ref_result, sample, build_local):
    """Same host arrays through tabmat_b200 (row-sharded over the ranks like the timed run) vs
    the reference result computed on rank 0.  Returns the parity dict on rank 0."""
    import torch
    import torch.distributed as dist

    from tabmat_b200.distributed import RowShardedMatrix, shard_bounds

    rows = len(sample["d"])
    lo, hi = shard_bounds(rows, world, rank)
    Xl = build_local(wl.ours_from_sample(sample, lo, hi))
    S = RowShardedMatrix(Xl, rows, pack=True,
                         reduce_dtype=torch.float32 if wl.dtype == "f32" else None)
    d_local = torch.from_numpy(sample["d"][lo:hi]).to(device)
    got = S.sandwich(d_local)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if rank != 0:
        return None
    got = wl.dense_result(got)
    ref = ref_result
    scale = float(np.abs(ref).max())
    err = float(np.abs(got - ref).max() / scale)
    par = {"max_normwise_err": err, "tol": PARITY_TOL[wl.dtype], "ok": bool(err <= PARITY_TOL[wl.dtype]),
           "rows": rows, "ranks": world,
           "against": "reference CPU result on the identical host arrays (cpu_baseline leg)",
           "definition": "max|ours - ref| / max|ref| over the whole result"}
    if wl.key == "c5":
        # the dense self block is the one computed in TF32 on the tensor cores
        dd = np.abs(got[:P_DENSE, :P_DENSE] - ref[:P_DENSE, :P_DENSE])
        rr = np.abs(ref[:P_DENSE, :P_DENSE])
        par["dense_block_normwise_err"] = float(dd.max() / rr.max())
        par["dense_block_max_elementwise_rel_err"] = float((dd / np.maximum(rr, 1e-30)).max())
        rest = np.abs(got - ref)
        rest[:P_DENSE, :P_DENSE] = 0
        par["other_blocks_normwise_err"] = float(rest.max() / scale)
    elif wl.key == "c2":
        dd, rr = np.abs(got - ref), np.abs(ref)
        offd = ~np.eye(ref.shape[0], dtype=bool)
        par["max_elementwise_rel_err"] = float((dd / np.maximum(rr, 1e-30)).max())
        par["offdiag_err_over_max_offdiag"] = float(dd[offd].max() / rr[offd].max())
    return par


def main():
    args = parse_args()
    wl = WORKLOADS[args.config](args.n)
    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    import torch
    import torch.distributed as dist

    import tabmat_b200 as tm
    from tabmat_b200.distributed import RowShardedMatrix, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (tabmat_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650"
    tdt = torch.float32 if wl.dtype == "f32" else torch.float64
    fsize = 4 if wl.dtype == "f32" else 8

    sorted_rows = wl.key == "c5" and args.row_order == "sorted"

    def build_local(X):
        if sorted_rows:
            return tm.RowSortedMatrix.from_split(X)
        return X

    lo, hi = shard_bounds(wl.n, world, rank)
    n_local = hi - lo
    Xs, d, flops_local, info = wl.device_matrix(n_local, seed=1000 + rank, device=device)
    nnz_local = info.get("nnz", 0)
    Xs = build_local(Xs)
    torch.cuda.empty_cache()
    S = RowShardedMatrix(Xs, wl.n, pack=True, reduce_dtype=tdt)
    p = Xs.shape[1]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # inputs smaller than the 126 MB L2 (c3: 80 MB) are evicted between timed iterations by
    # writing a 256 MB buffer; the steps are then timed one by one (the flush is outside the
    # event pairs) and the K times added up
    flush = None
    if wl.key == "c3":
        flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device)

    def timed_loop(fn, steps):
        barrier()
        if flush is None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
            total = e0.elapsed_time(e1)
        else:
            evs = []
            for _ in range(steps):
                flush.fill_(1.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                evs.append((e0, e1))
            barrier()
            total = sum(a.elapsed_time(b) for a, b in evs)
        ms = torch.tensor([total], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step():
        return S.sandwich(d)

    warm = max(3, args.warmup)
    for _ in range(warm):
        step()
    graph_note = None
    if wl.key == "c3" and world == 1 and os.environ.get("TABMAT_B200_BENCH_GRAPH", "1") != "0":
        # The 80 MB histogram runs ~25 us on the device, less than the Python + ctypes cost of one
        # call: the step (memset + kernel, through the same C-ABI entry) is captured once in a
        # CUDA graph and replayed, so that the timed loop measures the device work
        try:
            gstream = torch.cuda.Stream()
            gstream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(gstream):
                step()
            torch.cuda.current_stream().wait_stream(gstream)
            graph = torch.cuda.CUDAGraph()
            tm.reset_launch_count()
            with torch.cuda.graph(graph):
                graph_out = S.sandwich(d)
            graph_kernels = tm.launch_count()   # this library's kernels inside one replay
            eager = step

            def step():  # noqa: F811
                graph.replay()
                return graph_out

            step()
            torch.cuda.synchronize()
            assert torch.allclose(graph_out, eager(), rtol=1e-4), "graph replay differs from the eager call"
            graph_note = "step = replay of a CUDA graph capturing X.sandwich(d) (memset + kernel)"
        except Exception as e:  # pragma: no cover - capture not possible: time the eager call
            graph_note = f"CUDA graph capture failed ({e!r}); eager calls timed"
            step = eager if "eager" in dir() else step
    sampler = ClockSampler(local_rank) if rank == 0 else None
    tm.reset_launch_count()
    tm._lib.lib.tm_split_profile_enable(1)   # CUDA events around the passes of every step
    total_ms = timed_loop(step, args.steps)
    launches = tm.launch_count()
    if graph_note and graph_note.startswith("step = replay"):
        launches = graph_kernels * args.steps   # replays do not pass through the launch counter
    pass_arr = (ctypes.c_float * 3)()
    tm._lib.lib.tm_split_profile_read(pass_arr)
    tm._lib.lib.tm_split_profile_enable(0)
    pass_ms = {"tensor": float(pass_arr[0]), "scatter": float(pass_arr[1]), "index": float(pass_arr[2])}
    plan = int(tm._lib.lib.tm_split_last_plan())   # which forms the step used (see the header)
    clocks = sampler.stop() if sampler else None

    # ---- end to end through the public API: d from pinned host memory, result to pinned host
    d_host = torch.empty(n_local, dtype=tdt).pin_memory()
    d_host.copy_(d)
    e2e_note = ""
    shared = None
    if wl.key == "c5" and world > 1 and os.environ.get("TABMAT_B200_E2E_SHARED", "1") != "0":
        # The p x p result is one object per job.  Every rank places its own 1/N row band of it
        # after the allreduce and copies the band into ONE page-locked host buffer shared by the
        # ranks (/dev/shm + cudaHostRegister): N PCIe links carry the 8 p^2 bytes instead of one.
        from tabmat_b200.distributed import SharedHostResult

        shared = SharedHostResult.for_group(p)

        def e2e_step():
            S.sandwich_into_shared(d_host, shared)

        d2h = p * p * 8
        e2e_note = ("d (each rank's shard) from pinned host memory every step; the p x p float64 "
                    f"result lands in one pinned host buffer shared by the {world} ranks, each "
                    "rank copying its own row band over its own PCIe link (a 1-element allreduce "
                    "after the copies is the completion fence)")
    elif wl.key == "c5":
        out_host = torch.empty((p, p), dtype=torch.float64).pin_memory()
        # one rank; or TABMAT_B200_E2E_SHARED=0: reduce to rank 0 and one copy from there
        e2e_dst = 0 if world > 1 else None

        def e2e_step():
            S.sandwich_into(d_host, out_host, dst=e2e_dst)

        d2h = p * p * 8
        e2e_note = ("d (this rank's shard) from pinned host memory every step; the p x p float64 "
                    "result read to pinned host memory"
                    + (" on rank 0 (reduce to rank 0)" if world > 1 else ""))
    else:
        res0 = S.sandwich(d)
        out_host = torch.empty(tuple(res0.shape), dtype=res0.dtype).pin_memory()

        def e2e_step():
            dd = torch.empty_like(d)
            dd.copy_(d_host, non_blocking=True)
            r = S.sandwich(dd)
            out_host.copy_(r, non_blocking=True)

        d2h = out_host.numel() * out_host.element_size()
        e2e_note = "d from pinned host memory every step; the result read to pinned host memory"

    for _ in range(2):
        e2e_step()
    e2e_ms = timed_loop(e2e_step, args.steps)

    # ---- fused IRLS pass, outside the timed region: Hessian + score from one
    # pass over the dense block vs the reference's two calls
    irls = None
    if wl.key == "c5" and not os.environ.get("TABMAT_B200_BENCH_NO_IRLS"):
        vvec = torch.randn(n_local, device=device, dtype=tdt)

        def ev_time(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            e1.synchronize()
            return e0.elapsed_time(e1) / reps

        t_sand = ev_time(lambda: S.sandwich(d))
        t_fused = ev_time(lambda: S.sandwich_and_transpose_matvec(d, vvec))
        t_tmv = ev_time(lambda: S.transpose_matvec(vvec))
        Hf, gf = S.sandwich_and_transpose_matvec(d, vvec)
        gs = S.transpose_matvec(vvec)
        irls = {"sandwich_ms": t_sand, "sandwich_and_transpose_matvec_ms": t_fused,
                "separate_transpose_matvec_ms": t_tmv,
                "fused_over_sandwich": t_fused / t_sand,
                "score_fused_vs_separate_normwise": float(((gf - gs).abs().max() / gs.abs().max()).item()),
                "note": "X.sandwich_and_transpose_matvec(d, v): Hessian and score of one IRLS step; the "
                        "dense block's share of X^T v rides in the tcgen05 kernel's scale warps"}
        del vvec, Hf, gf, gs
        if world == 1 and hasattr(Xs, "to_stored_order"):
            # a whole IRLS step (eta = X beta, logistic weights on the device, Hessian + score):
            # caller's row order (three n-vector permutations per step) vs stored order
            yv = (torch.rand(n_local, device=device) < 0.4).to(tdt)
            ys = Xs.to_stored_order(yv)
            beta_t = torch.zeros(Xs.shape[1], device=device, dtype=tdt)

            def logistic(resp):
                def fn(eta):
                    mu = torch.sigmoid(eta)
                    return mu * (1 - mu), resp - mu
                return fn

            irls["irls_step_ms"] = ev_time(lambda: tm.irls_step(Xs, beta_t, logistic(yv)))
            irls["irls_step_stored_order_ms"] = ev_time(
                lambda: tm.irls_step(Xs, beta_t, logistic(ys), stored_order=True))
            del yv, ys
        # glum's active-set call: X.sandwich(d, cols=half of the columns) through the selecting
        # assembly (the passes compute whole blocks; only the placement changes)
        if world == 1:
            p_tot = S.shape[1]
            half = np.sort(np.random.default_rng(11).choice(p_tot, size=p_tot // 2,
                                                            replace=False)).astype(np.int32)
            t_half = ev_time(lambda: S.sandwich(d, cols=half))
            irls["cols_half_ms"] = t_half
            irls["cols_half_over_full"] = t_half / t_sand
    if shared is not None:
        shared.close(unlink=rank == 0)
    fl = torch.tensor([float(flops_local), float(nnz_local)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(fl)
    flops, nnz = float(fl[0].item()), float(fl[1].item())

    bd = None
    if rank == 0 and args.breakdown and wl.key == "c5":
        bd = block_breakdown(Xs.mat, Xs._gather(d)) if sorted_rows else block_breakdown(Xs, d)
    if world > 1:
        dist.barrier()

    # ---- reference CPU leg (rank 0) + parity on the identical host arrays (all ranks) --------
    cpu = None
    parity = None
    want_cpu = not args.no_cpu_baseline
    want_parity = want_cpu and not args.no_parity
    del d_host
    sample = None
    if want_parity or (want_cpu and rank == 0):
        sample = wl.host_sample(args.cpu_rows, seed=4)
    if want_cpu and rank == 0:
        try:
            cpu = cpu_reference_run(wl, args.cpu_rows, 3, 1, keep_result=want_parity, sample=sample)
        except Exception as e:  # pragma: no cover
            cpu = {"value": None, "unit": "GFLOP/s", "cores": 0, "omp_threads": 0, "kind": "port",
                   "sample": f"failed: {e!r}"}
    if want_parity:
        ok = torch.tensor([1 if (rank != 0 or (cpu and cpu.get("result") is not None)) else 0],
                          device=device)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 1:
            # the timed matrices are not needed any more: make room for the sample
            del S, Xs, d
            torch.cuda.empty_cache()
            parity = parity_check(wl, args, world, rank, device,
                                  cpu["result"] if rank == 0 else None, sample, build_local)

    if rank == 0:
        ms_step = total_ms / args.steps
        e2e_step_ms = e2e_ms / args.steps
        unit = "GFLOP/s"
        value = flops / (ms_step * 1e-3) / 1e9
        e2e_value = flops / (e2e_step_ms * 1e-3) / 1e9
        config = {"workload": wl.describe(), "rows_per_gpu": n_local, "p": p,
                  "algorithmic_flop_per_step": flops}
        line_extra = {}
        if wl.key == "c5":
            n_oh = sum(1 for K in CAT_LEVELS if K <= 256)
            ps = SPARSE_BLOCKS * SPARSE_COLS
            cats_in_tc, sparse_in_tc, gather = bool(plan & 2), bool(plan & 4), bool(plan & 8)
            n_big = len(CAT_LEVELS) - n_oh
            xb = P_DENSE * 4 + 4   # one dense row + its weight
            pass_bytes = {
                # X + d + the codes of the categoricals it serves in; dense self + those cross
                # blocks (+ the CSR arrays and the dense x sparse block in the fully fused form)
                "tensor": n_local * (xb + 4 * (n_oh + (n_big if cats_in_tc else 0)))
                + (nnz_local * 8 + 4 * (n_local + 1) + 4 * ps * P_DENSE if sparse_in_tc else 0)
                + 4 * (P_DENSE * P_DENSE + sum(K for K in CAT_LEVELS if K <= 256 or cats_in_tc) * P_DENSE),
                # gather form: X + d + the row-blocked CSC (value, row id) in, dense x sparse out;
                # RED form: X + d + many-level codes + CSR in, their cross blocks out
                "scatter": (n_local * xb + nnz_local * 8 + 4 * ps * P_DENSE) if gather else
                (n_local * (xb + 4 * n_big) + nnz_local * 8 + 4 * (n_local + 1)
                 + 4 * (ps + sum(K for K in CAT_LEVELS if K > 256)) * P_DENSE),
                # CSC (value, row id, packed codes) + d + every code vector in; sparse self,
                # cat x sparse, cat x cat out
                "index": nnz_local * 12 + 4 * (n_local + 1) + n_local * (4 * len(CAT_LEVELS) + 4)
                + 4 * (ps * ps + sum(CAT_LEVELS) * ps + sum(CAT_LEVELS)
                       + sum(a * b for i, a in enumerate(CAT_LEVELS) for b in CAT_LEVELS[i + 1:])),
            }
            fused = sparse_in_tc
            kernel_names = {
                "tensor": "k_dense_syrk_tc (tcgen05 SYRK + one-hot MMAs"
                          + (" + scatter warps: run sums of the many-level categoricals" if cats_in_tc else "")
                          + (" + the vector REDs of dense x sparse" if sparse_in_tc else "") + ")",
                "scatter": ("k_csc_dense_gather (dense x sparse by row-blocked gather, one RED per run)"
                            if gather else
                            "k_dense_cross_runs (dense x many-level cats + dense x sparse, vector RED)"),
                "index": "index pass (k_cat_pairs, k_cat_sparse_cols, k_sparse_sandwich)",
            }
            live = {k: v for k, v in pass_ms.items() if v > 0}
            top = max(live, key=live.get)
            top_ms, top_bytes, top_kernel = live[top], pass_bytes[top], kernel_names[top]
            traffic_key = {"tensor": "tensor_cats" if cats_in_tc else "tensor",
                           "scatter": "gather" if gather else "scatter", "index": "index"}[top]
            # L2 RED payload: one 512-byte row per sparse non-zero (+ per row and many-level
            # categorical in the caller's row order); measured L2 atomic peak 6.0 TB/s
            n_red_cat = 0 if sorted_rows else len(CAT_LEVELS) - n_oh
            red_bytes = (n_local * n_red_cat + (0 if gather else nnz_local)) * P_DENSE * 4
            red_ms = pass_ms["tensor"] if (fused or (gather and cats_in_tc)) else pass_ms["scatter"]
            whole_bytes = split_bytes(n_local, nnz_local)
            config.update({
                "nnz_sparse_total": nnz, "l2": "inputs (>20 GB per step) exceed the 126 MB L2",
                "parallelism": f"row-shard x{world} + NCCL allreduce of the flat block workspace (f32)",
                "row_order": (args.row_order if not sorted_rows else
                              "sorted by (cat2000, cat1000) at construction; d permuted inside "
                              "the timed region"),
                "pass_schedule": "serial (tensor, index, scatter)",
                "forms": {"cats_in_tcgen05_kernel": cats_in_tc, "sparse_in_tcgen05_kernel": sparse_in_tc,
                          "dense_x_sparse": "gather" if gather else "red"},
                "whole_step_hbm_gbs": whole_bytes / (ms_step * 1e-3) / 1e9,
                "whole_step_hbm_frac": whole_bytes / (ms_step * 1e-3) / 1e9 / hbm_peak})
            line_extra["passes_ms"] = pass_ms
            line_extra["passes_hbm_frac"] = {
                k: (pass_bytes[k] / (v * 1e-3) / 1e9 / hbm_peak if v > 0 else None)
                for k, v in pass_ms.items()}
            l2_red = None
            if not gather:
                l2_red = {"payload_bytes": red_bytes,
                          "achieved_TBs": red_bytes / (red_ms * 1e-3) / 1e12 if red_ms > 0 else None,
                          "measured_peak_TBs": 6.0,
                          "note": "the vector REDs of dense x sparse are bound by the L2 atomic "
                                  "units (1.9e11 sector-ops/s), not by HBM"}
            else:
                # the gather form reads one 512-byte X row per non-zero through L2
                line_extra["l2_gather"] = {
                    "bytes": nnz_local * P_DENSE * 4,
                    "achieved_TBs": nnz_local * P_DENSE * 4 / (pass_ms["scatter"] * 1e-3) / 1e12,
                    "note": "k_csc_dense_gather: X rows gathered through L2, one RED per (row block, column) run"}
        else:
            top_ms, top_bytes = ms_step, wl.bytes(n_local, info)
            top_kernel = {"c2": "k_dense_syrk_tc (tcgen05 weighted SYRK, 3 lower-triangular 128x128 tiles)",
                          "c3": "k_cat_hist_vec (weighted histogram of the codes, 16-byte loads)",
                          "c4": "whole step: k_sparse_sandwich (CSR outer products, scalar REDs) + k_csr_dense (vector REDs)",
                          }[wl.key]
            traffic_key = wl.key
            l2_red = None
            config["l2"] = ("256 MB written between timed iterations (inputs 80 MB < 126 MB L2), "
                            "outside the per-step event pairs" if flush is not None else
                            "inputs exceed the 126 MB L2")
            config["parallelism"] = f"row-shard x{world}" + (" + NCCL allreduce" if world > 1 else "")
            if graph_note:
                config["launch"] = graph_note
            if wl.key == "c3":
                unit = "GB/s"
                value = wl.bytes(n_local, info) * world / (ms_step * 1e-3) / 1e9
                e2e_value = wl.bytes(n_local, info) * world / (e2e_step_ms * 1e-3) / 1e9
            if wl.key == "c2":
                bf16 = float(peaks.get("bf16_tflops", 1590.0))
                full = 2.0 * n_local * C2.P * C2.P
                line_extra["tensor"] = {
                    "executed_TFLOPs(3 of 4 tiles)": 0.75 * full / (ms_step * 1e-3) / 1e12,
                    "tf32_peak_assumed_TF": bf16 / 2,
                    "frac_of_assumed_tf32_peak": 0.75 * full / (ms_step * 1e-3) / 1e12 / (bf16 / 2),
                    "note": "no TF32 peak is measured on this pool; bf16 / 2 assumed"}
        achieved = top_bytes / (top_ms * 1e-3) / 1e9
        traffic, traffic_src = load_traffic(traffic_key, n_local)
        roofline = {"bound": "hbm", "kernel": top_kernel, "achieved": achieved, "peak": hbm_peak,
                    "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                    "traffic_source": traffic_src, "peak_source": peak_src, "launch_ms": top_ms,
                    "algorithmic_bytes": top_bytes}
        if l2_red:
            roofline["l2_red"] = l2_red
        line = {
            "metric": wl.metric, "value": value, "unit": unit,
            "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
            "config": config,
            "e2e": {"value": e2e_value, "unit": unit, "ms_per_step": e2e_step_ms,
                    "h2d_bytes_per_step": int(wl.n * fsize), "d2h_bytes_per_step": int(d2h),
                    "h2d_bytes_per_step_per_gpu": int(n_local * fsize),
                    "note": e2e_note},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
        }
        line.update(line_extra)
        if bd is not None:
            line["breakdown_ms"] = bd
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu.get(k) for k in CPU_KEYS}
            if wl.key == "c3" and cpu.get("ms_per_step"):
                line["cpu_baseline"]["value"] = wl.bytes(cpu["rows"], {}) / (cpu["ms_per_step"] * 1e-3) / 1e9
                line["cpu_baseline"]["unit"] = "GB/s"
            if "ms_extrapolated_full" in cpu:
                line["cpu_baseline"]["ms_extrapolated_to_full_n"] = cpu["ms_extrapolated_full"]
        if parity is not None:
            line["parity"] = parity
        if irls is not None:
            line["irls_step"] = irls
        if args.breakdown and bd is not None:
            for k, v in sorted(bd.items(), key=lambda kv: -kv[1]):
                print(f"  {k:24s} {v:9.3f} ms", file=sys.stderr)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
