#!/usr/bin/env python
"""bench.py — SplitMatrix sandwich (X^T diag(d) X) throughput on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[4], SURVEY.md §8d "C5"): SplitMatrix with 128 dense columns,
3 CSC blocks of 1000 columns at density 1e-3 (merged by the constructor into one 3000-column
sparse block) and 5 categorical blocks with 10/50/200/1000/2000 levels, p = 6388, float32,
n = 4e7 rows in total, row-sharded contiguously over the N ranks (strong scaling: the total
row count is fixed), one NCCL allreduce of the packed p x p per step.

A "step" is one ``X.sandwich(d)``.  ``value`` = algorithmic GFLOP/s with everything resident
in HBM; ``e2e`` = the same through the public API with ``d`` in pinned host memory and the
p x p result copied back to pinned host memory inside the timed region.  The matrix X is the
resident operator (built once, like the reference's cached CSR); ``d`` is the per-step input.

``--impl reference`` times the reference's own CPU kernels (oracle/_ref, driven by
oracle/ref_split.py) on a bounded row sample of the same workload on the host cores.
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
METRIC = "SplitMatrix sandwich GFLOP/s"
ROW_ORDER_DEFAULT = "sorted"
PASS_KERNEL = {
    "tensor": "k_dense_syrk_tc (tcgen05 SYRK + one-hot MMAs: dense self, dense x few-level cats)",
    "scatter": "k_dense_cross_fused / k_dense_cross_runs (dense x many-level cats + dense x sparse, vector RED)",
    "index": "index pass (k_pack_records, k_cat_pairs, k_cat_sparse_csc, k_sparse_sandwich)",
}
WORKLOAD = ("SplitMatrix 128 dense + 3x1000 CSC @1e-3 + cat{10,50,200,1000,2000}, p=6388, "
            "f32, n=%d total rows")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=N_TOTAL, help="total rows (default 4e7)")
    ap.add_argument("--cpu-rows", type=int, default=1_000_000,
                    help="rows of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="print per-block times to stderr")
    ap.add_argument("--row-order", default=os.environ.get("TABMAT_B200_BENCH_ROW_ORDER", ROW_ORDER_DEFAULT),
                    choices=["original", "sorted"],
                    help="'sorted': the resident matrix is stored with its rows sorted by the "
                         "many-level categorical codes (tabmat_b200.RowSortedMatrix, built once "
                         "like the reference's cached CSR); d arrives in the caller's order and "
                         "is permuted inside the timed region")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------
# algorithmic work: FLOPs = sum_k r_k (r_k + 1), r_k = non-zeros of row k
# bytes = every block array + d read once + the p x p float64 result written once
# ---------------------------------------------------------------------------------------
def split_flops(sparse_row_counts_sum, sparse_row_counts_sq_sum, n):
    base = P_DENSE + len(CAT_LEVELS)
    # sum (base + s)(base + s + 1) = n*base*(base+1) + (2*base+1)*sum s + sum s^2
    return n * base * (base + 1) + (2 * base + 1) * sparse_row_counts_sum + sparse_row_counts_sq_sum


def split_bytes(n, nnz, fsize=4):
    return (n * (P_DENSE * fsize + len(CAT_LEVELS) * 4 + fsize) + nnz * (fsize + 4) + 4 * (n + 1)
            + P_TOTAL * P_TOTAL * 8)


# ---------------------------------------------------------------------------------------
# reference arm / cpu baseline (host cores)
# ---------------------------------------------------------------------------------------
def host_blocks(n, seed):
    import scipy.sparse as sps

    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, P_DENSE), dtype=np.float32)
    blocks = [("dense", X)]
    mats = []
    for _ in range(SPARSE_BLOCKS):
        nnz = int(n * SPARSE_COLS * SPARSE_DENSITY)
        r = rng.integers(0, n, size=nnz)
        c = rng.integers(0, SPARSE_COLS, size=nnz)
        v = rng.standard_normal(nnz, dtype=np.float32)
        mats.append(sps.csc_matrix((v, (r, c)), shape=(n, SPARSE_COLS)))
    # the reference's SplitMatrix constructor merges all sparse blocks into one
    A = sps.hstack(mats, format="csc")
    blocks.append(("sparse", A))
    for K in CAT_LEVELS:
        blocks.append(("cat", rng.integers(0, K, size=n).astype(np.int32), K))
    d = rng.random(n, dtype=np.float32)
    counts = np.diff(A.tocsr().indptr).astype(np.float64)
    return blocks, d, split_flops(counts.sum(), (counts ** 2).sum(), n)


def cpu_reference_run(n_rows, steps, warmup):
    """Time the reference's CPU kernels on an n_rows sample.  Returns a dict."""
    from oracle import ref_loader
    from oracle.ref_split import RefSplit

    kind = "reference"
    try:
        ext = ref_loader.load_ext()
    except Exception as e:  # oracle/_ref missing or not loadable -> the C port
        ext = None
        kind = "port"
        why = repr(e)
    cores = len(os.sched_getaffinity(0))
    if ext is None:
        return _cpu_port_run(min(n_rows, 100_000), steps, warmup, why)
    blocks, d, flops = host_blocks(n_rows, seed=4)
    S = RefSplit(blocks, ext)
    for _ in range(max(1, warmup)):
        S.sandwich(d)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        S.sandwich(d)
        times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    return dict(
        value=flops / t / 1e9, unit="GFLOP/s", cores=cores, kind=kind,
        sample=(f"{n_rows} rows of the same SplitMatrix workload (f32), mean of {steps} after "
                f"{max(1, warmup)} warm-up; reference Cython/C++ kernels built from source "
                "(stand-in xsimd/jemalloc layer), OpenMP on all host cores; block loop of "
                "split_matrix.py:324-356 restated in oracle/ref_split.py"),
        ms_per_step=t * 1e3, ms_min=float(np.min(times)) * 1e3, rows=n_rows,
        ms_extrapolated_full=t * 1e3 * (N_TOTAL / n_rows),
    )


def _cpu_port_run(n_rows, steps, warmup, why):
    """Fallback: the single-threaded C restatement (oracle/tabmat_oracle.c) block by block."""
    from oracle import c_oracle as orc

    blocks, d, flops = host_blocks(n_rows, seed=4)
    X, A = blocks[0][1], blocks[1][1]
    cats = blocks[2:]

    def once():
        orc.dense_sandwich(X, d)
        orc.sparse_sandwich(A, d)
        orc.csr_dense_sandwich(A, X, d)
        for _, codes, K in cats:
            orc.cat_sandwich(codes, d, None, K)
            orc.cat_dense_sandwich(codes, K, d, X)
            orc.cat_sparse_sandwich(codes, K, d, A)
        for i in range(len(cats)):
            for j in range(i + 1, len(cats)):
                orc.cat_cat_sandwich(cats[i][1], cats[j][1], cats[i][2], cats[j][2], d)

    once()
    times = []
    for _ in range(max(1, min(steps, 3))):
        t0 = time.perf_counter()
        once()
        times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    return dict(value=flops / t / 1e9, unit="GFLOP/s", cores=1, kind="port",
                sample=f"{n_rows} rows, single-threaded C restatement (oracle/_ref unusable: {why})",
                ms_per_step=t * 1e3, rows=n_rows,
                ms_extrapolated_full=t * 1e3 * (N_TOTAL / n_rows))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.cpu_rows, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "GFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % args.n, "sample_rows": r["rows"],
                   "ms_extrapolated_to_full_n": r["ms_extrapolated_full"]},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
def device_split_matrix(n, seed, device):
    """The benchmark SplitMatrix for an n-row shard, generated directly in HBM."""
    import torch

    import tabmat_b200 as tm

    g = torch.Generator(device=device).manual_seed(seed)
    X = torch.randn((n, P_DENSE), device=device, dtype=torch.float32, generator=g)
    mats = [tm.DenseMatrix(X)]
    s_sum = 0.0
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
    return Xs, d, split_flops(s_sum, s_sq, n), nnz_total


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


def block_bytes(label, n, nnz, fsize=4):
    """Algorithmic bytes of one block computation."""
    def K_of(s):
        return int(s[3:])
    a, _, b = label.partition("x")
    if label == "dense.cross_fused":
        ps = SPARSE_BLOCKS * SPARSE_COLS
        return (n * (P_DENSE * fsize + len(CAT_LEVELS) * 4 + fsize) + nnz * (fsize + 4)
                + 4 * (n + 1) + (ps + sum(CAT_LEVELS)) * P_DENSE * fsize)
    if label == "dense.self":
        return n * P_DENSE * fsize + n * fsize + P_DENSE * P_DENSE * fsize
    if label == "sparse.self":
        ps = SPARSE_BLOCKS * SPARSE_COLS
        return nnz * (fsize + 4) + 4 * (n + 1) + n * fsize + ps * ps * fsize
    if label.endswith(".self"):
        return n * (4 + fsize) + K_of(a[:-5]) * fsize
    if a == "dense" and b == "sparse":
        ps = SPARSE_BLOCKS * SPARSE_COLS
        return n * P_DENSE * fsize + nnz * (fsize + 4) + 4 * (n + 1) + n * fsize + ps * P_DENSE * fsize
    if a == "dense":
        return n * (P_DENSE * fsize + 4 + fsize) + K_of(b) * P_DENSE * fsize
    if a == "sparse":
        ps = SPARSE_BLOCKS * SPARSE_COLS
        return n * (4 + fsize) + nnz * (fsize + 4) + 4 * (n + 1) + K_of(b) * ps * fsize
    return n * (8 + fsize) + K_of(a) * K_of(b) * fsize


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
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

    lo, hi = shard_bounds(args.n, world, rank)
    n_local = hi - lo
    Xs, d, flops_local, nnz_local = device_split_matrix(n_local, seed=1000 + rank, device=device)
    if args.row_order == "sorted":
        Xo = Xs
        Xs = tm.RowSortedMatrix.from_split(Xo)
        del Xo
        torch.cuda.empty_cache()
    S = RowShardedMatrix(Xs, args.n, pack=True, reduce_dtype=torch.float32)
    p = Xs.shape[1]
    assert p == P_TOTAL

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    step = lambda: S.sandwich(d)  # noqa: E731
    for _ in range(max(3, args.warmup)):
        step()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    tm.reset_launch_count()
    tm._lib.lib.tm_split_profile_enable(1)   # CUDA events around the three passes of every step
    total_ms = timed_loop(step, args.steps)
    launches = tm.launch_count()
    pass_ms = (ctypes.c_float * 3)()
    tm._lib.lib.tm_split_profile_read(pass_ms)
    tm._lib.lib.tm_split_profile_enable(0)
    pass_ms = {"tensor": float(pass_ms[0]), "scatter": float(pass_ms[1]), "index": float(pass_ms[2])}
    clocks = sampler.stop() if sampler else None

    # end to end through the public API: d from pinned host memory, result to pinned host
    d_host = torch.empty(n_local, dtype=torch.float32).pin_memory()
    d_host.copy_(d)
    out_host = torch.empty((p, p), dtype=torch.float64).pin_memory()

    # The p x p result is one object per job: with N > 1 ranks it is reduced to rank 0 and read
    # to the host there (dst=0), instead of 8 redundant 326 MB device->host copies.
    e2e_dst = 0 if world > 1 else None

    def e2e_step():
        # host buffers in, host buffer out: H2D of d, every kernel, D2H of the result (at one
        # rank the copy of the finished blocks overlaps the dense-operand passes)
        S.sandwich_into(d_host, out_host, dst=e2e_dst)

    for _ in range(2):
        e2e_step()
    e2e_ms = timed_loop(e2e_step, args.steps)

    fl = torch.tensor([float(flops_local), float(nnz_local)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(fl)
    flops, nnz = float(fl[0].item()), float(fl[1].item())

    bd = None
    if rank == 0 and args.breakdown:
        if args.row_order == "sorted":
            bd = block_breakdown(Xs.mat, Xs._gather(d))
        else:
            bd = block_breakdown(Xs, d)
    if world > 1:
        dist.barrier()

    if rank == 0:
        ms_step = total_ms / args.steps
        e2e_step_ms = e2e_ms / args.steps
        # dominant kernel = the longest of the three passes of tm_split_sandwich_blocks, timed
        # with CUDA events on its own stream inside the timed region (the tensor-core pass runs
        # concurrently on a side stream)
        n_oh = sum(1 for K in CAT_LEVELS if K <= 256)
        ps = SPARSE_BLOCKS * SPARSE_COLS
        pass_bytes = {
            # X + d + the few-level categoricals' codes in; dense self + their cross blocks out
            "tensor": n_local * (P_DENSE * 4 + 4 + 4 * n_oh)
            + 4 * (P_DENSE * P_DENSE + sum(K for K in CAT_LEVELS if K <= 256) * P_DENSE),
            # X + d + many-level codes + CSR (data, indices, indptr) in; their cross blocks out
            "scatter": n_local * (P_DENSE * 4 + 4 + 4 * (len(CAT_LEVELS) - n_oh)) + nnz_local * 8
            + 4 * (n_local + 1) + 4 * (ps + sum(K for K in CAT_LEVELS if K > 256)) * P_DENSE,
            # CSR + row ids + every code vector + d in; sparse self, cat x sparse, cat x cat out
            "index": nnz_local * 12 + 4 * (n_local + 1) + n_local * (4 * len(CAT_LEVELS) + 4)
            + 4 * (ps * ps + sum(CAT_LEVELS) * ps + sum(CAT_LEVELS)
                   + sum(a * b for i, a in enumerate(CAT_LEVELS) for b in CAT_LEVELS[i + 1:])),
        }
        top = max(pass_ms, key=pass_ms.get)
        top_bytes = pass_bytes[top]
        achieved = top_bytes / (pass_ms[top] * 1e-3) / 1e9
        # L2 RED payload of the scatter pass: one 512-byte row per many-level categorical and
        # per sparse non-zero; the measured L2 atomic peak is 6.0 TB/s (tools/micro/red_bench.cu)
        # (row-sorted storage: the many-level categorical runs collapse to one RED per run, so
        # only the sparse non-zeros are left)
        n_red_cat = 0 if args.row_order == "sorted" else len(CAT_LEVELS) - n_oh
        red_bytes = (n_local * n_red_cat + nnz_local) * P_DENSE * 4
        whole_bytes = split_bytes(n_local, nnz_local)
        try:
            traffic = json.loads((ROOT / "profiles" / "traffic.json").read_text()).get(top)
        except Exception:
            traffic = None
        line = {
            "metric": METRIC, "value": flops / (ms_step * 1e-3) / 1e9, "unit": "GFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD % args.n, "rows_per_gpu": n_local, "p": p,
                       "nnz_sparse_total": nnz, "l2": "inputs (>20 GB per step) exceed the 126 MB L2",
                       "parallelism": f"row-shard x{world} + NCCL allreduce of the flat block workspace (f32, 90 MB)",
                       "row_order": (args.row_order if args.row_order == "original" else
                                     "sorted by (cat2000, cat1000) at construction; d permuted "
                                     "inside the timed region"),
                       "pass_schedule": "serial (tensor, index, scatter): overlapping them on side "
                                        "streams measured within 1 ms of the serial sum",
                       "algorithmic_flop_per_step": flops,
                       "whole_step_hbm_gbs": whole_bytes / (ms_step * 1e-3) / 1e9,
                       "whole_step_hbm_frac": whole_bytes / (ms_step * 1e-3) / 1e9 / hbm_peak},
            "e2e": {"value": flops / (e2e_step_ms * 1e-3) / 1e9, "unit": "GFLOP/s",
                    "ms_per_step": e2e_step_ms, "h2d_bytes_per_step": int(n_local * 4),
                    "d2h_bytes_per_step": int(p * p * 8),
                    "note": "d (this rank's shard) from pinned host memory every step; the p x p "
                            "float64 result read to pinned host memory"
                            + (" on rank 0 (reduce to rank 0)" if world > 1 else "")},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": PASS_KERNEL[top], "achieved": achieved,
                         "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "peak_source": peak_src, "launch_ms": pass_ms[top],
                         "algorithmic_bytes": top_bytes,
                         "l2_red": {"payload_bytes": red_bytes,
                                    "achieved_TBs": red_bytes / (pass_ms["scatter"] * 1e-3) / 1e12
                                    if pass_ms["scatter"] > 0 else None,
                                    "measured_peak_TBs": 6.0,
                                    "note": "the scatter pass is bound by the L2 atomic units, "
                                            "not by HBM"}},
            "passes_ms": pass_ms,
            "passes_hbm_frac": {k: (pass_bytes[k] / (v * 1e-3) / 1e9 / hbm_peak if v > 0 else None)
                                for k, v in pass_ms.items()},
        }
        if bd is not None:
            line["breakdown_ms"] = bd
        if world == 1 and not args.no_cpu_baseline:
            try:
                r = cpu_reference_run(args.cpu_rows, 3, 1)
                line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
                line["cpu_baseline"]["ms_extrapolated_to_full_n"] = r["ms_extrapolated_full"]
            except Exception as e:  # pragma: no cover
                line["cpu_baseline"] = {"value": None, "unit": "GFLOP/s", "cores": 0,
                                        "kind": "port", "sample": f"failed: {e!r}"}
        if args.breakdown and bd is not None:
            for k, v in sorted(bd.items(), key=lambda kv: -kv[1]):
                print(f"  {k:24s} {v:9.3f} ms", file=sys.stderr)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
