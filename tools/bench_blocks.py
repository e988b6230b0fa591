"""Per-kernel timings at the BASELINE.json parity/bench shapes (C2 dense, C3 categorical,
C4 sparse) with CUDA events; prints one JSON line per case with achieved GB/s / TFLOP/s.

    python tools/bench_blocks.py [c2] [c3] [c4] [--reps 10]
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import tabmat_b200 as tm  # noqa: E402
from tabmat_b200.ext import categorical as ecat  # noqa: E402
from tabmat_b200.ext import dense as edense  # noqa: E402

PEAKS = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
HBM = float(PEAKS.get("hbm_gbs", 6650.0))
BF16 = float(PEAKS.get("bf16_tflops", 1590.0))
reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 10
which = [a for a in sys.argv[1:] if a in ("c2", "c3", "c4", "c2f64", "c5mv")] or ["c2", "c3", "c4"]
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")


def timeit(fn, flush_l2=False):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        if flush_l2:
            flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def emit(name, ms, ms_min, nbytes, flops, extra=None):
    gbs = nbytes / (ms * 1e-3) / 1e9
    line = {"case": name, "ms_median": ms, "ms_min": ms_min, "algorithmic_bytes": nbytes,
            "GB/s": gbs, "hbm_frac": gbs / HBM, "algorithmic_flop": flops,
            "TFLOP/s": flops / (ms * 1e-3) / 1e12}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


g = torch.Generator(device="cuda").manual_seed(1)
if "c2" in which:
    n, p = 10_000_000, 256
    X = torch.randn((n, p), device="cuda", dtype=torch.float32, generator=g)
    d = torch.rand(n, device="cuda", dtype=torch.float32, generator=g)
    for mode, nm in ((0, "tcgen05"), (1, "cuda-core")):
        tm._lib.lib.tm_set_dense_f32_mode(mode)
        ms, mn = timeit(lambda: edense.dense_sandwich(X, d, None, None))
        fl_full = 2.0 * n * p * p
        emit(f"C2 dense f32 n=1e7 p=256 [{nm}]", ms, mn, n * p * 4 + n * 4 + p * p * 4,
             n * p * (p + 1), {"executed_TFLOP/s_full_square": fl_full / (ms * 1e-3) / 1e12,
                               "tf32_peak_assumed_TF": BF16 / 2,
                               "tensor_frac_executed(3of4 tiles)": 0.75 * fl_full / (ms * 1e-3) / 1e12 / (BF16 / 2)})
    tm._lib.lib.tm_set_dense_f32_mode(0)
    del X, d
    n, p = 10_000_000, 128
    X = torch.randn((n, p), device="cuda", dtype=torch.float32, generator=g)
    d = torch.rand(n, device="cuda", dtype=torch.float32, generator=g)
    ms, mn = timeit(lambda: edense.dense_sandwich(X, d, None, None))
    emit("dense f32 n=1e7 p=128 [tcgen05]", ms, mn, n * p * 4 + n * 4 + p * p * 4, n * p * (p + 1))
    del X, d
if "c2f64" in which:
    n, p = 2_000_000, 256
    X = torch.randn((n, p), device="cuda", dtype=torch.float64, generator=g)
    d = torch.rand(n, device="cuda", dtype=torch.float64, generator=g)
    ms, mn = timeit(lambda: edense.dense_sandwich(X, d, None, None))
    emit("dense f64 n=2e6 p=256 [cuda-core]", ms, mn, n * p * 8 + n * 8 + p * p * 8, n * p * (p + 1))
    del X, d
if "c3" in which:
    n, K = 10_000_000, 2000
    codes = torch.randint(0, K, (n,), device="cuda", dtype=torch.int32, generator=g)
    d = torch.rand(n, device="cuda", dtype=torch.float32, generator=g)
    ms, mn = timeit(lambda: ecat.sandwich_categorical(codes, d, None, K, False), flush_l2=True)
    emit("C3 cat sandwich f32 n=1e7 K=2000 (L2 flushed)", ms, mn, n * 8 + K * 4, n)
    out = torch.zeros(n, device="cuda")
    v = torch.rand(K, device="cuda", generator=g)
    ms, mn = timeit(lambda: ecat.matvec(codes, v, n, None, K, out, False), flush_l2=True)
    emit("C3 cat matvec f32 (L2 flushed)", ms, mn, n * 12 + K * 4, n)
    del codes, d
if "c4" in which:
    n, p, nnz_t = 10_000_000, 5000, 50_000_000
    r = torch.randint(0, n, (nnz_t,), device="cuda", generator=g, dtype=torch.int64)
    c = torch.randint(0, p, (nnz_t,), device="cuda", generator=g, dtype=torch.int64)
    key = torch.unique(r * p + c)
    del r, c
    rows = torch.div(key, p, rounding_mode="floor")
    cols = (key - rows * p).to(torch.int32)
    cnt = torch.bincount(rows, minlength=n)
    indptr = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    indptr[1:] = torch.cumsum(cnt, 0)
    nnz = int(cols.numel())
    vals = torch.randn(nnz, device="cuda", dtype=torch.float64, generator=g)
    A = tm.SparseMatrix.from_device_csr(vals, cols, indptr.to(torch.int32), (n, p))
    del key, rows, cnt
    d = torch.rand(n, device="cuda", dtype=torch.float64, generator=g)
    ms, mn = timeit(lambda: A.sandwich(d))
    cf = (indptr[1:] - indptr[:-1]).double()
    flops = float((cf * (cf + 1)).sum().item())
    emit("C4 sparse self f64 n=1e7 p=5000", ms, mn, nnz * 12 + 4 * (n + 1) + n * 8 + p * p * 8, flops)
    B = torch.randn((n, 128), device="cuda", dtype=torch.float64, generator=g)
    Bm = tm.DenseMatrix(B)
    ms, mn = timeit(lambda: A._cross_sandwich(Bm, d, None, None, None))
    emit("C4 sparse x dense(128) f64", ms, mn, n * 128 * 8 + nnz * 12 + 4 * (n + 1) + n * 8 + p * 128 * 8,
         2.0 * nnz * 128)
    v = torch.randn(p, device="cuda", dtype=torch.float64, generator=g)
    ms, mn = timeit(lambda: A.matvec(v))
    emit("C4 sparse matvec f64", ms, mn, nnz * 12 + 4 * (n + 1) + n * 8 + p * 8, 2.0 * nnz)
    w = torch.randn(n, device="cuda", dtype=torch.float64, generator=g)
    ms, mn = timeit(lambda: A.transpose_matvec(w))
    emit("C4 sparse transpose_matvec f64", ms, mn, nnz * 12 + 4 * (p + 1) + n * 8 + p * 8, 2.0 * nnz)

if "c5mv" in which:
    sys.path.insert(0, str(ROOT))
    import bench as B

    n = 40_000_000
    Xs, d, flops, nnz = B.device_split_matrix(n, seed=1000, device=torch.device("cuda", 0))
    p = Xs.shape[1]
    v = torch.randn(p, device="cuda", dtype=torch.float32, generator=g)
    nbytes = n * (B.P_DENSE * 4 + len(B.CAT_LEVELS) * 4 + 4) + nnz * 8 + 4 * (n + 1)
    ms, mn = timeit(lambda: Xs.matvec(v))
    emit("C5 split matvec f32 n=4e7 p=6388", ms, mn, nbytes, 2.0 * (n * B.P_DENSE + nnz) + n * 5)
    ms, mn = timeit(lambda: Xs.transpose_matvec(d))
    emit("C5 split transpose_matvec f32 n=4e7", ms, mn, nbytes, 2.0 * (n * B.P_DENSE + nnz) + n * 5)
