"""f64 timings: dense sandwich (CUDA-core / DMMA path) and the C5-layout SplitMatrix at n=4e6."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import tabmat_b200 as tm  # noqa: E402
from tabmat_b200.ext import dense as edense  # noqa: E402


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


g = torch.Generator(device="cuda").manual_seed(1)
for n, p in ((2_000_000, 256), (4_000_000, 128), (10_000, 64)):
    X = torch.randn((n, p), device="cuda", dtype=torch.float64, generator=g)
    d = torch.rand(n, device="cuda", dtype=torch.float64, generator=g)
    ms = timeit(lambda: edense.dense_sandwich(X, d, None, None))
    print(json.dumps({"case": f"dense f64 n={n} p={p}", "ms": ms,
                      "GB/s": n * p * 8 / ms / 1e6, "TFLOP/s(n p^2 x2)": 2.0 * n * p * p / ms / 1e9}))
    del X, d
# C5 layout in f64 at n = 4e6
n = 4_000_000
X = torch.randn((n, 128), device="cuda", dtype=torch.float64, generator=g)
mats = [tm.DenseMatrix(X)]
for _ in range(3):
    nnz = int(n * 1000 * 1e-3)
    key = torch.unique(torch.randint(0, n, (nnz,), device="cuda", generator=g) * 1000
                       + torch.randint(0, 1000, (nnz,), device="cuda", generator=g))
    rows = torch.div(key, 1000, rounding_mode="floor")
    cols = (key - rows * 1000).to(torch.int32)
    indptr = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    indptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
    vals = torch.randn(cols.numel(), device="cuda", dtype=torch.float64, generator=g)
    mats.append(tm.SparseMatrix.from_device_csr(vals, cols, indptr.to(torch.int32), (n, 1000)))
for K in (10, 50, 200, 1000, 2000):
    codes = torch.randint(0, K, (n,), device="cuda", generator=g, dtype=torch.int32)
    mats.append(tm.CategoricalMatrix(codes, categories=np.arange(K), dtype=np.float64))
S = tm.SplitMatrix(mats)
d = torch.rand(n, device="cuda", dtype=torch.float64, generator=g)
tm._lib.lib.tm_split_profile_enable(1)
ms = timeit(lambda: S.sandwich(d))
import ctypes
pm = (ctypes.c_float * 3)()
tm._lib.lib.tm_split_profile_read(pm)
print(json.dumps({"case": "C5 layout f64 n=4e6 SplitMatrix.sandwich", "ms": ms,
                  "passes_ms": {"tensor": pm[0], "scatter": pm[1], "index": pm[2]}}))
