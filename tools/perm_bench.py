"""d -> stored row order: gather form (out[i] = d[perm[i]]) against scatter form
(out[inv[j]] = d[j]) at n = 4e7; the permutation is that of a matrix sorted by two keys."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import tabmat_b200 as tm  # noqa: E402
from tabmat_b200 import _dev  # noqa: E402
from tabmat_b200._lib import check, fn  # noqa: E402

n = 40_000_000
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
key = torch.randint(0, 2000, (n,), device=dev, generator=g) * 1000 + torch.randint(0, 1000, (n,), device=dev, generator=g)
perm = torch.argsort(key, stable=True).to(torch.int32)
inv = torch.empty(n, dtype=torch.int32, device=dev)
inv[perm.to(torch.int64)] = torch.arange(n, dtype=torch.int32, device=dev)
d = torch.rand(n, device=dev)
out = torch.empty_like(d)
flush = torch.empty(64 << 20, device=dev)


def run(name, idx):
    ts = []
    for _ in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(fn(name, "f32")(_dev.ptr(d), _dev.ptr(idx), n, _dev.ptr(out), 0, _dev.stream_ptr()))
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


a = run("tm_permute_gather", perm)
ref = out.clone()
b = run("tm_permute_scatter", inv)
assert torch.equal(ref, out)
print(f"gather (d[perm[i]]) {a:.3f} ms   scatter (out[inv[j]] = d[j]) {b:.3f} ms")
