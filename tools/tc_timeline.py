"""Pipeline timeline of the tcgen05 SYRK kernel (CTA 0): cycle stamps per iteration for
0 producer got emptyR | 1 scale got emptyB | 2 scale got full | 3 scale arrived | 4 MMA got scaled | 5 MMA issued+commit"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tabmat_b200._lib import lib  # noqa: E402
from tabmat_b200.ext.dense import dense_sandwich  # noqa: E402

lib.tm_debug_set_tc_buffer.argtypes = [ctypes.c_void_p]
n, p = int(sys.argv[1]), int(sys.argv[2])
TL_OFF = 65536 + 3 * 128 * 128
dbg = torch.zeros(TL_OFF + 2 * 8 * 1024 + 64, dtype=torch.float32, device="cuda")
X = torch.randn((n, p), device="cuda")
d = torch.rand(n, device="cuda")
lib.tm_set_dense_f32_mode(2)
for _ in range(2):
    dense_sandwich(X, d, None, None)
lib.tm_debug_set_tc_buffer(dbg.data_ptr())
dense_sandwich(X, d, None, None)
torch.cuda.synchronize()
lib.tm_debug_set_tc_buffer(None)
T = dbg[TL_OFF:TL_OFF + 2 * 8 * 1024].cpu().numpy().view(np.int64).reshape(1024, 8)
its = min(1024, (n // 32 + 147) // 148)
T = T[:its]
t0 = T[0, 0]
T = T - t0
np.set_printoptions(linewidth=200)
print(f"n={n} p={p} iterations of CTA0 recorded: {its}")
print("first 4 iterations (cycles since start):\n", T[:4])
mid = T[its // 2: its // 2 + 8]
print("mid iterations:\n", mid)
per = np.diff(T[its // 4: 3 * its // 4], axis=0)
print("steady-state period per role (cycles, median):", np.median(per, axis=0))
seg = T[its // 4: 3 * its // 4]
print("median  math+stores(6-2) wait::st(7-6) fence.proxy+arrive(3-7):",
      np.median(seg[:, 6] - seg[:, 2]), np.median(seg[:, 7] - seg[:, 6]), np.median(seg[:, 3] - seg[:, 7]))
print("median  full-wait(2-1) scale-work(3-2) scaled->mma(4-3) mma-issue(5-4):",
      np.median(seg[:, 2] - seg[:, 1]), np.median(seg[:, 3] - seg[:, 2]),
      np.median(seg[:, 4] - seg[:, 3]), np.median(seg[:, 5] - seg[:, 4]))
# how far ahead is the producer: iteration index of producer when scale starts iteration i
prod = T[:, 0]
ahead = [np.searchsorted(prod, seg[i, 2]) - (its // 4 + i) for i in range(len(seg))]
print("producer lead over scale (iterations), median:", np.median(ahead))
