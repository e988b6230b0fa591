"""Bring-up diagnostics for the tcgen05 SYRK kernel: dumps stage-0 smem (after scaling) and the
raw TMEM accumulator tile of CTA 0 through the tm_debug_set_tc_buffer test hook."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tabmat_b200._lib import lib  # noqa: E402
from tabmat_b200.ext.dense import dense_sandwich  # noqa: E402

lib.tm_debug_set_tc_buffer.argtypes = [ctypes.c_void_p]
lib.tm_debug_set_tc_buffer.restype = None

n, p = int(sys.argv[1]) if len(sys.argv) > 1 else 32, int(sys.argv[2]) if len(sys.argv) > 2 else 32
X = (np.arange(n)[:, None] * p + np.arange(p)[None, :]).astype(np.float32) % 1024
d = np.ones(n, np.float32)
dbg = torch.zeros(65536 + 3 * 128 * 128, dtype=torch.float32, device="cuda")
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lib.tm_debug_set_tc_variant(variant)
print(f"==== n={n} p={p} variant={variant}")
lib.tm_debug_set_tc_buffer(dbg.data_ptr())
lib.tm_set_dense_f32_mode(2)
Xd, dd = torch.from_numpy(X).cuda(), torch.from_numpy(d).cuda()
out = dense_sandwich(Xd, dd, None, None)
torch.cuda.synchronize()
lib.tm_debug_set_tc_buffer(None)
D = dbg.cpu().numpy()
ref = X.astype(np.float64).T @ X.astype(np.float64)
np.set_printoptions(linewidth=200, suppress=True)
print("out[:4,:8]\n", out.cpu().numpy()[:4, :8])
print("ref[:4,:8]\n", ref[:4, :8])
mt = (p + 127) // 128
half = mt * 4 * 4096 // 4
R, B = D[:half], D[half:2 * half]
print("R smem row0 (32 floats):", R[:32])
print("R smem row1 (32 floats):", R[32:64])
print("R smem row9 (32 floats):", R[9 * 32:10 * 32])
print("B smem row1 (32 floats):", B[32:64])
print("R nonzeros:", np.count_nonzero(R), "B nonzeros:", np.count_nonzero(B))
T = D[65536:65536 + 128 * 128].reshape(128, 128)
print("TMEM tile0 nonzeros:", np.count_nonzero(T))
print("TMEM tile0 [:4,:8]\n", T[:4, :8])
if np.count_nonzero(T):
    nz = np.argwhere(T != 0)
    print("nonzero row range", nz[:, 0].min(), nz[:, 0].max(), "col range", nz[:, 1].min(), nz[:, 1].max())
    q = min(p, 128)
    print("max |T - ref| over p x p:", np.abs(T[:q, :q] - ref[:q, :q]).max(), "max|ref|", np.abs(ref).max())
    print("max |T.T - ref|:", np.abs(T[:q, :q].T - ref[:q, :q]).max())
