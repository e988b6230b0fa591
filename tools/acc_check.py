"""Full-size accuracy of the tcgen05 dense sandwich against a float64 reference computed on the
GPU in row chunks (the CPU reference only sees 1e6-row samples): normwise error and the relative
error of the diagonal (sums of like-signed terms, where an accumulator bias shows).

    python tools/acc_check.py [n] [p]
"""

import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import tabmat_b200 as tm  # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 40_000_000
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    X = torch.randn((n, p), device=dev, dtype=torch.float32, generator=g)
    d = torch.rand(n, device=dev, dtype=torch.float32, generator=g)
    ref = torch.zeros((p, p), device=dev, dtype=torch.float64)
    step = 1_000_000
    for lo in range(0, n, step):
        Xc = X[lo:lo + step].double()
        ref += Xc.t() @ (Xc * d[lo:lo + step].double()[:, None])
        del Xc
    D = tm.DenseMatrix(X)
    lib = tm._lib.lib
    for mode, name in ((0, "tf32"), (3, "3xtf32"), (1, "fp32 cuda cores")):
        lib.tm_set_dense_f32_mode(mode)
        got = D.sandwich(d).double()
        err = (got - ref).abs()
        diag_rel = ((got.diagonal() - ref.diagonal()) / ref.diagonal())
        print(f"n={n} p={p} {name:16s} normwise {float(err.max() / ref.abs().max()):.3e}  "
              f"diag rel: mean {float(diag_rel.mean()):+.3e} max|.| {float(diag_rel.abs().max()):.3e}",
              flush=True)
    lib.tm_set_dense_f32_mode(0)


if __name__ == "__main__":
    main()
