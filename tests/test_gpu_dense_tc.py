"""The tcgen05 / TMEM / TMA weighted-SYRK kernel (dense_tc.cu) against the oracle and against
the CUDA-core kernel, forced through ``tm_set_dense_f32_mode``."""

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


@pytest.fixture
def force_mode():
    from tabmat_b200._lib import lib

    def set_mode(m):
        lib.tm_set_dense_f32_mode(m)

    yield set_mode
    lib.tm_set_dense_f32_mode(0)


@pytest.mark.parametrize("n,p", [(1, 8), (31, 32), (32, 64), (33, 128), (1000, 132), (4097, 256),
                                 (300_001, 256), (70_000, 96), (5000, 200), (257, 12)])
def test_tcgen05_syrk_matches_oracle(n, p, force_mode):
    from tabmat_b200._lib import lib
    from tests.gpu_runner import run_cuda

    if not lib.tm_has_tcgen05():
        pytest.fail("tcgen05 path unavailable on this device (expected sm_100)")
    rng = np.random.default_rng(n + p)
    X = rng.standard_normal((n, p)).astype(np.float32)
    d = rng.standard_normal(n).astype(np.float32)
    args = dict(X=X, d=d, rows=None, cols=None)
    force_mode(2)
    got = run_cuda("dense_sandwich", args)
    force_mode(1)
    core = run_cuda("dense_sandwich", args)
    Xd = X.astype(np.float64)
    ref = Xd.T @ (d.astype(np.float64)[:, None] * Xd)
    cases.assert_close(got, ref, np.float32, f"tcgen05 n={n} p={p}")
    cases.assert_close(core, ref, np.float32, f"cuda-core n={n} p={p}")
    assert np.array_equal(got, got.T), "output must be exactly symmetric"
    if n <= 5000:
        cases.assert_close(got, cases.run_oracle("dense_sandwich", args), np.float32, "vs oracle")


def test_tcgen05_syrk_row_restriction(force_mode):
    from tests.gpu_runner import run_cuda

    rng = np.random.default_rng(3)
    n, p = 10_000, 64
    X = rng.standard_normal((n, p)).astype(np.float32)
    d = rng.random(n).astype(np.float32)
    rows = np.sort(rng.choice(n, size=n // 3, replace=False)).astype(np.int32)
    force_mode(2)
    got = run_cuda("dense_sandwich", dict(X=X, d=d, rows=rows, cols=None))
    Xd = X[rows].astype(np.float64)
    cases.assert_close(got, Xd.T @ (d[rows].astype(np.float64)[:, None] * Xd), np.float32, "rows")


def test_tcgen05_linearity_at_scale(force_mode):
    """Full-size property check (no oracle at this size): sandwich is linear in d and equals
    the column sums of d*X^2 on the diagonal."""
    import torch

    from tabmat_b200.ext.dense import dense_sandwich

    n, p = 2_000_000, 256
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn((n, p), device="cuda", dtype=torch.float32, generator=g)
    d1 = torch.rand(n, device="cuda", generator=g)
    d2 = torch.rand(n, device="cuda", generator=g)
    force_mode(2)
    s1, s2, s12 = dense_sandwich(X, d1, None, None), dense_sandwich(X, d2, None, None), \
        dense_sandwich(X, d1 + d2, None, None)
    scale = s12.abs().max().item()
    assert (s1 + s2 - s12).abs().max().item() / scale < 1e-3
    diag = (d1[:, None].double() * X.double() ** 2).sum(0)
    assert ((s1.diagonal().double() - diag).abs().max() / diag.abs().max()).item() < 1e-3
