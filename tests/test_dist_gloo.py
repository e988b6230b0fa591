"""world_size=2 on CPU (gloo): the row-sharding + allreduce host logic of
``tabmat_b200.distributed`` — shard bounds, restriction rebasing, packed-triangle payload.
The local shard compute is a tiny numpy stand-in here (the CUDA classes need a GPU); the
same wrapper runs the CUDA classes in bench.py / the GPU tests."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _HostShard:
    """Dense host shard with the MatrixBase methods on CPU torch tensors."""

    def __init__(self, X):
        self.X = torch.from_numpy(X)
        self.shape = X.shape
        self.dtype = X.dtype

    def _sel(self, rows, cols):
        M = self.X
        if rows is not None:
            M = M[torch.from_numpy(np.asarray(rows, dtype=np.int64))]
        if cols is not None:
            M = M[:, torch.from_numpy(np.asarray(cols, dtype=np.int64))]
        return M

    def sandwich(self, d, rows=None, cols=None):
        M = self._sel(rows, cols)
        dd = d if rows is None else d[torch.from_numpy(np.asarray(rows, dtype=np.int64))]
        return M.T @ (dd[:, None] * M)

    def transpose_matvec(self, v, rows=None, cols=None):
        M = self._sel(rows, cols)
        vv = v if rows is None else v[torch.from_numpy(np.asarray(rows, dtype=np.int64))]
        return M.T @ vv

    def matvec(self, v, cols=None):
        M = self._sel(None, cols)
        return M @ (v if cols is None else v[torch.from_numpy(np.asarray(cols, dtype=np.int64))])

    def _get_col_stds(self, weights, col_means):
        return torch.sqrt(((self.X - col_means) ** 2 * weights[:, None]).sum(0))


class SparseMatrix(_HostShard):
    """Stand-in with the NON-additive std formula of the sparse / categorical blocks
    (sqrt(max(0, sum w x^2 - mean^2)) with the global mean, sparse_matrix.py:305-315); the
    class name is what RowShardedMatrix._dense_column_mask looks at."""

    def _get_col_stds(self, weights, col_means):
        return torch.sqrt(torch.clamp_min((self.X ** 2 * weights[:, None]).sum(0) - col_means ** 2, 0))


def _worker(rank, world, port, n, p, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tabmat_b200.distributed import RowShardedMatrix, shard_bounds

        rng = np.random.default_rng(0)  # same global data on every rank
        X = rng.standard_normal((n, p))
        d = rng.standard_normal(n)
        v = rng.standard_normal(n)
        rows = np.sort(rng.choice(n, size=n // 2, replace=False))
        cols = np.array([0, 2, 3])
        lo, hi = shard_bounds(n, world, rank)
        res = {}
        for pack in (True, False):
            S = RowShardedMatrix(_HostShard(X[lo:hi]), n, pack=pack)
            dl, vl = torch.from_numpy(d[lo:hi]), torch.from_numpy(v[lo:hi])
            res[f"sand{pack}"] = S.sandwich(dl).numpy()
            res[f"sand_rc{pack}"] = S.sandwich(dl, rows, cols).numpy()
            res[f"tmv{pack}"] = S.transpose_matvec(vl, rows, cols).numpy()
            res[f"mv{pack}"] = S.matvec(torch.from_numpy(np.arange(p, dtype=np.float64))).numpy()
            only0 = S.sandwich(dl, dst=0)
            assert (only0 is None) == (rank != 0)
            if rank == 0:
                np.testing.assert_allclose(only0.numpy(), res[f"sand{pack}"], rtol=1e-10, atol=1e-10)
        full = _HostShard(X)
        dt, vt = torch.from_numpy(d), torch.from_numpy(v)
        for pack in (True, False):
            np.testing.assert_allclose(res[f"sand{pack}"], full.sandwich(dt).numpy(), rtol=1e-10,
                                       atol=1e-10)
            np.testing.assert_allclose(res[f"sand_rc{pack}"],
                                       full.sandwich(dt, rows, cols).numpy(), rtol=1e-10,
                                       atol=1e-10)
            np.testing.assert_allclose(res[f"tmv{pack}"],
                                       full.transpose_matvec(vt, rows, cols).numpy(), rtol=1e-10,
                                       atol=1e-10)
            np.testing.assert_allclose(
                res[f"mv{pack}"],
                full.matvec(torch.from_numpy(np.arange(p, dtype=np.float64))).numpy()[lo:hi])
        # row-sharded standardize: moments by allreduce, corrections after the collective
        w = np.abs(rng.standard_normal(n))
        w /= w.sum()
        Xoff = X + 3.0   # big means: a shard's second moment is far below mean^2
        Ssp = RowShardedMatrix(SparseMatrix(Xoff[lo:hi]), n)
        mu_off = torch.from_numpy(w @ Xoff)
        np.testing.assert_allclose(Ssp._get_col_stds(torch.from_numpy(w[lo:hi]), mu_off).numpy(),
                                   np.sqrt(w @ (Xoff - w @ Xoff) ** 2), rtol=1e-8)
        S = RowShardedMatrix(_HostShard(X[lo:hi]), n)
        for center, scale in ((True, True), (True, False), (False, True)):
            Z, means, stds = S.standardize(torch.from_numpy(w[lo:hi]), center, scale)
            mu = w @ X
            sd = np.sqrt(w @ (X - mu) ** 2)
            np.testing.assert_allclose(means.numpy(), mu if center else 0 * mu, atol=1e-12)
            if scale:
                np.testing.assert_allclose(stds.numpy(), sd, rtol=1e-10)
            mult = 1 / sd if scale else np.ones(p)
            Xs = (X - (mu if center else 0)) * mult
            np.testing.assert_allclose(
                Z.transpose_matvec(torch.from_numpy(v[lo:hi]), rows, cols).numpy(),
                Xs[rows][:, cols].T @ v[rows], rtol=1e-9, atol=1e-9)
            beta = np.arange(1, p + 1, dtype=np.float64)
            np.testing.assert_allclose(Z.matvec(torch.from_numpy(beta)).numpy(), (Xs @ beta)[lo:hi],
                                       rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(Z.matvec(torch.from_numpy(beta), cols).numpy(),
                                       (Xs[:, cols] @ beta[cols])[lo:hi], rtol=1e-9, atol=1e-9)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n", [101, 64])
def test_row_sharded_world2_gloo(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, 5, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    out = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(out) == [(0, "ok"), (1, "ok")], out


def test_shard_bounds_and_rows():
    from tabmat_b200.distributed import shard_bounds, shard_rows

    n = 10
    spans = [shard_bounds(n, 4, r) for r in range(4)]
    assert spans == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert shard_bounds(2, 4, 3) == (2, 2)
    rows = np.array([0, 2, 3, 7, 9])
    got = [shard_rows(rows, lo, hi).tolist() for lo, hi in spans]
    assert got == [[0, 2], [0], [1], [0]]
    assert shard_rows(None, 0, 3) is None
