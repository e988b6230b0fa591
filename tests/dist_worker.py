"""Worker of tests/test_gpu_dist_nccl.py: one process per GPU (launched by
torch.distributed.run), NCCL.  Every rank builds the SAME seeded host arrays, uploads its
contiguous row shard, and the row-sharded results are compared with the single-GPU result of
the whole matrix computed on rank 0 ("run 1 vs N GPUs on identical inputs")."""

import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    import torch
    import torch.distributed as dist

    import bench
    import tabmat_b200 as tm
    from tabmat_b200.distributed import RowShardedMatrix, shard_bounds

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    n = int(os.environ.get("TM_DIST_ROWS", "200003"))
    wl = bench.C5(None)
    s = wl.host_sample(n, seed=31)
    lo, hi = shard_bounds(n, world, rank)
    rng = np.random.default_rng(7)
    rows = np.sort(rng.choice(n, size=n // 2, replace=False)).astype(np.int32)   # spans the shards
    v = rng.standard_normal(n).astype(np.float32)
    beta = rng.standard_normal(bench.P_TOTAL).astype(np.float32)
    fails = []

    def check(what, got, ref, tol=1e-3):
        err = float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))
        if not err <= tol:
            fails.append(f"{what}: {err:.3e}")
        return err

    full = wl.ours_from_sample(s, 0, n) if rank == 0 else None
    for order in ("sorted", "original"):
        Xl = wl.ours_from_sample(s, lo, hi)
        if order == "sorted":
            Xl = tm.RowSortedMatrix.from_split(Xl)
        S = RowShardedMatrix(Xl, n, pack=True, reduce_dtype=torch.float32)
        dl = torch.from_numpy(s["d"][lo:hi]).to(dev)
        got = S.sandwich(dl).cpu().numpy()
        got_r = S.sandwich(dl, rows=rows).cpu().numpy()
        S.overlap = False     # whole workspace reduced at the end (no allreduce/compute overlap)
        got_plain = S.sandwich(dl).cpu().numpy()
        S.overlap = True
        only0 = S.sandwich(dl, dst=0)
        assert (only0 is None) == (rank != 0)
        tmv = S.transpose_matvec(torch.from_numpy(v[lo:hi]).to(dev), rows=rows).cpu().numpy()
        mv = S.matvec(torch.from_numpy(beta).to(dev)).cpu().numpy()
        out_host = np.zeros((bench.P_TOTAL, bench.P_TOTAL))
        S.sandwich_into(s["d"][lo:hi], out_host, dst=0)
        torch.cuda.synchronize()
        if rank == 0:
            ref = full.sandwich(s["d"])
            e1 = check(f"{order}: sandwich", got, ref)
            check(f"{order}: sandwich, no overlap", got_plain, ref)
            check(f"{order}: sandwich rows", got_r, full.sandwich(s["d"], rows=rows))
            check(f"{order}: sandwich dst=0", only0.cpu().numpy(), ref)
            check(f"{order}: sandwich_into", out_host, ref)
            check(f"{order}: transpose_matvec", tmv, full.transpose_matvec(v, rows=rows))
            check(f"{order}: matvec", mv, full.matvec(beta)[lo:hi])
            print(f"{order}: {world} ranks vs 1 GPU, normwise sandwich error {e1:.3e}", flush=True)
        # result delivered band by band into a host buffer shared by the ranks
        from tabmat_b200.distributed import SharedHostResult

        shared = SharedHostResult.for_group(bench.P_TOTAL, tag=f"tm_test_{order}")
        S.sandwich_into_shared(s["d"][lo:hi], shared)
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            check(f"{order}: sandwich_into_shared", np.array(shared.array), ref)
        dist.barrier()
        shared.close(unlink=rank == 0)
        # fused IRLS pass and the row-sharded standardized sandwich against one GPU
        Hs, gs = S.sandwich_and_transpose_matvec(dl, torch.from_numpy(v[lo:hi]).to(dev), rows=rows)
        w = np.full(n, 1.0 / n, dtype=np.float32)
        Z, _, _ = S.standardize(torch.from_numpy(w[lo:hi]).to(dev), True, True)
        zs = Z.sandwich(dl, rows=rows).cpu().numpy()
        ztm = Z.transpose_matvec(torch.from_numpy(v[lo:hi]).to(dev)).cpu().numpy()
        if rank == 0:
            Hf, gf = full.sandwich_and_transpose_matvec(s["d"], v, rows)
            check(f"{order}: fused Hessian", Hs.cpu().numpy(), Hf)
            check(f"{order}: fused score", gs.cpu().numpy(), gf)
            Zf, _, _ = full.standardize(w, True, True)
            check(f"{order}: standardized sandwich", zs, Zf.sandwich(s["d"], rows), tol=5e-3)
            check(f"{order}: standardized transpose_matvec", ztm, Zf.transpose_matvec(v))
        dist.barrier()
    ok = torch.tensor([0 if fails else 1], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_OK" if int(ok.item()) == 1 else "DIST_FAIL " + "; ".join(fails), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(ok.item()) == 1 else 1)


if __name__ == "__main__":
    main()
