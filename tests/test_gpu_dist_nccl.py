"""N-GPU parity on hardware: ``RowShardedMatrix`` over 2 NCCL ranks == the single-GPU result on
identical inputs (sandwich with and without a ``rows`` restriction that spans the shard
boundary, reduce-to-one-rank, ``sandwich_into``, transpose_matvec, matvec; both row orders).
Skipped on a box with fewer than 2 GPUs (run it with ``gpurun --gpus 2``).  Reference behaviour
matched: tests/test_split_matrix.py:229-288 (sandwich / matvec against the unsplit matrix)."""

import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_row_sharded_two_ranks_equals_single_gpu():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", str(ROOT / "tests" / "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    print(r.stdout[-3000:])
    assert r.returncode == 0 and "DIST_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
