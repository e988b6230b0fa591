set -u
mkdir -p gpurun_out
echo "== dist tests (2 GPUs)"
timeout -s KILL 900 python -m pytest tests/test_gpu_dist_nccl.py tests/test_gpu_dense_tc.py -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_r2d_n2.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_r2d_n2.log | cut -c1-300
grep -h "ranks vs 1 GPU\|DIST_" gpurun_out/pytest_r2d_n2.log | head
echo "== bench n=1"
TAG=r2d bash tools/gpu.sh bench --steps 10
echo "== bench n=2"
TAG=r2d CUT=3500 bash tools/gpu.sh benchn 2 --steps 10
echo "== bench n=2 e2e via reduce to rank 0 (round-1 form)"
TABMAT_B200_E2E_SHARED=0 TAG=r2d_r1e2e bash tools/gpu.sh benchn 2 --steps 10 --no-cpu-baseline | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print(j['ms_per_step'], j['e2e']['ms_per_step'])"
echo "== reference arm n=2"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
