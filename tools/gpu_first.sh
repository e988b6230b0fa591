#!/bin/bash
# First GPU bring-up: CUDA-core paths first, then the tcgen05 kernel in its own process
# under a hard timeout (a hung kernel must not take the box down).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke (cuda-core dense)"; TABMAT_B200_DENSE_F32_MODE=1 timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_core.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/smoke_core.log
echo "== pytest gpu (cuda-core dense)"; TABMAT_B200_DENSE_F32_MODE=1 timeout -s KILL 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dense_tc.py -p no:cacheprovider > gpurun_out/pytest_core.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/pytest_core.log
echo "== pytest tcgen05"; timeout -s KILL 300 python -m pytest tests/test_gpu_dense_tc.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_tc.log
echo "== smoke (auto)"; timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_auto.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/smoke_auto.log
echo "== bench n=4e6 (cuda-core dense)"; TABMAT_B200_DENSE_F32_MODE=1 timeout -s KILL 900 python bench.py --n 4000000 --steps 3 --breakdown --cpu-rows 200000 > gpurun_out/bench_4e6_core.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/bench_4e6_core.log
