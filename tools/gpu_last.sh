#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout -s KILL 600 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-200
echo "== row-blocked opt-in: scale + index tests"; TABMAT_B200_CSC_ROW_BLOCKS=1 timeout -s KILL 300 python -m pytest tests/test_gpu_scale.py tests/test_gpu_index_fused.py -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -3 | cut -c1-200
echo "== smoke"; timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n1.log | cut -c1-400
