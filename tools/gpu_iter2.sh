#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_iter.sh
echo "== blocks c3"; timeout -s KILL 200 python tools/bench_blocks.py c3 --reps 20 > gpurun_out/bench_blocks_c3.log 2>&1; echo "rc=$?"; cut -c1-200 gpurun_out/bench_blocks_c3.log
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -q -p no:cacheprovider --tb=short -x 2>&1 | tail -5 | cut -c1-200
