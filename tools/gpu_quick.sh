#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tc"; timeout -s KILL 600 python -m pytest tests/test_gpu_dense_tc.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1; rc=$?; echo "rc=$rc"; tail -25 gpurun_out/pytest_tc.log
if [ $rc -ne 0 ]; then exit 0; fi
echo "== pytest gpu"; timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== bench full"; timeout -s KILL 1500 python bench.py --steps 5 --breakdown --no-cpu-baseline > gpurun_out/bench_full.log 2>&1; echo "rc=$?"; tail -32 gpurun_out/bench_full.log | cut -c1-700
if [ -n "$BLOCKS" ]; then echo "== bench blocks"; timeout -s KILL 900 python tools/bench_blocks.py $BLOCKS > gpurun_out/bench_blocks.log 2>&1; echo "rc=$?"; cut -c1-330 gpurun_out/bench_blocks.log; fi
