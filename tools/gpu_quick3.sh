#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tc"; timeout -s KILL 600 python -m pytest tests/test_gpu_dense_tc.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1; rc=$?; echo "rc=$rc"; tail -25 gpurun_out/pytest_tc.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout -s KILL 300 python tools/tc_timeline.py 4000000 128 2>&1 | tail -6
timeout -s KILL 300 python tools/tc_timeline.py 4000000 256 2>&1 | tail -6
echo "== bench full"; timeout -s KILL 1500 python bench.py --steps 5 --breakdown --no-cpu-baseline > gpurun_out/bench_full.log 2>&1; echo "rc=$?"; head -8 gpurun_out/bench_full.log
echo "== bench blocks"; timeout -s KILL 900 python tools/bench_blocks.py c2 > gpurun_out/bench_blocks.log 2>&1; echo "rc=$?"; cut -c1-250 gpurun_out/bench_blocks.log
