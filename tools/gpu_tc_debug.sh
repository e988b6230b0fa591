#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/tc_debug.log
for args in "32 32 0" "32 32 1" "64 128 0"; do
  timeout -s KILL 120 python tools/tc_debug.py $args >> gpurun_out/tc_debug.log 2>&1
done
grep -E "====|nonzeros|max \||range" gpurun_out/tc_debug.log
echo "== pytest tcgen05"; timeout -s KILL 600 python -m pytest tests/test_gpu_dense_tc.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_tc.log
