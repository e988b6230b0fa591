#!/bin/bash
# full GPU pass: tests, bench (full n), optional per-kernel benches
mkdir -p gpurun_out
echo "== pytest gpu"; timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench full"; timeout -s KILL 1500 python bench.py --steps 5 --breakdown > gpurun_out/bench_full.log 2>&1; echo "rc=$?"; tail -32 gpurun_out/bench_full.log | cut -c1-1200
if [ -n "$BLOCKS" ]; then echo "== bench blocks"; timeout -s KILL 900 python tools/bench_blocks.py $BLOCKS > gpurun_out/bench_blocks.log 2>&1; echo "rc=$?"; cut -c1-400 gpurun_out/bench_blocks.log; fi
