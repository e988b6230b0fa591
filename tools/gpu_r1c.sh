#!/bin/bash
# call 1 of the row-sorted work: new tests, parity suite, bench in both row orders, C3 block bench
mkdir -p gpurun_out
summ='import sys,json
for ln in sys.stdin:
    ln=ln.strip()
    if not ln.startswith("{"): continue
    d=json.loads(ln); print("ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), d["passes_ms"], "frac", round(d["roofline"]["frac"],3), d["roofline"]["kernel"][:30], "launches", d["gpu_launches"])'
echo "== new tests"; timeout -s KILL 400 python -m pytest tests/test_gpu_row_order.py -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/pytest_row_order.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/pytest_row_order.log | cut -c1-220
echo "== bench sorted"; timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --row-order sorted --no-cpu-baseline > gpurun_out/bench_sorted.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench_sorted.log | cut -c1-600; python -c "$summ" < gpurun_out/bench_sorted.log
echo "== bench original"; timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --row-order original --no-cpu-baseline > gpurun_out/bench_original.log 2>&1; echo "rc=$?"; python -c "$summ" < gpurun_out/bench_original.log
echo "== bench sorted, scatter first"; TABMAT_B200_SCATTER_FIRST=1 timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --row-order sorted --no-cpu-baseline > gpurun_out/bench_sorted_sf.log 2>&1; echo "rc=$?"; python -c "$summ" < gpurun_out/bench_sorted_sf.log
echo "== bench original, runs kernel forced"; TABMAT_B200_CROSS_RUNS=1 timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --row-order original --no-cpu-baseline > gpurun_out/bench_original_runs.log 2>&1; echo "rc=$?"; python -c "$summ" < gpurun_out/bench_original_runs.log
echo "== blocks c3"; timeout -s KILL 200 python tools/bench_blocks.py c3 --reps 20 > gpurun_out/bench_blocks_c3.log 2>&1; echo "rc=$?"; cut -c1-260 gpurun_out/bench_blocks_c3.log
echo "== blocks c3 (v1 kernels)"; TABMAT_B200_CAT_V1=1 timeout -s KILL 200 python tools/bench_blocks.py c3 --reps 20 > gpurun_out/bench_blocks_c3_v1.log 2>&1; echo "rc=$?"; cut -c1-260 gpurun_out/bench_blocks_c3_v1.log
echo "== pytest gpu (without the full-size file)"; timeout -s KILL 500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short --ignore tests/test_gpu_scale.py --deselect tests/test_gpu_row_order.py > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-220
