#!/bin/bash
# pass-schedule sweep of tm_split_sandwich_blocks (TABMAT_B200_SCHED), row-sorted storage
mkdir -p gpurun_out
summ='import sys,json
for ln in sys.stdin:
    ln=ln.strip()
    if not ln.startswith("{"): continue
    d=json.loads(ln); print("ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), {k: round(v,2) for k,v in d["passes_ms"].items()}, "launches", d["gpu_launches"])'
for fused in 1 0; do
for s in 0 1 2 3 4 5; do
  echo "== sched $s fused=$fused"
  TABMAT_B200_INDEX_FUSED=$fused TABMAT_B200_SCHED=$s timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --row-order ${ORDER:-sorted} --no-cpu-baseline > gpurun_out/bench_sched_${s}_${fused}.log 2>&1; echo "rc=$?"; python -c "$summ" < gpurun_out/bench_sched_${s}_${fused}.log
done
done
timeout -s KILL 300 python -m pytest tests/test_gpu_index_fused.py tests/test_gpu_row_order.py tests/test_gpu_classes.py -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -3
