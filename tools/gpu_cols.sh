#!/bin/bash
mkdir -p gpurun_out
summ='import sys,json
for ln in sys.stdin:
    ln=ln.strip()
    if not ln.startswith("{"): continue
    d=json.loads(ln); print("ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), {k: round(v,2) for k,v in d["passes_ms"].items()}, "launches", d["gpu_launches"])'
run() {
  echo "== row blocks=$1 block=$2"
  TABMAT_B200_CSC_ROW_BLOCKS=$1 TABMAT_B200_CSC_ROW_BLOCK=$2 timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cols_$1_$2.log 2>&1
  echo "rc=$?"
  python -c "$summ" < gpurun_out/bench_cols_$1_$2.log
}
run 1 524288
run 1 1048576
run 1 262144
