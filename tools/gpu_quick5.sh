#!/bin/bash
for mode in "index-first" "scatter-first" "no-side-stream"; do
  echo "== $mode"
  if [ "$mode" = "scatter-first" ]; then export TABMAT_B200_SCATTER_FIRST=1; else unset TABMAT_B200_SCATTER_FIRST; fi
  if [ "$mode" = "no-side-stream" ]; then export TABMAT_B200_SIDE_STREAM=0; else unset TABMAT_B200_SIDE_STREAM; fi
  timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['passes_ms'])"
done
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "native or fused" 2>&1 | tail -2
