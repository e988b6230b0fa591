set -u
mkdir -p gpurun_out
TAG=r2a bash tools/gpu.sh tests tests/test_gpu_fused_scatter.py tests/test_gpu_index_fused.py tests/test_gpu_row_order.py
bash tools/gpu.sh smoke
for cfg in "4 4" "8 4" "0 4" "4 2" "8 2"; do
  set -- $cfg
  echo "== SCW=$1 SB=$2"
  TABMAT_B200_TC_SCW=$1 TABMAT_B200_TC_SB=$2 timeout -s KILL 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_r2a_scw$1_sb$2.log 2>&1
  echo "rc=$?"; tail -1 gpurun_out/bench_r2a_scw$1_sb$2.log | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.readline()); print({k:l.get(k) for k in ('ms_per_step','passes_ms')}, l.get('e2e',{}).get('ms_per_step'))
except Exception as e: print('no json', e)"
  grep -E "Traceback|Error" gpurun_out/bench_r2a_scw$1_sb$2.log | head -3
done
