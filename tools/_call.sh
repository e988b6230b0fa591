set -u
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_r2a.log 2>&1
echo "pytest rc=$?"; tail -40 gpurun_out/pytest_r2a.log | cut -c1-300
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
echo "== legacy index path (records, plain CSC)"
TABMAT_B200_CSC_PACKED=0 TABMAT_B200_CSC_ROW_BLOCKS=0 timeout -s KILL 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_r2a_legacyidx.log 2>&1
tail -1 gpurun_out/bench_r2a_legacyidx.log | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print({k:l.get(k) for k in ('ms_per_step','passes_ms')})"
echo "== packed, plain CSC"
TABMAT_B200_CSC_ROW_BLOCKS=0 timeout -s KILL 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_r2a_packedplain.log 2>&1
tail -1 gpurun_out/bench_r2a_packedplain.log | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print({k:l.get(k) for k in ('ms_per_step','passes_ms')})"
for c in c2 c3 c4; do
  echo "== config $c"
  timeout -s KILL 600 python bench.py --config $c --steps 10 --cpu-rows 200000 > gpurun_out/bench_r2a_$c.log 2>&1
  echo "rc=$?"; tail -1 gpurun_out/bench_r2a_$c.log | cut -c1-1800
  grep -E "Traceback|Error" gpurun_out/bench_r2a_$c.log | head -3
done
