set -u
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -s > gpurun_out/pytest_r2c.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/pytest_r2c.log | cut -c1-300
grep -h "^parity\|normwise / centred\|split f32, tf32\|ranks vs 1 GPU" gpurun_out/pytest_r2c.log | head -30
bash tools/gpu.sh smoke
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout -s KILL 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_r2c_$tag.log 2>&1
  echo "== $tag ($*) rc=$?"; tail -1 gpurun_out/bench_r2c_$tag.log | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.readline()); print({k:l.get(k) for k in ('ms_per_step','passes_ms')}, l.get('e2e',{}).get('ms_per_step'))
except Exception as e: print('no json', e)"
  grep -E "Traceback|Error" gpurun_out/bench_r2c_$tag.log | head -3
}
run gather X=1
run red TABMAT_B200_DXS=red
run gather_scw0 TABMAT_B200_TC_SCW=0
run gather_mb16 TABMAT_B200_GATHER_MB=16
run gather_mb64 TABMAT_B200_GATHER_MB=64
run gather_lag1 TABMAT_B200_GATHER_LAG=1
run gather_lag4 TABMAT_B200_GATHER_LAG=4
run gather_lag64 TABMAT_B200_GATHER_LAG=100000
for c in c2 c4; do
  echo "== config $c"
  timeout -s KILL 600 python bench.py --config $c --steps 10 --cpu-rows 200000 > gpurun_out/bench_r2c_$c.log 2>&1
  echo "rc=$?"; tail -1 gpurun_out/bench_r2c_$c.log | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print(l['ms_per_step'], l['value'], l['unit'], 'e2e', l['e2e']['ms_per_step'], 'roof', l['roofline']['frac'], l.get('parity',{}).get('max_normwise_err'))"
  grep -E "Traceback|Error" gpurun_out/bench_r2c_$c.log | head -3
done
echo "== c4 red"
TABMAT_B200_DXS=red timeout -s KILL 600 python bench.py --config c4 --steps 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print(l['ms_per_step'])"
echo "== launches"
C=200 TAG=r2c bash tools/gpu.sh launches
python - <<'PY'
import csv,collections
rows=list(csv.reader(l for l in open('gpurun_out/launches_r2c.csv') if l.startswith('"')))
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    try: agg[r[ki][:70]][0]+=1; agg[r[ki][:70]][1]+=float(r[vi].replace(',',''))
    except Exception: pass
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]: print(f"{t/1e6:9.3f} ms total {c:4d}x  {t/c/1e6:8.3f} ms each  {k}")
PY
