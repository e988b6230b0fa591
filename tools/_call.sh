set -u
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_r2b.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/pytest_r2b.log | cut -c1-300
grep -h "parity\|normwise\|tf32" gpurun_out/pytest_r2b.log | head -20
bash tools/gpu.sh smoke
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout -s KILL 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_r2b_$tag.log 2>&1
  echo "== $tag ($*) rc=$?"; tail -1 gpurun_out/bench_r2b_$tag.log | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.readline()); print({k:l.get(k) for k in ('ms_per_step','passes_ms')}, l.get('e2e',{}).get('ms_per_step'))
except Exception as e: print('no json', e)"
  grep -E "Traceback|Error" gpurun_out/bench_r2b_$tag.log | head -3
}
run base X=1
run scw8ilp TABMAT_B200_TC_SCW=8
run scw4ilp TABMAT_B200_TC_SCW=4
run scw8ilp_sb2 TABMAT_B200_TC_SCW=8 TABMAT_B200_TC_SB=2
run scatilp TABMAT_B200_SCATTER_ILP=1
run scatilp3 TABMAT_B200_SCATTER_ILP=1 TABMAT_B200_SCATTER_CTAS=3
run sched4_c4 TABMAT_B200_SCHED=4 TABMAT_B200_SCATTER_CTAS=4
run sched4_c4_ilp TABMAT_B200_SCHED=4 TABMAT_B200_SCATTER_CTAS=3 TABMAT_B200_SCATTER_ILP=1
run sched4_c2_ilp TABMAT_B200_SCHED=4 TABMAT_B200_SCATTER_CTAS=2 TABMAT_B200_SCATTER_ILP=1
run sched2_ilp TABMAT_B200_SCHED=2 TABMAT_B200_SCATTER_CTAS=2 TABMAT_B200_SCATTER_ILP=1
run tf32x3 TABMAT_B200_DENSE_F32_MODE=3
for c in c2 c3 c4; do
  echo "== config $c"
  timeout -s KILL 600 python bench.py --config $c --steps 10 --cpu-rows 200000 > gpurun_out/bench_r2b_$c.log 2>&1
  echo "rc=$?"; tail -1 gpurun_out/bench_r2b_$c.log | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print(l['ms_per_step'], l['value'], l['unit'], 'e2e', l['e2e']['ms_per_step'], 'roof', l['roofline']['frac'], l['config'].get('launch'), l.get('parity',{}).get('max_normwise_err'))"
  grep -E "Traceback|Error" gpurun_out/bench_r2b_$c.log | head -3
done
echo "== launches"
C=300 TAG=r2b bash tools/gpu.sh launches
python - <<'PY'
import csv,collections
rows=list(csv.reader(l for l in open('gpurun_out/launches_r2b.csv') if l.startswith('"')))
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    try: agg[r[ki][:60]][0]+=1; agg[r[ki][:60]][1]+=float(r[vi].replace(',',''))
    except Exception: pass
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]: print(f"{t/1e6:9.3f} ms {c:4d}x {k}")
PY
