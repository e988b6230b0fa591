#!/bin/bash
# quick iteration: index/row-order tests, bench at two schedules, per-kernel launch list
mkdir -p gpurun_out
summ='import sys,json
for ln in sys.stdin:
    ln=ln.strip()
    if not ln.startswith("{"): continue
    d=json.loads(ln); print("ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), {k: round(v,2) for k,v in d["passes_ms"].items()}, "launches", d["gpu_launches"])'
timeout -s KILL 300 python -m pytest tests/test_gpu_index_fused.py tests/test_gpu_row_order.py tests/test_gpu_classes.py -m gpu -q -p no:cacheprovider --tb=short -x 2>&1 | tail -15 | cut -c1-200
for s in ${SCHEDS:-0 3}; do
  echo "== sched $s"
  TABMAT_B200_SCHED=$s timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --row-order ${ORDER:-sorted} --no-cpu-baseline > gpurun_out/bench_iter_$s.log 2>&1; echo "rc=$?"; python -c "$summ" < gpurun_out/bench_iter_$s.log
done
echo "== ncu launch list"; TABMAT_B200_SCHED=3 timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|tmb|tc::' --csv --log-file gpurun_out/launches_iter.csv python bench.py --steps 1 --warmup 1 --row-order ${ORDER:-sorted} --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_iter.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; ix={h:i for i,h in enumerate(hdr)}
seq=[]
for r in rows[hi+1:]:
    if len(r)<len(hdr) or r[ix['Metric Name']]!='gpu__time_duration.sum': continue
    name=r[ix['Kernel Name']].split('(')[0][-50:]
    v=float(r[ix['Metric Value']].replace(',','')); unit=r[ix['Metric Unit']]
    ms = v/1e6 if unit in('ns','nsecond') else v/1e3 if unit in ('us','usecond') else v
    seq.append((name,ms))
idxs=[i for i,(n,_) in enumerate(seq) if 'k_permute' in n]
print("total launches", len(seq))
if idxs:
    a=idxs[-1]; b=len(seq)
    for n,ms in seq[a:b]:
        if ms >= 0.03: print(f"{ms:9.3f} ms  {n}")
    print("sum", sum(ms for _,ms in seq[a:b]))
PY
