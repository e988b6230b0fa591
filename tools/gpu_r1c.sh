#!/bin/bash
# call 1 of the row-sorted work: new tests, parity suite, bench in both row orders, C3 block bench
mkdir -p gpurun_out
summ='import sys,json
for ln in sys.stdin:
    ln=ln.strip()
    if not ln.startswith("{"): continue
    d=json.loads(ln); print("ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), d["passes_ms"], "frac", round(d["roofline"]["frac"],3), d["roofline"]["kernel"][:30], "launches", d["gpu_launches"])'
echo "== new tests"; timeout -s KILL 400 python -m pytest tests/test_gpu_row_order.py tests/test_gpu_index_fused.py -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/pytest_row_order.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/pytest_row_order.log | cut -c1-220
echo "== bench sorted"; timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --row-order sorted --no-cpu-baseline > gpurun_out/bench_sorted.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench_sorted.log | cut -c1-600; python -c "$summ" < gpurun_out/bench_sorted.log
echo "== bench original"; timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --row-order original --no-cpu-baseline > gpurun_out/bench_original.log 2>&1; echo "rc=$?"; python -c "$summ" < gpurun_out/bench_original.log
echo "== bench sorted, index pass unfused"; TABMAT_B200_INDEX_FUSED=0 timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --row-order sorted --no-cpu-baseline > gpurun_out/bench_sorted_unfused.log 2>&1; echo "rc=$?"; python -c "$summ" < gpurun_out/bench_sorted_unfused.log
echo "== bench sorted, scatter first"; TABMAT_B200_SCATTER_FIRST=1 timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --row-order sorted --no-cpu-baseline > gpurun_out/bench_sorted_sf.log 2>&1; echo "rc=$?"; python -c "$summ" < gpurun_out/bench_sorted_sf.log
echo "== blocks c3"; timeout -s KILL 200 python tools/bench_blocks.py c3 --reps 20 > gpurun_out/bench_blocks_c3.log 2>&1; echo "rc=$?"; cut -c1-260 gpurun_out/bench_blocks_c3.log
echo "== pytest gpu (without the full-size file)"; timeout -s KILL 500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short --ignore tests/test_gpu_scale.py --ignore tests/test_gpu_row_order.py --ignore tests/test_gpu_index_fused.py > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-220
echo "== ncu launch list (sorted)"; timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|tmb|tc::' --csv --log-file gpurun_out/launches_sorted.csv python bench.py --steps 1 --warmup 1 --row-order sorted --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_sorted.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; ix={h:i for i,h in enumerate(hdr)}
seq=[]
for r in rows[hi+1:]:
    if len(r)<len(hdr) or r[ix['Metric Name']]!='gpu__time_duration.sum': continue
    name=r[ix['Kernel Name']].split('(')[0][-50:]
    v=float(r[ix['Metric Value']].replace(',','')); unit=r[ix['Metric Unit']]
    ms = v/1e6 if unit in('ns','nsecond') else v/1e3 if unit in ('us','usecond') else v
    seq.append((name,ms))
idxs=[i for i,(n,_) in enumerate(seq) if 'k_dense_syrk_tc' in n]
print("total launches", len(seq), "syrk at", idxs[:8])
if idxs:
    a=idxs[-1]-2; b=min(len(seq), a+40)
    for n,ms in seq[a:b]: print(f"{ms:9.3f} ms  {n}")
PY
