#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_native.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_native.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; ix={h:i for i,h in enumerate(hdr)}
seq=[]
for r in rows[hi+1:]:
    if len(r)<len(hdr) or r[ix['Metric Name']]!='gpu__time_duration.sum': continue
    name=r[ix['Kernel Name']].split('(')[0][-60:]
    v=float(r[ix['Metric Value']].replace(',','')); unit=r[ix['Metric Unit']]
    ms = v/1e6 if unit in('ns','nsecond') else v/1e3 if unit in ('us','usecond') else v
    seq.append((name,ms))
# find the last occurrence of the syrk kernel and print a window of one step around it
idxs=[i for i,(n,_) in enumerate(seq) if 'k_dense_syrk_tc' in n]
print("syrk launches at", idxs[:10], "total launches", len(seq))
if len(idxs)>=2:
    a,b=idxs[1]-3, idxs[2]-3 if len(idxs)>2 else idxs[1]+60
    tot=0
    for n,ms in seq[a:b]:
        print(f"{ms:9.3f} ms  {n}")
        tot+=ms
    print("window total", tot)
PY
