#!/bin/bash
mkdir -p gpurun_out
echo "== ncu launch list of bench.py (full n, 2 steps)"
timeout -s KILL 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"
wc -l gpurun_out/launches_r1.csv
echo "== ncu full: tcgen05 SYRK at C2 (n=1e7,p=256)"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:k_dense_syrk_tc -s 3 -c 2 -o gpurun_out/prof_syrk_tc_r1 -f python tools/bench_blocks.py c2 --reps 2 > gpurun_out/ncu_syrk.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_syrk.log
ls -la gpurun_out/*.ncu-rep
