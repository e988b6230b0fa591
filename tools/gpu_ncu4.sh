#!/bin/bash
mkdir -p gpurun_out
echo "== bench"; timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n1.log | cut -c1-3000
echo "== ncu full on the two big kernels"
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k "regex:k_dense_cross_fused|k_dense_syrk_tc" -s 4 -c 2 -o gpurun_out/prof_split_r1 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_split.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_split.log
