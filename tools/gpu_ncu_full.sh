#!/bin/bash
# ncu --set full of the step's big kernels (first launch of each) + the C3 histogram
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:'k_cat_sparse_csc|k_cat_pairs|k_dense_cross_runs|k_sparse_sandwich|k_dense_syrk_tc' -c 5 -o gpurun_out/prof_step_r1c -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_step.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:'k_cat_hist2' -s 3 -c 1 -o gpurun_out/prof_hist_r1c -f python tools/bench_blocks.py c3 --reps 3 > gpurun_out/ncu_hist.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_hist.log
ls -la gpurun_out/*.ncu-rep
