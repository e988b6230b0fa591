#!/bin/bash
# final pass of the round: full GPU test-suite, smoke, bench (both arms), block benches, ncu launch list + full profile
mkdir -p gpurun_out
echo "== pytest gpu"; timeout -s KILL 900 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-200
echo "== smoke"; timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n1.log | cut -c1-2500
echo "== bench original order"; timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --row-order original --no-cpu-baseline > gpurun_out/bench_n1_original.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n1_original.log | cut -c1-600
echo "== reference arm"; timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-700
echo "== blocks"; timeout -s KILL 600 python tools/bench_blocks.py c2 c3 c4 --reps 20 > gpurun_out/bench_blocks.log 2>&1; echo "rc=$?"; cut -c1-330 gpurun_out/bench_blocks.log
echo "== ncu launch list"; timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1c_step.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:'k_cat_sparse_csc|k_cat_pairs|k_dense_cross_runs|k_sparse_sandwich|k_dense_syrk_tc' -c 5 -o gpurun_out/prof_step_r1c -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1; echo "rc=$?"
