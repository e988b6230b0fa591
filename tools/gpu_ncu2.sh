#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/run_tc.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from tabmat_b200._lib import lib
from tabmat_b200.ext.dense import dense_sandwich
n, p = 4000000, int(sys.argv[1])
X = torch.randn((n, p), device="cuda"); d = torch.rand(n, device="cuda")
lib.tm_set_dense_f32_mode(2)
for _ in range(3):
    dense_sandwich(X, d, None, None)
torch.cuda.synchronize()
PY
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_syrk_tc -s 2 -c 1 -o gpurun_out/prof_tc_p128 -f python /tmp/run_tc.py 128 > gpurun_out/ncu_tc128.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_tc128.log
