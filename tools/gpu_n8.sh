#!/bin/bash
mkdir -p gpurun_out
N=${N:-8}
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n$N.log | cut -c1-1200
grep -i -E "error|Traceback" gpurun_out/bench_n$N.log | head -5
