#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n2.log | cut -c1-1500
grep -i -E "error|Traceback" gpurun_out/bench_n2.log | head -5
