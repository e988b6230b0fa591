#!/bin/bash
# usage: tools/gpu_multi.sh "<N list>"   (run under gpurun --gpus maxN)
mkdir -p gpurun_out
if [ -n "$TESTS" ]; then echo "== pytest gpu"; timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log; fi
for N in $1; do
  echo "== bench N=$N"
  if [ "$N" = "1" ]; then
    timeout -s KILL 900 python bench.py --gpus 1 --steps 10 --warmup 3 --breakdown > gpurun_out/bench_n1.log 2>&1
  else
    timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2>&1
  fi
  echo "rc=$?"; tail -30 gpurun_out/bench_n$N.log | cut -c1-1500
done
