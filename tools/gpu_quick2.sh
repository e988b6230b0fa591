#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.mem,clocks.max.mem,clocks.sm,temperature.gpu,power.draw --format=csv
echo "== bench onehot ON"; timeout -s KILL 900 python bench.py --steps 5 --breakdown --no-cpu-baseline 2>&1 | head -12
echo "== bench onehot OFF"; TABMAT_B200_ONEHOT=0 timeout -s KILL 900 python bench.py --steps 5 --breakdown --no-cpu-baseline 2>&1 | head -12
