#!/bin/bash
echo "== side stream ON"; timeout -s KILL 900 python bench.py --steps 5 --breakdown --no-cpu-baseline 2>&1 | head -4
echo "== side stream OFF"; TABMAT_B200_SIDE_STREAM=0 timeout -s KILL 900 python bench.py --steps 5 --breakdown --no-cpu-baseline 2>&1 | head -4
