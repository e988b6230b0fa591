#!/bin/bash
# One entry point for everything that runs on the B200 box through gpurun:
#   gpurun --timeout 1500 -- 'bash tools/gpu.sh <mode> [args]'
# Every mode writes its logs under gpurun_out/ (merged back by gpurun).
#   tests [pytest args]   pytest -m gpu (default: whole suite)
#   smoke                 __graft_entry__.smoke()
#   bench [bench args]    bench.py at N=1 -> gpurun_out/bench_$TAG.log
#   benchn N [args]       bench.py under torch.distributed.run on N GPUs
#   ref                   bench.py --impl reference
#   blocks [c2 c3 c4]     tools/bench_blocks.py
#   launches [args]       ncu launch list (gpu__time_duration) of a short bench.py run
#   ncu REGEX [args]      ncu --set full of the first launches matching REGEX in bench.py
#   sweep VAR v1 v2 ...   bench.py once per value of the environment variable VAR
#   sanitize              compute-sanitizer memcheck + racecheck on the small parity cases
# TAG (env) names the log files; default = mode.
set -u
mkdir -p gpurun_out
mode=${1:-tests}; shift || true
TAG=${TAG:-$mode}
case "$mode" in
  tests)
    args=${*:-tests}
    timeout -s KILL ${T:-1500} python -m pytest $args -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_$TAG.log 2>&1
    echo "pytest rc=$?"; tail -15 gpurun_out/pytest_$TAG.log ;;
  smoke)
    timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ;;
  bench)
    timeout -s KILL ${T:-900} python bench.py "$@" > gpurun_out/bench_$TAG.log 2>&1
    echo "bench rc=$?"; tail -1 gpurun_out/bench_$TAG.log | cut -c1-${CUT:-2500}
    grep -E "Traceback|Error" gpurun_out/bench_$TAG.log | head -5 ;;
  benchn)
    N=$1; shift
    timeout -s KILL ${T:-900} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/bench_${TAG}_n$N.log 2>&1
    echo "bench n=$N rc=$?"; tail -1 gpurun_out/bench_${TAG}_n$N.log | cut -c1-${CUT:-2500}
    grep -i -E "Traceback|Error" gpurun_out/bench_${TAG}_n$N.log | head -5 ;;
  ref)
    timeout -s KILL ${T:-900} python bench.py --impl reference "$@" > gpurun_out/bench_ref_$TAG.log 2>&1
    echo "ref rc=$?"; tail -1 gpurun_out/bench_ref_$TAG.log | cut -c1-1500 ;;
  blocks)
    timeout -s KILL ${T:-900} python tools/bench_blocks.py "$@" > gpurun_out/bench_blocks_$TAG.log 2>&1
    echo "blocks rc=$?"; cut -c1-600 gpurun_out/bench_blocks_$TAG.log ;;
  launches)
    # only this library's kernels (namespaces tmb:: / tmb::tc::), so that the construction of the
    # matrix (torch sort / index kernels) does not fill the launch list
    timeout -s KILL ${T:-900} ncu --metrics gpu__time_duration.sum --clock-control none \
      --kernel-name-base demangled -k "regex:tmb::" -c ${C:-400} --csv \
      --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline "$@" \
      > gpurun_out/launches_$TAG.log 2>&1
    echo "launches rc=$?"; tail -2 gpurun_out/launches_$TAG.log | cut -c1-300 ;;
  ncu)
    rx=$1; shift
    timeout -s KILL ${T:-1200} ncu --set full --clock-control none --import-source on -k "regex:$rx" \
      -s ${S:-0} -c ${C:-4} -o gpurun_out/prof_$TAG -f python ${PROG:-bench.py --steps 1 --warmup 1 --no-cpu-baseline} "$@" \
      > gpurun_out/ncu_$TAG.log 2>&1
    echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$TAG.log | cut -c1-300; ls -la gpurun_out/prof_$TAG.ncu-rep
    # gpurun_out/ is capped at 64 MiB: keep the raw-metric CSV (every metric of every captured
    # launch) and drop the report unless KEEP_REP=1
    ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_raw.csv 2>/dev/null
    if [ -z "${KEEP_REP:-}" ]; then rm -f gpurun_out/prof_$TAG.ncu-rep; fi
    ls -la gpurun_out/ncu_${TAG}_raw.csv ;;
  sweep)
    var=$1; shift
    for v in "$@"; do
      env $var=$v timeout -s KILL ${T:-600} python bench.py ${BENCH_ARGS:---steps 5 --no-cpu-baseline --no-parity} > gpurun_out/bench_${TAG}_${var}_$v.log 2>&1
      echo "== $var=$v rc=$?"; tail -1 gpurun_out/bench_${TAG}_${var}_$v.log | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.readline()); print({k:l.get(k) for k in ('ms_per_step','passes_ms')}, l.get('e2e',{}).get('ms_per_step'))
except Exception as e: print('no json', e)"
    done ;;
  sanitize)
    bash tools/sanitize.sh ;;
  *) echo "unknown mode $mode"; exit 2 ;;
esac
