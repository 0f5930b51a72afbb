#!/bin/bash
# One parametrised script for every GPU-side job (run through `gpurun -- 'bash tools/gpu.sh <job> [args]'`).
# Everything a job prints goes to gpurun_out/<job>*.log so that it comes back with the call.
#   pcg [L] [iters]      per-iteration timing of the solver forms + converged solves (tools/pcg_iter_bench.py)
#   ccl                  labeling microbenchmark (tools/ccl_bench.py)
#   test [pytest args]   pytest -m gpu
#   bench [args]         bench.py
#   smoke                __graft_entry__.smoke()
#   launches [args]      ncu launch list of a short bench (gpu__time_duration per launch)
#   ncu <regex> <out> <cmd...>   one `ncu --set full` capture of the kernels matching <regex>
set -u
mkdir -p gpurun_out
job=${1:-test}; shift || true
case "$job" in
pcg)
    L=${1:-4096}; it=${2:-600}
    timeout 300 python tools/pcg_iter_bench.py --L "$L" --iters "$it" --converge > gpurun_out/pcg_L$L.log 2>&1
    tail -n 30 gpurun_out/pcg_L$L.log ;;
ccl)
    timeout 200 python tools/ccl_bench.py "$@" > gpurun_out/ccl.log 2>&1; cat gpurun_out/ccl.log ;;
test)
    timeout 1500 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/pytest_gpu.log 2>&1; tail -n 25 gpurun_out/pytest_gpu.log ;;
bench)
    timeout 900 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err ;;
smoke)
    timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -n 10 gpurun_out/smoke.log ;;
launches)
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 1 --warmup 1 "$@" > gpurun_out/launches.log 2>&1; tail -n 3 gpurun_out/launches.log ;;
ncu)
    rx=$1; out=$2; shift 2
    timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$rx" -c 3 -o gpurun_out/$out -f "$@" > gpurun_out/$out.log 2>&1
    tail -n 5 gpurun_out/$out.log ;;
*)  echo "unknown job $job"; exit 2 ;;
esac
