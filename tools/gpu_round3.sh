#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 200 python tools/ccl_bench.py > gpurun_out/ccl_bench.log 2>&1
SHORT="python bench.py --steps 1 --warmup 0 --itmax 150 --no-cpu-baseline --e2e-steps 0"
timeout 300 $SHORT > gpurun_out/short_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pcg_spmv2|pcg_update2' -s 20 -c 4 -o gpurun_out/prof_pcg2 -f $SHORT > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/pytest_gpu.log gpurun_out/ccl_bench.log gpurun_out/short_plain.log; tail -3 gpurun_out/ncu_full.log
