#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 300 python tools/probe2.py > gpurun_out/probe2.log 2>&1
timeout 200 python tools/ccl_bench.py > gpurun_out/ccl_bench.log 2>&1
timeout 200 python tools/ccl_bench.py quick > gpurun_out/ccl_quick.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ccl_|build_mask' -c 8 -o gpurun_out/prof_ccl -f python tools/ccl_bench.py quick > gpurun_out/ncu_ccl.log 2>&1
cat gpurun_out/pytest_gpu.log gpurun_out/probe2.log gpurun_out/ccl_bench.log; tail -3 gpurun_out/ncu_ccl.log
