#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
timeout 300 python tools/ccl_bench.py > gpurun_out/ccl_bench.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/ccl_launches.csv python tools/ccl_bench.py quick > gpurun_out/ccl_ncu.log 2>&1
cat gpurun_out/pytest_gpu.log gpurun_out/ccl_bench.log
