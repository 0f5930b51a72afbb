#!/bin/bash
# GPU session: parity tests + labeling microbenchmark + per-kernel launch list + source-level profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 300 python tools/ccl_bench.py > gpurun_out/ccl_bench.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/ccl_launches.csv python tools/ccl_bench.py quick > gpurun_out/ccl_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ccl_local|ccl_rootfix|ccl_merge' -c 3 -o gpurun_out/prof_ccl2 -f python tools/ccl_bench.py quick > gpurun_out/ncu_ccl2.log 2>&1
cat gpurun_out/pytest_gpu.log gpurun_out/ccl_bench.log
