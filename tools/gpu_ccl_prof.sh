#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ccl_local|ccl_rootfix|ccl_span' -c 3 -o gpurun_out/prof_ccl2 -f python tools/ccl_bench.py quick > gpurun_out/ncu_ccl2.log 2>&1
tail -5 gpurun_out/ncu_ccl2.log
