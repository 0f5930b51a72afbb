#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pcg_pipe' -s 200 -c 2 -o gpurun_out/prof_pcg3 -f python bench.py --steps 1 --warmup 0 --itmax 400 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_pcg3.log 2>&1
tail -3 gpurun_out/ncu_pcg3.log
