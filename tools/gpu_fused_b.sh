#!/bin/bash
# final check of the one-pass PCG kernel (default variant): the conductance tests that reach it, then one
# ncu --set full capture of one launch at the bench size
mkdir -p gpurun_out
timeout 50 python -m pytest $(cat tools/gpu_final_ids.txt) -q --maxfail=5 2>&1 | tail -15 > gpurun_out/pytest_gpu_conduct.log
cat gpurun_out/pytest_gpu_conduct.log
timeout 30 ncu --set full --import-source on --clock-control none -k regex:pcg_fused_kernel -s 8 -c 1 -f -o gpurun_out/prof_fused_v3 \
    python tools/pcg_iter_bench.py --L 4096 --iters 12 --default-only > gpurun_out/ncu_fused_v3.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_fused_v3.log
tail -4 gpurun_out/ncu_fused_v3.log
