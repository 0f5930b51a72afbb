#!/bin/bash
# tile configurations of the one-pass PCG kernel at the bench size, then one ncu --set full capture of each
mkdir -p gpurun_out
timeout 60 python tools/pcg_iter_bench.py --L 4096 --iters 600 --configs > gpurun_out/fused_cfgs.log 2>&1
echo "rc=$?" >> gpurun_out/fused_cfgs.log
cat gpurun_out/fused_cfgs.log
timeout 55 ncu --set full --import-source on --clock-control none --kernel-id ::regex:pcg_fused_kernel:8 -f -o gpurun_out/prof_fused \
    python tools/pcg_iter_bench.py --L 4096 --iters 12 --configs --fused-only > gpurun_out/ncu_fused.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_fused.log
tail -5 gpurun_out/ncu_fused.log
ls -la gpurun_out/prof_fused.ncu-rep
