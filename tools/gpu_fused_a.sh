#!/bin/bash
# per-iteration timing of the one-pass PCG kernel's variants at the bench size (and their agreement with the
# two-kernel form after 600 iterations)
mkdir -p gpurun_out
timeout 40 python tools/pcg_iter_bench.py --L 4096 --iters 600 --configs > gpurun_out/fused_variants.log 2>&1
echo "rc=$?" >> gpurun_out/fused_variants.log
cat gpurun_out/fused_variants.log
