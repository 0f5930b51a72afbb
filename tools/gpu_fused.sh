#!/bin/bash
# one-pass PCG kernel: per-iteration timing of its variants at the bench size first (short timeout: a hang
# costs seconds, not the budget), then the GPU test-suite and smoke
mkdir -p gpurun_out
timeout 40 python tools/pcg_iter_bench.py --L 4096 --iters 600 --configs > gpurun_out/fused_variants.log 2>&1
echo "rc=$?" >> gpurun_out/fused_variants.log
cat gpurun_out/fused_variants.log
timeout 150 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 40 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/smoke.log
