#!/bin/bash
# one-pass PCG kernel: quick check first (short timeout: a hang costs seconds, not the budget), then the GPU
# test-suite, smoke and the per-iteration timing at the bench size
mkdir -p gpurun_out
timeout 40 python tools/pcg_iter_bench.py --L 1024 --iters 300 --converge > gpurun_out/fused_quick.log 2>&1
echo "rc=$?" >> gpurun_out/fused_quick.log
cat gpurun_out/fused_quick.log
if grep -q FUSED_OK gpurun_out/fused_quick.log; then
    timeout 40 python tools/pcg_iter_bench.py --L 4096 --iters 800 > gpurun_out/fused_4096.log 2>&1
    cat gpurun_out/fused_4096.log
    timeout 120 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
    cat gpurun_out/pytest_gpu.log
    timeout 40 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
    tail -3 gpurun_out/smoke.log
fi
