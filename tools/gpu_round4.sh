#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
python bench.py --steps 1 --warmup 1 --itmax 3000 --no-cpu-baseline --e2e-steps 0 > gpurun_out/short_plain.log 2>&1
cat gpurun_out/pytest_gpu.log gpurun_out/short_plain.log
