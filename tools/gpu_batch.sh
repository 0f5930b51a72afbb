#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "batch" 2>&1 | tail -15 > gpurun_out/pytest_batch.log
timeout 600 python tools/batch_bench.py > gpurun_out/batch_bench.log 2>&1
cat gpurun_out/pytest_batch.log gpurun_out/batch_bench.log
