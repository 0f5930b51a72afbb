#!/bin/bash
# round-end evidence on one B200 (no full ncu captures): GPU tests, smoke, full bench (both arms), labeling bench, launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 python tools/ccl_bench.py > gpurun_out/ccl_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 0 --itmax 300 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_launches.log 2>&1
cat gpurun_out/pytest_gpu.log gpurun_out/smoke.log; cut -c1-300 gpurun_out/bench.json; cut -c1-200 gpurun_out/bench_ref.json; cat gpurun_out/ccl_bench.log
