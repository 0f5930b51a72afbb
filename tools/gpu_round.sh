#!/bin/bash
# one GPU-box session: tests, smoke, bench (both arms), ncu launch list + one full capture
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
SHORT="python bench.py --steps 1 --warmup 0 --itmax 150 --no-cpu-baseline --e2e-steps 0"
timeout 300 $SHORT > gpurun_out/short_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_launches.log 2>&1
timeout 300 $SHORT > gpurun_out/short_plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pcg_spmv -s 20 -c 2 -o gpurun_out/prof_spmv -f $SHORT > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_ref.json; tail -5 gpurun_out/ncu_launches.log; tail -5 gpurun_out/ncu_full.log; ls -la gpurun_out
