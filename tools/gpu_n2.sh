#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 1 --warmup 1 --itmax 3000 --e2e-steps 1 > gpurun_out/bench_n2.log 2>&1
echo "rc=$?"; grep "^{" gpurun_out/bench_n2.log | cut -c1-300; tail -5 gpurun_out/bench_n2.log | cut -c1-300
timeout 600 python -m pytest tests/test_slab_gpu.py -q -m gpu 2>&1 | tail -3
