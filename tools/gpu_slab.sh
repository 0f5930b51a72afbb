#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29533 tests/slab_gpu_worker.py > gpurun_out/slab_worker.log 2>&1
echo "rc=$?" >> gpurun_out/slab_worker.log
tail -40 gpurun_out/slab_worker.log
