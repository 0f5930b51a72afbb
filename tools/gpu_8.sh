#!/bin/bash
# 8-GPU session: slab mode at L=65536 (C5) and L=16384, slab parity worker on 8 ranks, realization-sharded bench
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/gpus8.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node=8 --master-port 29541 tests/slab_gpu_worker.py > gpurun_out/slab_worker8.log 2>&1; echo "rc=$?" >> gpurun_out/slab_worker8.log
timeout 600 $TR --nproc-per-node=8 --master-port 29542 tools/slab_bench.py --L 65536 --p 0.5927 --iters 100 > gpurun_out/slab_bench_65536_n8.log 2>&1; echo "rc=$?" >> gpurun_out/slab_bench_65536_n8.log
timeout 600 $TR --nproc-per-node=8 --master-port 29543 tools/slab_bench.py --L 16384 --p 0.5927 --iters 200 > gpurun_out/slab_bench_16384_n8.log 2>&1
timeout 600 $TR --nproc-per-node=2 --master-port 29544 tools/slab_bench.py --L 16384 --p 0.5927 --iters 200 > gpurun_out/slab_bench_16384_n2.log 2>&1
timeout 600 $TR --nproc-per-node=1 --master-port 29545 tools/slab_bench.py --L 16384 --p 0.5927 --iters 200 > gpurun_out/slab_bench_16384_n1.log 2>&1
timeout 900 $TR --nproc-per-node=8 --master-port 29546 bench.py --gpus 8 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench_n8.log 2>&1
tail -3 gpurun_out/slab_worker8.log; grep -h "^{" gpurun_out/slab_bench_*.log gpurun_out/bench_n8.log | cut -c1-900
