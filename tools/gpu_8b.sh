#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node=8 --master-port 29542 tools/slab_bench.py --L 65536 --p 0.60 --iters 100 > gpurun_out/slab_pcg_65536_n8.log 2>&1; echo "rc=$?" >> gpurun_out/slab_pcg_65536_n8.log
timeout 600 $TR --nproc-per-node=8 --master-port 29543 tools/slab_bench.py --L 16384 --p 0.60 --iters 200 > gpurun_out/slab_pcg_16384_n8.log 2>&1
timeout 600 $TR --nproc-per-node=4 --master-port 29547 tools/slab_bench.py --L 16384 --p 0.60 --iters 200 > gpurun_out/slab_pcg_16384_n4.log 2>&1
timeout 600 $TR --nproc-per-node=2 --master-port 29544 tools/slab_bench.py --L 16384 --p 0.60 --iters 200 > gpurun_out/slab_pcg_16384_n2.log 2>&1
timeout 600 $TR --nproc-per-node=1 --master-port 29545 tools/slab_bench.py --L 16384 --p 0.60 --iters 200 > gpurun_out/slab_pcg_16384_n1.log 2>&1
grep -h "^{" gpurun_out/slab_pcg_*.log | cut -c1-900; tail -3 gpurun_out/slab_pcg_65536_n8.log | cut -c1-300
