#!/bin/bash
# FIRST GPU call of the next round (about 3 GPU-minutes): time and check the opt-in kernel variants that were
# written after round 1's GPU budget had ended (bit-identical to the defaults on the host emulations):
#   one-pass PCG kernel FtCfgA4 (perc_set_solver 15 / PERC_FUSED_CFG=6), labeling tile kernel PERC_CCL_VAR=1 | 2
mkdir -p gpurun_out
timeout 60 python tools/pcg_iter_bench.py --L 4096 --iters 600 --configs > gpurun_out/next_pcg_variants.log 2>&1
for v in 0 1 2; do
    PERC_CCL_VAR=$v timeout 60 python tools/ccl_bench.py > gpurun_out/next_ccl_var$v.log 2>&1
done
# parity of the variants on the GPU: the labeling tests under each labeling variant, the conductance tests that
# reach the one-pass kernel under FtCfgA4
for v in 1 2; do
    PERC_CCL_VAR=$v timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
        -k "labels or png or large_lattice or full_size or batch_equals or first_span" 2>&1 | tail -3 > gpurun_out/next_pytest_ccl_var$v.log
done
PERC_FUSED_CFG=6 timeout 120 python -m pytest $(cat tools/gpu_final_ids.txt) -q -x 2>&1 | tail -3 > gpurun_out/next_pytest_pcg_a4.log
tail -n +1 gpurun_out/next_*.log
