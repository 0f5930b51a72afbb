for b in 0 1.5 2 2.5 3; do echo "== PERC_DEFL_BCOST=$b"; PERC_DEFL_BCOST=$b timeout 120 python tools/pcg_iter_bench.py --L 4096 --iters 600 --default-only 2>&1 | grep -E "used_fused" | tail -1; done
