#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --itmax 3000 --no-cpu-baseline --e2e-steps 0 > gpurun_out/short_plain.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/short_plain.log').read().strip().splitlines()[-1])
print("spmv ms", d["roofline"]["avg_launch_ms"], "frac", d["roofline"]["frac"])
for k,v in d["extra"].items():
    if isinstance(v,dict) and "avg_launch_ms" in v: print(k, v["avg_launch_ms"], v["frac"])
print("ms/iter", d["extra"]["pcg_ms_per_iteration"], "G", d["extra"]["mean_G"])
PY
