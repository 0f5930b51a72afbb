"""slab-mode benchmark (SURVEY 8e mode 2): one square site lattice at p_c decomposed over the ranks;
labeling + stitch timed per call, then a fixed number of distributed PCG iterations.
   torchrun --nproc-per-node N tools/slab_bench.py --L 16384 [--iters 200]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import percolation_b200 as P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=16384)
    ap.add_argument("--p", type=float, default=0.60)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    dist.init_process_group("gloo")
    r, G = dist.get_rank(), dist.get_world_size()
    dev = int(os.environ.get("LOCAL_RANK", r))
    torch.cuda.set_device(dev)
    uid = torch.from_numpy(P.comm_unique_id() if r == 0 else np.zeros(128, np.uint8))
    dist.broadcast(uid, 0)
    Lsz = args.L
    S = P.SlabLattice(P.SQUARE, Lsz, Lsz, 0, dev, G, r, unique_id=uid.numpy())
    t = Lsz * Lsz
    t0 = time.perf_counter()
    S.generate(20240611, 0, int(args.p * t), -1)
    t_gen = time.perf_counter() - t0
    times = []
    for rep in range(args.reps):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        S.label(P.SITE)
        torch.cuda.synchronize(); dist.barrier()
        times.append(time.perf_counter() - t0)
    ph = S.phase_ms()
    sm = S.summary()
    out = {"L": Lsz, "ranks": G, "p": args.p, "generate_s": t_gen, "label_wall_ms": 1e3 * min(times),
           "label_device_ms": float(ph[1] + ph[2] + ph[3] + ph[4]), "mask_ms": float(ph[0]),
           "gsites_per_s_wall": t / min(times) / 1e9, "ncl": sm["ncl"], "maxcs": sm["maxcs"], "nspan": sm["nspan"]}
    if sm["nspan"] and args.iters > 0:
        S.conduct(0, tol=1e-30, itmax=15, voltages=False)      # warm-up: NCCL sets its channels up on first use
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = S.conduct(0, tol=1e-30, itmax=args.iters - 1, voltages=False)
        torch.cuda.synchronize(); dist.barrier()
        dt = time.perf_counter() - t0
        ph = S.phase_ms()
        own = (S.yb - S.ya) * Lsz
        out.update({"pcg_iters": res["iter"], "pcg_ms_per_iter": float(ph[5]) / res["iter"], "pcg_wall_ms_per_iter": 1e3 * dt / res["iter"],
                    "pcg_gbs_per_gpu": 50.0 * own / (float(ph[5]) / res["iter"] * 1e-3) / 1e9,
                    "spmv_ms": float(ph[6]), "update_ms": float(ph[7]), "G_after_iters": res["Gtop"], "err": res["err"]})
    if r == 0:
        print(json.dumps(out), flush=True)
    S.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
