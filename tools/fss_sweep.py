"""BASELINE configs[3]: triangular-lattice site percolation finite-size-scaling sweep.
For L in 1024 .. 16384 and 17 occupation fractions p = p_c + j * 0.25 * L^(-3/4), j = -8 .. 8 (p_c = 1/2), `--nreal`
realizations per point are labeled on the device (perc_batch: occupancy + labeling + spanning + largest cluster, no
host synchronisation inside the batch); the table gives the spanning probability Pi(p, L) and the mean largest-cluster
fraction P_max(p, L) -- the quantities whose crossing / collapse with L^(-1/nu), L^(-beta/nu) the sweep is for.

One GPU:      python tools/fss_sweep.py [--Lmax 16384] [--nreal 32]
Several GPUs: torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/fss_sweep.py ...
              realizations are sharded round-robin over the ranks (disjoint Philox streams), no data-path collective,
              ONE all-reduce of the integer statistics per table (percolation_b200.shard, multi-GPU mode 1).
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import percolation_b200 as P
from percolation_b200.shard import my_realizations, STREAM_STRIDE

ap = argparse.ArgumentParser()
ap.add_argument("--Lmin", type=int, default=1024)
ap.add_argument("--Lmax", type=int, default=16384)
ap.add_argument("--nreal", type=int, default=32, help="realizations per (L, p) point, all ranks together")
ap.add_argument("--seed", type=int, default=20240611)
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
dist = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

PC, NPTS = 0.5, 17
Ls = []
L_ = args.Lmin
while L_ <= args.Lmax:
    Ls.append(L_)
    L_ *= 2
mine = my_realizations(args.nreal, rank, world)          # realization indices of this rank
FIELDS = ("realizations", "spanning", "sum_maxcs", "sum_ncl", "failed")
table = np.zeros((len(Ls), NPTS, len(FIELDS)), np.int64)
t_all = time.perf_counter()
for a, Lsz in enumerate(Ls):
    with P.Lattice(P.TRIANGULAR, Lsz, Lsz, 0, device=local) as lat:
        t = lat.t
        for j in range(NPTS):
            p = PC + (j - NPTS // 2) * 0.25 * Lsz ** -0.75
            ks = int(p * t)                                  # tsites = ps * t, truncated (Tri/site.f:164)
            if not mine:
                continue
            # one device-resident batch per point; this rank's realizations are the streams rank * STRIDE + i
            _, st = lat.batch(P.SITE, len(mine), args.seed + 1000 * a + j, rank * STREAM_STRIDE, ks, 0, 0)
            table[a, j] = [st[f] for f in FIELDS]
if dist is not None:
    import torch
    tt = torch.from_numpy(table).cuda()
    dist.all_reduce(tt, op=dist.ReduceOp.SUM)                # integer sums: identical for any number of GPUs
    table = tt.cpu().numpy()
elapsed = time.perf_counter() - t_all
if rank == 0:
    sites = sum(int(table[a, :, 0].sum()) * Lsz * Lsz for a, Lsz in enumerate(Ls))
    print(json.dumps({"config": "C4 triangular site FSS sweep", "L": Ls, "points": NPTS, "nreal": args.nreal, "gpus": world,
                      "seconds": elapsed, "gsites_per_s": sites / elapsed / 1e9}))
    for a, Lsz in enumerate(Ls):
        for j in range(NPTS):
            n, sp, mx, ncl, failed = (int(v) for v in table[a, j])
            p = PC + (j - NPTS // 2) * 0.25 * Lsz ** -0.75
            print("L=%6d  p=%.6f  Pi=%.4f  Pmax=%.5f  clusters/site=%.5f  n=%d%s"
                  % (Lsz, p, sp / max(n, 1), mx / max(n, 1) / (Lsz * Lsz), ncl / max(n, 1) / (Lsz * Lsz), n,
                     "  FAILED=%d" % failed if failed else ""))
if dist is not None:
    dist.destroy_process_group()
