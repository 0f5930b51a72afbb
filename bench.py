#!/usr/bin/env python
"""bench.py -- headline benchmark of the percolation-realization hot path on B200.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on, as SURVEY 8(d) C3 specifies it and as the
reference's p-sweep driver runs it, Sq/bond_cond.f:84-97,208-485): square-lattice mixed site/bond percolation, L = 4096,
sites fixed at ps = 0.80, bonds added in one order.  One "step" = one realization =
    occupancy ranks (K1 Philox generator) -> first-spanning bond count k* (perc_first_span: bisection, K1-K5 per probe)
    -> for each of the 9 sweep points pb* + 0.005 j, j = 0 .. 8:  labeling + sizes + spanning (K2-K5) and the Kirchhoff
       conductance of the spanning cluster (K6-K8, fp64, tol 1e-10; deflated one-pass Jacobi-PCG kernel).
Metric: conductance realizations / s.  `extra` keeps round 1's single-point line (one solve at pb = 0.70) and short
legs for the other BASELINE configurations (C1, C2, C4; C5 in slabs when several GPUs run).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, through the C-ABI)
  python bench.py --impl reference ...                            CPU arm: the oracle port of the
        reference's algorithm on ALL host cores (the Fortran itself cannot be built here: no
        Fortran compiler in the image), each step a bounded, really timed sample.

For N > 1 (torchrun, one rank per GPU) realizations are sharded across ranks with no data-path
collective; one NCCL all-reduce merges the statistics at the end (weak scaling).  After that the same ranks
run ONE lattice decomposed into row slabs (multi-GPU mode 2) as `extra.slab`.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "1")     # CPU arm: one realization per core, no nested OpenMP teams in the oracle

METRIC = "conductance realizations/sec at L=4096 near p_c; CCL Gsites/s; % HBM peak"
UNIT = "realizations/s"
SEED = 20240611
NPTS = 9                      # sweep points per realization (j = 0 .. 8)
DPB = 5.0e-3                  # Sq/bond_cond.f:89-91
ITERS_FILE = os.path.join(ROOT, "profiles", "bench_iters.json")


def ncu_traffic_bytes(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json, written by tools/ncu_traffic.py)"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return float(json.load(f)[kernel_key]["dram_bytes_per_launch"])
    except Exception:
        return None


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def sweep_fill(kstar, nb, j):
    """bond count of sweep point j: pb accumulated in fp64 by repeated + 5e-3 from pb* = k*/nb and truncated
    (the rule of Sq/bond_cond.f:89-94); point 0 is the first-spanning fill itself"""
    pb = float(kstar) / float(nb)
    for _ in range(j):
        pb = pb + DPB
    return min(nb, max(int(kstar), int(pb * nb)))


# ------------------------------------------------------------------------------------------------
# CPU arm / cpu_baseline: the oracle port timed on a bounded sample, one realization per host core
# ------------------------------------------------------------------------------------------------
def host_occupancy(rng, t, nb, ks, kb):
    socc = np.zeros(t, np.uint8)
    socc[np.argpartition(rng.random(t), ks)[:ks]] = 1
    bocc = np.zeros(nb, np.uint8)
    bocc[np.argpartition(rng.random(nb), kb)[:kb]] = 1
    return socc, bocc


def plain_iters_table(npts, Lsz=4096, fallback_mean=None):
    """iterations the reference's solver (plain Jacobi-PCG = linbcg on a symmetric matrix) needs at each sweep point of
    the bench's realizations: measured once on the GPU with the plain one-pass kernel (`bench.py --measure-plain`,
    committed as profiles/bench_iters.json) -- the CPU arm cannot run 10^5 iterations of a 2^24-site system per
    point inside a bench run, so its solve time is (measured seconds per iteration) x (these counts)."""
    try:
        with open(ITERS_FILE) as f:
            d = json.load(f)
        it = [float(v) for v in d["plain_iters_per_point"]][:npts]
        if len(it) == npts and int(d.get("L", 0)) == int(Lsz):
            return it, "profiles/bench_iters.json (plain Jacobi-PCG counts of the bench realizations, GPU-measured)"
    except Exception:
        pass
    v = float(fallback_mean) if fallback_mean else 66000.0 * Lsz / 4096.0          # (Jacobi-PCG iterations grow about linearly with L)
    return [v] * npts, "documented constant (about 66 000 x L / 4096 at pb = 0.70, DESIGN.md)"


def cpu_sample(Lsz, socc, bocc, nthreads, cg_iters):
    """every thread runs one realization's bounded sample concurrently (one realization per core): union-find labeling
    + spanning + `cg_iters` Jacobi-PCG iterations, all really timed"""
    from oracle import pyoracle as O
    m = n = Lsz
    b1, b2 = O.bondlist(O.SQUARE, m, n, 0)
    out = [None] * nthreads

    def work(k):
        t0 = time.perf_counter()
        ws, wb, wsz, ncl, wmax = O.label_uf(O.MIXED, O.SQUARE, m, n, 0, b1, b2, site_occ=socc, bond_occ=bocc)
        ids = O.spanning(O.MIXED, m, n, b1, b2, ws, wb)
        t_label = time.perf_counter() - t0
        cid = int(ids[0]) if len(ids) else int(np.argmax(wsz))
        w = O.weights(O.MIXED, b1, b2, ws, wb, cid)
        t_cg = O.cg_time_iters(m, n, b1, b2, w, cg_iters)
        out[k] = (t_label, t_cg / cg_iters)

    ths = [threading.Thread(target=work, args=(k,)) for k in range(nthreads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    wall = time.perf_counter() - t0
    return {"t_label_s": float(np.mean([o[0] for o in out])), "t_iter_s": float(np.mean([o[1] for o in out])),
            "sample_wall_s": wall}


def cpu_extrapolate(r, nthreads, iters_per_point):
    """realizations/s of the whole C3 step on `nthreads` cores from the timed sample: the first-spanning search counted
    as ONE labeling pass (what an incremental Newman-Ziff fill costs), then per sweep point one labeling and the
    reference solver's iteration count at that point"""
    per_real = (1 + len(iters_per_point)) * r["t_label_s"] + float(np.sum(iters_per_point)) * r["t_iter_s"]
    return nthreads / per_real, per_real


def cpu_c1_literal(nthreads, per_thread=2):
    """BASELINE configs[0] in the LITERAL flavour (the reference's own algorithm with its O(t) relabel scans, Sq/site.f:162-289,
    spanning scan :309-344, linbcg with tol 1e-8 / itmax 2500 :545): square site L = 100, p = 0.60, `per_thread`
    realizations on each host core, seeds from the reference's seed table (Sq/site_perc.f:69-75)"""
    from oracle import pyoracle as O
    m = n = 100
    t = m * n
    b1, b2 = O.bondlist(O.SQUARE, m, n, 0)
    seeds = O.seed_table(58302, nthreads * per_thread, 1000000)
    done = [0] * nthreads

    def work(k):
        for j in range(per_thread):
            order = O.shuffle_sites(int(seeds[k * per_thread + j]), t)
            s, c, res = O.site_literal(O.SQUARE, m, n, 0, order, O.fill_count(0.60, t))
            perccln = int(res.get("perccln", 0))
            if perccln > 0:
                w = O.weights(O.SITE, b1, b2, s, np.zeros(len(b1), np.int32), perccln)
                O.conduct_literal(m, n, b1, b2, w, tol=1e-8, itmax=2500)
            done[k] += 1

    ths = [threading.Thread(target=work, args=(k,)) for k in range(nthreads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    return {"realizations_per_s": sum(done) / dt, "realizations": sum(done), "seconds": dt, "cores": nthreads,
            "flavour": "literal (reference algorithm incl. relabel scans and linbcg tol 1e-8 / itmax 2500), C port, gcc -O2"}


def host_threads(args):
    """one realization per host core: every core this process may run on, bounded by host memory (a realization of the
    CPU port holds about 150 B per site: labels, bond weights, the sparse matrix and five vectors)"""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0)) or n
    except Exception:
        pass
    if args.cpu_threads > 0:
        n = min(n, args.cpu_threads)
    try:
        import psutil
        per_thread = 150.0 * args.L * args.L + 64e6
        n = min(n, max(1, int(0.6 * psutil.virtual_memory().available / per_thread)))
    except Exception:
        pass
    return max(1, n)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    Lsz = args.L
    t, nb = Lsz * Lsz, 2 * Lsz * Lsz - 2 * Lsz
    ks, kb = int(args.ps * t), int(args.pb * nb)
    nthreads = host_threads(args)
    iters, src = plain_iters_table(args.npts, Lsz)
    rng = np.random.default_rng(SEED)
    vals, t0 = [], None
    for step in range(args.warmup + args.steps):
        socc, bocc = host_occupancy(rng, t, nb, ks, kb)
        if step == args.warmup:
            t0 = time.perf_counter()
        r = cpu_sample(Lsz, socc, bocc, nthreads, args.cpu_cg_iters)
        if step >= args.warmup:
            vals.append(r)
    wall = time.perf_counter() - t0
    ext = [cpu_extrapolate(v, nthreads, iters) for v in vals]
    value = float(np.mean([e[0] for e in ext]))
    sample = ("MEASURED per step, on each of %d host threads concurrently: union-find labeling + spanning of one L=%d "
              "realization at pb=%.2f (%.2f s) + %d Jacobi-PCG iterations (%.3f s/iter).  `value` is EXTRAPOLATED from "
              "those two timings to the whole step (first-spanning search counted as one labeling pass + %d sweep points, "
              "each one labeling + the reference solver's iteration count at that point: %s; source: %s): %.0f s per "
              "realization and core.  C port of the reference algorithm, gcc -O2 (no Fortran compiler in the image)"
              % (nthreads, Lsz, args.pb, vals[-1]["t_label_s"], args.cpu_cg_iters, vals[-1]["t_iter_s"], args.npts,
                 [int(v) for v in iters], src, ext[-1][1]))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        # the time the timed region REALLY took per step (the bounded sample), not the extrapolated step
        "ms_per_step": 1e3 * wall / max(args.steps, 1),
        "value_is_extrapolated": True, "extrapolated_s_per_realization_per_core": ext[-1][1],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "bench_wall_s": wall,
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": "C3 square mixed site/bond L=%d, ps=%.2f: first-spanning search + %d-point p-sweep (pb* + 0.005 j), "
                        "labeling+spanning+Kirchhoff PCG at every point" % (args.L, args.ps, args.npts),
            "lattice": "square", "L": args.L, "ps": args.ps, "sweep_points": args.npts, "dpb": DPB, "tol": args.tol,
            "occupancy": "Philox-4x32-10 exact-count generator, one stream per realization (ranks: one order per realization)",
            "l2": "per-iteration working set %.2f GB > 126 MB L2 (no flush needed)" % (4 * 8 * args.L * args.L / 1e9),
            "conduct": "perc_conduct (voltages kept)" if getattr(args, "voltages", False) else
                       "perc_conduct_g (Gtop/Gbot of the p-sweep drivers; interior voltages not formed)",
            "parallelism": "realizations sharded over %d GPU(s), one final NCCL all-reduce of statistics" % args.gpus}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def e2e_host_inputs(rng, t, nb, b1, b2, count, alloc):
    """The reference drivers' shuffled fill orders (site order(t), Sq/sitebond.f:131-147; bond order as the two
    columns of border(nb,2), :165-176) in host buffers made by `alloc(n)` (pinned int32 tensors in the bench),
    prepared BEFORE the timed region: like the CPU arm, a timed step starts from its inputs."""
    out = []
    for _ in range(count):
        hs, hb = alloc(t), alloc(2 * nb)
        hs.numpy()[:] = rng.permutation(t).astype(np.int32) + 1
        bp = rng.permutation(nb)
        hb.numpy()[:nb] = b1[bp]
        hb.numpy()[nb:] = b2[bp]
        out.append((hs, hb))
    return out


def extra_legs(P, local, peak):
    """short runs of the other BASELINE configurations (rank 0, one GPU), so that each has a driver-run number"""
    out = {}
    try:
        # C1: square site L=100 p=0.60, 1000 realizations: labeling + spanning + conductance at the reference's tol / itmax
        with P.Lattice(P.SQUARE, 100, 100, 0, device=local) as L:
            ks = int(0.60 * L.t)
            L.batch_conduct(P.SITE, 8, 1, 0, ks, 0, tol=1e-8, itmax=2500)
            dt = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                G, iters, st = L.batch_conduct(P.SITE, 1000, SEED, 0, ks, 0, tol=1e-8, itmax=2500)
                dt = min(dt, time.perf_counter() - t0)
            sp = iters >= 0
            out["C1 square site L=100 p=0.60 x1000: labeling+spanning+conductance (tol 1e-8, itmax 2500; perc_batch_conduct)"] = {
                "realizations_per_s": 1000 / dt, "seconds": dt, "spanning_fraction": float(sp.mean()),
                "mean_iters": float(iters[sp].mean()) if sp.any() else 0.0, "failed_selections": st["failed"]}
            L.batch(P.SITE, 8, 1, 0, ks, 0, 64)
            t0 = time.perf_counter()
            hist, st = L.batch(P.SITE, 1000, SEED, 0, ks, 0, 64)
            dt = time.perf_counter() - t0
            out["C1 labeling+spanning only (perc_batch)"] = {"realizations_per_s": 1000 / dt, "seconds": dt,
                                                              "spanning_fraction": st["spanning"] / 1000.0}
    except Exception as e:                                                     # an extra leg must not cost the headline
        out["C1 error"] = repr(e)
    try:
        # C2: triangular bond L=1024 p=0.35, 4096 batched realizations: labeling + cluster-size histogram
        with P.Lattice(P.TRIANGULAR, 1024, 1024, 0, device=local) as L:
            kb = int(0.35 * L.nb)
            L.batch(P.BOND, 8, 1, 0, 0, kb, 64)
            t0 = time.perf_counter()
            hist, st = L.batch(P.BOND, 4096, SEED, 0, 0, kb, 64)
            dt = time.perf_counter() - t0
            out["C2 triangular bond L=1024 p=0.35 x4096: labeling + size histogram (perc_batch)"] = {
                "realizations_per_s": 4096 / dt, "seconds": dt, "gsites_per_s": 4096 * L.t / dt / 1e9,
                "hist_total_clusters": int(hist.sum()), "failed_selections": st["failed"]}
    except Exception as e:
        out["C2 error"] = repr(e)
    try:
        # C4: triangular site L=16384 at p_c = 0.5: one labeling (mask build included), then a short finite-size-scaling
        # sweep at L = 1024 (17 p-points, 8 realizations each) -- tools/fss_sweep.py runs the full table
        with P.Lattice(P.TRIANGULAR, 16384, 16384, 0, device=local) as L:
            acc = np.zeros(8)
            for r in range(3):
                L.generate(SEED, r, int(0.5 * L.t), -1)
                L.label(P.SITE)
                if r:
                    acc += L.phase_ms()
            acc /= 2
            tot = float(acc[0] + acc[1] + acc[2] + acc[3] + acc[4])
            ccl = float(acc[1] + acc[2] + acc[3])
            out["C4 triangular site L=16384 p=0.5: one labeling"] = {
                "ms_mask_ccl_span": tot, "ms_ccl": ccl, "gsites_per_s": L.t / (tot * 1e-3) / 1e9,
                "ccl_frac_of_hbm_peak_at_5B_per_site": 5.0 * L.t / (ccl * 1e-3) / 1e9 / peak, "nspan": L.summary()["nspan"]}
        with P.Lattice(P.TRIANGULAR, 1024, 1024, 0, device=local) as L:
            pts, t0 = [], time.perf_counter()
            for j in range(-8, 9):
                p = 0.5 + j * 0.25 * 1024 ** (-0.75)
                hist, st = L.batch(P.SITE, 8, SEED, 100 * (j + 8), int(p * L.t), 0, 0)
                pts.append({"p": p, "spanning_fraction": st["spanning"] / 8.0, "mean_maxcs": st["sum_maxcs"] / 8.0})
            out["C4 FSS sweep L=1024: 17 p-points x 8 realizations (perc_batch)"] = {
                "seconds": time.perf_counter() - t0, "points": pts}
    except Exception as e:
        out["C4 error"] = repr(e)
    return out


def slab_leg(P, torch, dist, rank, world, local, args):
    """multi-GPU mode 2: ONE square site lattice at p_c decomposed into row slabs over the ranks
    (C5: L = 65536 on 8 GPUs; 32768 on 4, 16384 on 2): labeling + stitch, then a fixed number of distributed PCG
    iterations; where the lattice fits one GPU, rank 0 repeats the labeling undecomposed and compares the counts"""
    Lsz = args.slab_L if args.slab_L > 0 else {2: 16384, 4: 32768, 8: 65536}.get(world, 8192 * world)
    uid = torch.from_numpy(P.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)).cuda()
    dist.broadcast(uid, 0)
    S = P.SlabLattice(P.SQUARE, Lsz, Lsz, 0, local, world, rank, unique_id=uid.cpu().numpy())
    t = Lsz * Lsz
    S.generate(SEED, 0, int(0.592746 * t), -1)
    times = []
    for _ in range(3):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        S.label(P.SITE)
        torch.cuda.synchronize(); dist.barrier()
        times.append(time.perf_counter() - t0)
    sm = S.summary()
    out = {"L": Lsz, "ranks": world, "p": 0.592746, "label_stitch_wall_ms": 1e3 * min(times),
           "gsites_per_s": t / min(times) / 1e9, "ncl": sm["ncl"], "maxcs": sm["maxcs"], "nspan": sm["nspan"]}
    if sm["nspan"] and args.slab_iters > 0:
        S.conduct(0, tol=1e-30, itmax=15, voltages=False)          # warm-up: NCCL sets its channels up on first use
        dist.barrier(); torch.cuda.synchronize()
        res = S.conduct(0, tol=1e-30, itmax=args.slab_iters - 1, voltages=False)
        torch.cuda.synchronize(); dist.barrier()
        ph = S.phase_ms()
        ms_it = torch.tensor([float(ph[5]) / max(res["iter"], 1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(ms_it, op=dist.ReduceOp.MAX)
        own = (S.yb - S.ya) * Lsz
        out.update({"pcg_iters": res["iter"], "pcg_ms_per_iter": float(ms_it.item()),
                    "pcg_gbs_per_gpu": 50.0 * own / (float(ms_it.item()) * 1e-3) / 1e9,
                    "pcg_bytes_per_site_iter": 50, "pcg_form": "two-kernel form, 2 all-reduces + 1 halo exchange per iteration"})
    S.close()
    if rank == 0 and t >= 2 ** 30:
        out["equals_single_gpu"] = "not checked: 2^%d sites do not fit a single-GPU handle (int32 labels)" % (t.bit_length() - 1)
    if rank == 0 and t < 2 ** 30:
        # the same lattice undecomposed on one GPU: counts must be bit-identical
        try:
            with P.Lattice(P.SQUARE, Lsz, Lsz, 0, device=local) as L1:
                L1.generate(SEED, 0, int(0.592746 * t), -1)
                L1.label(P.SITE)
                s1 = L1.summary()
                out["equals_single_gpu"] = bool(s1["ncl"] == sm["ncl"] and s1["maxcs"] == sm["maxcs"] and s1["nspan"] == sm["nspan"])
        except Exception as e:
            out["equals_single_gpu"] = "not checked: " + repr(e)
    return out


def run_ours(args):
    import torch
    import percolation_b200 as P
    from percolation_b200.shard import Stats, stream_id
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    Lsz = args.L
    L = P.Lattice(P.SQUARE, Lsz, Lsz, 0, device=local)
    t, nb = L.t, L.nb
    ks = int(args.ps * t)
    lib = P.load()
    sptr = C.c_uint64(0)
    lib.perc_stream(C.byref(L._h), C.byref(sptr))
    ext = torch.cuda.ExternalStream(sptr.value, device=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    solver_mode = {"auto": 0, "classic": 1, "plain": 2}[args.solver]
    L.set_solver(solver_mode)
    stats = {"G": [], "iters": [], "kernel_ms": [], "upd_ms": [], "ccl_ms": [], "mask_ms": [], "pcg_ms": [], "fused": [],
             "iters_pt": [[] for _ in range(args.npts)], "pb_star": [], "search_ms": [], "incremental": [], "inc_ms": []}

    def sweep_step(i, record):
        """one realization as the p-sweep driver runs it: ranks -> k* -> NPTS x (label, conduct)"""
        L.generate(SEED, stream_id(rank, i), ks, nb)
        t0 = time.perf_counter()
        fs = L.first_span(P.MIXED, P.BOND)
        t_search = time.perf_counter() - t0
        if record:
            stats["pb_star"].append(fs["kstar"] / nb)
            stats["search_ms"].append(1e3 * t_search)
        for j in range(args.npts):
            L.set_fill(kb=sweep_fill(fs["kstar"], nb, j))
            # (the first point follows the search's last probe, the others the previous point: only the added bonds are united)
            inc = L.label_incremental(P.MIXED)
            if record:
                stats["incremental"].append(bool(inc))
            ph = L.phase_ms().copy()
            r = L.conduct(0, tol=args.tol, itmax=args.itmax, voltages=args.voltages)
            ph2 = L.phase_ms()
            if record:
                stats["G"].append(0.5 * (r["Gtop"] + r["Gbot"]))
                stats["iters"].append(r["iter"])
                stats["iters_pt"][j].append(r["iter"])
                (stats["inc_ms"] if inc else stats["ccl_ms"]).append(float(ph[1] + ph[2] + ph[3]))
                stats["mask_ms"].append(float(ph[0]))
                stats["kernel_ms"].append(float(ph2[6]))
                stats["upd_ms"].append(float(ph2[7]))
                stats["pcg_ms"].append(float(ph2[5]))
                stats["fused"].append(L.solver_used())

    for i in range(args.warmup):
        sweep_step(i, False)
    clocks = ClockSampler(local)
    launches0 = L.launch_count()
    barrier()
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for i in range(args.steps):
        sweep_step(args.warmup + i, True)
    e1.record(ext)
    barrier()
    clk = clocks.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    launches = L.launch_count() - launches0
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())

    # ---- e2e: the same step through the reference-shaped C-ABI calls with HOST buffers: the driver's shuffled site
    # and bond orders uploaded from pinned host memory (Sq/sitebond.f:131-176), k* and the sweep's G values read back
    # (what Sq/bond_cond.f:481-482 writes), plus labels and sizes of the last sweep point (Sq/sitebond.f:469-477)
    rng = np.random.default_rng(SEED + rank)
    b1, b2 = P.geom_bondlist(P.SQUARE, Lsz, Lsz, 0)
    pin = lambda n_, dt: torch.empty(n_, dtype=dt, pin_memory=True)
    ip = lambda ten: C.cast(ten.data_ptr(), C.POINTER(C.c_int32))
    i32 = lambda v: C.byref(C.c_int32(int(v)))
    f64 = lambda v: C.byref(C.c_double(float(v)))
    e2e_steps = args.e2e_steps if args.e2e_steps >= 0 else args.steps
    e2e_val, e2e_G = None, []
    h2d = d2h = 0
    if e2e_steps > 0:
        h_s, h_b3, h_c = pin(t, torch.int32), pin(nb, torch.int32), pin(t, torch.int32)
        inputs = e2e_host_inputs(rng, t, nb, b1, b2, min(e2e_steps, 2), lambda n_: pin(n_, torch.int32))

        def ck(rc):
            if rc:
                raise RuntimeError("C-ABI call failed: %d" % rc)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            h_sorder, h_border = inputs[i % len(inputs)]
            ck(lib.perc_set_site_order(C.byref(L._h), ip(h_sorder)))
            ck(lib.perc_set_bond_order(C.byref(L._h), ip(h_border)))
            ck(lib.perc_set_fill(C.byref(L._h), i32(ks), i32(nb)))
            kstar, f_, mx, pcs = C.c_int32(0), C.c_float(0), C.c_int32(0), C.c_int32(0)
            ck(lib.perc_first_span(C.byref(L._h), i32(P.MIXED), i32(P.BOND), C.byref(kstar), C.byref(f_), C.byref(mx), C.byref(pcs)))
            h2d = 4 * t + 8 * nb
            d2h = 16
            for j in range(args.npts):
                ck(lib.perc_set_fill(C.byref(L._h), i32(ks), i32(sweep_fill(kstar.value, nb, j))))
                ck(lib.perc_label_incremental(C.byref(L._h), i32(P.MIXED), None))
                Gt, Gb, er, it = C.c_double(0), C.c_double(0), C.c_double(0), C.c_int32(0)
                conduct = lib.perc_conduct if args.voltages else lib.perc_conduct_g
                ck(conduct(C.byref(L._h), i32(0), f64(1.0), f64(1.0), f64(1e-12), f64(args.tol),
                           i32(args.itmax), f64(1e-10), C.byref(Gt), C.byref(Gb), C.byref(it), C.byref(er)))
                e2e_G.append(0.5 * (Gt.value + Gb.value))
                d2h += 28
            ck(lib.perc_get_site_labels(C.byref(L._h), ip(h_s)))
            ck(lib.perc_get_bond_labels(C.byref(L._h), ip(h_b3)))
            ck(lib.perc_get_sizes(C.byref(L._h), ip(h_c)))
            d2h += 8 * t + 4 * nb
        torch.cuda.synchronize()
        el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        e2e_val = world * e2e_steps / float(el.item())

    # ---- statistics: one NCCL all-reduce of the statistics block (multi-GPU mode 1)
    merged = Stats()
    for G_, it_ in zip(stats["G"], stats["iters"]):
        merged.add(G=G_, iters=it_, spans=True)
    if dist is not None:
        # the library's communicator (NCCL, bootstrapped with an id broadcast by the host program) and its
        # single reduction of the statistics block -- perc_comm_init_rank / perc_allreduce_stats
        uid = torch.from_numpy(P.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)).cuda()
        dist.broadcast(uid, 0)
        L.comm_init_rank(world, rank, uid.cpu().numpy())
        merged.allreduce_native(L)
    sd = merged.asdict()
    st = [sd["sum_G"], sd["sum_G2"], sd["count"], sd["iters"]]

    # ---- round 1's single-point line (one solve at pb = 0.70 per realization), kept for continuity
    single = None
    if rank == 0 and args.single_steps > 0:
        kb1 = int(args.pb * nb)
        res, ev = [], [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for i in range(args.single_steps + 1):
            if i == 1:
                ev[0].record(ext)
            L.generate(SEED, stream_id(0, i), ks, kb1)
            L.label(P.MIXED)
            ph = L.phase_ms()
            stats["ccl_ms"].append(float(ph[1] + ph[2] + ph[3]))          # (the from-scratch labeling: extra.ccl)
            res.append(L.conduct(0, tol=args.tol, itmax=args.itmax, voltages=False))
        ev[1].record(ext)
        torch.cuda.synchronize()
        dt = ev[0].elapsed_time(ev[1]) * 1e-3
        single = {"workload": "one realization = one solve at pb=%.2f (round 1's bench step)" % args.pb,
                  "realizations_per_s": args.single_steps / dt, "mean_pcg_iterations": float(np.mean([r["iter"] for r in res[1:]])),
                  "G": [0.5 * (r["Gtop"] + r["Gbot"]) for r in res[1:]]}

    # ---- the same sweep WITHOUT deflation on the same lattice (plain one-pass kernel, capped): its launch time / roofline
    plain = None
    if rank == 0 and args.plain_iters > 0:
        kb1 = int(args.pb * nb)
        L.generate(SEED, stream_id(0, 1), ks, kb1)
        L.label(P.MIXED)
        L.set_solver(2)
        rp = L.conduct(0, tol=1e-30, itmax=args.plain_iters, voltages=False)
        php = L.phase_ms()
        L.set_solver(solver_mode)
        plain = {"kernel": "pcg_fused_kernel<FtCfgA3> (perc_set_solver 2: the same one-pass sweep without deflation, linbcg's iterates)",
                 "iterations_timed": rp["iter"], "avg_launch_ms": float(php[6])}

    # ---- accuracy of the bench's tolerance: the first timed realization's sweep point at a 1e-13 solve
    acc = None
    if rank == 0 and args.check_tol:
        kb1 = int(args.pb * nb)
        L.generate(SEED, stream_id(0, 1), ks, kb1)
        L.label(P.MIXED)
        ra = L.conduct(0, tol=args.tol, itmax=args.itmax, voltages=False)
        rb = L.conduct(0, tol=1e-13, itmax=args.itmax, voltages=False)
        Ga, Gb_ = 0.5 * (ra["Gtop"] + ra["Gbot"]), 0.5 * (rb["Gtop"] + rb["Gbot"])
        acc = {"tol": args.tol, "rel_dG_vs_tol_1e-13": abs(Ga - Gb_) / abs(Gb_), "rel_Gtop_minus_Gbot": abs(ra["Gtop"] - ra["Gbot"]) / abs(Gb_),
               "iters": [ra["iter"], rb["iter"]]}

    # ---- plain Jacobi-PCG iteration counts of the bench realizations (input of the CPU arm's extrapolation)
    if rank == 0 and args.measure_plain:
        L.set_solver(2)
        per_pt = [[] for _ in range(args.npts)]
        for i in range(args.steps):
            L.generate(SEED, stream_id(0, args.warmup + i), ks, nb)
            fs = L.first_span(P.MIXED, P.BOND)
            for j in range(args.npts):
                L.set_fill(kb=sweep_fill(fs["kstar"], nb, j))
                L.label(P.MIXED)
                per_pt[j].append(L.conduct(0, tol=args.tol, itmax=4000000, voltages=False)["iter"])
        L.set_solver(solver_mode)
        os.makedirs(os.path.dirname(ITERS_FILE), exist_ok=True)
        with open(ITERS_FILE, "w") as f:
            json.dump({"plain_iters_per_point": [float(np.mean(v)) for v in per_pt], "L": Lsz, "ps": args.ps, "tol": args.tol,
                       "realizations": args.steps, "first_stream": args.warmup,
                       "deflated_iters_per_point": [float(np.mean(v)) for v in stats["iters_pt"]],
                       "note": "iterations of the plain one-pass Jacobi-PCG (perc_set_solver 2: linbcg's iterates) at the sweep points"}, f)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        interior = t - 2 * Lsz
        fused = stats["fused"][0] if stats["fused"] and all(f == stats["fused"][0] for f in stats["fused"]) else -1
        # average duration of an iteration-kernel launch over the timed region, weighted by launches
        it_arr, km = np.array(stats["iters"], float), np.array(stats["kernel_ms"], float)
        kernel_ms = float((it_arr * km).sum() / max(it_arr.sum(), 1))
        upd_ms = float(np.mean(stats["upd_ms"]))
        if not stats["ccl_ms"]:                                  # (no from-scratch labeling was timed: take one now)
            L.generate(SEED, stream_id(0, 0), ks, int(args.pb * nb))
            L.label(P.MIXED)
            ph = L.phase_ms()
            stats["ccl_ms"].append(float(ph[1] + ph[2] + ph[3]))
        ccl_ms, mask_ms = float(np.mean(stats["ccl_ms"])), float(np.mean(stats["mask_ms"]))
        if fused >= 1:
            # one-pass iteration kernel: u 8 + s 8 + conduct byte 1 read; u 8 + s 8 written
            spmv_bytes = 33.0 * interior
            roof_key = "pcg_fused_kernel<FtCfgD>" if fused == 2 else "pcg_fused_kernel"
            roof_name = ("pcg_fused_kernel (one Jacobi-PCG iteration in one persistent TMA-fed pass over u = D^-1 r and s = A p: "
                         "s = A u + beta s, u -= alpha D^-1 s, sums r.u, r.r and the bond energy u.A u; p and x only on the read-out rows"
                         + ("; deflated: block-constant coarse space, the coarse solve inside the same launch)" if fused == 2 else ")"))
        else:
            spmv_bytes = 25.0 * interior          # r 8 + p_old 8 + conduct byte 1 read; p 8 written (q = A p is never stored)
            roof_key = "pcg_pipe_kernel<0>"
            roof_name = "pcg_pipe_kernel<0> (persistent TMA tile pipeline: p = r/d + bk p, p.Ap as bond energies; q = A p is never stored)"
        upd_bytes = (41.0 if args.voltages else 25.0) * interior   # p 8 + r 8 + byte read, r 8 written (+ x 8 + 8)
        ach = spmv_bytes / (max(kernel_ms, 1e-9) * 1e-3) / 1e9
        value = world * args.steps / (ms_total * 1e-3)
        mean_iters = st[3] / max(st[2], 1)
        nsolve = len(stats["iters"])
        ccl_all = ccl_ms + mask_ms
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "clocks": clk,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": roof_name,
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": ncu_traffic_bytes(roof_key) if Lsz == 4096 else None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": spmv_bytes,
                         "avg_launch_ms": kernel_ms,
                         "avg_launch_ms_source": "CUDA events on the library's stream around whole chunks of launches of the timed solves / launches in them"},
            "extra": {
                "pcg_solver": {2: "deflated one-pass (33 B per site and iteration)", 1: "one-pass (33 B per site and iteration)",
                               0: "two-kernel (50 B per site and iteration)"}.get(fused, "mixed"),
                "conductance_solves_per_s": world * nsolve / (ms_total * 1e-3),
                "ccl": {"gsites_per_s": t / (ccl_ms * 1e-3) / 1e9, "ms": ccl_ms,
                        "achieved_gbs": 5.0 * t / (ccl_ms * 1e-3) / 1e9, "frac": 5.0 * t / (ccl_ms * 1e-3) / 1e9 / peak,
                        "with_mask_build": {"ms": ccl_all, "gsites_per_s": t / (ccl_all * 1e-3) / 1e9,
                                            "frac": 5.0 * t / (ccl_all * 1e-3) / 1e9 / peak}},
                "mean_pcg_iterations": mean_iters, "mean_pcg_iterations_per_sweep_point": [float(np.mean(v)) for v in stats["iters_pt"]],
                "relabeling": {"what": "perc_label_incremental at the sweep points: only the added bonds are united; extra.ccl is the from-scratch pass",
                               "incremental_passes": int(np.sum(stats["incremental"])), "of": len(stats["incremental"]),
                               "ms": float(np.mean(stats["inc_ms"])) if stats["inc_ms"] else None},
                "mean_pb_star": float(np.mean(stats["pb_star"])), "first_span_search_ms": float(np.mean(stats["search_ms"])),
                "mean_G": st[0] / max(st[2], 1),
                "all_solves_converged": bool(all(it <= args.itmax for it in stats["iters"])),
                "pcg_ms_per_iteration": float(np.sum(stats["pcg_ms"])) / max(float(np.sum(stats["iters"])), 1),
                "solves": int(st[2]),
            },
        }
        if single:
            line["extra"]["single_point"] = single
        if acc:
            line["extra"]["tolerance_check"] = acc
        if plain:
            pa = 33.0 * interior / (max(plain["avg_launch_ms"], 1e-9) * 1e-3) / 1e9
            plain.update({"achieved_gbs": pa, "frac": pa / peak, "algorithmic_bytes_per_launch": 33.0 * interior,
                          "traffic": ncu_traffic_bytes("pcg_fused_kernel") if Lsz == 4096 else None})
            line["extra"]["plain_one_pass_kernel"] = plain
        if fused == 0 and upd_ms > 0:
            line["extra"]["pcg_pipe_kernel<1> (residual update, A p recomputed)"] = {
                "achieved_gbs": upd_bytes / (upd_ms * 1e-3) / 1e9, "frac": upd_bytes / (upd_ms * 1e-3) / 1e9 / peak,
                "avg_launch_ms": upd_ms}
        if world == 1 and not args.no_extra_legs:
            line["extra"]["configs"] = extra_legs(P, local, peak)
        if world == 1 and not args.no_cpu_baseline:
            kb1 = int(args.pb * nb)
            L.generate(SEED, stream_id(0, args.warmup), ks, kb1)
            socc, bocc = L.get_occupancy()
            nthreads = host_threads(args)
            iters_tab, src = plain_iters_table(args.npts, Lsz)
            r = cpu_sample(Lsz, socc, bocc, nthreads, args.cpu_cg_iters)
            v, per_real = cpu_extrapolate(r, nthreads, iters_tab)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": nthreads, "kind": "port",
                "sample": ("MEASURED on each of %d host threads concurrently: union-find labeling + spanning of the first timed "
                           "realization at pb=%.2f (%.2f s) + %d Jacobi-PCG iterations (%.3f s/iter); EXTRAPOLATED to the whole "
                           "step (search as one labeling pass + %d sweep points x (labeling + the reference solver's iterations "
                           "at that point %s [%s])) = %.0f s per realization and core; C port of the reference algorithm, "
                           "gcc -O2 (no Fortran compiler in the image)"
                           % (nthreads, args.pb, r["t_label_s"], args.cpu_cg_iters, r["t_iter_s"], args.npts,
                              [int(x) for x in iters_tab], src, per_real)),
                "sample_wall_s": r["sample_wall_s"]}
            try:
                line["cpu_baseline"]["c1_literal"] = cpu_c1_literal(nthreads)
            except Exception as e:
                line["cpu_baseline"]["c1_literal"] = "failed: " + repr(e)
    L.close()
    slab = None
    if dist is not None and not args.no_slab:
        try:
            slab = slab_leg(P, torch, dist, rank, world, local, args)
        except Exception as e:
            slab = {"error": repr(e)}
    if rank == 0:
        if slab is not None:
            line["extra"]["slab"] = slab
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--L", type=int, default=4096)
    ap.add_argument("--ps", type=float, default=0.80)
    ap.add_argument("--pb", type=float, default=0.70, help="bond fraction of the single-point leg and of the CPU sample")
    ap.add_argument("--npts", type=int, default=NPTS, help="sweep points per realization")
    ap.add_argument("--tol", type=float, default=1e-10)
    ap.add_argument("--itmax", type=int, default=2000000,
                    help="iteration cap of a solve (the cap bounds the run time; all_solves_converged reports whether it was hit)")
    ap.add_argument("--e2e-steps", type=int, default=-1)
    ap.add_argument("--plain-iters", type=int, default=3000, help="iterations of the plain one-pass kernel timed for extra.plain_one_pass_kernel (0: skip)")
    ap.add_argument("--single-steps", type=int, default=2, help="realizations of the single-point leg (extra.single_point)")
    ap.add_argument("--cpu-threads", type=int, default=0, help="host threads of the CPU arm (0 = every core this process may use)")
    ap.add_argument("--cpu-cg-iters", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true")
    ap.add_argument("--no-slab", action="store_true")
    ap.add_argument("--slab-L", type=int, default=0)
    ap.add_argument("--slab-iters", type=int, default=200)
    ap.add_argument("--check-tol", action="store_true", default=True)
    ap.add_argument("--measure-plain", action="store_true",
                    help="also solve the timed realizations with the plain one-pass kernel and write profiles/bench_iters.json")
    ap.add_argument("--solver", default="auto", choices=["auto", "classic", "plain"],
                    help="auto: deflated one-pass kernel (perc_set_solver 0); plain: one-pass without deflation; classic: two-kernel form")
    ap.add_argument("--voltages", action="store_true",
                    help="form the interior voltages too (perc_conduct instead of perc_conduct_g; same G)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
