"""Per-iteration device time of the two PCG solver forms (perc_conduct_g) on one realization of BASELINE
configs[2] (square mixed site/bond, ps 0.80, pb 0.70): the two-kernel form (pcg_pipe_kernel<0|1>, 50 B per
site and iteration) against the one-pass kernel (pcg_fused_kernel, 33 B).  torch-free (ctypes binding only).
Prints FUSED_OK when both forms agree (same iterates: G after a fixed number of iterations to 1e-9)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import percolation_b200 as P

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=4096)
ap.add_argument("--iters", type=int, default=600)
ap.add_argument("--ps", type=float, default=0.80)
ap.add_argument("--pb", type=float, default=0.70)
ap.add_argument("--lattice", type=int, default=1)
ap.add_argument("--pbc", type=int, default=0, help="periodic wrap in x")
ap.add_argument("--converge-fast", action="store_true", help="with --converge: the one-pass kernels at tol 1e-10 only, no two-kernel solve")
ap.add_argument("--converge", action="store_true", help="also run both forms to tol 1e-10")
ap.add_argument("--configs", action="store_true", help="time every tile configuration of the one-pass kernel (perc_set_solver 10 / 12 / 14)")
ap.add_argument("--fused-only", action="store_true", help="skip the two-kernel form (profiling runs)")
ap.add_argument("--default-only", action="store_true", help="one solve with the default solver (profiling runs)")
args = ap.parse_args()
HBM = 6455.6
with P.Lattice(args.lattice, args.L, args.L, args.pbc) as L:
    t = L.t
    L.generate(20240611, 0, int(args.ps * t), int(args.pb * L.nb))
    L.label(P.MIXED)
    ids, _ = L.span()
    print("L=%d lattice=%d nspan=%d" % (args.L, args.lattice, len(ids)), flush=True)
    if not len(ids):
        sys.exit("no spanning cluster")
    interior = t - 2 * args.L
    res = {}
    runs = (("two-kernel", 1), ("one-pass", 2), ("deflated", 0), ("two-kernel", 1), ("one-pass", 2), ("deflated", 0))
    if args.configs:
        runs = tuple(("one-pass-" + nm, md) for nm, md in (("A", 10), ("A3", 12), ("D", 14)))
        runs = (() if args.fused_only else (("two-kernel", 1),)) + runs + (() if args.fused_only else (("two-kernel", 1),) + runs)
    if args.default_only:
        runs = (("one-pass", 0),)
    for name, mode in runs:
        try:
            L.set_solver(mode)
            t0 = time.perf_counter()
            r = L.conduct(0, tol=1e-300, itmax=args.iters - 1, voltages=False)
            wall = time.perf_counter() - t0
        except P.PercError as e:
            print("%-11s FAILED: %s" % (name, e), flush=True)
            continue
        ph = L.phase_ms()
        used = L.solver_used()
        per_it = float(ph[5]) / r["iter"]
        bytes_it = (33.0 if used else 50.0) * interior
        print("%-11s used_fused=%d iters=%d  %.4f ms/iter (solve %.1f ms, wall %.1f ms)  kernels %.4f + %.4f ms  "
              "%.0f GB/s algorithmic = %.1f%% of %.1f | G=%.12e err=%.3e"
              % (name, used, r["iter"], per_it, ph[5], wall * 1e3, ph[6], ph[7], bytes_it / (per_it * 1e-3) / 1e9,
                 100 * bytes_it / (per_it * 1e-3) / 1e9 / HBM, HBM, r["Gtop"], r["err"]), flush=True)
        res[name] = r
    if args.default_only:
        sys.exit(0)
    if args.configs:
        names = [n for n in res if n not in ("two-kernel", "one-pass-D")]
        a = res.get("two-kernel", res[names[0]])
        ok = all(a["iter"] == res[n]["iter"] and abs(a["Gtop"] - res[n]["Gtop"]) <= 1e-9 * abs(a["Gtop"]) for n in names)
        print("FUSED_OK" if ok else "FUSED_MISMATCH", flush=True)
        sys.exit(0)
    a, b = res["two-kernel"], res["one-pass"]
    ok = a["iter"] == b["iter"] and abs(a["Gtop"] - b["Gtop"]) <= 1e-9 * abs(a["Gtop"]) and abs(a["Gbot"] - b["Gbot"]) <= 1e-9 * abs(a["Gbot"])
    ok = ok and abs(a["err"] - b["err"]) <= 1e-6 * abs(a["err"])
    if args.converge and args.converge_fast:
        for name, mode in (("deflated", 0), ("one-pass", 2)):
            L.set_solver(mode)
            t0 = time.perf_counter()
            r = L.conduct(0, tol=1e-10, itmax=4000000, voltages=False)
            print("%-10s converged (tol 1e-10): used=%d iters=%d Gtop=%.13e Gbot=%.13e err=%.3e  %.2f s" % (name, L.solver_used(), r["iter"], r["Gtop"], r["Gbot"], r["err"], time.perf_counter() - t0), flush=True)
            res[name + "-c"] = r
        d, p1 = res["deflated-c"], res["one-pass-c"]
        ok = ok and abs(d["Gtop"] - p1["Gtop"]) <= 1e-7 * p1["Gtop"] and d["iter"] <= p1["iter"]
    elif args.converge:
        for name, mode in (("deflated", 0), ("one-pass", 2), ("two-kernel", 1)):
            for tol in ((1e-10, 1e-13) if name != "two-kernel" else (1e-10,)):
                L.set_solver(mode)
                t0 = time.perf_counter()
                r = L.conduct(0, tol=tol, itmax=4000000, voltages=False)
                print("%-10s converged (tol %.0e): iters=%d Gtop=%.13e Gbot=%.13e err=%.3e  %.2f s  %.4f ms/iter" % (name, tol, r["iter"], r["Gtop"], r["Gbot"], r["err"], time.perf_counter() - t0, L.phase_ms()[6]), flush=True)
                res[name + "-c%.0e" % tol] = r
        res["two-kernel-c"] = res["two-kernel-c1e-10"]; res["one-pass-c"] = res["one-pass-c1e-10"]
        d13, p13 = res["deflated-c1e-13"], res["one-pass-c1e-13"]
        print("deflated vs one-pass at tol 1e-13: relG = %.2e ; iterations %d vs %d (x%.2f)" % (abs(d13["Gtop"] - p13["Gtop"]) / p13["Gtop"], d13["iter"], p13["iter"], p13["iter"] / d13["iter"]), flush=True)
        ok = ok and abs(d13["Gtop"] - p13["Gtop"]) <= 1e-9 * p13["Gtop"]
        a, b = res["two-kernel-c"], res["one-pass-c"]
        ok = ok and abs(a["Gtop"] - b["Gtop"]) <= 1e-8 * abs(a["Gtop"]) and abs(a["iter"] - b["iter"]) <= max(3, a["iter"] // 100)
    print("FUSED_OK" if ok else "FUSED_MISMATCH", flush=True)
