"""BASELINE configs[2] as the reference's drivers run it (Sq/bond_cond.f:208-485, Sq/sb_perc.f): square mixed
site/bond, sites fixed at ps, bonds added in one order; per realization find the first-spanning bond count k*
(perc_first_span) and evaluate the conductance at pb* + 0.005 j, j = 0..npts-1, cold and warm-started."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import percolation_b200 as P

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=1024)
ap.add_argument("--ps", type=float, default=0.80)
ap.add_argument("--npts", type=int, default=5)
ap.add_argument("--tol", type=float, default=1e-10)
args = ap.parse_args()
with P.Lattice(P.SQUARE, args.L, args.L, 0) as L:
    ks = int(args.ps * L.t)
    L.generate(20240611, 0, ks, L.nb)
    t0 = time.perf_counter()
    fs = L.first_span(P.MIXED, P.BOND)
    t_first = time.perf_counter() - t0
    pbs = fs["kstar"] / L.nb
    out = {"L": args.L, "ps": args.ps, "kstar": fs["kstar"], "pb_star": pbs, "first_span_s": t_first, "points": []}
    for mode in ("cold", "warm"):
        tot_it, t0 = 0, time.perf_counter()
        for j in range(args.npts):
            kb = int((pbs + 0.005 * (j + 1)) * L.nb)
            L.set_fill(kb=kb)
            L.label(P.MIXED)
            r = L.conduct(0, tol=args.tol, itmax=4000000, warm=(mode == "warm" and j > 0))
            tot_it += r["iter"]
            out["points"].append({"mode": mode, "pb": kb / L.nb, "G": 0.5 * (r["Gtop"] + r["Gbot"]), "iter": r["iter"]})
        out[mode + "_s"] = time.perf_counter() - t0
        out[mode + "_iters"] = tot_it
    print(json.dumps(out), flush=True)
