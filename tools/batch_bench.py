"""batched-realization benchmark: BASELINE configs[0] (square site L=100, p=0.60, 1000 realizations) and
configs[1] (triangular bond L=1024, p=0.35): labeling + spanning + size histogram, device-resident"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import percolation_b200 as P

def run(name, lat, kind, Lsz, ps, pb, nreal):
    with P.Lattice(lat, Lsz, Lsz, 0) as L:
        ks, kb = int(ps * L.t), int(pb * L.nb)
        L.batch(kind, 8, 1, 0, ks, kb, 64)                 # warm-up
        t0 = time.perf_counter()
        hist, st = L.batch(kind, nreal, 20240611, 0, ks, kb, 64)
        dt = time.perf_counter() - t0
        print(json.dumps({"config": name, "L": Lsz, "realizations": nreal, "seconds": dt, "realizations_per_s": nreal / dt,
                          "gsites_per_s": nreal * L.t / dt / 1e9, "spanning_fraction": st["spanning"] / nreal,
                          "mean_ncl": st["sum_ncl"] / nreal, "failed": st["failed"], "launches": L.launch_count()}), flush=True)

def run_conduct(name, lat, kind, Lsz, ps, pb, nreal, tol=1e-8, itmax=2500):
    with P.Lattice(lat, Lsz, Lsz, 0) as L:
        ks, kb = int(ps * L.t), int(pb * L.nb)
        L.batch_conduct(kind, 8, 1, 0, ks, kb, tol=tol, itmax=itmax)
        dt = 1e9
        for rep in range(3):                                     # best of 3 (the first call also pays allocations)
            t0 = time.perf_counter()
            G, iters, st = L.batch_conduct(kind, nreal, 20240611, 0, ks, kb, tol=tol, itmax=itmax)
            dt = min(dt, time.perf_counter() - t0)
        sp = iters >= 0
        print(json.dumps({"config": name, "L": Lsz, "realizations": nreal, "seconds": dt, "realizations_per_s": nreal / dt,
                          "spanning_fraction": float(sp.mean()), "mean_G": float(G[sp, 0].mean()) if sp.any() else 0.0,
                          "mean_iters": float(iters[sp].mean()) if sp.any() else 0.0, "tol": tol, "itmax": itmax}), flush=True)


if __name__ == "__main__":
    run_conduct("C1 square site L=100 p=0.60: labeling + spanning + conductance (reference tol 1e-8, itmax 2500)",
                P.SQUARE, P.SITE, 100, 0.60, 0.0, 1000)
    run("C1 square site L=100 p=0.60", P.SQUARE, P.SITE, 100, 0.60, 0.0, 1000)
    run("C2 triangular bond L=1024 p=0.35", P.TRIANGULAR, P.BOND, 1024, 0.0, 0.35, 512)
    run("square site L=4096 p=0.5927", P.SQUARE, P.SITE, 4096, 0.5927, 0.0, 32)
