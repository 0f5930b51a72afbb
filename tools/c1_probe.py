import sys, os, time
sys.path.insert(0, "/root/repo")
import percolation_b200 as P
with P.Lattice(P.SQUARE, 100, 100, 0) as L:
    ks = 6000
    for rep in range(3):
        t0 = time.perf_counter()
        G, it, st = L.batch_conduct(P.SITE, 1000, 20240611, 0, ks, 0, tol=1e-8, itmax=2500)
        print("rep", rep, time.perf_counter() - t0, (it >= 0).mean(), it[it >= 0].mean(), (it > 2500).sum(), flush=True)
