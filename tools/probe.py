"""exploratory timing probe (not the bench): phases of one realization at several sizes"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import percolation_b200 as P

def run(lat, kind, Lsz, ps, pb, itmax, tol=1e-10):
    with P.Lattice(lat, Lsz, Lsz, 0) as L:
        t, nb = L.t, L.nb
        ks = int(ps * t) if kind != P.BOND else -1
        kb = int(pb * nb) if kind != P.SITE else -1
        t0 = time.time(); L.generate(4242, 0, ks, kb); L.sync(); tg = time.time() - t0
        for rep in range(3):
            t0 = time.time(); L.label(kind); tl = time.time() - t0
        ms = L.phase_ms()
        sm = L.summary()
        print("L=%d lat=%d kind=%d gen %.2f ms label(wall) %.3f ms phases mask %.3f local %.3f merge %.3f flat+sum %.3f span %.3f | ncl %d maxcs %d nspan %d"
              % (Lsz, lat, kind, tg * 1e3, tl * 1e3, ms[0], ms[1], ms[2], ms[3], ms[4], sm["ncl"], sm["maxcs"], sm["nspan"]), flush=True)
        if sm["nspan"] and itmax:
            t0 = time.time(); r = L.conduct(0, tol=tol, itmax=itmax); tc = time.time() - t0
            ms = L.phase_ms()
            print("   conduct: iters %d err %.3e Gtop %.9e Gbot %.9e wall %.1f ms  pcg %.1f ms  spmv %.4f ms update %.4f ms  per-iter %.4f ms"
                  % (r["iter"], r["err"], r["Gtop"], r["Gbot"], tc * 1e3, ms[5], ms[6], ms[7], ms[5] / max(r["iter"], 1)), flush=True)

if __name__ == "__main__":
    run(P.SQUARE, P.SITE, 1024, 0.60, 0, 3000)
    run(P.SQUARE, P.MIXED, 1024, 0.80, 0.70, 3000)
    run(P.SQUARE, P.MIXED, 4096, 0.80, 0.66, 2000)
    run(P.SQUARE, P.MIXED, 4096, 0.80, 0.70, 2000)
    run(P.TRIANGULAR, P.BOND, 1024, 0, 0.36, 3000)
    run(P.TRIANGULAR, P.SITE, 4096, 0.51, 0, 2000)
    run(P.SQUARE, P.SITE, 8192, 0.5935, 0, 500)
