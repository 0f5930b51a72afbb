"""labeling-only microbenchmark (K1 mask + K2-K5) at the BASELINE sizes"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import percolation_b200 as P

def run(lat, kind, Lsz, ps, pb, reps=5):
    with P.Lattice(lat, Lsz, Lsz, 0) as L:
        t, nb = L.t, L.nb
        L.generate(99, 0, int(ps * t) if kind != P.BOND else -1, int(pb * nb) if kind != P.SITE else -1)
        acc = np.zeros(8)
        for r in range(reps):
            L.label(kind)
            if r: acc += L.phase_ms()
        acc /= (reps - 1)
        tot = acc[1] + acc[2] + acc[3]
        print("L=%d lat=%d kind=%d: mask %.3f local %.3f merge %.3f flatten+summary %.3f span %.3f | CCL %.3f ms = %.1f Gsites/s (%.1f%% of 6455.6 GB/s at 5 B/site)"
              % (Lsz, lat, kind, acc[0], acc[1], acc[2], acc[3], acc[4], tot, t / tot / 1e6, 100 * 5 * t / tot / 1e6 / 6455.6), flush=True)

if __name__ == "__main__":
    quick = len(sys.argv) > 1
    run(P.SQUARE, P.MIXED, 4096, 0.80, 0.70, 3 if quick else 5)
    if not quick:
        run(P.SQUARE, P.SITE, 4096, 0.5927, 0)
        run(P.TRIANGULAR, P.BOND, 1024, 0, 0.35)
        run(P.TRIANGULAR, P.SITE, 16384, 0.5, 0)
