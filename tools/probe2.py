"""convergence probe: PCG iterations to tol at several sizes near p_c"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import percolation_b200 as P

def run(lat, kind, Lsz, ps, pb, itmax, tol):
    with P.Lattice(lat, Lsz, Lsz, 0) as L:
        t, nb = L.t, L.nb
        ks = int(ps * t) if kind != P.BOND else -1
        kb = int(pb * nb) if kind != P.SITE else -1
        L.generate(4242, 0, ks, kb)
        if pb < 0:
            r = L.first_span(kind, P.BOND); print("first span kb*=%d f=%.6f" % (r["kstar"], r["f"]))
        L.label(kind)
        sm = L.summary()
        if not sm["nspan"]:
            print("L=%d no span" % Lsz); return
        t0 = time.time(); r = L.conduct(0, tol=tol, itmax=itmax); tc = time.time() - t0
        ms = L.phase_ms()
        print("L=%d lat=%d kind=%d ps=%.3f pb=%.3f tol=%.0e: iters %d err %.3e Gtop %.12e Gbot %.12e wall %.1f ms per-iter %.4f ms"
              % (Lsz, lat, kind, ps, pb, tol, r["iter"], r["err"], r["Gtop"], r["Gbot"], tc * 1e3, ms[5] / max(r["iter"], 1)), flush=True)

if __name__ == "__main__":
    for Lsz in (128, 256, 512, 1024, 2048):
        run(P.SQUARE, P.MIXED, Lsz, 0.80, 0.70, 400000, 1e-10)
    run(P.SQUARE, P.MIXED, 4096, 0.80, 0.70, 100000, 1e-10)
