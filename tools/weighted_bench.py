"""per-iteration device time of the weighted Kirchhoff solver (perc_set_bond_conductance, csrc/pcg_weighted.cu) on one BASELINE
configs[2] realization, next to the uniform-g0 two-kernel form on the same lattice"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import percolation_b200 as P

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=4096)
ap.add_argument("--iters", type=int, default=300)
ap.add_argument("--lattice", type=int, default=1)
args = ap.parse_args()
with P.Lattice(args.lattice, args.L, args.L, 0) as L:
    L.generate(20240611, 3, int(0.8 * L.t), int((0.70 if args.lattice == 1 else 0.5) * L.nb))
    L.label(P.MIXED)
    interior = L.t - 2 * args.L
    L.set_solver(1)
    r = L.conduct(0, tol=1e-30, itmax=args.iters)
    ph = L.phase_ms()
    print("uniform g0, two-kernel form with voltages: %.4f + %.4f ms per iteration = %.0f GB/s at 66 B per site" %
          (ph[6], ph[7], 66.0 * interior / ((ph[6] + ph[7]) * 1e-3) / 1e9))
    w = 0.5 + np.random.default_rng(1).random(L.nb)
    L.set_bond_conductance(w)
    r = L.conduct(0, tol=1e-30, itmax=args.iters)
    ph = L.phase_ms()
    B = 112.0 if args.lattice == 1 else 144.0
    print("per-bond conductances: %.4f (p update + SpMV) + %.4f (update) ms per iteration = %.0f GB/s at %.0f B per site = %.1f %% of 6455.6" %
          (ph[6], ph[7], B * interior / ((ph[6] + ph[7]) * 1e-3) / 1e9, B, 100 * B * interior / ((ph[6] + ph[7]) * 1e-3) / 1e9 / 6455.6))
