"""How well is G determined at all?  Independent high-accuracy solve of a golden case (tests/golden/make_conduct_fixtures.py):
sparse LU (scipy splu) + iterative refinement with the residual b - A x accumulated in 80-bit long double, so that x is
accurate far beyond what any fp64 Krylov solve can reach on this matrix (condition number ~ 1e8 .. 1e9 after Jacobi scaling,
1e12 before: the leak bonds).  Prints G of the refined solution next to the oracle's plain-CG value; results are stored in
conduct_fixtures.json as Gtop_direct / Gbot_direct.   usage: python tests/golden/truth_check.py NAME [NAME ...]"""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from oracle import pyoracle as O  # noqa: E402
from philox_np import generate_occupancy  # noqa: E402
from make_conduct_fixtures import CASES, OUT  # noqa: E402


def run(name):
    lat, kind, m, n, ps, pb, seed, stream = CASES[name]
    t = m * n
    b1, b2 = O.bondlist(lat, m, n, 0)
    nb = len(b1)
    ks = int(ps * t) if kind != 2 else -1
    kb = int(pb * nb) if kind != 1 else -1
    socc, bocc = generate_occupancy(seed, stream, lat, m, n, 0, ks, kb, b1, b2)
    ws, wb, wsz, ncl, wmax = O.label_uf(kind, lat, m, n, 0, b1, b2, site_occ=socc, bond_occ=bocc)
    ids = O.spanning(kind, m, n, b1, b2, ws, wb)
    w = O.weights(kind, b1, b2, ws, wb, int(ids[0]))
    a, b = b1.astype(np.int64) - 1, b2.astype(np.int64) - 1
    # full-lattice weighted Laplacian (MATLAB/ConductCalc.m:150-196 route), unknowns = rows 1 .. n-2
    W = sp.coo_matrix((np.concatenate([w, w]), (np.concatenate([a, b]), np.concatenate([b, a]))), shape=(t, t)).tocsr()
    d = np.asarray(W.sum(axis=1)).ravel()
    A = (sp.diags(d) - W).tocsr()
    inter = np.arange(m, t - m)
    Aii = A[inter][:, inter].tocsc()
    V = np.zeros(t); V[t - m:] = 1.0
    rhs = -(A[inter] @ V)
    t0 = time.time()
    lu = spla.splu(Aii)
    x = lu.solve(rhs)
    print("%s: LU in %.0f s" % (name, time.time() - t0), flush=True)
    # long-double residual: r_i = rhs_i - d_i x_i + sum_j w_ij x_j over the interior edges
    ia, ib = a - m, b - m
    keep = (ia >= 0) & (ia < len(inter)) & (ib >= 0) & (ib < len(inter))
    ea, eb, ew = ia[keep], ib[keep], w[keep].astype(np.longdouble)
    dl, rl = d[inter].astype(np.longdouble), rhs.astype(np.longdouble)
    xl = x.astype(np.longdouble)
    for k in range(6):
        acc = dl * xl
        np.subtract.at(acc, ea, ew * xl[eb])
        np.subtract.at(acc, eb, ew * xl[ea])
        r = rl - acc
        dx = lu.solve(np.asarray(r, np.float64))
        xl = xl + dx.astype(np.longdouble)
        print("  refinement %d: |r|/|D^-1 b| = %.2e  |dx|/|x| = %.2e" % (k, float(np.sqrt((r * r).sum()) / np.sqrt(((rl / dl) ** 2).sum())),
                                                                          float(np.sqrt((dx * dx).sum()) / np.sqrt(float((xl * xl).sum())))), flush=True)
    x = np.asarray(xl, np.float64)
    chk = O.conduct_check(m, n, b1, b2, w, x)            # the oracle's own read-out (Sq/bondc.f:554-592) of this x
    rec = json.load(open(OUT))
    ref = rec[name]
    print("%s: direct Gtop %.13e Gbot %.13e | oracle CG (tol 1e-13) Gtop %.13e Gbot %.13e | rel diff %.2e / %.2e"
          % (name, chk["Gtop"], chk["Gbot"], ref["Gtop"], ref["Gbot"], abs(chk["Gtop"] - ref["Gtop"]) / chk["Gtop"],
             abs(chk["Gbot"] - ref["Gbot"]) / chk["Gbot"]), flush=True)
    ref["Gtop_direct"], ref["Gbot_direct"] = chk["Gtop"], chk["Gbot"]
    rec = json.load(open(OUT)); rec[name] = ref
    json.dump(rec, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    for nm in sys.argv[1:]:
        run(nm)
