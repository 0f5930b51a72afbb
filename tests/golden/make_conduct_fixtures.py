"""Golden conductance values of the ORACLE (oracle/perc_oracle.c: orc_conduct_cg, plain Jacobi-PCG = linbcg on a symmetric
matrix, Sq/bondc.f:750-838) on realizations the GPU generator draws, computed WITHOUT a GPU: the occupancy comes from the
numpy restatement of the K1 generator (tests/philox_np.py), labels / spanning cluster / bond weights from the oracle.
The GPU tests (tests/test_gpu_parity.py::test_conductance_golden) regenerate the same realization on the device
(checked by occupancy checksums) and must reproduce Gtop / Gbot to 1e-9 with every solver form -- including the bench
configuration (square mixed L = 4096, ps 0.80, pb 0.70), where a CPU solve takes hours and cannot run inside a test.

usage: python tests/golden/make_conduct_fixtures.py NAME [NAME ...]     (results are merged into conduct_fixtures.json)
       OMP_NUM_THREADS=6 python tests/golden/make_conduct_fixtures.py sq_mixed_4096
Committed run times on the 8-core container: L=1024 a few minutes, L=2048 ~1/2 h, L=4096 a few hours (OpenMP)."""
import json
import os
import sys
import time
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from oracle import pyoracle as O  # noqa: E402
from philox_np import generate_occupancy  # noqa: E402

# name: lattice, kind, m, n, ps, pb, seed, stream  (seed / stream of sq_mixed_4096 = bench.py's first realization on rank 0)
CASES = {
    "sq_mixed_512": (1, 3, 512, 512, 0.80, 0.70, 20240611, 0),
    "sq_mixed_1024": (1, 3, 1024, 1024, 0.80, 0.70, 20240611, 0),
    "tri_site_1024": (2, 1, 1024, 1024, 0.52, 0.0, 20240611, 0),
    "sq_bond_1024x512": (1, 2, 1024, 512, 0.0, 0.52, 20240611, 1),
    "sq_mixed_2048": (1, 3, 2048, 2048, 0.80, 0.70, 20240611, 0),
    "sq_mixed_4096": (1, 3, 4096, 4096, 0.80, 0.70, 20240611, 0),
}
OUT = os.path.join(HERE, "conduct_fixtures.json")


def run(name):
    lat, kind, m, n, ps, pb, seed, stream = CASES[name]
    t = m * n
    b1, b2 = O.bondlist(lat, m, n, 0)
    nb = len(b1)
    ks = int(ps * t) if kind != 2 else -1
    kb = int(pb * nb) if kind != 1 else -1
    t0 = time.time()
    socc, bocc = generate_occupancy(seed, stream, lat, m, n, 0, ks, kb, b1, b2)
    ws, wb, wsz, ncl, wmax = O.label_uf(kind, lat, m, n, 0, b1, b2, site_occ=socc, bond_occ=bocc)
    ids = O.spanning(kind, m, n, b1, b2, ws, wb)
    assert len(ids), "realization does not span"
    w = O.weights(kind, b1, b2, ws, wb, int(ids[0]))
    print("%s: labeled in %.0f s, ncl %d, spanning cluster %d" % (name, time.time() - t0, ncl, int(ids[0])), flush=True)
    t0 = time.time()
    ref = O.conduct_cg(m, n, b1, b2, w, tol=1e-13, itmax=10000000)
    rec = dict(lattice=lat, kind=kind, m=m, n=n, ps=ps, pb=pb, seed=seed, stream=stream, ks=ks, kb=kb,
               site_crc=zlib.crc32(socc.tobytes()) if socc is not None else 0,
               bond_crc=zlib.crc32(bocc.tobytes()) if bocc is not None else 0,
               ncl=int(ncl), maxcs=int(wmax), cluster=int(ids[0]), tol=1e-13,
               Gtop=ref["Gtop"], Gbot=ref["Gbot"], iter=int(ref["iter"]), err=ref["err"],
               cpu_seconds=round(time.time() - t0, 1), omp_threads=os.environ.get("OMP_NUM_THREADS", "default"))
    print(name, rec, flush=True)
    allrec = json.load(open(OUT)) if os.path.exists(OUT) else {}
    allrec[name] = rec
    json.dump(allrec, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    for nm in sys.argv[1:]:
        run(nm)
