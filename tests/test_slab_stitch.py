"""Slab decomposition (one lattice over several GPUs), host side: the redundant union-find that
stitches the rank-local labelings across the slab interfaces (percolation_b200/csrc/slab.cu,
stitch_host, through the C-ABI perc_stitch_host).  The rank-local labelings come from the oracle
here (no GPU); the result must equal the oracle's labeling of the whole lattice.  One test drives
it with world_size-2 gloo processes exchanging the interface blocks the way the NCCL all-gather does."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def P():
    import percolation_b200 as P
    from percolation_b200 import build
    build.build()
    P.load()
    return P


def slab_rows(n, nranks, r):
    ya, yb = n * r // nranks, n * (r + 1) // nranks
    hb, ha = (1 if r > 0 else 0), (1 if r + 1 < nranks else 0)
    return ya, yb, ya - hb, yb + ha


def local_block(O, lat, m, n, pbc, socc, nranks, r):
    """what one rank contributes to the all-gather: its slab (+ halo rows) labeled as a lattice of its own"""
    ya, yb, y0, y1 = slab_rows(n, nranks, r)
    nl = y1 - y0
    b1, b2 = O.bondlist(lat, m, nl, pbc)
    loc = socc[y0 * m:y1 * m]
    ws, _, _, _, _ = O.label_uf(O.SITE, lat, m, nl, pbc, b1, b2, site_occ=loc)
    ws = ws.astype(np.int64)
    owned = np.zeros(nl * m, bool)
    owned[(ya - y0) * m:(yb - y0) * m] = True
    size = np.bincount(ws[owned & (ws > 0)], minlength=nl * m + 1)          # owned rows only
    off = y0 * m
    gid = lambda lab: np.where(lab > 0, lab + off, 0)
    blk = np.zeros(5 * m + 8, np.int64)
    rowA, rowB = ws[(ya - y0) * m:(ya - y0 + 1) * m], ws[(yb - y0) * m:(yb - y0 + 1) * m] if r + 1 < nranks else None
    if r > 0:
        blk[0:m] = gid(rowA)
        blk[2 * m:3 * m] = size[rowA]
    if r + 1 < nranks:
        blk[m:2 * m] = gid(rowB)
        blk[3 * m:4 * m] = size[rowB]
    else:
        blk[4 * m:5 * m] = gid(ws[(nl - 1) * m:])
    roots = np.nonzero(size > 0)[0]
    blk[5 * m + 0] = len(roots)
    if len(roots):
        best = roots[np.argmax(size[roots])]                                  # first (smallest label) among the largest
        blk[5 * m + 2], blk[5 * m + 3] = size[best], best + off
    return blk, ws, (ya, yb, y0)


def check_rank(P, O, lat, m, n, pbc, socc, nranks, r, gathered, want, wsz, ncl, wmax, ids):
    _, ws, (ya, yb, y0) = local_block(O, lat, m, n, pbc, socc, nranks, r)
    res = P.stitch_host(nranks, r, m, gathered)
    assert res["ncl"] == ncl and res["maxcs"] == wmax
    assert list(res["span_ids"]) == list(ids) and list(res["span_sizes"]) == [wsz[i] for i in ids]
    if wmax:
        assert wsz[res["maxcn"]] == wmax and (wsz[1:res["maxcn"]] < wmax).all()
    # lattice-wide labels of the owned rows: interface clusters through the class table
    cls = {int(p[0]): (int(p[2]), int(p[3])) for p in res["pairs"]}
    lab = ws[(ya - y0) * m:(yb - y0) * m]
    glob = np.where(lab > 0, lab + y0 * m, 0)
    out = np.array([cls[g][0] if g in cls else g for g in glob], np.int64)
    assert (out == want[ya * m:yb * m]).all()
    for g, (cid, tot) in cls.items():
        assert wsz[cid] == tot


@pytest.mark.parametrize("lat,m,n,pbc,nranks", [(1, 32, 24, 0, 2), (1, 50, 31, 1, 3), (2, 34, 40, 0, 4), (2, 48, 23, 1, 2),
                                               (1, 64, 64, 0, 8), (1, 6, 4, 0, 2)])
def test_stitch_matches_whole_lattice_labeling(P, O, lat, m, n, pbc, nranks):
    rng = np.random.default_rng(7 * m + n + nranks)
    b1, b2 = O.bondlist(lat, m, n, pbc)
    for p in (0.0, 0.35, 0.5, 0.6, 0.72, 1.0):
        socc = (rng.random(m * n) < p).astype(np.uint8)
        want, _, wsz, ncl, wmax = O.label_uf(O.SITE, lat, m, n, pbc, b1, b2, site_occ=socc)
        ids = O.spanning(O.SITE, m, n, b1, b2, want, None)
        gathered = np.concatenate([local_block(O, lat, m, n, pbc, socc, nranks, r)[0] for r in range(nranks)])
        for r in range(nranks):
            check_rank(P, O, lat, m, n, pbc, socc, nranks, r, gathered, want.astype(np.int64), wsz, ncl, wmax, ids)


def test_stitch_detects_inconsistent_interface(P, O):
    m, n, nranks = 16, 8, 2
    socc = np.ones(m * n, np.uint8)
    blocks = [local_block(O, 1, m, n, 0, socc, nranks, r)[0] for r in range(nranks)]
    blocks[1][3] = 0                                    # rank 1 claims an interface site is empty
    with pytest.raises(P.PercError):
        P.stitch_host(nranks, 0, m, np.concatenate(blocks))


WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
import percolation_b200 as P
from oracle import pyoracle as O
from test_slab_stitch import local_block, check_rank
dist.init_process_group("gloo")
r, G = dist.get_rank(), dist.get_world_size()
lat, m, n, pbc = 2, 40, 30, 1
rng = np.random.default_rng(99)                       # same realization on every rank
socc = (rng.random(m * n) < 0.52).astype(np.uint8)
blk = torch.from_numpy(local_block(O, lat, m, n, pbc, socc, G, r)[0])
out = [torch.zeros_like(blk) for _ in range(G)]
dist.all_gather(out, blk)                             # what ncclAllGather does on the device
gathered = torch.cat(out).numpy()
b1, b2 = O.bondlist(lat, m, n, pbc)
want, _, wsz, ncl, wmax = O.label_uf(O.SITE, lat, m, n, pbc, b1, b2, site_occ=socc)
ids = O.spanning(O.SITE, m, n, b1, b2, want, None)
check_rank(P, O, lat, m, n, pbc, socc, G, r, gathered, want.astype(np.int64), wsz, ncl, wmax, ids)
dist.barrier()
sys.stdout.write("rank%d-ok\n" % r); sys.stdout.flush()
'''


def test_stitch_over_gloo_world_size_2(P, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "rank0-ok" in out.stdout and "rank1-ok" in out.stdout
