"""CPU check of the CUDA labeling algorithm: tests/ccl_emul.cpp compiles the kernels' shared
source (percolation_b200/csrc/ccl_tile.cuh) with g++ and runs every phase thread by thread;
labels, sizes, cluster count and largest cluster must be bit-exact against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emul():
    out = os.path.join(HERE, "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libccl_emul.so")
    src = os.path.join(HERE, "ccl_emul.cpp")
    deps = [src, os.path.join(ROOT, "percolation_b200", "csrc", "ccl_tile.cuh"),
            os.path.join(ROOT, "percolation_b200", "csrc", "geometry.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", so])
    return C.CDLL(so)


def run_emul(lib, lat, m, n, pbc, kind, socc, bocc):
    t = m * n
    label = np.zeros(t, np.int32)
    csize = np.zeros(t, np.int32)
    out = np.zeros(8, np.int64)
    bp = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint8)) if a is not None else None
    rc = lib.ccl_emul(lat, m, n, pbc, kind, bp(socc), bp(bocc), label.ctypes.data_as(C.POINTER(C.c_int32)),
                      csize.ctypes.data_as(C.POINTER(C.c_int32)), out.ctypes.data_as(C.POINTER(C.c_int64)))
    assert rc == 0
    return label, csize, out


SHAPES = [(50, 50), (70, 45), (32, 32), (34, 66), (128, 96), (6, 3), (160, 130), (256, 128), (130, 70), (400, 200)]
CASES = [(lat, m, n, pbc) for lat in (1, 2) for (m, n) in SHAPES for pbc in (0, 1)]


@pytest.mark.parametrize("var", [0, 1, 2])
@pytest.mark.parametrize("lat,m,n,pbc", CASES)
def test_emulated_ccl_matches_oracle(emul, O, lat, m, n, pbc, var):
    """var = 2: the tile kernel as the library runs it (per-site roots derived in the label phase, two runs per trip of the
    per-run loop); var = 0 / 1: its earlier forms (roots staged per run / one run per trip), kept as a cross-check"""
    emul.ccl_emul_set_variant(var)
    t = m * n
    b1, b2 = O.bondlist(lat, m, n, pbc)
    nb = len(b1)
    rng = np.random.default_rng(1000 * m + 10 * n + pbc + lat)
    for kind in (O.SITE, O.BOND, O.MIXED):
        for ps, pb in ((0.0, 0.0), (0.45, 0.3), (0.6, 0.5), (0.8, 0.65), (1.0, 0.0), (0.3, 1.0), (1.0, 1.0)):
            socc = (rng.random(t) < ps).astype(np.uint8) if kind != O.BOND else None
            bocc = (rng.random(nb) < pb).astype(np.uint8) if kind != O.SITE else None
            ws, wb, wsz, ncl, wmax = O.label_uf(kind, lat, m, n, pbc, b1, b2, site_occ=socc, bond_occ=bocc)
            label, csize, out = run_emul(emul, lat, m, n, pbc, kind, socc, bocc)
            if kind != O.BOND:
                assert (label == ws).all(), (kind, ps, pb)
            else:
                # bond problem: every occupied bond carries the label of its lower end site
                occ = np.nonzero(bocc)[0]
                assert (label[b1[occ] - 1] == wb[occ]).all() and (label[b2[occ] - 1] == wb[occ]).all()
            assert (csize == wsz[1:t + 1]).all(), (kind, ps, pb)
            assert out[0] + out[1] == ncl, (kind, ps, pb, out, ncl)
            maxcs = out[2] if out[2] else (1 if out[1] else 0)
            assert maxcs == wmax
            if out[2]:
                assert csize[out[3] - 1] == out[2] and (csize[:out[3] - 1] < out[2]).all()   # min label among the largest


def run_emul_slab(lib, lat, m, n, pbc, kind, nranks, rank, socc, bocc):
    out = np.zeros(16, np.int64)
    rows = n // nranks + 3
    label = np.zeros(m * rows, np.int32)
    csize = np.zeros(m * rows, np.int32)
    bp = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint8)) if a is not None else None
    rc = lib.ccl_emul_slab(lat, m, n, pbc, kind, nranks, rank, bp(socc), bp(bocc), label.ctypes.data_as(C.POINTER(C.c_int32)),
                           csize.ctypes.data_as(C.POINTER(C.c_int32)), out.ctypes.data_as(C.POINTER(C.c_int64)))
    assert rc == 0
    y0, nl, lo, hi = (int(v) for v in out[5:9])
    return label[:m * nl], csize[:m * nl], int(out[1]), (y0, nl, lo, hi)


@pytest.mark.parametrize("lat,m,n,pbc,nranks", [(1, 48, 40, 0, 2), (2, 64, 70, 1, 3), (1, 144, 140, 1, 2), (2, 32, 24, 0, 4)])
def test_emulated_slab_labeling_plus_stitch_matches_oracle(emul, O, lat, m, n, pbc, nranks):
    """slab decomposition, end to end on the CPU: the kernels' labeling source on every rank's slab (sizes of
    the owned rows only), the interface blocks the GPU would all-gather, the library's host stitch -- against
    the oracle's labeling of the whole lattice, for the three problem kinds"""
    import percolation_b200 as P
    from percolation_b200 import build
    build.build()
    emul.ccl_emul_set_variant(0)
    t = m * n
    b1, b2 = O.bondlist(lat, m, n, pbc)
    nb = len(b1)
    rng = np.random.default_rng(5 * m + n + nranks)
    for kind in (O.SITE, O.BOND, O.MIXED):
        for ps, pb in ((0.62, 0.55), (0.8, 0.7), (1.0, 0.3), (0.35, 1.0)):
            socc = (rng.random(t) < ps).astype(np.uint8) if kind != O.BOND else None
            bocc = (rng.random(nb) < pb).astype(np.uint8) if kind != O.SITE else None
            ws, wb, wsz, ncl, wmax = O.label_uf(kind, lat, m, n, pbc, b1, b2, site_occ=socc, bond_occ=bocc)
            blocks, locs, nlone = [], [], 0
            for r in range(nranks):
                lab, csz, lone, (y0, nl, lo, hi) = run_emul_slab(emul, lat, m, n, pbc, kind, nranks, r, socc, bocc)
                nlone += lone
                lab = lab.astype(np.int64)
                off = y0 * m
                gid = lambda a: np.where(a > 0, a + off, 0)
                blk = np.zeros(5 * m + 8, np.int64)
                rowA, rowB = lab[lo * m:(lo + 1) * m], lab[hi * m:(hi + 1) * m] if r + 1 < nranks else None
                if r > 0:
                    blk[0:m] = gid(rowA)
                    blk[2 * m:3 * m] = np.where(rowA > 0, csz[np.maximum(rowA, 1) - 1], 0)
                if r + 1 < nranks:
                    blk[m:2 * m] = gid(rowB)
                    blk[3 * m:4 * m] = np.where(rowB > 0, csz[np.maximum(rowB, 1) - 1], 0)
                else:
                    blk[4 * m:5 * m] = gid(lab[(nl - 1) * m:])
                roots = np.nonzero(csz > 0)[0]
                blk[5 * m + 0] = len(roots)
                blk[5 * m + 1] = lone
                if len(roots):
                    best = roots[np.argmax(csz[roots])]
                    blk[5 * m + 2], blk[5 * m + 3] = csz[best], best + 1 + off
                blocks.append(blk)
                locs.append((lab, csz, y0, lo, hi))
            gathered = np.concatenate(blocks)
            for r in range(nranks):
                res = P.stitch_host(nranks, r, m, gathered)
                assert res["ncl"] + res["nlone"] == ncl, (kind, ps, pb, res["ncl"], res["nlone"], ncl)
                assert (res["maxcs"] if res["maxcs"] else (1 if res["nlone"] else 0)) == wmax
                lab, csz, y0, lo, hi = locs[r]
                cls = {int(p[0]): (int(p[2]), int(p[3])) for p in res["pairs"]}
                own = lab[lo * m:hi * m]
                glob = np.where(own > 0, own + y0 * m, 0)
                out = np.array([cls[g][0] if g in cls else g for g in glob], np.int64)
                ya = y0 + lo
                if kind != O.BOND:
                    assert (out == ws[ya * m:ya * m + len(out)]).all(), (kind, ps, pb, r)
                # sizes: interface classes carry the lattice-wide total; the others are complete on their rank
                for g, (cid, tot) in cls.items():
                    assert wsz[cid] == tot, (kind, ps, pb, r, g, cid, tot, wsz[cid])
                for root in np.nonzero(csz > 0)[0]:
                    g = int(root) + 1 + y0 * m
                    if g not in cls:
                        assert wsz[g] == csz[root], (kind, ps, pb, r, g)
