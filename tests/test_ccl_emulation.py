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


@pytest.mark.parametrize("lat,m,n,pbc", CASES)
def test_emulated_ccl_matches_oracle(emul, O, lat, m, n, pbc):
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
