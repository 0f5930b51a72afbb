"""CPU check of the one-pass PCG kernel: tests/pcg_fused_emul.cpp compiles the kernel's shared source
(percolation_b200/csrc/pcg_fused_tile.cuh) with g++ and runs every phase of every tile thread by thread.
Gtop / Gbot must agree with the oracle's Jacobi-PCG (same matrix, same stopping rule, Sq/bondc.f:465-595)
within the north star's 1e-9, and the iteration counts must be those of linbcg (the recurrences are the same
iterates in exact arithmetic)."""
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
    so = os.path.join(out, "libpcg_fused_emul.so")
    src = os.path.join(HERE, "pcg_fused_emul.cpp")
    deps = [src, os.path.join(ROOT, "percolation_b200", "csrc", "pcg_fused_tile.cuh"),
            os.path.join(ROOT, "percolation_b200", "csrc", "geometry.cuh"),
            os.path.join(ROOT, "percolation_b200", "csrc", "pcg_defl_host.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-x", "c++", src, "-o", so])
    return C.CDLL(so)


def run_emul(lib, lat, m, n, w, tol, itmax, Va=1.0, g0=1.0, gleak=1e-12, read_thresh=1e-10, cfg=0, pbc=0):
    Gt, Gb, err = C.c_double(), C.c_double(), C.c_double()
    it, fast = C.c_int(), C.c_int()
    w = np.ascontiguousarray(w, np.float64)
    rc = lib.fused_emul_solve(C.c_int(lat), C.c_int(m), C.c_int(n), w.ctypes.data_as(C.POINTER(C.c_double)),
                              C.c_double(Va), C.c_double(g0), C.c_double(gleak), C.c_double(tol), C.c_int(itmax),
                              C.c_double(read_thresh), C.byref(Gt), C.byref(Gb), C.byref(it), C.byref(err), C.byref(fast), C.c_int(cfg), C.c_int(pbc))
    assert rc == 0
    return {"Gtop": Gt.value, "Gbot": Gb.value, "iter": it.value, "err": err.value, "tiles_fast": fast.value}


def spanning_case(O, lat, kind, m, n, ps, pb, seed, pbc=0):
    """one realization with a spanning cluster -> per-bond weights of the Kirchhoff problem (None if nothing spans)"""
    t = m * n
    b1, b2 = O.bondlist(lat, m, n, pbc)
    nb = len(b1)
    rng = np.random.default_rng(seed)
    socc = (rng.random(t) < ps).astype(np.uint8) if kind != O.BOND else None
    bocc = (rng.random(nb) < pb).astype(np.uint8) if kind != O.SITE else None
    ws, wb, wsz, ncl, wmax = O.label_uf(kind, lat, m, n, pbc, b1, b2, site_occ=socc, bond_occ=bocc)
    ids = O.spanning(kind, m, n, b1, b2, ws, wb)
    if len(ids) == 0:
        return None
    return b1, b2, O.weights(kind, b1, b2, ws, wb, int(ids[0]))


# shapes: one partial tile; several tiles in x with interior (fast-path) tiles; tile rows that straddle; m = 16
CASES = [
    (1, "MIXED", 48, 40, 0.85, 0.75), (1, "BOND", 400, 70, 0.0, 0.56), (1, "SITE", 144, 100, 0.65, 0.0),
    (1, "MIXED", 528, 101, 0.85, 0.72), (1, "BOND", 16, 130, 0.0, 0.75), (1, "SITE", 256, 33, 0.66, 0.0),
    (2, "BOND", 48, 40, 0.0, 0.42), (2, "SITE", 400, 70, 0.56, 0.0), (2, "MIXED", 144, 100, 0.8, 0.6),
    (2, "BOND", 528, 101, 0.0, 0.40), (2, "SITE", 16, 130, 0.8, 0.0), (2, "MIXED", 256, 33, 0.85, 0.6),
]


# every case with the plain one-pass variant (pcg_fused_tile.cuh: FtCfgA3 = 2); two cases with interior tiles for the
# first variant (0 = FtCfgA: r as the state vector, phase U, per-tile reductions)
DEFAULT_CFG = 2
CFG_CASES = [(DEFAULT_CFG,) + c for c in CASES] + [(0, 1, "MIXED", 400, 84, 0.85, 0.72), (0, 2, "BOND", 400, 84, 0.0, 0.40)]


@pytest.mark.parametrize("cfg,lat,kind,m,n,ps,pb", CFG_CASES)
def test_emulated_fused_pcg_matches_oracle(emul, O, cfg, lat, kind, m, n, ps, pb):
    kind = getattr(O, kind)
    for seed in range(20):
        case = spanning_case(O, lat, kind, m, n, ps, pb, 7000 + 13 * m + n + seed)
        if case is not None:
            break
    else:
        pytest.fail("no spanning realization among the seeds")
    b1, b2, w = case
    ref = O.conduct_cg(m, n, b1, b2, w, tol=1e-13, itmax=200000)
    got = run_emul(emul, lat, m, n, w, 1e-13, 200000, cfg=cfg)
    assert abs(got["Gtop"] - ref["Gtop"]) <= 1e-9 * abs(ref["Gtop"]), (got, ref["Gtop"])
    assert abs(got["Gbot"] - ref["Gbot"]) <= 1e-9 * abs(ref["Gbot"]), (got, ref["Gbot"])
    assert abs(got["iter"] - ref["iter"]) <= max(3, ref["iter"] // 100), (got["iter"], ref["iter"])
    assert got["err"] <= 1e-13
    if m >= 400 and n >= 70:
        assert got["tiles_fast"] > 0          # the geometry-free fast path was exercised
    if cfg != DEFAULT_CFG:
        return
    # the reference's own defaults (tol 1e-8, itmax 2500, Sq/bondc.f:545): same iteration count, same G to 1e-6
    ref8 = O.conduct_cg(m, n, b1, b2, w)
    got8 = run_emul(emul, lat, m, n, w, 1e-8, 2500, cfg=cfg)
    assert abs(got8["iter"] - ref8["iter"]) <= 1
    assert abs(got8["Gtop"] - ref8["Gtop"]) <= 1e-6 * abs(ref8["Gtop"])


# periodic wrap in x (pbc = 1): the plain one-pass kernel when the seam falls on a tile border (m a multiple of 128) -- one tile
# per lattice row (both halos wrap), two and three tiles, both lattices, tile rows that straddle
PBC_CASES = [(1, "MIXED", 128, 40, 0.85, 0.72), (1, "BOND", 256, 70, 0.0, 0.55), (1, "SITE", 128, 33, 0.66, 0.0),
             (2, "SITE", 128, 50, 0.56, 0.0), (2, "MIXED", 384, 66, 0.8, 0.6), (2, "BOND", 256, 36, 0.0, 0.40)]


@pytest.mark.parametrize("lat,kind,m,n,ps,pb", PBC_CASES)
def test_emulated_fused_pcg_periodic_wrap(emul, O, lat, kind, m, n, ps, pb):
    kind = getattr(O, kind)
    for seed in range(20):
        case = spanning_case(O, lat, kind, m, n, ps, pb, 9000 + 13 * m + n + seed, pbc=1)
        if case is not None:
            break
    else:
        pytest.fail("no spanning realization among the seeds")
    b1, b2, w = case
    ref = O.conduct_cg(m, n, b1, b2, w, tol=1e-13, itmax=200000)
    got = run_emul(emul, lat, m, n, w, 1e-13, 200000, cfg=2, pbc=1)
    assert abs(got["Gtop"] - ref["Gtop"]) <= 1e-9 * abs(ref["Gtop"]), (got, ref["Gtop"])
    assert abs(got["Gbot"] - ref["Gbot"]) <= 1e-9 * abs(ref["Gbot"]), (got, ref["Gbot"])
    assert abs(got["iter"] - ref["iter"]) <= max(3, ref["iter"] // 100), (got["iter"], ref["iter"])
    assert got["err"] <= 1e-13
    # without the wrap the same bonds give a different conductance: the wrap bonds matter in these realizations
    lit = O.conduct_literal(m, n, b1, b2, w)
    got8 = run_emul(emul, lat, m, n, w, 1e-8, 2500, cfg=2, pbc=1)
    assert abs(got8["iter"] - lit["iter"]) <= 1 and abs(got8["Gtop"] - lit["Gtop"]) <= 1e-6 * abs(lit["Gtop"])


@pytest.mark.parametrize("lat,kind,m,n,ps,pb", PBC_CASES)
def test_emulated_deflated_pcg_periodic_wrap(emul, O, lat, kind, m, n, ps, pb, bw=0, bh=0):
    """the deflated sweep with pbc = 1: the coarse operator couples the first and the last block column, the crossing currents
    and the mu records wrap; same G as the oracle's Jacobi-PCG, fewer iterations (the library's own block choice: one tile per
    block at these sizes, 1 / 2 / 3 block columns)"""
    kind = getattr(O, kind)
    for seed in range(20):
        case = spanning_case(O, lat, kind, m, n, ps, pb, 9000 + 13 * m + n + seed, pbc=1)
        if case is not None:
            break
    else:
        pytest.fail("no spanning realization among the seeds")
    b1, b2, w = case
    ref = O.conduct_cg(m, n, b1, b2, w, tol=1e-13, itmax=200000)
    got = run_emul_defl(emul, lat, m, n, w, 1e-13, 200000, bw=bw, bh=bh, pbc=1)
    assert abs(got["Gtop"] - ref["Gtop"]) <= 1e-9 * abs(ref["Gtop"]), (got, ref["Gtop"])
    assert abs(got["Gbot"] - ref["Gbot"]) <= 1e-9 * abs(ref["Gbot"]), (got, ref["Gbot"])
    assert got["err"] <= 1e-13 and got["iter"] <= ref["iter"] + 3


@pytest.mark.parametrize("lat,m,n,bw,bh", [(1, 768, 70, 2, 1), (2, 512, 130, 2, 2), (2, 256, 70, 2, 1)])
def test_emulated_deflated_pcg_periodic_wrap_wide_blocks(emul, O, lat, m, n, bw, bh):
    """periodic wrap with blocks of several tiles: the seam lies between the last and the first block column (3 resp. 2 of them);
    the third shape has ONE block column of two tiles: the tiles wrap, no block border lies on the seam"""
    for seed in range(20):
        case = spanning_case(O, lat, O.MIXED, m, n, 0.85, 0.7, 7100 + m + seed, pbc=1)
        if case is not None:
            break
    else:
        pytest.fail("no spanning realization among the seeds")
    b1, b2, w = case
    ref = O.conduct_cg(m, n, b1, b2, w, tol=1e-13, itmax=200000)
    got = run_emul_defl(emul, lat, m, n, w, 1e-13, 200000, bw=bw, bh=bh, pbc=1)
    assert abs(got["Gtop"] - ref["Gtop"]) <= 1e-9 * abs(ref["Gtop"]), (got, ref["Gtop"])
    assert abs(got["Gbot"] - ref["Gbot"]) <= 1e-9 * abs(ref["Gbot"]), (got, ref["Gbot"])
    assert got["err"] <= 1e-13 and got["iter"] <= ref["iter"] + 3


def test_emulated_fused_full_lattice_closed_form(emul, O):
    # every bond conducting: G = m / (n - 1) on the square lattice (series / parallel)
    m, n = 272, 67
    b1, b2 = O.bondlist(O.SQUARE, m, n, 0)
    got = run_emul(emul, O.SQUARE, m, n, np.ones(len(b1)), 1e-13, 100000)
    assert abs(got["Gtop"] - m / (n - 1)) < 1e-9 and abs(got["Gbot"] - m / (n - 1)) < 1e-9


def run_emul_defl(lib, lat, m, n, w, tol, itmax, Va=1.0, g0=1.0, gleak=1e-12, read_thresh=1e-10, bw=0, bh=0, pbc=0):
    Gt, Gb, err = C.c_double(), C.c_double(), C.c_double()
    it, k = C.c_int(), C.c_int()
    w = np.ascontiguousarray(w, np.float64)
    rc = lib.fused_emul_solve_defl(C.c_int(lat), C.c_int(m), C.c_int(n), w.ctypes.data_as(C.POINTER(C.c_double)),
                                   C.c_double(Va), C.c_double(g0), C.c_double(gleak), C.c_double(tol), C.c_int(itmax),
                                   C.c_double(read_thresh), C.byref(Gt), C.byref(Gb), C.byref(it), C.byref(err), C.byref(k),
                                   C.c_int(bw), C.c_int(bh), None, C.c_int(pbc))
    assert rc == 0, rc
    return {"Gtop": Gt.value, "Gbot": Gb.value, "iter": it.value, "err": err.value, "coarse": k.value}


# deflated one-pass solver (FtCfgD; blocks of bw x bh tiles): same G as the oracle's Jacobi-PCG to 1e-9, err is the TRUE
# relative residual |r| / |D^-1 b| of linbcg's stopping rule, and never more iterations than the plain solver
DEFL_CASES = [
    (1, "MIXED", 528, 101, 0.85, 0.72, 1, 1), (1, "BOND", 400, 70, 0.0, 0.56, 1, 1), (1, "SITE", 144, 100, 0.65, 0.0, 1, 2),
    (1, "MIXED", 48, 40, 0.85, 0.75, 1, 1), (1, "BOND", 16, 130, 0.0, 0.75, 1, 1), (1, "SITE", 256, 33, 0.66, 0.0, 1, 1),
    (1, "MIXED", 272, 194, 0.85, 0.70, 0, 0), (1, "MIXED", 528, 130, 0.85, 0.72, 2, 1),
    (2, "BOND", 528, 101, 0.0, 0.40, 1, 1), (2, "SITE", 400, 70, 0.56, 0.0, 1, 1), (2, "MIXED", 144, 100, 0.8, 0.6, 1, 2),
    (2, "BOND", 48, 40, 0.0, 0.42, 1, 1), (2, "SITE", 16, 130, 0.8, 0.0, 1, 1), (2, "MIXED", 256, 33, 0.85, 0.6, 1, 1),
    (2, "MIXED", 272, 194, 0.85, 0.58, 0, 0), (2, "SITE", 528, 130, 0.56, 0.0, 2, 1),
]


@pytest.mark.parametrize("lat,kind,m,n,ps,pb,bw,bh", DEFL_CASES)
def test_emulated_deflated_pcg_matches_oracle(emul, O, lat, kind, m, n, ps, pb, bw, bh):
    kind = getattr(O, kind)
    for seed in range(20):
        case = spanning_case(O, lat, kind, m, n, ps, pb, 7000 + 13 * m + n + seed)
        if case is not None:
            break
    else:
        pytest.fail("no spanning realization among the seeds")
    b1, b2, w = case
    ref = O.conduct_cg(m, n, b1, b2, w, tol=1e-13, itmax=200000)
    got = run_emul_defl(emul, lat, m, n, w, 1e-13, 200000, bw=bw, bh=bh)
    assert abs(got["Gtop"] - ref["Gtop"]) <= 1e-9 * abs(ref["Gtop"]), (got, ref["Gtop"])
    assert abs(got["Gbot"] - ref["Gbot"]) <= 1e-9 * abs(ref["Gbot"]), (got, ref["Gbot"])
    assert got["err"] <= 1e-13
    assert got["iter"] <= ref["iter"] + 3, (got["iter"], ref["iter"])
    print("coarse %d: %d iterations (plain %d)" % (got["coarse"], got["iter"], ref["iter"]))


def test_emulated_deflated_full_lattice_closed_form(emul, O):
    m, n = 272, 67
    b1, b2 = O.bondlist(O.SQUARE, m, n, 0)
    got = run_emul_defl(emul, O.SQUARE, m, n, np.ones(len(b1)), 1e-13, 100000, bw=1, bh=1)
    assert abs(got["Gtop"] - m / (n - 1)) < 1e-9 and abs(got["Gbot"] - m / (n - 1)) < 1e-9
