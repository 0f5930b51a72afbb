"""ctypes front-end of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product (percolation_b200) never
does.  See oracle/perc_oracle.h for what the oracle restates and how it is
pinned.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libperc_oracle.so")

SQUARE, TRIANGULAR = 1, 2
SITE, BOND, MIXED = 1, 2, 3


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("perc_oracle.c", "perc_oracle.h", "Makefile")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src)):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class Result(C.Structure):
    _fields_ = [("cln", C.c_int), ("maxcs", C.c_int), ("maxcn", C.c_int),
                ("perccln", C.c_int), ("perccls", C.c_int), ("filled", C.c_int)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_rand.restype = C.c_float
        _lib.orc_fraction.restype = C.c_float
        _lib.orc_cg_time_iters.restype = C.c_double
    return _lib


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _bp(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8)) if a is not None else None


def nb(lattice, m, n, pbc):
    return lib().orc_nb(lattice, m, n, pbc)


def nearestn(lattice, m, n, pbc, rn):
    nn = np.zeros(10, np.int32)
    lib().orc_nearestn(lattice, m, n, pbc, int(rn), _ip(nn))
    return nn[: lib().orc_scn(lattice)].copy()


def bondlist(lattice, m, n, pbc):
    cnt = nb(lattice, m, n, pbc)
    b1 = np.zeros(cnt + 8, np.int32)
    b2 = np.zeros(cnt + 8, np.int32)
    got = lib().orc_bondlist(lattice, m, n, pbc, _ip(b1), _ip(b2))
    assert got == cnt, (got, cnt)
    return b1[:cnt].copy(), b2[:cnt].copy()


def srand(seed):
    lib().orc_srand(int(seed))


def rand():
    return np.float32(lib().orc_rand())


def irand():
    return lib().orc_irand()


def shuffle_sites(seed, t):
    order = np.zeros(t, np.int32)
    lib().orc_shuffle_sites(int(seed), t, _ip(order))
    return order


def shuffle_bonds(seed, b1, b2):
    bo1 = np.ascontiguousarray(b1, np.int32).copy()
    bo2 = np.ascontiguousarray(b2, np.int32).copy()
    lib().orc_shuffle_bonds(int(seed), len(bo1), _ip(bo1), _ip(bo2))
    return bo1, bo2


def seed_table(master, count, mult):
    out = np.zeros(count, np.int32)
    lib().orc_seed_table(int(master), count, int(mult), _ip(out))
    return out


def sb_seed_tables(master, which, iters, npseed=100):
    pseed = np.zeros(npseed, np.int32)
    ss = np.zeros(iters, np.int32)
    bs = np.zeros(iters, np.int32)
    lib().orc_sb_seed_tables(int(master), npseed, which, iters, _ip(pseed), _ip(ss), _ip(bs))
    return pseed, ss, bs


def fill_count(p, N):
    return lib().orc_fill_count(C.c_double(p), int(N))


def sweep_table(p0, dp, npts, nbonds):
    pb = np.zeros(npts, np.float64)
    nbarr = np.zeros(npts, np.int32)
    lib().orc_sweep_table(C.c_double(p0), C.c_double(dp), npts, nbonds, _dp(pb), _ip(nbarr))
    return pb, nbarr


def fraction(filled, total):
    return np.float32(lib().orc_fraction(int(filled), int(total)))


def site_literal(lattice, m, n, pbc, order, k, stop_at_span=False):
    t = m * n
    s = np.zeros(t, np.int32)
    c = np.zeros(t + 2, np.int32)
    res = Result()
    order = np.ascontiguousarray(order, np.int32)
    lib().orc_site_literal(lattice, m, n, pbc, _ip(order), int(k), int(stop_at_span), _ip(s), _ip(c), C.byref(res))
    return s, c, res.asdict()


def bond_literal(lattice, m, n, pbc, b1, b2, bo1, bo2, k, stop_at_span=False):
    cnt = len(b1)
    b3 = np.zeros(cnt, np.int32)
    c = np.zeros(cnt + 2, np.int32)
    res = Result()
    lib().orc_bond_literal(lattice, m, n, pbc, cnt, _ip(b1), _ip(b2), _ip(bo1), _ip(bo2), int(k),
                           int(stop_at_span), _ip(b3), _ip(c), C.byref(res))
    return b3, c, res.asdict()


def sitebond_literal(lattice, m, n, pbc, b1, b2, sorder, ks, bo1, bo2, kb, stop_at_span=False):
    t, cnt = m * n, len(b1)
    s = np.zeros(t, np.int32)
    b3 = np.zeros(cnt, np.int32)
    c = np.zeros(t + cnt + 2, np.int32)
    res = Result()
    sorder = np.ascontiguousarray(sorder, np.int32)
    lib().orc_sitebond_literal(lattice, m, n, pbc, cnt, _ip(b1), _ip(b2), _ip(sorder), int(ks),
                               _ip(bo1), _ip(bo2), int(kb), int(stop_at_span), _ip(s), _ip(b3), _ip(c), C.byref(res))
    return s, b3, c, res.asdict()


def bondsite_literal(lattice, m, n, pbc, b1, b2, bo1, bo2, kb, sorder, ks, stop_at_span=False):
    t, cnt = m * n, len(b1)
    s = np.zeros(t, np.int32)
    b3 = np.zeros(cnt, np.int32)
    c = np.zeros(t + cnt + 2, np.int32)
    res = Result()
    sorder = np.ascontiguousarray(sorder, np.int32)
    lib().orc_bondsite_literal(lattice, m, n, pbc, cnt, _ip(b1), _ip(b2), _ip(bo1), _ip(bo2), int(kb),
                               _ip(sorder), int(ks), int(stop_at_span), _ip(s), _ip(b3), _ip(c), C.byref(res))
    return s, b3, c, res.asdict()


def label_uf(kind, lattice, m, n, pbc, b1, b2, site_occ=None, bond_occ=None):
    """canonical labels: returns (s_can[t], b3_can[nb], csize[t+1], ncl, maxcs)"""
    t, cnt = m * n, len(b1)
    s_can = np.zeros(t, np.int32)
    b3_can = np.zeros(cnt, np.int32)
    csize = np.zeros(t + 1, np.int32)
    ncl = C.c_int64(0)
    maxcs = C.c_int(0)
    so = None if site_occ is None else np.ascontiguousarray(site_occ, np.uint8)
    bo = None if bond_occ is None else np.ascontiguousarray(bond_occ, np.uint8)
    lib().orc_label_uf(kind, lattice, m, n, pbc, cnt, _ip(b1), _ip(b2), _bp(so), _bp(bo),
                       _ip(s_can), _ip(b3_can), _ip(csize), C.byref(ncl), C.byref(maxcs))
    return s_can, b3_can, csize, ncl.value, maxcs.value


def canonicalise(kind, t, b1, b2, s_ref, b3_ref, c_ref, cln):
    cnt = len(b1)
    s_can = np.zeros(t, np.int32)
    b3_can = np.zeros(cnt, np.int32)
    csize = np.zeros(t + 1, np.int32)
    rc = lib().orc_canonicalise(kind, t, cnt, _ip(b1), _ip(b2), _ip(s_ref), _ip(b3_ref), _ip(c_ref), int(cln),
                                _ip(s_can), _ip(b3_can), _ip(csize))
    return rc, s_can, b3_can, csize


def spanning(kind, m, n, b1, b2, s_can, b3_can, max_ids=4096):
    ids = np.zeros(max_ids, np.int32)
    cnt = lib().orc_spanning(kind, m, n, len(b1), _ip(b1), _ip(b2), _ip(s_can), _ip(b3_can), _ip(ids), max_ids)
    return ids[:min(cnt, max_ids)].copy()


def size_hist(kind, t, csize, b3_can, maxsize):
    hist = np.zeros(maxsize + 1, np.int64)
    lib().orc_size_hist(kind, t, 0 if b3_can is None else len(b3_can), _ip(csize), _ip(b3_can),
                        hist.ctypes.data_as(C.POINTER(C.c_int64)), int(maxsize))
    return hist


def weights(kind, b1, b2, s, b3, perccln, g0=1.0, gleak=1e-12):
    w = np.zeros(len(b1), np.float64)
    lib().orc_weights(kind, len(b1), _ip(b1), _ip(b2), _ip(s), _ip(b3), int(perccln),
                      C.c_double(g0), C.c_double(gleak), _dp(w))
    return w


def _conduct(fn, m, n, b1, b2, w, Va, tol, itmax, read_thresh, x0):
    N = m * n - 2 * m
    Vint = np.zeros(N, np.float64) if x0 is None else np.ascontiguousarray(x0, np.float64).copy()
    Gtop, Gbot, err = C.c_double(0), C.c_double(0), C.c_double(0)
    it = C.c_int(0)
    fn(m, n, len(b1), _ip(b1), _ip(b2), _dp(w), C.c_double(Va), C.c_double(tol), int(itmax),
       C.c_double(read_thresh), _dp(Vint), C.byref(Gtop), C.byref(Gbot), C.byref(it), C.byref(err))
    return dict(Gtop=Gtop.value, Gbot=Gbot.value, iter=it.value, err=err.value, Vint=Vint)


def conduct_literal(m, n, b1, b2, w, Va=1.0, tol=1e-8, itmax=2500, read_thresh=1e-10, x0=None):
    return _conduct(lib().orc_conduct_literal, m, n, b1, b2, w, Va, tol, itmax, read_thresh, x0)


def conduct_cg(m, n, b1, b2, w, Va=1.0, tol=1e-8, itmax=2500, read_thresh=1e-10, x0=None):
    return _conduct(lib().orc_conduct_cg, m, n, b1, b2, w, Va, tol, itmax, read_thresh, x0)


def conduct_check(m, n, b1, b2, w, Vint, Va=1.0, read_thresh=1e-10):
    Gtop, Gbot, err = C.c_double(0), C.c_double(0), C.c_double(0)
    Vint = np.ascontiguousarray(Vint, np.float64)
    lib().orc_conduct_check(m, n, len(b1), _ip(b1), _ip(b2), _dp(w), C.c_double(Va), C.c_double(read_thresh),
                            _dp(Vint), C.byref(Gtop), C.byref(Gbot), C.byref(err))
    return dict(Gtop=Gtop.value, Gbot=Gbot.value, err=err.value)


def cg_time_iters(m, n, b1, b2, w, iters, Va=1.0):
    return lib().orc_cg_time_iters(m, n, len(b1), _ip(b1), _ip(b2), _dp(w), C.c_double(Va), int(iters))
