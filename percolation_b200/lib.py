"""ctypes binding of libperc_b200.so (the C-ABI of include/perc_abi.h).

Every call goes through the C-ABI exactly as the Fortran ISO_C_BINDING drivers do
(scalars by reference, 1-based ids, column-major 2-D arrays).  There is no CPU
fallback: if the library is missing or no CUDA device is present, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("PERC_B200_SO") or os.path.join(_HERE, "libperc_b200.so")     # (override: kernel experiments)

SQUARE, TRIANGULAR = 1, 2
SITE, BOND, MIXED = 1, 2, 3

E_ARG, E_ODD_M, E_HANDLE, E_STATE, E_SIZE, E_NOSPAN, E_NCCL, E_IFACE, E_SELECT = -1, -2, -3, -4, -5, -6, -7, -8, -9
_ENAMES = {E_ARG: "PERC_E_ARG", E_ODD_M: "PERC_E_ODD_M", E_HANDLE: "PERC_E_HANDLE", E_STATE: "PERC_E_STATE",
           E_SIZE: "PERC_E_SIZE", E_NOSPAN: "PERC_E_NOSPAN", E_NCCL: "PERC_E_NCCL", E_IFACE: "PERC_E_IFACE", E_SELECT: "PERC_E_SELECT"}

# every symbol include/perc_abi.h declares
SYMBOLS = [
    "perc_geom_nb", "perc_geom_bondlist", "perc_geom_nearestn",
    "perc_create", "perc_destroy", "perc_sync",
    "perc_set_site_order", "perc_set_bond_order", "perc_set_fill", "perc_set_occupancy",
    "perc_generate", "perc_get_occupancy",
    "perc_label", "perc_summary", "perc_get_site_labels", "perc_get_bond_labels", "perc_get_sizes",
    "perc_span", "perc_hist", "perc_hist_log2", "perc_site", "perc_bond", "perc_sitebond", "perc_first_span",
    "perc_conduct", "perc_conduct_warm", "perc_conduct_g", "perc_get_voltage", "perc_launch_count", "perc_phase_ms", "perc_stream", "perc_set_solver", "perc_solver_used",
    "perc_create_slab", "perc_comm_unique_id", "perc_comm_init", "perc_slab_rows", "perc_generate_i8",
    "perc_summary_i8", "perc_span_i8", "perc_get_site_labels_i8", "perc_stitch_host",
    "perc_batch", "perc_batch_conduct", "perc_comm_init_rank", "perc_allreduce_stats", "perc_set_bond_conductance", "perc_clear_bond_conductance", "perc_write_txt", "perc_label_incremental",
]


class PercError(RuntimeError):
    def __init__(self, fn, code):
        self.code = code
        what = _ENAMES.get(code, "cudaError %d" % code if code > 0 else "error %d" % code)
        super().__init__("%s failed: %s" % (fn, what))


_lib = None


def load():
    """load the shared library; fails loudly when it has not been built"""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError("libperc_b200.so is not built (run `python -m percolation_b200.build` "
                              "or __graft_entry__.build()); there is no CPU fallback")
        _lib = C.CDLL(SO_PATH)
        for s in SYMBOLS:
            getattr(_lib, s).restype = C.c_int32
    return _lib


def _i32(v):
    return C.byref(C.c_int32(int(v)))


def _i64(v):
    return C.byref(C.c_int64(int(v)))


def _f64(v):
    return C.byref(C.c_double(float(v)))


def _ptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None else None


def _ck(fn, rc):
    if rc != 0:
        raise PercError(fn, rc)


def geom_nb(lattice, m, n, pbc):
    out = C.c_int32(0)
    _ck("perc_geom_nb", load().perc_geom_nb(_i32(lattice), _i32(m), _i32(n), _i32(pbc), C.byref(out)))
    return out.value


def geom_bondlist(lattice, m, n, pbc):
    """(b1, b2): the reference's bond list b(nb,1:2) (Sq/bond.f:112-129)"""
    nb = geom_nb(lattice, m, n, pbc)
    b = np.zeros(2 * nb, np.int32)
    _ck("perc_geom_bondlist", load().perc_geom_bondlist(_i32(lattice), _i32(m), _i32(n), _i32(pbc), _ptr(b, C.c_int32)))
    return b[:nb].copy(), b[nb:].copy()


def geom_nearestn(lattice, m, n, pbc, rn):
    nn = np.zeros(6, np.int32)
    _ck("perc_geom_nearestn", load().perc_geom_nearestn(_i32(lattice), _i32(m), _i32(n), _i32(pbc), _i32(rn),
                                                        _ptr(nn, C.c_int32)))
    return nn[:4 if lattice == SQUARE else 6].copy()


class Lattice:
    """one handle = the reference's parameter block (lattice, m, n, pbc) on one GPU"""

    def __init__(self, lattice, m, n, pbc=0, device=0):
        self.lattice, self.m, self.n, self.pbc = int(lattice), int(m), int(n), int(pbc)
        self.t = self.m * self.n
        self._h = C.c_int64(0)
        self._lib = load()
        _ck("perc_create", self._lib.perc_create(C.byref(self._h), _i32(lattice), _i32(m), _i32(n), _i32(pbc),
                                                 _i32(device)))
        self.nb = geom_nb(lattice, m, n, pbc)

    def close(self):
        if self._h.value:
            self._lib.perc_destroy(C.byref(self._h))
            self._h = C.c_int64(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _call(self, name, *args):
        _ck(name, getattr(self._lib, name)(C.byref(self._h), *args))

    # ---- occupancy
    def set_site_order(self, order):
        order = np.ascontiguousarray(order, np.int32)
        assert order.size == self.t
        self._call("perc_set_site_order", _ptr(order, C.c_int32))

    def set_bond_order(self, bo1, bo2=None):
        """border(nb,2) column-major: pass (lo, hi) columns or one flat array of 2*nb"""
        border = np.ascontiguousarray(np.concatenate([bo1, bo2]) if bo2 is not None else bo1, np.int32)
        assert border.size == 2 * self.nb
        self._call("perc_set_bond_order", _ptr(border, C.c_int32))

    def set_fill(self, ks=-1, kb=-1):
        self._call("perc_set_fill", _i32(ks), _i32(kb))

    def set_occupancy(self, socc=None, bocc=None):
        s = None if socc is None else np.ascontiguousarray(socc, np.uint8)
        b = None if bocc is None else np.ascontiguousarray(bocc, np.uint8)
        self._call("perc_set_occupancy", _ptr(s, C.c_uint8), _ptr(b, C.c_uint8))

    def generate(self, seed, stream=0, ks=-1, kb=-1):
        self._call("perc_generate", _i64(seed), _i64(stream), _i32(ks), _i32(kb))

    def get_occupancy(self, sites=True, bonds=True):
        s = np.zeros(self.t, np.uint8) if sites else None
        b = np.zeros(self.nb, np.uint8) if bonds else None
        self._call("perc_get_occupancy", _ptr(s, C.c_uint8), _ptr(b, C.c_uint8))
        return s, b

    # ---- labeling
    def label(self, kind):
        self._call("perc_label", _i32(kind))

    def label_incremental(self, kind):
        """re-labeling after perc_set_fill RAISED the fill of a labeled handle (same order / generator stream): only the added
        elements are united (Newman-Ziff along a sweep, Sq/bond_cond.f:208-485); otherwise the full pass.  Returns True when
        the incremental pass ran."""
        inc = C.c_int32(0)
        self._call("perc_label_incremental", _i32(kind), C.byref(inc))
        return bool(inc.value)

    def summary(self):
        ncl, maxcs, maxcn, nspan = C.c_int64(0), C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self._call("perc_summary", C.byref(ncl), C.byref(maxcs), C.byref(maxcn), C.byref(nspan))
        return dict(ncl=ncl.value, maxcs=maxcs.value, maxcn=maxcn.value, nspan=nspan.value)

    def site_labels(self):
        s = np.zeros(self.t, np.int32)
        self._call("perc_get_site_labels", _ptr(s, C.c_int32))
        return s

    def bond_labels(self):
        b3 = np.zeros(self.nb, np.int32)
        self._call("perc_get_bond_labels", _ptr(b3, C.c_int32))
        return b3

    def sizes(self):
        c = np.zeros(self.t, np.int32)
        self._call("perc_get_sizes", _ptr(c, C.c_int32))
        return c

    def span(self, max_ids=4096):
        ids = np.zeros(max_ids, np.int32)
        sizes = np.zeros(max_ids, np.int32)
        nspan = C.c_int32(0)
        self._call("perc_span", _i32(max_ids), C.byref(nspan), _ptr(ids, C.c_int32), _ptr(sizes, C.c_int32))
        k = min(nspan.value, max_ids)
        return ids[:k].copy(), sizes[:k].copy()

    def hist(self, nbins):
        h = np.zeros(nbins, np.int64)
        self._call("perc_hist", _i32(nbins), _ptr(h, C.c_int64))
        return h

    def hist_log2(self, nbins=32):
        h = np.zeros(nbins, np.int64)
        self._call("perc_hist_log2", _i32(nbins), _ptr(h, C.c_int64))
        return h

    def site(self, order, k):
        """Sq/site.f:162-344 in one call: returns s, c, dict(maxcs, perccln, perccls)"""
        order = np.ascontiguousarray(order, np.int32)
        s = np.zeros(self.t, np.int32)
        c = np.zeros(self.t, np.int32)
        a, b, d = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self._call("perc_site", _ptr(order, C.c_int32), _i32(k), _ptr(s, C.c_int32), _ptr(c, C.c_int32),
                   C.byref(a), C.byref(b), C.byref(d))
        return s, c, dict(maxcs=a.value, perccln=b.value, perccls=d.value)

    def bond(self, bo1, bo2, k):
        border = np.ascontiguousarray(np.concatenate([bo1, bo2]), np.int32)
        b3 = np.zeros(self.nb, np.int32)
        c = np.zeros(self.t, np.int32)
        a, b, d = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self._call("perc_bond", _ptr(border, C.c_int32), _i32(k), _ptr(b3, C.c_int32), _ptr(c, C.c_int32),
                   C.byref(a), C.byref(b), C.byref(d))
        return b3, c, dict(maxcs=a.value, perccln=b.value, perccls=d.value)

    def sitebond(self, sorder, ks, bo1, bo2, kb):
        sorder = np.ascontiguousarray(sorder, np.int32)
        border = np.ascontiguousarray(np.concatenate([bo1, bo2]), np.int32)
        s = np.zeros(self.t, np.int32)
        b3 = np.zeros(self.nb, np.int32)
        c = np.zeros(self.t, np.int32)
        a, b, d = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self._call("perc_sitebond", _ptr(sorder, C.c_int32), _i32(ks), _ptr(border, C.c_int32), _i32(kb),
                   _ptr(s, C.c_int32), _ptr(b3, C.c_int32), _ptr(c, C.c_int32), C.byref(a), C.byref(b), C.byref(d))
        return s, b3, c, dict(maxcs=a.value, perccln=b.value, perccls=d.value)

    def first_span(self, kind, which):
        kstar, f, maxcs, perccls = C.c_int32(0), C.c_float(0), C.c_int32(0), C.c_int32(0)
        self._call("perc_first_span", _i32(kind), _i32(which), C.byref(kstar), C.byref(f), C.byref(maxcs),
                   C.byref(perccls))
        return dict(kstar=kstar.value, f=np.float32(f.value), maxcs=maxcs.value, perccls=perccls.value)

    # ---- batches of independent realizations
    BATCH_STATS = ("realizations", "sum_ncl", "sum_maxcs", "spanning", "sum_nspan", "sum_perccls", "failed",
                   "sum_maxcs2", "sum_sites", "sum_bonds")

    def batch(self, kind, nreal, seed, stream0=0, ks=0, kb=0, nbins=0):
        """nreal realizations (streams stream0 .. stream0+nreal-1) labeled on the device; returns (hist, stats)"""
        hist = np.zeros(max(nbins, 1), np.int64)
        stats = np.zeros(16, np.int64)
        self._call("perc_batch", _i32(kind), _i32(nreal), _i64(seed), _i64(stream0), _i32(ks), _i32(kb), _i32(nbins),
                   _ptr(hist, C.c_int64), _ptr(stats, C.c_int64))
        return hist[:nbins].copy(), dict(zip(self.BATCH_STATS, (int(v) for v in stats)))

    def batch_conduct(self, kind, nreal, seed, stream0=0, ks=0, kb=0, Va=1.0, g0=1.0, gleak=1e-12, tol=1e-8, itmax=2500,
                      read_thresh=1e-10):
        """nreal realizations: labeling + Kirchhoff conductance of each default spanning cluster; returns
        (G[nreal, 2] = Gtop, Gbot; iters[nreal] (-1: no spanning cluster); stats)"""
        G = np.zeros(2 * nreal, np.float64)
        iters = np.zeros(nreal, np.int32)
        stats = np.zeros(16, np.int64)
        self._call("perc_batch_conduct", _i32(kind), _i32(nreal), _i64(seed), _i64(stream0), _i32(ks), _i32(kb), _f64(Va),
                   _f64(g0), _f64(gleak), _f64(tol), _i32(itmax), _f64(read_thresh), _ptr(G, C.c_double),
                   _ptr(iters, C.c_int32), _ptr(stats, C.c_int64))
        return G.reshape(nreal, 2), iters, dict(zip(self.BATCH_STATS, (int(v) for v in stats)))

    def comm_init_rank(self, nranks, rank, unique_id):
        uid = np.ascontiguousarray(unique_id, np.uint8)
        self._call("perc_comm_init_rank", _i32(nranks), _i32(rank), _ptr(uid, C.c_uint8))

    def allreduce_stats(self, ivals=None, dvals=None):
        """one NCCL reduction (sum) of the statistics of a sharded run; in place, returns the arrays"""
        iv = np.ascontiguousarray(ivals if ivals is not None else [], np.int64)
        dv = np.ascontiguousarray(dvals if dvals is not None else [], np.float64)
        self._call("perc_allreduce_stats", _i32(iv.size), _ptr(iv, C.c_int64), _i32(dv.size), _ptr(dv, C.c_double))
        return iv, dv

    # ---- conductance
    def conduct(self, cluster_id=0, Va=1.0, g0=1.0, gleak=1e-12, tol=1e-8, itmax=2500, read_thresh=1e-10,
                voltages=True, warm=False):
        """Sq/bondc.f:465-595.  voltages=False -> perc_conduct_g (same G, interior voltages not formed);
        warm=True -> perc_conduct_warm (initial guess = the handle's previous voltages)"""
        Gtop, Gbot, err, it = C.c_double(0), C.c_double(0), C.c_double(0), C.c_int32(0)
        self._call("perc_conduct_warm" if warm else "perc_conduct" if voltages else "perc_conduct_g", _i32(cluster_id), _f64(Va), _f64(g0), _f64(gleak), _f64(tol), _i32(itmax),
                   _f64(read_thresh), C.byref(Gtop), C.byref(Gbot), C.byref(it), C.byref(err))
        return dict(Gtop=Gtop.value, Gbot=Gbot.value, iter=it.value, err=err.value)

    TXT_FILES = {"site.txt": 1, "bond.txt": 2, "sbsite.txt": 3, "sbbond.txt": 4, "bondlist.txt": 5}

    def write_txt(self, which, path):
        """one of the reference programs' output files (site.txt, bond.txt, sbsite.txt, sbbond.txt, bondlist.txt) in the
        reference's record format, from the current labeling (Sq/site.f:354-359, Sq/bond.f:443-448, Sq/sitebond.f:469-477)"""
        w = self.TXT_FILES[which] if isinstance(which, str) else int(which)
        p = os.fsencode(path)
        self._call("perc_write_txt", _i32(w), C.c_char_p(p), _i32(len(p)))

    def set_bond_conductance(self, w):
        """per-bond conductances (MATLAB/ConductCalc.m condtype = 2): w[nb] in the reference's bond-row order, used for the
        bonds that conduct; None restores the uniform g0"""
        if w is None:
            self._call("perc_clear_bond_conductance")
            return
        w = np.ascontiguousarray(w, np.float64)
        assert w.size == self.nb
        self._call("perc_set_bond_conductance", _ptr(w, C.c_double))

    def voltage(self):
        v = np.zeros(self.t - 2 * self.m, np.float64)
        self._call("perc_get_voltage", _ptr(v, C.c_double))
        return v

    # ---- instrumentation
    def set_solver(self, mode):
        """0 = automatic (deflated one-pass iteration kernel whenever it applies), 1 = always the two-kernel form
        (plain Jacobi-PCG, linbcg's iteration count), 2 = one-pass kernel without deflation"""
        self._call("perc_set_solver", _i32(mode))

    def solver_used(self):
        """which iteration kernel the handle's last conductance solve ran: 0 two-kernel form, 1 one-pass, 2 deflated one-pass"""
        f = C.c_int32(0)
        self._call("perc_solver_used", C.byref(f))
        return f.value

    def launch_count(self):
        n = C.c_int64(0)
        self._call("perc_launch_count", C.byref(n))
        return n.value

    def phase_ms(self):
        ms = np.zeros(8, np.float32)
        self._call("perc_phase_ms", _i32(8), _ptr(ms, C.c_float))
        return ms

    def sync(self):
        self._call("perc_sync")


# ---- one lattice decomposed into row slabs over several GPUs ---------------------------------------
IFACE_WORDS = lambda m: 5 * m + 8          # block a rank contributes to the interface all-gather (csrc/slab.h)


def matlab_variable_conductances(conducting, g0=1.0, seed=1838534):
    """The draw of MATLAB/ConductCalc.m with condtype = 2: `rand('twister', 1838534)` (:44-46: MT19937 seeded by
    init_genrand, 53-bit doubles -- numpy's legacy RandomState(seed).random_sample() is the same generator), one draw
    per CONDUCTING bond in bond-list order (:94-96, :117-119, :139-141: G(i,j) = -g0*rand inside the loop over the bonds).
    conducting: bool[nb].  Returns w[nb] for perc_set_bond_conductance (entries of non-conducting bonds are not used)."""
    conducting = np.asarray(conducting, bool)
    w = np.zeros(conducting.size, np.float64)
    w[conducting] = g0 * np.random.RandomState(seed).random_sample(int(conducting.sum()))
    return w


def comm_unique_id():
    """NCCL bootstrap id (128 bytes); rank 0 creates it, the host program distributes it"""
    buf = np.zeros(128, np.uint8)
    _ck("perc_comm_unique_id", load().perc_comm_unique_id(_ptr(buf, C.c_uint8)))
    return buf


def stitch_host(nranks, rank, m, gathered, max_span=4096):
    """the stitch's redundant host union-find on its own (no device needed)"""
    gathered = np.ascontiguousarray(gathered, np.int64)
    assert gathered.size == nranks * IFACE_WORDS(m)
    summary = np.zeros(5, np.int64)
    ids, sizes = np.zeros(max_span, np.int64), np.zeros(max_span, np.int64)
    max_pairs = 2 * m
    pairs = np.zeros(4 * max_pairs, np.int64)
    npairs = C.c_int32(0)
    _ck("perc_stitch_host", load().perc_stitch_host(_i32(nranks), _i32(rank), _i32(m), _ptr(gathered, C.c_int64),
                                                    _ptr(summary, C.c_int64), _i32(max_span), _ptr(ids, C.c_int64),
                                                    _ptr(sizes, C.c_int64), _i32(max_pairs), C.byref(npairs),
                                                    _ptr(pairs, C.c_int64)))
    ns = int(summary[4])
    return dict(ncl=int(summary[0]), nlone=int(summary[1]), maxcs=int(summary[2]), maxcn=int(summary[3]), nspan=ns,
                span_ids=ids[:ns].copy(), span_sizes=sizes[:ns].copy(), pairs=pairs[:4 * npairs.value].reshape(-1, 4).copy())


class SlabLattice(Lattice):
    """rank `rank` of `nranks`: rows [n*rank/nranks, n*(rank+1)/nranks) of one m x n lattice (+ halo rows)"""

    def __init__(self, lattice, m, n, pbc, device, nranks, rank, unique_id=None):
        self.lattice, self.m, self.n, self.pbc = int(lattice), int(m), int(n), int(pbc)
        self.t = self.m * self.n
        self.nranks, self.rank = int(nranks), int(rank)
        self._h = C.c_int64(0)
        self._lib = load()
        _ck("perc_create_slab", self._lib.perc_create_slab(C.byref(self._h), _i32(lattice), _i32(m), _i32(n), _i32(pbc),
                                                           _i32(device), _i32(nranks), _i32(rank)))
        ya, yb = C.c_int32(0), C.c_int32(0)
        self._call("perc_slab_rows", C.byref(ya), C.byref(yb))
        self.ya, self.yb = ya.value, yb.value
        if lattice == SQUARE:
            self.nb = m * (2 * n - 1) if pbc else 2 * m * n - m - n
        else:
            self.nb = m * (3 * n - 2) if pbc else 3 * m * n - 2 * m - 2 * n + 1
        if nranks > 1:
            assert unique_id is not None, "every rank needs the NCCL id created by rank 0 (comm_unique_id)"
            uid = np.ascontiguousarray(unique_id, np.uint8)
            self._call("perc_comm_init", _ptr(uid, C.c_uint8))

    def generate(self, seed, stream=0, ks=-1, kb=-1):
        self._call("perc_generate_i8", _i64(seed), _i64(stream), _i64(ks), _i64(kb))

    def summary(self):
        ncl, maxcs, maxcn, nspan = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self._call("perc_summary_i8", C.byref(ncl), C.byref(maxcs), C.byref(maxcn), C.byref(nspan))
        return dict(ncl=ncl.value, maxcs=maxcs.value, maxcn=maxcn.value, nspan=nspan.value)

    def span(self, max_ids=4096):
        ids, sizes = np.zeros(max_ids, np.int64), np.zeros(max_ids, np.int64)
        nspan = C.c_int32(0)
        self._call("perc_span_i8", _i32(max_ids), C.byref(nspan), _ptr(ids, C.c_int64), _ptr(sizes, C.c_int64))
        k = min(nspan.value, max_ids)
        return ids[:k].copy(), sizes[:k].copy()

    def site_labels(self):
        """lattice-wide canonical labels (int64) of the rows this rank owns"""
        s = np.zeros((self.yb - self.ya) * self.m, np.int64)
        self._call("perc_get_site_labels_i8", _ptr(s, C.c_int64))
        return s
