"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on the
same seeded inputs.  Integer / index results must be bit-exact; conductance must agree
to 1e-9 relative with both sides converged to <= 1e-12."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def P():
    import percolation_b200 as P
    P.load()
    return P


def occ_from(order, k, total):
    occ = np.zeros(total, np.uint8)
    occ[np.asarray(order[:k]) - 1] = 1
    return occ


def bond_occ_from(b1, b2, bo1, bo2, k, t):
    key = b1.astype(np.int64) * (t + 1) + b2
    srt = np.argsort(key)
    q = bo1[:k].astype(np.int64) * (t + 1) + bo2[:k]
    rows = srt[np.searchsorted(key[srt], q)]
    occ = np.zeros(len(b1), np.uint8)
    occ[rows] = 1
    return occ


SHAPES = [(50, 50), (70, 45), (32, 32), (34, 66), (128, 96), (6, 3)]
CASES = [(lat, m, n, pbc) for lat in (1, 2) for (m, n) in SHAPES for pbc in (0, 1)]


@pytest.mark.parametrize("lat,m,n,pbc", CASES)
def test_site_labels_sizes_span(P, O, lat, m, n, pbc):
    t = m * n
    b1, b2 = O.bondlist(lat, m, n, pbc)
    with P.Lattice(lat, m, n, pbc) as L:
        order = O.shuffle_sites(1080115 + m, t)
        for frac in (0.0, 0.3, 0.5, 0.6, 0.75, 1.0):
            k = O.fill_count(frac, t)
            s, c, res = L.site(order, k)
            want, _, wsz, ncl, wmax = O.label_uf(O.SITE, lat, m, n, pbc, b1, b2, site_occ=occ_from(order, k, t))
            assert (s == want).all()
            assert (c == wsz[1:]).all()
            sm = L.summary()
            assert sm["ncl"] == ncl and sm["maxcs"] == wmax
            ids = O.spanning(O.SITE, m, n, b1, b2, want, None)
            gids, gsz = L.span()
            assert list(gids) == list(ids) and list(gsz) == [wsz[i] for i in ids]
            assert res["perccln"] == (ids[0] if len(ids) else 0)
            nbins = 64
            h = L.hist(nbins)
            oh = O.size_hist(O.SITE, t, wsz, None, max(t, nbins))
            assert (h[:nbins - 1] == oh[1:nbins]).all() and h[nbins - 1] == oh[nbins:].sum()
            hl = L.hist_log2(8)                      # log-binned: [2^b, 2^(b+1)), the last bin takes the rest
            sz = wsz[wsz > 0].astype(np.int64)
            want_l = np.bincount(np.minimum(np.floor(np.log2(sz)).astype(int), 7), minlength=8) if len(sz) else np.zeros(8, int)
            assert (hl == want_l).all()


@pytest.mark.parametrize("lat,m,n,pbc", CASES)
def test_bond_labels_sizes_span(P, O, lat, m, n, pbc):
    t = m * n
    b1, b2 = O.bondlist(lat, m, n, pbc)
    nb = len(b1)
    with P.Lattice(lat, m, n, pbc) as L:
        bo1, bo2 = O.shuffle_bonds(184489 + n, b1, b2)
        for frac in (0.0, 0.2, 0.35, 0.5, 0.7, 1.0):
            k = O.fill_count(frac, nb)
            b3, c, res = L.bond(bo1, bo2, k)
            occ = bond_occ_from(b1, b2, bo1, bo2, k, t)
            _, want, wsz, ncl, wmax = O.label_uf(O.BOND, lat, m, n, pbc, b1, b2, bond_occ=occ)
            assert (b3 == want).all()
            assert (c == wsz[1:]).all()
            sm = L.summary()
            assert sm["ncl"] == ncl and sm["maxcs"] == wmax
            ids = O.spanning(O.BOND, m, n, b1, b2, None, want)
            gids, gsz = L.span()
            assert list(gids) == list(ids) and list(gsz) == [wsz[i] for i in ids]
            _, bocc = L.get_occupancy(sites=False)
            assert (bocc == occ).all()


@pytest.mark.parametrize("lat,m,n,pbc", CASES)
def test_mixed_labels_sizes_span(P, O, lat, m, n, pbc):
    t = m * n
    b1, b2 = O.bondlist(lat, m, n, pbc)
    nb = len(b1)
    with P.Lattice(lat, m, n, pbc) as L:
        sorder = O.shuffle_sites(8811064 % 100000 + m, t)
        bo1, bo2 = O.shuffle_bonds(3784068 + n, b1, b2)
        for fs, fb in ((0.0, 0.4), (0.5, 0.5), (0.8, 0.6), (0.8, 0.0), (1.0, 1.0), (0.7, 0.9)):
            ks, kb = O.fill_count(fs, t), O.fill_count(fb, nb)
            s, b3, c, res = L.sitebond(sorder, ks, bo1, bo2, kb)
            socc = occ_from(sorder, ks, t)
            bocc = bond_occ_from(b1, b2, bo1, bo2, kb, t)
            ws, wb, wsz, ncl, wmax = O.label_uf(O.MIXED, lat, m, n, pbc, b1, b2, site_occ=socc, bond_occ=bocc)
            assert (s == ws).all()
            assert (b3 == wb).all()
            assert (c == wsz[1:]).all()
            sm = L.summary()
            assert sm["ncl"] == ncl and sm["maxcs"] == wmax
            ids = O.spanning(O.MIXED, m, n, b1, b2, ws, wb)
            gids, gsz = L.span()
            assert list(gids) == list(ids)
            h = L.hist(32)
            oh = O.size_hist(O.MIXED, t, wsz, wb, max(t + nb, 32))
            assert (h[:31] == oh[1:32]).all() and h[31] == oh[32:].sum()


def test_png_fixtures_through_gpu(P, O):
    """the reference's own renderings: largest cluster (green) == GPU max cluster"""
    z = np.load(os.path.join(HERE, "golden", "png_fixtures.npz"))
    M = N = 50
    handles = {lat: P.Lattice(lat, M, N, 0) for lat in (1, 2)}
    lists = {lat: O.bondlist(lat, M, N, 0) for lat in (1, 2)}
    checked = 0
    for row in z["index"]:
        key, kind, lat, p, name = str(row).split("|")
        kind, lat = int(kind), int(lat)
        dots = np.unpackbits(z[key + "_dots"])[: M * N]
        cls = z[key + "_cls"]
        L = handles[lat]
        b1, b2 = lists[lat]
        if kind == O.SITE:
            L.set_occupancy(socc=dots)
            L.label(P.SITE)
            s = L.site_labels()
            lab_of_bond = s[b1 - 1]
            occb = (dots[b1 - 1] == 1) & (dots[b2 - 1] == 1)
        else:
            L.set_occupancy(bocc=(cls > 0).astype(np.uint8))
            L.label(P.BOND)
            lab_of_bond = L.bond_labels()
            occb = cls > 0
        assert ((lab_of_bond > 0) & occb == occb).all(), name
        c = L.sizes()
        sm = L.summary()
        green = cls == 2
        if green.any():
            labs = np.unique(lab_of_bond[green])
            assert all(c[l - 1] == sm["maxcs"] for l in labs), name
            assert (green == (occb & np.isin(lab_of_bond, labs))).all(), name
        checked += 1
    assert checked == 65
    for L in handles.values():
        L.close()


# m = 40: scalar kernels (stored q); m % 16 == 0: staged kernels (cp.async tiles, q recomputed), incl.
# partial tiles (m = 48, 144) and several tiles per row / column (144 x 70)
@pytest.mark.parametrize("m,n", [(40, 36), (48, 36), (144, 70)])
@pytest.mark.parametrize("lat,kind", [(1, 2), (2, 2), (1, 1), (2, 1), (1, 3), (2, 3)])
@pytest.mark.parametrize("pbc", [0, 1])
def test_conductance_vs_oracle(P, O, lat, kind, pbc, m, n):
    t = m * n
    b1, b2 = O.bondlist(lat, m, n, pbc)
    nb = len(b1)
    done = 0
    with P.Lattice(lat, m, n, pbc) as L:
        for seed in range(5):
            sorder = O.shuffle_sites(626504 + seed, t)
            bo1, bo2 = O.shuffle_bonds(184489 + seed, b1, b2)
            if kind == O.BOND:
                kb = O.fill_count(0.54 if lat == 1 else 0.38, nb)
                b3, c, res = L.bond(bo1, bo2, kb)
                s = None
            elif kind == O.SITE:
                ks = O.fill_count(0.63 if lat == 1 else 0.54, t)
                s, c, res = L.site(sorder, ks)
                b3 = None
            else:
                ks, kb = O.fill_count(0.85, t), O.fill_count(0.68 if lat == 1 else 0.5, nb)
                s, b3, c, res = L.sitebond(sorder, ks, bo1, bo2, kb)
            if not res["perccln"]:
                continue
            cid = res["perccln"]
            w = O.weights(kind, b1, b2, s, b3, cid)
            # (a) reference defaults: same recurrences -> same iteration count, same G to rounding
            ref = O.conduct_literal(m, n, b1, b2, w)
            got = L.conduct(cid)
            # (at tol = 1e-8 the iterate itself carries ~1e-7 relative error -- SURVEY F6 -- so two
            # correct implementations differing in summation order agree only to about that)
            assert abs(got["iter"] - ref["iter"]) <= 1
            assert abs(got["Gtop"] - ref["Gtop"]) <= 1e-7 * abs(ref["Gtop"])
            assert abs(got["Gbot"] - ref["Gbot"]) <= 1e-7 * abs(ref["Gbot"])
            # (b) both converged: 1e-9 relative (tolerance of the north star), in fact ~1e-12
            ref = O.conduct_cg(m, n, b1, b2, w, tol=1e-13, itmax=200000)
            got = L.conduct(cid, tol=1e-13, itmax=200000)
            assert got["err"] <= 1e-13
            assert abs(got["Gtop"] - ref["Gtop"]) <= 1e-9 * abs(ref["Gtop"])
            assert abs(got["Gbot"] - ref["Gbot"]) <= 1e-9 * abs(ref["Gbot"])
            # (c) independent residual check of the GPU voltages on the CPU
            chk = O.conduct_check(m, n, b1, b2, w, L.voltage())
            assert chk["err"] <= 2e-13
            assert abs(chk["Gtop"] - got["Gtop"]) <= 1e-12 * abs(got["Gtop"])
            # (d) the sweep-driver entry point (no interior voltages), two-kernel form: bit-identical G, iter, err
            L.set_solver(1)
            gonly = L.conduct(cid, tol=1e-13, itmax=200000, voltages=False)
            assert gonly == got and L.solver_used() == 0
            # (e) ... and with the one-pass iteration kernel where it applies (same recurrences in the
            # Chronopoulos-Gear arrangement): 1e-9 against the oracle, linbcg's iteration count
            applies = pbc == 0 and m % 16 == 0
            L.set_solver(2)
            gf = L.conduct(cid, tol=1e-13, itmax=200000, voltages=False)
            assert L.solver_used() == (1 if applies else 0)
            assert gf["err"] <= 1e-13
            assert abs(gf["Gtop"] - ref["Gtop"]) <= 1e-9 * abs(ref["Gtop"])
            assert abs(gf["Gbot"] - ref["Gbot"]) <= 1e-9 * abs(ref["Gbot"])
            assert abs(gf["iter"] - got["iter"]) <= max(3, got["iter"] // 100)
            # (f) the default: the DEFLATED one-pass kernel -- same x to the tolerance, fewer iterations
            L.set_solver(0)
            gdf = L.conduct(cid, tol=1e-13, itmax=200000, voltages=False)
            assert L.solver_used() == (2 if applies else 0)
            assert gdf["err"] <= 1e-13
            assert abs(gdf["Gtop"] - ref["Gtop"]) <= 1e-9 * abs(ref["Gtop"])
            assert abs(gdf["Gbot"] - ref["Gbot"]) <= 1e-9 * abs(ref["Gbot"])
            assert gdf["iter"] <= got["iter"] + 3
            L.set_solver(2)
            gd = L.conduct(cid, voltages=False)             # reference defaults 1e-8 / 2500
            refd = O.conduct_literal(m, n, b1, b2, w)
            assert abs(gd["iter"] - refd["iter"]) <= 1
            assert abs(gd["Gtop"] - refd["Gtop"]) <= 1e-7 * abs(refd["Gtop"])
            with pytest.raises(P.PercError) as e:
                L.voltage()
            assert e.value.code == P.E_STATE
            L.set_solver(0)
            done += 1
            if done == 2:
                break
    assert done >= 1


# ---- per-bond conductances (MATLAB/ConductCalc.m condtype = 2) -----------------------------------
@pytest.mark.parametrize("m,n", [(40, 36), (64, 50)])
@pytest.mark.parametrize("lat,kind", [(1, 2), (2, 2), (1, 1), (2, 1), (1, 3), (2, 3)])
@pytest.mark.parametrize("pbc", [0, 1])
def test_variable_bond_conductance_vs_oracle(P, O, lat, kind, pbc, m, n):
    """conducting bonds get g0 * rand drawn as ConductCalc.m draws it (twister seed 1838534, one draw per conducting bond
    in bond order, :44-46, :94-96); Gtop / Gbot / voltages against the oracle's CG on the same weights: 1e-9"""
    t = m * n
    b1, b2 = O.bondlist(lat, m, n, pbc)
    nb = len(b1)
    done = 0
    with P.Lattice(lat, m, n, pbc) as L:
        for seed in range(5):
            sorder = O.shuffle_sites(626504 + seed, t)
            bo1, bo2 = O.shuffle_bonds(184489 + seed, b1, b2)
            if kind == O.BOND:
                b3, c, res = L.bond(bo1, bo2, O.fill_count(0.54 if lat == 1 else 0.38, nb)); s = None
            elif kind == O.SITE:
                s, c, res = L.site(sorder, O.fill_count(0.63 if lat == 1 else 0.54, t)); b3 = None
            else:
                s, b3, c, res = L.sitebond(sorder, O.fill_count(0.85, t), bo1, bo2, O.fill_count(0.68 if lat == 1 else 0.5, nb))
            if not res["perccln"]:
                continue
            cid = res["perccln"]
            w0 = O.weights(kind, b1, b2, s, b3, cid, g0=1.0, gleak=1e-12)     # 1.0 on the bonds that conduct
            wv = P.matlab_variable_conductances(w0 == 1.0, g0=1.0)
            w = np.where(w0 == 1.0, wv, 1e-12)
            uniform = L.conduct(cid, tol=1e-13, itmax=400000)
            L.set_bond_conductance(wv)
            got = L.conduct(cid, tol=1e-13, itmax=400000)
            ref = O.conduct_cg(m, n, b1, b2, w, tol=1e-13, itmax=400000)
            assert got["err"] <= 1e-13 and abs(got["iter"] - ref["iter"]) <= max(3, ref["iter"] // 100)
            assert abs(got["Gtop"] - ref["Gtop"]) <= 1e-9 * abs(ref["Gtop"])
            assert abs(got["Gbot"] - ref["Gbot"]) <= 1e-9 * abs(ref["Gbot"])
            assert got["Gtop"] < uniform["Gtop"]                                 # every conducting bond got weaker
            chk = O.conduct_check(m, n, b1, b2, w, L.voltage())
            assert chk["err"] <= 2e-13 and abs(chk["Gtop"] - got["Gtop"]) <= 1e-12 * abs(got["Gtop"])
            gonly = L.conduct(cid, tol=1e-13, itmax=400000, voltages=False)     # the sweep drivers' entry point: same solve
            assert gonly["Gtop"] == got["Gtop"] and gonly["Gbot"] == got["Gbot"]
            # the reference's own tolerance / iteration cap against the literal linbcg
            lit = O.conduct_literal(m, n, b1, b2, w)
            gd = L.conduct(cid)
            assert abs(gd["iter"] - lit["iter"]) <= 1 and abs(gd["Gtop"] - lit["Gtop"]) <= 1e-6 * abs(lit["Gtop"])
            # all weights equal to g0: the uniform solve, and dropping the weights restores it exactly
            L.set_bond_conductance(np.ones(nb))
            same = L.conduct(cid, tol=1e-13, itmax=400000)
            assert abs(same["Gtop"] - uniform["Gtop"]) <= 1e-9 * abs(uniform["Gtop"])
            L.set_bond_conductance(None)
            again = L.conduct(cid, tol=1e-13, itmax=400000)
            assert again == uniform
            done += 1
            if done == 2:
                break
    assert done >= 1


def test_full_lattice_closed_form(P, O):
    """every bond occupied: square G = m/(n-1) exactly"""
    for (m, n, pbc) in ((10, 10, 0), (16, 8, 1), (50, 50, 0)):
        with P.Lattice(1, m, n, pbc) as L:
            L.set_occupancy(bocc=np.ones(L.nb, np.uint8))
            L.label(P.BOND)
            r = L.conduct(0, tol=1e-13, itmax=100000)
            assert abs(r["Gtop"] - m / (n - 1)) < 1e-10 and abs(r["Gbot"] - m / (n - 1)) < 1e-10
    for (m, n, pbc, want) in ((10, 10, 0, 1.662318388765), (16, 8, 1, 3.576045037457)):
        with P.Lattice(2, m, n, pbc) as L:
            L.set_occupancy(socc=np.ones(L.t, np.uint8))
            L.label(P.SITE)
            r = L.conduct(0, tol=1e-13, itmax=100000)
            assert abs(r["Gtop"] - want) < 2e-11 and abs(r["Gbot"] - want) < 2e-11


@pytest.mark.parametrize("lat,kind,m,n,ps,pb", [(1, 3, 528, 101, 0.85, 0.70), (2, 1, 400, 70, 0.56, 0.0), (1, 2, 16, 130, 0.0, 0.75),
                                                   (1, 3, 1024, 1024, 0.80, 0.70), (2, 2, 1024, 512, 0.0, 0.37)])
def test_one_pass_solver_equals_two_kernel_solver(P, O, lat, kind, m, n, ps, pb):
    """perc_conduct_g with the one-pass iteration kernel (pcg_fused_kernel: tiles with and without the
    geometry-free fast path, partial tiles, both lattices) against the two-kernel form on the same handle;
    the small shapes also against the oracle"""
    with P.Lattice(lat, m, n, 0) as L:
        t, nb = L.t, L.nb
        found = 0
        for stream in range(6):
            L.generate(4711, stream, int(ps * t) if kind != 2 else -1, int(pb * nb) if kind != 1 else -1)
            L.label(kind)
            ids, _ = L.span()
            if not len(ids):
                continue
            found += 1
            for tol, itmax in ((1e-13, 2000000), (1e-8, 2500)):
                L.set_solver(1)
                a = L.conduct(0, tol=tol, itmax=itmax, voltages=False)
                assert L.solver_used() == 0
                # 2 = the one-pass kernel (12 names its variant explicitly); 10 = its first variant; 0 / 14 = deflated
                for mode in (2, 10, 12, 0, 14):
                    L.set_solver(mode)
                    b = L.conduct(0, tol=tol, itmax=itmax, voltages=False)
                    defl = mode in (0, 14)
                    assert L.solver_used() == (2 if defl else 1)
                    rel = 1e-9 if tol < 1e-10 else 1e-6
                    if a["iter"] > itmax:
                        # stopped by itmax, not converged: the two arrangements of the recurrences drift apart
                        # like any two finite-precision CG runs; they agree to about the residual they stopped at
                        # (the deflated solve is a different, faster converging sequence: only its own residual is checked)
                        rel = max(rel, a["err"])
                        if defl:
                            assert b["err"] <= a["err"] * 1.01, (mode, tol, a, b)
                            continue
                    assert abs(a["Gtop"] - b["Gtop"]) <= rel * abs(a["Gtop"]), (mode, tol, a, b)
                    assert abs(a["Gbot"] - b["Gbot"]) <= rel * abs(a["Gbot"]), (mode, tol, a, b)
                    if defl:
                        assert b["iter"] <= a["iter"] + 3, (mode, tol, a, b)
                    else:
                        assert abs(a["iter"] - b["iter"]) <= max(3, a["iter"] // 100), (mode, tol, a, b)
                    if a["iter"] <= itmax:
                        assert b["err"] <= tol
            L.set_solver(0)
            if t <= 60000:
                b1, b2 = O.bondlist(lat, m, n, 0)
                socc, bocc = L.get_occupancy(sites=kind != 2, bonds=kind != 1)
                ws, wb, wsz, ncl, wmax = O.label_uf(kind, lat, m, n, 0, b1, b2, site_occ=socc, bond_occ=bocc)
                w = O.weights(kind, b1, b2, ws, wb, int(ids[0]))
                ref = O.conduct_cg(m, n, b1, b2, w, tol=1e-13, itmax=2000000)
                b = L.conduct(0, tol=1e-13, itmax=2000000, voltages=False)
                assert abs(b["Gtop"] - ref["Gtop"]) <= 1e-9 * abs(ref["Gtop"])
                assert abs(b["Gbot"] - ref["Gbot"]) <= 1e-9 * abs(ref["Gbot"])
            if found == 2 or t > 60000:
                break
        assert found >= 1


@pytest.mark.parametrize("lat,kind,m,n,ps,pb", [(1, 3, 128, 60, 0.85, 0.70), (1, 2, 256, 101, 0.0, 0.55), (2, 1, 256, 70, 0.56, 0.0),
                                                   (2, 3, 384, 66, 0.8, 0.6), (1, 3, 1024, 256, 0.80, 0.70)])
def test_one_pass_solver_with_periodic_wrap(P, O, lat, kind, m, n, ps, pb):
    """pbc = 1 with the seam on a tile border (m a multiple of 128): perc_conduct_g runs the one-pass kernels (the halo columns
    beyond the seam are patched into the TMA boxes; in the deflated sweep the coarse operator, the crossing currents and the
    mu records wrap as well: the last block column is the west neighbour of the first).  Against the two-kernel form on the
    same handle and, for the small shapes, the oracle."""
    with P.Lattice(lat, m, n, 1) as L:
        t, nb = L.t, L.nb
        found = 0
        for stream in range(6):
            L.generate(4712, stream, int(ps * t) if kind != 2 else -1, int(pb * nb) if kind != 1 else -1)
            L.label(kind)
            ids, _ = L.span()
            if not len(ids):
                continue
            found += 1
            for tol, itmax in ((1e-13, 2000000), (1e-8, 2500)):
                L.set_solver(1)
                a = L.conduct(0, tol=tol, itmax=itmax, voltages=False)
                assert L.solver_used() == 0
                for mode in (2, 0):
                    L.set_solver(mode)
                    b = L.conduct(0, tol=tol, itmax=itmax, voltages=False)
                    assert L.solver_used() == (1 if mode == 2 else 2)
                    if mode == 0:
                        # the deflated sweep: never more iterations than plain Jacobi-PCG; G compared when both converged
                        assert b["iter"] <= a["iter"] + 3, (tol, a, b)
                        if a["iter"] > itmax:
                            continue
                    rel = 1e-9 if tol < 1e-10 else 1e-6
                    if a["iter"] > itmax:
                        rel = max(rel, a["err"])
                    assert abs(a["Gtop"] - b["Gtop"]) <= rel * abs(a["Gtop"]), (mode, tol, a, b)
                    assert abs(a["Gbot"] - b["Gbot"]) <= rel * abs(a["Gbot"]), (mode, tol, a, b)
                    if mode == 2:
                        assert abs(a["iter"] - b["iter"]) <= max(3, a["iter"] // 100), (mode, tol, a, b)
            if t <= 60000:
                b1, b2 = O.bondlist(lat, m, n, 1)
                socc, bocc = L.get_occupancy(sites=kind != 2, bonds=kind != 1)
                ws, wb, wsz, ncl, wmax = O.label_uf(kind, lat, m, n, 1, b1, b2, site_occ=socc, bond_occ=bocc)
                w = O.weights(kind, b1, b2, ws, wb, int(ids[0]))
                ref = O.conduct_cg(m, n, b1, b2, w, tol=1e-13, itmax=2000000)
                L.set_solver(0)
                b = L.conduct(0, tol=1e-13, itmax=2000000, voltages=False)
                assert abs(b["Gtop"] - ref["Gtop"]) <= 1e-9 * abs(ref["Gtop"])
                assert abs(b["Gbot"] - ref["Gbot"]) <= 1e-9 * abs(ref["Gbot"])
            if found == 2 or t > 60000:
                break
        assert found >= 1


def test_first_span_matches_literal_fill(P, O):
    for lat in (1, 2):
        m = n = 24
        t = m * n
        b1, b2 = O.bondlist(lat, m, n, 0)
        seeds = O.seed_table(58302, 6, 1000000)              # Sq/site_perc.f:69-75
        with P.Lattice(lat, m, n, 0) as L:
            for sd in seeds:
                order = O.shuffle_sites(int(sd), t)
                _, c, res = O.site_literal(lat, m, n, 0, order, t, stop_at_span=True)
                L.set_site_order(order)
                got = L.first_span(P.SITE, P.SITE)
                assert got["kstar"] == res["filled"]
                assert got["f"] == O.fraction(res["filled"], t)
                assert got["maxcs"] == res["maxcs"] and got["perccls"] == res["perccls"]
                bo1, bo2 = O.shuffle_bonds(int(sd), b1, b2)
                _, c, res = O.bond_literal(lat, m, n, 0, b1, b2, bo1, bo2, len(b1), stop_at_span=True)
                L.set_bond_order(bo1, bo2)
                got = L.first_span(P.BOND, P.BOND)
                assert got["kstar"] == res["filled"] and got["maxcs"] == res["maxcs"]
                # mixed, sites fixed at ps = 0.8, bonds added (Sq/sb_perc.f)
                ks = O.fill_count(0.8, t)
                s, b3, c, res = O.sitebond_literal(lat, m, n, 0, b1, b2, order, ks, bo1, bo2, len(b1), stop_at_span=True)
                L.set_fill(ks=ks)
                got = L.first_span(P.MIXED, P.BOND)
                assert got["kstar"] == (res["filled"] if res["perccln"] else 0)


# ---- K1 generator -------------------------------------------------------------------------
from philox_np import philox_pairs, site_keys, bond_keys_and_ids  # noqa: E402  (numpy restatement of csrc/philox.cuh + K1 selection)


def test_philox_known_answer():
    # Random123 KAT: counter 0, key 0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8
    A, B = philox_pairs(0, 0, 0, [0])
    assert int(A[0]) == (0x6627e8d5 << 32) | 0xe169c58d and int(B[0]) == (0xbc57ac4c << 32) | 0x9b00dbd8


@pytest.mark.parametrize("lat,m,n,pbc", [(1, 50, 50, 0), (2, 64, 40, 1), (1, 257, 129, 1)])
def test_generator_exact_count_and_keys(P, lat, m, n, pbc):
    seed, stream = 20240611, 3
    with P.Lattice(lat, m, n, pbc) as L:
        t, nb = L.t, L.nb
        b1, b2 = P.geom_bondlist(lat, m, n, pbc)
        for fs, fb in ((0.6, 0.35), (0.0, 1.0), (1.0, 0.0), (0.5927, 0.5)):
            ks, kb = int(fs * t), int(fb * nb)
            L.generate(seed, stream, ks, kb)
            socc, bocc = L.get_occupancy()
            assert int(socc.sum()) == ks and int(bocc.sum()) == kb
            keys = site_keys(seed, stream, t)
            order = np.lexsort((np.arange(t), keys))
            want = np.zeros(t, np.uint8)
            want[order[:ks]] = 1
            assert (socc == want).all()
            bkey, bid = bond_keys_and_ids(seed, stream, lat, m, n, pbc, b1, b2)
            order = np.lexsort((bid, bkey))
            want = np.zeros(nb, np.uint8)
            want[order[:kb]] = 1
            assert (bocc == want).all()
        # nesting: a larger fill count keeps every previously occupied element (sweep property)
        L.generate(seed, stream, int(0.5 * t), int(0.4 * nb))
        s1, bb1 = L.get_occupancy()
        L.set_fill(int(0.6 * t), int(0.5 * nb))
        s2, bb2 = L.get_occupancy()
        assert (s2 >= s1).all() and (bb2 >= bb1).all()
        assert int(s2.sum()) == int(0.6 * t) and int(bb2.sum()) == int(0.5 * nb)
        # different realizations differ
        L.generate(seed, stream + 1, int(0.5 * t), -1)
        s3, _ = L.get_occupancy(bonds=False)
        assert (s3 != s1).any()


# ---- larger lattices -----------------------------------------------------------------------
@pytest.mark.parametrize("lat,kind,L_", [(1, 1, 1024), (2, 2, 1024), (1, 3, 1024), (2, 1, 2048)])
def test_large_lattice_vs_oracle(P, O, lat, kind, L_):
    m = n = L_
    t = m * n
    b1, b2 = O.bondlist(lat, m, n, 0)
    nb = len(b1)
    with P.Lattice(lat, m, n, 0) as L:
        ks = int((0.5927 if lat == 1 else 0.5) * t) if kind == 1 else int(0.8 * t)
        kb = int((0.5 if lat == 1 else 0.3473) * nb) if kind == 2 else int(0.66 * nb)
        L.generate(777, 0, ks if kind != 2 else -1, kb if kind != 1 else -1)
        L.label(kind)
        socc, bocc = L.get_occupancy(sites=kind != 2, bonds=kind != 1)
        ws, wb, wsz, ncl, wmax = O.label_uf(kind, lat, m, n, 0, b1, b2, site_occ=socc, bond_occ=bocc)
        if kind != 2:
            assert (L.site_labels() == ws).all()
        if kind != 1:
            assert (L.bond_labels() == wb).all()
        assert (L.sizes() == wsz[1:]).all()
        sm = L.summary()
        assert sm["ncl"] == ncl and sm["maxcs"] == wmax
        ids = O.spanning(kind, m, n, b1, b2, ws, wb)
        assert list(L.span()[0]) == list(ids)


def test_full_size_properties_L4096(P):
    """BASELINE config sizes: size-independent properties (sum of sizes == occupied count,
    labels are fixed points, idempotence, spanning monotone in the fill)"""
    m = n = 4096
    with P.Lattice(1, m, n, 0) as L:
        t, nb = L.t, L.nb
        ks, kb = int(0.8 * t), int(0.72 * nb)
        L.generate(12345, 0, ks, kb)
        L.label(P.MIXED)
        s = L.site_labels()
        c = L.sizes()
        sm = L.summary()
        occ = s > 0
        assert int(occ.sum()) == ks
        roots = np.nonzero(c > 0)[0]
        assert (s[roots] == roots + 1).all()                    # canonical label = own id at the root
        assert (s[occ] <= np.nonzero(occ)[0] + 1).all()          # label is the minimum member id
        assert (c[s[occ] - 1] > 0).all()
        h = L.hist(8)
        nlone = sm["ncl"] - len(roots)
        assert int(c.astype(np.int64).sum()) + nlone == ks + kb  # every element in exactly one cluster
        assert h.sum() == sm["ncl"]
        L.label(P.MIXED)                                         # idempotent
        assert (L.site_labels() == s).all() and (L.sizes() == c).all()
        sp_hi = L.summary()["nspan"]
        L.set_fill(kb=int(0.45 * nb))
        L.label(P.MIXED)
        assert L.summary()["nspan"] == 0 and sp_hi >= 1


def test_c_driver_full_path(P, tmp_path):
    """Fortran-convention C driver: enumerate, shuffle, perc_bond, perc_conduct"""
    import subprocess
    from test_abi_cpu import _build_c_driver
    exe = _build_c_driver(str(tmp_path))
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("OK nb="), out.stdout + out.stderr


# ---- batches of realizations (device-resident selection + statistics) ------------------------
@pytest.mark.parametrize("lat,kind,m,n,pbc,ps,pb", [(1, 1, 100, 100, 0, 0.60, 0.0), (2, 2, 128, 96, 1, 0.0, 0.35),
                                                   (1, 3, 144, 80, 0, 0.8, 0.62), (2, 1, 64, 64, 0, 0.5, 0.0)])
def test_batch_equals_realization_loop(P, lat, kind, m, n, pbc, ps, pb):
    """perc_batch (no host synchronisation inside) == the same realizations one call at a time"""
    nreal, nbins, seed, stream0 = 7, 48, 58302, 11
    with P.Lattice(lat, m, n, pbc) as L:
        ks, kb = int(ps * L.t), int(pb * L.nb)
        hist = np.zeros(nbins, np.int64)
        want = dict(realizations=0, sum_ncl=0, sum_maxcs=0, spanning=0, sum_nspan=0, sum_perccls=0, failed=0, sum_maxcs2=0,
                    sum_sites=0, sum_bonds=0)
        for i in range(nreal):
            L.generate(seed, stream0 + i, ks if kind != 2 else -1, kb if kind != 1 else -1)
            L.label(kind)
            hist += L.hist(nbins)
            sm = L.summary()
            ids, sizes = L.span()
            want["realizations"] += 1
            want["sum_ncl"] += sm["ncl"]
            want["sum_maxcs"] += sm["maxcs"]
            want["sum_maxcs2"] += sm["maxcs"] ** 2
            want["spanning"] += len(ids) > 0
            want["sum_nspan"] += len(ids)
            want["sum_perccls"] += int(sizes[0]) if len(ids) else 0
            want["sum_sites"] += ks if kind != 2 else 0
            want["sum_bonds"] += kb if kind != 1 else 0
        got_hist, got = L.batch(kind, nreal, seed, stream0, ks, kb, nbins)
        assert got == want
        assert (got_hist == hist).all()
        # a batch consumes the handle's occupancy input ...
        with pytest.raises(P.PercError) as e:
            L.label(kind)
        assert e.value.code == P.E_STATE
        # ... and the handle is usable as before once it gets a new one
        L.generate(seed, stream0, ks if kind != 2 else -1, kb if kind != 1 else -1)
        L.label(kind)
        assert L.summary()["ncl"] > 0


@pytest.mark.parametrize("lat,kind,m,n,pbc,ps,pb", [(1, 1, 100, 100, 0, 0.62, 0.0), (2, 2, 64, 48, 1, 0.0, 0.40),
                                                   (1, 3, 96, 60, 0, 0.85, 0.68), (1, 2, 144, 128, 0, 0.0, 0.53)])
def test_batch_conduct_equals_realization_loop(P, lat, kind, m, n, pbc, ps, pb):
    """perc_batch_conduct (small lattices: one CTA per realization, the whole solve in one launch; the last
    case is too large for that and goes realization by realization) == perc_generate + perc_label +
    perc_conduct_g one realization at a time, both converged to 1e-13"""
    nreal, seed, stream0 = 6, 626504, 3
    with P.Lattice(lat, m, n, pbc) as L:
        ks, kb = int(ps * L.t), int(pb * L.nb)
        want, spans = [], 0
        for i in range(nreal):
            L.generate(seed, stream0 + i, ks if kind != 2 else -1, kb if kind != 1 else -1)
            L.label(kind)
            L.set_solver(2)                         # plain Jacobi-PCG, as the one-CTA batch solver runs it (same iteration counts)
            if len(L.span()[0]):
                r = L.conduct(0, tol=1e-13, itmax=200000, voltages=False)
                want.append((r["Gtop"], r["Gbot"], r["iter"]))
                spans += 1
            else:
                want.append((0.0, 0.0, -1))
        G, iters, st = L.batch_conduct(kind, nreal, seed, stream0, ks, kb, tol=1e-13, itmax=200000)
        assert st["realizations"] == nreal and st["spanning"] == spans and st["failed"] == 0
        assert spans >= 1
        for i, (gt, gb, it) in enumerate(want):
            if it < 0:
                assert iters[i] == -1 and G[i, 0] == 0.0 and G[i, 1] == 0.0
            else:
                assert abs(G[i, 0] - gt) <= 1e-9 * abs(gt) and abs(G[i, 1] - gb) <= 1e-9 * abs(gb), (i, G[i], gt, gb)
                assert abs(int(iters[i]) - it) <= max(3, it // 50)


@pytest.mark.parametrize("lat,kind,m,n", [(1, 2, 64, 48), (1, 3, 144, 80), (2, 2, 50, 40)])
def test_warm_start_sweep(P, lat, kind, m, n):
    """p-sweep on one bond order (Sq/bond_cond.f:208-485): perc_conduct_warm starts from the previous point's
    voltages and must reach the same converged conductance as a cold start"""
    with P.Lattice(lat, m, n, 0) as L:
        ks = int(0.85 * L.t) if kind == 3 else -1
        p0 = 0.54 if lat == 1 and kind == 2 else 0.70 if kind == 3 else 0.38
        fills = [int((p0 + 0.005 * j) * L.nb) for j in range(5)]
        L.generate(31337, 0, ks, fills[0])
        cold, warm = [], []
        for kb in fills:
            L.set_fill(kb=kb)
            L.label(kind)
            assert len(L.span()[0]) >= 1
            cold.append(L.conduct(0, tol=1e-13, itmax=400000))
        for j, kb in enumerate(fills):
            L.set_fill(kb=kb)
            L.label(kind)
            warm.append(L.conduct(0, tol=1e-13, itmax=400000, warm=j > 0))
        for c, w in zip(cold, warm):
            assert abs(c["Gtop"] - w["Gtop"]) <= 1e-9 * abs(c["Gtop"]) and abs(c["Gbot"] - w["Gbot"]) <= 1e-9 * abs(c["Gbot"])
            assert w["err"] <= 1e-13
        assert sum(w["iter"] for w in warm[1:]) < sum(c["iter"] for c in cold[1:])


# ---- re-labeling along a sweep ------------------------------------------------------------------
@pytest.mark.parametrize("lat,kind,which,m,n,pbc", [(1, 1, 1, 70, 45, 0), (2, 1, 1, 64, 40, 1), (1, 2, 2, 130, 70, 0), (2, 2, 2, 48, 36, 1),
                                                       (1, 3, 2, 144, 70, 0), (2, 3, 2, 160, 96, 1), (1, 3, 1, 50, 41, 1), (2, 3, 1, 64, 50, 0),
                                                       (1, 3, 2, 1024, 512, 0)])
@pytest.mark.parametrize("source", ["order", "generator"])
def test_incremental_labeling_equals_full_labeling(P, O, lat, kind, which, m, n, pbc, source):
    """perc_label_incremental along a sweep (elements of one kind added at rising fill, the other kind fixed) against perc_label
    at every sweep point on a second handle: labels, bond labels, sizes, counts, spanning clusters bit-identical; the small
    lattices also against the oracle"""
    t = m * n
    b1, b2 = O.bondlist(lat, m, n, pbc)
    nb = len(b1)
    with P.Lattice(lat, m, n, pbc) as A, P.Lattice(lat, m, n, pbc) as B:
        for L in (A, B):
            if source == "order" and t <= 100000:
                L.set_site_order(O.shuffle_sites(626504, t))
                L.set_bond_order(*O.shuffle_bonds(184489, b1, b2))
            elif source == "order":                          # (the reference's float32 shuffle is for its own sizes, t <= 10^6)
                rng = np.random.default_rng(5)
                L.set_site_order(rng.permutation(t).astype(np.int32) + 1)
                bp = rng.permutation(nb)
                L.set_bond_order(b1[bp], b2[bp])
            else:
                L.generate(99, 7, t, nb)
        # sweep of the `which` elements from empty to full in uneven steps; the other kind fixed at 0.8 (mixed)
        N = t if which == 1 else nb
        fills = sorted(set([0, 1, 2, int(0.1 * N), int(0.3 * N), int(0.45 * N), int(0.5 * N), int(0.5 * N) + 1, int(0.55 * N),
                            int(0.6 * N), int(0.75 * N), int(0.9 * N), N - 1, N]))
        ks0, kb0 = (-1, int(0.8 * nb)) if which == 1 else (int(0.8 * t), -1)
        nin = 0
        for k in fills:
            ks, kb = (k, kb0) if which == 1 else (ks0, k)
            for L in (A, B):
                L.set_fill(ks if kind != 2 else -1, kb if kind != 1 else -1)
            B.label(kind)
            nin += A.label_incremental(kind)
            assert A.summary() == B.summary(), (k, A.summary(), B.summary())
            sa, sb = A.span(), B.span()
            assert list(sa[0]) == list(sb[0]) and list(sa[1]) == list(sb[1])
            if kind != 2:
                assert (A.site_labels() == B.site_labels()).all()
            if kind != 1:
                assert (A.bond_labels() == B.bond_labels()).all()
            assert (A.sizes() == B.sizes()).all()
            assert (A.hist(16) == B.hist(16)).all()
        assert nin == len(fills) - 1                        # everything after the first point was incremental
        if t <= 20000:
            socc, bocc = A.get_occupancy(sites=kind != 2, bonds=kind != 1)
            ws, wb, wsz, ncl, wmax = O.label_uf(kind, lat, m, n, pbc, b1, b2, site_occ=socc, bond_occ=bocc)
            if kind != 2:
                assert (A.site_labels() == ws).all()
            assert (A.sizes() == wsz[1:]).all()
        # a smaller fill, or another order, takes the full pass again
        A.set_fill(0 if kind != 2 else -1, 0 if kind != 1 else -1)
        assert A.label_incremental(kind) is False


# ---- the reference programs' output files -------------------------------------------------------
def _read_txt(path, ncol):
    lines = open(path).read().split("\n")
    assert lines[-1] == ""
    rows = lines[:-1]
    assert all(len(r) == 11 * ncol - 1 and r.count(",") == ncol - 1 for r in rows)      # i10 fields joined by commas
    return np.array([[int(f) for f in r.split(",")] for r in rows], np.int64)


def test_reference_output_files(P, O, tmp_path):
    """perc_write_txt: site.txt / bond.txt / sbsite.txt / sbbond.txt / bondlist.txt in the reference's record formats
    (Sq/site.f:354-359, Sq/bond.f:443-448, Sq/sitebond.f:469-477), contents = the getters' arrays = the oracle's"""
    for lat in (1, 2):
        m, n = 24, 20
        t = m * n
        b1, b2 = O.bondlist(lat, m, n, 0)
        nb = len(b1)
        sorder = O.shuffle_sites(1080115, t)
        bo1, bo2 = O.shuffle_bonds(184489, b1, b2)
        with P.Lattice(lat, m, n, 0) as L:
            s, c, res = L.site(sorder, O.fill_count(0.6, t))
            L.write_txt("site.txt", str(tmp_path / "site.txt"))
            a = _read_txt(tmp_path / "site.txt", 3)
            assert (a[:, 0] == np.arange(1, t + 1)).all() and (a[:, 1] == s).all() and (a[:, 2] == c).all()
            socc = np.zeros(t, np.uint8); socc[sorder[:O.fill_count(0.6, t)] - 1] = 1
            ws, wb, wsz, ncl, wmax = O.label_uf(O.SITE, lat, m, n, 0, b1, b2, site_occ=socc)
            assert (a[:, 1] == ws).all() and (a[:, 2] == wsz[1:]).all()
            with pytest.raises(P.PercError):
                L.write_txt("bond.txt", str(tmp_path / "x.txt"))                 # no bond labeling on the handle
            L.write_txt("bondlist.txt", str(tmp_path / "bondlist.txt"))
            a = _read_txt(tmp_path / "bondlist.txt", 2)
            assert (a[:, 0] == b1).all() and (a[:, 1] == b2).all()
            b3, c, res = L.bond(bo1, bo2, O.fill_count(0.5, nb))
            L.write_txt("bond.txt", str(tmp_path / "bond.txt"))
            a = _read_txt(tmp_path / "bond.txt", 5)
            assert (a[:, 0] == b1).all() and (a[:, 1] == b2).all() and (a[:, 2] == b3).all() and (a[:, 3] == np.arange(1, nb + 1)).all()
            assert (a[:t, 4] == c).all() and (a[t:, 4] == 0).all()
            s, b3, c, res = L.sitebond(sorder, O.fill_count(0.8, t), bo1, bo2, O.fill_count(0.6, nb))
            L.write_txt("sbsite.txt", str(tmp_path / "sbsite.txt"))
            L.write_txt("sbbond.txt", str(tmp_path / "sbbond.txt"))
            a, b = _read_txt(tmp_path / "sbsite.txt", 3), _read_txt(tmp_path / "sbbond.txt", 3)
            assert (a[:, 1] == s).all() and (a[:, 2] == c).all()
            assert (b[:, 0] == b1).all() and (b[:, 1] == b2).all() and (b[:, 2] == b3).all()


# ---- the bench configuration (square mixed L = 4096, ps 0.80, pb 0.70): every solver form on one realization ---------------
def test_bench_configuration_solver_forms_agree(P):
    """bench.py's first timed realization at its single-point fill: deflated one-pass, plain one-pass and two-kernel form at
    tol 1e-13 agree to 1e-9 (the north star's tolerance) in Gtop and Gbot; the plain forms take linbcg's iteration count,
    the deflated one several times fewer; a tol 1e-10 solve (the bench's tolerance) is within 1e-9 of the 1e-13 one"""
    from percolation_b200.shard import stream_id
    Lsz = 4096
    with P.Lattice(1, Lsz, Lsz, 0) as L:
        L.generate(20240611, stream_id(0, 3), int(0.80 * L.t), int(0.70 * L.nb))
        L.label(P.MIXED)
        assert L.summary()["nspan"] >= 1
        res = {}
        for mode in (0, 2, 1):
            L.set_solver(mode)
            res[mode] = L.conduct(0, tol=1e-13, itmax=4000000, voltages=False)
            assert L.solver_used() == {0: 2, 2: 1, 1: 0}[mode] and res[mode]["err"] <= 1e-13
        L.set_solver(0)
        r10 = L.conduct(0, tol=1e-10, itmax=4000000, voltages=False)
    a = res[1]
    for mode in (0, 2):
        b = res[mode]
        assert abs(a["Gtop"] - b["Gtop"]) <= 1e-9 * abs(a["Gtop"]) and abs(a["Gbot"] - b["Gbot"]) <= 1e-9 * abs(a["Gbot"]), (mode, a, b)
    assert abs(res[2]["iter"] - a["iter"]) <= max(3, a["iter"] // 100)
    assert res[0]["iter"] * 3 < a["iter"]
    assert abs(r10["Gtop"] - res[0]["Gtop"]) <= 1e-9 * abs(res[0]["Gtop"])


# ---- conductance against the oracle's golden values (tests/golden/conduct_fixtures.json) -------------------------
def _golden():
    import json
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "conduct_fixtures.json")
    return json.load(open(p)) if os.path.exists(p) else {}


@pytest.mark.parametrize("name", sorted(_golden()) or ["none"])
def test_conductance_golden(P, name):
    """Gtop / Gbot of orc_conduct_cg (oracle, tol 1e-13; computed on the CPU by tests/golden/make_conduct_fixtures.py from the
    numpy restatement of the generator) against the GPU on the SAME realization -- the occupancy the device draws is
    pinned by CRC -- for the deflated one-pass, the plain one-pass and (L <= 2048) the two-kernel solver.  Includes the bench
    configuration (square mixed L = 4096, ps 0.80, pb 0.70, bench.py's first realization)."""
    import zlib
    gold = _golden()
    if name not in gold:
        pytest.skip("no golden conductance fixtures")
    f = gold[name]
    with P.Lattice(f["lattice"], f["m"], f["n"], 0) as L:
        L.generate(f["seed"], f["stream"], f["ks"], f["kb"])
        socc, bocc = L.get_occupancy(sites=f["kind"] != 2, bonds=f["kind"] != 1)
        if f["kind"] != 2:
            assert zlib.crc32(socc.tobytes()) == f["site_crc"]
        if f["kind"] != 1:
            assert zlib.crc32(bocc.tobytes()) == f["bond_crc"]
        L.label(f["kind"])
        sm = L.summary()
        assert sm["ncl"] == f["ncl"] and sm["maxcs"] == f["maxcs"]
        ids, _ = L.span()
        assert int(ids[0]) == f["cluster"]
        modes = [(0, 2), (2, 1)] + ([(1, 0)] if f["m"] * f["n"] <= 2048 * 2048 else [])
        for mode, used in modes:
            L.set_solver(mode)
            r = L.conduct(0, tol=1e-13, itmax=10000000, voltages=False)
            assert L.solver_used() == used
            assert r["err"] <= 1e-13
            assert abs(r["Gtop"] - f["Gtop"]) <= 1e-9 * f["Gtop"], (mode, r, f["Gtop"])
            assert abs(r["Gbot"] - f["Gbot"]) <= 1e-9 * f["Gbot"], (mode, r, f["Gbot"])
            if mode == 0:
                assert r["iter"] < f["iter"]
            else:
                assert abs(r["iter"] - f["iter"]) <= max(3, f["iter"] // 100)
