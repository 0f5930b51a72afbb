"""literal (history-dependent) fills == order-independent union-find partition
after canonicalisation, for all four problem programs x both lattices x pbc."""
import numpy as np
import pytest

CASES = [(lat, m, n, pbc) for lat in (1, 2) for (m, n) in ((12, 10), (16, 7)) for pbc in (0, 1)]


def occ_from(order, k, total):
    occ = np.zeros(total, np.uint8)
    occ[np.asarray(order[:k]) - 1] = 1
    return occ


def bond_occ_from(b1, b2, bo1, bo2, k):
    idx = {(int(a), int(b)): r for r, (a, b) in enumerate(zip(b1, b2))}
    occ = np.zeros(len(b1), np.uint8)
    for a, b in zip(bo1[:k], bo2[:k]):
        occ[idx[(int(a), int(b))]] = 1
    return occ


@pytest.mark.parametrize("lat,m,n,pbc", CASES)
def test_site(O, lat, m, n, pbc):
    t = m * n
    b1, b2 = O.bondlist(lat, m, n, pbc)
    for seed in (11, 12, 13):
        order = O.shuffle_sites(seed, t)
        for k in (0, 1, t // 3, int(0.6 * t), t):
            s, c, res = O.site_literal(lat, m, n, pbc, order, k)
            rc, s_can, _, csz = O.canonicalise(O.SITE, t, b1, b2, s, None, c, res["cln"])
            assert rc == 0
            want, _, wsz, ncl, wmax = O.label_uf(O.SITE, lat, m, n, pbc, b1, b2, site_occ=occ_from(order, k, t))
            assert (s_can == want).all() and (csz == wsz).all()
            assert res["maxcs"] == wmax
            ids = O.spanning(O.SITE, m, n, b1, b2, want, None)
            assert (res["perccln"] > 0) == (len(ids) > 0)
            if res["perccln"]:
                canon = s_can[s == res["perccln"]][0]
                assert canon in ids and res["perccls"] == wsz[canon]


@pytest.mark.parametrize("lat,m,n,pbc", CASES)
def test_bond(O, lat, m, n, pbc):
    t = m * n
    b1, b2 = O.bondlist(lat, m, n, pbc)
    nb = len(b1)
    for seed in (21, 22):
        bo1, bo2 = O.shuffle_bonds(seed, b1, b2)
        for k in (0, 1, nb // 4, nb // 2, nb):
            b3, c, res = O.bond_literal(lat, m, n, pbc, b1, b2, bo1, bo2, k)
            rc, _, b3_can, csz = O.canonicalise(O.BOND, t, b1, b2, None, b3, c, res["cln"])
            assert rc == 0
            occ = bond_occ_from(b1, b2, bo1, bo2, k)
            _, want, wsz, ncl, wmax = O.label_uf(O.BOND, lat, m, n, pbc, b1, b2, bond_occ=occ)
            assert (b3_can == want).all() and (csz == wsz).all() and res["maxcs"] == wmax
            ids = O.spanning(O.BOND, m, n, b1, b2, None, want)
            assert (res["perccln"] > 0) == (len(ids) > 0)


@pytest.mark.parametrize("lat,m,n,pbc", CASES)
def test_mixed_both_orders(O, lat, m, n, pbc):
    t = m * n
    b1, b2 = O.bondlist(lat, m, n, pbc)
    nb = len(b1)
    for seed in (31, 32):
        sorder = O.shuffle_sites(seed, t)
        bo1, bo2 = O.shuffle_bonds(seed + 100, b1, b2)
        for ks, kb in ((0, nb // 3), (t // 2, nb // 2), (int(0.8 * t), int(0.6 * nb)), (t, nb), (t // 2, 0)):
            socc = occ_from(sorder, ks, t)
            bocc = bond_occ_from(b1, b2, bo1, bo2, kb)
            ws, wb, wsz, ncl, wmax = O.label_uf(O.MIXED, lat, m, n, pbc, b1, b2, site_occ=socc, bond_occ=bocc)
            s, b3, c, res = O.sitebond_literal(lat, m, n, pbc, b1, b2, sorder, ks, bo1, bo2, kb)
            rc, s_can, b3_can, csz = O.canonicalise(O.MIXED, t, b1, b2, s, b3, c, res["cln"])
            assert rc == 0
            assert (s_can == ws).all() and (b3_can == wb).all() and (csz == wsz).all()
            if ks + kb > 0:
                assert res["maxcs"] == max(wmax, 1)
            s2, b32, c2, res2 = O.bondsite_literal(lat, m, n, pbc, b1, b2, bo1, bo2, kb, sorder, ks)
            rc, s_can2, b3_can2, csz2 = O.canonicalise(O.MIXED, t, b1, b2, s2, b32, c2, res2["cln"])
            assert rc == 0
            assert (s_can2 == ws).all() and (b3_can2 == wb).all() and (csz2 == wsz).all()
            ids = O.spanning(O.MIXED, m, n, b1, b2, ws, wb)
            big = [i for i in ids if wsz[i] >= 2 * n - 1]
            assert (res["perccln"] > 0) == (len(big) > 0) == (res2["perccln"] > 0)


def test_first_spanning_stop(O):
    """_perc variants: stop at the first step where a spanning cluster exists"""
    lat, m, n, pbc = 1, 14, 14, 0
    t = m * n
    b1, b2 = O.bondlist(lat, m, n, pbc)
    order = O.shuffle_sites(456293, t)
    s, c, res = O.site_literal(lat, m, n, pbc, order, t, stop_at_span=True)
    kstar = res["filled"]
    for k in (kstar - 1, kstar):
        want, _, wsz, _, _ = O.label_uf(O.SITE, lat, m, n, pbc, b1, b2, site_occ=occ_from(order, k, t))
        assert (len(O.spanning(O.SITE, m, n, b1, b2, want, None)) > 0) == (k == kstar)
    bo1, bo2 = O.shuffle_bonds(916125, b1, b2)
    b3, c, res = O.bond_literal(lat, m, n, pbc, b1, b2, bo1, bo2, len(b1), stop_at_span=True)
    kstar = res["filled"]
    for k in (kstar - 1, kstar):
        _, want, wsz, _, _ = O.label_uf(O.BOND, lat, m, n, pbc, b1, b2, bond_occ=bond_occ_from(b1, b2, bo1, bo2, k))
        assert (len(O.spanning(O.BOND, m, n, b1, b2, None, want)) > 0) == (k == kstar)


def test_thresholds_statistical(O):
    """mean first-spanning fraction approaches the p_c quoted in the headers
    (Sq/site.f:24-25 0.593; Tri/site.f:24-25 0.5) -- loose finite-size check"""
    for lat, pc in ((1, 0.593), (2, 0.5)):
        m = n = 24
        t = m * n
        seeds = O.seed_table(58302, 24, 1000000)
        fr = []
        for sd in seeds:
            order = O.shuffle_sites(int(sd), t)
            _, _, res = O.site_literal(lat, m, n, 0, order, t, stop_at_span=True)
            fr.append(float(O.fraction(res["filled"], t)))
        assert abs(np.mean(fr) - pc) < 0.04, (lat, np.mean(fr))
