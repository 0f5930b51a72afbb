"""Oracle vs the reference's own SampleOutput renderings (decoded by
tests/golden/make_png_fixtures.py into tests/golden/png_fixtures.npz):
fill counts int(p*N), nesting within a series, occupied-bond rule and
largest-cluster membership (green bonds)."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
M = N = 50


def load_series():
    z = np.load(os.path.join(HERE, "golden", "png_fixtures.npz"))
    series = {}
    for row in z["index"]:
        key, kind, lattice, p, name = str(row).split("|")
        dots = np.unpackbits(z[key + "_dots"])[: M * N]
        series.setdefault((int(kind), int(lattice)), []).append((float(p), name, dots, z[key + "_cls"]))
    return series


SERIES = load_series()


def test_all_65_images_decoded():
    assert sum(len(v) for v in SERIES.values()) == 65
    assert {k: len(v) for k, v in SERIES.items()} == {(1, 1): 19, (1, 2): 14, (2, 1): 14, (2, 2): 18}


@pytest.mark.parametrize("kind,lattice", sorted(SERIES))
def test_counts_and_nesting(O, kind, lattice):
    prev = None
    nbonds = O.nb(lattice, M, N, 0)
    for p, name, dots, cls in SERIES[(kind, lattice)]:
        occ = dots if kind == O.SITE else (cls > 0)
        total = M * N if kind == O.SITE else nbonds
        assert int(occ.sum()) == O.fill_count(p, total), name      # tsites = ps*t, Sq/site.f:164
        if prev is not None:
            assert (occ >= prev).all(), name                        # one nested run per series
        prev = occ.copy()


@pytest.mark.parametrize("kind,lattice", sorted(SERIES))
def test_largest_cluster_membership(O, kind, lattice):
    b1, b2 = O.bondlist(lattice, M, N, 0)
    for p, name, dots, cls in SERIES[(kind, lattice)]:
        if kind == O.SITE:
            both = (dots[b1 - 1] == 1) & (dots[b2 - 1] == 1)
            assert ((cls > 0) == both).all(), name                  # MATLAB/Square/SitePlot.m:73-85
            s_can, _, csize, ncl, maxcs = O.label_uf(O.SITE, lattice, M, N, 0, b1, b2, site_occ=dots)
            lab_of_bond = s_can[b1 - 1]
        else:
            occ = (cls > 0).astype(np.uint8)
            ends = np.zeros(M * N, np.uint8)
            ends[b1[occ == 1] - 1] = 1
            ends[b2[occ == 1] - 1] = 1
            assert (ends == dots).all(), name                       # MATLAB/Square/BondPlot.m:88-95
            _, b3_can, csize, ncl, maxcs = O.label_uf(O.BOND, lattice, M, N, 0, b1, b2, bond_occ=occ)
            lab_of_bond = b3_can
        if maxcs == 0:
            assert (cls == 0).all()
            continue
        green = cls == 2
        big = np.nonzero(csize == maxcs)[0]
        if green.any():
            # every green cluster is a maximum-size cluster and is green throughout
            # (ties: the renderings colour one or several of the equal-size maxima)
            labs = np.unique(lab_of_bond[green])
            assert all(l in big for l in labs), name
            assert (green == ((cls > 0) & np.isin(lab_of_bond, labs))).all(), name
        else:
            assert kind == O.SITE and maxcs == 1, name              # largest cluster has no bond to draw


@pytest.mark.parametrize("kind,lattice", sorted(SERIES))
def test_literal_fill_reproduces_images(O, kind, lattice):
    """feed the literal (history-dependent) labeler an order consistent with the
    nested series; at every image's count its partition must equal the image's."""
    b1, b2 = O.bondlist(lattice, M, N, 0)
    ser = SERIES[(kind, lattice)]
    rng = np.random.default_rng(7)
    order = []
    prev = np.zeros_like(ser[0][2] if kind == O.SITE else (ser[0][3] > 0).astype(np.uint8))
    for p, name, dots, cls in ser:
        occ = dots if kind == O.SITE else (cls > 0).astype(np.uint8)
        new = np.nonzero((occ == 1) & (prev == 0))[0]
        order.extend(rng.permutation(new).tolist())
        prev = occ
    rest = np.nonzero(prev == 0)[0]
    order.extend(rng.permutation(rest).tolist())
    order = np.array(order, np.int64)
    for p, name, dots, cls in ser[1::3]:
        if kind == O.SITE:
            k = int(dots.sum())
            s, c, res = O.site_literal(lattice, M, N, 0, (order + 1).astype(np.int32), k)
            rc, s_can, _, csz = O.canonicalise(O.SITE, M * N, b1, b2, s, None, c, res["cln"])
            assert rc == 0
            want, _, wsz, _, wmax = O.label_uf(O.SITE, lattice, M, N, 0, b1, b2, site_occ=dots)
            assert (s_can == want).all() and (csz == wsz).all() and res["maxcs"] == wmax, name
        else:
            occ = (cls > 0).astype(np.uint8)
            k = int(occ.sum())
            bo1 = np.ascontiguousarray(b1[order], np.int32)
            bo2 = np.ascontiguousarray(b2[order], np.int32)
            b3, c, res = O.bond_literal(lattice, M, N, 0, b1, b2, bo1, bo2, k)
            rc, _, b3_can, csz = O.canonicalise(O.BOND, M * N, b1, b2, None, b3, c, res["cln"])
            assert rc == 0
            _, want, wsz, _, wmax = O.label_uf(O.BOND, lattice, M, N, 0, b1, b2, bond_occ=occ)
            assert (b3_can == want).all() and (csz == wsz).all() and res["maxcs"] == wmax, name
