"""Oracle conductance: closed forms the reference implies, literal linbcg vs
plain Jacobi-PCG, current conservation (MATLAB/ConductCalc.m:198)."""
import numpy as np
import pytest


@pytest.mark.parametrize("m,n,pbc", [(10, 10, 0), (16, 8, 0), (16, 8, 1), (50, 50, 0)])
def test_full_square_lattice_closed_form(O, m, n, pbc):
    b1, b2 = O.bondlist(O.SQUARE, m, n, pbc)
    w = np.ones(len(b1))
    r = O.conduct_literal(m, n, b1, b2, w, tol=1e-13, itmax=20000)
    assert abs(r["Gtop"] - m / (n - 1)) < 1e-10 and abs(r["Gbot"] - m / (n - 1)) < 1e-10


@pytest.mark.parametrize("m,n,pbc,want", [(10, 10, 0, 1.662318388765), (20, 20, 0, 1.576112235950),
                                          (16, 8, 0, 3.497858355433), (10, 10, 1, 1.721896523400),
                                          (16, 8, 1, 3.576045037457)])
def test_full_triangular_lattice(O, m, n, pbc, want):
    b1, b2 = O.bondlist(O.TRIANGULAR, m, n, pbc)
    w = np.ones(len(b1))
    r = O.conduct_cg(m, n, b1, b2, w, tol=1e-13, itmax=20000)
    assert abs(r["Gtop"] - want) < 2e-11 and abs(r["Gbot"] - want) < 2e-11


@pytest.mark.parametrize("lat", [1, 2])
def test_literal_bicg_equals_pcg_and_conserves_current(O, lat):
    m = n = 20
    b1, b2 = O.bondlist(lat, m, n, 0)
    nb = len(b1)
    pb = 0.56 if lat == 1 else 0.40
    for seed in (626504, 184489, 77):
        bo1, bo2 = O.shuffle_bonds(seed, b1, b2)
        k = O.fill_count(pb, nb)
        b3, c, res = O.bond_literal(lat, m, n, 0, b1, b2, bo1, bo2, k)
        if not res["perccln"]:
            continue
        w = O.weights(O.BOND, b1, b2, None, b3, res["perccln"])
        a = O.conduct_literal(m, n, b1, b2, w)                     # reference defaults 1e-8 / 2500
        b = O.conduct_cg(m, n, b1, b2, w)
        assert a["iter"] == b["iter"]
        assert abs(a["Gtop"] - b["Gtop"]) < 1e-9 * abs(a["Gtop"])
        c1 = O.conduct_cg(m, n, b1, b2, w, tol=1e-13, itmax=100000)
        # Gtop - Gbot artefact: leak bonds kept in the diagonal but dropped from the
        # read-out off-diagonals (Sq/bondc.f:576) -> O(m * 1e-12) absolute
        assert abs(c1["Gtop"] - c1["Gbot"]) < 10 * m * 1e-12
        chk = O.conduct_check(m, n, b1, b2, w, c1["Vint"])
        assert chk["err"] < 1e-12 and chk["Gtop"] == c1["Gtop"]
        return
    pytest.fail("no spanning realization among the seeds")
