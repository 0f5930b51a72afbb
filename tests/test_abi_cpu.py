"""CPU-side checks of the product library: it loads, exports every symbol that
include/perc_abi.h declares, and its closed-form geometry (host index arithmetic,
no device) reproduces the oracle's literal nearestn / bond-list restatement."""
import os
import re

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


def test_exports_every_declared_symbol(P):
    hdr = open(os.path.join(ROOT, "include", "perc_abi.h")).read()
    declared = sorted(set(re.findall(r"^int32_t\s+(perc_\w+)\s*\(", hdr, re.M)))
    assert declared == sorted(P.SYMBOLS)
    lib = P.load()
    for s in declared:
        assert hasattr(lib, s), s


def test_fortran_interface_binds_every_entry_point():
    """fortran/perc_iface.f90 (the module a reference driver would `use`) declares one bind(C) interface per entry point of
    include/perc_abi.h -- except perc_stitch_host, the host-executed stitch that exists for the CPU tests only -- with the
    header's argument count"""
    hdr = open(os.path.join(ROOT, "include", "perc_abi.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    f90 = open(os.path.join(ROOT, "fortran", "perc_iface.f90")).read()
    decl = {nm: args for nm, args in re.findall(r"^int32_t\s+(perc_\w+)\s*\(([^;]*?)\)\s*;", hdr, re.M | re.S)}
    bound = {nm: args for args, nm in re.findall(r"function\s+perc_\w+\s*\(([^)]*)\)\s*(?:&\s*)?bind\(C,\s*name=\"(perc_\w+)\"\)", f90, re.S)}
    assert set(decl) - set(bound) == {"perc_stitch_host"}
    assert not set(bound) - set(decl)
    for nm, args in bound.items():
        nc = 0 if decl[nm].strip() in ("", "void") else decl[nm].count(",") + 1
        nf = len([a for a in re.sub(r"[&\s]", "", args).split(",") if a])
        assert nc == nf, (nm, nc, nf)


GEOMS = [(lat, m, n, pbc) for lat in (1, 2) for (m, n) in ((4, 3), (6, 5), (8, 2), (10, 10), (34, 7), (66, 35))
         for pbc in (0, 1)]


@pytest.mark.parametrize("lat,m,n,pbc", GEOMS)
def test_bond_numbering_matches_reference_enumeration(P, O, lat, m, n, pbc):
    assert P.geom_nb(lat, m, n, pbc) == O.nb(lat, m, n, pbc)
    b1, b2 = P.geom_bondlist(lat, m, n, pbc)
    r1, r2 = O.bondlist(lat, m, n, pbc)
    assert (b1 == r1).all() and (b2 == r2).all()


@pytest.mark.parametrize("lat,m,n,pbc", [g for g in GEOMS if g[1] <= 10])
def test_nearestn_matches_reference(P, O, lat, m, n, pbc):
    for rn in range(1, m * n + 1):
        assert list(P.geom_nearestn(lat, m, n, pbc, rn)) == list(O.nearestn(lat, m, n, pbc, rn)), rn


def test_argument_errors(P):
    with pytest.raises(P.PercError) as e:
        P.geom_nb(2, 7, 6, 0)                 # odd m on the triangular lattice
    assert e.value.code == -2
    with pytest.raises(P.PercError):
        P.geom_nb(3, 8, 8, 0)
    with pytest.raises(P.PercError):
        P.geom_nb(1, 1, 8, 0)


def test_no_device_fails_loudly(P):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(P.PercError):
        P.Lattice(1, 8, 8)


def _build_c_driver(tmpdir):
    import subprocess
    exe = os.path.join(tmpdir, "abi_c_driver")
    so_dir = os.path.join(ROOT, "percolation_b200")
    subprocess.check_call(["gcc", "-O2", "-o", exe, os.path.join(ROOT, "tests", "abi_c_driver.c"),
                           "-L" + so_dir, "-lperc_b200", "-Wl,-rpath," + so_dir])
    return exe


def test_c_driver_links_and_geometry(P, tmp_path):
    """a C driver written with Fortran calling conventions links against the C-ABI"""
    import subprocess
    exe = _build_c_driver(str(tmp_path))
    out = subprocess.run([exe, "geometry-only"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("OK geometry"), out.stdout + out.stderr
