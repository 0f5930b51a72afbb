"""numpy restatement of the K1 generator (percolation_b200/csrc/philox.cuh + occupancy.cu): Philox-4x32-10 element
keys and the exact-count selection "the k elements with the smallest (key, id)".  TEST INFRASTRUCTURE: lets the CPU side
reproduce the occupancy the GPU draws for (seed, stream, ks, kb) -- tests/test_gpu_parity.py checks the GPU against it,
tests/golden/make_conduct_fixtures.py uses it to build the oracle's inputs without a GPU."""
import numpy as np


def philox_pairs(seed, stream, typ, counters):
    """numpy restatement of csrc/philox.cuh (Philox-4x32-10): the two 64-bit keys (A, B) of each call"""
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    ids = np.asarray(counters, np.uint64)
    c0 = ids & np.uint64(0xffffffff)
    c1 = ids >> np.uint64(32)
    c2 = np.full_like(ids, np.uint64(stream & 0xffffffff))
    c3 = np.full_like(ids, np.uint64((stream >> 32) & 0xffffffff))
    k0 = np.uint64(seed & 0xffffffff)
    k1 = np.uint64(((seed >> 32) & 0xffffffff) ^ (0, 0x5bd1e995, 0x2545f491)[typ])
    mask = np.uint64(0xffffffff)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & mask
        n1 = p1 & mask
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ k1) & mask
        n3 = p0 & mask
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return (c0 << np.uint64(32)) | c1, (c2 << np.uint64(32)) | c3


def site_keys(seed, stream, t):
    """site i takes half i & 1 of call (type 0, counter i >> 1)"""
    A, B = philox_pairs(seed, stream, 0, np.arange((t + 1) // 2))
    return np.stack([A, B], 1).reshape(-1)[:t]


def bond_keys_and_ids(seed, stream, lat, m, n, pbc, b1, b2):
    """(key, tie-break id) of every bond in reference row order: E/N from call (type 1, owner), NW/NE from
    call (type 2, owner); id = dir * t + owner"""
    t = m * n
    lo, hi = b1.astype(np.int64) - 1, b2.astype(np.int64) - 1
    x1, y1, y2 = lo % m, lo // m, hi // m
    d = hi - lo
    dirn = np.full(len(lo), -1)
    owner = lo.copy()
    same = y1 == y2
    dirn[same & (d == 1)] = 0
    wrap = same & (d == m - 1) & (m > 2)                       # periodic E bond, owned by the row's last site
    dirn[wrap] = 0
    owner[wrap] = hi[wrap]
    up = y2 == y1 + 1
    dirn[up & (d == m)] = 1
    if lat == 2:
        dirn[up & (d == m + 1)] = 3
        dirn[up & ((d == m - 1) | ((x1 == 0) & (d == 2 * m - 1)))] = 2
    assert (dirn >= 0).all()
    A1, B1 = philox_pairs(seed, stream, 1, owner)
    A2, B2 = philox_pairs(seed, stream, 2, owner)
    key = np.where(dirn == 0, A1, np.where(dirn == 1, B1, np.where(dirn == 2, A2, B2)))
    return key, dirn * t + owner


def generate_occupancy(seed, stream, lat, m, n, pbc, ks, kb, b1, b2):
    """(site_occ[t] or None, bond_occ[nb] or None) of perc_generate(seed, stream, ks, kb); k < 0: that element type is not drawn"""
    t = m * n
    socc = bocc = None
    if ks >= 0:
        keys = site_keys(seed, stream, t)
        order = np.lexsort((np.arange(t), keys))
        socc = np.zeros(t, np.uint8)
        socc[order[:ks]] = 1
    if kb >= 0:
        bkey, bid = bond_keys_and_ids(seed, stream, lat, m, n, pbc, b1, b2)
        order = np.lexsort((bid, bkey))
        bocc = np.zeros(len(b1), np.uint8)
        bocc[order[:kb]] = 1
    return socc, bocc
