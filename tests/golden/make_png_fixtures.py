#!/usr/bin/env python
"""Decode the reference's SampleOutput/*.png renderings (50x50 lattices) into
golden fixtures: tests/golden/png_fixtures.npz.

Runs only in the build container (reads /root/reference, which does not exist
on the GPU box); the committed .npz is what the tests use.

The images were drawn by MATLAB/{Square,Triangular}/{Site,Bond}Plot.m from the
Fortran programs' text outputs: a black dot per occupied site (site problem) or
per end point of an occupied bond (bond problem); every lattice bond as a line,
grey (153,153,153) = unoccupied, blue (0,0,204) = occupied, green (0,204,0) =
occupied and in the largest cluster (legend MATLAB/Square/SitePlot.m:34-37).
Colours are exact (no anti-aliasing), so decoding is exact.

Per image we store: p (from the file name), site dots (t bytes), and the colour
class of every reference bond row (0 grey, 1 blue, 2 green) in the order of the
reference's own bond list (Sq/site.f:106-120) -- produced here from the oracle's
literal nearestn restatement.
"""
import glob
import os
import re
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import pyoracle as O  # noqa: E402

REF = "/root/reference/SampleOutput"
M = N = 50


def site_xy(lattice):
    """plot coordinates (a = 1) of sites 1..t, MATLAB/*/SitePlot.m"""
    t = M * N
    X = np.zeros(t)
    Y = np.zeros(t)
    for rn in range(1, t + 1):
        x = (rn - 1) % M + 1
        y = (rn - 1) // M
        if lattice == O.SQUARE:
            X[rn - 1], Y[rn - 1] = x, 1 + y
        else:
            X[rn - 1] = 0.5 * np.sqrt(3.0) * x
            Y[rn - 1] = (0.5 if x % 2 == 0 else 1.0) + y
    return X, Y


def centres(proj):
    idx = np.nonzero(proj > 0)[0]
    groups = np.split(idx, np.nonzero(np.diff(idx) > 1)[0] + 1)
    return np.array([(g[0] + g[-1]) / 2.0 for g in groups])


def calibrate(img, lattice):
    """affine plot->pixel map from the dots of the densest image of a series"""
    dark = img.sum(2) < 120
    cx = centres(dark.sum(0))
    X, Y = site_xy(lattice)
    ux = np.unique(np.round(X, 6))
    assert len(cx) == len(ux), (len(cx), len(ux))
    bx, ax = np.polyfit(ux, cx, 1)
    # rows: project only strips around the odd columns (x = 1, 3, ...): on the
    # triangular lattice even columns sit half a spacing lower and would merge.
    strip = np.zeros(img.shape[1], bool)
    for c in cx[0::2]:
        strip[int(round(c)) - 1:int(round(c)) + 2] = True
    cy = centres((dark & strip[None, :]).sum(1))
    uy = np.unique(np.round(Y[0::2][(np.arange(len(Y[0::2])) % (M // 2)) == 0], 6))
    assert len(cy) == len(uy), (len(cy), len(uy))
    by, ay = np.polyfit(uy, cy[::-1], 1)      # pixel rows grow downwards
    return ax, bx, ay, by


def classify(img, px, py, r):
    """colour class in a (2r+1)^2 window: 2 green, 1 blue, 0 grey/none; -1 ambiguous"""
    h, w, _ = img.shape
    x0, x1 = max(0, int(round(px)) - r), min(w, int(round(px)) + r + 1)
    y0, y1 = max(0, int(round(py)) - r), min(h, int(round(py)) + r + 1)
    win = img[y0:y1, x0:x1].reshape(-1, 3)
    g = int(((win[:, 0] == 0) & (win[:, 1] == 204) & (win[:, 2] == 0)).sum())
    b = int(((win[:, 0] == 0) & (win[:, 1] == 0) & (win[:, 2] == 204)).sum())
    if g and b:
        return 2 if g > 2 * b else (1 if b > 2 * g else -1)
    return 2 if g else (1 if b else 0)


def decode_series(kind, lattice):
    sub = ("Site" if kind == O.SITE else "Bond") + "/" + ("Square" if lattice == O.SQUARE else "Triangular")
    files = sorted(glob.glob(os.path.join(REF, sub, "*.png")))
    named = []
    for f in files:
        mm = re.search(r"_p([0-9.]+)\.png$", f)
        named.append((float(mm.group(1)) if mm else 0.0, f))
    named.sort()
    dense = np.array(Image.open(named[-1][1]).convert("RGB")).astype(int)
    ax, bx, ay, by = calibrate(dense, lattice)
    X, Y = site_xy(lattice)
    b1, b2 = O.bondlist(lattice, M, N, 0)
    out = []
    for p, f in named:
        img = np.array(Image.open(f).convert("RGB")).astype(int)
        px, py = ax + bx * X, ay + by * Y
        dots = np.zeros(M * N, np.uint8)
        for i in range(M * N):
            xi, yi = int(round(px[i])), int(round(py[i]))
            win = img[yi - 1:yi + 2, xi - 1:xi + 2].sum(2)
            dots[i] = 1 if (win < 120).sum() >= 5 else 0
        cls = np.zeros(len(b1), np.int8)
        r = 2 if img.shape[0] > 1000 else 2
        for k in range(len(b1)):
            mx = 0.5 * (px[b1[k] - 1] + px[b2[k] - 1])
            my = 0.5 * (py[b1[k] - 1] + py[b2[k] - 1])
            cls[k] = classify(img, mx, my, r)
        assert (cls >= 0).all(), (f, int((cls < 0).sum()))
        out.append((p, os.path.basename(f), dots, cls))
    return out


def main():
    data = {}
    index = []
    for kind in (O.SITE, O.BOND):
        for lattice in (O.SQUARE, O.TRIANGULAR):
            series = decode_series(kind, lattice)
            for j, (p, name, dots, cls) in enumerate(series):
                key = "k%d_l%d_%02d" % (kind, lattice, j)
                data[key + "_dots"] = np.packbits(dots)
                data[key + "_cls"] = cls
                index.append("%s|%d|%d|%.4f|%s" % (key, kind, lattice, p, name))
                print(key, name, "p=%.3f" % p, "dots", int(dots.sum()), "occ bonds", int((cls > 0).sum()),
                      "green", int((cls == 2).sum()))
    data["index"] = np.array(index)
    np.savez_compressed(os.path.join(HERE, "png_fixtures.npz"), **data)


if __name__ == "__main__":
    main()
