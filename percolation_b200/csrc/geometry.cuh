// geometry.cuh -- lattice index arithmetic shared by every kernel and by the host side.
//
// Replaces SUBROUTINE nearestn (reference Fortran/Square/site.f:371-469,
// Fortran/Triangular/site.f:373-558) and the bond-list enumeration block
// (Fortran/Square/bond.f:112-129) by closed forms; nothing is materialised on the device.
//
// Internal conventions (0-based): site i = y*m + x, x in [0,m) fastest, y = 0 is the
// grounded bottom row, y = n-1 the top row at Va.  Reference site number rn = i + 1.
// Triangular lattice (even m only): reference "odd rn" columns (x even here) are the
// up-type sites with neighbours rn-m, rn-1, rn+1, rn+m-1, rn+m, rn+m+1.
//
// Every lattice bond is OWNED by exactly one site through one of four directions:
//   dir 0 = E  -> (x+1, y)   (wraps to x = 0 when pbc)
//   dir 1 = N  -> (x,   y+1)
//   dir 2 = NW -> (x-1, y+1) (triangular, x even; wraps to x = m-1 when pbc)
//   dir 3 = NE -> (x+1, y+1) (triangular, x even)
// Occupancy mask byte per site: bit0 = site active, bit(1+dir) = owned bond occupied.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define PERC_HD __host__ __device__ __forceinline__
#else
#define PERC_HD inline
#endif

namespace perc {

enum : int { LAT_SQUARE = 1, LAT_TRIANGULAR = 2 };
enum : int { KIND_SITE = 1, KIND_BOND = 2, KIND_MIXED = 3 };
enum : int { DIR_E = 0, DIR_N = 1, DIR_NW = 2, DIR_NE = 3 };

enum : unsigned {
    MASK_SITE = 1u, MASK_E = 2u, MASK_N = 4u, MASK_NW = 8u, MASK_NE = 16u, MASK_BONDS = 30u
};

struct Geom {
    int lattice, m, n, pbc;
    int64_t t;        // m * n
    int64_t nb;       // number of lattice bonds (Sq/site.f:89-93, Tri/site.f:91-95)
    int ndir;         // owned directions per site: 2 (square) or 4 (triangular)
    int row_bonds;    // reference bond rows contributed by one lattice row y < n-1
};

PERC_HD Geom make_geom(int lattice, int m, int n, int pbc)
{
    Geom g;
    g.lattice = lattice; g.m = m; g.n = n; g.pbc = pbc;
    g.t = (int64_t)m * n;
    g.ndir = lattice == LAT_SQUARE ? 2 : 4;
    if (lattice == LAT_SQUARE) {
        g.row_bonds = pbc ? 2 * m : 2 * m - 1;
        g.nb = pbc ? (int64_t)m * (2 * n - 1) : 2 * (int64_t)m * n - m - n;
    } else {
        g.row_bonds = pbc ? 3 * m : 3 * m - 2;
        g.nb = pbc ? (int64_t)m * (3 * n - 2) : 3 * (int64_t)m * n - 2 * m - 2 * n + 1;
    }
    return g;
}

// does site (x,y) own a bond in direction dir?
PERC_HD bool bond_exists(const Geom& g, int x, int y, int dir)
{
    switch (dir) {
    case DIR_E:  return x + 1 < g.m || g.pbc;
    case DIR_N:  return y + 1 < g.n;
    case DIR_NW: return g.lattice == LAT_TRIANGULAR && !(x & 1) && y + 1 < g.n && (x > 0 || g.pbc);
    case DIR_NE: return g.lattice == LAT_TRIANGULAR && !(x & 1) && y + 1 < g.n;
    }
    return false;
}

// bit set (MASK_E..MASK_NE) of the bonds site (x,y) owns
PERC_HD unsigned owned_bond_bits(const Geom& g, int x, int y)
{
    unsigned b = 0;
    if (x + 1 < g.m || g.pbc) b |= MASK_E;
    if (y + 1 < g.n) {
        b |= MASK_N;
        if (g.lattice == LAT_TRIANGULAR && !(x & 1)) {
            b |= MASK_NE;
            if (x > 0 || g.pbc) b |= MASK_NW;
        }
    }
    return b;
}

// other end of the bond owned by (x,y) in direction dir
PERC_HD int64_t bond_other_end(const Geom& g, int x, int y, int dir)
{
    int xx = x, yy = y;
    switch (dir) {
    case DIR_E:  xx = x + 1 == g.m ? 0 : x + 1; break;
    case DIR_N:  yy = y + 1; break;
    case DIR_NW: xx = x == 0 ? g.m - 1 : x - 1; yy = y + 1; break;
    default:     xx = x + 1; yy = y + 1; break;
    }
    return (int64_t)yy * g.m + xx;
}

// ---- reference bond numbering <-> (owner site, direction) ------------------------------
// Reference row r (0-based here) = r-th pair produced by
//   do i = 1, t-1; do j = 1, scn; if (nn(j) > i)          (Sq/bond.f:114-129)
// i.e. bonds are listed under their LOWER end point in nearestn order.  Forward order:
//   square:      [E, N]  (+ wrap bond to x = m-1 of the same row, listed under x = 0, pbc)
//   tri x even:  x = 0: [E, N, NE] (+ pbc: [same-row wrap, NW wrap]);  x > 0: [E, NW, N, NE]
//   tri x odd:   [E, N]
// The same-row wrap bond listed under x = 0 is OWNED here by site (m-1, y) as its E bond.

// number of reference rows listed under sites (0..x-1, y)
PERC_HD int ref_row_prefix(const Geom& g, int x, int y)
{
    bool top = (y == g.n - 1);
    if (top) return (g.pbc && x >= 1) ? x + 1 : x;      // top row: [E] (+ wrap under x = 0)
    if (g.lattice == LAT_SQUARE) return g.pbc ? (x >= 1 ? 2 * x + 1 : 0) : 2 * x;
    if (x == 0) return 0;
    if (g.pbc) return (x & 1) ? 3 * x + 2 : 3 * x + 1;
    return (x & 1) ? 3 * x : 3 * x - 1;
}

// reference row (0-based) of the bond owned by (x,y) in direction dir; bond must exist
PERC_HD int64_t bond_ref_row(const Geom& g, int x, int y, int dir)
{
    int64_t base = (int64_t)y * g.row_bonds;
    bool top = (y == g.n - 1);
    if (dir == DIR_E && x == g.m - 1) {                 // same-row wrap, listed under x = 0
        int pos;
        if (top) pos = 1;
        else pos = g.lattice == LAT_SQUARE ? 2 : 3;
        return base + pos;
    }
    int pre = ref_row_prefix(g, x, y);
    int pos = 0;
    if (top) pos = 0;
    else if (g.lattice == LAT_SQUARE) pos = (dir == DIR_E || x == g.m - 1) ? 0 : 1;   // x = m-1 lists [N] only
    else if (x & 1) pos = (dir == DIR_E || x == g.m - 1) ? 0 : 1;
    else if (x == 0) pos = dir == DIR_E ? 0 : dir == DIR_N ? 1 : dir == DIR_NE ? 2 : 4;
    else pos = dir == DIR_E ? 0 : dir == DIR_NW ? 1 : dir == DIR_N ? 2 : 3;
    return base + pre + pos;
}

// inverse: reference row r -> owner site index and direction
PERC_HD void ref_row_to_owner(const Geom& g, int64_t r, int64_t* site, int* dir)
{
    int y = (int)(r / g.row_bonds);
    int rem = (int)(r - (int64_t)y * g.row_bonds);
    if (y >= g.n) { y = g.n - 1; rem = (int)(r - (int64_t)y * g.row_bonds); }
    bool top = (y == g.n - 1);
    int x, pos;
    if (top) {
        if (g.pbc) {
            if (rem == 0) { x = 0; pos = 0; } else if (rem == 1) { x = 0; pos = 1; } else { x = rem - 1; pos = 0; }
            if (x == 0 && pos == 1) { *site = (int64_t)y * g.m + (g.m - 1); *dir = DIR_E; return; }
        } else { x = rem; pos = 0; }
        *site = (int64_t)y * g.m + x; *dir = DIR_E; return;
    }
    if (g.lattice == LAT_SQUARE) {
        if (g.pbc) {
            if (rem < 3) { x = 0; pos = rem; } else { x = (rem - 1) / 2; pos = (rem - 1) % 2; }
            if (x == 0 && pos == 2) { *site = (int64_t)y * g.m + (g.m - 1); *dir = DIR_E; return; }
        } else { x = rem / 2; pos = rem % 2; }
        *site = (int64_t)y * g.m + x; *dir = (pos == 0 && x != g.m - 1) ? DIR_E : DIR_N; return;
    }
    // triangular: pairs (even x, odd x+1) hold 6 rows, the first pair 5 (non-pbc) / 7 (pbc)
    int first = g.pbc ? 5 : 3;                          // rows under x = 0
    if (rem < first) {
        x = 0; pos = rem;
        if (pos == 0) *dir = DIR_E; else if (pos == 1) *dir = DIR_N; else if (pos == 2) *dir = DIR_NE;
        else if (pos == 3) { *site = (int64_t)y * g.m + (g.m - 1); *dir = DIR_E; return; }
        else *dir = DIR_NW;
        *site = (int64_t)y * g.m; return;
    }
    if (rem < first + 2) { x = 1; pos = rem - first; }
    else {
        int q = rem - (first + 2);                      // rows after the first pair
        int pair = q / 6, off = q % 6;
        if (off < 4) { x = 2 + 2 * pair; pos = off; } else { x = 3 + 2 * pair; pos = off - 4; }
    }
    if (x & 1) *dir = (pos == 0 && x != g.m - 1) ? DIR_E : DIR_N;
    else *dir = pos == 0 ? DIR_E : pos == 1 ? DIR_NW : pos == 2 ? DIR_N : DIR_NE;
    *site = (int64_t)y * g.m + x;
}

// ---- full 8-direction neighbourhood (conductance stencil) --------------------------------
// bits: 0 E, 1 N, 2 NW, 3 NE, 4 W, 5 S, 6 SW, 7 SE.  Up-type (x even) triangular sites use
// E,N,NW,NE,W,S; down-type use E,N,W,S,SW,SE; square sites E,N,W,S.
enum : unsigned { NB_E = 1u, NB_N = 2u, NB_NW = 4u, NB_NE = 8u, NB_W = 16u, NB_S = 32u, NB_SW = 64u, NB_SE = 128u };

PERC_HD unsigned neighbour_bits(const Geom& g, int x, int y)
{
    unsigned b = 0;
    bool xl = x > 0 || g.pbc, xr = x + 1 < g.m || g.pbc, yu = y + 1 < g.n, yd = y > 0;
    if (xr) b |= NB_E;
    if (xl) b |= NB_W;
    if (yu) b |= NB_N;
    if (yd) b |= NB_S;
    if (g.lattice == LAT_TRIANGULAR) {
        if (!(x & 1)) { if (yu && xl) b |= NB_NW; if (yu) b |= NB_NE; }
        else          { if (yd) b |= NB_SW; if (yd && xr) b |= NB_SE; }
    }
    return b;
}

}  // namespace perc
