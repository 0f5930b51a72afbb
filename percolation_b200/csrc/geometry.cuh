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
#define PERC_HD_COLD __host__ __device__ __forceinline__
#else
#define PERC_HD inline
#define PERC_HD_COLD inline
#endif

namespace perc {

enum : int { LAT_SQUARE = 1, LAT_TRIANGULAR = 2 };
enum : int { KIND_SITE = 1, KIND_BOND = 2, KIND_MIXED = 3 };
enum : int { DIR_E = 0, DIR_N = 1, DIR_NW = 2, DIR_NE = 3 };

enum : unsigned {
    MASK_SITE = 1u, MASK_E = 2u, MASK_N = 4u, MASK_NW = 8u, MASK_NE = 16u, MASK_BONDS = 30u
};

struct Geom {
    int lattice, m, n, pbc;   // n = lattice rows HELD by this handle (the whole lattice, or a slab + halo rows)
    int64_t t;        // m * n  (sites held)
    int64_t nb;       // number of bonds of the WHOLE lattice (Sq/site.f:89-93, Tri/site.f:91-95)
    int ndir;         // owned directions per site: 2 (square) or 4 (triangular)
    int row_bonds;    // reference bond rows contributed by one lattice row y < n-1
    // slab decomposition of one lattice over several GPUs (SURVEY 8e mode 2).  A handle holds the
    // rows [y0, y0 + n) of a lattice of ng rows; rows [own_lo, own_hi) (local numbering) belong to it,
    // the one row below / above is a halo copy of the neighbour's boundary row.  Whole lattice on one
    // GPU: y0 = 0, ng = n, own_lo = 0, own_hi = n.
    int y0, ng, own_lo, own_hi;
    int64_t tg;       // m * ng (sites of the whole lattice)
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
    g.y0 = 0; g.ng = n; g.own_lo = 0; g.own_hi = n; g.tg = g.t;
    return g;
}

// slab of rank `rank` of `nranks`: rows [ya, yb) of the ng-row lattice plus one halo row on each inner side
PERC_HD Geom make_slab_geom(int lattice, int m, int ng, int pbc, int nranks, int rank)
{
    Geom g = make_geom(lattice, m, ng, pbc);                 // nb, row_bonds, tg of the whole lattice
    const int ya = (int)((int64_t)ng * rank / nranks), yb = (int)((int64_t)ng * (rank + 1) / nranks);
    const int hb = rank > 0 ? 1 : 0, ha = rank + 1 < nranks ? 1 : 0;
    g.y0 = ya - hb;
    g.n = yb - ya + hb + ha;
    g.t = (int64_t)m * g.n;
    g.own_lo = hb; g.own_hi = hb + (yb - ya);
    return g;
}

// does site (x,y) own a bond in direction dir?  (x, y: LOCAL coordinates of the handle; the vertical
// extent is that of the whole lattice)
PERC_HD bool bond_exists(const Geom& g, int x, int y, int dir)
{
    switch (dir) {
    case DIR_E:  return x + 1 < g.m || g.pbc;
    case DIR_N:  return g.y0 + y + 1 < g.ng;
    case DIR_NW: return g.lattice == LAT_TRIANGULAR && !(x & 1) && g.y0 + y + 1 < g.ng && (x > 0 || g.pbc);
    case DIR_NE: return g.lattice == LAT_TRIANGULAR && !(x & 1) && g.y0 + y + 1 < g.ng;
    }
    return false;
}

// bit set (MASK_E..MASK_NE) of the bonds site (x,y) owns
PERC_HD unsigned owned_bond_bits(const Geom& g, int x, int y)
{
    unsigned b = 0;
    if (x + 1 < g.m || g.pbc) b |= MASK_E;
    if (g.y0 + y + 1 < g.ng) {
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

// ---- rows of a handle (LOCAL row y; the whole lattice on one GPU: every interior row is both) ------
// rows whose unknowns this handle solves: owned, and not a Dirichlet row (0 and ng-1) of the lattice
PERC_HD bool solve_row(const Geom& g, int y)
{
    return y >= g.own_lo && y < g.own_hi && g.y0 + y >= 1 && g.y0 + y <= g.ng - 2;
}
// rows that carry a value of the search direction p: the solve rows plus the halo copies of the neighbours'
PERC_HD bool p_row(const Geom& g, int y)
{
    return y >= 0 && y < g.n && g.y0 + y >= 1 && g.y0 + y <= g.ng - 2;
}

// ---- full 8-direction neighbourhood (conductance stencil) --------------------------------
// bits: 0 E, 1 N, 2 NW, 3 NE, 4 W, 5 S, 6 SW, 7 SE.  Up-type (x even) triangular sites use
// E,N,NW,NE,W,S; down-type use E,N,W,S,SW,SE; square sites E,N,W,S.
enum : unsigned { NB_E = 1u, NB_N = 2u, NB_NW = 4u, NB_NE = 8u, NB_W = 16u, NB_S = 32u, NB_SW = 64u, NB_SE = 128u };

PERC_HD unsigned neighbour_bits(const Geom& g, int x, int y)
{
    unsigned b = 0;
    bool xl = x > 0 || g.pbc, xr = x + 1 < g.m || g.pbc, yu = g.y0 + y + 1 < g.ng, yd = g.y0 + y > 0;
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


// ---- diagonal of the conductance matrix, bit for bit as the reference forms it -------------------------------------
// G(i,i) = -rowsum with rowsum accumulated over j = 1 .. t in ASCENDING j (Sq/bondc.f:499-505): the order of the additions
// is the order of the neighbours' site numbers.  On this matrix the last bit matters: a diagonal that differs from the
// reference's by one rounding (e.g. the correctly rounded nc*g0 + nl*gleak) is an inconsistent perturbation of relative
// size 1e-16, which the smallest eigenvalue of the cluster (1e-8 .. 1e-9 after Jacobi scaling at L = 1024 .. 4096)
// amplifies to 1e-9 .. 1e-8 in the voltages -- above the 1e-9 the conductance has to agree to.
// Directions in ascending neighbour number for a site away from the periodic seam:
//   square S W E N;  triangular, x even (up-type) S W E NW N NE;  x odd (down-type) SW S SE W E N.
// At the seam (pbc) the wrapped neighbours move: W of x = 0 lies at +m-1 (after E), E of x = m-1 at -(m-1) (before W),
// NW of x = 0 at +2m-1 (after NE), SE of x = m-1 at -(2m-1) (before SW).
PERC_HD double diag_seq(const Geom& g, unsigned cf, unsigned ex, int x, double g0, double gleak)
{
    // weight of the bond in direction `bit`, +0.0 if the neighbour does not exist (d + 0.0 == d: a missing neighbour
    // leaves the running sum as the reference's loop over G(i,j) = 0 does)
#define PERC_W(bit) ((ex & (bit)) ? ((cf & (bit)) ? g0 : gleak) : 0.0)
    const bool tri = g.lattice == LAT_TRIANGULAR;
    if (!(g.pbc && (x == 0 || x == g.m - 1))) {
        if (!tri) return ((PERC_W(NB_S) + PERC_W(NB_W)) + PERC_W(NB_E)) + PERC_W(NB_N);
        if (!(x & 1)) return ((((PERC_W(NB_S) + PERC_W(NB_W)) + PERC_W(NB_E)) + PERC_W(NB_NW)) + PERC_W(NB_N)) + PERC_W(NB_NE);
        return ((((PERC_W(NB_SW) + PERC_W(NB_S)) + PERC_W(NB_SE)) + PERC_W(NB_W)) + PERC_W(NB_E)) + PERC_W(NB_N);
    }
    // periodic seam: the wrapped neighbours change their place in the order
    const bool seam0 = x == 0;
    double d = 0.0;
    if (tri && (x & 1)) {                                  // x = m - 1 (m even): SE wraps to the front
        d = ((d + PERC_W(NB_SE)) + PERC_W(NB_SW)) + PERC_W(NB_S);
        d = ((d + PERC_W(NB_E)) + PERC_W(NB_W)) + PERC_W(NB_N);
        return d;
    }
    d = d + PERC_W(NB_S);
    if (seam0) d = (d + PERC_W(NB_E)) + PERC_W(NB_W);      // W of x = 0 lies at +m-1
    else d = (d + PERC_W(NB_E)) + PERC_W(NB_W);            // E of x = m-1 lies at -(m-1), before W at -1
    if (tri) d = ((d + PERC_W(NB_N)) + PERC_W(NB_NE)) + PERC_W(NB_NW);   // x = 0 (up-type): NW wraps behind NE
    else d = d + PERC_W(NB_N);
    return d;
#undef PERC_W
}
// The same ordered sum together with its rounding residue rho = d - (exact sum of the weights), recovered exactly by
// error-free additions (Knuth's TwoSum; no multiplications, so no contraction can disturb it).  The reference's matrix is
// the exact-row-sum Laplacian plus diag(rho): the deflated solver needs rho wherever it multiplies the matrix with a
// block-constant vector analytically (pcg_fused_tile.cuh).
PERC_HD_COLD double diag_seq_rho(const Geom& g, unsigned cf, unsigned ex, int x, double g0, double gleak, double* rho)
{
    double d = 0.0, err = 0.0;
    // d += w with the rounding error of the addition added to err (w = +0.0 for a missing neighbour: no error)
#define PERC_ADD(bit) { const double w_ = (ex & (bit)) ? ((cf & (bit)) ? g0 : gleak) : 0.0; const double t_ = d + w_, bb_ = t_ - d; \
                        err += (d - (t_ - bb_)) + (w_ - bb_); d = t_; }
    const bool tri = g.lattice == LAT_TRIANGULAR;
    if (!(g.pbc && (x == 0 || x == g.m - 1))) {
        if (!tri) { PERC_ADD(NB_S) PERC_ADD(NB_W) PERC_ADD(NB_E) PERC_ADD(NB_N) }
        else if (!(x & 1)) { PERC_ADD(NB_S) PERC_ADD(NB_W) PERC_ADD(NB_E) PERC_ADD(NB_NW) PERC_ADD(NB_N) PERC_ADD(NB_NE) }
        else { PERC_ADD(NB_SW) PERC_ADD(NB_S) PERC_ADD(NB_SE) PERC_ADD(NB_W) PERC_ADD(NB_E) PERC_ADD(NB_N) }
    } else if (tri && (x & 1)) {
        PERC_ADD(NB_SE) PERC_ADD(NB_SW) PERC_ADD(NB_S) PERC_ADD(NB_E) PERC_ADD(NB_W) PERC_ADD(NB_N)
    } else {
        PERC_ADD(NB_S) PERC_ADD(NB_E) PERC_ADD(NB_W)
        if (tri) { PERC_ADD(NB_N) PERC_ADD(NB_NE) PERC_ADD(NB_NW) } else { PERC_ADD(NB_N) }
    }
#undef PERC_ADD
    *rho = -err;
    return d;
}

// right-hand side of a site of row n-2: Itemp = sum over its bonds into the top row of w Va, in bond-list order =
// ascending neighbour number NW, N, NE (Sq/bondc.f:490-497; the seam moves NW of x = 0 behind NE)
PERC_HD double rhs_seq(const Geom& g, unsigned cf, unsigned ex, int x, double g0, double gleak, double Va)
{
    const bool seam0 = g.pbc && x == 0;
    double b = 0.0;
    if (!seam0 && (ex & NB_NW)) b = b + ((cf & NB_NW) ? g0 : gleak) * Va;
    if (ex & NB_N) b = b + ((cf & NB_N) ? g0 : gleak) * Va;
    if (ex & NB_NE) b = b + ((cf & NB_NE) ? g0 : gleak) * Va;
    if (seam0 && (ex & NB_NW)) b = b + ((cf & NB_NW) ? g0 : gleak) * Va;
    return b;
}

}  // namespace perc
