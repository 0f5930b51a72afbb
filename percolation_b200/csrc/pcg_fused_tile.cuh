// pcg_fused_tile.cuh -- one Jacobi-PCG iteration in ONE pass over the lattice, written once for device and host.
//
// Replaces the iteration body of linbcg (Sq/bondc.f:780-833: atimes, asolve, the two dot products and the
// three vector updates) for the entry point that needs only Gtop / Gbot (perc_conduct_g, the p-sweep drivers
// Sq/bond_cond.f:392-482).  The recurrences are those of Chronopoulos & Gear's CG (J. Comput. Appl. Math.
// 25, 1989), mathematically the same iterates as linbcg's on a symmetric matrix:
//
//     u = D^-1 r            w = A u
//     s <- w + beta s       (= A p; the search direction p itself is only needed where x is: rows 1 and n-2)
//     r <- r - alpha s      u' = D^-1 r
//     gamma' = r.u'   rr = r.r   delta' = u'.A u'  (energy of the bonds: sum w (u'_i - u'_j)^2)
//     beta' = gamma'/gamma      alpha' = gamma' / (delta' - beta' gamma'/alpha)
//
// so ONE grid-wide reduction per iteration instead of two, and the per-site traffic is r 8 + s 8 + conduct
// byte 1 read, r 8 + s 8 written = 33 B instead of 50 B for the two-kernel form (pcg_pipe_kernel<0|1>).
// r and s are double-buffered in HBM (a tile reads a 2-site halo of r and a 1-site halo of s that its
// neighbours rewrite).  delta' needs u' on the sites right of / above the tile: the tile recomputes r' on
// its own east / north ring (one extra row, two extra columns) instead of waiting for the neighbours.
//
// Every function is __host__ __device__: pcg.cu wraps the phases into the persistent TMA-fed kernel;
// tests/pcg_fused_emul.cpp compiles the SAME source with g++ and runs the phases thread by thread, so the
// arithmetic and the indexing are checked against the oracle on a machine without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include "geometry.cuh"

namespace perc {

// tile configuration: TY tile rows, RPT consecutive rows per thread (sliding 3-row window), DC private copies of
// the diagonal table (32 = one per lane, conflict-free), V = 1: the ring columns are done by the first threads
// after their own rows and every tile ends with a block reduction; V = 2: two extra warps own the ring columns
// (the others do not wait for them) and the sums stay in registers until the CTA has done all its tiles; V = 3: as
// V = 2, and the state vector in HBM is u = D^-1 r instead of r (u' = u - alpha D^-1 s'): phase U and its shared
// array disappear, the residual is formed only inside the sums (r = d u)
constexpr int FT_KMAX = 1152;                              // largest coarse (deflation) dimension: E^-1 is dense, FT_KMAX^2 doubles
                                                           // (L = 4096 with 31-row tiles: 32 x 34 blocks of 1 x 4 tiles = 1088)

template <int TY_, int RPT_, int DC_, int CTAS_, int V_>
struct FtCfg {
    static constexpr int TX = 128, TY = TY_, RPT = RPT_, DC = DC_, CTAS = CTAS_, V = V_;   // CTAS: resident CTAs per SM
    static constexpr int NPAT_SQ = 16, NPAT_TRI = 128;      // conduct patterns of a site with all its neighbours (ft_pat)
    static constexpr int NG = (TY + 1) / RPT;               // row groups
    static_assert(NG * RPT == TY + 1, "compute rows (tile + north ring row) must split evenly");
    static constexpr int MAIN_THREADS = 64 * NG;            // 2 columns per thread
    static constexpr int CR = TY + 1;                       // compute rows: gy = y0 + lr, lr = 0 .. TY (last = north ring)
    static constexpr int RING_T0 = V == 1 ? 0 : MAIN_THREADS;    // first thread that works on the ring columns
    static constexpr int RING_NT = V == 1 ? 2 * CR : 64;         // ... and how many of them share the ring sites (32: ONE ring warp -- measured: no gain,
                                                                 // warps are allocated four at a time, so 23 warps get no more registers than 24)
    static constexpr int THREADS = V == 1 ? MAIN_THREADS : MAIN_THREADS + RING_NT;
    static constexpr int RR = TY + 3;                       // staged rows of r, conduct bytes, u: gy = y0 - 1 + pr
    static constexpr int SR = TY + 1;                       // staged rows of s:                   gy = y0 + ps
    static constexpr int LD = TX + 4;                       // doubles per staged row: column c <-> gx = x0 - 2 + c
    static constexpr int CLD = TX + 32;                     // bytes per staged conduct row: byte b <-> gx = x0 - 16 + b
    static constexpr int R_BYTES = (RR * LD * 8 + 127) / 128 * 128;
    static constexpr int S_BYTES = (SR * LD * 8 + 127) / 128 * 128;
    static constexpr int CF_BYTES = (RR * CLD + 127) / 128 * 128;
    static constexpr int STAGE_BYTES = R_BYTES + S_BYTES + CF_BYTES;
    static constexpr int TAB_BYTES = 64 * DC * 16;
    static constexpr bool USTATE = V >= 3;
    static constexpr bool DEFL = V >= 5;                    // deflated iteration (block-constant coarse space, see FtDefl)
    static constexpr int U_BYTES = USTATE ? 0 : R_BYTES;    // shared u array of phase U
    static constexpr int SFT_N = RR * 4, REC_N = CR * 8;    // shift table (mu per staged row and block column) and its per-row records
    static constexpr int SHIFT_BYTES = DEFL ? 2 * (SFT_N + REC_N) * 8 : 0;               // two tiles in flight
    static constexpr int RHO_BYTES = DEFL ? NPAT_TRI * (DC / 2) * 8 : 0;                 // rho per table slot
    // mu / Z^T A u' of every block, row partials of the coarse product (+ scratch), rho u' per main thread, slot totals of a
    // tile and the running sums of its block
    static constexpr int COARSE_BYTES = DEFL ? FT_KMAX * 8 + 256 * 8 + THREADS * 8 + 16 * 8 + RHO_BYTES : 0;
    static constexpr int SMEM = 2 * STAGE_BYTES + U_BYTES + TAB_BYTES + 96 * 8 + 16 + 64 * 8 + 5 * 16 + SHIFT_BYTES + COARSE_BYTES;
    // table of diagonals of the sites with ALL their neighbours, indexed by which of them conduct (ft_pat: 16 patterns on
    // the square lattice, 2 x 64 on the triangular one) -- the diagonal depends on the pattern, not only on the counts: it is
    // the reference's sum in ascending neighbour order (diag_seq).  Private copies per lane: 32 (square) / 16 (triangular).
    static PERC_HD int tabp(int lat, int pat, int lane) { return lat == LAT_SQUARE ? pat * DC + (lane & (DC - 1)) : pat * (DC / 2) + (lane & (DC / 2 - 1)); }
};

typedef FtCfg<32, 3, 32, 1, 1> FtCfgA;      // 704 threads
typedef FtCfg<32, 3, 32, 1, 3> FtCfgA3;     // 768 threads (two ring warps), one reduction per CTA, u = D^-1 r as the state vector
// deflated sweep: 31-row tiles, four rows per thread -- 16 main warps (four per scheduler) + two ring warps = 576 threads with
// 96 registers each instead of 768 with 80: 0.143 ms per iteration at L = 4096 against 0.151 for the 32-row shape (FtCfgD32,
// PERC_FUSED_TILE32=1), which sits at its register cap; without deflation the two shapes time the same (0.1015 / 0.1007)
typedef FtCfg<31, 4, 32, 1, 5> FtCfgD;
typedef FtCfg<32, 3, 32, 1, 5> FtCfgD32;

#ifdef __CUDACC__
typedef double2 ft_d2;
#else
struct alignas(16) ft_d2 { double x, y; };
#endif
struct alignas(16) FtDiag { double d, inv; };            // diagonal of a site and its reciprocal (0 for d = 0)

struct FtScalars { double g0, gleak, alpha, beta; };

PERC_HD ft_d2 ft_ld2(const double* p) { return *reinterpret_cast<const ft_d2*>(p); }
PERC_HD void ft_st2(double* p, double a, double b) { ft_d2 v; v.x = a; v.y = b; *reinterpret_cast<ft_d2*>(p) = v; }
PERC_HD int ft_popc(unsigned v)
{
#ifdef __CUDA_ARCH__
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
// acc += v if bit != 0 (device: one predicated DADD, which ptxas turns into DADD + two selects)
PERC_HD void ft_padd(double& acc, double v, unsigned bit)
{
#ifdef __CUDA_ARCH__
    asm("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n @p add.f64 %0, %0, %1;\n}" : "+d"(acc) : "d"(v), "r"(bit));
#else
    if (bit) acc += v;
#endif
}
template <int LAT> PERC_HD unsigned ft_interior_ex(int gx)
{
    if (LAT == LAT_SQUARE) return NB_E | NB_N | NB_W | NB_S;
    return (gx & 1) ? (NB_E | NB_N | NB_W | NB_S | NB_SW | NB_SE) : (NB_E | NB_N | NB_W | NB_S | NB_NW | NB_NE);
}
// pattern of a site's conducting bonds as a table index; parity = x & 1 (triangular: up- / down-type site)
template <int LAT> PERC_HD int ft_pat(unsigned cf, int parity)
{
    if (LAT == LAT_SQUARE) return (int)((cf & 3u) | ((cf >> 2) & 12u));                 // E N W S
    return parity ? 64 + (int)((cf & 3u) | ((cf >> 2) & 60u)) : (int)(cf & 63u);        // E N NW NE W S  /  E N W S SW SE
}
// ... and back: conduct bits of table entry `pat`
template <int LAT> PERC_HD unsigned ft_pat_bits(int pat)
{
    if (LAT == LAT_SQUARE) return (unsigned)((pat & 3) | ((pat & 12) << 2));
    return pat >= 64 ? (unsigned)(((pat - 64) & 3) | (((pat - 64) & 60) << 2)) : (unsigned)pat;
}
// table entry: diagonal of an interior site with conduct pattern `pat`, summed as the reference sums it, and its reciprocal
template <int LAT> PERC_HD FtDiag ft_diag_entry(const Geom& g, int pat, double g0, double gleak)
{
    const int parity = LAT == LAT_TRIANGULAR && pat >= 64;
    FtDiag e;
    e.d = diag_seq(g, ft_pat_bits<LAT>(pat), ft_interior_ex<LAT>(parity), 2 + parity, g0, gleak);
    e.inv = e.d > 0.0 ? 1.0 / e.d : 0.0;
    return e;
}
// slot k of the shared-memory table (all private copies of an entry are adjacent)
template <int LAT, class C> PERC_HD int ft_tab_slots() { return LAT == LAT_SQUARE ? C::NPAT_SQ * C::DC : C::NPAT_TRI * (C::DC / 2); }
template <int LAT, class C> PERC_HD FtDiag ft_tab_slot(const Geom& g, int k, double g0, double gleak)
{
    return ft_diag_entry<LAT>(g, LAT == LAT_SQUARE ? k / C::DC : k / (C::DC / 2), g0, gleak);
}
// ... and the rounding residue rho of the same entry (deflation)
template <int LAT, class C> PERC_HD double ft_rho_slot(const Geom& g, int k, double g0, double gleak)
{
    const int pat = LAT == LAT_SQUARE ? k / C::DC : k / (C::DC / 2), parity = LAT == LAT_TRIANGULAR && pat >= 64;
    double rho;
    diag_seq_rho(g, ft_pat_bits<LAT>(pat), ft_interior_ex<LAT>(parity), 2 + parity, g0, gleak, &rho);
    return rho;
}
// diagonal of ANY site (boundary tiles): the same ordered sum over the neighbours that exist
// The reciprocal (the Jacobi scaling: any positive number close to 1/d serves, it does not enter the matrix) comes from a
// small table by bond counts, cinv[(#conducting << 3) | #leaking] (ft_cinv_entry), instead of a division per site.
PERC_HD double ft_cinv_entry(int idx, double g0, double gleak)
{
    const double d = fma((double)(idx & 7), gleak, (double)(idx >> 3) * g0);
    return d > 0.0 ? 1.0 / d : 0.0;
}
PERC_HD_COLD FtDiag ft_diag_site(const Geom& g, unsigned cf, unsigned ex, int gx, double g0, double gleak, const double* cinv)
{
    FtDiag e;
    e.d = diag_seq(g, cf, ex, gx, g0, gleak);
    const int nc = ft_popc(cf & ex);
    e.inv = cinv[(nc << 3) | (ft_popc(ex) - nc)];
    return e;
}
// every site the tile touches (2-site halo) is an unknown with its full neighbourhood, or lies on a Dirichlet
// row where r = 0: the fast paths below need no per-site geometry
template <class C>
PERC_HD bool ft_interior(const Geom& g, int x0, int y0)
{
    return x0 - 2 >= 1 && x0 + C::TX + 1 <= g.m - 2 && y0 >= 1 && y0 + C::TY <= g.n - 2;
}

// ---- deflation ("better preconditioning") -------------------------------------------------------------
// Deflated Jacobi-PCG after Saad, Yeung, Erhel, Guyomarc'h (SIAM J. Sci. Comput. 21, 2000), in the one-pass arrangement:
// coarse space Z = indicator vectors of blocks of bw x bh tiles (restricted to the unknown rows 1 .. n-2), E = Z^T A Z
// (a weighted 5-point / 7-point graph Laplacian of the blocks; its dense inverse, <= FT_KMAX^2 doubles, lives in L2),
//     x0 = Z E^-1 Z^T b,   every search direction  p = u - Z mu + beta p  with  E mu = Z^T A u   (=> Z^T A p = 0),
//     s = A p = A (u - Z mu) + beta s,   p.A p = u.A u - mu.Z^T A u - beta gamma / alpha.
// The sweep is the same 33 B per site.  A Z mu is applied ANALYTICALLY: the matrix is the exact-row-sum Laplacian plus the
// rounding residue rho of its diagonal (diag_seq_rho), so (A Z mu)_i = rho_i mu_i + sum_j w_ij (mu_i - mu_j), and the sum
// has terms only where a bond crosses a block border or ends on a Dirichlet row (the shift table of a tile: mu per staged
// row and west / own / east block column).  Shifting u by mu on its way into the stencil instead would evaluate large
// cancelling sums: every site of one conduct pattern in a block would commit the SAME rounding error, a block-coherent
// residual that the deflated iteration cannot remove and the cluster's small eigenvalues amplify (2e-10 in G at L = 512,
// 2e-8 at L = 4096).  Z^T A u' -- the net current leaving each block -- comes from the bonds that CROSS tile borders
// (slots per tile: into the tile to the east / north / west / north-west, into the Dirichlet rows) plus sum rho_i u'_i.
struct FtDefl {
    int bw, bh;        // tiles per block
    int nbx, nby;      // blocks per lattice row / column
    int ntx, nty;      // tiles per lattice row / column
    int k;             // nbx * nby <= FT_KMAX
    int sw, sh;        // log2 bw / log2 bh when they are powers of two (the library's own choice always is), else -1
    int pbc;           // periodic wrap in x (m a multiple of TX): the last tile / block column is the west neighbour of the first
};
// FS_R: sum over the tile's unknown sites of rho_i u'_i (rho: rounding residue of the matrix diagonal, diag_seq_rho) -- the part of
// Z^T A u' that is not a current between blocks; in the weight pass (UNIT) the sum of rho_i itself: the entry it adds to E
enum : int { FS_E = 0, FS_N = 1, FS_W = 2, FS_NW = 3, FS_D = 4, FS_R = 5, FS_SLOTS = 6, FS_STRIDE = 8 };

// blocks start as single tiles and grow (towards square blocks) until the coarse dimension fits; kmax <= 0: no limit
// (the kernels without deflation walk the lattice with bw = bh = 1: block = tile)
PERC_HD FtDefl ft_defl_make(const Geom& g, int TX, int TY, int kmax, int bw0 = 0, int bh0 = 0)
{
    FtDefl D;
    D.ntx = (g.m + TX - 1) / TX; D.nty = (g.n + TY - 1) / TY;
    D.bw = bw0 > 0 ? bw0 : 1; D.bh = bh0 > 0 ? bh0 : 1;
    for (;;) {
        D.nbx = (D.ntx + D.bw - 1) / D.bw; D.nby = (D.nty + D.bh - 1) / D.bh;
        D.k = D.nbx * D.nby;
        if (kmax <= 0 || D.k <= kmax) break;
        if (D.bw * TX < D.bh * TY) D.bw *= 2; else D.bh *= 2;
    }
    D.sw = D.sh = -1; D.pbc = g.pbc;
    for (int b = 0; b < 30; ++b) { if (D.bw == (1 << b)) D.sw = b; if (D.bh == (1 << b)) D.sh = b; }
    return D;
}
PERC_HD int ft_defl_block(const FtDefl& D, int ix, int iy)
{
    return (D.sh >= 0 ? iy >> D.sh : iy / D.bh) * D.nbx + (D.sw >= 0 ? ix >> D.sw : ix / D.bw);
}

#if defined(__CUDA_ARCH__)
#define FT_LDCG(p) __ldcg(p)           // written by other CTAs of the same launch: read through L2
#else
#define FT_LDCG(p) (*(p))
#endif

// entry (staged row pr, block column cls: 0 west / 1 own / 2 east) of a tile's shift table: mu of the block the sites
// of that row and column class lie in; 0 on rows that are not unknowns (Dirichlet rows, outside the lattice)
template <class C>
PERC_HD double ft_defl_shift_entry(const Geom& g, const FtDefl& D, const double* mu, int ix, int iy, int pr, int cls)
{
    const int gy = iy * C::TY - 1 + pr;
    if (gy < 1 || gy > g.n - 2) return 0.0;
    const int ty = pr == 0 ? iy - 1 : (pr > C::TY ? iy + 1 : iy);
    int tx = ix + cls - 1;
    if (ty < 0 || ty >= D.nty) return 0.0;
    if (tx < 0 || tx >= D.ntx) { if (!D.pbc) return 0.0; tx = tx < 0 ? D.ntx - 1 : 0; }
    return mu[ft_defl_block(D, tx, ty)];
}
// class of a staged column (column c <-> gx = x0 - 2 + c)
template <class C> PERC_HD int ft_defl_cls(int col) { return col < 2 ? 0 : (col >= 2 + C::TX ? 2 : 1); }
// entry e of the record of compute row lr, from the tile's shift table: 0 mu of the row's own sites, then mu minus the mu of
// the neighbour to the 1 north, 2 south, (3: always 0) 4 west, 5 east, 6 north-west across the west border, 7 south-east
// across the east border -- what ft_phase_main adds to the neighbours' values
template <class C>
PERC_HD double ft_defl_rec_entry(const double* sft, int lr, int e)
{
    const double mc = sft[(lr + 1) * 4 + 1];
    switch (e) {
    case 0: return mc;
    case 1: return mc - sft[(lr + 2) * 4 + 1];
    case 2: return mc - sft[lr * 4 + 1];
    case 4: return mc - sft[(lr + 1) * 4 + 0];
    case 5: return mc - sft[(lr + 1) * 4 + 2];
    case 6: return mc - sft[(lr + 2) * 4 + 0];
    case 7: return mc - sft[lr * 4 + 2];
    default: return 0.0;
    }
}

// The order in which a CTA visits the lattice: block by block (block bid, bid + G, ...), the tiles of a block one after the
// other, so that the crossing currents of a block are summed by the CTA that owns it; rev: the same sequence backwards
// (consecutive iterations sweep in opposite directions: each starts on what L2 still holds).
struct FtWalk {
    int B, bx, by, w, h, jx, jy;
    int pos, lo, hi;                   // with a schedule: position in this CTA's list of blocks [lo, hi)
    PERC_HD bool valid(const FtDefl& D) const { return B >= 0 && B < D.k; }
    PERC_HD void enter(const FtDefl& D, int rev)
    {
        bx = B % D.nbx; by = B / D.nbx;
        w = D.ntx - bx * D.bw < D.bw ? D.ntx - bx * D.bw : D.bw;
        h = D.nty - by * D.bh < D.bh ? D.nty - by * D.bh : D.bh;
        jx = rev ? w - 1 : 0; jy = rev ? h - 1 : 0;
    }
    // sched = nullptr: block bid, bid + G, ...; else sched[0 .. G] = offsets of the CTAs' lists, which follow (ascending block
    // numbers; built on the host so that the slower boundary tiles are spread evenly: ft_defl_schedule)
    PERC_HD void start(const FtDefl& D, int bid, int G, int rev, const int* sched = nullptr)
    {
        pos = lo = hi = 0;
        if (sched) {
            lo = sched[bid]; hi = sched[bid + 1];
            pos = rev ? hi - 1 : lo;
            B = lo < hi ? sched[G + 1 + pos] : -1;
        } else B = bid < D.k ? (rev ? bid + ((D.k - 1 - bid) / G) * G : bid) : -1;
        if (valid(D)) enter(D, rev);
    }
    PERC_HD void next_block(const FtDefl& D, int G, int rev, const int* sched)
    {
        if (sched) { pos += rev ? -1 : 1; B = (pos >= lo && pos < hi) ? sched[G + 1 + pos] : -1; }
        else B += rev ? -G : G;
        if (valid(D)) enter(D, rev);
    }
    PERC_HD void next(const FtDefl& D, int G, int rev, const int* sched = nullptr)
    {
        if (!rev) { if (++jx == w) { jx = 0; if (++jy == h) next_block(D, G, 0, sched); } }
        else      { if (--jx < 0) { jx = w - 1; if (--jy < 0) next_block(D, G, 1, sched); } }
    }
    PERC_HD int ix(const FtDefl& D) const { return bx * D.bw + jx; }
    PERC_HD int iy(const FtDefl& D) const { return by * D.bh + jy; }
    // which of the tile's slots leave the block, and whether the tile is the first / last of its block in walking order
    PERC_HD int info(const FtDefl& D, int rev) const
    {
        const int x = bx * D.bw + jx, y = by * D.bh + jy;
        // (periodic wrap: the tile east of the last column is the first one -- another block unless there is only one block column)
        const int wrapx = D.pbc && D.nbx > 1;
        const int cE = jx == w - 1 && (x + 1 < D.ntx || wrapx), cN = jy == h - 1 && y + 1 < D.nty, cW = jx == 0 && (x >= 1 || wrapx);
        const int nwx = jx == 0 && (x >= 1 || wrapx), nwy = jy == h - 1, cNW = (x >= 1 || D.pbc) && y + 1 < D.nty && (nwx | nwy);
        const int first = rev ? (jx == w - 1 && jy == h - 1) : (jx == 0 && jy == 0), last = rev ? (jx == 0 && jy == 0) : (jx == w - 1 && jy == h - 1);
        return cE | cN << 1 | cW << 2 | cNW << 3 | (nwx & cNW) << 4 | (nwy & cNW) << 5 | first << 6 | last << 7;
    }
};
enum : int { FW_E = 1, FW_N = 2, FW_W = 4, FW_NW = 8, FW_NWX = 16, FW_NWY = 32, FW_FIRST = 64, FW_LAST = 128 };
// running sums of a block: what leaves it altogether, and what enters the block to the east / north / west / north-west
enum : int { FB_OUT = 0, FB_E = 1, FB_N = 2, FB_W = 3, FB_NW = 4, FB_PLANES = 5 };

// slot totals ts[FS_*] of one tile into the running sums of its block (info: FtWalk::info of the tile)
PERC_HD void ft_defl_block_add(int info, const double* ts, double* acc)
{
    double out = ts[FS_D] + ts[FS_R];
    if (info & FW_E) { out += ts[FS_E]; acc[FB_E] += ts[FS_E]; }
    if (info & FW_N) { out += ts[FS_N]; acc[FB_N] += ts[FS_N]; }
    if (info & FW_W) { out += ts[FS_W]; acc[FB_W] += ts[FS_W]; }
    if (info & FW_NW) { out += ts[FS_NW]; acc[(info & FW_NWX) ? ((info & FW_NWY) ? FB_NW : FB_W) : FB_N] += ts[FS_NW]; }
    acc[FB_OUT] += out;
}
// Z^T (A u') of block B -- the net current leaving it -- from the block sums Fb[plane * FT_KMAX + block]
// (square lattice: nothing ever enters from the east / south-east, those planes are not read)
template <int LAT>
PERC_HD double ft_defl_block_f(const FtDefl& D, const double* Fb, int B)
{
    const int bx = B % D.nbx, by = B / D.nbx;
    // (independent loads: the addresses are clamped, the values selected afterwards)
    const int wrapx = D.pbc && D.nbx > 1;
    const bool he = bx > 0 || wrapx, hn = by > 0, hw = bx + 1 < D.nbx || wrapx, hnw = hw && hn;
    const int Bw = bx > 0 ? B - 1 : B + D.nbx - 1, Be = bx + 1 < D.nbx ? B + 1 : B - (D.nbx - 1);     // west / east neighbour block (wrapped)
    const double o = FT_LDCG(&Fb[FB_OUT * FT_KMAX + B]);
    double e = FT_LDCG(&Fb[FB_E * FT_KMAX + (he ? Bw : B)]);
    double n = FT_LDCG(&Fb[FB_N * FT_KMAX + (hn ? B - D.nbx : B)]);
    e = he ? e : 0.0; n = hn ? n : 0.0;
    if (LAT == LAT_SQUARE) return (o - e) - n;
    double w = FT_LDCG(&Fb[FB_W * FT_KMAX + (hw ? Be : B)]);
    double nw = FT_LDCG(&Fb[FB_NW * FT_KMAX + (hnw ? Be - D.nbx : B)]);
    w = hw ? w : 0.0; nw = hnw ? nw : 0.0;
    return (((o - e) - n) - w) - nw;
}

// one work item q of a tile's crossing currents: the E bonds of its east column, the N / NE / NW bonds of its top row,
// (triangular) the NW bonds of its west column, and the bonds of rows 1 / n-2 into the Dirichlet rows.  Only bonds between
// two unknown rows cross blocks; a bond into a Dirichlet row leaves the block for good (slot FS_D).  UNIT: u_i = 1, u_j = 0
// (the bond weights themselves: the entries of E).  A: accessor with cf(gx, gy) and u(gx, gy).
template <class C> struct FtFluxItems { static constexpr int N = 2 * C::TY + 3 * C::TX; };
template <int LAT, class C, bool UNIT, class A>
PERC_HD void ft_flux_item(const Geom& g, double g0, double gleak, const A& a, int x0, int y0, int q, double* f, bool interior = false)
{
#define FT_W(cf, bit) (((cf) & (bit)) ? g0 : gleak)
#define FT_UNK(gy) ((gy) >= 1 && (gy) <= g.n - 2)
    if (interior && !UNIT) {
        // tile without lattice borders (ft_interior): every bond exists, every row is an unknown, no Dirichlet items
        if (q < C::TY) {
            const int gy = y0 + q, gx = x0 + C::TX - 1;
            f[FS_E] += FT_W(a.cf(gx, gy), NB_E) * (a.u(gx, gy) - a.u(gx + 1, gy));
        } else if (q < C::TY + C::TX) {
            const int gx = x0 + q - C::TY, gy = y0 + C::TY - 1;
            const unsigned cf = a.cf(gx, gy);
            const double ui = a.u(gx, gy);
            f[FS_N] += FT_W(cf, NB_N) * (ui - a.u(gx, gy + 1));
            if (LAT == LAT_TRIANGULAR && !(gx & 1)) {
                f[FS_N] += FT_W(cf, NB_NE) * (ui - a.u(gx + 1, gy + 1));
                f[q == C::TY ? FS_NW : FS_N] += FT_W(cf, NB_NW) * (ui - a.u(gx - 1, gy + 1));
            }
        } else if (LAT == LAT_TRIANGULAR && q < 2 * C::TY + C::TX - 1) {
            const int gx = x0, gy = y0 + q - C::TY - C::TX;
            f[FS_W] += FT_W(a.cf(gx, gy), NB_NW) * (a.u(gx, gy) - a.u(gx - 1, gy + 1));
        }
        return;
    }
    if (q < C::TY) {
        const int gy = y0 + q, gx = x0 + C::TX - 1;
        if ((gx + 1 < g.m || g.pbc) && FT_UNK(gy)) f[FS_E] += FT_W(a.cf(gx, gy), NB_E) * (UNIT ? 1.0 : a.u(gx, gy) - a.u(gx + 1, gy));
        return;
    }
    q -= C::TY;
    if (q < C::TX) {
        const int gx = x0 + q, gy = y0 + C::TY - 1;
        if (gx < g.m && FT_UNK(gy) && FT_UNK(gy + 1)) {
            const unsigned cf = a.cf(gx, gy), ex = neighbour_bits(g, gx, gy);
            const double ui = UNIT ? 1.0 : a.u(gx, gy);
            if (ex & NB_N) f[FS_N] += FT_W(cf, NB_N) * (UNIT ? 1.0 : ui - a.u(gx, gy + 1));
            if (LAT == LAT_TRIANGULAR) {
                if (ex & NB_NE) f[FS_N] += FT_W(cf, NB_NE) * (UNIT ? 1.0 : ui - a.u(gx + 1, gy + 1));
                if (ex & NB_NW) f[q == 0 ? FS_NW : FS_N] += FT_W(cf, NB_NW) * (UNIT ? 1.0 : ui - a.u(gx - 1, gy + 1));
            }
        }
        return;
    }
    q -= C::TX;
    if (q < C::TY) {
        if (LAT == LAT_TRIANGULAR && q < C::TY - 1) {
            const int gx = x0, gy = y0 + q;
            if (FT_UNK(gy) && FT_UNK(gy + 1) && (neighbour_bits(g, gx, gy) & NB_NW))
                f[FS_W] += FT_W(a.cf(gx, gy), NB_NW) * (UNIT ? 1.0 : a.u(gx, gy) - a.u(gx - 1, gy + 1));
        }
        return;
    }
    q -= C::TY;
    const int top = q >= C::TX;
    q -= top * C::TX;
    const int gy = top ? g.n - 2 : 1, gx = x0 + q;
    if (gy < y0 || gy >= y0 + C::TY || !FT_UNK(gy) || gx >= g.m) return;
    const unsigned cf = a.cf(gx, gy), ex = neighbour_bits(g, gx, gy);
    const double ui = UNIT ? 1.0 : a.u(gx, gy);
    double acc = 0.0;
    if (top) {
        if (ex & NB_N) acc += FT_W(cf, NB_N);
        if (ex & NB_NW) acc += FT_W(cf, NB_NW);
        if (ex & NB_NE) acc += FT_W(cf, NB_NE);
    } else {
        if (ex & NB_S) acc += FT_W(cf, NB_S);
        if (ex & NB_SW) acc += FT_W(cf, NB_SW);
        if (ex & NB_SE) acc += FT_W(cf, NB_SE);
    }
    // n = 3: row 1 is also row n-2 and both items add their share
    f[FS_D] += acc * ui;
#undef FT_W
#undef FT_UNK
}

// accessor over a tile's shared-memory stage: u' of the compute rows (ss) and the staged conduct bytes
template <class C> struct FtStageAcc {
    const double* ss; const uint8_t* scf; int x0, y0;
    PERC_HD unsigned cf(int gx, int gy) const { return scf[(gy - y0 + 1) * C::CLD + (gx - x0 + 16)]; }
    PERC_HD double u(int gx, int gy) const { return ss[(gy - y0) * C::LD + (gx - x0 + 2)]; }
};
// accessor over the conduct bytes in global memory (UNIT items only)
struct FtGlobalAcc {
    const uint8_t* cfull; int m;
    PERC_HD unsigned cf(int gx, int gy) const { return cfull[(int64_t)gy * m + gx]; }
    PERC_HD double u(int, int) const { return 0.0; }
};

// crossing currents of a tile by the ring threads (rl = 0 .. RING_NT-1) while the others sum the bond energies.  Every slot has ONE
// producer warp: the first takes the east column (FS_E), the west column (FS_W, triangular), the bonds into the Dirichlet
// rows (FS_D) and sum rho u' (FS_R: the per-thread sums the main threads left in sru -- they run over all tiles of a block
// and are collected with its last tile: with_r); the second the top row (FS_N, FS_NW).  Configurations with ONE ring warp
// (RING_NT = 32): that warp plays both.  Partial sums of this thread into f[FS_SLOTS]; the caller folds the lanes of each warp.
template <int LAT, class C>
PERC_HD void ft_flux_thread(const Geom& g, const FtScalars& sc, const double* ss, const uint8_t* scf, const double* sru, int x0, int y0,
                            int rl, bool interior, bool with_r, double* f)
{
    const FtStageAcc<C> a{ss, scf, x0, y0};
    const int l = rl & 31;
    const bool first = rl < 32, second = C::RING_NT == 32 || rl >= 32;      // (one ring warp plays both)
    if (first) {
        for (int q = l; q < C::TY; q += 32) ft_flux_item<LAT, C, false>(g, sc.g0, sc.gleak, a, x0, y0, q, f, interior);
        if (LAT == LAT_TRIANGULAR)
            for (int q = C::TY + C::TX + l; q < 2 * C::TY + C::TX; q += 32) ft_flux_item<LAT, C, false>(g, sc.g0, sc.gleak, a, x0, y0, q, f, interior);
        if (!interior && (y0 <= 1 || y0 + C::TY >= g.n - 2))
            for (int q = 2 * C::TY + C::TX + l; q < FtFluxItems<C>::N; q += 32) ft_flux_item<LAT, C, false>(g, sc.g0, sc.gleak, a, x0, y0, q, f, false);
        if (with_r) for (int j = l; j < C::MAIN_THREADS; j += 32) f[FS_R] += sru[j];
    }
    if (second)
        for (int q = C::TY + l; q < C::TY + C::TX; q += 32) ft_flux_item<LAT, C, false>(g, sc.g0, sc.gleak, a, x0, y0, q, f, interior);
}

// deflated start: x0 = Z nu with E nu = Z^T b; returns u0 = D^-1 (b - A Z nu) of the unknown site (x, y) (cf: its conduct
// byte); *bi = its right-hand side (bonds into the top row at Va), *nui = nu of its block
template <class C>
PERC_HD double ft_defl_u0(const Geom& g, const FtDefl& D, unsigned cf, const double* nu, int x, int y, double Va, double g0,
                          double gleak, double* bi, double* nui)
{
    const unsigned ex = neighbour_bits(g, x, y);
    cf &= ex;
    double rho;
    const double d = diag_seq_rho(g, cf, ex, x, g0, gleak, &rho);
    const double ni = nu[ft_defl_block(D, x / C::TX, y / C::TY)];
    const unsigned bits[8] = {NB_E, NB_N, NB_NW, NB_NE, NB_W, NB_S, NB_SW, NB_SE};
    const int ddx[8] = {1, 0, -1, 1, -1, 0, -1, 1}, ddy[8] = {0, 1, 1, 1, 0, -1, -1, -1};
    const double b = y == g.n - 2 ? rhs_seq(g, cf, ex, x, g0, gleak, Va) : 0.0;
    // (A Z nu)_i = rho_i nu_i + sum_j w_ij (nu_i - nu_j): its block sums are E nu exactly as E is assembled (weights of the
    // crossing bonds + sum of rho on the diagonal), so Z^T r0 = 0 holds in the arithmetic the coarse stage works in
    double acc = rho * ni;
    for (int k = 0; k < 8; ++k) {
        if (!(ex & bits[k])) continue;
        const double w = (cf & bits[k]) ? g0 : gleak;
        const int xx = x + ddx[k], yy = y + ddy[k];
        if (yy >= 1 && yy <= g.n - 2) acc += w * (ni - nu[ft_defl_block(D, (xx < 0 ? xx + g.m : (xx >= g.m ? xx - g.m : xx)) / C::TX, yy / C::TY)]);
        else acc += w * ni;
    }
    *bi = b; *nui = ni;
    return d > 0.0 ? (b - acc) / d : 0.0;
}

// ---- periodic wrap in x (pbc = 1; one-pass kernel with u as the state vector, m a multiple of TX) ----------------------
// a site of the first / last lattice column with pbc: its wrapped neighbours change their place in the reference's ordered
// diagonal sum (diag_seq), so its diagonal is computed, not taken from the pattern table
PERC_HD bool ft_seam(const Geom& g, int gx) { return g.pbc && (gx == 0 || gx == g.m - 1); }
// The TMA boxes of the first / last tile of a lattice row hold zeros in the halo columns beyond the seam; before the tile
// phases run, the ring threads (rl of nthr) overwrite them with the wrapped columns read straight from the input vectors.
// Work item q = (side, staged row): side 0 = west halo of a tile with x0 = 0 (lattice columns m-2, m-1), side 1 = east halo
// of a tile with x0 + TX = m (lattice columns 0, 1).
template <class C>
PERC_HD void ft_wrap_patch(const Geom& g, int x0, int y0, const double* u_in, const double* s_in, const uint8_t* cfull,
                           double* sr, double* ss, uint8_t* scf, int rl, int nthr)
{
    for (int q = rl; q < 2 * C::RR; q += nthr) {
        const int side = q >= C::RR, pr = q - side * C::RR, gy = y0 - 1 + pr;
        if (side == 0 ? x0 != 0 : x0 + C::TX != g.m) continue;
        if (gy < 0 || gy >= g.n) continue;
        const int gx = side ? 0 : g.m - 2, col = side ? 2 + C::TX : 0;     // lattice columns gx, gx + 1 -> box columns col, col + 1
        const int64_t o = (int64_t)gy * g.m + gx;
        sr[pr * C::LD + col] = u_in[o]; sr[pr * C::LD + col + 1] = u_in[o + 1];
        scf[pr * C::CLD + (side ? 16 + C::TX : 14)] = cfull[o]; scf[pr * C::CLD + (side ? 17 + C::TX : 15)] = cfull[o + 1];
        if (pr >= 1 && pr <= C::SR) { ss[(pr - 1) * C::LD + col] = s_in[o]; ss[(pr - 1) * C::LD + col + 1] = s_in[o + 1]; }
    }
}

// ---- phase U: u = r / d on the staged box (tile + 2-site halo; rows y0-1 .. y0+TY+1) -------------------------
template <int LAT, class C>
PERC_HD void ft_phase_u(const Geom& g, const double* sr, const uint8_t* scf, double* su, const FtDiag* dtab,
                        int x0, int y0, bool interior, int tid, double g0, double gleak, const double* cinv)
{
    constexpr int NCP = C::LD / 2, TRIPS = (C::RR * NCP + C::THREADS - 1) / C::THREADS;
    const int lane = tid & 31;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int it = 0; it < TRIPS; ++it) {
        const int k = tid + it * C::THREADS;
        if (k >= C::RR * NCP) break;
        const int pr = k / NCP, cp = k - pr * NCP;
        const int gy = y0 - 1 + pr, gx = x0 - 2 + 2 * cp;
        double u0 = 0.0, u1 = 0.0;
        if (gy >= 1 && gy <= g.n - 2) {
            const ft_d2 r2 = ft_ld2(&sr[pr * C::LD + 2 * cp]);
            const unsigned c01 = *reinterpret_cast<const unsigned short*>(&scf[pr * C::CLD + 14 + 2 * cp]);
            unsigned e0, e1;
            if (interior) { e0 = ft_interior_ex<LAT>(gx); e1 = ft_interior_ex<LAT>(gx + 1); }
            else {
                e0 = (gx >= 0 && gx < g.m) ? neighbour_bits(g, gx, gy) : 0u;
                e1 = (gx + 1 >= 0 && gx + 1 < g.m) ? neighbour_bits(g, gx + 1, gy) : 0u;
            }
            const unsigned f0 = c01 & 0xffu & e0, f1 = (c01 >> 8) & e1;
            u0 = r2.x * ((interior || e0 == ft_interior_ex<LAT>(gx)) ? dtab[C::tabp(LAT, ft_pat<LAT>(f0, gx & 1), lane)] : ft_diag_site(g, f0, e0, gx, g0, gleak, cinv)).inv;
            u1 = r2.y * ((interior || e1 == ft_interior_ex<LAT>(gx + 1)) ? dtab[C::tabp(LAT, ft_pat<LAT>(f1, (gx + 1) & 1), lane)] : ft_diag_site(g, f1, e1, gx + 1, g0, gleak, cinv)).inv;
        }
        ft_st2(&su[pr * C::LD + 2 * cp], u0, u1);
    }
}

// ---- phase M: w = A u, s' = w + beta s, r' = r - alpha s', u' = r'/d on the compute rows; stores r', s' of
// the tile rows; u' replaces s in shared memory (each thread overwrites only what it has read itself);
// p / x of the two read-out rows; sums r'.u' and r'.r' over the tile's sites ---------------------------------
// INT: the tile has no lattice border within its 2-site halo (ft_interior) -- its own instantiation: no geometry, no
// validity tests, no read-out rows (they lie in boundary tiles), every diagonal from the pattern table; the boundary tiles
// (a few per cent) run the general code without weighing on the registers and the instruction count of this one
template <int LAT, class C, bool INT>
PERC_HD void ft_phase_main(const Geom& g, const FtScalars& sc, const double* sr, double* ss, const uint8_t* scf,
                           const double* su, const FtDiag* dtab, const double* cinv, int x0, int y0, bool interior_flag, int tid,
                           double* __restrict__ r_out, double* __restrict__ s_out, double* __restrict__ xrow,
                           double* __restrict__ prow, double& acc_rz, double& acc_rr, const double* srec = nullptr,
                           const double* rtab = nullptr, double* acc_ru = nullptr)
{
    if (tid >= C::MAIN_THREADS) return;
    constexpr bool interior = INT;
    (void)interior_flag;
    const int tx = tid & 63, ty = tid >> 6, lane = tid & 31;
    const int gx = x0 + 2 * tx, lr0 = ty * C::RPT;
    const double alpha = sc.alpha, beta = sc.beta;
    const double* c = &su[(lr0 + 1) * C::LD + 2 + 2 * tx];
    ft_d2 dn = ft_ld2(c - C::LD), cc = ft_ld2(c);
    double drt = c[-C::LD + 2];                              // row below, x+2: SE neighbour of the odd column
#if defined(__CUDA_ARCH__)
#ifdef FT_NOUNROLL
#pragma unroll 1
#else
#pragma unroll
#endif
#endif
    for (int j = 0; j < C::RPT; ++j, c += C::LD) {
        const int lr = lr0 + j, gy = y0 + lr;
        const ft_d2 up = ft_ld2(c + C::LD);
        const double lf = c[-1], rt = c[2];
        const double nw = LAT == LAT_TRIANGULAR ? c[C::LD - 1] : 0.0;
        const bool valid = interior || (gy >= 1 && gy <= g.n - 2 && gx < g.m);
        const unsigned c01 = *reinterpret_cast<const unsigned short*>(&scf[(lr + 1) * C::CLD + 16 + 2 * tx]);
        unsigned cf0 = c01 & 0xffu, cf1 = c01 >> 8;
        // the neighbours' values as the two sites see them
        double vE0 = cc.y, vW0 = lf, vN0 = up.x, vS0 = dn.x, vE1 = rt, vW1 = cc.x, vN1 = up.y, vS1 = dn.y;
        double vNW0 = nw, vNE0 = up.y, vSW1 = dn.x, vSE1 = drt;
        double mc = 0.0;
        if (C::DEFL) {
            // minus (A Z mu)_i = rho_i mu_i + sum_j w_ij (mu_i - mu_j): every neighbour j enters the stencil as u_j + (mu_i - mu_j).
            // The differences vanish unless the bond crosses a block border or ends on a Dirichlet row (mu = 0 there), so
            // nothing is added -- and nothing rounded -- in the interior of a block.  Record of the row (ft_defl_rec_entry):
            // mu, the north / south differences (uniform over the row), west / east (first / last thread of the row only).
            const double* rec = srec + lr * 8;
            const ft_d2 m01 = ft_ld2(rec);
            mc = m01.x;
            const double dN = m01.y, dS = rec[2];
            vW0 += rec[tx == 0 ? 4 : 3]; vE1 += rec[tx == 63 ? 5 : 3];
            vN0 += dN; vN1 += dN; vS0 += dS; vS1 += dS;
            if (LAT == LAT_TRIANGULAR) {
                // up-type column: NW -> (x-1, y+1), NE -> (x+1, y+1) (the thread's own second column); down-type column:
                // SW -> (x-1, y-1) (the thread's first column), SE -> (x+1, y-1)
                vNW0 += tx == 0 ? rec[6] : dN; vNE0 += dN;
                vSW1 += dS; vSE1 += tx == 63 ? rec[7] : dS;
            }
        }
        unsigned e0, e1;
        double all0, all1;
        if (interior) {
            e0 = ft_interior_ex<LAT>(gx); e1 = ft_interior_ex<LAT>(gx + 1);
            all0 = (vE0 + vW0) + (vN0 + vS0);
            all1 = (vE1 + vW1) + (vN1 + vS1);
            if (LAT == LAT_TRIANGULAR) { all0 += vNW0 + vNE0; all1 += vSW1 + vSE1; }
        } else {
            e0 = valid ? neighbour_bits(g, gx, gy) : 0u; e1 = valid ? neighbour_bits(g, gx + 1, gy) : 0u;
            cf0 &= e0; cf1 &= e1;
            all0 = 0.0; all1 = 0.0;
            if (e0 & NB_E) all0 += vE0;  if (e0 & NB_W) all0 += vW0; if (e0 & NB_N) all0 += vN0;  if (e0 & NB_S) all0 += vS0;
            if (e1 & NB_E) all1 += vE1;  if (e1 & NB_W) all1 += vW1; if (e1 & NB_N) all1 += vN1;  if (e1 & NB_S) all1 += vS1;
            if (LAT == LAT_TRIANGULAR) {
                if (e0 & NB_NW) all0 += vNW0; if (e0 & NB_NE) all0 += vNE0;
                if (e1 & NB_SW) all1 += vSW1; if (e1 & NB_SE) all1 += vSE1;
            }
        }
        double con0 = 0.0, con1 = 0.0;                       // conducting neighbours
        ft_padd(con0, vE0, cf0 & NB_E); ft_padd(con0, vW0, cf0 & NB_W); ft_padd(con0, vN0, cf0 & NB_N); ft_padd(con0, vS0, cf0 & NB_S);
        ft_padd(con1, vE1, cf1 & NB_E); ft_padd(con1, vW1, cf1 & NB_W); ft_padd(con1, vN1, cf1 & NB_N); ft_padd(con1, vS1, cf1 & NB_S);
        if (LAT == LAT_TRIANGULAR) {
            ft_padd(con0, vNW0, cf0 & NB_NW); ft_padd(con0, vNE0, cf0 & NB_NE);
            ft_padd(con1, vSW1, cf1 & NB_SW); ft_padd(con1, vSE1, cf1 & NB_SE);
        }
        // (gx is even: the thread's first column is an up-type site on the triangular lattice, its second a down-type one)
        const int i0 = C::tabp(LAT, ft_pat<LAT>(cf0, 0), lane), i1 = C::tabp(LAT, ft_pat<LAT>(cf1, 1), lane);
        // (boundary tiles: only the sites ON the lattice border lack neighbours; every other site is in the table as well)
        const bool full0 = interior || (e0 == ft_interior_ex<LAT>(gx) && !ft_seam(g, gx)), full1 = interior || (e1 == ft_interior_ex<LAT>(gx + 1) && !ft_seam(g, gx + 1));
        const FtDiag t0 = full0 ? dtab[i0] : ft_diag_site(g, cf0, e0, gx, sc.g0, sc.gleak, cinv);
        const FtDiag t1 = full1 ? dtab[i1] : ft_diag_site(g, cf1, e1, gx + 1, sc.g0, sc.gleak, cinv);
        // off-diagonal part with the weights g0 and gleak THEMSELVES (g0 con + gleak (all - con)): a rounded g0 - gleak would be
        // a bond weight the diagonal does not contain
        double w0 = t0.d * cc.x - (sc.g0 * con0 + sc.gleak * (all0 - con0));
        double w1 = t1.d * cc.y - (sc.g0 * con1 + sc.gleak * (all1 - con1));
        double rho0 = 0.0, rho1 = 0.0;
        if (C::DEFL) {
            if (full0) rho0 = rtab[i0]; else diag_seq_rho(g, cf0, e0, gx, sc.g0, sc.gleak, &rho0);
            if (full1) rho1 = rtab[i1]; else diag_seq_rho(g, cf1, e1, gx + 1, sc.g0, sc.gleak, &rho1);
            w0 -= rho0 * mc; w1 -= rho1 * mc;
        }
        double* sp = &ss[lr * C::LD + 2 + 2 * tx];
        const ft_d2 s2 = ft_ld2(sp), r2 = ft_ld2(&sr[(lr + 1) * C::LD + 2 + 2 * tx]);
        const double sn0 = w0 + beta * s2.x, sn1 = w1 + beta * s2.y;
        double rn0, rn1, un0, un1;
        if (C::USTATE) {                                     // sr / su hold u: u' = u - alpha D^-1 s', r' = d u'
            un0 = cc.x - alpha * (sn0 * t0.inv); un1 = cc.y - alpha * (sn1 * t1.inv);
            rn0 = t0.d * un0; rn1 = t1.d * un1;
        } else {
            rn0 = r2.x - alpha * sn0; rn1 = r2.y - alpha * sn1;
            un0 = rn0 * t0.inv; un1 = rn1 * t1.inv;
        }
        ft_st2(sp, valid ? un0 : 0.0, valid ? un1 : 0.0);
        if (valid && lr < C::TY) {
            const int64_t i = (int64_t)gy * g.m + gx;
            ft_st2(s_out + i, sn0, sn1);
            if (C::USTATE) ft_st2(r_out + i, un0, un1); else ft_st2(r_out + i, rn0, rn1);
            acc_rz += rn0 * un0 + rn1 * un1;
            acc_rr += rn0 * rn0 + rn1 * rn1;
            if (C::DEFL) *acc_ru += rho0 * un0 + rho1 * un1;
            // the rows the read-out consumes: p = (u - Z mu) + beta p, x += alpha p
            if (!INT && (gy == 1 || gy == g.n - 2)) {
                const int64_t o = (gy == 1 ? 0 : g.m) + gx;
                const ft_d2 p2 = ft_ld2(prow + o), x2 = ft_ld2(xrow + o);
                const double p0 = (cc.x - mc) + beta * p2.x, p1 = (cc.y - mc) + beta * p2.y;
                ft_st2(prow + o, p0, p1);
                ft_st2(xrow + o, x2.x + alpha * p0, x2.y + alpha * p1);
            }
        }
        drt = rt;
        dn = cc; cc = up;
    }
}

// east / west ring columns (gx = x0 - 1 and x0 + TX) of the compute rows: u' only (one thread per site).  SIDE (0 west,
// 1 east) is a template parameter: the block-column classes of the site and of its neighbours are then constants.
// Who reads these values: the E bonds of the tile's east column and (triangular) the NW bonds of its west column -- rows
// 0 .. TY-1 -- and the NW bond of the tile's top-left site, which ends on the west ring site of row TY; the east ring
// site of row TY is never read (the last tile column is a down-type column: no NE bond).
template <int LAT, class C, bool INT, int SIDE>
PERC_HD void ft_ring_site(const Geom& g, const FtScalars& sc, const double* sr, double* ss, const uint8_t* scf,
                          const double* su, const FtDiag* dtab, const double* cinv, int x0, int y0, int lr, int lane,
                          const double* sft, const double* rtab)
{
    constexpr bool interior = INT;
    constexpr int col = SIDE ? 2 + C::TX : 1;
    const int gy = y0 + lr;
    int gx = SIDE ? x0 + C::TX : x0 - 1;
    if (!interior && g.pbc) gx = gx < 0 ? g.m - 1 : (gx >= g.m ? 0 : gx);      // (the wrapped column: its values were patched into the box)
    double un = 0.0;
    if (interior || (gy >= 1 && gy <= g.n - 2 && gx >= 0 && gx < g.m)) {
        const unsigned ex = interior ? ft_interior_ex<LAT>(gx) : neighbour_bits(g, gx, gy);
        const unsigned cf = scf[(lr + 1) * C::CLD + (SIDE ? 16 + C::TX : 15)] & ex;
        const double* c = &su[(lr + 1) * C::LD + col];
        double all = 0.0, con = 0.0;
        // (deflation: as in ft_phase_main, neighbour j enters as u_j + (mu_i - mu_j); mu from the shift table by row and
        // block column -- a neighbour in the site's own row and block column needs no look-up)
        const double mi = C::DEFL ? sft[(lr + 1) * 4 + ft_defl_cls<C>(col)] : 0.0;
#define FT_NB(bit, dc, dr) if (interior || (ex & bit)) { double v = c[(dr) * C::LD + (dc)]; \
                                           if (C::DEFL && ((dr) != 0 || ft_defl_cls<C>(col + (dc)) != ft_defl_cls<C>(col))) \
                                               v += mi - sft[(lr + 1 + (dr)) * 4 + ft_defl_cls<C>(col + (dc))]; \
                                           all += v; if (cf & bit) con += v; }
        FT_NB(NB_E, 1, 0) FT_NB(NB_W, -1, 0) FT_NB(NB_N, 0, 1) FT_NB(NB_S, 0, -1)
        if (LAT == LAT_TRIANGULAR) {
            // (x0 - 1 is odd, a down-type column; x0 + TX even, an up-type one -- but only when x0 is even, which TX = 128 makes it)
            if (gx & 1) { FT_NB(NB_SW, -1, -1) FT_NB(NB_SE, 1, -1) } else { FT_NB(NB_NW, -1, 1) FT_NB(NB_NE, 1, 1) }
        }
#undef FT_NB
        const int it = C::tabp(LAT, ft_pat<LAT>(cf, gx & 1), lane);
        const bool full = interior || (ex == ft_interior_ex<LAT>(gx) && !ft_seam(g, gx));
        const FtDiag t = full ? dtab[it] : ft_diag_site(g, cf, ex, gx, sc.g0, sc.gleak, cinv);
        double w = t.d * c[0] - (sc.g0 * con + sc.gleak * (all - con));
        if (C::DEFL) {
            double rho;
            if (full) rho = rtab[it]; else diag_seq_rho(g, cf, ex, gx, sc.g0, sc.gleak, &rho);
            w -= rho * mi;
        }
        const double sn = w + sc.beta * ss[lr * C::LD + col];
        if (C::USTATE) un = c[0] - sc.alpha * (sn * t.inv);
        else un = (sr[(lr + 1) * C::LD + col] - sc.alpha * sn) * t.inv;
    }
    ss[lr * C::LD + col] = un;
}
template <int LAT, class C, bool INT>
PERC_HD void ft_phase_ringcols(const Geom& g, const FtScalars& sc, const double* sr, double* ss, const uint8_t* scf,
                               const double* su, const FtDiag* dtab, const double* cinv, int x0, int y0, int tid, const double* sft = nullptr,
                               const double* rtab = nullptr)
{
    const int lane = tid & 31;
    if (C::V >= 2 && C::TY == 32 && C::RING_NT == 64) {
        // two ring warps: the first takes the west column, the second the east one, lane <-> tile row
        const int rl = tid - C::RING_T0;
        if (rl < 0 || rl >= 64) return;
        if (rl < 32) {
            // (square lattice: the tile owns E and N bonds only -- nothing reads the west ring column)
            if (LAT == LAT_TRIANGULAR) {
                ft_ring_site<LAT, C, INT, 0>(g, sc, sr, ss, scf, su, dtab, cinv, x0, y0, rl, lane, sft, rtab);
                if (rl == 0) ft_ring_site<LAT, C, INT, 0>(g, sc, sr, ss, scf, su, dtab, cinv, x0, y0, C::TY, lane, sft, rtab);
            }
        } else {
            ft_ring_site<LAT, C, INT, 1>(g, sc, sr, ss, scf, su, dtab, cinv, x0, y0, rl - 32, lane, sft, rtab);
        }
        return;
    }
    if (C::V >= 2 && C::TY == 32 && C::RING_NT == 32) {
        // one ring warp: lane <-> tile row of the east column, then (triangular) of the west one
        const int rl = tid - C::RING_T0;
        if (rl < 0 || rl >= 32) return;
        ft_ring_site<LAT, C, INT, 1>(g, sc, sr, ss, scf, su, dtab, cinv, x0, y0, rl, lane, sft, rtab);
        if (LAT == LAT_TRIANGULAR) {
            ft_ring_site<LAT, C, INT, 0>(g, sc, sr, ss, scf, su, dtab, cinv, x0, y0, rl, lane, sft, rtab);
            if (rl == 0) ft_ring_site<LAT, C, INT, 0>(g, sc, sr, ss, scf, su, dtab, cinv, x0, y0, C::TY, lane, sft, rtab);
        }
        return;
    }
    for (int q = tid - C::RING_T0; q >= 0 && q < 2 * C::CR; q += C::RING_NT) {
        const int side = q >= C::CR, lr = q - side * C::CR;
        if ((lr == C::TY && side) || (LAT == LAT_SQUARE && !side)) continue;       // never read
        if (side) ft_ring_site<LAT, C, INT, 1>(g, sc, sr, ss, scf, su, dtab, cinv, x0, y0, lr, lane, sft, rtab);
        else ft_ring_site<LAT, C, INT, 0>(g, sc, sr, ss, scf, su, dtab, cinv, x0, y0, lr, lane, sft, rtab);
    }
}

// ---- phase E: u'.A u' as the energy of the bonds OWNED by the tile's sites (E, N, NW, NE), u' from shared
// memory (zeros on Dirichlet rows and outside the lattice) ---------------------------------------------------
template <int LAT, class C, bool INT>
PERC_HD void ft_phase_energy(const Geom& g, const FtScalars& sc, const double* ss, const uint8_t* scf, int x0, int y0,
                             int tid, double& acc_e)
{
    if (tid >= C::MAIN_THREADS) return;
    constexpr bool interior = INT;
    const int tx = tid & 63, ty = tid >> 6;
    const int gx = x0 + 2 * tx, lr0 = ty * C::RPT;
#if defined(__CUDA_ARCH__)
#ifdef FT_NOUNROLL
#pragma unroll 1
#else
#pragma unroll
#endif
#endif
    for (int j = 0; j < C::RPT; ++j) {
        const int lr = lr0 + j, gy = y0 + lr;
        if (lr >= C::TY) continue;
        if (!interior && !(gy >= 0 && gy < g.n && gx < g.m)) continue;
        const double* c = &ss[lr * C::LD + 2 + 2 * tx];
        const ft_d2 cc = ft_ld2(c), up = ft_ld2(c + C::LD);
        const double rt = c[2];
        const unsigned c01 = *reinterpret_cast<const unsigned short*>(&scf[(lr + 1) * C::CLD + 16 + 2 * tx]);
        const unsigned cf0 = c01 & 0xffu, cf1 = c01 >> 8;
        const double dE0 = cc.x - cc.y, dE1 = cc.y - rt, dN0 = cc.x - up.x, dN1 = cc.y - up.y;
        double all = 0.0, con = 0.0;
        if (interior) {
            all = (dE0 * dE0 + dE1 * dE1) + (dN0 * dN0 + dN1 * dN1);
            ft_padd(con, dE0 * dE0, cf0 & NB_E); ft_padd(con, dE1 * dE1, cf1 & NB_E);
            ft_padd(con, dN0 * dN0, cf0 & NB_N); ft_padd(con, dN1 * dN1, cf1 & NB_N);
            if (LAT == LAT_TRIANGULAR) {
                const double dNW = cc.x - c[C::LD - 1], dNE = cc.x - up.y;
                all += dNW * dNW + dNE * dNE;
                ft_padd(con, dNW * dNW, cf0 & NB_NW); ft_padd(con, dNE * dNE, cf0 & NB_NE);
            }
        } else {
            const unsigned e0 = neighbour_bits(g, gx, gy), e1 = neighbour_bits(g, gx + 1, gy);
            const bool rowE = gy >= 1 && gy <= g.n - 2;      // an E bond joins two sites of one row
            if (rowE && (e0 & NB_E)) { all += dE0 * dE0; ft_padd(con, dE0 * dE0, cf0 & NB_E); }
            if (rowE && (e1 & NB_E)) { all += dE1 * dE1; ft_padd(con, dE1 * dE1, cf1 & NB_E); }
            if (e0 & NB_N) { all += dN0 * dN0; ft_padd(con, dN0 * dN0, cf0 & NB_N); }
            if (e1 & NB_N) { all += dN1 * dN1; ft_padd(con, dN1 * dN1, cf1 & NB_N); }
            if (LAT == LAT_TRIANGULAR) {
                const double dNW = cc.x - c[C::LD - 1], dNE = cc.x - up.y;
                if (e0 & NB_NW) { all += dNW * dNW; ft_padd(con, dNW * dNW, cf0 & NB_NW); }
                if (e0 & NB_NE) { all += dNE * dNE; ft_padd(con, dNE * dNE, cf0 & NB_NE); }
            }
        }
        acc_e += sc.g0 * con + sc.gleak * (all - con);
    }
}

// ---- the scalar recurrences, run once per iteration by whoever holds the three lattice-wide sums -------------
struct FtState { double gamma, alpha, beta, bnrm, err, rr, tol; int iter, itmax, done; };

// prime = 1: the pass that only forms s = A u0 and delta0 (alpha = beta = 0 on entry): alpha0 = gamma0 / delta0
PERC_HD void ft_scalar_step(FtState& st, double rz, double rr, double en, int prime)
{
    if (prime) { st.gamma = rz; st.alpha = rz / en; st.beta = 0.0; st.rr = rr; return; }
    const int it = st.iter + 1;
    const double err = sqrt(rr) / st.bnrm;
    st.iter = it; st.err = err; st.rr = rr;
    const double beta = rz / st.gamma;
    st.alpha = rz / (en - beta * rz / st.alpha);
    st.beta = beta; st.gamma = rz;
    if (!(err > st.tol) || it > st.itmax) st.done = 1;       // linbcg's loop guard iter <= itmax (Sq/bondc.f:780)
}

}  // namespace perc
