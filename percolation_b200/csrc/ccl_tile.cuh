// ccl_tile.cuh -- the cluster-labeling algorithm, written once for device and host.
//
// Replaces the incremental fills of the reference (Sq/site.f:162-289, Sq/bond.f:165-369,
// Sq/sitebond.f:187-400, Sq/bondsite.f:182-354) by a static, bit-parallel, run-based
// connected-component labeling.  Every function here is __host__ __device__: ccl.cu wraps the
// phases into kernels (one CUDA thread per 32-site word of a 128x64 tile); tests/ccl_emul.cpp
// compiles the SAME source with g++ and runs the phases thread by thread, so the bit logic is
// checked against the oracle on a machine without a GPU.
//
//   phase 0  mask bytes -> bit planes S (site), E, N, NW, NE (owned occupied bonds) in shared memory
//   phase 1  horizontal runs inside each 32-bit word; a run start is a union-find NODE
//   phase 2  unions: run continues into the next word; N / NW / NE bonds into the row above,
//            one union per overlapping pair of runs (redundant columns are masked out bitwise),
//            done hierarchically (row pairs, then 4-row groups, ...) so the trees stay shallow
//   phase 3  node -> root, per-root size (weights by problem kind, popcounts of bit planes)
//   phase 4  provisional labels (global index of the tile-local root + 1), root sizes, root list
// then, on the global table (label = parent + 1):
//   merge    unions across tile borders and the periodic wrap (same run-pair pruning)
//   rootfix  every tile-local root -> its global root, sizes folded into the global root
//   flatten  every site: one hop to the global root
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include "geometry.cuh"

#ifdef __CUDA_ARCH__
#define PERC_DEV 1
#else
#define PERC_DEV 0
#endif

namespace perc {

constexpr int CT_TW = 128, CT_TH = 64;          // tile, sites
constexpr int CT_NW = CT_TW / 32;               // 32-bit words per tile row
constexpr int CT_PR = CT_TH + 2;                // plane rows incl. one halo row below / above
constexpr int CT_THREADS = CT_NW * CT_TH;       // one thread per word
constexpr int MAX_SPAN = 4096;

// device-side summary written by the labeling pipeline
struct Summary {
    unsigned long long ncl;        // clusters with a site-id label
    unsigned long long nlone;      // mixed problem: occupied bonds with no occupied end (size-1 clusters)
    unsigned long long maxpack;    // (size << 32) | (0xffffffff - label)  -> max size, then min label
    unsigned long long nocc_sites; // occupied sites in the mask
    unsigned long long nocc_bonds; // occupied bonds in the mask
    unsigned long long span_best;  // ((0xffffffff - id) << 32) | size of the spanning cluster with the SMALLEST id (0: none);
                                   // does not depend on the order or the capacity of the list below
    unsigned int nroots;           // tile-local roots appended to the root list
    int nspan;                     // spanning clusters found
    int span_overflow;             // more than MAX_SPAN spanning clusters: the list holds MAX_SPAN of them (always incl. the smallest id)
    int pad;
    int32_t span_ids[MAX_SPAN];
    int32_t span_sizes[MAX_SPAN];
};

struct TileSmem {
    int lab[CT_TH * CT_TW];                     // union-find parents (node positions only); after phase 3
                                                // re-used as the per-site root id staged for the coalesced output
    uint32_t cnt[CT_TH * CT_TW / 2];            // per-root sizes, two 16-bit counters per word (a tile-local cluster
                                                // weighs at most 8192 sites x 7 = 57344 < 2^16)
    uint32_t pS[CT_PR][CT_NW], pE[CT_PR][CT_NW], pN[CT_PR][CT_NW];
    union {
        uint32_t pC[CT_TH][CT_NW];              // phases 1-2: bit x: site x joined to site x-1 (inside the tile)
        uint32_t ringbits[CT_TH * CT_TW / 32];  // phases 3-4: bit per root: the cluster touches the tile's border ring
    };
    uint32_t pT[CT_TH][CT_NW];                  // node starts
    uint8_t hL[CT_PR + 2], hR[CT_PR + 2];       // mask bytes left / right of the tile's columns (wrap-aware)
    unsigned long long best;                    // largest closed cluster of the tile, packed like Summary::maxpack
    unsigned lone, nroot, rootbase, nclosed;
    // triangular lattice only -- kept last: the square kernels allocate the struct without them and fit
    // four CTAs per SM
    uint32_t pNW[CT_PR][CT_NW], pNE[CT_PR][CT_NW];
};

template <int LAT> constexpr size_t tile_smem_bytes() { return LAT == LAT_TRIANGULAR ? sizeof(TileSmem) : offsetof(TileSmem, pNW); }

struct TileRegs { uint32_t rootbits; int nloc, off; };   // per-thread values that live across phases


// per-root counters: 16-bit halves of 32-bit words (no carry can cross: every total stays below 2^16)
PERC_HD void cnt_add(TileSmem& s, int root, int w)
{
#if PERC_DEV
    atomicAdd(&s.cnt[root >> 1], (uint32_t)w << ((root & 1) * 16));
#else
    s.cnt[root >> 1] += (uint32_t)w << ((root & 1) * 16);
#endif
}
PERC_HD int cnt_get(const TileSmem& s, int root) { return (int)((s.cnt[root >> 1] >> ((root & 1) * 16)) & 0xffffu); }
PERC_HD void ring_set(TileSmem& s, int root)
{
#if PERC_DEV
    atomicOr(&s.ringbits[root >> 5], 1u << (root & 31));
#else
    s.ringbits[root >> 5] |= 1u << (root & 31);
#endif
}
PERC_HD bool ring_get(const TileSmem& s, int root) { return (s.ringbits[root >> 5] >> (root & 31)) & 1u; }

// ---- bit helpers ----------------------------------------------------------------------------
PERC_HD int hibit(uint32_t v)
{
#if PERC_DEV
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}
PERC_HD int lobit(uint32_t v)
{
#if PERC_DEV
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
PERC_HD int popc32(uint32_t v)
{
#if PERC_DEV
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
// warp-uniform trip count for per-lane loops over set bits: all lanes run max-over-the-warp iterations
// and re-converge at the top of every iteration (a divergent `while (bits)` loop lets the lanes drift
// apart for good -- measured 7 of 32 lanes active)
PERC_HD int warp_max_count(int n)
{
#if PERC_DEV
    return __reduce_max_sync(0xffffffffu, n);
#else
    return n;
#endif
}
PERC_HD void warp_converge()
{
#if PERC_DEV
    __syncwarp();
#endif
}
PERC_HD uint32_t le_mask(int b) { return 0xFFFFFFFFu >> (31 - b); }      // bits 0..b

// 4 mask bytes -> 4 plane bits (bit k of each byte, byte order = site order)
PERC_HD uint32_t nib(uint32_t w, int k) { return ((((w >> k) & 0x01010101u) * 0x01020408u) >> 24) & 0xFu; }

template <int LAT>
PERC_HD void planes_from_words(const uint32_t q[8], uint32_t& S, uint32_t& E, uint32_t& N, uint32_t& NW, uint32_t& NE)
{
    S = E = N = NW = NE = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t w = q[k];
        S |= nib(w, 0) << (4 * k);
        E |= nib(w, 1) << (4 * k);
        N |= nib(w, 2) << (4 * k);
        if (LAT == LAT_TRIANGULAR) { NW |= nib(w, 3) << (4 * k); NE |= nib(w, 4) << (4 * k); }
    }
}

// planes of the 32 sites (gx0 .. gx0+31, gy); columns >= m read as 0.  vec: rows are 16-byte aligned
template <int LAT>
PERC_HD void load_planes(const Geom& g, const uint8_t* __restrict__ mask, int gx0, int gy, bool vec,
                         uint32_t& S, uint32_t& E, uint32_t& N, uint32_t& NW, uint32_t& NE)
{
    S = E = N = NW = NE = 0;
    if (gy < 0 || gy >= g.n || gx0 >= g.m) return;
    const uint8_t* p = mask + (int64_t)gy * g.m + gx0;
    if (vec && gx0 + 32 <= g.m) {
        uint32_t q[8];
#if PERC_DEV
        uint4 a = __ldg(reinterpret_cast<const uint4*>(p)), b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
        q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
#else
        memcpy(q, p, 32);
#endif
        planes_from_words<LAT>(q, S, E, N, NW, NE);
        return;
    }
    int lim = g.m - gx0 < 32 ? g.m - gx0 : 32;
    for (int k = 0; k < lim; ++k) {
        uint32_t v = p[k];
        S |= (v & 1u) << k; E |= ((v >> 1) & 1u) << k; N |= ((v >> 2) & 1u) << k;
        if (LAT == LAT_TRIANGULAR) { NW |= ((v >> 3) & 1u) << k; NE |= ((v >> 4) & 1u) << k; }
    }
}

// ---- union-find on the tile (parents point to smaller indices; the root is the minimum) -------
PERC_HD int tile_find(int* lab_, int a)
{
    volatile int* lab = lab_;
    for (;;) {
        int p = lab[a];
        if (p == a) return a;
        int gp = lab[p];
        if (gp == p) return p;
        lab[a] = gp;            // path halving; races benignly with atomicMin (always an ancestor)
        a = gp;
    }
}

// read-only find (no path-halving stores): used once the owner threads start writing roots
PERC_HD int tile_find_ro(const int* lab_, int a)
{
    const volatile int* lab = lab_;
    for (;;) {
        int p = lab[a];
        if (p == a) return a;
        a = p;
    }
}

// two read-only finds walked in lockstep (the two shared-memory loads of a step are independent)
PERC_HD void tile_find_ro2(const int* lab_, int a, int b, int& ra, int& rb)
{
    const volatile int* lab = lab_;
    int pa = lab[a], pb = lab[b];
    while (pa != a || pb != b) {
        a = pa; b = pb;
        pa = lab[a]; pb = lab[b];
    }
    ra = a; rb = b;
}

PERC_HD void tile_unite(int* lab, int a, int b)
{
    for (;;) {
        a = tile_find(lab, a);
        b = tile_find(lab, b);
        if (a == b) return;
        if (a < b) { int tmp = a; a = b; b = tmp; }
#if PERC_DEV
        int old = atomicMin(&lab[a], b);
#else
        int old = lab[a]; if (b < old) lab[a] = b;
#endif
        if (old == a) return;
        a = old;
    }
}

// node (run start) that holds the ACTIVE site (row ly, column col) of the tile
PERC_HD int node_of(const TileSmem& s, int ly, int col)
{
    int w = col >> 5, b = col & 31;
    return ly * CT_TW + (w << 5) + hibit(s.pT[ly][w] & le_mask(b));
}

// ---- phase 0 -------------------------------------------------------------------------------
template <int LAT, int KIND>
PERC_HD void tile_phase0(TileSmem& s, const Geom& g, const uint8_t* __restrict__ mask, int x0, int y0, int tid, bool vec)
{
    for (int j = tid; j < CT_PR * CT_NW; j += CT_THREADS) {
        int pr = j / CT_NW, w = j % CT_NW;
        uint32_t S = 0, E = 0, N = 0, NW = 0, NE = 0;
        bool halo = pr == 0 || pr == CT_PR - 1;
        if (!halo || KIND == KIND_MIXED) load_planes<LAT>(g, mask, x0 + 32 * w, y0 + pr - 1, vec, S, E, N, NW, NE);
        s.pS[pr][w] = S; s.pE[pr][w] = E; s.pN[pr][w] = N;
        if (LAT == LAT_TRIANGULAR) { s.pNW[pr][w] = NW; s.pNE[pr][w] = NE; }
    }
    if (tid < 2 * CT_PR) {
        int side = tid >= CT_PR, pr = tid - side * CT_PR;
        uint8_t v = 0;
        if (KIND == KIND_MIXED) {
            int gy = y0 + pr - 1;
            int xe = g.m - x0 < CT_TW ? g.m - x0 : CT_TW;
            int gx = side ? x0 + xe : x0 - 1;
            if (gx < 0) gx = g.pbc ? g.m - 1 : -1;
            if (gx >= g.m) gx = g.pbc ? 0 : -1;
            if (gx >= 0 && gy >= 0 && gy < g.n) v = mask[(int64_t)gy * g.m + gx];
        }
        if (side) s.hR[pr] = v; else s.hL[pr] = v;
    }
    if (tid == 0) { s.lone = 0; s.nroot = 0; s.rootbase = 0; s.nclosed = 0; s.best = 0; }
}

// ---- phase 1: runs --------------------------------------------------------------------------
PERC_HD void tile_phase1(TileSmem& s, int tid)
{
    const int w = tid % CT_NW, ly = tid / CT_NW, pr = ly + 1;
    const uint32_t S = s.pS[pr][w], E = s.pE[pr][w];
    uint32_t connE = S & E & (S >> 1);                            // bit x: x joined to x+1 (same word)
    if (w + 1 < CT_NW) connE |= S & E & (s.pS[pr][w + 1] << 31);
    uint32_t carry = 0;
    if (w > 0) carry = ((s.pS[pr][w - 1] & s.pE[pr][w - 1]) >> 31) & S & 1u;
    const uint32_t connL = (connE << 1) | carry;
    const uint32_t T = S & (~connL | 1u);                         // bit 0 is always a node if active
    s.pC[ly][w] = connL;
    s.pT[ly][w] = T;
    const int base = ly * CT_TW + (w << 5);
    for (uint32_t t = T; t; t &= t - 1) {
        int a = lobit(t);
        s.lab[base + a] = base + a;
        s.cnt[(base + a) >> 1] = 0;      // both halves of the word belong to this thread's 32 sites
    }
}

// ---- phase 2: unions --------------------------------------------------------------------------
// level 0: a run that continues into the next 32-bit word (one thread per word)
PERC_HD void tile_phase2_words(TileSmem& s, int tid)
{
    const int w = tid % CT_NW, ly = tid / CT_NW;
    if (w > 0 && (s.pC[ly][w] & 1u)) tile_unite(s.lab, ly * CT_TW + (w << 5), node_of(s, ly, (w << 5) - 1));
}

// level k = 1 .. log2(TH): bonds N / NW / NE between row ly and ly + 1 where ly + 1 is an odd multiple
// of 2^(k-1), i.e. the two halves of every 2^k-row group are joined.  Trees stay shallow (depth grows
// by ~1 per level instead of ~1 per row), and the TH/2^k active rows are spread over all threads:
// 2^k threads share one word, each takes a 32/2^k-bit slice of its columns.
constexpr int CT_LEVELS = 6;
static_assert((1 << CT_LEVELS) == CT_TH, "CT_LEVELS = log2(CT_TH)");

template <int LAT, int K>
PERC_HD void tile_phase2_level(TileSmem& s, int tid)
{
    constexpr int SH = K < 5 ? K : 5, NSL = 1 << SH, WIDTH = 32 / NSL;
    const int item = tid >> SH, sl = tid & (NSL - 1);
    if (item >= (CT_TH >> K) * CT_NW) return;
    const int w = item % CT_NW, ly = (((item / CT_NW) * 2 + 1) << (K - 1)) - 1, pr = ly + 1;
    const uint32_t slice = (WIDTH == 32 ? 0xFFFFFFFFu : ((1u << WIDTH) - 1u)) << (sl * WIDTH);
    const int base = ly * CT_TW + (w << 5);
    const uint32_t S = s.pS[pr][w], connL = s.pC[ly][w], T = s.pT[ly][w];
    const uint32_t Su = s.pS[pr + 1][w], Cu = s.pC[ly + 1][w], Tu = s.pT[ly + 1][w];
    const uint32_t v = S & s.pN[pr][w] & Su;
    // a column is redundant when the column to its left makes the same union (both rows continue a run)
    for (uint32_t need = v & ~((v << 1) & connL & Cu) & slice; need; need &= need - 1) {
        int x = lobit(need);
        tile_unite(s.lab, base + hibit(T & le_mask(x)), base + CT_TW + hibit(Tu & le_mask(x)));
    }
    if (LAT == LAT_TRIANGULAR) {
        uint32_t Sul = Su << 1, Sur = Su >> 1, Cur = Cu >> 1;
        if (w > 0) Sul |= s.pS[pr + 1][w - 1] >> 31;
        if (w + 1 < CT_NW) { Sur |= s.pS[pr + 1][w + 1] << 31; Cur |= s.pC[ly + 1][w + 1] << 31; }
        // NW / NE are redundant when the N bond exists and the upper row joins the two columns
        for (uint32_t need = (S & s.pNW[pr][w] & Sul) & ~(v & Cu) & slice; need; need &= need - 1) {
            int x = lobit(need);
            tile_unite(s.lab, base + hibit(T & le_mask(x)), node_of(s, ly + 1, (w << 5) + x - 1));
        }
        for (uint32_t need = (S & s.pNE[pr][w] & Sur) & ~(v & Cur) & slice; need; need &= need - 1) {
            int x = lobit(need);
            tile_unite(s.lab, base + hibit(T & le_mask(x)), node_of(s, ly + 1, (w << 5) + x + 1));
        }
    }
}

// between the last union level and phase 3 (separate barrier on both sides): the run-connectivity plane is
// dead, its storage becomes the ring bitmap
PERC_HD void tile_clear_ring(TileSmem& s, int tid)
{
    for (int k = tid; k < CT_TH * CT_TW / 32; k += CT_THREADS) s.ringbits[k] = 0;
}

// ---- phase 3: node -> root, sizes -------------------------------------------------------------
// Size weights: site problem = sites; bond problem = bonds (counted at their owner
// site); mixed (Sq/sitebond.f:231-305) = site + owned occupied bonds + incoming occupied bonds whose
// owner site is unoccupied (dangling onto this site).  A bond with no occupied end is a lone
// size-1 cluster (Sq/sitebond.f:231-242), counted into s.lone.
// VAR = 0: the first form (per-site roots staged per run); VAR = 1: the staging is skipped, tile_phase4_labels<1> derives
// each site's root from its run start instead.  The library runs VAR = 2 (tile_phase3_pair below: VAR = 1 with two runs
// per trip of the per-run loop; on the GPU 0.166 ms against 0.197 for VAR = 0 at L = 4096 mixed); 0 and 1 are kept for the
// host emulation's cross-check (tests/test_ccl_emulation.py).
template <int LAT, int KIND, int VAR = 0>
PERC_HD void tile_phase3(TileSmem& s, const Geom& g, int x0, int y0, int tid, TileRegs& r)
{
    const int w = tid % CT_NW, ly = tid / CT_NW, pr = ly + 1;
    const int base = ly * CT_TW + (w << 5);
    const uint32_t S = s.pS[pr][w], T = s.pT[ly][w];
    const uint32_t E = s.pE[pr][w], N = s.pN[pr][w];
    const uint32_t NW = LAT == LAT_TRIANGULAR ? s.pNW[pr][w] : 0u, NE = LAT == LAT_TRIANGULAR ? s.pNE[pr][w] : 0u;
    uint32_t inW = 0, inS = 0, inSW = 0, inSE = 0;
    // slab handles: only the rows this rank owns are counted (halo rows are the neighbour's)
    const bool owned = y0 + ly >= g.own_lo && y0 + ly < g.own_hi;
    if (KIND == KIND_MIXED && owned) {
        const int xe = g.m - x0 < CT_TW ? g.m - x0 : CT_TW;
        const int wl = (xe - 1) >> 5, bl = (xe - 1) & 31;          // word / bit of the last real column
        auto dang = [](uint8_t v, unsigned bit) -> uint32_t { return ((v & bit) && !(v & MASK_SITE)) ? 1u : 0u; };
        uint32_t c = w > 0 ? (s.pE[pr][w - 1] & ~s.pS[pr][w - 1]) >> 31 : dang(s.hL[pr], MASK_E);
        inW = ((E & ~S) << 1) | c;
        inS = s.pN[pr - 1][w] & ~s.pS[pr - 1][w];
        const uint32_t Sup = s.pS[pr + 1][w];
        uint32_t Sright = S >> 1;
        if (w + 1 < CT_NW) Sright |= s.pS[pr][w + 1] << 31;
        if (w == wl && (s.hR[pr] & MASK_SITE)) Sright |= 1u << bl;
        unsigned lone = popc32(E & ~S & ~Sright) + popc32(N & ~S & ~Sup);
        if (LAT == LAT_TRIANGULAR) {
            uint32_t dne = s.pNE[pr - 1][w] & ~s.pS[pr - 1][w];
            c = w > 0 ? (s.pNE[pr - 1][w - 1] & ~s.pS[pr - 1][w - 1]) >> 31 : dang(s.hL[pr - 1], MASK_NE);
            inSW = (dne << 1) | c;
            uint32_t dnw = s.pNW[pr - 1][w] & ~s.pS[pr - 1][w];
            inSE = dnw >> 1;
            if (w + 1 < CT_NW) inSE |= (s.pNW[pr - 1][w + 1] & ~s.pS[pr - 1][w + 1]) << 31;
            if (w == wl && dang(s.hR[pr - 1], MASK_NW)) inSE |= 1u << bl;
            uint32_t Supl = Sup << 1, Supr = Sup >> 1;
            if (w > 0) Supl |= s.pS[pr + 1][w - 1] >> 31; else if (s.hL[pr + 1] & MASK_SITE) Supl |= 1u;
            if (w + 1 < CT_NW) Supr |= s.pS[pr + 1][w + 1] << 31;
            if (w == wl && (s.hR[pr + 1] & MASK_SITE)) Supr |= 1u << bl;
            lone += popc32(NW & ~S & ~Supl) + popc32(NE & ~S & ~Supr);
        }
        if (lone) {
#if PERC_DEV
            atomicAdd(&s.lone, lone);
#else
            s.lone += lone;
#endif
        }
    }
    // border ring of the tile: only clusters with a site on it can continue into another tile
    uint32_t ring = 0;
    {
        const int xe = g.m - x0 < CT_TW ? g.m - x0 : CT_TW;
        if (ly == 0 || ly == CT_TH - 1) ring = 0xFFFFFFFFu;
        if (w == 0) ring |= 1u;
        if (w == CT_NW - 1) ring |= 0x80000000u;
        if (g.pbc && w == ((xe - 1) >> 5)) ring |= 1u << ((xe - 1) & 31);      // wrap column of a partial tile
    }
    uint32_t rootbits = 0;
    int prev = -1, acc = 0;
    bool accring = false;
    auto flush = [&]() {
        if (prev < 0) return;
        if (acc) cnt_add(s, prev, acc);
        if (accring) ring_set(s, prev);
    };
    uint32_t t = T;
    for (int it = warp_max_count(popc32(T)); it > 0; --it) {
        warp_converge();
        if (!t) continue;
        const int a = lobit(t);
        t &= t - 1;
        const uint32_t seg = S & (t ? ((1u << lobit(t)) - 1u) : 0xFFFFFFFFu) & ~((1u << a) - 1u);   // sites of this node
        // read-only find: a path-halving store of another thread could overwrite the root written here
        const int root = tile_find_ro(s.lab, base + a);
        s.lab[base + a] = root;
        // spread the root over the other sites of the run: their entries of s.lab are no union-find nodes
        // (no find ever reads them), they stage the per-site root for the coalesced label output
        if (VAR == 0)
            for (uint32_t sg = seg & ~(1u << a); sg; sg &= sg - 1) s.lab[base + lobit(sg)] = root;
        if (root == base + a) rootbits |= 1u << a;
        int wgt;
        if (!owned) wgt = 0;
        else if (KIND == KIND_SITE) wgt = popc32(seg);
        else {
            wgt = popc32(seg & E) + popc32(seg & N);
            if (LAT == LAT_TRIANGULAR) wgt += popc32(seg & NW) + popc32(seg & NE);
            if (KIND == KIND_MIXED) {
                wgt += popc32(seg) + popc32(seg & inW) + popc32(seg & inS);
                if (LAT == LAT_TRIANGULAR) wgt += popc32(seg & inSW) + popc32(seg & inSE);
            }
        }
        if (root != prev) { flush(); prev = root; acc = 0; accring = false; }     // neighbouring runs often share a root
        acc += wgt;
        accring |= (seg & ring) != 0;
    }
    flush();
    r.rootbits = rootbits;
}

// VAR = 2 of phase 3 (the form the library runs): tile_phase3<.., 1> with a per-run loop that takes two runs per trip
// (two find chains in flight)
template <int LAT, int KIND>
PERC_HD void tile_phase3_pair(TileSmem& s, const Geom& g, int x0, int y0, int tid, TileRegs& r)
{
    const int w = tid % CT_NW, ly = tid / CT_NW, pr = ly + 1;
    const int base = ly * CT_TW + (w << 5);
    const uint32_t S = s.pS[pr][w], T = s.pT[ly][w];
    const uint32_t E = s.pE[pr][w], N = s.pN[pr][w];
    const uint32_t NW = LAT == LAT_TRIANGULAR ? s.pNW[pr][w] : 0u, NE = LAT == LAT_TRIANGULAR ? s.pNE[pr][w] : 0u;
    uint32_t inW = 0, inS = 0, inSW = 0, inSE = 0;
    // slab handles: only the rows this rank owns are counted (halo rows are the neighbour's)
    const bool owned = y0 + ly >= g.own_lo && y0 + ly < g.own_hi;
    if (KIND == KIND_MIXED && owned) {
        const int xe = g.m - x0 < CT_TW ? g.m - x0 : CT_TW;
        const int wl = (xe - 1) >> 5, bl = (xe - 1) & 31;          // word / bit of the last real column
        auto dang = [](uint8_t v, unsigned bit) -> uint32_t { return ((v & bit) && !(v & MASK_SITE)) ? 1u : 0u; };
        uint32_t c = w > 0 ? (s.pE[pr][w - 1] & ~s.pS[pr][w - 1]) >> 31 : dang(s.hL[pr], MASK_E);
        inW = ((E & ~S) << 1) | c;
        inS = s.pN[pr - 1][w] & ~s.pS[pr - 1][w];
        const uint32_t Sup = s.pS[pr + 1][w];
        uint32_t Sright = S >> 1;
        if (w + 1 < CT_NW) Sright |= s.pS[pr][w + 1] << 31;
        if (w == wl && (s.hR[pr] & MASK_SITE)) Sright |= 1u << bl;
        unsigned lone = popc32(E & ~S & ~Sright) + popc32(N & ~S & ~Sup);
        if (LAT == LAT_TRIANGULAR) {
            uint32_t dne = s.pNE[pr - 1][w] & ~s.pS[pr - 1][w];
            c = w > 0 ? (s.pNE[pr - 1][w - 1] & ~s.pS[pr - 1][w - 1]) >> 31 : dang(s.hL[pr - 1], MASK_NE);
            inSW = (dne << 1) | c;
            uint32_t dnw = s.pNW[pr - 1][w] & ~s.pS[pr - 1][w];
            inSE = dnw >> 1;
            if (w + 1 < CT_NW) inSE |= (s.pNW[pr - 1][w + 1] & ~s.pS[pr - 1][w + 1]) << 31;
            if (w == wl && dang(s.hR[pr - 1], MASK_NW)) inSE |= 1u << bl;
            uint32_t Supl = Sup << 1, Supr = Sup >> 1;
            if (w > 0) Supl |= s.pS[pr + 1][w - 1] >> 31; else if (s.hL[pr + 1] & MASK_SITE) Supl |= 1u;
            if (w + 1 < CT_NW) Supr |= s.pS[pr + 1][w + 1] << 31;
            if (w == wl && (s.hR[pr + 1] & MASK_SITE)) Supr |= 1u << bl;
            lone += popc32(NW & ~S & ~Supl) + popc32(NE & ~S & ~Supr);
        }
        if (lone) {
#if PERC_DEV
            atomicAdd(&s.lone, lone);
#else
            s.lone += lone;
#endif
        }
    }
    // border ring of the tile: only clusters with a site on it can continue into another tile
    uint32_t ring = 0;
    {
        const int xe = g.m - x0 < CT_TW ? g.m - x0 : CT_TW;
        if (ly == 0 || ly == CT_TH - 1) ring = 0xFFFFFFFFu;
        if (w == 0) ring |= 1u;
        if (w == CT_NW - 1) ring |= 0x80000000u;
        if (g.pbc && w == ((xe - 1) >> 5)) ring |= 1u << ((xe - 1) & 31);      // wrap column of a partial tile
    }
    uint32_t rootbits = 0;
    int prev = -1, acc = 0;
    bool accring = false;
    auto flush = [&]() {
        if (prev < 0) return;
        if (acc) cnt_add(s, prev, acc);
        if (accring) ring_set(s, prev);
    };
    {
        // two runs per trip (half the warp-uniform trips, two find chains in flight); per-site roots derived in the
        // label phase as with VAR = 1.  Same roots, counters and ring bits as tile_phase3's loop.
        auto account = [&](int a, uint32_t seg, int root) {
            s.lab[base + a] = root;
            if (root == base + a) rootbits |= 1u << a;
            int wgt;
            if (!owned) wgt = 0;
            else if (KIND == KIND_SITE) wgt = popc32(seg);
            else {
                wgt = popc32(seg & E) + popc32(seg & N);
                if (LAT == LAT_TRIANGULAR) wgt += popc32(seg & NW) + popc32(seg & NE);
                if (KIND == KIND_MIXED) {
                    wgt += popc32(seg) + popc32(seg & inW) + popc32(seg & inS);
                    if (LAT == LAT_TRIANGULAR) wgt += popc32(seg & inSW) + popc32(seg & inSE);
                }
            }
            if (root != prev) { flush(); prev = root; acc = 0; accring = false; }
            acc += wgt;
            accring |= (seg & ring) != 0;
        };
        uint32_t t2 = T;
        for (int it = warp_max_count((popc32(T) + 1) >> 1); it > 0; --it) {
            warp_converge();
            if (!t2) continue;
            const int a1 = lobit(t2);
            t2 &= t2 - 1;
            const uint32_t seg1 = S & (t2 ? ((1u << lobit(t2)) - 1u) : 0xFFFFFFFFu) & ~((1u << a1) - 1u);
            int a2 = -1;
            uint32_t seg2 = 0;
            if (t2) {
                a2 = lobit(t2);
                t2 &= t2 - 1;
                seg2 = S & (t2 ? ((1u << lobit(t2)) - 1u) : 0xFFFFFFFFu) & ~((1u << a2) - 1u);
            }
            int r1, r2;
            tile_find_ro2(s.lab, base + a1, base + (a2 >= 0 ? a2 : a1), r1, r2);
            account(a1, seg1, r1);
            if (a2 >= 0) account(a2, seg2, r2);
        }
        flush();
        r.rootbits = rootbits;
    }
}

// ---- phase 4 --------------------------------------------------------------------------------
PERC_HD int32_t tile_global_label(const Geom& g, int x0, int y0, int node)
{
    return (y0 + node / CT_TW) * g.m + x0 + (node % CT_TW) + 1;       // t < 2^31 (check_geom)
}

// Every thread sorts its tile-local roots into CLOSED clusters (no site on the border ring: final, counted
// here) and border roots (root list).  Runs after the barrier that completes the per-root sizes.
PERC_HD void tile_phase4_fill(TileSmem& s, const Geom& g, int x0, int y0, int tid, TileRegs& r, int32_t* __restrict__ size)
{
    const int w = tid % CT_NW, ly = tid / CT_NW;
    const int base = ly * CT_TW + (w << 5);
    int nloc = 0;
    unsigned nclosed = 0;
    unsigned long long best = 0;
    for (uint32_t t = r.rootbits; t; t &= t - 1) {
        const int node = base + lobit(t);
        if (ring_get(s, node)) { ++nloc; continue; }
        const int c = cnt_get(s, node);
        const int32_t gl = tile_global_label(g, x0, y0, node);
        size[gl - 1] = c;
        if (c == 0) continue;                    // lives in halo rows only: the neighbour rank counts it
        ++nclosed;
        unsigned long long pk = ((unsigned long long)(unsigned)c << 32) | (unsigned long long)(0xffffffffu - (unsigned)gl);
        if (pk > best) best = pk;
    }
    r.nloc = nloc;
    r.off = 0;
#if PERC_DEV
    if (nloc) r.off = (int)atomicAdd(&s.nroot, (unsigned)nloc);
    if (nclosed) { atomicAdd(&s.nclosed, nclosed); atomicMax(&s.best, best); }
#else
    r.off = (int)s.nroot; s.nroot += nloc;
    s.nclosed += nclosed;
    if (best > s.best) s.best = best;
#endif
}

// thread 0: reserve the block's segment of the global root list, publish the closed clusters
PERC_HD void tile_phase4_reserve(TileSmem& s, Summary* sum)
{
#if PERC_DEV
    if (s.nroot) s.rootbase = atomicAdd(&sum->nroots, s.nroot);
    if (s.lone) atomicAdd(&sum->nlone, (unsigned long long)s.lone);
    if (s.nclosed) { atomicAdd(&sum->ncl, (unsigned long long)s.nclosed); atomicMax(&sum->maxpack, s.best); }
#else
    s.rootbase = sum->nroots; sum->nroots += s.nroot;
    sum->nlone += s.lone;
    sum->ncl += s.nclosed;
    if (s.best > sum->maxpack) sum->maxpack = s.best;
#endif
}

// provisional labels: 32 threads per tile row, 4 consecutive sites each (128-bit stores)
// VAR = 0: s.lab holds the root of every active site (staged by phase 3); VAR = 1: only of the run starts -- a site's run
// start is the highest node bit at or below it (an active site that is no run start continues the run of its left
// neighbour, which the same thread has just resolved)
template <int VAR = 0>
PERC_HD void tile_phase4_labels(const TileSmem& s, const Geom& g, int x0, int y0, int tid, int32_t* __restrict__ label, bool vec)
{
    const int lane = tid & 31;
    const int w = lane >> 3, b0 = (lane & 7) << 2;
    for (int ly = tid >> 5; ly < CT_TH; ly += CT_THREADS / 32) {
        const int gy = y0 + ly, gx = x0 + 4 * lane;
        if (gy >= g.n || gx >= g.m) continue;
        const uint32_t S4 = (s.pS[ly + 1][w] >> b0) & 0xFu;
        const int* st = &s.lab[ly * CT_TW + 4 * lane];
        int32_t out[4];
        if (VAR == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) out[k] = (S4 >> k) & 1u ? tile_global_label(g, x0, y0, st[k]) : 0;
        } else {
            const uint32_t Tw = s.pT[ly][w];
            const int* wl = &s.lab[ly * CT_TW + (w << 5)];
            int node = -1;
            int32_t lab = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!((S4 >> k) & 1u)) { out[k] = 0; continue; }
                const int b = b0 + k;
                if (k == 0 || ((Tw >> b) & 1u) || !((S4 >> (k - 1)) & 1u)) {
                    const int nd = hibit(Tw & le_mask(b));
                    if (nd != node) { node = nd; lab = tile_global_label(g, x0, y0, wl[nd]); }
                }
                out[k] = lab;
            }
        }
        int32_t* dst = label + (int64_t)gy * g.m + gx;
        if (vec) {
#if PERC_DEV
            *reinterpret_cast<int4*>(dst) = make_int4(out[0], out[1], out[2], out[3]);
#else
            memcpy(dst, out, 16);
#endif
        } else {
            for (int k = 0; k < 4 && gx + k < g.m; ++k) dst[k] = out[k];
        }
    }
}

// border roots: size at the root's own site index, and the root list entry
PERC_HD void tile_phase4_roots(const TileSmem& s, const Geom& g, int x0, int y0, int tid, const TileRegs& r,
                               int32_t* __restrict__ size, int32_t* __restrict__ rootlist)
{
    if (!r.nloc) return;
    const int w = tid % CT_NW, ly = tid / CT_NW;
    const int base = ly * CT_TW + (w << 5);
    int k = 0;
    for (uint32_t t = r.rootbits; t; t &= t - 1) {
        const int node = base + lobit(t);
        if (!ring_get(s, node)) continue;
        const int c = cnt_get(s, node);
        const int32_t gl = tile_global_label(g, x0, y0, node);
        size[gl - 1] = c;
        rootlist[s.rootbase + r.off + k] = gl - 1;
        ++k;
    }
}

// ---- global table: label[i] = parent + 1 (0 = inactive site) ---------------------------------
PERC_HD int32_t ld_label(const int32_t* label, int64_t a)
{
#if PERC_DEV
    return __ldcg(&label[a]);
#else
    return label[a];
#endif
}

PERC_HD int gl_find(const int32_t* label, int a)
{
    int p;
    while ((p = ld_label(label, a) - 1) != a) a = p;
    return a;
}

// find with path halving (merge only: plain stores of an ancestor race benignly with atomicMin)
PERC_HD int gl_find_halve(int32_t* label, int a)
{
    for (;;) {
        int p = ld_label(label, a) - 1;
        if (p == a) return a;
        int gp = ld_label(label, p) - 1;
        if (gp == p) return p;
        label[a] = gp + 1;
        a = gp;
    }
}

PERC_HD void gl_unite(int32_t* label, int a, int b)
{
    for (;;) {
        a = gl_find_halve(label, a);
        b = gl_find_halve(label, b);
        if (a == b) return;
        if (a < b) { int tmp = a; a = b; b = tmp; }
#if PERC_DEV
        int old = atomicMin(&label[a], b + 1);
#else
        int old = label[a]; if (b + 1 < old) label[a] = b + 1;
#endif
        if (old == a + 1) return;
        a = old - 1;
    }
}

// merge work: the first nrowb * nwords * 32 ids = (word of the last row of a tile row, lane): the lane
// makes the union of its column (bonds N / NW / NE into the next tile row) if the column is not
// redundant; then ncolb * n ids = one row of a vertical tile border (or of the periodic wrap column).
template <int LAT>
PERC_HD void merge_item(const Geom& g, const uint8_t* __restrict__ mask, int32_t* label, int nrowb, int nwords,
                        int ncolb, int64_t id, bool vec)
{
    const int64_t na = (int64_t)nrowb * nwords * 32;
    if (id < na) {
        const int x = (int)(id & 31);
        const int64_t wid = id >> 5;
        const int k = (int)(wid / nwords), wi = (int)(wid % nwords);
        const int y = (k + 1) * CT_TH - 1, gx0 = wi << 5;
        uint32_t S, E, N, NW, NE, Su, Eu, Nu, NWu, NEu;
        load_planes<LAT>(g, mask, gx0, y, vec, S, E, N, NW, NE);
        if (!((S >> x) & 1u)) return;
        load_planes<LAT>(g, mask, gx0, y + 1, vec, Su, Eu, Nu, NWu, NEu);
        const uint32_t connL = (S & E & (S >> 1)) << 1, Cu = (Su & Eu & (Su >> 1)) << 1;
        const uint32_t v = S & N & Su;
        const int64_t i = (int64_t)y * g.m + gx0 + x;
        if (((v & ~((v << 1) & connL & Cu)) >> x) & 1u) gl_unite(label, (int)i, (int)(i + g.m));
        if (LAT == LAT_TRIANGULAR) {
            if ((((S & NW) & ~(v & Cu)) >> x) & 1u) {
                int64_t j = bond_other_end(g, gx0 + x, y, DIR_NW);
                if (mask[j] & MASK_SITE) gl_unite(label, (int)i, (int)j);
            }
            if ((((S & NE) & ~(v & (Cu >> 1))) >> x) & 1u) {
                int64_t j = bond_other_end(g, gx0 + x, y, DIR_NE);
                if (mask[j] & MASK_SITE) gl_unite(label, (int)i, (int)j);
            }
        }
        return;
    }
    id -= na;
    if (id >= (int64_t)ncolb * g.n) return;
    const int kb = (int)(id / g.n), y = (int)(id % g.n);
    // border kb: between column xl = (kb+1)*TW - 1 and xr = xl + 1; the last one may be the wrap
    int xl = (kb + 1) * CT_TW - 1, xr = xl + 1;
    if (xl >= g.m - 1) { xl = g.m - 1; xr = 0; if (!g.pbc) return; }
    const int64_t il = (int64_t)y * g.m + xl, ir = (int64_t)y * g.m + xr;
    const unsigned ml = mask[il], mr = mask[ir];
    if ((ml & MASK_SITE) && (ml & MASK_E) && (mr & MASK_SITE)) gl_unite(label, (int)il, (int)ir);
    if (LAT == LAT_TRIANGULAR && y + 1 < g.n) {
        // NW bond of the right column (x even) reaches the left column one row up
        if (!(xr & 1) && (mr & MASK_SITE) && (mr & MASK_NW)) {
            int64_t j = (int64_t)(y + 1) * g.m + xl;
            if (mask[j] & MASK_SITE) gl_unite(label, (int)ir, (int)j);
        }
        // NE bond of the left column (x even; only when a tile border falls on an odd column count)
        if (!(xl & 1) && (ml & MASK_SITE) && (ml & MASK_NE) && xl + 1 < g.m) {
            int64_t j = (int64_t)(y + 1) * g.m + xl + 1;
            if (mask[j] & MASK_SITE) gl_unite(label, (int)il, (int)j);
        }
    }
}

// root list entry k: tile-local root j -> global root r; fold the tile-local size into size[r].
// Lanes of a warp that reach the same global root add once (the largest cluster owns a root in almost
// every tile: without this its size entry takes one serialised atomic per tile-local root).
// Returns (size << 32) | ~label candidates for the largest cluster; *isroot = 1 for a global root.
PERC_HD unsigned long long rootfix_item(int32_t* label, int32_t* size, const int32_t* __restrict__ rootlist,
                                        int64_t k, bool valid, int* isroot)
{
    int j = -1, r = -1, s = 0;
    *isroot = 0;
    if (valid) {
        j = rootlist[k];
        r = gl_find(label, j);
        if (r != j) { label[j] = r + 1; s = size[j]; }
        else *isroot = 1;
    }
    long long cand = 0;
#if PERC_DEV
    const unsigned peers = __match_any_sync(0xffffffffu, r);
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    int tot = s;
    for (unsigned m = peers & (peers - 1); m; m &= m - 1) {          // peers other than the leader
        int v = __shfl_sync(peers, s, __ffs(m) - 1);
        if (lane == leader) tot += v;
    }
    if (valid && lane == leader) {
        if (tot) cand = (long long)atomicAdd(&size[r], tot) + tot;
        else cand = __ldcg(&size[r]);
    }
#else
    if (valid) { size[r] += s; cand = size[r]; }
#endif
    if (!valid || cand == 0) return 0;
    return ((unsigned long long)cand << 32) | (unsigned long long)(0xffffffffu - (unsigned)(r + 1));
}

// one hop to the global root (valid after rootfix: every tile-local root points at its global root)
PERC_HD int32_t flatten_one(const int32_t* label, int32_t l)
{
    return l ? ld_label(label, l - 1) : 0;
}

}  // namespace perc
