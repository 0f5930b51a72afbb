// pcg_fused_emul.cpp -- host emulation of the one-pass PCG kernel (TEST INFRASTRUCTURE).
//
// Compiles percolation_b200/csrc/pcg_fused_tile.cuh -- the very source pcg_fused_kernel is built from -- with
// g++ and executes the phases of every tile thread by thread (a __syncthreads() becomes the end of a loop
// over tid; a TMA box copy becomes a loop with zero fill outside the lattice).  Lets the CPU test-suite check
// the recurrences, the halo / ring indexing and the bond-energy sum against the oracle without a GPU.
// Never used by the product path.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../percolation_b200/csrc/pcg_fused_tile.cuh"

using namespace perc;

namespace {

// conduct bytes (8 direction bits per site) from per-bond weights in reference row order
void build_cfull(const Geom& g, const double* w, double gleak, std::vector<uint8_t>& cf)
{
    static const unsigned fwd[4] = {NB_E, NB_N, NB_NW, NB_NE}, back[4] = {NB_W, NB_S, NB_SE, NB_SW};
    cf.assign((size_t)g.t, 0);
    for (int64_t r = 0; r < g.nb; ++r) {
        if (!(w[r] > gleak)) continue;
        int64_t a; int dir;
        ref_row_to_owner(g, r, &a, &dir);
        const int64_t b = bond_other_end(g, (int)(a % g.m), (int)(a / g.m), dir);
        cf[a] |= (uint8_t)fwd[dir];
        cf[b] |= (uint8_t)back[dir];
    }
}

double diag_of(const Geom& g, const std::vector<uint8_t>& cf, int x, int y, double g0, double gleak)
{
    const unsigned ex = neighbour_bits(g, x, y), c = cf[(size_t)y * g.m + x] & ex;
    const int nc = __builtin_popcount(c), ne = __builtin_popcount(ex);
    return fma((double)(ne - nc), gleak, (double)nc * g0);
}

// "TMA": box of `rows` x `cols` elements at element coordinates (cx, cy) of a row-major m x n array
template <typename T>
void box_copy(T* dst, const T* src, int m, int n, int cx, int cy, int cols, int rows)
{
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            const int x = cx + c, y = cy + r;
            dst[(size_t)r * cols + c] = (x >= 0 && x < m && y >= 0 && y < n) ? src[(size_t)y * m + x] : T(0);
        }
}

template <int LAT, class C>
int solve(const Geom& g, const std::vector<uint8_t>& cf, double Va, double g0, double gleak, double tol, int itmax,
          double read_thresh, double* Gtop, double* Gbot, int* iter, double* err, int* tiles_fast)
{
    const int m = g.m, n = g.n;
    const int64_t t = g.t;
    std::vector<double> r[2], s[2], xrow((size_t)2 * m, 0.0), prow((size_t)2 * m, 0.0);
    for (int k = 0; k < 2; ++k) { r[k].assign((size_t)t, 0.0); s[k].assign((size_t)t, 0.0); }
    // r = b (Sq/bondc.f:490-497), bnrm = |D^-1 b|
    double bn = 0.0;
    for (int x = 0; x < m; ++x) {
        const int y = n - 2;
        const unsigned ex = neighbour_bits(g, x, y), c = cf[(size_t)y * m + x];
        double b = 0.0;
        if (ex & NB_N)  b += ((c & NB_N)  ? g0 : gleak) * Va;
        if (ex & NB_NW) b += ((c & NB_NW) ? g0 : gleak) * Va;
        if (ex & NB_NE) b += ((c & NB_NE) ? g0 : gleak) * Va;
        const double z = b / diag_of(g, cf, x, y, g0, gleak);
        r[0][(size_t)y * m + x] = C::USTATE ? z : b;          // V = 3: the state vector is u = D^-1 r
        bn += z * z;
    }
    FtState st{};
    st.bnrm = sqrt(bn); st.tol = tol; st.itmax = itmax;

    std::vector<FtDiag> dtab(64 * C::DC);
    for (int k = 0; k < 64 * C::DC; ++k) dtab[k] = ft_diag_entry(k / C::DC, g0, gleak);
    const int ntx = (m + C::TX - 1) / C::TX, nty = (n + C::TY - 1) / C::TY;
    std::vector<double> sr((size_t)C::RR * C::LD), ss((size_t)C::SR * C::LD), su((size_t)C::RR * C::LD);
    std::vector<uint8_t> scf((size_t)C::RR * C::CLD);
    *tiles_fast = 0;
    int cur = 0;
    for (int pass = 0; !st.done; ++pass) {
        const int prime = pass == 0;
        const FtScalars sc{g0, gleak, prime ? 0.0 : st.alpha, prime ? 0.0 : st.beta};
        double rz = 0.0, rr = 0.0, en = 0.0;
        for (int tl = 0; tl < ntx * nty; ++tl) {
            const int x0 = (tl % ntx) * C::TX, y0 = (tl / ntx) * C::TY;
            box_copy(sr.data(), r[cur].data(), m, n, x0 - 2, y0 - 1, C::LD, C::RR);
            box_copy(ss.data(), s[cur].data(), m, n, x0 - 2, y0, C::LD, C::SR);
            box_copy(scf.data(), cf.data(), m, n, x0 - 16, y0 - 1, C::CLD, C::RR);
            for (auto& v : su) v = NAN;                        // shared memory starts as garbage
            const bool interior = ft_interior<C>(g, x0, y0);
            if (prime && interior) ++*tiles_fast;
            if (!C::USTATE)
                for (int tid = 0; tid < C::THREADS; ++tid) ft_phase_u<LAT, C>(g, sr.data(), scf.data(), su.data(), dtab.data(), x0, y0, interior, tid);
            const double* up = C::USTATE ? sr.data() : su.data();
            for (int tid = C::THREADS - 1; tid >= 0; --tid) {
                if (C::SPLIT && interior)
                    ft_phase_main<LAT, C, true>(g, sc, sr.data(), ss.data(), scf.data(), up, dtab.data(), x0, y0, true, tid,
                                                r[cur ^ 1].data(), s[cur ^ 1].data(), xrow.data(), prow.data(), rz, rr);
                else
                    ft_phase_main<LAT, C, false>(g, sc, sr.data(), ss.data(), scf.data(), up, dtab.data(), x0, y0, interior, tid,
                                                 r[cur ^ 1].data(), s[cur ^ 1].data(), xrow.data(), prow.data(), rz, rr);
                ft_phase_ringcols<LAT, C>(g, sc, sr.data(), ss.data(), scf.data(), up, dtab.data(), x0, y0, tid);
            }
            for (int tid = 0; tid < C::THREADS; ++tid) {
                if (C::SPLIT && interior) ft_phase_energy<LAT, C, true>(g, sc, ss.data(), scf.data(), x0, y0, true, tid, en);
                else ft_phase_energy<LAT, C, false>(g, sc, ss.data(), scf.data(), x0, y0, interior, tid, en);
            }
        }
        ft_scalar_step(st, rz, rr, en, prime);
        cur ^= 1;
        if (pass > 50 * 1000 * 1000) return -1;
    }
    // read-out (Sq/bondc.f:554-592): rows 0 and n-1 of G~ V, off-diagonals below read_thresh dropped
    double top = 0.0, bot = 0.0;
    for (int e = 0; e < 2; ++e)
        for (int x = 0; x < m; ++x) {
            const int y = e == 0 ? 0 : n - 1;
            const unsigned ex = neighbour_bits(g, x, y), c = cf[(size_t)y * m + x];
            const double vi = e == 0 ? 0.0 : Va;
            double acc = diag_of(g, cf, x, y, g0, gleak) * vi;
            const int xl = x > 0 ? x - 1 : m - 1, xr = x + 1 < m ? x + 1 : 0;
            auto val = [&](int xx, int yy) { return yy == 0 ? 0.0 : yy == n - 1 ? Va : yy == 1 ? xrow[xx] : xrow[(size_t)m + xx]; };
#define NBR(bit, xx, yy) if (ex & bit) { const double wt = (c & bit) ? g0 : gleak; if (fabs(wt) >= read_thresh) acc -= wt * val(xx, yy); }
            NBR(NB_E, xr, y) NBR(NB_W, xl, y) NBR(NB_N, x, y + 1) NBR(NB_S, x, y - 1)
            NBR(NB_NW, xl, y + 1) NBR(NB_NE, xr, y + 1) NBR(NB_SW, xl, y - 1) NBR(NB_SE, xr, y - 1)
#undef NBR
            if (e == 0) bot += acc; else top += acc;
        }
    *Gtop = top / Va; *Gbot = fabs(bot) / Va; *iter = st.iter; *err = st.err;
    return 0;
}

}  // namespace

// w: per-bond weights in reference row order (g0 for conducting bonds, gleak otherwise); cfg: tile configuration
// (the ones pcg.cu instantiates)
extern "C" int fused_emul_solve(int lattice, int m, int n, const double* w, double Va, double g0, double gleak,
                                double tol, int itmax, double read_thresh, double* Gtop, double* Gbot, int* iter,
                                double* err, int* tiles_fast, int cfg)
{
    if (m % 16 || n < 4) return -2;
    const Geom g = make_geom(lattice, m, n, 0);
    std::vector<uint8_t> cf;
    build_cfull(g, w, gleak, cf);
#define RUN(CFG) (lattice == LAT_SQUARE ? solve<LAT_SQUARE, CFG>(g, cf, Va, g0, gleak, tol, itmax, read_thresh, Gtop, Gbot, iter, err, tiles_fast) \
                                        : solve<LAT_TRIANGULAR, CFG>(g, cf, Va, g0, gleak, tol, itmax, read_thresh, Gtop, Gbot, iter, err, tiles_fast))
    switch (cfg) {
    case 0: return RUN(FtCfgA);
    case 2: return RUN(FtCfgA3);
    case 5: return RUN(FtCfgA4);
    }
#undef RUN
    return -3;
}
