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
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../percolation_b200/csrc/pcg_fused_tile.cuh"
#include "../percolation_b200/csrc/pcg_defl_host.h"

using namespace perc;

// the kernel picks the instantiation of the tile phases per tile (interior / boundary): same dispatch here
#define EMUL_MAIN(...) do { if (interior) ft_phase_main<LAT, C, true>(__VA_ARGS__); else ft_phase_main<LAT, C, false>(__VA_ARGS__); } while (0)
#define EMUL_RING(...) do { if (interior) ft_phase_ringcols<LAT, C, true>(__VA_ARGS__); else ft_phase_ringcols<LAT, C, false>(__VA_ARGS__); } while (0)
#define EMUL_ENERGY(...) do { if (interior) ft_phase_energy<LAT, C, true>(__VA_ARGS__); else ft_phase_energy<LAT, C, false>(__VA_ARGS__); } while (0)

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
    return diag_seq(g, c, ex, x, g0, gleak);
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
        const double b = rhs_seq(g, c, ex, x, g0, gleak, Va);
        const double z = b / diag_of(g, cf, x, y, g0, gleak);
        r[0][(size_t)y * m + x] = C::USTATE ? z : b;          // V = 3: the state vector is u = D^-1 r
        bn += z * z;
    }
    FtState st{};
    st.bnrm = sqrt(bn); st.tol = tol; st.itmax = itmax;

    std::vector<FtDiag> dtab(64 * C::DC);
    for (int k = 0; k < ft_tab_slots<LAT, C>(); ++k) dtab[k] = ft_tab_slot<LAT, C>(g, k, g0, gleak);
    std::vector<double> cinv(64);
    for (int k = 0; k < 64; ++k) cinv[k] = ft_cinv_entry(k, g0, gleak);
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
            if (C::USTATE && g.pbc && (x0 == 0 || x0 + C::TX == m))   // periodic wrap: the ring threads patch the halo columns
                for (int rl = 0; rl < C::RING_NT; ++rl)
                    ft_wrap_patch<C>(g, x0, y0, r[cur].data(), s[cur].data(), cf.data(), sr.data(), ss.data(), scf.data(), rl, C::RING_NT);
            const bool interior = ft_interior<C>(g, x0, y0);
            if (prime && interior) ++*tiles_fast;
            if (!C::USTATE)
                for (int tid = 0; tid < C::THREADS; ++tid) ft_phase_u<LAT, C>(g, sr.data(), scf.data(), su.data(), dtab.data(), x0, y0, interior, tid, g0, gleak, cinv.data());
            const double* up = C::USTATE ? sr.data() : su.data();
            for (int tid = C::THREADS - 1; tid >= 0; --tid) {
                EMUL_MAIN(g, sc, sr.data(), ss.data(), scf.data(), up, dtab.data(), cinv.data(), x0, y0, interior, tid,
                                      r[cur ^ 1].data(), s[cur ^ 1].data(), xrow.data(), prow.data(), rz, rr);
                EMUL_RING(g, sc, sr.data(), ss.data(), scf.data(), up, dtab.data(), cinv.data(), x0, y0, tid, nullptr);
            }
            for (int tid = 0; tid < C::THREADS; ++tid) EMUL_ENERGY(g, sc, ss.data(), scf.data(), x0, y0, tid, en);
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


// the deflated one-pass solve (FtCfgD), phase by phase as pcg_fused_kernel<.., FtCfgD> runs it: shift table per tile,
// main / ring / energy phases, the crossing currents by the ring threads, then the coarse stage (assemble Z^T A u',
// mu = E^-1 ., the scalar recurrences with delta - mu . Z^T A u')
template <int LAT, class C>
int solve_defl(const Geom& g, const std::vector<uint8_t>& cf, double Va, double g0, double gleak, double tol, int itmax,
               double read_thresh, double* Gtop, double* Gbot, int* iter, double* err, int* coarse_dim, int bw0, int bh0, double* xfull = nullptr)
{
    const int m = g.m, n = g.n;
    const int64_t t = g.t;
    const FtDefl D = ft_defl_make(g, C::TX, C::TY, FT_KMAX, bw0, bh0);
    *coarse_dim = D.k;
    const int ntiles = D.ntx * D.nty;
    // E from the crossing bond weights, dense inverse
    std::vector<double> W((size_t)ntiles * FS_STRIDE, 0.0), Einv((size_t)D.k * D.k), Fb((size_t)FB_PLANES * FT_KMAX, 0.0);
    const FtGlobalAcc ga{cf.data(), m};
    for (int tl = 0; tl < ntiles; ++tl)
        for (int q = 0; q < FtFluxItems<C>::N; ++q)
            ft_flux_item<LAT, C, true>(g, g0, gleak, ga, (tl % D.ntx) * C::TX, (tl / D.ntx) * C::TY, q, &W[(size_t)tl * FS_STRIDE]);
    for (int y = 1; y <= n - 2; ++y)
        for (int x = 0; x < m; ++x) {
            double rho;
            diag_seq_rho(g, cf[(size_t)y * m + x], neighbour_bits(g, x, y), x, g0, gleak, &rho);
            W[(size_t)((y / C::TY) * D.ntx + x / C::TX) * FS_STRIDE + FS_R] += rho;
        }
    if (ft_defl_build_einv(D, W.data(), Einv.data(), 2)) return -4;
    // x0 = Z nu, E nu = Z^T b; u0 = D^-1 (b - A Z nu)
    std::vector<double> r[2], s[2], xrow((size_t)2 * m, 0.0), prow((size_t)2 * m, 0.0), fb((size_t)D.k, 0.0), nu((size_t)D.k, 0.0), mu((size_t)D.k, 0.0);
    for (int k = 0; k < 2; ++k) { r[k].assign((size_t)t, 0.0); s[k].assign((size_t)t, 0.0); }
    double bn = 0.0;
    {
        const std::vector<double> zero((size_t)D.k, 0.0);
        for (int x = 0; x < m; ++x) {
            double b, ni;
            const double z = ft_defl_u0<C>(g, D, cf[(size_t)(n - 2) * m + x], zero.data(), x, n - 2, Va, g0, gleak, &b, &ni);
            fb[ft_defl_block(D, x / C::TX, (n - 2) / C::TY)] += b;
            bn += z * z;
        }
    }
    for (int i = 0; i < D.k; ++i) { double a = 0.0; for (int j = 0; j < D.k; ++j) a += Einv[(size_t)i * D.k + j] * fb[j]; nu[i] = a; }
    for (int y = 1; y <= n - 2; ++y)
        for (int x = 0; x < m; ++x) {
            double b, ni;
            r[0][(size_t)y * m + x] = ft_defl_u0<C>(g, D, cf[(size_t)y * m + x], nu.data(), x, y, Va, g0, gleak, &b, &ni);
            if (y == 1) xrow[x] = ni;
            if (y == n - 2) xrow[(size_t)m + x] = ni;
        }
    FtState st{};
    st.bnrm = sqrt(bn); st.tol = tol; st.itmax = itmax;
    // (accuracy study: the full iterate x = Z nu + sum alpha p -- the kernel keeps it on two rows only)
    std::vector<double> pfull;
    if (xfull) {
        pfull.assign((size_t)t, 0.0);
        for (size_t q = 0; q < (size_t)t; ++q) xfull[q] = 0.0;
        for (int y = 1; y <= n - 2; ++y) for (int x = 0; x < m; ++x) xfull[(size_t)y * m + x] = nu[ft_defl_block(D, x / C::TX, y / C::TY)];
    }
    std::vector<FtDiag> dtab(64 * C::DC);
    for (int k = 0; k < ft_tab_slots<LAT, C>(); ++k) dtab[k] = ft_tab_slot<LAT, C>(g, k, g0, gleak);
    std::vector<double> cinv(64);
    for (int k = 0; k < 64; ++k) cinv[k] = ft_cinv_entry(k, g0, gleak);
    std::vector<double> sr((size_t)C::RR * C::LD), ss((size_t)C::SR * C::LD), sft((size_t)C::SFT_N), srec((size_t)C::REC_N), sru((size_t)C::THREADS), rtab((size_t)ft_tab_slots<LAT, C>());
    for (int k = 0; k < ft_tab_slots<LAT, C>(); ++k) rtab[k] = ft_rho_slot<LAT, C>(g, k, g0, gleak);
    std::vector<uint8_t> scf((size_t)C::RR * C::CLD);
    int cur = 0;
    double esum = 0.0;
    std::vector<int> sched;
    ft_defl_schedule<C>(g, D, 5, 2.0, sched);
    for (int pass = 0; !st.done; ++pass) {
        const int prime = pass == 0;
        const FtScalars sc{g0, gleak, prime ? 0.0 : st.alpha, prime ? 0.0 : st.beta};
        if (xfull && !prime)
            for (int y = 1; y <= n - 2; ++y) for (int x = 0; x < m; ++x) {
                const size_t q = (size_t)y * m + x;
                pfull[q] = (r[cur][q] - mu[ft_defl_block(D, x / C::TX, y / C::TY)]) + sc.beta * pfull[q];
                xfull[q] += sc.alpha * pfull[q];
            }
        double rz = 0.0, rr = 0.0, en = 0.0;
        // the kernel's walk (FtWalk) as a grid of GRID CTAs runs it, forwards and backwards in turn: block by block, the slot
        // totals of a tile folded into the running sums of its block, which go out with the block's last tile
        // (every third pass with the host-built schedule of the library -- ft_defl_schedule: boundary tiles spread over the CTAs)
        const int GRID = 5, rev = pass & 1;
        const int* schp = (pass % 3 == 2) ? sched.data() : nullptr;
        int visited = 0;
        for (int bid = 0; bid < GRID; ++bid) {
            FtWalk wk;
            double bacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (wk.start(D, bid, GRID, rev, schp); wk.valid(D); wk.next(D, GRID, rev, schp)) {
                ++visited;
                const int ix = wk.ix(D), iy = wk.iy(D), x0 = ix * C::TX, y0 = iy * C::TY, info = wk.info(D, rev);
                if (ft_defl_block(D, ix, iy) != wk.B) return -6;
                box_copy(sr.data(), r[cur].data(), m, n, x0 - 2, y0 - 1, C::LD, C::RR);
                box_copy(ss.data(), s[cur].data(), m, n, x0 - 2, y0, C::LD, C::SR);
                box_copy(scf.data(), cf.data(), m, n, x0 - 16, y0 - 1, C::CLD, C::RR);
                if (g.pbc && (x0 == 0 || x0 + C::TX == m))           // periodic wrap: the ring threads patch the halo columns
                    for (int rl = 0; rl < C::RING_NT; ++rl)
                        ft_wrap_patch<C>(g, x0, y0, r[cur].data(), s[cur].data(), cf.data(), sr.data(), ss.data(), scf.data(), rl, C::RING_NT);
                for (int j = 0; j < C::SFT_N; ++j) sft[j] = (j & 3) == 3 ? 0.0 : ft_defl_shift_entry<C>(g, D, mu.data(), ix, iy, j >> 2, j & 3);
                for (int j = 0; j < C::REC_N; ++j) srec[j] = ft_defl_rec_entry<C>(sft.data(), j >> 3, j & 7);
                const bool interior = ft_interior<C>(g, x0, y0);
                for (int tid = C::THREADS - 1; tid >= 0; --tid) {
                    double ru = 0.0;
                    EMUL_MAIN(g, sc, sr.data(), ss.data(), scf.data(), sr.data(), dtab.data(), cinv.data(), x0, y0, interior, tid,
                                                 r[cur ^ 1].data(), s[cur ^ 1].data(), xrow.data(), prow.data(), rz, rr, srec.data(), rtab.data(), &ru);
                    EMUL_RING(g, sc, sr.data(), ss.data(), scf.data(), sr.data(), dtab.data(), cinv.data(), x0, y0, tid, sft.data(), rtab.data());
                    sru[tid] = ru;
                }
                for (int tid = 0; tid < C::THREADS; ++tid) EMUL_ENERGY(g, sc, ss.data(), scf.data(), x0, y0, tid, en);
                // crossing currents by the 64 ring threads: per-thread partial sums, folded lane by lane (the kernel: a shuffle tree)
                double tsl[FS_STRIDE] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (int rl = 0; rl < C::RING_NT; ++rl) {
                    double fl[FS_SLOTS] = {0, 0, 0, 0, 0, 0};
                    ft_flux_thread<LAT, C>(g, sc, ss.data(), scf.data(), sru.data(), x0, y0, rl, interior, true, fl);
                    for (int k = 0; k < FS_SLOTS; ++k) tsl[k] += fl[k];
                }
                ft_defl_block_add(info, tsl, bacc);
                if (info & FW_LAST) for (int q = 0; q < FB_PLANES; ++q) { Fb[(size_t)q * FT_KMAX + wk.B] = bacc[q]; bacc[q] = 0.0; }
            }
        }
        if (visited != ntiles) return -7;                     // every tile exactly once (the sums below would not notice a repeat)
        // coarse stage
        std::vector<double> f((size_t)D.k);
        for (int B = 0; B < D.k; ++B) f[B] = ft_defl_block_f<LAT>(D, Fb.data(), B);
        double mf = 0.0;
        for (int i = 0; i < D.k; ++i) { double a = 0.0; for (int j = 0; j < D.k; ++j) a += Einv[(size_t)i * D.k + j] * f[j]; mu[i] = a; mf += a * f[i]; }
        if (!prime) esum += st.alpha * st.gamma;            // the step just applied: |e_k|_A^2 - |e_k+1|_A^2 = alpha_k gamma_k
        ft_scalar_step(st, rz, rr, en - mf, prime);
        cur ^= 1;
        if (pass > 50 * 1000 * 1000) return -1;
    }
    if (getenv("EMUL_FINALC")) {
        // final coarse correction: x += Z E^-1 Z^T r with the recursive residual r = d u
        std::vector<double> fr((size_t)D.k, 0.0), mr((size_t)D.k, 0.0);
        for (int y = 1; y <= n - 2; ++y) for (int x = 0; x < m; ++x)
            fr[ft_defl_block(D, x / C::TX, y / C::TY)] += diag_of(g, cf, x, y, g0, gleak) * r[cur][(size_t)y * m + x];
        double nf = 0.0;
        for (int i = 0; i < D.k; ++i) { double a2 = 0.0; for (int j = 0; j < D.k; ++j) a2 += Einv[(size_t)i * D.k + j] * fr[j]; mr[i] = a2; nf += fr[i] * fr[i]; }
        printf("FINALC |Z^T r| = %.3e  max |mu_r| = %.3e\n", sqrt(nf), *std::max_element(mr.begin(), mr.end()));
        for (int x = 0; x < m; ++x) { xrow[x] += mr[ft_defl_block(D, x / C::TX, 1 / C::TY)]; xrow[(size_t)m + x] += mr[ft_defl_block(D, x / C::TX, (n - 2) / C::TY)]; }
        if (xfull) for (int y = 1; y <= n - 2; ++y) for (int x = 0; x < m; ++x) xfull[(size_t)y * m + x] += mr[ft_defl_block(D, x / C::TX, y / C::TY)];
    }
    if (getenv("EMUL_ENERGY")) {
        double nufb = 0.0;
        for (int i = 0; i < D.k; ++i) nufb += nu[i] * fb[i];
        printf("ENERGY nu.fb=%.17g esum=%.17g bx=%.17g\n", nufb, esum, nufb + esum);
    }
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
                                double* err, int* tiles_fast, int cfg, int pbc)
{
    if (m % 16 || n < 4) return -2;
    if (pbc && (cfg != 2 || m % FtCfgA3::TX)) return -2;     // (what pcg_fused_applies admits)
    const Geom g = make_geom(lattice, m, n, pbc);
    std::vector<uint8_t> cf;
    build_cfull(g, w, gleak, cf);
#define RUN(CFG) (lattice == LAT_SQUARE ? solve<LAT_SQUARE, CFG>(g, cf, Va, g0, gleak, tol, itmax, read_thresh, Gtop, Gbot, iter, err, tiles_fast) \
                                        : solve<LAT_TRIANGULAR, CFG>(g, cf, Va, g0, gleak, tol, itmax, read_thresh, Gtop, Gbot, iter, err, tiles_fast))
    switch (cfg) {
    case 0: return RUN(FtCfgA);
    case 2: return RUN(FtCfgA3);
    }
#undef RUN
    return -3;
}

// the deflated solver (FtCfgD); bw, bh: tiles per block (0 = the library's choice); coarse_dim: number of blocks
extern "C" int fused_emul_solve_defl(int lattice, int m, int n, const double* w, double Va, double g0, double gleak,
                                     double tol, int itmax, double read_thresh, double* Gtop, double* Gbot, int* iter,
                                     double* err, int* coarse_dim, int bw, int bh, double* xfull, int pbc)
{
    if (m % 16 || n < 4) return -2;
    if (pbc && m % FtCfgD::TX) return -2;                   // (what pcg_fused_applies admits)
    const Geom g = make_geom(lattice, m, n, pbc);
    std::vector<uint8_t> cf;
    build_cfull(g, w, gleak, cf);
    return lattice == LAT_SQUARE ? solve_defl<LAT_SQUARE, FtCfgD>(g, cf, Va, g0, gleak, tol, itmax, read_thresh, Gtop, Gbot, iter, err, coarse_dim, bw, bh, xfull)
                                 : solve_defl<LAT_TRIANGULAR, FtCfgD>(g, cf, Va, g0, gleak, tol, itmax, read_thresh, Gtop, Gbot, iter, err, coarse_dim, bw, bh, xfull);
}
