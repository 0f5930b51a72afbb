// pcg_defl_host.h -- host side of the deflated one-pass PCG (pcg_fused_tile.cuh, FtDefl): the coarse operator
// E = Z^T A Z from the per-tile sums of the bond weights that cross tile borders, its banded Cholesky factor and its
// dense inverse.  E is the weighted graph Laplacian of the blocks (5-point on the square lattice, 7-point on the
// triangular one; block (bx, by) couples to (bx +- 1, by), (bx, by +- 1) and, triangular, (bx - 1, by + 1) / (bx + 1, by - 1))
// plus the weights of the bonds into the two Dirichlet rows on its diagonal: symmetric positive definite, half
// bandwidth nbx + 1 in the ordering B = by * nbx + bx.  Plain C++ (no CUDA): pcg.cu calls it between the weight kernel
// and the first sweep; tests/pcg_fused_emul.cpp calls the same code on the CPU.
#pragma once
#include <algorithm>
#include <cmath>
#include <thread>
#include <vector>
#include "pcg_fused_tile.cuh"

namespace perc {

// W[tile * FS_STRIDE + slot]: sum of the weights of the tile's crossing bonds per slot (ft_flux_item<.., UNIT = true>).
// Einv: k * k doubles, row-major, symmetric.  Blocks without any unknown get the identity row (their mu stays 0).
// Returns 0, or -1 if E is not positive definite (cannot happen for a lattice with n >= 3).
inline int ft_defl_build_einv(const FtDefl& D, const double* W, double* Einv, int nthreads = 8)
{
    // half bandwidth: (bx - 1, by + 1) is nbx - 1 away; with the periodic wrap (bx = 0) <-> (nbx - 1, by + 1) is 2 nbx - 1 away
    const int k = D.k, hb = D.pbc ? 2 * D.nbx : D.nbx + 1;
    // lower band: L[i * (hb + 1) + (hb - (i - j))] = E[i][j], j = i - hb .. i
    const int ld = hb + 1;
    std::vector<double> L((size_t)k * ld, 0.0);
    auto at = [&](int i, int j) -> double& { return L[(size_t)i * ld + (hb - (i - j))]; };
    const int dx[4] = {1, 0, -1, -1}, dy[4] = {0, 1, 0, 1};
    for (int iy = 0; iy < D.nty; ++iy)
        for (int ix = 0; ix < D.ntx; ++ix) {
            const int tl = iy * D.ntx + ix, B = ft_defl_block(D, ix, iy);
            at(B, B) += W[(size_t)tl * FS_STRIDE + FS_D] + W[(size_t)tl * FS_STRIDE + FS_R];
            for (int s = 0; s < 4; ++s) {
                int jx = ix + dx[s];
                const int jy = iy + dy[s];
                if (D.pbc) jx = jx < 0 ? D.ntx - 1 : (jx >= D.ntx ? 0 : jx);
                if (jx < 0 || jx >= D.ntx || jy >= D.nty) continue;
                const int B2 = ft_defl_block(D, jx, jy);
                if (B2 == B) continue;
                const double w = W[(size_t)tl * FS_STRIDE + s];
                at(B, B) += w; at(B2, B2) += w;
                if (B2 > B) at(B2, B) -= w; else at(B, B2) -= w;
            }
        }
    for (int i = 0; i < k; ++i) if (!(at(i, i) > 0.0)) at(i, i) = 1.0;
    // banded Cholesky E = L L^T, in place
    for (int j = 0; j < k; ++j) {
        double d = at(j, j);
        const int lo = j - hb > 0 ? j - hb : 0;
        for (int p = lo; p < j; ++p) d -= at(j, p) * at(j, p);
        if (!(d > 0.0)) return -1;
        d = std::sqrt(d);
        at(j, j) = d;
        const int hi = j + hb < k - 1 ? j + hb : k - 1;
        for (int i = j + 1; i <= hi; ++i) {
            double v = at(i, j);
            const int lo2 = i - hb > lo ? i - hb : lo;
            for (int p = lo2; p < j; ++p) v -= at(i, p) * at(j, p);
            at(i, j) = v / d;
        }
    }
    // inverse, column by column: L y = e_c, L^T x = y (columns are independent: spread over host threads)
    auto column_range = [&](int c0, int c1) {
        std::vector<double> y((size_t)k);
        for (int c = c0; c < c1; ++c) {
            for (int i = 0; i < c; ++i) y[i] = 0.0;
            for (int i = c; i < k; ++i) {
                double v = i == c ? 1.0 : 0.0;
                const int lo = i - hb > c ? i - hb : c;
                for (int p = lo; p < i; ++p) v -= at(i, p) * y[p];
                y[i] = v / at(i, i);
            }
            for (int i = k - 1; i >= 0; --i) {
                double v = y[i];
                const int hi = i + hb < k - 1 ? i + hb : k - 1;
                for (int p = i + 1; p <= hi; ++p) v -= at(p, i) * y[p];
                y[i] = v / at(i, i);
            }
            for (int i = 0; i < k; ++i) Einv[(size_t)i * k + c] = y[i];
        }
    };
    if (nthreads > k / 64) nthreads = k / 64;
    if (nthreads <= 1) column_range(0, k);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back(column_range, (int)((int64_t)k * t / nthreads), (int)((int64_t)k * (t + 1) / nthreads));
        for (auto& t : th) t.join();
    }
    return 0;
}

// Which CTA walks which blocks: tiles that touch the lattice border run the general code of the tile phases (about twice the
// time of an interior tile), and a block on the left / right edge consists of such tiles only.  Longest-processing-time-first
// over the blocks' costs, then every CTA's list in ascending block order.  out = offsets [G + 1] followed by the lists [k].
template <class C>
inline void ft_defl_schedule(const Geom& g, const FtDefl& D, int G, double boundary_cost, std::vector<int>& out)
{
    std::vector<double> cost((size_t)D.k, 0.0);
    for (int iy = 0; iy < D.nty; ++iy)
        for (int ix = 0; ix < D.ntx; ++ix)
            cost[ft_defl_block(D, ix, iy)] += ft_interior<C>(g, ix * C::TX, iy * C::TY) ? 1.0 : boundary_cost;
    std::vector<int> order((size_t)D.k);
    for (int B = 0; B < D.k; ++B) order[B] = B;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
    std::vector<double> load((size_t)G, 0.0);
    std::vector<std::vector<int>> lists((size_t)G);
    for (int B : order) {
        int best = 0;
        for (int c = 1; c < G; ++c) if (load[c] < load[best]) best = c;
        load[best] += cost[B];
        lists[best].push_back(B);
    }
    out.assign((size_t)G + 1 + D.k, 0);
    int off = 0;
    for (int c = 0; c < G; ++c) {
        std::sort(lists[c].begin(), lists[c].end());
        out[c] = off;
        for (int B : lists[c]) out[(size_t)G + 1 + off++] = B;
    }
    out[G] = off;
}

}  // namespace perc
