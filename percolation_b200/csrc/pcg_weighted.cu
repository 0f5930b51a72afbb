// pcg_weighted.cu -- Kirchhoff solve with PER-BOND conductances.
//
// The reference's MATLAB post-processor can draw the conductance of every bond of the spanning cluster at random
// (MATLAB/ConductCalc.m:38-47 `condtype = 2`; :94-96, :117-119, :139-141: G(i,j) = -g0*rand for a conducting bond, -1e-12 for
// every other lattice bond; diagonal = -rowsum :163-166; right-hand side :100-104).  With perc_set_bond_conductance the
// caller hands over w(nb), the conductance bond row i of the reference's bond list has WHEN it conducts (which bonds
// conduct is still decided by the labels: build_cfull_kernel); everything else -- unknowns, leak bonds, linbcg's
// stopping rule (Sq/bondc.f:750-838), the read-out with its 1e-10 drop rule (:554-592) -- is as in pcg.cu.
//
// Layout: the matrix is four weight planes W[dir][t] (dir = E, N, NW, NE: the bond OWNED by the site in that direction,
// 0 where the lattice has no such bond) plus the diagonal D[t]; a site reads its own planes and the planes of the west /
// south (/ south-west / south-east) neighbours that own its other bonds.  Algorithmic traffic per site and iteration
// (square): SpMV  r 8 + p 8 + D 8 + W 16 read, p 8 + q 8 written = 56 B;  update  p 8 + q 8 + r 8 + x 8 + D 8 read,
// r 8 + x 8 written = 56 B  => 112 B (144 B triangular) against 50 B for the uniform-g0 kernels, whose matrix is one
// byte per site.  Plain two-kernel Jacobi-PCG with a stored q; no tile pipeline yet.
#include <cmath>
#include "context.h"

namespace perc {

namespace {

constexpr int W_THREADS = 256;

__device__ __forceinline__ double w_block_sum(double v, double* sh)
{
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += sh[k];
    return s;    // valid in thread 0
}
__device__ __forceinline__ bool w_last_block(unsigned* ticket)
{
    __shared__ unsigned s_last;
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
        __threadfence();
    }
    __syncthreads();
    return s_last != 0;
}
__device__ __forceinline__ double w_fold(const double* partial, int cnt, int nq, int q, double* sh)
{
    double v = 0.0;
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) v += __ldcg(&partial[(int64_t)k * nq + q]);
    return w_block_sum(v, sh);
}

// the eight neighbours of site i = (x, y): index and the (plane, owner) that holds the weight of the bond to it
struct WNb { int64_t j[8]; int64_t wi[8]; };
__device__ __forceinline__ void w_neighbours(const Geom& g, int x, int64_t i, WNb& nb)
{
    const int xl = x > 0 ? x - 1 : g.m - 1, xr = x + 1 < g.m ? x + 1 : 0;
    const int64_t row = i - x, t = g.t, m = g.m;
    // order of the bits NB_E, NB_N, NB_NW, NB_NE, NB_W, NB_S, NB_SW, NB_SE
    nb.j[0] = row + xr;      nb.wi[0] = DIR_E * t + i;
    nb.j[1] = i + m;         nb.wi[1] = DIR_N * t + i;
    nb.j[2] = row + m + xl;  nb.wi[2] = DIR_NW * t + i;
    nb.j[3] = row + m + xr;  nb.wi[3] = DIR_NE * t + i;
    nb.j[4] = row + xl;      nb.wi[4] = DIR_E * t + nb.j[4];
    nb.j[5] = i - m;         nb.wi[5] = DIR_N * t + nb.j[5];
    nb.j[6] = row - m + xl;  nb.wi[6] = DIR_NE * t + nb.j[6];      // the south-west neighbour owns the bond as its NE
    nb.j[7] = row - m + xr;  nb.wi[7] = DIR_NW * t + nb.j[7];      // the south-east neighbour owns it as its NW
}

// planes: bond row r of the reference's list -> (owner, dir); conducting -> the caller's w(r), else the leak
__global__ void __launch_bounds__(W_THREADS)
w_planes_kernel(Geom g, const uint8_t* __restrict__ cfull, const double* __restrict__ wb, double gleak, double* __restrict__ W)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= g.nb) return;
    int64_t a; int dir;
    ref_row_to_owner(g, r, &a, &dir);
    const unsigned bit = dir == DIR_E ? NB_E : dir == DIR_N ? NB_N : dir == DIR_NW ? NB_NW : NB_NE;
    W[(int64_t)dir * g.t + a] = (cfull[a] & bit) ? wb[r] : gleak;
}

// diagonal = sum of the weights of the site's bonds, in ascending neighbour number like the reference's row sums
// (geometry.cuh: diag_seq; away from the periodic seam: square S W E N; triangular even x S W E NW N NE, odd x SW S SE W E N)
__global__ void __launch_bounds__(W_THREADS)
w_diag_kernel(Geom g, const double* __restrict__ W, double* __restrict__ D)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.t) return;
    const int x = (int)(i % g.m), y = (int)(i / g.m);
    const unsigned ex = neighbour_bits(g, x, y);
    WNb nb; w_neighbours(g, x, i, nb);
    const unsigned bits[8] = {NB_E, NB_N, NB_NW, NB_NE, NB_W, NB_S, NB_SW, NB_SE};
    double w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = (ex & bits[k]) ? W[nb.wi[k]] : 0.0;
    const double E = w[0], N = w[1], NW = w[2], NE = w[3], Wt = w[4], S = w[5], SW = w[6], SE = w[7];
    double d;
    if (g.lattice != LAT_TRIANGULAR) d = ((S + Wt) + E) + N;
    else if (!(x & 1)) d = ((((S + Wt) + E) + NW) + N) + NE;
    else d = ((((SW + S) + SE) + Wt) + E) + N;
    D[i] = d;
}

// r = b (bonds from row n-2 into the top row at Va), x = p = q = 0; bnrm = |D^-1 b|
__global__ void __launch_bounds__(W_THREADS)
w_init_kernel(Geom g, double Va, const double* __restrict__ W, const double* __restrict__ D, double* __restrict__ vx,
              double* __restrict__ vr, double* __restrict__ vp, double* __restrict__ vq, double* __restrict__ partial,
              PcgState* __restrict__ st, double tol, int itmax)
{
    __shared__ double sh[32];
    double s_b = 0.0, s_rz = 0.0, s_rr = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < g.t; i += stride) {
        const int x = (int)(i % g.m), y = (int)(i / g.m);
        double b = 0.0;
        if (y == g.n - 2 && g.n >= 3) {
            const unsigned ex = neighbour_bits(g, x, y);
            WNb nb; w_neighbours(g, x, i, nb);
            if (ex & NB_NW) b = b + W[nb.wi[2]] * Va;
            if (ex & NB_N) b = b + W[nb.wi[1]] * Va;
            if (ex & NB_NE) b = b + W[nb.wi[3]] * Va;
            const double z = b / D[i];
            s_b += z * z; s_rz += b * z; s_rr += b * b;
        }
        vx[i] = 0.0; vr[i] = b; vp[i] = 0.0; vq[i] = 0.0;
    }
    const double a = w_block_sum(s_b, sh), c = w_block_sum(s_rz, sh), e = w_block_sum(s_rr, sh);
    if (threadIdx.x == 0) { partial[blockIdx.x * 3 + 0] = a; partial[blockIdx.x * 3 + 1] = c; partial[blockIdx.x * 3 + 2] = e; }
    if (w_last_block(&st->ticket_a)) {
        const double fa = w_fold(partial, gridDim.x, 3, 0, sh), fc = w_fold(partial, gridDim.x, 3, 1, sh), fe = w_fold(partial, gridDim.x, 3, 2, sh);
        if (threadIdx.x == 0) {
            st->bnrm = sqrt(fa);
            st->bknum = fc; st->bkden = 1.0; st->bk = 0.0; st->rr = fe;
            st->akden = 0.0; st->ak = 0.0; st->err = 0.0;
            st->iter = 0; st->itmax = itmax; st->tol = tol; st->done = 0;
            st->Itop = 0.0; st->Ibot = 0.0;
        }
    }
}

// p = r / d + bk p  (linbcg :784-800), one pass over the unknown rows
__global__ void __launch_bounds__(W_THREADS)
w_pupdate_kernel(Geom g, const double* __restrict__ D, const double* __restrict__ vr, double* __restrict__ vp, const PcgState* __restrict__ st)
{
    if (st->done) return;
    const double bk = st->bk;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + g.m;
    if (i >= g.t - g.m) return;
    vp[i] = vr[i] / D[i] + bk * vp[i];
}

// q = A p, p.q
__global__ void __launch_bounds__(W_THREADS)
w_spmv_kernel(Geom g, const double* __restrict__ W, const double* __restrict__ D, const double* __restrict__ vp,
              double* __restrict__ vq, double* __restrict__ partial, PcgState* __restrict__ st)
{
    if (st->done) return;
    __shared__ double sh[32];
    double dot = 0.0;
    const int64_t lo = g.m, hi = g.t - g.m, stride = (int64_t)gridDim.x * blockDim.x;
    const unsigned bits[8] = {NB_E, NB_N, NB_NW, NB_NE, NB_W, NB_S, NB_SW, NB_SE};
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
        const int x = (int)(i % g.m), y = (int)(i / g.m);
        const unsigned ex = neighbour_bits(g, x, y);
        WNb nb; w_neighbours(g, x, i, nb);
        const double pc = vp[i];
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (ex & bits[k]) {
                const int64_t j = nb.j[k];
                const double pj = (j >= lo && j < hi) ? vp[j] : 0.0;       // Dirichlet rows: p = 0
                acc += W[nb.wi[k]] * pj;
            }
        const double qv = D[i] * pc - acc;
        vq[i] = qv;
        dot += pc * qv;
    }
    const double bs = w_block_sum(dot, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = bs;
    if (w_last_block(&st->ticket_a)) {
        const double tot = w_fold(partial, gridDim.x, 1, 0, sh);
        if (threadIdx.x == 0) { st->akden = tot; st->ak = st->bknum / tot; }
    }
}

// x += ak p; r -= ak q; next bknum = sum r.r/d; err = |r| / bnrm   (linbcg :808-816)
__global__ void __launch_bounds__(W_THREADS)
w_update_kernel(Geom g, const double* __restrict__ D, double* __restrict__ vx, double* __restrict__ vr,
                const double* __restrict__ vp, const double* __restrict__ vq, double* __restrict__ partial, PcgState* __restrict__ st)
{
    if (st->done) return;
    __shared__ double sh[32];
    const double ak = st->ak;
    double s_rz = 0.0, s_rr = 0.0;
    const int64_t lo = g.m, hi = g.t - g.m, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
        const double r = vr[i] - ak * vq[i];
        vx[i] += ak * vp[i];
        vr[i] = r;
        s_rz += r * (r / D[i]);
        s_rr += r * r;
    }
    const double a = w_block_sum(s_rz, sh), c = w_block_sum(s_rr, sh);
    if (threadIdx.x == 0) { partial[blockIdx.x * 2 + 0] = a; partial[blockIdx.x * 2 + 1] = c; }
    if (w_last_block(&st->ticket_b)) {
        const double fa = w_fold(partial, gridDim.x, 2, 0, sh), fc = w_fold(partial, gridDim.x, 2, 1, sh);
        if (threadIdx.x == 0) {
            const int it = st->iter + 1;
            const double err = sqrt(fc) / st->bnrm;
            st->iter = it; st->err = err; st->rr = fc;
            st->bkden = st->bknum; st->bknum = fa; st->bk = fa / st->bkden;
            if (!(err > st->tol) || it > st->itmax) st->done = 1;     // loop guard iter <= itmax (:780)
        }
    }
}

// Iout = G~ V on rows 0 and n-1 (full diagonal, off-diagonals with |g| >= read_thresh only: second sprsin, :576)
__global__ void __launch_bounds__(256)
w_readout_kernel(Geom g, double Va, double read_thresh, const double* __restrict__ W, const double* __restrict__ D,
                 const double* __restrict__ vx, PcgState* __restrict__ st)
{
    __shared__ double sh[32];
    double stop = 0.0, sbot = 0.0;
    const unsigned bits[8] = {NB_E, NB_N, NB_NW, NB_NE, NB_W, NB_S, NB_SW, NB_SE};
    for (int x = threadIdx.x; x < g.m; x += blockDim.x)
        for (int e = 0; e < 2; ++e) {
            const int y = e == 0 ? 0 : g.n - 1;
            const int64_t i = (int64_t)y * g.m + x;
            const unsigned ex = neighbour_bits(g, x, y);
            WNb nb; w_neighbours(g, x, i, nb);
            double acc = D[i] * (e == 0 ? 0.0 : Va);
            for (int k = 0; k < 8; ++k)
                if (ex & bits[k]) {
                    const double w = W[nb.wi[k]];
                    const int yj = (int)(nb.j[k] / g.m);
                    const double vj = yj == 0 ? 0.0 : (yj == g.n - 1 ? Va : vx[nb.j[k]]);
                    if (fabs(w) >= read_thresh) acc -= w * vj;
                }
            if (e == 0) sbot += acc; else stop += acc;
        }
    const double a = w_block_sum(stop, sh), b = w_block_sum(sbot, sh);
    if (threadIdx.x == 0) { st->Itop = a; st->Ibot = b; }
}

unsigned w_nblk(int64_t n) { return (unsigned)((n + W_THREADS - 1) / W_THREADS); }

}  // namespace

// the caller's per-bond conductances (host array, reference bond-row order); NULL drops them
int pcg_set_bond_weights(Ctx* c, const double* w)
{
    if (c->nranks > 1) return -1;
    if (!w) { c->have_bond_w = false; return 0; }
    const size_t bytes = sizeof(double) * (size_t)c->g.nb;
    if (!c->bond_w) PERC_CUDA(cudaMalloc(&c->bond_w, bytes));
    PERC_CUDA(cudaMemcpyAsync(c->bond_w, w, bytes, cudaMemcpyHostToDevice, c->stream));
    PERC_CUDA(cudaStreamSynchronize(c->stream));           // the caller's array may be pageable and reused
    c->have_bond_w = true;
    return 0;
}

// the solve (called by pcg_solve after build_cfull_kernel when the handle holds per-bond conductances); always keeps x
int pcg_solve_weighted(Ctx* c, double Va, double gleak, double tol, int itmax, double read_thresh)
{
    const Geom& g = c->g;
    cudaStream_t s = c->stream;
    if (!c->wplane) PERC_CUDA(cudaMalloc(&c->wplane, sizeof(double) * 4 * (size_t)g.t));
    if (!c->wdiag) PERC_CUDA(cudaMalloc(&c->wdiag, sizeof(double) * (size_t)g.t));
    if (!c->vq) PERC_CUDA(cudaMalloc(&c->vq, sizeof(double) * (size_t)g.t));
    const int grid = 148 * 8;
    if (grid * 3 > c->partial_cap) {
        if (c->partial) cudaFree(c->partial);
        c->partial = nullptr; c->partial_cap = 0;
        PERC_CUDA(cudaMalloc(&c->partial, sizeof(double) * grid * 3));
        c->partial_cap = grid * 3;
    }
    PERC_CUDA(cudaMemsetAsync(c->wplane, 0, sizeof(double) * 4 * (size_t)g.t, s));
    w_planes_kernel<<<w_nblk(g.nb), W_THREADS, 0, s>>>(g, c->cfull, c->bond_w, gleak, c->wplane);
    w_diag_kernel<<<w_nblk(g.t), W_THREADS, 0, s>>>(g, c->wplane, c->wdiag);
    w_init_kernel<<<grid, W_THREADS, 0, s>>>(g, Va, c->wplane, c->wdiag, c->vx, c->vr, c->vp, c->vq, c->partial, c->d_pcg, tol, itmax);
    c->launches += 3;
    PERC_CUDA(cudaGetLastError());
    const int64_t interior = g.t - 2 * (int64_t)g.m;
    float sp_ms = 0.f, up_ms = 0.f; int nsamp = 0;
    int chunk = 16, iters_before = 0;
    for (;;) {
        for (int k = 0; k < chunk; ++k) {
            const bool sample = k == chunk / 2;
            if (sample) PERC_CUDA(cudaEventRecord(c->ev[8], s));
            w_pupdate_kernel<<<w_nblk(interior), W_THREADS, 0, s>>>(g, c->wdiag, c->vr, c->vp, c->d_pcg);
            w_spmv_kernel<<<grid, W_THREADS, 0, s>>>(g, c->wplane, c->wdiag, c->vp, c->vq, c->partial, c->d_pcg);
            if (sample) PERC_CUDA(cudaEventRecord(c->ev[9], s));
            w_update_kernel<<<grid, W_THREADS, 0, s>>>(g, c->wdiag, c->vx, c->vr, c->vp, c->vq, c->partial, c->d_pcg);
            if (sample) PERC_CUDA(cudaEventRecord(c->ev[10], s));
            c->launches += 3;
            PERC_CUDA(cudaGetLastError());
        }
        PERC_CUDA(cudaMemcpyAsync(c->h_pcg, c->d_pcg, sizeof(PcgState), cudaMemcpyDeviceToHost, s));
        PERC_CUDA(cudaStreamSynchronize(s));
        if (!c->h_pcg->done && c->h_pcg->iter == iters_before) return (int)cudaErrorLaunchFailure;
        if (!c->h_pcg->done || c->h_pcg->iter > iters_before + chunk / 2) {
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, c->ev[8], c->ev[9]);
            cudaEventElapsedTime(&b, c->ev[9], c->ev[10]);
            sp_ms += a; up_ms += b; nsamp++;
        }
        iters_before = c->h_pcg->iter;
        if (c->h_pcg->done) break;
        if (chunk < 256) chunk *= 2;
    }
    w_readout_kernel<<<1, 256, 0, s>>>(g, Va, read_thresh, c->wplane, c->wdiag, c->vx, c->d_pcg);
    c->launches++;
    c->phase_ms[6] = nsamp ? sp_ms / nsamp : 0.f;
    c->phase_ms[7] = nsamp ? up_ms / nsamp : 0.f;
    return (int)cudaGetLastError();
}

}  // namespace perc
