// pcg.cu -- Kirchhoff conductance of a spanning cluster: matrix-free Jacobi-PCG in fp64.
//
// Replaces Sq/bondc.f:465-595: dense G(t,t) assembly (:482-505), sprsin (:723-746), linbcg
// (:750-838, itol = 2, Jacobi preconditioner asolve :855-864) and the G*V read-out (:554-592).
// On a symmetric matrix with rr = r the reference's BiCG is plain Jacobi-PCG doing every SpMV
// twice; the recurrences below are the same with one SpMV per iteration.
//
// The matrix is never formed.  EVERY lattice bond is in it: weight g0 if the bond
// belongs to the chosen spanning cluster, else the leak gleak (1e-12), unknowns = all sites of
// rows 1..n-2, rows 0 / n-1 are Dirichlet (0 / Va).  One byte per site (cfull) says which of
// its up-to-6 bonds conduct; weights and the diagonal are rebuilt from it on the fly.
//
//   K6  pcg_spmv_kernel    p <- r/d + bk*p (tile + halo, shared memory), q = A p, sum p.q
//   K7  pcg_update_kernel  x += ak p, r -= ak q, sums r.r/d (next bknum) and r.r (err)
//   K8  pcg_readout_kernel literal Gtop / Gbot incl. the 1e-10 drop rule of the 2nd sprsin
// Reductions are two-stage and fixed-order (per-block partial, last block folds them), so a
// solve is bit-reproducible for a given lattice size.
#include <cmath>
#include "context.h"

namespace perc {

constexpr int SP_TX = 64, SP_TY = 16, SP_THREADS = 256;      // SpMV tile
constexpr int SP_HX = SP_TX + 2, SP_HY = SP_TY + 2;
constexpr int UP_THREADS = 256;

struct PcgParams {
    double g0, gleak, Va, read_thresh;
};

__device__ __forceinline__ double diag_of(unsigned cf, unsigned ex, double g0, double gleak)
{
    int nc = __popc(cf & ex), ne = __popc(ex);
    return (double)nc * g0 + (double)(ne - nc) * gleak;
}

// ------------------------------------------------------------------------------------------
// which bonds conduct (bond: Sq/bondc.f:482-489; site: MATLAB/ConductCalc.m:88-109;
// mixed: MATLAB/ConductCalc.m:134-160) -> 8-direction byte per site
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
build_cfull_kernel(Geom g, int kind, int32_t cid, const uint8_t* __restrict__ mask, const int32_t* __restrict__ label,
                   uint8_t* __restrict__ cfull)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.t) return;
    int x = (int)(i % g.m), y = (int)(i / g.m);
    unsigned ex = neighbour_bits(g, x, y);
    unsigned out = 0;
    bool mine = label[i] == cid;
    unsigned mk = mask[i];
    int xl = x > 0 ? x - 1 : g.m - 1, xr = x + 1 < g.m ? x + 1 : 0;
    int64_t row = i - x;
#define OWN(bit, nbbit, j)                                                                                  \
    if (ex & nbbit) {                                                                                       \
        bool on;                                                                                            \
        if (kind == KIND_SITE) on = mine && label[j] == cid;                                                \
        else if (kind == KIND_BOND) on = (mk & bit) && mine;                                                \
        else on = (mk & bit) && mine && label[j] == cid;                                                    \
        if (on) out |= nbbit;                                                                               \
    }
#define INC(bit, nbbit, j)                                                                                  \
    if (ex & nbbit) {                                                                                       \
        bool on;                                                                                            \
        if (kind == KIND_SITE) on = mine && label[j] == cid;                                                \
        else if (kind == KIND_BOND) on = (mask[j] & bit) && label[j] == cid;                                \
        else on = (mask[j] & bit) && mine && label[j] == cid;                                               \
        if (on) out |= nbbit;                                                                               \
    }
    OWN(MASK_E, NB_E, row + xr)
    OWN(MASK_N, NB_N, i + g.m)
    OWN(MASK_NW, NB_NW, row + g.m + xl)
    OWN(MASK_NE, NB_NE, row + g.m + xr)
    INC(MASK_E, NB_W, row + xl)
    INC(MASK_N, NB_S, i - g.m)
    INC(MASK_NE, NB_SW, row - g.m + xl)
    INC(MASK_NW, NB_SE, row - g.m + xr)
#undef OWN
#undef INC
    cfull[i] = (uint8_t)out;
}

// ------------------------------------------------------------------------------------------
// block reduction helpers (fixed order)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* sh)
{
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += sh[k];
    return s;    // valid in thread 0
}

// last block folds `cnt` partials (stride `stride` doubles apart, `nq` interleaved quantities)
__device__ __forceinline__ bool last_block(unsigned* ticket)
{
    __shared__ unsigned s_last;
    __threadfence();
    if (threadIdx.x == 0) s_last = atomicInc(ticket, gridDim.x * gridDim.y - 1) == gridDim.x * gridDim.y - 1;
    __syncthreads();
    return s_last != 0;
}

__device__ __forceinline__ double fold_partials(const double* partial, int cnt, int nq, int q, double* sh)
{
    double v = 0.0;
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) v += __ldcg(&partial[(int64_t)k * nq + q]);
    return block_sum(v, sh);
}

// ------------------------------------------------------------------------------------------
// init: r = b (Sq/bondc.f:490-497), x = 0, p = 0, q = 0; bnrm = |D^-1 b| (:769-770)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(UP_THREADS)
pcg_init_kernel(Geom g, PcgParams prm, const uint8_t* __restrict__ cfull, double* __restrict__ vx,
                double* __restrict__ vr, double* __restrict__ vp, double* __restrict__ vq,
                double* __restrict__ partial, PcgState* __restrict__ st, double tol, int itmax)
{
    __shared__ double sh[32];
    double s_b = 0.0, s_rz = 0.0, s_rr = 0.0;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < g.t; i += stride) {
        int x = (int)(i % g.m), y = (int)(i / g.m);
        double b = 0.0;
        if (y == g.n - 2 && y >= 1) {
            unsigned ex = neighbour_bits(g, x, y), cf = cfull[i];
            // bonds into the top row: N, NW, NE
            if (ex & NB_N)  b += ((cf & NB_N)  ? prm.g0 : prm.gleak) * prm.Va;
            if (ex & NB_NW) b += ((cf & NB_NW) ? prm.g0 : prm.gleak) * prm.Va;
            if (ex & NB_NE) b += ((cf & NB_NE) ? prm.g0 : prm.gleak) * prm.Va;
            double d = diag_of(cf, ex, prm.g0, prm.gleak);
            double z = b / d;
            s_b += z * z; s_rz += b * z; s_rr += b * b;
        }
        vx[i] = 0.0; vr[i] = b; vp[i] = 0.0; vq[i] = 0.0;
    }
    double a = block_sum(s_b, sh), c = block_sum(s_rz, sh), e = block_sum(s_rr, sh);
    if (threadIdx.x == 0) { partial[blockIdx.x * 3 + 0] = a; partial[blockIdx.x * 3 + 1] = c; partial[blockIdx.x * 3 + 2] = e; }
    if (last_block(&st->ticket_a)) {
        double fa = fold_partials(partial, gridDim.x, 3, 0, sh);
        double fc = fold_partials(partial, gridDim.x, 3, 1, sh);
        double fe = fold_partials(partial, gridDim.x, 3, 2, sh);
        if (threadIdx.x == 0) {
            st->bnrm = sqrt(fa);
            st->bknum = fc; st->bkden = 1.0; st->bk = 0.0; st->rr = fe;
            st->akden = 0.0; st->ak = 0.0; st->err = 0.0;
            st->iter = 0; st->itmax = itmax; st->tol = tol; st->done = 0;
            st->Itop = 0.0; st->Ibot = 0.0;
        }
    }
}

// ------------------------------------------------------------------------------------------
// K6: fused p-update + stencil SpMV + dot
// ------------------------------------------------------------------------------------------
template <int LAT>
__global__ void __launch_bounds__(SP_THREADS)
pcg_spmv_kernel(Geom g, PcgParams prm, const uint8_t* __restrict__ cfull, const double* __restrict__ vr,
                const double* __restrict__ vp_old, double* __restrict__ vp, double* __restrict__ vq,
                double* __restrict__ partial, PcgState* __restrict__ st)
{
    if (st->done) return;
    __shared__ double pn[SP_HY * SP_HX];
    __shared__ double sh[32];
    const double bk = st->bk;
    const int x0 = blockIdx.x * SP_TX, y0 = blockIdx.y * SP_TY;
    // p_new on tile + halo
    for (int k = threadIdx.x; k < SP_HY * SP_HX; k += SP_THREADS) {
        int ly = k / SP_HX - 1, lx = k % SP_HX - 1;
        int gy = y0 + ly, gx = x0 + lx;
        if (g.pbc) { if (gx == -1) gx = g.m - 1; else if (gx == g.m) gx = 0; }
        double v = 0.0;
        if (gx >= 0 && gx < g.m && gy >= 1 && gy < g.n - 1) {
            int64_t j = (int64_t)gy * g.m + gx;
            unsigned ex = neighbour_bits(g, gx, gy), cf = cfull[j];
            double d = diag_of(cf, ex, prm.g0, prm.gleak);
            v = vr[j] / d + bk * vp_old[j];      // p is double-buffered: neighbours' tiles write vp
        }
        pn[k] = v;
    }
    __syncthreads();
    double dot = 0.0;
    for (int k = threadIdx.x; k < SP_TX * SP_TY; k += SP_THREADS) {
        int ly = k / SP_TX, lx = k % SP_TX;
        int gy = y0 + ly, gx = x0 + lx;
        if (gx >= g.m || gy < 1 || gy >= g.n - 1) continue;
        int64_t i = (int64_t)gy * g.m + gx;
        unsigned ex = neighbour_bits(g, gx, gy), cf = cfull[i];
        const double* c = &pn[(ly + 1) * SP_HX + lx + 1];
        double pc = c[0];
        double acc = 0.0;
#define NB(bit, off) if (ex & bit) acc += ((cf & bit) ? prm.g0 : prm.gleak) * c[off];
        NB(NB_E, 1) NB(NB_W, -1) NB(NB_N, SP_HX) NB(NB_S, -SP_HX)
        if (LAT == LAT_TRIANGULAR) { NB(NB_NW, SP_HX - 1) NB(NB_NE, SP_HX + 1) NB(NB_SW, -SP_HX - 1) NB(NB_SE, -SP_HX + 1) }
#undef NB
        double qv = diag_of(cf, ex, prm.g0, prm.gleak) * pc - acc;
        vp[i] = pc;
        vq[i] = qv;
        dot += pc * qv;
    }
    double bs = block_sum(dot, sh);
    int bid = blockIdx.y * gridDim.x + blockIdx.x;
    if (threadIdx.x == 0) partial[bid] = bs;
    if (last_block(&st->ticket_a)) {
        double tot = fold_partials(partial, gridDim.x * gridDim.y, 1, 0, sh);
        if (threadIdx.x == 0) { st->akden = tot; st->ak = st->bknum / tot; }
    }
}

// ------------------------------------------------------------------------------------------
// K7: x += ak p; r -= ak q; next bknum = sum r.r/d; err = |r| / bnrm   (linbcg :808-816)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(UP_THREADS)
pcg_update_kernel(Geom g, PcgParams prm, const uint8_t* __restrict__ cfull, double* __restrict__ vx,
                  double* __restrict__ vr, const double* __restrict__ vp, const double* __restrict__ vq,
                  double* __restrict__ partial, PcgState* __restrict__ st)
{
    if (st->done) return;
    __shared__ double sh[32];
    const double ak = st->ak;
    double s_rz = 0.0, s_rr = 0.0;
    int64_t lo = g.m, hi = g.t - g.m;
    int64_t per = (hi - lo + gridDim.x - 1) / gridDim.x;
    int64_t b0 = lo + (int64_t)blockIdx.x * per, b1 = b0 + per < hi ? b0 + per : hi;
    for (int64_t i = b0 + threadIdx.x; i < b1; i += UP_THREADS) {
        int x = (int)(i % g.m), y = (int)(i / g.m);
        double p = vp[i], q = vq[i];
        double r = vr[i] - ak * q;
        vx[i] += ak * p;
        vr[i] = r;
        double d = diag_of(cfull[i], neighbour_bits(g, x, y), prm.g0, prm.gleak);
        s_rz += r * (r / d);
        s_rr += r * r;
    }
    double a = block_sum(s_rz, sh), c = block_sum(s_rr, sh);
    if (threadIdx.x == 0) { partial[blockIdx.x * 2 + 0] = a; partial[blockIdx.x * 2 + 1] = c; }
    if (last_block(&st->ticket_b)) {
        double fa = fold_partials(partial, gridDim.x, 2, 0, sh);
        double fc = fold_partials(partial, gridDim.x, 2, 1, sh);
        if (threadIdx.x == 0) {
            int it = st->iter + 1;
            double err = sqrt(fc) / st->bnrm;
            st->iter = it; st->err = err; st->rr = fc;
            st->bkden = st->bknum; st->bknum = fa; st->bk = fa / st->bkden;
            if (!(err > st->tol) || it > st->itmax) st->done = 1;     // loop guard iter <= itmax (:780)
        }
    }
}

// ------------------------------------------------------------------------------------------
// K8: read-out (Sq/bondc.f:554-592): Iout = G~ V on rows 0 and n-1, G~ keeps the full
// diagonal but only off-diagonals with |g| >= read_thresh (second sprsin, :576)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pcg_readout_kernel(Geom g, PcgParams prm, const uint8_t* __restrict__ cfull, const double* __restrict__ vx,
                   PcgState* __restrict__ st)
{
    __shared__ double sh[32];
    double stop = 0.0, sbot = 0.0;
    for (int x = threadIdx.x; x < g.m; x += blockDim.x) {
        for (int e = 0; e < 2; ++e) {
            int y = e == 0 ? 0 : g.n - 1;
            int64_t i = (int64_t)y * g.m + x;
            unsigned ex = neighbour_bits(g, x, y), cf = cfull[i];
            double vi = e == 0 ? 0.0 : prm.Va;
            double acc = diag_of(cf, ex, prm.g0, prm.gleak) * vi;
            int xl = x > 0 ? x - 1 : g.m - 1, xr = x + 1 < g.m ? x + 1 : 0;
            int64_t row = i - x;
#define VAL(j) ((j) < g.m ? 0.0 : ((j) >= g.t - g.m ? prm.Va : vx[j]))
#define NB(bit, j) if (ex & bit) { double w = (cf & bit) ? prm.g0 : prm.gleak; if (fabs(w) >= prm.read_thresh) acc -= w * VAL(j); }
            NB(NB_E, row + xr) NB(NB_W, row + xl) NB(NB_N, i + g.m) NB(NB_S, i - g.m)
            NB(NB_NW, row + g.m + xl) NB(NB_NE, row + g.m + xr) NB(NB_SW, row - g.m + xl) NB(NB_SE, row - g.m + xr)
#undef NB
#undef VAL
            if (e == 0) sbot += acc; else stop += acc;
        }
    }
    double a = block_sum(stop, sh), b = block_sum(sbot, sh);
    if (threadIdx.x == 0) { st->Itop = a; st->Ibot = b; }
}

// ------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------
static unsigned nblk(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

int pcg_solve(Ctx* c, int cluster_id, double Va, double g0, double gleak, double tol, int itmax,
              double read_thresh, double* Gtop, double* Gbot, int* iter, double* err)
{
    const Geom& g = c->g;
    cudaStream_t s = c->stream;
    PcgParams prm{g0, gleak, Va, read_thresh};
    dim3 sgrid((g.m + SP_TX - 1) / SP_TX, (g.n + SP_TY - 1) / SP_TY);
    int ugrid = 148 * 8;
    int64_t interior = g.t - 2 * (int64_t)g.m;
    if (ugrid > (interior + UP_THREADS - 1) / UP_THREADS) ugrid = (int)((interior + UP_THREADS - 1) / UP_THREADS);
    if (ugrid < 1) ugrid = 1;
    int need = (int)(sgrid.x * sgrid.y);
    if (need < ugrid * 3) need = ugrid * 3;
    if (need > c->partial_cap) {
        if (c->partial) cudaFree(c->partial);
        PERC_CUDA(cudaMalloc(&c->partial, sizeof(double) * need));
        c->partial_cap = need;
    }
    PERC_CUDA(cudaMemsetAsync(c->d_pcg, 0, sizeof(PcgState), s));
    PERC_CUDA(cudaEventRecord(c->ev[6], s));
    build_cfull_kernel<<<nblk(g.t), 256, 0, s>>>(g, c->kind, cluster_id, c->mask, c->label, c->cfull);
    pcg_init_kernel<<<ugrid, UP_THREADS, 0, s>>>(g, prm, c->cfull, c->vx, c->vr, c->vp, c->vq, c->partial, c->d_pcg, tol, itmax);
    c->launches += 2;
    double* pold = c->vp; double* pnew = c->vp2;
    float sp_ms = 0.f, up_ms = 0.f; int nsamp = 0;
    int chunk = 16;
    for (;;) {
        for (int k = 0; k < chunk; ++k) {
            // the first iteration of a chunk is always live (we stop launching once done is seen):
            // bracket its two kernels with events -> average per-kernel device time
            bool sample = (k == 0);
            if (sample) PERC_CUDA(cudaEventRecord(c->ev[8], s));
            if (g.lattice == LAT_SQUARE)
                pcg_spmv_kernel<LAT_SQUARE><<<sgrid, SP_THREADS, 0, s>>>(g, prm, c->cfull, c->vr, pold, pnew, c->vq, c->partial, c->d_pcg);
            else
                pcg_spmv_kernel<LAT_TRIANGULAR><<<sgrid, SP_THREADS, 0, s>>>(g, prm, c->cfull, c->vr, pold, pnew, c->vq, c->partial, c->d_pcg);
            if (sample) PERC_CUDA(cudaEventRecord(c->ev[9], s));
            pcg_update_kernel<<<ugrid, UP_THREADS, 0, s>>>(g, prm, c->cfull, c->vx, c->vr, pnew, c->vq, c->partial, c->d_pcg);
            { double* tmp = pold; pold = pnew; pnew = tmp; }
            if (sample) PERC_CUDA(cudaEventRecord(c->ev[10], s));
            c->launches += 2;
        }
        PERC_CUDA(cudaMemcpyAsync(c->h_pcg, c->d_pcg, sizeof(PcgState), cudaMemcpyDeviceToHost, s));
        PERC_CUDA(cudaStreamSynchronize(s));
        {
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, c->ev[8], c->ev[9]);
            cudaEventElapsedTime(&b, c->ev[9], c->ev[10]);
            sp_ms += a; up_ms += b; nsamp++;
        }
        if (c->h_pcg->done) break;
        if (chunk < 256) chunk *= 2;
    }
    pcg_readout_kernel<<<1, 256, 0, s>>>(g, prm, c->cfull, c->vx, c->d_pcg);
    c->launches++;
    PERC_CUDA(cudaEventRecord(c->ev[7], s));
    PERC_CUDA(cudaMemcpyAsync(c->h_pcg, c->d_pcg, sizeof(PcgState), cudaMemcpyDeviceToHost, s));
    PERC_CUDA(cudaStreamSynchronize(s));
    PERC_CUDA(cudaGetLastError());
    cudaEventElapsedTime(&c->phase_ms[5], c->ev[6], c->ev[7]);
    c->phase_ms[6] = nsamp ? sp_ms / nsamp : 0.f;
    c->phase_ms[7] = nsamp ? up_ms / nsamp : 0.f;
    *Gtop = c->h_pcg->Itop / Va;
    *Gbot = fabs(c->h_pcg->Ibot) / Va;
    *iter = c->h_pcg->iter;
    *err = c->h_pcg->err;
    c->solved = true;
    return 0;
}

}  // namespace perc
