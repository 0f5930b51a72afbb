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
//   K6  pcg_pipe_kernel<0> p <- r/d + bk*p (tile + halo, shared memory), sum p.(A p)
//   K7  pcg_pipe_kernel<1> r -= ak (A p) with A p recomputed (q is never stored), x += ak p,
//                          sums r.r/d (next bknum) and r.r (err)
//       (pcg_spmv_kernel / pcg_update_kernel: scalar fallback with a stored q when m is not a multiple of 16)
//   K8  pcg_readout_kernel literal Gtop / Gbot incl. the 1e-10 drop rule of the 2nd sprsin
// Reductions are two-stage and fixed-order (per-block partial, last block folds them), so a
// solve is bit-reproducible for a given lattice size.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cooperative_groups.h>
#include "context.h"
#include "pcg_fused_tile.cuh"
#include "pcg_defl_host.h"

namespace perc {

static unsigned nblk64(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

constexpr int SP_TX = 64, SP_TY = 16, SP_THREADS = 256;      // SpMV tile
constexpr int SP_HX = SP_TX + 2, SP_HY = SP_TY + 2;
constexpr int UP_THREADS = 256;

struct PcgParams {
    double g0, gleak, Va, read_thresh;
};

// diagonal of site (x, .): the reference's G(i,i) = -rowsum, summed in ascending neighbour number (geometry.cuh: diag_seq)
__device__ __forceinline__ double diag_of(const Geom& g, unsigned cf, unsigned ex, int x, double g0, double gleak)
{
    return diag_seq(g, cf & ex, ex, x, g0, gleak);
}

// ------------------------------------------------------------------------------------------
// which bonds conduct (bond: Sq/bondc.f:482-489; site: MATLAB/ConductCalc.m:88-109;
// mixed: MATLAB/ConductCalc.m:134-160) -> 8-direction byte per site
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
build_cfull_kernel(Geom g, int kind, int32_t cid, const uint8_t* __restrict__ mask, const int32_t* __restrict__ label,
                   uint8_t* __restrict__ cfull, const Summary* __restrict__ dsum)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.t) return;
    if (dsum) {
        // batch mode: the default spanning cluster (smallest canonical id) is chosen on the device; 0 = none spans
        cid = 0;
        if (dsum->nspan > 0) cid = (int32_t)(0xffffffffu - (unsigned)(dsum->span_best >> 32));
    }
    int x = (int)(i % g.m), y = (int)(i / g.m);
    if (y < g.own_lo || y >= g.own_hi) return;          // halo rows: copied from the neighbour rank
    unsigned ex = neighbour_bits(g, x, y);
    unsigned out = 0;
    bool mine = cid != 0 && label[i] == cid;
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
    if (threadIdx.x == 0) {
        __threadfence();          // publish this block's partial (written by thread 0) before taking a ticket
        s_last = atomicInc(ticket, gridDim.x * gridDim.y - 1) == gridDim.x * gridDim.y - 1;
        __threadfence();
    }
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
                double* __restrict__ partial, PcgState* __restrict__ st, double tol, int itmax, int dist, int warm)
{
    __shared__ double sh[32];
    double s_b = 0.0, s_rz = 0.0, s_rr = 0.0;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < g.t; i += stride) {
        int x = (int)(i % g.m), y = (int)(i / g.m);
        double b = 0.0;
        if (g.y0 + y == g.ng - 2 && solve_row(g, y)) {
            unsigned ex = neighbour_bits(g, x, y), cf = cfull[i];
            // bonds into the top row, summed in the reference's order (geometry.cuh: rhs_seq)
            b = rhs_seq(g, cf, ex, x, prm.g0, prm.gleak, prm.Va);
            double d = diag_of(g, cf, ex, x, prm.g0, prm.gleak);
            double z = b / d;
            s_b += z * z;
            if (!warm) { s_rz += b * z; s_rr += b * b; }
        }
        double r = b;
        if (warm) {
            // linbcg with a non-zero initial guess (Sq/bondc.f:759-763: r = b - A x): x holds the previous
            // solution (rows 0 / n-1 of x are 0, so Dirichlet neighbours drop out; their share is in b)
            if (solve_row(g, y)) {
                const unsigned ex = neighbour_bits(g, x, y), cf = cfull[i];
                const int xl = x > 0 ? x - 1 : g.m - 1, xr = x + 1 < g.m ? x + 1 : 0;
                const int64_t row = i - x;
                const double d = diag_of(g, cf, ex, x, prm.g0, prm.gleak);
                double acc = d * vx[i];
#define NB(bit, j) if (ex & bit) acc -= ((cf & bit) ? prm.g0 : prm.gleak) * vx[j];
                NB(NB_E, row + xr) NB(NB_W, row + xl) NB(NB_N, i + g.m) NB(NB_S, i - g.m)
                NB(NB_NW, row + g.m + xl) NB(NB_NE, row + g.m + xr) NB(NB_SW, row - g.m + xl) NB(NB_SE, row - g.m + xr)
#undef NB
                r = b - acc;
                s_rz += r * (r / d); s_rr += r * r;
            }
        } else vx[i] = 0.0;
        vr[i] = r; vp[i] = 0.0;
        if (vq) vq[i] = 0.0;          // only the odd-m fallback stores q
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
            if (dist) { st->red[0] = fa; st->red[1] = fc; st->red[2] = fe; }   // summed over the ranks, then pcg_post
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
            double d = diag_of(g, cf, ex, gx, prm.g0, prm.gleak);
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
        double qv = diag_of(g, cf, ex, gx, prm.g0, prm.gleak) * pc - acc;
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
        double d = diag_of(g, cfull[i], neighbour_bits(g, x, y), x, prm.g0, prm.gleak);
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
// helpers of the pipelined kernels: 2 sites per thread with 128-bit accesses, and a constant neighbourhood
// for tiles that do not touch the lattice boundary.
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }

template <int LAT>
__device__ __forceinline__ unsigned interior_ex(int gx)
{
    if (LAT == LAT_SQUARE) return NB_E | NB_N | NB_W | NB_S;
    return (gx & 1) ? (NB_E | NB_N | NB_W | NB_S | NB_SW | NB_SE) : (NB_E | NB_N | NB_W | NB_S | NB_NW | NB_NE);
}

// ------------------------------------------------------------------------------------------
// K6'/K7' (m % 16 == 0): persistent, double-buffered tile pipeline; the q = A p vector is never stored.
// One CTA per SM walks over the 128 x 32 tiles of the lattice.  While the CTA computes tile k from
// one shared-memory stage, the TMA tensor copies (cp.async.bulk.tensor.2d, completion on an mbarrier,
// zero-filled outside the lattice) of tile k+1 (+ halo) are already in flight into the other stage, so
// ~75 KB per SM are always outstanding and no load latency is exposed to the arithmetic.
//   MODE 0: p_new = r / d + bk * p_old in place on tile + halo, q = A p_new, p.q      (K6)
//   MODE 1: q = A p RECOMPUTED, r -= ak q, x += ak p, sums r.r/d and r.r              (K7)
// HBM traffic per site and iteration:
//   MODE 0: r 8 + p_old 8 + cfull 1 read, p 8 written                      = 25 B
//   MODE 1: p 8 + r 8 + cfull 1 read, r 8 written (+ x 8 read, 8 written)  = 25 B (41 B with keep_x)
// against 82 B for a stored-q pair.  The fp64 stencil is done twice; the SMs have the headroom.
// keep_x = 0 keeps x only on rows 1 and n-2, the rows the read-out (K8) consumes: Gtop / Gbot come
// out bit-identical, the interior voltages are simply not formed.
// Partial sums are stored per TILE and folded in tile order, so the result does not depend on the
// number of CTAs.
// ------------------------------------------------------------------------------------------
constexpr int PT_TX = 128, PT_TY = 36, PT_THREADS = 768, PT_ROWS = PT_TY + 2;
constexpr int PT_RPT = PT_TY / (PT_THREADS / 64);      // consecutive tile rows per thread in the stencil
constexpr int PT_LD = PT_TX + 4;            // doubles per staged row: [1] left halo, [2..129] tile, [130] right halo
constexpr int PT_CLD = PT_TX + 32;          // bytes per staged cfull row: [12..15] left halo word, [16..143] tile, [144..147] right halo word
constexpr int PT_VEC_BYTES = (PT_ROWS * PT_LD * 8 + 127) / 128 * 128;      // one staged fp64 tile (TMA destinations: 128-byte aligned)
constexpr int PT_CF_BYTES = (PT_ROWS * PT_CLD + 127) / 128 * 128;
constexpr int PT_STAGE_BYTES = 2 * PT_VEC_BYTES + PT_CF_BYTES;
// (d, 1/d) table, one private copy per lane (entry [idx][lane]): a lookup is conflict-free whatever the indices
constexpr size_t PT_SMEM = 2 * (size_t)PT_STAGE_BYTES + sizeof(double2) * 64 * 32 + sizeof(double) * (32 + 4 + 64) + 16;

struct PtStage { double* sp; double* sr; uint8_t* scf; };

// acc += v if bit != 0, as ONE predicated DADD (the compiler's select-then-add costs two FSEL more)
__device__ __forceinline__ void padd(double& acc, double v, unsigned bit)
{
    asm("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n @p add.f64 %0, %0, %1;\n}" : "+d"(acc) : "d"(v), "r"(bit));
}

template <int DIST> __device__ __forceinline__ bool pt_solve_row(const Geom& g, int y)
{
    return DIST ? solve_row(g, y) : (y >= 1 && y < g.n - 1);
}
template <int DIST> __device__ __forceinline__ bool pt_p_row(const Geom& g, int y)
{
    return DIST ? p_row(g, y) : (y >= 1 && y < g.n - 1);
}

__device__ __forceinline__ PtStage pt_stage(unsigned char* raw, int k)
{
    PtStage s;
    unsigned char* base = raw + (size_t)k * PT_STAGE_BYTES;
    s.sp = reinterpret_cast<double*>(base);
    s.sr = reinterpret_cast<double*>(base + PT_VEC_BYTES);
    s.scf = reinterpret_cast<uint8_t*>(base + 2 * PT_VEC_BYTES);
    return s;
}

// ---- TMA bulk copies (cp.async.bulk, completion counted on an mbarrier) ---------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one TMA tensor copy: box (w x PT_ROWS) of a row-major 2-D array at element coordinates (cx, cy); elements
// outside the array arrive as zeros
__device__ __forceinline__ void tma_box_g2s(void* dst, const CUtensorMap* map, int cx, int cy, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(cx), "r"(cy), "r"(smem_u32(bar)) : "memory");
}

// Issue the asynchronous copies of one tile (+ halo) into a stage: ONE thread, three TMA tensor copies
// (cp.async.bulk.tensor.2d, SASS UTMALDG): the 132 x 34 fp64 boxes of p and r at (x0-2, y0-1) and the
// 160 x 34 byte box of the conduct mask at (x0-16, y0-1).  The boxes carry the halo rows / columns with
// them, and whatever lies outside the lattice arrives as zeros.  The stage's mbarrier counts the bytes.
template <int MODE>
__device__ __forceinline__ void pt_issue(const PtStage& s, unsigned long long* bar, int x0, int y0,
                                         const CUtensorMap* tm_p, const CUtensorMap* tm_r, const CUtensorMap* tm_cf)
{
    if (threadIdx.x != 0) return;
    mbar_arrive_expect(bar, 2u * PT_ROWS * PT_LD * 8u + PT_ROWS * PT_CLD);
    tma_box_g2s(s.sp, tm_p, x0 - 2, y0 - 1, bar);
    tma_box_g2s(s.sr, tm_r, x0 - 2, y0 - 1, bar);
    tma_box_g2s(s.scf, tm_cf, x0 - 16, y0 - 1, bar);
}

// periodic wrap: the halo column of an edge tile lives at the other end of the lattice row (plain loads:
// two tiles per lattice row, only with pbc)
template <int MODE, int DIST>
__device__ __forceinline__ void pt_wrap_halo(const Geom& g, const PtStage& s, int x0, int y0, const uint8_t* __restrict__ cfull,
                                             const double* __restrict__ vr, const double* __restrict__ vp_in)
{
    const int tid = threadIdx.x;
    if (!g.pbc || tid >= 2 * PT_ROWS) return;
    const int side = tid >= PT_ROWS, pr = tid - side * PT_ROWS;
    const int gy = y0 + pr - 1;
    const int xe = g.m - x0 < PT_TX ? g.m - x0 : PT_TX;
    if (side ? x0 + PT_TX < g.m : x0 != 0) return;                            // inner halo columns came with the row
    if (gy < 0 || gy >= g.n) return;
    const int hx = side ? 0 : g.m - 1;
    const int64_t j = (int64_t)gy * g.m + hx;
    const int col = side ? 2 + xe : 1;
    s.sp[pr * PT_LD + col] = vp_in[j];
    if (MODE == 0) {
        s.sr[pr * PT_LD + col] = vr[j];
        s.scf[pr * PT_CLD + (side ? 16 + xe : 15)] = cfull[j];
    }
}

// DIST = 0: whole lattice on this GPU (row tests fold to 1 <= y <= n-2); DIST = 1: slab of a decomposed lattice
template <int LAT, int MODE, int DIST>
__global__ void __launch_bounds__(PT_THREADS, 1)
pcg_pipe_kernel(const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tm_r,
                const __grid_constant__ CUtensorMap tm_cf, Geom g, PcgParams prm, const uint8_t* __restrict__ cfull,
                double* __restrict__ vr, const double* __restrict__ vp_in, double* __restrict__ vp_out, double* __restrict__ vx,
                double* __restrict__ partial, PcgState* __restrict__ st, int keep_x, int ntx, int ntiles)
{
    if (st->done) return;
    extern __shared__ __align__(128) unsigned char pt_raw[];
    double2* dtab = reinterpret_cast<double2*>(pt_raw + 2 * (size_t)PT_STAGE_BYTES);    // [64][32] (d, 1/d)
    double* sh = reinterpret_cast<double*>(dtab + 64 * 32);
    const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6, lane = tid & 31;
    // (d, 1/d) of the sites with all their neighbours, by conduct pattern (pcg_fused_tile.cuh: ft_pat; 16 x 32 private
    // copies on the square lattice, 128 x 16 on the triangular one); every other site sums its diagonal in place (pt_dsite)
    for (int k = tid; k < ft_tab_slots<LAT, FtCfgA3>(); k += PT_THREADS) {
        const FtDiag e = ft_tab_slot<LAT, FtCfgA3>(g, k, prm.g0, prm.gleak);
        dtab[k] = make_double2(e.d, e.inv);
    }
    double* cinv = sh + 36;                                                             // [64] 1/d by bond counts (any positive scaling serves as the Jacobi preconditioner)
    if (tid >= PT_THREADS - 64) cinv[tid - (PT_THREADS - 64)] = ft_cinv_entry(tid - (PT_THREADS - 64), prm.g0, prm.gleak);
    auto pt_dsite = [&](unsigned cf, unsigned ex, int gx_) {
        const FtDiag e = ft_diag_site(g, cf, ex, gx_, prm.g0, prm.gleak, cinv);
        return make_double2(e.d, e.inv);
    };
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(sh + 32);          // one mbarrier per stage
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const double bk = st->bk, ak = st->ak;          // slab mode: set by pcg_post_kernel after the all-reduce
    // MODE 0 sweeps the tiles from the END of the lattice to its start, MODE 1 front to back: each
    // kernel starts on the data the previous one touched last (still resident in the 126 MB L2)
    auto tile_of = [&](int t) { return MODE == 0 ? ntiles - 1 - t : t; };

    int t = blockIdx.x;
    if (t < ntiles) { int tl = tile_of(t); pt_issue<MODE>(pt_stage(pt_raw, 0), &bars[0], (tl % ntx) * PT_TX, (tl / ntx) * PT_TY, &tm_p, &tm_r, &tm_cf); }
    for (int k = 0; t < ntiles; t += gridDim.x, ++k) {
        const int tl = tile_of(t), x0 = (tl % ntx) * PT_TX, y0 = (tl / ntx) * PT_TY;
        const PtStage s = pt_stage(pt_raw, k & 1);
        const int tn = t + gridDim.x;
        if (tn < ntiles) {                              // the other stage was released by the barriers of the last tile
            int tnl = tile_of(tn);
            pt_issue<MODE>(pt_stage(pt_raw, (k + 1) & 1), &bars[(k + 1) & 1], (tnl % ntx) * PT_TX, (tnl / ntx) * PT_TY, &tm_p, &tm_r, &tm_cf);
        }
        mbar_wait(&bars[k & 1], (unsigned)((k >> 1) & 1));
        pt_wrap_halo<MODE, DIST>(g, s, x0, y0, cfull, vr, vp_in);
        if (g.pbc) __syncthreads();
        // every cell of tile + halo has its full neighbourhood inside the lattice
        const bool interior = x0 >= 2 && x0 + PT_TX <= g.m - 2 && (DIST ? g.y0 : 0) + y0 >= 2 && (DIST ? g.y0 : 0) + y0 + PT_TY <= (DIST ? g.ng : g.n) - 2;
        const int xe = g.m - x0 < PT_TX ? g.m - x0 : PT_TX;

        // ---- MODE 0: p_new = r / d + bk * p_old, in place on tile + halo (p is double-buffered in
        // HBM: neighbouring tiles still read p_old) ---------------------------------------------------
        if (MODE == 0) {
            const int cx = (tid & 63) * 2, prb = tid >> 6, gxc = x0 + cx;
            if (gxc < g.m) {
                const unsigned e0i = interior_ex<LAT>(gxc), e1i = interior_ex<LAT>(gxc + 1);
#pragma unroll
                for (int it = 0; it < (PT_ROWS + PT_THREADS / 64 - 1) / (PT_THREADS / 64); ++it) {
                    const int pr = prb + it * (PT_THREADS / 64), gy = y0 + pr - 1;
                    if (pr >= PT_ROWS) continue;
                    double2 v = make_double2(0.0, 0.0);
                    if (pt_p_row<DIST>(g, gy)) {
                        const double2 r2 = ld2(&s.sr[pr * PT_LD + 2 + cx]), p2 = ld2(&s.sp[pr * PT_LD + 2 + cx]);
                        const unsigned c01 = *reinterpret_cast<const unsigned short*>(&s.scf[pr * PT_CLD + 16 + cx]);
                        const unsigned e0 = interior ? e0i : neighbour_bits(g, gxc, gy);
                        const unsigned e1 = interior ? e1i : neighbour_bits(g, gxc + 1, gy);
                        const unsigned f0 = c01 & 0xffu, f1 = c01 >> 8;
                        v.x = r2.x * ((interior || e0 == e0i) ? dtab[FtCfgA3::tabp(LAT, ft_pat<LAT>(f0, 0), lane)] : pt_dsite(f0, e0, gxc)).y + bk * p2.x;
                        v.y = r2.y * ((interior || e1 == e1i) ? dtab[FtCfgA3::tabp(LAT, ft_pat<LAT>(f1, 1), lane)] : pt_dsite(f1, e1, gxc + 1)).y + bk * p2.y;
                    }
                    st2(&s.sp[pr * PT_LD + 2 + cx], v);
                    // slab mode: the halo copies of p are advanced here (pointwise recurrence) and stored for MODE 1
                    if (DIST && pr >= 1 && pr <= PT_TY && pt_p_row<DIST>(g, gy) && !pt_solve_row<DIST>(g, gy)) st2(vp_out + (int64_t)gy * g.m + gxc, v);
                }
            }
            if (tid < 2 * PT_ROWS) {
                const int side = tid >= PT_ROWS, pr = tid - side * PT_ROWS;
                const int gy = y0 + pr - 1;
                int hx = side ? x0 + xe : x0 - 1;
                if (g.pbc) { if (hx == -1) hx = g.m - 1; else if (hx == g.m) hx = 0; }
                const int col = side ? 2 + xe : 1;
                double v = 0.0;
                if (pt_p_row<DIST>(g, gy) && hx >= 0 && hx < g.m) {
                    const unsigned cf = s.scf[pr * PT_CLD + (side ? 16 + xe : 15)];
                    const unsigned ex = neighbour_bits(g, hx, gy);
                    v = s.sr[pr * PT_LD + col] * pt_dsite(cf, ex, hx).y + bk * s.sp[pr * PT_LD + col];
                }
                s.sp[pr * PT_LD + col] = v;
            }
            __syncthreads();
        }

        // ---- MODE 0: p.Ap as bond energies; MODE 1: q = A p from shared memory, r -= ak q, x += ak p.
        // A thread owns 2 columns x PT_RPT consecutive rows and slides a 3-row window up the tile ------
        const int gx = x0 + 2 * tx;
        double acc0 = 0.0, acc1 = 0.0;
        {
            const int r0 = ty * PT_RPT;
            const double* c = &s.sp[(r0 + 1) * PT_LD + 2 + 2 * tx];
            double2 dn = ld2(c - PT_LD), cc = ld2(c);
            double dlf = c[-PT_LD - 1], drt = c[-PT_LD + 2];      // row below: x-1 and x+2 (triangular diagonals)
#pragma unroll
            for (int j = 0; j < PT_RPT; ++j, c += PT_LD) {
                const int ly = r0 + j, gy = y0 + ly;
                const double2 up = ld2(c + PT_LD);
                const double lf = c[-1], rt = c[2];
                const bool valid = pt_solve_row<DIST>(g, gy) && gx < g.m;
                const unsigned c01 = *reinterpret_cast<const unsigned short*>(&s.scf[(ly + 1) * PT_CLD + 16 + 2 * tx]);
                const unsigned cf0 = c01 & 0xffu, cf1 = c01 >> 8;
                if (MODE == 0) {
                    // p.Ap as the energy of the bonds OWNED by this pair of sites, sum over E, N (NW, NE) of
                    // w (p_i - p_j)^2: half the neighbours of the stencil form, no diagonal lookup, and >= 0
                    // by construction.  Dirichlet rows hold p = 0 in shared memory, so a bond into them
                    // comes out as w p_i^2; bonds of Dirichlet-row owners (row 0 -> row 1) are included.
                    const int G = (DIST ? g.y0 : 0) + gy, NG = DIST ? g.ng : g.n;
                    const bool owner = (DIST ? (gy >= g.own_lo && gy < g.own_hi) : (gy >= 0 && gy < g.n)) && gx < g.m;
                    if (owner) {
                        const double dE0 = cc.x - cc.y, dE1 = cc.y - rt, dN0 = cc.x - up.x, dN1 = cc.y - up.y;
                        double all = 0.0, con = 0.0;
                        if (interior) {
                            all = (dE0 * dE0 + dE1 * dE1) + (dN0 * dN0 + dN1 * dN1);
                            padd(con, dE0 * dE0, cf0 & NB_E); padd(con, dE1 * dE1, cf1 & NB_E);
                            padd(con, dN0 * dN0, cf0 & NB_N); padd(con, dN1 * dN1, cf1 & NB_N);
                            if (LAT == LAT_TRIANGULAR) {
                                const double dNW = cc.x - c[PT_LD - 1], dNE = cc.x - up.y;
                                all += dNW * dNW + dNE * dNE;
                                padd(con, dNW * dNW, cf0 & NB_NW); padd(con, dNE * dNE, cf0 & NB_NE);
                            }
                        } else {
                            const unsigned e0 = neighbour_bits(g, gx, gy), e1 = neighbour_bits(g, gx + 1, gy);
                            const bool rowE = G >= 1 && G <= NG - 2;      // an E bond joins two sites of one row
                            if (rowE && (e0 & NB_E)) { all += dE0 * dE0; padd(con, dE0 * dE0, cf0 & NB_E); }
                            if (rowE && (e1 & NB_E)) { all += dE1 * dE1; padd(con, dE1 * dE1, cf1 & NB_E); }
                            if (e0 & NB_N) { all += dN0 * dN0; padd(con, dN0 * dN0, cf0 & NB_N); }
                            if (e1 & NB_N) { all += dN1 * dN1; padd(con, dN1 * dN1, cf1 & NB_N); }
                            if (LAT == LAT_TRIANGULAR) {
                                const double dNW = cc.x - c[PT_LD - 1], dNE = cc.x - up.y;
                                if (e0 & NB_NW) { all += dNW * dNW; padd(con, dNW * dNW, cf0 & NB_NW); }
                                if (e0 & NB_NE) { all += dNE * dNE; padd(con, dNE * dNE, cf0 & NB_NE); }
                            }
                        }
                        acc0 += prm.g0 * con + prm.gleak * (all - con);
                    }
                    if (valid) st2(vp_out + (int64_t)gy * g.m + gx, cc);
                } else {
                unsigned e0, e1;
                double all0, all1;
                if (interior) {
                    e0 = interior_ex<LAT>(gx); e1 = interior_ex<LAT>(gx + 1);
                    all0 = (cc.y + lf) + (up.x + dn.x);
                    all1 = (rt + cc.x) + (up.y + dn.y);
                    if (LAT == LAT_TRIANGULAR) { all0 += c[PT_LD - 1] + up.y; all1 += dn.x + drt; }
                } else {
                    e0 = neighbour_bits(g, gx, gy); e1 = neighbour_bits(g, gx + 1, gy);
                    all0 = 0.0; all1 = 0.0;
                    if (e0 & NB_E) all0 += cc.y;  if (e0 & NB_W) all0 += lf;  if (e0 & NB_N) all0 += up.x;  if (e0 & NB_S) all0 += dn.x;
                    if (e1 & NB_E) all1 += rt;    if (e1 & NB_W) all1 += cc.x; if (e1 & NB_N) all1 += up.y; if (e1 & NB_S) all1 += dn.y;
                    if (LAT == LAT_TRIANGULAR) {
                        if (e0 & NB_NW) all0 += c[PT_LD - 1]; if (e0 & NB_NE) all0 += up.y;
                        if (e1 & NB_SW) all1 += dn.x;         if (e1 & NB_SE) all1 += drt;
                    }
                }
                double con0 = 0.0, con1 = 0.0;    // conducting neighbours (cfull bits only on existing bonds)
                padd(con0, cc.y, cf0 & NB_E); padd(con0, lf, cf0 & NB_W);   padd(con0, up.x, cf0 & NB_N); padd(con0, dn.x, cf0 & NB_S);
                padd(con1, rt, cf1 & NB_E);   padd(con1, cc.x, cf1 & NB_W); padd(con1, up.y, cf1 & NB_N); padd(con1, dn.y, cf1 & NB_S);
                if (LAT == LAT_TRIANGULAR) {
                    padd(con0, c[PT_LD - 1], cf0 & NB_NW); padd(con0, up.y, cf0 & NB_NE);
                    padd(con1, dn.x, cf1 & NB_SW);         padd(con1, drt, cf1 & NB_SE);
                }
                const double2 t0 = (interior || e0 == interior_ex<LAT>(gx)) ? dtab[FtCfgA3::tabp(LAT, ft_pat<LAT>(cf0, 0), lane)] : pt_dsite(cf0, e0, gx);
                const double2 t1 = (interior || e1 == interior_ex<LAT>(gx + 1)) ? dtab[FtCfgA3::tabp(LAT, ft_pat<LAT>(cf1, 1), lane)] : pt_dsite(cf1, e1, gx + 1);
                double2 q;
                q.x = t0.x * cc.x - (prm.g0 * con0 + prm.gleak * (all0 - con0));
                q.y = t1.x * cc.y - (prm.g0 * con1 + prm.gleak * (all1 - con1));
                if (valid) {
                    const int64_t i = (int64_t)gy * g.m + gx;
                    double2 r = ld2(&s.sr[(ly + 1) * PT_LD + 2 + 2 * tx]);
                    r.x -= ak * q.x; r.y -= ak * q.y;
                    st2(vr + i, r);
                    if (keep_x || (DIST ? g.y0 : 0) + gy == 1 || (DIST ? g.y0 : 0) + gy == (DIST ? g.ng : g.n) - 2) {
                        double2 x = ld2(vx + i);
                        x.x += ak * cc.x; x.y += ak * cc.y;
                        st2(vx + i, x);
                    }
                    acc0 += r.x * r.x * t0.y + r.y * r.y * t1.y;
                    acc1 += r.x * r.x + r.y * r.y;
                }
                }
                dlf = lf; drt = rt;
                dn = cc; cc = up;
            }
            (void)dlf;
        }
        // generic-proxy accesses to this stage are ordered before the bulk copies that will refill it
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // per-tile partial sums (block_sum synchronises: every thread is done with this stage afterwards)
        if (MODE == 0) {
            double bs = block_sum(acc0, sh);
            if (tid == 0) partial[tl] = bs;
        } else {
            double a = block_sum(acc0, sh), b = block_sum(acc1, sh);
            if (tid == 0) { partial[tl * 2 + 0] = a; partial[tl * 2 + 1] = b; }
        }
    }
    if (MODE == 0) {
        if (last_block(&st->ticket_a)) {
            double tot = fold_partials(partial, ntiles, 1, 0, sh);
            if (threadIdx.x == 0) { st->akden = tot; st->ak = st->bknum / tot; }
        }
    } else {
        if (last_block(&st->ticket_b)) {
            double fa = fold_partials(partial, ntiles, 2, 0, sh);
            double fc = fold_partials(partial, ntiles, 2, 1, sh);
            if (threadIdx.x == 0 && DIST) { st->red[0] = fa; st->red[1] = fc; }      // summed over the ranks, then pcg_post
            if (threadIdx.x == 0 && !DIST) {
                int it = st->iter + 1;
                double err = sqrt(fc) / st->bnrm;
                st->iter = it; st->err = err; st->rr = fc;
                st->bkden = st->bknum; st->bknum = fa; st->bk = fa / st->bkden;
                if (!(err > st->tol) || it > st->itmax) st->done = 1;     // loop guard iter <= itmax (:780)
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// K6+K7 in one pass (pcg_fused_tile.cuh): the whole iteration in ONE persistent TMA-fed kernel with ONE
// grid-wide reduction.  Same pipeline as pcg_pipe_kernel (one CTA per SM, two shared-memory stages, one
// thread issues the three tensor copies of the next tile while the CTA computes the current one); the
// per-tile arithmetic lives in pcg_fused_tile.cuh, shared with the host emulation of the CPU tests.
//   HBM traffic per site and iteration: r 8 + s 8 + conduct byte 1 read, r 8 + s 8 written = 33 B.
// Single GPU, no periodic wrap, Gtop / Gbot only (x and p exist on rows 1 and n-2: xrow / prow, 2 m doubles
// each); everything else takes the two-kernel path.  Consecutive iterations sweep the tiles in opposite
// directions, so each starts on the part of r / s the previous one wrote last (still in L2).
// ------------------------------------------------------------------------------------------
static_assert(FT_KMAX <= 2 * FtCfgD::THREADS && FT_KMAX <= 2 * FtCfgD32::THREADS, "the coarse stage takes at most two blocks per thread");
static_assert(FtCfgA::SMEM <= 227 * 1024 && FtCfgA3::SMEM <= 227 * 1024 && FtCfgD::SMEM <= 227 * 1024 && FtCfgD32::SMEM <= 227 * 1024, "fused PCG tile does not fit the shared memory of one SM");

// block sums of three values at once (fixed order: lanes by shuffles, then warp 0 folds the per-warp partials by
// shuffles); result valid in thread 0
__device__ __forceinline__ void block_sum3(double& a, double& b, double& c, double* sh)
{
    for (int o = 16; o; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
        c += __shfl_down_sync(0xffffffffu, c, o);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (int)(blockDim.x >> 5);
    __syncthreads();
    if (lane == 0) { sh[w] = a; sh[32 + w] = b; sh[64 + w] = c; }
    __syncthreads();
    if (w == 0) {
        a = lane < nw ? sh[lane] : 0.0; b = lane < nw ? sh[32 + lane] : 0.0; c = lane < nw ? sh[64 + lane] : 0.0;
        for (int o = 16; o; o >>= 1) {
            a += __shfl_down_sync(0xffffffffu, a, o);
            b += __shfl_down_sync(0xffffffffu, b, o);
            c += __shfl_down_sync(0xffffffffu, c, o);
        }
    }
}

// device arrays of the deflated iteration (FtCfgD; the others use D only: the order in which the tiles are visited):
// mu[k] of the blocks (read by the sweep, rewritten by its coarse stage), Fb[FB_PLANES][FT_KMAX] crossing currents
// summed per block, Einv[k][k] dense inverse of E = Z^T A Z
struct FtDeflDev { FtDefl D; double* mu; double* Fb; const double* Einv; const int* sched; };
constexpr int FW_VALID = 256;           // tcoord.z: FtWalk::info of the tile | FW_VALID

template <int LAT, class C>
__global__ void __launch_bounds__(C::THREADS, C::CTAS)
pcg_fused_kernel(const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_s,
                 const __grid_constant__ CUtensorMap tm_cf, Geom g, PcgParams prm, double* __restrict__ r_out,
                 double* __restrict__ s_out, double* __restrict__ xrow, double* __restrict__ prow,
                 double* __restrict__ partial, PcgState* __restrict__ st, int rev, int prime, FtDeflDev dd,
                 const double* __restrict__ u_in, const double* __restrict__ s_in, const uint8_t* __restrict__ cfull)
{
    if (st->done) return;
    extern __shared__ __align__(128) unsigned char ft_raw[];
    double* su = reinterpret_cast<double*>(ft_raw + 2 * (size_t)C::STAGE_BYTES);                      // phase U (not with V = 3)
    FtDiag* dtab = reinterpret_cast<FtDiag*>(ft_raw + 2 * (size_t)C::STAGE_BYTES + C::U_BYTES);      // [64][DC]
    double* sh = reinterpret_cast<double*>(dtab + 64 * C::DC);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(sh + 96);                      // one mbarrier per stage
    double* cinv = reinterpret_cast<double*>(bars + 2);                                             // [64] 1/d by bond counts (boundary tiles)
    int4* tcoord = reinterpret_cast<int4*>(cinv + 64);                                              // [2] tile (ix, iy, info | FW_VALID, block) of the two stages
    FtWalk* wk_sm = reinterpret_cast<FtWalk*>(tcoord + 2);                                          // the producer's position on the lattice (48 bytes; in registers only while it moves)
    double* sft = reinterpret_cast<double*>(tcoord + 5);                                            // [2][SFT_N] shift tables (deflation)
    double* srec = sft + 2 * C::SFT_N;                                                              // [2][REC_N] ... and their per-row records
    double* sfl = srec + 2 * C::REC_N;                                                              // [256] row partials of the coarse product; scratch
    double* sru = sfl + 256;                                                                        // [THREADS] sum rho u' per main thread
    double* tsl = sru + C::THREADS;                                                                 // [8] slot totals of the tile just done
    double* bacc = tsl + 8;                                                                         // [8] running sums of the block being walked
    double* sf = bacc + 8;                                                                          // [FT_KMAX] mu of every block during the sweep, Z^T A u' in the coarse stage
    double* rtab = sf + FT_KMAX;                                                                    // rounding residue of the diagonal per table slot
    const int tid = threadIdx.x;
    const int G = (int)gridDim.x;
    const FtScalars sc{prm.g0, prm.gleak, prime ? 0.0 : st->ak, prime ? 0.0 : st->bk};
    auto stage_r = [&](int k) { return reinterpret_cast<double*>(ft_raw + (size_t)k * C::STAGE_BYTES); };
    auto stage_s = [&](int k) { return reinterpret_cast<double*>(ft_raw + (size_t)k * C::STAGE_BYTES + C::R_BYTES); };
    auto stage_cf = [&](int k) { return reinterpret_cast<uint8_t*>(ft_raw + (size_t)k * C::STAGE_BYTES + C::R_BYTES + C::S_BYTES); };
    // the producer thread (first lane of the first ring warp, which has time to spare: on the square lattice it has no ring column) walks the lattice (FtWalk: block by
    // block) one tile ahead of the CTA: it publishes the coordinates of the next tile and issues its three TMA tensor copies:
    // r (tile + 2-site halo), s (tile + east / north / west ring), conduct bytes
    constexpr int PROD = C::V == 1 ? 0 : C::RING_T0;
    constexpr int BSTEP = C::RING_NT == 32 ? C::RING_T0 + 1 : C::RING_T0 + 32;    // folds the slot totals of a finished tile into its block's sums                     // first lane of the second ring warp: folds the slot totals of a finished tile
    auto publish_issue = [&](const FtWalk& wk, int k) {
        int4 nt = make_int4(0, 0, 0, 0);
        if (wk.valid(dd.D)) nt = make_int4(wk.ix(dd.D), wk.iy(dd.D), wk.info(dd.D, rev) | FW_VALID, wk.B);
        tcoord[k] = nt;
        if (!(nt.z & FW_VALID)) return;
        const int x0 = nt.x * C::TX, y0 = nt.y * C::TY;
        mbar_arrive_expect(&bars[k], (unsigned)(C::RR * C::LD * 8 + C::SR * C::LD * 8 + C::RR * C::CLD));
        tma_box_g2s(stage_r(k), &tm_r, x0 - 2, y0 - 1, &bars[k]);
        tma_box_g2s(stage_s(k), &tm_s, x0 - 2, y0, &bars[k]);
        tma_box_g2s(stage_cf(k), &tm_cf, x0 - 16, y0 - 1, &bars[k]);
    };
    if (tid == PROD) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        FtWalk wk;
        wk.start(dd.D, (int)blockIdx.x, G, rev, dd.sched);
        publish_issue(wk, 0);
        *wk_sm = wk;
    }
    // shift table of tile (ix, iy) into buffer b, by ONE warp (lane l): the nine mu of the 3 x 3 tiles around the tile
    // first, then one lane per staged row: mu of the west / own / east block column (0 on rows that are not unknowns),
    // then one lane per compute row: its record (ft_defl_rec_entry: the same differences, formed here without the switch)
    auto fill_shift = [&](int b, int ix, int iy, int l) {
        double* m9 = sfl + 240;
        double* T = sft + b * C::SFT_N;
        double* R = srec + b * C::REC_N;
        if (l < 9) {
            int tx = ix + l % 3 - 1;
            const int ty = iy + l / 3 - 1;
            if (dd.D.pbc) tx = tx < 0 ? dd.D.ntx - 1 : (tx >= dd.D.ntx ? 0 : tx);        // periodic wrap: the tile column beyond the seam
            m9[l] = (tx >= 0 && tx < dd.D.ntx && ty >= 0 && ty < dd.D.nty) ? sf[ft_defl_block(dd.D, tx, ty)] : 0.0;
        }
        __syncwarp();
        for (int pr = l; pr < C::RR; pr += 32) {
            const int gy = iy * C::TY - 1 + pr;
            const bool unk = gy >= 1 && gy <= g.n - 2;
            const double* mrow = m9 + (pr == 0 ? 0 : (pr > C::TY ? 6 : 3));      // tile row below / own / above
            ft_st2(&T[pr * 4 + 0], unk ? mrow[0] : 0.0, unk ? mrow[1] : 0.0);
            ft_st2(&T[pr * 4 + 2], unk ? mrow[2] : 0.0, 0.0);
        }
        __syncwarp();
        for (int lr = l; lr < C::CR; lr += 32) {
            const double* a = T + lr * 4;                       // rows pr = lr (below), lr + 1 (own), lr + 2 (above)
            const double mc = a[5];
            ft_st2(&R[lr * 8 + 0], mc, mc - a[9]);              // mu, north
            ft_st2(&R[lr * 8 + 2], mc - a[1], 0.0);             // south, (zero)
            ft_st2(&R[lr * 8 + 4], mc - a[4], mc - a[6]);       // west, east
            ft_st2(&R[lr * 8 + 6], mc - a[8], mc - a[2]);       // north-west across the west border, south-east across the east border
        }
        __syncwarp();
    };
    // slot totals of a finished tile into the running sums of its block; the block's sums go out with its last tile
    auto block_step = [&](int info, int B) {
        ft_defl_block_add(info, tsl, bacc);
        if (info & FW_LAST) {
#pragma unroll
            for (int q = 0; q < FB_PLANES; ++q) { dd.Fb[q * FT_KMAX + B] = bacc[q]; bacc[q] = 0.0; }
        }
    };

    for (int k = tid; k < ft_tab_slots<LAT, C>(); k += C::THREADS) dtab[k] = ft_tab_slot<LAT, C>(g, k, prm.g0, prm.gleak);
    if (tid >= C::THREADS - 64) cinv[tid - (C::THREADS - 64)] = ft_cinv_entry(tid - (C::THREADS - 64), prm.g0, prm.gleak);
    if (C::DEFL) {
        for (int k = tid; k < ft_tab_slots<LAT, C>(); k += C::THREADS) rtab[k] = ft_rho_slot<LAT, C>(g, k, prm.g0, prm.gleak);
        // mu of every block (8 KB) stays in shared memory for the sweep
        for (int B = tid; B < dd.D.k; B += C::THREADS) sf[B] = dd.mu[B];
        if (tid < 16) tsl[tid] = 0.0;                                    // (tsl and bacc)
    }
    __syncthreads();
    if (C::DEFL && tid >= PROD && tid < PROD + 32) { const int4 t0 = tcoord[0]; if (t0.z & FW_VALID) fill_shift(0, t0.x, t0.y, tid - PROD); }
    __syncthreads();
    double rz = 0.0, rr = 0.0, en = 0.0;
    int pinfo = 0, pB = 0;                                       // the tile before this one
    for (int k = 0;; ++k) {
        const int4 tc = tcoord[k & 1];
        if (!(tc.z & FW_VALID)) break;
        const int ix = tc.x, iy = tc.y, x0 = ix * C::TX, y0 = iy * C::TY;
        if (tid == PROD) {                                       // the other stage was released by the barriers of the last tile
            FtWalk wk = *wk_sm;
            wk.next(dd.D, G, rev, dd.sched);
            publish_issue(wk, (k + 1) & 1);
            *wk_sm = wk;
        }
        if (C::DEFL && tid >= PROD && tid < PROD + 32) {
            // ... and its warp prepares the next tile's shift table while the others wait for this tile's data (the other
            // buffer: its readers finished before the last tile's barriers)
            __syncwarp();
            const int4 nt = tcoord[(k + 1) & 1];
            if (nt.z & FW_VALID) fill_shift((k + 1) & 1, nt.x, nt.y, tid - PROD);
        }
        if (C::DEFL && tid == BSTEP && (pinfo & FW_VALID)) block_step(pinfo, pB);
        mbar_wait(&bars[k & 1], (unsigned)((k >> 1) & 1));
        const double* sr = stage_r(k & 1);
        double* ss = stage_s(k & 1);
        const uint8_t* scf = stage_cf(k & 1);
        const double* sftk = C::DEFL ? sft + (k & 1) * C::SFT_N : nullptr;
        const double* sreck = C::DEFL ? srec + (k & 1) * C::REC_N : nullptr;
        const bool interior = ft_interior<C>(g, x0, y0);
        if (C::USTATE && g.pbc && (x0 == 0 || x0 + C::TX == g.m)) {
            // periodic wrap: the halo columns beyond the seam come straight from the input vectors (uniform over the CTA)
            if (tid >= C::RING_T0) ft_wrap_patch<C>(g, x0, y0, u_in, s_in, cfull, const_cast<double*>(sr), ss, const_cast<uint8_t*>(scf), tid - C::RING_T0, C::RING_NT);
            __syncthreads();
        }
        if (!C::USTATE) {
            ft_phase_u<LAT, C>(g, sr, scf, su, dtab, x0, y0, interior, tid, prm.g0, prm.gleak, cinv);
            __syncthreads();
        }
        const double* uu = C::USTATE ? sr : su;                 // V = 3: the staged vector IS u
        if (C::V == 1) { rz = 0.0; rr = 0.0; en = 0.0; }
        double ru = 0.0;                                        // sum rho u' over this thread's sites of the tile (deflation)
        // (the deflated main phase needs every register: the running bond-energy sum waits in shared memory meanwhile)
        if (C::DEFL && tid < C::MAIN_THREADS) sru[tid] = en;
        if (interior) {                                         // (uniform over the CTA: two instantiations of the tile phases)
            ft_phase_main<LAT, C, true>(g, sc, sr, ss, scf, uu, dtab, cinv, x0, y0, true, tid, r_out, s_out, xrow, prow, rz, rr, sreck, rtab, &ru);
            ft_phase_ringcols<LAT, C, true>(g, sc, sr, ss, scf, uu, dtab, cinv, x0, y0, tid, sftk, rtab);
        } else {
            ft_phase_main<LAT, C, false>(g, sc, sr, ss, scf, uu, dtab, cinv, x0, y0, false, tid, r_out, s_out, xrow, prow, rz, rr, sreck, rtab, &ru);
            ft_phase_ringcols<LAT, C, false>(g, sc, sr, ss, scf, uu, dtab, cinv, x0, y0, tid, sftk, rtab);
        }
        if (C::DEFL && tid < C::MAIN_THREADS) { en = sru[tid]; sru[tid] = ru; }
        __syncthreads();
        if (interior) ft_phase_energy<LAT, C, true>(g, sc, ss, scf, x0, y0, tid, en);
        else ft_phase_energy<LAT, C, false>(g, sc, ss, scf, x0, y0, tid, en);
        if (C::DEFL && tid >= C::RING_T0) {
            // meanwhile the two ring warps form the currents through the bonds that cross the tile's borders: every slot
            // has one producer warp (ft_flux_thread), lanes folded by a fixed shuffle tree
            const int rl = tid - C::RING_T0;
            double f[FS_SLOTS] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            ft_flux_thread<LAT, C>(g, sc, ss, scf, sru, x0, y0, rl, interior, true, f);
            const bool dirichlet = !interior && (y0 <= 1 || y0 + C::TY >= g.n - 2);
#pragma unroll
            for (int q = 0; q < FS_SLOTS; ++q) {
                const bool second = q == FS_N || q == FS_NW;                     // slots of the second ring warp
                if (C::RING_NT == 64 && second != (rl >= 32)) continue;
                if (LAT == LAT_SQUARE && (q == FS_W || q == FS_NW)) continue;    // (stay 0)
                double v = f[q];
                if (q != FS_D || dirichlet)
                    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                if ((rl & 31) == 0) tsl[q] = v;
            }
        }
        // generic-proxy accesses to this stage are ordered before the bulk copies that will refill it
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (C::V == 1) {
            // per-tile partial sums, folded in tile order: the result does not depend on the number of CTAs
            const int tl = iy * dd.D.ntx + ix;
            double a = rz, b = rr, c = en;
            block_sum3(a, b, c, sh);                            // synchronises: every thread is done with the stage
            if (tid == 0) { partial[tl * 3 + 0] = a; partial[tl * 3 + 1] = b; partial[tl * 3 + 2] = c; }
        } else {
            __syncthreads();                                    // every thread is done with the stage
        }
        pinfo = tc.z; pB = tc.w;
    }
    if (C::DEFL && tid == BSTEP && (pinfo & FW_VALID)) block_step(pinfo, pB);
    int nparts = dd.D.ntx * dd.D.nty;
    constexpr int NQ = C::DEFL ? 4 : 3;
    if (C::V >= 2) {
        // the sums stayed in registers over all tiles of this CTA: one reduction per CTA, folded in CTA order
        block_sum3(rz, rr, en, sh);
        if (tid == 0) { partial[blockIdx.x * NQ + 0] = rz; partial[blockIdx.x * NQ + 1] = rr; partial[blockIdx.x * NQ + 2] = en; }
        nparts = G;
    }
    if (C::DEFL) {
        // coarse stage (cooperative launch: every CTA is resident): when all tiles of the lattice are done -- ONE grid-wide
        // barrier -- every CTA forms Z^T A u' of all blocks from the block sums (five loads per block, 40 KB from L2), its
        // share of the rows of mu' = E^-1 (Z^T A u') -- E^-1 streams from L2 -- and its share of mu'.(Z^T A u'), which the
        // scalar recurrences subtract from u'.A u'
        __threadfence();
        cooperative_groups::grid_group grid = cooperative_groups::this_grid();
        grid.sync();
        const int kk = dd.D.k;
        const int lane = tid & 31, w = tid >> 5, nw = C::THREADS / 32;
        {   // (k <= FT_KMAX < 2 THREADS: at most two blocks per thread, their ten loads in flight together)
            // (every CTA reads the same 30 KB: each starts at another block so that they do not queue on the same L2 sectors)
            const int rot = (int)(((unsigned)blockIdx.x * 37u) % (unsigned)kk);
            int B0 = tid + rot, B1 = tid + C::THREADS + rot;
            const bool v0 = tid < kk, v1 = tid + C::THREADS < kk;
            B0 = B0 >= kk ? B0 - kk : B0; B1 = B1 >= kk ? B1 - kk : B1; B1 = B1 >= kk ? B1 - kk : B1;
            const double f0 = ft_defl_block_f<LAT>(dd.D, dd.Fb, v0 ? B0 : 0), f1 = ft_defl_block_f<LAT>(dd.D, dd.Fb, v1 ? B1 : 0);
            if (v0) sf[B0] = f0;
            if (v1) sf[B1] = f1;
        }
        __syncthreads();
        const int nrows = (kk - (int)blockIdx.x + G - 1) / G;               // rows blockIdx.x + r G < kk
        const int qlen = ((kk + 3) / 4 + 31) / 32 * 32;
        double* part = sfl;                                                  // [nrows][4] (nrows * 4 <= 240: checked on the host)
        // unit (row r of this CTA, quarter q of the row): 8 independent loads per lane, one L2 latency per unit; the four
        // quarter sums of a row are folded in fixed order afterwards
        for (int u = w; u < nrows * 4; u += nw) {
            const int r = u >> 2, q = u & 3, j = (int)blockIdx.x + r * G;
            const double* row = dd.Einv + (size_t)j * kk;
            double a = 0.0;
#pragma unroll 8
            for (int i = q * qlen + lane; i < (q + 1) * qlen && i < kk; i += 32) a += row[i] * sf[i];
            for (int o = 16; o; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
            if (lane == 0) part[u] = a;
        }
        __syncthreads();
        double mf = 0.0;
        if (tid < nrows) {
            const int j = (int)blockIdx.x + tid * G;
            const double a = (part[tid * 4] + part[tid * 4 + 1]) + (part[tid * 4 + 2] + part[tid * 4 + 3]);
            dd.mu[j] = a;
            mf = a * sf[j];
        }
        if (w == 0) {
            for (int o = 16; o; o >>= 1) mf += __shfl_down_sync(0xffffffffu, mf, o);
            if (nrows > 32) mf = nan("");                                    // (never: nrows * 4 <= 240 on the host)
            if (lane == 0) partial[blockIdx.x * NQ + 3] = mf;
        }
    }
    if (last_block(&st->ticket_a)) {
        const double fz = fold_partials(partial, nparts, NQ, 0, sh);
        const double fr = fold_partials(partial, nparts, NQ, 1, sh);
        double fe = fold_partials(partial, nparts, NQ, 2, sh);
        const double fm = C::DEFL ? fold_partials(partial, nparts, NQ, 3, sh) : 0.0;
        if (threadIdx.x == 0) {
            if (C::DEFL) fe -= fm;                              // p.A p = u.A u - mu.Z^T A u - beta gamma / alpha
            FtState f;
            f.gamma = st->bknum; f.alpha = st->ak; f.beta = st->bk; f.bnrm = st->bnrm; f.err = st->err; f.rr = st->rr;
            f.tol = st->tol; f.iter = st->iter; f.itmax = st->itmax; f.done = st->done;
            ft_scalar_step(f, fz, fr, fe, prime);
            st->bkden = st->bknum; st->akden = fe;
            st->bknum = f.gamma; st->ak = f.alpha; st->bk = f.beta; st->err = f.err; st->rr = f.rr;
            st->iter = f.iter; st->done = f.done;
        }
    }
}

// ---- deflation set-up kernels -----------------------------------------------------------------------------------
// sums of the weights of the bonds that cross each tile's borders (the entries of E = Z^T A Z), one CTA per tile
template <int LAT, class C>
__global__ void __launch_bounds__(64)
defl_weights_kernel(Geom g, PcgParams prm, const uint8_t* __restrict__ cfull, int ntx, double* __restrict__ W)
{
    __shared__ double sw[2][8];
    const int tl = blockIdx.x, x0 = (tl % ntx) * C::TX, y0 = (tl / ntx) * C::TY;
    const FtGlobalAcc a{cfull, g.m};
    double f[FS_SLOTS] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int q = threadIdx.x; q < FtFluxItems<C>::N; q += 64) ft_flux_item<LAT, C, true>(g, prm.g0, prm.gleak, a, x0, y0, q, f);
    // ... and the sum of the diagonal's rounding residues over the tile's unknown sites (the matrix is the exact-row-sum
    // Laplacian + diag(rho): E gets sum rho on its diagonal)
    for (int q = threadIdx.x; q < C::TX * C::TY; q += 64) {
        const int gx = x0 + q % C::TX, gy = y0 + q / C::TX;
        if (gx >= g.m || gy < 1 || gy > g.n - 2) continue;
        double rho;
        diag_seq_rho(g, a.cf(gx, gy), neighbour_bits(g, gx, gy), gx, prm.g0, prm.gleak, &rho);
        f[FS_R] += rho;
    }
#pragma unroll
    for (int q = 0; q < FS_SLOTS; ++q) {
        for (int o = 16; o; o >>= 1) f[q] += __shfl_down_sync(0xffffffffu, f[q], o);
        if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5][q] = f[q];
    }
    __syncthreads();
    if (threadIdx.x < FS_STRIDE) W[(size_t)tl * FS_STRIDE + threadIdx.x] = threadIdx.x < FS_SLOTS ? sw[0][threadIdx.x] + sw[1][threadIdx.x] : 0.0;
}

// nu = E^-1 Z^T b: b lives on row n-2 only, so Z^T b has entries only in one block row (one CTA, thread B <-> block B)
template <class C>
__global__ void __launch_bounds__(1024)
defl_nu_kernel(Geom g, PcgParams prm, FtDefl D, const uint8_t* __restrict__ cfull, const double* __restrict__ Einv,
               double* __restrict__ nu, double* __restrict__ mu)
{
    __shared__ double fb[FT_KMAX];
    const int by = ((g.n - 2) / C::TY) / D.bh;
    for (int B = threadIdx.x; B < D.k; B += blockDim.x) {
        double acc = 0.0;
        if (B / D.nbx == by) {
            const int bx = B % D.nbx;
            const int xa = bx * D.bw * C::TX, xb = (bx + 1) * D.bw * C::TX < g.m ? (bx + 1) * D.bw * C::TX : g.m;
            for (int x = xa; x < xb; ++x) {
                const unsigned ex = neighbour_bits(g, x, g.n - 2), cf = cfull[(int64_t)(g.n - 2) * g.m + x];
                acc += rhs_seq(g, cf, ex, x, prm.g0, prm.gleak, prm.Va);
            }
        }
        fb[B] = acc;
    }
    __syncthreads();
    for (int B = threadIdx.x; B < D.k; B += blockDim.x) {
        double a = 0.0;
        for (int j = by * D.nbx; j < (by + 1) * D.nbx; ++j) a += Einv[(size_t)B * D.k + j] * fb[j];
        nu[B] = a;
        mu[B] = 0.0;                                             // the priming sweep runs unshifted
    }
}

// x0 = Z nu on the read-out rows, u0 = D^-1 (b - A Z nu) on every unknown row (vr held b on row n-2, 0 elsewhere)
template <class C>
__global__ void __launch_bounds__(256)
defl_init_kernel(Geom g, PcgParams prm, FtDefl D, const uint8_t* __restrict__ cfull, const double* __restrict__ nu,
                 double* __restrict__ vr, double* __restrict__ xrow)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + g.m;
    if (i >= g.t - g.m) return;
    const int x = (int)(i % g.m), y = (int)(i / g.m);
    double b, ni;
    vr[i] = ft_defl_u0<C>(g, D, cfull[i], nu, x, y, prm.Va, prm.g0, prm.gleak, &b, &ni);
    if (y == 1) xrow[x] = ni;
    if (y == g.n - 2) xrow[g.m + x] = ni;
}

// slab mode: the scalar recurrences, run by one thread after the all-reduce of the rank-local sums
//   phase 0: after pcg_init (red = |D^-1 b|^2, b.z, b.b)   phase 1: after MODE 0 (akden)
//   phase 2: after MODE 1 (red = r.z, r.r)                  phase 3: after the read-out (Itop, Ibot are summed in place)
__global__ void pcg_post_kernel(PcgState* st, int phase)
{
    if (phase == 0) { st->bnrm = sqrt(st->red[0]); st->bknum = st->red[1]; st->rr = st->red[2]; return; }
    if (st->done) return;
    if (phase == 1) { st->ak = st->bknum / st->akden; return; }
    const double fa = st->red[0], fc = st->red[1];
    const int it = st->iter + 1;
    const double err = sqrt(fc) / st->bnrm;
    st->iter = it; st->err = err; st->rr = fc;
    st->bkden = st->bknum; st->bknum = fa; st->bk = fa / st->bkden;
    if (!(err > st->tol) || it > st->itmax) st->done = 1;             // loop guard iter <= itmax (:780)
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
            // lattice rows 0 and ng-1: held by the first / last rank of a decomposed lattice
            if (e == 0 ? g.y0 != 0 : g.y0 + g.n != g.ng) continue;
            int y = e == 0 ? 0 : g.n - 1;
            int64_t i = (int64_t)y * g.m + x;
            unsigned ex = neighbour_bits(g, x, y), cf = cfull[i];
            double vi = e == 0 ? 0.0 : prm.Va;
            double acc = diag_of(g, cf, ex, x, prm.g0, prm.gleak) * vi;
            int xl = x > 0 ? x - 1 : g.m - 1, xr = x + 1 < g.m ? x + 1 : 0;
            int64_t row = i - x;
#define VAL(j) (g.y0 + (j) / g.m == 0 ? 0.0 : (g.y0 + (j) / g.m == g.ng - 1 ? prm.Va : vx[j]))
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
// K7s: small lattices, one CTA per realization (BASELINE configs[0]: L = 100, 1000 realizations).  The whole
// Jacobi-PCG solve of one realization runs inside one CTA: p and r live in shared memory, q = A p in
// registers, the two reductions of an iteration are block reductions, nothing returns to the host or to
// another kernel until the solve is done; a batch is one launch with grid = realizations.  Same
// recurrences, matrix, stopping rule and read-out as above (Sq/bondc.f:465-595).
// ------------------------------------------------------------------------------------------
constexpr int SM_THREADS = 1024, SM_KMAX = 13;           // up to 13 sites per thread: t <= 13312 (L <= 115)

// block sum of two values at once; every thread gets the result (fixed order: warp shuffles, then warp 0
// folds the 32 per-warp partials with shuffles and broadcasts through shared memory)
__device__ __forceinline__ double2 sm_block_sum2(double a, double b, double* sh)
{
    for (int o = 16; o; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();                                          // the previous result has been read by everyone
    if (lane == 0) { sh[w] = a; sh[32 + w] = b; }
    __syncthreads();
    if (w == 0) {
        a = sh[lane]; b = sh[32 + lane];
        for (int o = 16; o; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
        if (lane == 0) { sh[64] = a; sh[65] = b; }
    }
    __syncthreads();
    return make_double2(sh[64], sh[65]);
}
__device__ __forceinline__ double sm_block_sum(double v, double* sh) { return sm_block_sum2(v, 0.0, sh).x; }

template <int LAT>
__global__ void __launch_bounds__(SM_THREADS, 1)
pcg_small_kernel(Geom g, PcgParams prm, const uint8_t* __restrict__ cfbatch, double tol, int itmax,
                 double* __restrict__ Gout, int* __restrict__ iters, double* __restrict__ errs)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int t = (int)g.t, m = g.m;
    double* sp = reinterpret_cast<double*>(sm_raw);
    double* sr = sp + t;
    double* sx = sr + t;                                     // x on rows 1 and n-2 only (the read-out rows)
    double* sh = sx + 2 * m;                                 // 2 x 32 partials + 2 results
    double2* tab = reinterpret_cast<double2*>(sh + 66);      // 64 x (d, 1/d)
    uint8_t* scf = reinterpret_cast<uint8_t*>(tab + 64);
    uint8_t* sidx = scf + t;                                 // per site: index into dtab (sites with all their neighbours), 255 = compute
    double* dtab = reinterpret_cast<double*>(sm_raw + (((size_t)(sidx + t - sm_raw)) + 15) / 16 * 16);     // [128] diagonal by conduct pattern
    const int tid = threadIdx.x;
    const uint8_t* cfg = cfbatch + (size_t)blockIdx.x * t;
    bool any = false;
    for (int i = tid; i < t; i += SM_THREADS) { uint8_t c = cfg[i]; scf[i] = c; any |= c != 0; }
    const int K = (t + SM_THREADS - 1) / SM_THREADS;
    // no spanning cluster in this realization (the conduct mask is empty): report G = 0
    if (!__syncthreads_or(any)) {
        if (tid == 0) { Gout[2 * blockIdx.x] = 0.0; Gout[2 * blockIdx.x + 1] = 0.0; iters[blockIdx.x] = -1; errs[blockIdx.x] = 0.0; }
        return;
    }
    // The diagonal is the reference's ordered row sum (diag_seq): for a site with all its neighbours away from the periodic
    // seam it depends only on which of them conduct -- a table of 16 (square) / 2 x 64 (triangular: m is even, so the parity
    // of x is the parity of the site index) entries, looked up through a byte per site; the few border sites are computed.
    if (tid < 128) dtab[tid] = ft_diag_entry<LAT>(g, tid < (LAT == LAT_SQUARE ? 16 : 128) ? tid : 0, prm.g0, prm.gleak).d;
    for (int i = tid; i < t; i += SM_THREADS) {
        const int x = i % m, y = i / m, par = LAT == LAT_TRIANGULAR ? (x & 1) : 0;
        const unsigned ex = neighbour_bits(g, x, y);
        const bool table = ex == ft_interior_ex<LAT>(par) && !(g.pbc && (x == 0 || x == m - 1));
        sidx[i] = table ? (uint8_t)ft_pat<LAT>(scf[i] & ex, par) : (uint8_t)255;
    }
    __syncthreads();
    // (ex, conducting bits, diagonal) of site i
    auto site = [&](int i, unsigned& ex, unsigned& cf, double& d) {
        const unsigned idx = sidx[i];
        if (idx != 255u) { ex = ft_interior_ex<LAT>(LAT == LAT_TRIANGULAR ? (i & 1) : 0); cf = scf[i] & ex; d = dtab[idx]; }
        else { const int x = i % m; ex = neighbour_bits(g, x, i / m); cf = scf[i] & ex; d = diag_seq(g, cf, ex, x, prm.g0, prm.gleak); }
    };
    // r = b (bonds from row n-2 into the top row at Va), p = 0; bnrm = |D^-1 b|, bknum = b.z
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int i = tid; i < t; i += SM_THREADS) {
        double b = 0.0;
        if (i >= t - 2 * m && i < t - m && i >= m) {
            unsigned ex, cf; double d;
            site(i, ex, cf, d);
            b = rhs_seq(g, cf, ex, i % m, prm.g0, prm.gleak, prm.Va);
            const double z = b / d;
            s0 += z * z; s1 += b * z; s2 += b * b;
        }
        sr[i] = b; sp[i] = 0.0;
    }
    const double2 init = sm_block_sum2(s0, s1, sh);
    const double bnrm = sqrt(init.x);
    double bknum = init.y, bkden = 1.0, bk = 0.0, err = 0.0;
    (void)s2;
    for (int i = tid; i < 2 * m; i += SM_THREADS) sx[i] = 0.0;
    int iter = 0;
    for (;;) {
        // p = r / d + bk p  (interior rows)
        for (int i = tid + m; i < t - m; i += SM_THREADS) {
            unsigned ex, cf; double d;
            site(i, ex, cf, d);
            sp[i] = sr[i] / d + bk * sp[i];
        }
        __syncthreads();
        // q = A p, p.q
        double q[SM_KMAX], dot = 0.0;
#pragma unroll
        for (int k = 0; k < SM_KMAX; ++k) {
            q[k] = 0.0;
            const int i = tid + k * SM_THREADS;
            if (k >= K || i < m || i >= t - m) continue;
            unsigned ex, cf; double d;
            site(i, ex, cf, d);
            int xl = i - 1, xr = i + 1;                        // west / east neighbour (wrapped at the seam)
            if (sidx[i] == 255u) { const int x = i % m; if (x == 0) xl = i + m - 1; if (x == m - 1) xr = i - (m - 1); }
            double all = 0.0, con = 0.0;
#define NBR(bit, j) if (ex & bit) { const double v = sp[j]; all += v; if (cf & bit) con += v; }
            NBR(NB_E, xr) NBR(NB_W, xl) NBR(NB_N, i + m) NBR(NB_S, i - m)
            if (LAT == LAT_TRIANGULAR) { NBR(NB_NW, xl + m) NBR(NB_NE, xr + m) NBR(NB_SW, xl - m) NBR(NB_SE, xr - m) }
#undef NBR
            const double pc = sp[i];
            q[k] = d * pc - (prm.g0 * con + prm.gleak * (all - con));
            dot += pc * q[k];
        }
        const double akden = sm_block_sum(dot, sh);
        const double ak = bknum / akden;
        // x += ak p (read-out rows), r -= ak q, r.z and r.r
        double rz = 0.0, rr = 0.0;
#pragma unroll
        for (int k = 0; k < SM_KMAX; ++k) {
            const int i = tid + k * SM_THREADS;
            if (k >= K || i < m || i >= t - m) continue;
            unsigned ex, cf; double d;
            site(i, ex, cf, d);
            const double r = sr[i] - ak * q[k];
            sr[i] = r;
            if (i < 2 * m) sx[i - m] += ak * sp[i];
            if (i >= t - 2 * m) sx[m + i - (t - 2 * m)] += ak * sp[i];     // (n = 3: row 1 is both; the read-out uses sx[x])
            rz += r * r / d;
            rr += r * r;
        }
        const double2 fs = sm_block_sum2(rz, rr, sh);
        const double fa = fs.x, fc = fs.y;
        ++iter;
        err = sqrt(fc) / bnrm;
        bkden = bknum; bknum = fa; bk = fa / bkden;
        if (!(err > tol) || iter > itmax) break;             // loop guard iter <= itmax (:780)
    }
    // read-out (Sq/bondc.f:554-592)
    __syncthreads();
    double stop = 0.0, sbot = 0.0;
    for (int x = tid; x < m; x += SM_THREADS) {
        for (int e = 0; e < 2; ++e) {
            const int y = e == 0 ? 0 : g.n - 1;
            const int i = y * m + x, row = i - x;
            const unsigned ex = neighbour_bits(g, x, y), cf = scf[i];
            const double vi = e == 0 ? 0.0 : prm.Va;
            double acc = diag_of(g, cf, ex, x, prm.g0, prm.gleak) * vi;
            const int xl = x > 0 ? x - 1 : m - 1, xr = x + 1 < m ? x + 1 : 0;
#define VAL(j) ((j) / m == 0 ? 0.0 : ((j) / m == g.n - 1 ? prm.Va : ((j) / m == 1 ? sx[(j) - m] : sx[m + (j) - (g.n - 2) * m])))
#define NB(bit, j) if (ex & bit) { double w = (cf & bit) ? prm.g0 : prm.gleak; if (fabs(w) >= prm.read_thresh) acc -= w * VAL(j); }
            NB(NB_E, row + xr) NB(NB_W, row + xl) NB(NB_N, i + m) NB(NB_S, i - m)
            NB(NB_NW, row + m + xl) NB(NB_NE, row + m + xr) NB(NB_SW, row - m + xl) NB(NB_SE, row - m + xr)
#undef NB
#undef VAL
            if (e == 0) sbot += acc; else stop += acc;
        }
    }
    const double2 cur = sm_block_sum2(stop, sbot, sh);
    const double Itop = cur.x, Ibot = cur.y;
    if (tid == 0) {
        Gout[2 * blockIdx.x] = Itop / prm.Va; Gout[2 * blockIdx.x + 1] = fabs(Ibot) / prm.Va;
        iters[blockIdx.x] = iter; errs[blockIdx.x] = err;
    }
}

static size_t small_smem_bytes(const Geom& g) { return sizeof(double) * (2 * (size_t)g.t + 2 * (size_t)g.m + 66 + 128) + sizeof(double2) * 64 + 2 * (size_t)g.t + 32; }

bool pcg_small_fits(const Geom& g) { return g.t <= (int64_t)SM_KMAX * SM_THREADS && small_smem_bytes(g) <= 227 * 1024 && g.n >= 3; }

// conduct byte map of the realization just labeled (default spanning cluster chosen on the device) -> slot of a batch
int pcg_small_stage(Ctx* c, uint8_t* cfbatch, int slot)
{
    const Geom& g = c->g;
    build_cfull_kernel<<<nblk64(g.t), 256, 0, c->stream>>>(g, c->kind, 0, c->mask, c->label, cfbatch + (size_t)slot * g.t, c->d_sum);
    c->launches++;
    return (int)cudaGetLastError();
}

int pcg_small_solve(Ctx* c, const uint8_t* cfbatch, int nreal, double Va, double g0, double gleak, double tol, int itmax,
                    double read_thresh, double* d_G, int* d_iters, double* d_errs)
{
    const Geom& g = c->g;
    PcgParams prm{g0, gleak, Va, read_thresh};
    const size_t smem = small_smem_bytes(g);
    if (g.lattice == LAT_SQUARE) {
        PERC_CUDA(cudaFuncSetAttribute(pcg_small_kernel<LAT_SQUARE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pcg_small_kernel<LAT_SQUARE><<<nreal, SM_THREADS, smem, c->stream>>>(g, prm, cfbatch, tol, itmax, d_G, d_iters, d_errs);
    } else {
        PERC_CUDA(cudaFuncSetAttribute(pcg_small_kernel<LAT_TRIANGULAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pcg_small_kernel<LAT_TRIANGULAR><<<nreal, SM_THREADS, smem, c->stream>>>(g, prm, cfbatch, tol, itmax, d_G, d_iters, d_errs);
    }
    c->launches++;
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------
static unsigned nblk(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

// 2-D TMA descriptor of a row-major (cols x rows) array; box = box_cols x PT_ROWS elements.  The driver entry
// point is resolved at run time, so the library still links against nothing but the CUDA runtime.
static int make_tensor_map(CUtensorMap* map, void* base, int elem_bytes, int cols, int rows, int box_cols, int box_rows = PT_ROWS)
{
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess) return (int)e;
        if (!fn || q != cudaDriverEntryPointSuccess) return (int)cudaErrorNotSupported;
        encode = (EncodeFn)fn;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * elem_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(map, elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 900 + (int)r;
}

// which iteration kernel(s) a solve uses (perc_set_solver): 0 = automatic: the one-pass kernel whenever it applies, deflated;
// 1 = always the two-kernel form (linbcg's own sequence of operations); 2 = the one-pass kernel without deflation (linbcg's
// iterates in exact arithmetic).  PERC_PCG_SOLVER=classic|fused|deflated sets the default of the process.
static int pcg_default_mode()
{
    const char* e = getenv("PERC_PCG_SOLVER");
    if (e && !strcmp(e, "classic")) return 1;
    if (e && !strcmp(e, "fused")) return 2;
    return 0;
}

bool pcg_fused_applies(const Ctx* c, int keep_x, int warm)
{
    const Geom& g = c->g;
    const int mode = c->pcg_mode >= 0 ? c->pcg_mode : pcg_default_mode();
    // (periodic wrap: the plain one-pass kernel only, and only when the seam falls on a tile border)
    return mode != 1 && !keep_x && !warm && c->nranks == 1 && (!g.pbc || g.m % FtCfgA3::TX == 0) && (g.m % 16) == 0 && g.n >= 4;
}

// V = 3 of the one-pass kernel keeps u = D^-1 r in HBM: turn the initial residual into u (the unknown rows only;
// everything else is and stays 0)
__global__ void __launch_bounds__(256)
pcg_scale_u_kernel(Geom g, PcgParams prm, const uint8_t* __restrict__ cfull, double* __restrict__ vr)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + g.m;
    if (i >= g.t - g.m) return;
    const double r = vr[i];
    if (r == 0.0) return;
    const int x = (int)(i % g.m), y = (int)(i / g.m);
    vr[i] = r / diag_of(g, cfull[i], neighbour_bits(g, x, y), x, prm.g0, prm.gleak);
}

// deflation set-up of a solve: E = Z^T A Z from the conduct bytes (weights kernel -> host: banded Cholesky, dense
// inverse -> device), nu = E^-1 Z^T b, x0 = Z nu, u0 = D^-1 (b - A Z nu)
template <class C>
static int pcg_defl_setup(Ctx* c, const PcgParams& prm, const FtDefl& D, int ntiles, double* xrow)
{
    const Geom& g = c->g;
    cudaStream_t s = c->stream;
    // Einv[k][k], mu[FT_KMAX], nu[FT_KMAX], Fb[FB_PLANES][FT_KMAX], W[ntiles][FS_STRIDE]
    const size_t need = sizeof(double) * ((size_t)D.k * D.k + (2 + FB_PLANES) * (size_t)FT_KMAX + (size_t)ntiles * FS_STRIDE);
    if (need > c->defl_bytes) {
        if (c->d_defl) cudaFree(c->d_defl);
        if (c->h_defl) cudaFreeHost(c->h_defl);
        c->d_defl = nullptr; c->h_defl = nullptr; c->defl_bytes = 0;
        PERC_CUDA(cudaMalloc(&c->d_defl, need));
        PERC_CUDA(cudaMallocHost(&c->h_defl, sizeof(double) * ((size_t)D.k * D.k + (size_t)ntiles * FS_STRIDE)));
        c->defl_bytes = need;
    }
    double* d_einv = c->d_defl;
    double* d_mu = d_einv + (size_t)D.k * D.k;
    double* d_nu = d_mu + FT_KMAX;
    double* d_W = d_nu + (1 + FB_PLANES) * (size_t)FT_KMAX;
    double* h_einv = c->h_defl;
    double* h_W = h_einv + (size_t)D.k * D.k;
    if (g.lattice == LAT_SQUARE) defl_weights_kernel<LAT_SQUARE, C><<<ntiles, 64, 0, s>>>(g, prm, c->cfull, D.ntx, d_W);
    else defl_weights_kernel<LAT_TRIANGULAR, C><<<ntiles, 64, 0, s>>>(g, prm, c->cfull, D.ntx, d_W);
    PERC_CUDA(cudaGetLastError());
    PERC_CUDA(cudaMemcpyAsync(h_W, d_W, sizeof(double) * (size_t)ntiles * FS_STRIDE, cudaMemcpyDeviceToHost, s));
    PERC_CUDA(cudaStreamSynchronize(s));
    unsigned hw = std::thread::hardware_concurrency();
    if (ft_defl_build_einv(D, h_W, h_einv, hw > 16 ? 16 : (hw ? (int)hw : 1))) return (int)cudaErrorUnknown;
    PERC_CUDA(cudaMemcpyAsync(d_einv, h_einv, sizeof(double) * (size_t)D.k * D.k, cudaMemcpyHostToDevice, s));
    defl_nu_kernel<C><<<1, 1024, 0, s>>>(g, prm, D, c->cfull, d_einv, d_nu, d_mu);
    defl_init_kernel<C><<<nblk64(g.t - 2 * (int64_t)g.m), 256, 0, s>>>(g, prm, D, c->cfull, d_nu, c->vr, xrow);
    c->launches += 3;
    return (int)cudaGetLastError();
}

// the iteration loop of the one-pass kernel; on entry (after pcg_init_kernel) vr = b, vp = vx = 0, the scalars of
// the solve are initialised.  Buffers: r ping-pongs between vr and vp, s between vp2 and vx; x / p of rows 1 and
// n-2 live in xprow and are copied into vx for the read-out at the end.
template <class C>
static int pcg_fused_loop_t(Ctx* c, const PcgParams& prm)
{
    const Geom& g = c->g;
    cudaStream_t s = c->stream;
    const int ntx = (g.m + C::TX - 1) / C::TX, ntiles = ntx * ((g.n + C::TY - 1) / C::TY);
    // the order in which the tiles are visited: block by block (FtWalk); without deflation every tile is a block
    FtDeflDev dd{};
    if (C::DEFL) {
        int bw = 0, bh = 0;
        if (const char* e = getenv("PERC_DEFL_BLOCK")) sscanf(e, "%d,%d", &bw, &bh);      // tiles per block (experiments)
        dd.D = ft_defl_make(g, C::TX, C::TY, FT_KMAX, bw, bh);
    } else dd.D = ft_defl_make(g, C::TX, C::TY, 0, 1, 1);
    const int grid = dd.D.k < C::CTAS * c->num_sms ? dd.D.k : C::CTAS * c->num_sms;
    if (!c->xprow) PERC_CUDA(cudaMalloc(&c->xprow, sizeof(double) * 4 * g.m));
    PERC_CUDA(cudaMemsetAsync(c->xprow, 0, sizeof(double) * 4 * g.m, s));
    PERC_CUDA(cudaMemsetAsync(c->vp2, 0, sizeof(double) * g.t, s));
    // (function attributes are per device: set on every solve, the call is cheap)
    PERC_CUDA(cudaFuncSetAttribute(pcg_fused_kernel<LAT_SQUARE, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    PERC_CUDA(cudaFuncSetAttribute(pcg_fused_kernel<LAT_TRIANGULAR, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    double* rbuf[2] = {c->vr, c->vp};
    double* sbuf[2] = {c->vp2, c->vx};
    CUtensorMap tm_r[2], tm_s[2], tm_cf;
    int rc;
    for (int k = 0; k < 2; ++k) {
        rc = make_tensor_map(&tm_r[k], rbuf[k], 8, g.m, g.n, C::LD, C::RR); if (rc) return rc;
        rc = make_tensor_map(&tm_s[k], sbuf[k], 8, g.m, g.n, C::LD, C::SR); if (rc) return rc;
    }
    rc = make_tensor_map(&tm_cf, c->cfull, 1, g.m, g.n, C::CLD, C::RR); if (rc) return rc;
    double* xrow = c->xprow; double* prow = c->xprow + 2 * (size_t)g.m;
    if (C::DEFL) {
        rc = pcg_defl_setup<C>(c, prm, dd.D, ntiles, xrow); if (rc) return rc;
        dd.Einv = c->d_defl; dd.mu = c->d_defl + (size_t)dd.D.k * dd.D.k; dd.Fb = dd.mu + 2 * FT_KMAX;
        if ((dd.D.k + grid - 1) / grid > 32) return (int)cudaErrorInvalidConfiguration;   // (k <= 1024 and grid >= 32 SMs)
        c->defl_k = dd.D.k;
        // which CTA walks which blocks (boundary tiles cost about twice an interior one); built once per lattice / block shape
        double bcost = 2.0;
        if (const char* e = getenv("PERC_DEFL_BCOST")) bcost = atof(e);                   // (experiments; 0: blocks bid, bid + G, ...)
        const int key[6] = {g.m, g.n, dd.D.bw, dd.D.bh, grid, (int)(bcost * 1000)};
        if (bcost > 0.0) {
            if (!c->d_sched || memcmp(key, c->sched_key, sizeof(key)) != 0) {
                std::vector<int> sch;
                ft_defl_schedule<C>(g, dd.D, grid, bcost, sch);
                if (c->d_sched) cudaFree(c->d_sched);
                c->d_sched = nullptr;
                PERC_CUDA(cudaMalloc(&c->d_sched, sizeof(int) * sch.size()));
                PERC_CUDA(cudaMemcpyAsync(c->d_sched, sch.data(), sizeof(int) * sch.size(), cudaMemcpyHostToDevice, s));
                PERC_CUDA(cudaStreamSynchronize(s));            // `sch` is pageable and goes out of scope
                memcpy(c->sched_key, key, sizeof(key));
            }
            dd.sched = c->d_sched;
        }
    }
    int cur = 0, pass = 0;
    Geom garg = g; PcgParams parg = prm;
    auto launch = [&](int prime) -> cudaError_t {
        int rev = pass & 1;
        double* ro = rbuf[cur ^ 1]; double* so = sbuf[cur ^ 1];
        const double* ui = rbuf[cur]; const double* si = sbuf[cur]; const uint8_t* cfp = c->cfull;
        void* args[] = {&tm_r[cur], &tm_s[cur], &tm_cf, &garg, &parg, &ro, &so, &xrow, &prow, &c->partial, &c->d_pcg,
                        &rev, &prime, &dd, &ui, &si, &cfp};
        const void* fn = g.lattice == LAT_SQUARE ? (const void*)pcg_fused_kernel<LAT_SQUARE, C> : (const void*)pcg_fused_kernel<LAT_TRIANGULAR, C>;
        // the deflated sweep ends with a grid-wide stage: cooperative launch (all CTAs resident: one per SM)
        cudaError_t e = C::DEFL ? cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(C::THREADS), args, C::SMEM, s)
                                : cudaLaunchKernel(fn, dim3(grid), dim3(C::THREADS), args, C::SMEM, s);
        cur ^= 1; ++pass;
        c->launches++;
        return e;
    };
    if (C::USTATE && !C::DEFL) {
        pcg_scale_u_kernel<<<nblk64(g.t - 2 * (int64_t)g.m), 256, 0, s>>>(g, prm, c->cfull, c->vr);
        c->launches++;
    }
    PERC_CUDA(launch(1));          // s = A u0 and delta0: alpha0, beta0 = 0
    // average duration of a launch: CUDA events around whole chunks of launches (back to back on the stream, so the
    // bracket is the kernels' own device time incl. the gaps between them) / launches in them.  Chunks in which the
    // solve ended are left out (their trailing launches return at once) unless there is no other.
    float it_ms = 0.f; long nlaunch = 0;
    int chunk = 32, iters_before = 0;
    for (;;) {
        PERC_CUDA(cudaEventRecord(c->ev[8], s));
        for (int k = 0; k < chunk; ++k) PERC_CUDA(launch(0));
        PERC_CUDA(cudaEventRecord(c->ev[9], s));
        PERC_CUDA(cudaMemcpyAsync(c->h_pcg, c->d_pcg, sizeof(PcgState), cudaMemcpyDeviceToHost, s));
        PERC_CUDA(cudaStreamSynchronize(s));
        const int live = c->h_pcg->iter - iters_before;
        if (!c->h_pcg->done || nlaunch == 0) {
            float a = 0.f;
            cudaEventElapsedTime(&a, c->ev[8], c->ev[9]);
            it_ms += a; nlaunch += c->h_pcg->done ? (live > 0 ? live : 1) : chunk;
        }
        // a chunk that did not advance the iteration count means the launches did not run: never spin on that
        if (!c->h_pcg->done && live == 0) return (int)cudaErrorLaunchFailure;
        iters_before = c->h_pcg->iter;
        if (c->h_pcg->done) break;
        if (chunk < 512) chunk *= 2;
    }
    // the read-out consumes x on rows 1 and n-2 (vx was an s buffer: dead now)
    PERC_CUDA(cudaMemcpyAsync(c->vx + g.m, xrow, sizeof(double) * g.m, cudaMemcpyDeviceToDevice, s));
    PERC_CUDA(cudaMemcpyAsync(c->vx + (int64_t)(g.n - 2) * g.m, xrow + g.m, sizeof(double) * g.m, cudaMemcpyDeviceToDevice, s));
    c->phase_ms[6] = nlaunch ? it_ms / (float)nlaunch : 0.f;
    c->phase_ms[7] = 0.f;
    return 0;
}

// variant of the one-pass kernel: the deflated iteration FtCfgD by default; perc_set_solver(h, 2) / 12: FtCfgA3 (the same
// sweep without deflation), 10: FtCfgA (the first version: per-tile partial sums folded in tile order)
static int pcg_fused_loop(Ctx* c, const PcgParams& prm)
{
    const int mode = c->pcg_mode >= 0 ? c->pcg_mode : pcg_default_mode();
    int use = c->fused_cfg >= 0 ? c->fused_cfg : (mode == 2 ? 2 : 4);
    if (const char* e = getenv("PERC_FUSED_CFG")) if (c->fused_cfg < 0 && (*e == '1' || *e == '3' || *e == '5')) use = *e - '1';
    if (c->g.pbc && use != 4) use = 2;    // the first variant (r as the state vector) has no periodic wrap
    c->last_fused_cfg = use;
    if (use == 4) {
        if (const char* e = getenv("PERC_FUSED_TILE32")) if (*e == '1') return pcg_fused_loop_t<FtCfgD32>(c, prm);    // (the 32-row tile shape: timing comparisons)
        return pcg_fused_loop_t<FtCfgD>(c, prm);
    }
    if (use == 2) return pcg_fused_loop_t<FtCfgA3>(c, prm);
    return pcg_fused_loop_t<FtCfgA>(c, prm);
}

int pcg_solve(Ctx* c, int cluster_id, double Va, double g0, double gleak, double tol, int itmax,
              double read_thresh, int keep_x, double* Gtop, double* Gbot, int* iter, double* err, int warm)
{
    const Geom& g = c->g;
    // warm start: the voltages of the handle's previous solve are the initial guess (they must exist)
    if (warm && !(c->have_x && c->vx && c->nranks == 1)) warm = 0;
    if (warm) keep_x = 1;
    c->solved = false;                   // set again only when this solve succeeds (an error must not expose old voltages)
    cudaStream_t s = c->stream;
    PcgParams prm{g0, gleak, Va, read_thresh};
    dim3 sgrid((g.m + SP_TX - 1) / SP_TX, (g.n + SP_TY - 1) / SP_TY);
    const int ntx = (g.m + PT_TX - 1) / PT_TX, ntiles = ntx * ((g.n + PT_TY - 1) / PT_TY);
    const int pgrid = ntiles < c->num_sms ? ntiles : c->num_sms;     // one persistent CTA per SM
    const bool vec = (g.m % 16) == 0;    // staged kernels: rows of the fp64 vectors and of the byte mask are 16-byte aligned
    const int dist = c->nranks > 1;      // slab of a decomposed lattice: sums all-reduced, halo rows exchanged
    if (dist && !vec) return -1;         // the scalar fallback is single-GPU only
    int ugrid = 148 * 8;
    int64_t interior = g.t - 2 * (int64_t)g.m;
    if (ugrid > (interior + UP_THREADS - 1) / UP_THREADS) ugrid = (int)((interior + UP_THREADS - 1) / UP_THREADS);
    if (ugrid < 1) ugrid = 1;
    int need = (int)(sgrid.x * sgrid.y);
    if (need < ugrid * 3) need = ugrid * 3;
    if (need < 2 * ntiles) need = 2 * ntiles;
    { const int nf = 3 * ((g.m + FtCfgA::TX - 1) / FtCfgA::TX) * ((g.n + FtCfgA::TY - 1) / FtCfgA::TY); if (need < nf) need = nf; }
    const int want_x = keep_x;
    if (!vec) keep_x = 1;                // the scalar fallback always forms x
    if (need > c->partial_cap) {
        if (c->partial) cudaFree(c->partial);
        PERC_CUDA(cudaMalloc(&c->partial, sizeof(double) * need));
        c->partial_cap = need;
    }
    int rc = 0;
    if (vec && !(c->pcg_attr & 1u)) {      // per-device attribute, tracked per handle
        PERC_CUDA(cudaFuncSetAttribute(pcg_pipe_kernel<LAT_SQUARE, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PT_SMEM));
        PERC_CUDA(cudaFuncSetAttribute(pcg_pipe_kernel<LAT_SQUARE, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PT_SMEM));
        PERC_CUDA(cudaFuncSetAttribute(pcg_pipe_kernel<LAT_SQUARE, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PT_SMEM));
        PERC_CUDA(cudaFuncSetAttribute(pcg_pipe_kernel<LAT_SQUARE, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PT_SMEM));
        PERC_CUDA(cudaFuncSetAttribute(pcg_pipe_kernel<LAT_TRIANGULAR, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PT_SMEM));
        PERC_CUDA(cudaFuncSetAttribute(pcg_pipe_kernel<LAT_TRIANGULAR, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PT_SMEM));
        PERC_CUDA(cudaFuncSetAttribute(pcg_pipe_kernel<LAT_TRIANGULAR, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PT_SMEM));
        PERC_CUDA(cudaFuncSetAttribute(pcg_pipe_kernel<LAT_TRIANGULAR, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PT_SMEM));
        c->pcg_attr |= 1u;
    }
    PERC_CUDA(cudaMemsetAsync(c->d_pcg, 0, sizeof(PcgState), s));
    PERC_CUDA(cudaEventRecord(c->ev[6], s));
    build_cfull_kernel<<<nblk(g.t), 256, 0, s>>>(g, c->kind, cluster_id, c->mask, c->label, c->cfull, nullptr);
    if (dist) { rc = slab_halo_exchange(c, c->cfull, 1); if (rc) return rc; }           // conduct bytes of the halo rows
    pcg_init_kernel<<<ugrid, UP_THREADS, 0, s>>>(g, prm, c->cfull, c->vx, c->vr, c->vp, c->vq, c->partial, c->d_pcg, tol, itmax, dist, warm);
    c->launches += 2;
    PERC_CUDA(cudaMemsetAsync(c->vp2, 0, sizeof(double) * g.t, s));     // rows the kernels never write (Dirichlet) must read 0
    if (dist) {
        rc = slab_allreduce_f64(c, c->d_pcg->red, 3); if (rc) return rc;
        pcg_post_kernel<<<1, 1, 0, s>>>(c->d_pcg, 0);
        rc = slab_halo_exchange(c, c->vr, 8); if (rc) return rc;                        // r = b on the halo rows
        c->launches++;
    }
    // per-bond conductances (perc_set_bond_conductance): their own kernels (pcg_weighted.cu), voltages always formed
    const bool weighted = c->have_bond_w;
    if (weighted && (dist || warm)) return -1;
    const bool fused = !weighted && pcg_fused_applies(c, keep_x, warm);
    c->last_fused = fused;
    // TMA descriptors of the arrays the pipeline stages (row-major m x n, boxes of 34 rows)
    CUtensorMap tm_r{}, tm_pa{}, tm_pb{}, tm_cf{};
    if (vec) {
        rc = make_tensor_map(&tm_r, c->vr, 8, g.m, g.n, PT_LD); if (rc) return rc;
        rc = make_tensor_map(&tm_pa, c->vp, 8, g.m, g.n, PT_LD); if (rc) return rc;
        rc = make_tensor_map(&tm_pb, c->vp2, 8, g.m, g.n, PT_LD); if (rc) return rc;
        rc = make_tensor_map(&tm_cf, c->cfull, 1, g.m, g.n, PT_CLD); if (rc) return rc;
    }
#define PIPE_ARGS(pin, pout) ((pin) == c->vp ? tm_pa : tm_pb, tm_r, tm_cf, g, prm, c->cfull, c->vr, pin, pout, c->vx, c->partial, c->d_pcg, keep_x, ntx, ntiles)
#define PIPE_LAUNCH(MODE, pin, pout)                                                                             \
    do {                                                                                                         \
        if (g.lattice == LAT_SQUARE) {                                                                           \
            if (dist) pcg_pipe_kernel<LAT_SQUARE, MODE, 1><<<pgrid, PT_THREADS, PT_SMEM, s>>> PIPE_ARGS(pin, pout); \
            else pcg_pipe_kernel<LAT_SQUARE, MODE, 0><<<pgrid, PT_THREADS, PT_SMEM, s>>> PIPE_ARGS(pin, pout);      \
        } else {                                                                                                 \
            if (dist) pcg_pipe_kernel<LAT_TRIANGULAR, MODE, 1><<<pgrid, PT_THREADS, PT_SMEM, s>>> PIPE_ARGS(pin, pout); \
            else pcg_pipe_kernel<LAT_TRIANGULAR, MODE, 0><<<pgrid, PT_THREADS, PT_SMEM, s>>> PIPE_ARGS(pin, pout);  \
        }                                                                                                        \
    } while (0)
    double* pold = c->vp; double* pnew = c->vp2;
    float sp_ms = 0.f, up_ms = 0.f; int nsamp = 0;
    int chunk = 16, iters_before = 0;
    if (weighted) {
        rc = pcg_solve_weighted(c, Va, gleak, tol, itmax, read_thresh);
        if (rc) return rc;
    } else if (fused) {
        rc = pcg_fused_loop(c, prm);
        if (rc) return rc;
    } else
    for (;;) {
        for (int k = 0; k < chunk; ++k) {
            // bracket the two kernels of one mid-chunk iteration with events (the pipeline is full
            // there, so the bracket is the kernels' own device time) -> average per-kernel time
            bool sample = (k == chunk / 2);
            if (sample) PERC_CUDA(cudaEventRecord(c->ev[8], s));
            if (vec) {
                PIPE_LAUNCH(0, pold, pnew);
            } else if (g.lattice == LAT_SQUARE)
                pcg_spmv_kernel<LAT_SQUARE><<<sgrid, SP_THREADS, 0, s>>>(g, prm, c->cfull, c->vr, pold, pnew, c->vq, c->partial, c->d_pcg);
            else
                pcg_spmv_kernel<LAT_TRIANGULAR><<<sgrid, SP_THREADS, 0, s>>>(g, prm, c->cfull, c->vr, pold, pnew, c->vq, c->partial, c->d_pcg);
            if (dist) {
                rc = slab_allreduce_f64(c, &c->d_pcg->akden, 1); if (rc) return rc;      // p.Ap over all slabs
                pcg_post_kernel<<<1, 1, 0, s>>>(c->d_pcg, 1);
                c->launches++;
            }
            if (sample) PERC_CUDA(cudaEventRecord(c->ev[9], s));
            if (vec) {
                PIPE_LAUNCH(1, pnew, nullptr);
            } else
                pcg_update_kernel<<<ugrid, UP_THREADS, 0, s>>>(g, prm, c->cfull, c->vx, c->vr, pnew, c->vq, c->partial, c->d_pcg);
            if (dist) {
                rc = slab_allreduce_f64(c, c->d_pcg->red, 2); if (rc) return rc;         // r.z and r.r over all slabs
                pcg_post_kernel<<<1, 1, 0, s>>>(c->d_pcg, 2);
                rc = slab_halo_exchange(c, c->vr, 8); if (rc) return rc;                 // boundary rows of the new residual
                c->launches++;
            }
            { double* tmp = pold; pold = pnew; pnew = tmp; }
            if (sample) PERC_CUDA(cudaEventRecord(c->ev[10], s));
            c->launches += 2;
            PERC_CUDA(cudaGetLastError());               // a failed launch must not leave the host polling for `done`
        }
        PERC_CUDA(cudaMemcpyAsync(c->h_pcg, c->d_pcg, sizeof(PcgState), cudaMemcpyDeviceToHost, s));
        PERC_CUDA(cudaStreamSynchronize(s));
        if (!c->h_pcg->done && c->h_pcg->iter == iters_before) return (int)cudaErrorLaunchFailure;
        // a sample is valid if the solve was still live when it ran (always true for a chunk that
        // ended not-done; for the final chunk only if it finished after the sampled iteration)
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
    if (!weighted) {
        pcg_readout_kernel<<<1, 256, 0, s>>>(g, prm, c->cfull, c->vx, c->d_pcg);
        c->launches++;
    }
    if (dist) { rc = slab_allreduce_f64(c, &c->d_pcg->Itop, 2); if (rc) return rc; }     // Itop from the last rank, Ibot from the first
    PERC_CUDA(cudaEventRecord(c->ev[7], s));
    PERC_CUDA(cudaMemcpyAsync(c->h_pcg, c->d_pcg, sizeof(PcgState), cudaMemcpyDeviceToHost, s));
    PERC_CUDA(cudaStreamSynchronize(s));
    PERC_CUDA(cudaGetLastError());
    cudaEventElapsedTime(&c->phase_ms[5], c->ev[6], c->ev[7]);
    if (!fused && !weighted) {
        c->phase_ms[6] = nsamp ? sp_ms / nsamp : 0.f;
        c->phase_ms[7] = nsamp ? up_ms / nsamp : 0.f;
    }
    *Gtop = c->h_pcg->Itop / Va;
    *Gbot = fabs(c->h_pcg->Ibot) / Va;
    *iter = c->h_pcg->iter;
    *err = c->h_pcg->err;
    c->solved = true;
    c->have_x = want_x != 0 || weighted;
    return 0;
#undef PIPE_LAUNCH
#undef PIPE_ARGS
}

}  // namespace perc
