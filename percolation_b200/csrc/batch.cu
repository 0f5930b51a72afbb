// batch.cu -- many independent realizations on one handle without leaving the device.
//
// The reference's trial loops (`do ii = 1, numtrials`, Sq/site_perc.f:87, Sq/bond_cond.f:123,
// Sq/sb_perc.f:104) run one realization after the other and write one text row each.  Here every
// realization i = 0 .. nreal-1 draws its occupancy from the generator (stream = stream0 + i, exact
// fill counts ks / kb), is labeled, and is folded into statistics that stay on the device: the
// cluster-size histogram n_s and a block of integer sums.  No host synchronisation happens inside
// the loop (the exact-count selection runs device-resident, occupancy.cu); one download at the end.
// Integer sums are order independent, so shards of a batch on different GPUs all-reduce to the same
// result for any number of GPUs (SURVEY 8e mode 1).
#include <vector>
#include "context.h"

namespace perc {

// stats: [0] realizations  [1] sum ncl  [2] sum maxcs  [3] realizations with a spanning cluster
//        [4] sum nspan  [5] sum size of the default spanning cluster  [6] failed selections
//        [7] sum maxcs^2  [8] sum occupied sites  [9] sum occupied bonds
// A realization whose device-side exact-count selection failed (occupancy.cu: select_pick_kernel, counted in stats[6]) was
// labeled as an empty lattice: it is NOT folded into the statistics; the batch then returns PERC_E_SELECT.
__global__ void batch_accum_kernel(const Summary* __restrict__ sum, unsigned long long* __restrict__ hist, int nbins,
                                   long long* __restrict__ stats, const PhiloxThreshold* __restrict__ thr, int use_s, int use_b)
{
    if ((use_s && thr[0].failed) || (use_b && thr[1].failed)) return;
    const unsigned long long mp = sum->maxpack;
    long long ms = (long long)(mp >> 32);
    if (ms == 0 && sum->nlone > 0) ms = 1;
    stats[0] += 1;
    stats[1] += (long long)(sum->ncl + sum->nlone);
    stats[2] += ms;
    stats[7] += ms * ms;
    stats[8] += (long long)sum->nocc_sites;
    stats[9] += (long long)sum->nocc_bonds;
    int ns = sum->nspan < MAX_SPAN ? sum->nspan : MAX_SPAN;
    if (ns > 0) {
        stats[3] += 1;
        stats[4] += sum->nspan;
        stats[5] += (long long)(sum->span_best & 0xffffffffull);       // default choice: smallest canonical id
    }
    if (nbins > 0 && sum->nlone) hist[0] += sum->nlone;  // lone bonds are size-1 clusters
}

__global__ void __launch_bounds__(256)
batch_hist_kernel(int64_t t, const int32_t* __restrict__ label, const int32_t* __restrict__ size, int nbins,
                  unsigned long long* __restrict__ hist)
{
    extern __shared__ unsigned sh_hist[];
    int nsh = nbins < 4096 ? nbins : 4096;
    for (int k = threadIdx.x; k < nsh; k += blockDim.x) sh_hist[k] = 0;
    __syncthreads();
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t; i += stride) {
        if (label[i] != (int32_t)(i + 1)) continue;          // sizes live at the roots
        int32_t s = size[i];
        if (s > 0) {
            int b = s < nbins ? s - 1 : nbins - 1;
            if (b < nsh) atomicAdd(&sh_hist[b], 1u);
            else atomicAdd(&hist[b], 1ull);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nsh; k += blockDim.x)
        if (sh_hist[k]) atomicAdd(&hist[k], (unsigned long long)sh_hist[k]);
}

constexpr int BATCH_NSTATS = 16;
constexpr int BATCH_WAYS = 4;

// Small lattices do not fill the GPU (L = 1024 is 128 tiles for 148 SMs, L = 100 is one): a batch round-robins
// its realizations over up to BATCH_WAYS contexts, each with its own arrays and stream, so that the kernels
// of different realizations overlap.  ways[0] is the handle itself; the children are created once and kept.
static int batch_contexts(Ctx* c, int nreal, std::vector<Ctx*>& ways)
{
    ways.assign(1, c);
    int want = c->g.t <= ((int64_t)1 << 22) ? BATCH_WAYS : 1;
    if (want > nreal) want = nreal;
    while ((int)c->batch_kids.size() < want - 1) {
        Ctx* k = new Ctx();
        k->g = c->g; k->device = c->device;
        int rc = ctx_alloc(k);
        if (rc) { ctx_free(k); delete k; return rc; }
        c->batch_kids.push_back(k);
    }
    for (int j = 0; j < want - 1; ++j) ways.push_back(c->batch_kids[j]);
    return 0;
}

int batch_run(Ctx* c, int kind, int nreal, unsigned long long seed, unsigned long long stream0, int64_t ks, int64_t kb,
              int nbins, int64_t* hist, int64_t* stats)
{
    if (c->nranks > 1) return -4;
    std::vector<Ctx*> ways;
    int rc = batch_contexts(c, nreal, ways);
    if (rc) return rc;
    const int K = (int)ways.size();
    const int nb1 = nbins > 0 ? nbins : 1;
    const size_t hb = sizeof(unsigned long long) * (size_t)nb1;
    // per way: [nb1] histogram words, then BATCH_NSTATS statistics words
    unsigned long long* d_acc = nullptr;
    const size_t way_words = (size_t)nb1 + BATCH_NSTATS;
    PERC_CUDA(cudaMalloc(&d_acc, sizeof(unsigned long long) * way_words * K));
    {
        cudaError_t e = cudaMemset(d_acc, 0, sizeof(unsigned long long) * way_words * K);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();   // the ways' streams do not synchronise with the legacy stream
        if (e != cudaSuccess) { cudaFree(d_acc); return (int)e; }
    }
    const int nsh = nbins < 4096 ? nbins : 4096;
    for (Ctx* w : ways) w->batch_thr = true;
    for (int i = 0; i < nreal && rc == 0; ++i) {
        Ctx* w = ways[i % K];
        unsigned long long* d_hist = d_acc + way_words * (i % K);
        long long* d_stats = (long long*)(d_hist + nb1);
        rc = occ_generate_dev(w, seed, stream0 + (unsigned long long)i, kind != KIND_BOND ? ks : -1, kind != KIND_SITE ? kb : -1,
                              (unsigned long long*)(d_stats + 6));
        if (rc) break;
        rc = ccl_launch(w, kind);
        if (rc) break;
        if (nbins > 0) {
            batch_hist_kernel<<<148 * 4, 256, sizeof(unsigned) * nsh, w->stream>>>(w->g.t, w->label, w->size, nbins, d_hist);
            w->launches++;
        }
        batch_accum_kernel<<<1, 1, 0, w->stream>>>(w->d_sum, d_hist, nbins, d_stats, w->d_thr, kind != KIND_BOND, kind != KIND_SITE);
        w->launches++;
    }
    for (Ctx* w : ways) {
        w->batch_thr = false;
        w->labeled = false;          // the per-realization state on the host was never fetched
        w->site_src = SRC_NONE; w->bond_src = SRC_NONE;     // the batch's thresholds lived on the device only
        cudaError_t e = cudaStreamSynchronize(w->stream);
        if (!rc && e != cudaSuccess) rc = (int)e;
        if (w != c) { c->launches += w->launches; w->launches = 0; }
    }
    if (rc == 0) {
        std::vector<unsigned long long> h(way_words * K);
        {
            cudaError_t e = cudaMemcpy(h.data(), d_acc, sizeof(unsigned long long) * way_words * K, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { cudaFree(d_acc); return (int)e; }
        }
        for (int k = 0; k < BATCH_NSTATS; ++k) stats[k] = 0;
        if (nbins > 0 && hist) for (int b = 0; b < nbins; ++b) hist[b] = 0;
        for (int j = 0; j < K; ++j) {
            const unsigned long long* hw = h.data() + way_words * j;
            if (nbins > 0 && hist) for (int b = 0; b < nbins; ++b) hist[b] += (int64_t)hw[b];
            for (int k = 0; k < BATCH_NSTATS; ++k) stats[k] += (int64_t)hw[nb1 + k];
        }
        rc = (int)cudaGetLastError();
    }
    (void)hb;
    cudaFree(d_acc);
    return rc;
}

// batch of realizations WITH the Kirchhoff conductance of each one's default spanning cluster (the trial
// loop of Sq/bond_cond.f:123-498 at one fill).  Small lattices (t <= 13312): label every realization into
// its slot of a conduct-byte batch, then ONE launch solves them all, one CTA each (pcg_small_kernel).
// Larger lattices: realization by realization through the pipelined solver.
// G(2, nreal) = Gtop, Gbot (0, 0 and iters = -1 for a realization without a spanning cluster).
int batch_conduct_run(Ctx* c, int kind, int nreal, unsigned long long seed, unsigned long long stream0, int64_t ks, int64_t kb,
                      double Va, double g0, double gleak, double tol, int itmax, double read_thresh,
                      double* G, int32_t* iters, int64_t* stats)
{
    if (c->nranks > 1) return -4;
    cudaStream_t st = c->stream;
    const Geom& g = c->g;
    int rc = 0;
    for (int k = 0; k < BATCH_NSTATS; ++k) stats[k] = 0;
    if (!pcg_small_fits(g)) {
        for (int i = 0; i < nreal; ++i) {
            rc = occ_generate(c, seed, stream0 + (unsigned long long)i, kind != KIND_BOND ? ks : -1, kind != KIND_SITE ? kb : -1);
            if (rc) return rc;
            rc = ccl_run(c, kind);
            if (rc) return rc;
            stats[0] += 1;
            G[2 * i] = G[2 * i + 1] = 0.0; iters[i] = -1;
            if (c->h_span_ids.empty()) continue;
            stats[3] += 1;
            if (!c->vx) return -4;                        // the caller (abi) allocates the solver vectors first
            double err = 0.0; int it = 0;
            rc = pcg_solve(c, c->h_span_ids[0], Va, g0, gleak, tol, itmax, read_thresh, 0, &G[2 * i], &G[2 * i + 1], &it, &err);
            if (rc) return rc;
            iters[i] = it;
        }
        return 0;
    }
    uint8_t* d_cf = nullptr; double* d_G = nullptr; double* d_err = nullptr; int* d_it = nullptr; long long* d_stats = nullptr;
    long long* d_wstats = nullptr;
    struct Scratch {                                   // freed on every path out of this function
        void** p[6];
        ~Scratch() { for (void** q : p) if (*q) cudaFree(*q); }
    } scratch{{(void**)&d_cf, (void**)&d_G, (void**)&d_err, (void**)&d_it, (void**)&d_stats, (void**)&d_wstats}};
    PERC_CUDA(cudaMalloc(&d_cf, (size_t)nreal * g.t));
    PERC_CUDA(cudaMalloc(&d_G, sizeof(double) * 2 * nreal));
    PERC_CUDA(cudaMalloc(&d_err, sizeof(double) * nreal));
    PERC_CUDA(cudaMalloc(&d_it, sizeof(int) * nreal));
    PERC_CUDA(cudaMalloc(&d_stats, sizeof(long long) * BATCH_NSTATS));
    PERC_CUDA(cudaMemsetAsync(d_stats, 0, sizeof(long long) * BATCH_NSTATS, st));
    // labeling of the realizations: round-robin over the ways (statistics: one block of words per way)
    std::vector<Ctx*> ways;
    rc = batch_contexts(c, nreal, ways);
    const int K = (int)ways.size();
    if (!rc) { PERC_CUDA(cudaMalloc(&d_wstats, sizeof(long long) * BATCH_NSTATS * K)); PERC_CUDA(cudaMemset(d_wstats, 0, sizeof(long long) * BATCH_NSTATS * K)); PERC_CUDA(cudaDeviceSynchronize()); }
    for (Ctx* w : ways) w->batch_thr = true;
    for (int i = 0; i < nreal && rc == 0; ++i) {
        Ctx* w = ways[i % K];
        long long* ws = d_wstats + BATCH_NSTATS * (i % K);
        rc = occ_generate_dev(w, seed, stream0 + (unsigned long long)i, kind != KIND_BOND ? ks : -1, kind != KIND_SITE ? kb : -1,
                              (unsigned long long*)(ws + 6));
        if (!rc) rc = ccl_launch(w, kind);
        if (!rc) rc = pcg_small_stage(w, d_cf, i);
        if (!rc) { batch_accum_kernel<<<1, 1, 0, w->stream>>>(w->d_sum, nullptr, 0, ws, w->d_thr, kind != KIND_BOND, kind != KIND_SITE); w->launches++; }
    }
    for (Ctx* w : ways) {
        w->batch_thr = false;
        w->labeled = false;
        w->site_src = SRC_NONE; w->bond_src = SRC_NONE;
        cudaError_t e = cudaStreamSynchronize(w->stream);      // every conduct map is staged before the solve starts
        if (!rc && e != cudaSuccess) rc = (int)e;
        if (w != c) { c->launches += w->launches; w->launches = 0; }
    }
    if (!rc) {
        std::vector<long long> hs((size_t)BATCH_NSTATS * K);
        PERC_CUDA(cudaMemcpy(hs.data(), d_wstats, sizeof(long long) * BATCH_NSTATS * K, cudaMemcpyDeviceToHost));
        for (int k = 0; k < BATCH_NSTATS; ++k) { long long sum = 0; for (int j = 0; j < K; ++j) sum += hs[(size_t)j * BATCH_NSTATS + k]; hs[k] = sum; }
        PERC_CUDA(cudaMemcpy(d_stats, hs.data(), sizeof(long long) * BATCH_NSTATS, cudaMemcpyHostToDevice));
        PERC_CUDA(cudaDeviceSynchronize());
    }
    if (!rc) rc = pcg_small_solve(c, d_cf, nreal, Va, g0, gleak, tol, itmax, read_thresh, d_G, d_it, d_err);
    if (!rc) {
        long long h[BATCH_NSTATS];
        PERC_CUDA(cudaMemcpyAsync(G, d_G, sizeof(double) * 2 * nreal, cudaMemcpyDeviceToHost, st));
        PERC_CUDA(cudaMemcpyAsync(iters, d_it, sizeof(int) * nreal, cudaMemcpyDeviceToHost, st));
        PERC_CUDA(cudaMemcpyAsync(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost, st));
        PERC_CUDA(cudaStreamSynchronize(st));
        for (int k = 0; k < BATCH_NSTATS; ++k) stats[k] = h[k];
        rc = (int)cudaGetLastError();
    } else cudaStreamSynchronize(st);
    return rc;
}

}  // namespace perc
