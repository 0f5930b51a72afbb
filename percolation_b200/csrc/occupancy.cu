// occupancy.cu -- occupancy inputs and the K1 occupancy generator.
//
// Replaces the shuffle blocks of the reference (Fortran/permute.f:39-44, Sq/site.f:131-147,
// Sq/bond.f:137-150) and the "occupy the first k of the order" rule (Sq/site.f:164-176,
// Sq/bond.f:167, Sq/sitebond.f:187-196).  Two sources feed one mask builder:
//   * rank tables built from a caller-supplied order (identical inputs to the reference);
//   * counter-based Philox-4x32-10 keys, exactly-k smallest occupied (no table in HBM).
// The mask byte per site: bit0 site active, bits 1..4 owned bonds E, N, NW, NE occupied.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include "context.h"
#include "philox.cuh"

namespace perc {

// ------------------------------------------------------------------------------------------
// sources
// ------------------------------------------------------------------------------------------
struct RankSrc {
    const int32_t* srank;
    const int32_t* brank;
    int64_t t;
    int ks, kb;
    __device__ __forceinline__ bool site(int64_t i) const { return srank != nullptr && srank[i] < ks; }
    __device__ __forceinline__ bool bond(int dir, int64_t i) const { return brank != nullptr && brank[dir * t + i] < kb; }
};

struct PhiloxSrc {
    unsigned long long seed, stream;
    int64_t t;              // sites of the WHOLE lattice: element ids are global, so a slab of a decomposed
                            // lattice draws exactly the occupancy the single-GPU run draws
    PhiloxThreshold ts, tb;
    __device__ __forceinline__ bool site(int64_t i) const
    {
        if (!ts.enabled) return false;
        if (ts.all) return true;
        unsigned long long k = site_key(seed, stream, (unsigned long long)i);
        return k < ts.key || (k == ts.key && (unsigned long long)i <= ts.id);
    }
    __device__ __forceinline__ bool bond(int dir, int64_t i) const
    {
        if (!tb.enabled) return false;
        if (tb.all) return true;
        unsigned long long id = (unsigned long long)dir * t + i;
        unsigned long long k = bond_key(seed, stream, dir, (unsigned long long)i);
        return k < tb.key || (k == tb.key && id <= tb.id);
    }
};

// the E/N (pair 0) or NW/NE (pair 1) bonds owned by site i from ONE Philox call: bits 1 (first) | 2 (second)
__device__ __forceinline__ unsigned philox_bond_pair(const PhiloxSrc& p, int pair, int64_t i)
{
    if (!p.tb.enabled) return 0;
    if (p.tb.all) return 3u;
    unsigned long long A, B;
    elem_key_pair(p.seed, p.stream, 1 + pair, (unsigned long long)i, A, B);
    const unsigned long long idA = (unsigned long long)(2 * pair) * p.t + i, idB = (unsigned long long)(2 * pair + 1) * p.t + i;
    unsigned r = 0;
    if (A < p.tb.key || (A == p.tb.key && idA <= p.tb.id)) r |= 1u;
    if (B < p.tb.key || (B == p.tb.key && idB <= p.tb.id)) r |= 2u;
    return r;
}

// mixed source: sites and bonds may come from different inputs
struct AnySrc {
    RankSrc r;
    PhiloxSrc p;
    int site_src, bond_src;
    __device__ __forceinline__ bool site(int64_t i) const
    {
        return site_src == SRC_RANK ? r.site(i) : site_src == SRC_PHILOX ? p.site(i) : false;
    }
    __device__ __forceinline__ bool bond(int dir, int64_t i) const
    {
        return bond_src == SRC_RANK ? r.bond(dir, i) : bond_src == SRC_PHILOX ? p.bond(dir, i) : false;
    }
};

static AnySrc make_src(const Ctx* c)
{
    AnySrc s;
    s.r.srank = c->site_src == SRC_RANK ? c->srank : nullptr;
    s.r.brank = c->bond_src == SRC_RANK ? c->brank : nullptr;
    s.r.t = c->g.t; s.r.ks = (int)c->ks; s.r.kb = (int)c->kb;
    s.p.seed = c->seed; s.p.stream = c->stream_id; s.p.t = c->g.tg;
    s.p.ts = c->thr_site; s.p.tb = c->thr_bond;
    s.site_src = c->site_src; s.bond_src = c->bond_src;
    return s;
}

// ------------------------------------------------------------------------------------------
// mask builder (one thread per site)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) build_mask_kernel(Geom g, int kind, AnySrc src, uint8_t* __restrict__ mask,
                                                         Summary* __restrict__ sum, const PhiloxThreshold* __restrict__ dthr)
{
    // batch mode: the exact-count thresholds were selected on the device and never visited the host
    if (dthr) { src.p.ts = dthr[0]; src.p.tb = dthr[1]; }
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned ns = 0, nbd = 0;
    if (i < g.t) {
        int x = (int)(i % g.m), y = (int)(i / g.m);
        const int64_t ig = i + (int64_t)g.y0 * g.m;      // index in the whole lattice (= i unless this is a slab)
        const int gy = y + g.y0;
        const bool owned = y >= g.own_lo && y < g.own_hi;
        unsigned own = owned_bond_bits(g, x, y);
        unsigned bits = 0;
        bool site = false;
        if (kind == KIND_SITE) {
            site = src.site(ig);
            bits = own;                                   // every lattice bond is present
            ns = site && owned;
        } else {
            if (src.bond_src == SRC_PHILOX) {
                // E and N from one Philox call, NW and NE (triangular up-type sites) from a second one
                if (own & (MASK_E | MASK_N)) bits |= (philox_bond_pair(src.p, 0, ig) << 1) & own;
                if (own & (MASK_NW | MASK_NE)) bits |= (philox_bond_pair(src.p, 1, ig) << 3) & own;
            } else {
#pragma unroll
                for (int d = 0; d < 4; ++d)
                    if ((own >> (d + 1)) & 1u) if (src.bond(d, ig)) bits |= 2u << d;
            }
            nbd = owned ? __popc(bits) : 0;
            if (kind == KIND_MIXED) { site = src.site(ig); ns = site && owned; }
            else {
                // bond problem: a site is a cluster node iff one of its bonds is occupied
                site = bits != 0;
                if (!site && (x > 0 || g.pbc)) site = src.bond(DIR_E, ig - x + (x > 0 ? x - 1 : g.m - 1));
                if (!site && gy > 0) site = src.bond(DIR_N, ig - g.m);
                if (!site && g.lattice == LAT_TRIANGULAR && (x & 1) && gy > 0) {
                    site = src.bond(DIR_NE, ig - g.m - 1);
                    if (!site && (x + 1 < g.m || g.pbc)) site = src.bond(DIR_NW, ig - g.m - x + (x + 1 < g.m ? x + 1 : 0));
                }
            }
        }
        mask[i] = (uint8_t)(bits | (site ? 1u : 0u));
    }
    // occupied-element counts (exact-count check of the generator, perc_summary)
    ns = __reduce_add_sync(0xffffffffu, ns);
    nbd = __reduce_add_sync(0xffffffffu, nbd);
    __shared__ unsigned s_ns, s_nb;
    if (threadIdx.x == 0) { s_ns = 0; s_nb = 0; }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { if (ns) atomicAdd(&s_ns, ns); if (nbd) atomicAdd(&s_nb, nbd); }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_ns) atomicAdd(&sum->nocc_sites, (unsigned long long)s_ns);
        if (s_nb) atomicAdd(&sum->nocc_bonds, (unsigned long long)s_nb);
    }
}

int occ_build_mask(Ctx* c, int kind)
{
    AnySrc src = make_src(c);
    int64_t t = c->g.t;
    unsigned blocks = (unsigned)((t + 255) / 256);
    build_mask_kernel<<<blocks, 256, 0, c->stream>>>(c->g, kind, src, c->mask, c->d_sum, c->batch_thr ? c->d_thr : nullptr);
    c->launches++;
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// rank tables from caller-supplied orders / flags
// ------------------------------------------------------------------------------------------
__global__ void scatter_site_rank_kernel(const int32_t* __restrict__ order, int64_t t, int32_t* __restrict__ srank)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < t) {
        int32_t s = order[i];
        if (s >= 1 && s <= t) srank[s - 1] = (int32_t)i;
    }
}

// border(nb,2) column-major: lo ends then hi ends (Sq/bond.f:40); (lo,hi) -> owner slot
__global__ void scatter_bond_rank_kernel(Geom g, const int32_t* __restrict__ lo, const int32_t* __restrict__ hi,
                                         int32_t* __restrict__ brank, int* __restrict__ bad)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.nb) return;
    int64_t a = (int64_t)lo[i] - 1, b = (int64_t)hi[i] - 1;
    if (a < 0 || b < 0 || a >= g.t || b >= g.t || a >= b) { atomicAdd(bad, 1); return; }
    int x1 = (int)(a % g.m), y1 = (int)(a / g.m), x2 = (int)(b % g.m), y2 = (int)(b / g.m);
    int64_t owner = a; int dir = -1;
    if (y2 == y1) {
        if (x2 == x1 + 1) dir = DIR_E;
        else if (g.pbc && x1 == 0 && x2 == g.m - 1) { owner = b; dir = DIR_E; }
    } else if (y2 == y1 + 1) {
        int dx = x2 - x1;
        if (dx == 0) dir = DIR_N;
        else if (g.lattice == LAT_TRIANGULAR && !(x1 & 1)) {
            if (dx == 1) dir = DIR_NE;
            else if (dx == -1 || (g.pbc && x1 == 0 && x2 == g.m - 1)) dir = DIR_NW;
        }
    }
    if (dir < 0) { atomicAdd(bad, 1); return; }
    brank[(int64_t)dir * g.t + owner] = (int32_t)i;
}

__global__ void flags_to_site_rank_kernel(const uint8_t* __restrict__ socc, int64_t t, int32_t* __restrict__ srank)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < t) srank[i] = socc[i] ? 0 : RANK_NONE;
}

__global__ void flags_to_bond_rank_kernel(Geom g, const uint8_t* __restrict__ bocc, int32_t* __restrict__ brank)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= g.nb) return;
    int64_t site; int dir;
    ref_row_to_owner(g, r, &site, &dir);
    brank[(int64_t)dir * g.t + site] = bocc[r] ? 0 : RANK_NONE;
}

static unsigned nblk(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

int occ_upload_site_order(Ctx* c, const int32_t* order)
{
    int64_t t = c->g.t;
    { int rc = ctx_ensure_ranks(c, true, false); if (rc) return rc; }
    int32_t* d = (int32_t*)ctx_dev_stage(c, sizeof(int32_t) * t);
    if (!d) return (int)cudaErrorMemoryAllocation;
    PERC_CUDA(cudaMemcpyAsync(d, order, sizeof(int32_t) * t, cudaMemcpyHostToDevice, c->stream));
    PERC_CUDA(cudaMemsetAsync(c->srank, 0x7f, sizeof(int32_t) * t, c->stream));
    scatter_site_rank_kernel<<<nblk(t), 256, 0, c->stream>>>(d, t, c->srank);
    c->launches++;
    c->site_src = SRC_RANK;
    c->labeled = false; c->occ_epoch++;
    return (int)cudaGetLastError();
}

int occ_upload_bond_order(Ctx* c, const int32_t* border)
{
    int64_t nb = c->g.nb, t = c->g.t;
    { int rc = ctx_ensure_ranks(c, false, true); if (rc) return rc; }
    int32_t* d = (int32_t*)ctx_dev_stage(c, sizeof(int32_t) * 2 * nb + 64);
    if (!d) return (int)cudaErrorMemoryAllocation;
    int* bad = (int*)(d + 2 * nb);
    PERC_CUDA(cudaMemcpyAsync(d, border, sizeof(int32_t) * 2 * nb, cudaMemcpyHostToDevice, c->stream));
    PERC_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), c->stream));
    PERC_CUDA(cudaMemsetAsync(c->brank, 0x7f, sizeof(int32_t) * c->g.ndir * t, c->stream));
    scatter_bond_rank_kernel<<<nblk(nb), 256, 0, c->stream>>>(c->g, d, d + nb, c->brank, bad);
    c->launches++;
    int hbad = 0;
    PERC_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PERC_CUDA(cudaStreamSynchronize(c->stream));
    if (hbad) return -1;      // PERC_E_ARG: a pair that is not a lattice bond
    c->bond_src = SRC_RANK;
    c->labeled = false; c->occ_epoch++;
    return 0;
}

int occ_upload_flags(Ctx* c, const uint8_t* socc, const uint8_t* bocc)
{
    int64_t t = c->g.t, nb = c->g.nb;
    { int rc = ctx_ensure_ranks(c, socc != nullptr, bocc != nullptr); if (rc) return rc; }
    if (socc) {
        uint8_t* d = (uint8_t*)ctx_dev_stage(c, t);
        if (!d) return (int)cudaErrorMemoryAllocation;
        PERC_CUDA(cudaMemcpyAsync(d, socc, t, cudaMemcpyHostToDevice, c->stream));
        flags_to_site_rank_kernel<<<nblk(t), 256, 0, c->stream>>>(d, t, c->srank);
        c->launches++;
        PERC_CUDA(cudaStreamSynchronize(c->stream));
        c->site_src = SRC_RANK; c->ks = 1;
    }
    if (bocc) {
        uint8_t* d = (uint8_t*)ctx_dev_stage(c, nb);
        if (!d) return (int)cudaErrorMemoryAllocation;
        PERC_CUDA(cudaMemcpyAsync(d, bocc, nb, cudaMemcpyHostToDevice, c->stream));
        PERC_CUDA(cudaMemsetAsync(c->brank, 0x7f, sizeof(int32_t) * c->g.ndir * t, c->stream));
        flags_to_bond_rank_kernel<<<nblk(nb), 256, 0, c->stream>>>(c->g, d, c->brank);
        c->launches++;
        PERC_CUDA(cudaStreamSynchronize(c->stream));
        c->bond_src = SRC_RANK; c->kb = 1;
    }
    c->labeled = false; c->occ_epoch++;
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// export of the raw occupancy in reference layout
// ------------------------------------------------------------------------------------------
__global__ void export_sites_kernel(int64_t t, AnySrc src, uint8_t* __restrict__ socc)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < t) socc[i] = src.site(i) ? 1 : 0;
}

__global__ void export_bonds_kernel(Geom g, AnySrc src, uint8_t* __restrict__ bocc)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= g.nb) return;
    int64_t site; int dir;
    ref_row_to_owner(g, r, &site, &dir);
    bocc[r] = src.bond(dir, site) ? 1 : 0;
}

int occ_export(Ctx* c, uint8_t* socc, uint8_t* bocc)
{
    if (c->nranks > 1) return -4;          // whole-lattice layout: not available on a slab handle
    AnySrc src = make_src(c);
    int64_t t = c->g.t, nb = c->g.nb;
    { int rc = ctx_ensure_ranks(c, socc != nullptr, bocc != nullptr); if (rc) return rc; }
    if (socc) {
        uint8_t* d = (uint8_t*)ctx_dev_stage(c, t);
        if (!d) return (int)cudaErrorMemoryAllocation;
        export_sites_kernel<<<nblk(t), 256, 0, c->stream>>>(t, src, d);
        c->launches++;
        PERC_CUDA(cudaMemcpyAsync(socc, d, t, cudaMemcpyDeviceToHost, c->stream));
        PERC_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (bocc) {
        uint8_t* d = (uint8_t*)ctx_dev_stage(c, nb);
        if (!d) return (int)cudaErrorMemoryAllocation;
        export_bonds_kernel<<<nblk(nb), 256, 0, c->stream>>>(c->g, src, d);
        c->launches++;
        PERC_CUDA(cudaMemcpyAsync(bocc, d, nb, cudaMemcpyDeviceToHost, c->stream));
        PERC_CUDA(cudaStreamSynchronize(c->stream));
    }
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K1: exact-count selection of the k smallest Philox keys
// ------------------------------------------------------------------------------------------
constexpr int SEL_BINS = 4096;
constexpr int SEL_CAND = 8192;

// The selection kernels walk over Philox CALLS (two keys each, philox.cuh) and visit every element of
// the whole lattice once: visit(key, id).  type 0: sites; type 1: bonds (E/N call per site, NW/NE call
// per up-type site of the triangular lattice).
template <typename F>
__device__ __forceinline__ void for_each_key(const Geom& g, int type, unsigned long long seed, unsigned long long stream, F visit)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long A, B;
    if (type == 0) {
        for (int64_t c = first; 2 * c < g.t; c += stride) {
            elem_key_pair(seed, stream, 0, (unsigned long long)c, A, B);
            visit(A, (unsigned long long)(2 * c));
            if (2 * c + 1 < g.t) visit(B, (unsigned long long)(2 * c + 1));
        }
        return;
    }
    for (int64_t i = first; i < g.t; i += stride) {
        const int x = (int)(i % g.m), y = (int)(i / g.m);
        const unsigned own = owned_bond_bits(g, x, y);
        if (own & (MASK_E | MASK_N)) {
            elem_key_pair(seed, stream, 1, (unsigned long long)i, A, B);
            if (own & MASK_E) visit(A, (unsigned long long)i);
            if (own & MASK_N) visit(B, (unsigned long long)g.t + i);
        }
        if (own & (MASK_NW | MASK_NE)) {
            elem_key_pair(seed, stream, 2, (unsigned long long)i, A, B);
            if (own & MASK_NW) visit(A, 2ull * g.t + i);
            if (own & MASK_NE) visit(B, 3ull * g.t + i);
        }
    }
}

// histogram of keys in [lo, lo + SEL_BINS << shift); hist[SEL_BINS] = keys below lo
__global__ void __launch_bounds__(256) select_hist_kernel(Geom g, int type, unsigned long long seed,
                                                          unsigned long long stream, unsigned long long lo,
                                                          int shift, int whole, unsigned long long* __restrict__ hist)
{
    __shared__ unsigned sh[SEL_BINS + 1];
    for (int k = threadIdx.x; k <= SEL_BINS; k += blockDim.x) sh[k] = 0;
    __syncthreads();
    for_each_key(g, type, seed, stream, [&](unsigned long long k, unsigned long long) {
        if (k < lo) { atomicAdd(&sh[SEL_BINS], 1u); return; }
        unsigned long long b = (k - lo) >> shift;
        if (b < (unsigned long long)SEL_BINS) atomicAdd(&sh[(int)b], 1u);
    });
    __syncthreads();
    for (int k = threadIdx.x; k <= SEL_BINS; k += blockDim.x)
        if (sh[k]) atomicAdd(&hist[k], (unsigned long long)sh[k]);
}

// gather (key, id) of elements with key in [lo, lo + width)   (width = 0 means up to 2^64)
__global__ void __launch_bounds__(256) select_gather_kernel(Geom g, int type, unsigned long long seed,
                                                            unsigned long long stream, unsigned long long lo,
                                                            unsigned long long width, unsigned long long* __restrict__ cand,
                                                            int cap, unsigned long long* __restrict__ count)
{
    for_each_key(g, type, seed, stream, [&](unsigned long long k, unsigned long long id) {
        if (k < lo) return;
        if (width != 0 && k - lo >= width) return;
        unsigned long long pos = atomicAdd(count, 1ull);
        if (pos < (unsigned long long)cap) { cand[2 * pos] = k; cand[2 * pos + 1] = id; }
    });
}

// find the k-th smallest (1-based) (key, id) among the N elements of `type`
static int select_threshold(Ctx* c, int type, int64_t N, int64_t k, PhiloxThreshold* out)
{
    // keys are a pure function of the global element id: a slab handle selects over the WHOLE lattice
    // (redundantly on every rank, no communication) and so finds the threshold the single-GPU run finds
    const Geom gw = make_geom(c->g.lattice, c->g.m, c->g.ng, c->g.pbc);
    out->enabled = 1; out->all = 0; out->key = 0; out->id = 0;
    if (k <= 0) { out->enabled = 0; return 0; }
    if (k >= N) { out->all = 1; return 0; }
    std::vector<unsigned long long> hist(SEL_BINS + 1);
    unsigned grid = 148 * 8;
    // estimate window: the k-th smallest of N uniform keys sits near k/N * 2^64
    long double frac = (long double)k / (long double)N;
    long double sigma = sqrtl((long double)N * frac * (1.0L - frac));
    long double hw = (8.0L * sigma + 16.0L) / (long double)N * 18446744073709551616.0L;   // half width in key units
    int shift = 0;
    while (shift < 52 && ldexpl(1.0L, shift) * (SEL_BINS / 2) < hw) shift++;
    unsigned long long lo;
    {
        long double centre = frac * 18446744073709551616.0L;
        long double span_half = ldexpl(1.0L, shift) * (SEL_BINS / 2);
        long double l = centre - span_half;
        long double maxlo = 18446744073709551616.0L - ldexpl(1.0L, shift) * SEL_BINS;
        if (l < 0) l = 0;
        if (l > maxlo) l = maxlo;
        lo = (unsigned long long)l;
        if (shift >= 52) { shift = 52; lo = 0; }
    }
    int64_t below = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        bool ok = true;
        int sh = shift;
        unsigned long long cur_lo = lo;
        below = 0;
        bool first = true;
        for (;;) {
            PERC_CUDA(cudaMemsetAsync(c->d_hist, 0, sizeof(unsigned long long) * (SEL_BINS + 8), c->stream));
            select_hist_kernel<<<grid, 256, 0, c->stream>>>(gw, type, c->seed, c->stream_id, cur_lo, sh, 0, c->d_hist);
            c->launches++;
            PERC_CUDA(cudaMemcpyAsync(hist.data(), c->d_hist, sizeof(unsigned long long) * (SEL_BINS + 1),
                                      cudaMemcpyDeviceToHost, c->stream));
            PERC_CUDA(cudaStreamSynchronize(c->stream));
            if (first) { below = (int64_t)hist[SEL_BINS]; first = false; }
            int64_t inwin = 0;
            for (int b = 0; b < SEL_BINS; ++b) inwin += (int64_t)hist[b];
            if (!(below <= k - 1 && k - 1 < below + inwin)) { ok = false; break; }
            int b = 0; int64_t cum = below;
            while (cum + (int64_t)hist[b] <= k - 1) { cum += (int64_t)hist[b]; b++; }
            cur_lo += (unsigned long long)b << sh;
            below = cum;
            int64_t inbin = (int64_t)hist[b];
            if (inbin <= SEL_CAND || sh == 0) {
                if (inbin > c->cand_cap) return -5;
                unsigned long long width = sh >= 64 ? 0ull : (1ull << sh);
                unsigned long long* cnt = c->d_hist + SEL_BINS + 4;
                PERC_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), c->stream));
                select_gather_kernel<<<grid, 256, 0, c->stream>>>(gw, type, c->seed, c->stream_id, cur_lo, width,
                                                                  c->d_cand, c->cand_cap, cnt);
                c->launches++;
                std::vector<unsigned long long> cand(2 * (size_t)inbin);
                PERC_CUDA(cudaMemcpyAsync(cand.data(), c->d_cand, sizeof(unsigned long long) * 2 * inbin,
                                          cudaMemcpyDeviceToHost, c->stream));
                PERC_CUDA(cudaStreamSynchronize(c->stream));
                std::vector<std::pair<unsigned long long, unsigned long long>> v((size_t)inbin);
                for (int64_t j = 0; j < inbin; ++j) v[j] = {cand[2 * j], cand[2 * j + 1]};
                std::sort(v.begin(), v.end());
                auto pick = v[(size_t)(k - 1 - below)];
                out->key = pick.first; out->id = pick.second;
                return 0;
            }
            sh = sh >= 12 ? sh - 12 : 0;       // refine the chosen bin into SEL_BINS sub-bins
        }
        if (ok) break;
        shift = 52; lo = 0;                    // estimate window missed: restart on the whole key range
    }
    return -5;
}

// ------------------------------------------------------------------------------------------
// K1, device-resident variant for batches of realizations: the same window histogram, but the bin
// scan, the gather of the chosen bin and the pick of the k-th (key, id) all stay on the device, so
// a batch runs without a single host synchronisation.  The window is +-8 sigma around the expected
// key; if it misses, or the bin holds more than SEL_CAND keys, the threshold is flagged invalid and
// the batch reports the realization (the caller re-runs it through occ_generate).
// ------------------------------------------------------------------------------------------
struct SelState { unsigned long long lo, width, below, inbin, count; int ok; int pad; };

__global__ void __launch_bounds__(1024) select_scan_kernel(const unsigned long long* __restrict__ hist, long long k,
                                                           unsigned long long lo, int shift, int cap, SelState* __restrict__ st)
{
    __shared__ unsigned long long part[1024];
    const int tid = threadIdx.x;
    unsigned long long v[4], s = 0;
    for (int j = 0; j < 4; ++j) { v[j] = hist[tid * 4 + j]; s += v[j]; }
    part[tid] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {                     // inclusive scan of the per-thread sums
        unsigned long long add = tid >= o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += add;
        __syncthreads();
    }
    const unsigned long long below = hist[SEL_BINS], total = part[1023];
    const unsigned long long want = (unsigned long long)(k - 1);
    if (tid == 0) { st->ok = (below <= want && want < below + total) ? 1 : 0; st->count = 0; }
    __syncthreads();
    unsigned long long cum = below + part[tid] - s;          // keys before this thread's first bin
    for (int j = 0; j < 4; ++j) {
        if (cum <= want && want < cum + v[j]) {
            st->lo = lo + ((unsigned long long)(tid * 4 + j) << shift);
            st->width = 1ull << shift;
            st->below = cum;
            st->inbin = v[j];
            if (v[j] > (unsigned long long)cap) st->ok = 0;
        }
        cum += v[j];
    }
}

__global__ void __launch_bounds__(256) select_gather_dev_kernel(Geom g, int type, unsigned long long seed, unsigned long long stream,
                                                                SelState* __restrict__ st, unsigned long long* __restrict__ cand, int cap)
{
    if (!st->ok) return;
    const unsigned long long lo = st->lo, width = st->width;
    for_each_key(g, type, seed, stream, [&](unsigned long long k, unsigned long long id) {
        if (k < lo || k - lo >= width) return;
        unsigned long long pos = atomicAdd(&st->count, 1ull);
        if (pos < (unsigned long long)cap) { cand[2 * pos] = k; cand[2 * pos + 1] = id; }
    });
}

// the (k - below)-th smallest (key, id) of the gathered bin, by rank counting
__global__ void __launch_bounds__(1024) select_pick_kernel(long long k, const SelState* __restrict__ st,
                                                           const unsigned long long* __restrict__ cand, int cap,
                                                           PhiloxThreshold* __restrict__ thr, unsigned long long* __restrict__ nfail)
{
    if (threadIdx.x == 0) { thr->enabled = 1; thr->all = 0; thr->failed = 0; }
    if (!st->ok || st->count > (unsigned long long)cap || st->count != st->inbin) {
        if (threadIdx.x == 0) { thr->key = 0; thr->id = 0; thr->enabled = 0; thr->failed = 1; atomicAdd(nfail, 1ull); }
        return;
    }
    const int n = (int)st->count;
    const long long want = k - 1 - (long long)st->below;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long ki = cand[2 * i], ii = cand[2 * i + 1];
        long long rank = 0;
        for (int j = 0; j < n; ++j) {
            const unsigned long long kj = cand[2 * j], ij = cand[2 * j + 1];
            rank += (kj < ki) || (kj == ki && ij < ii);
        }
        if (rank == want) { thr->key = ki; thr->id = ii; }
    }
}

// device-resident threshold of realization (seed, stream): no host synchronisation
static int select_threshold_dev(Ctx* c, int type, int64_t N, int64_t k, unsigned long long seed, unsigned long long stream,
                                PhiloxThreshold* d_out, unsigned long long* d_nfail)
{
    PhiloxThreshold h{};
    h.enabled = 1;
    if (k <= 0 || k >= N) {
        h.enabled = k > 0; h.all = k >= N;
        PERC_CUDA(cudaMemcpyAsync(d_out, &h, sizeof(h), cudaMemcpyHostToDevice, c->stream));   // pageable: staged by the runtime
        return 0;
    }
    const Geom gw = make_geom(c->g.lattice, c->g.m, c->g.ng, c->g.pbc);
    long double frac = (long double)k / (long double)N;
    long double sigma = sqrtl((long double)N * frac * (1.0L - frac));
    long double hw = (8.0L * sigma + 16.0L) / (long double)N * 18446744073709551616.0L;
    int shift = 0;
    while (shift < 52 && ldexpl(1.0L, shift) * (SEL_BINS / 2) < hw) shift++;
    long double centre = frac * 18446744073709551616.0L, span_half = ldexpl(1.0L, shift) * (SEL_BINS / 2);
    long double l = centre - span_half, maxlo = 18446744073709551616.0L - ldexpl(1.0L, shift) * SEL_BINS;
    if (l < 0) l = 0;
    if (l > maxlo) l = maxlo;
    unsigned long long lo = (unsigned long long)l;
    if (shift >= 52) { shift = 52; lo = 0; }
    SelState* st = (SelState*)(c->d_hist + SEL_BINS + 1);
    unsigned grid = 148 * 8;
    PERC_CUDA(cudaMemsetAsync(c->d_hist, 0, sizeof(unsigned long long) * (SEL_BINS + 8), c->stream));
    select_hist_kernel<<<grid, 256, 0, c->stream>>>(gw, type, seed, stream, lo, shift, 0, c->d_hist);
    select_scan_kernel<<<1, 1024, 0, c->stream>>>(c->d_hist, (long long)k, lo, shift, c->cand_cap, st);
    select_gather_dev_kernel<<<grid, 256, 0, c->stream>>>(gw, type, seed, stream, st, c->d_cand, c->cand_cap);
    select_pick_kernel<<<1, 1024, 0, c->stream>>>((long long)k, st, c->d_cand, c->cand_cap, d_out, d_nfail);
    c->launches += 4;
    return (int)cudaGetLastError();
}

// thresholds of one realization of a batch -> c->d_thr[0] (sites), c->d_thr[1] (bonds)
int occ_generate_dev(Ctx* c, unsigned long long seed, unsigned long long stream, int64_t ks, int64_t kb,
                     unsigned long long* d_nfail)
{
    c->seed = seed; c->stream_id = stream;
    int rc = 0;
    if (ks >= 0) {
        rc = select_threshold_dev(c, 0, c->g.tg, ks, seed, stream, c->d_thr + 0, d_nfail);
        if (rc) return rc;
        c->site_src = SRC_PHILOX; c->ks = ks;
    }
    if (kb >= 0) {
        rc = select_threshold_dev(c, 1, c->g.nb, kb, seed, stream, c->d_thr + 1, d_nfail);
        if (rc) return rc;
        c->bond_src = SRC_PHILOX; c->kb = kb;
    }
    c->labeled = false; c->occ_epoch++;
    return 0;
}

int occ_generate(Ctx* c, unsigned long long seed, unsigned long long stream, int64_t ks, int64_t kb)
{
    c->seed = seed; c->stream_id = stream;
    int rc = 0;
    if (ks >= 0) {
        if (ks > c->g.tg) return -1;
        rc = select_threshold(c, 0, c->g.tg, ks, &c->thr_site);
        if (rc) return rc;
        c->site_src = SRC_PHILOX; c->ks = ks;
    }
    if (kb >= 0) {
        if (kb > c->g.nb) return -1;
        rc = select_threshold(c, 1, c->g.nb, kb, &c->thr_bond);
        if (rc) return rc;
        c->bond_src = SRC_PHILOX; c->kb = kb;
    }
    c->labeled = false;
    return 0;
}

}  // namespace perc
