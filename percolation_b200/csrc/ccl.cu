// ccl.cu -- cluster labeling, sizes, spanning test, size histogram.
//
// Replaces the reference's incremental Newman-Ziff-style fills with their O(N) relabel scans
// (Sq/site.f:162-289, Sq/bond.f:165-369, Sq/sitebond.f:187-400, Sq/bondsite.f:182-354), the
// spanning scans (Sq/site.f:309-344, Sq/bond.f:389-432, Sq/sitebond.f:423-458) and the c()
// size table (Sq/site.f:260,278-287) by a static connected-component labeling of the
// occupancy mask.  The partition is order independent; labels are canonical:
// label = smallest member site id (1-based).
//
//   K2  ccl_local_kernel    per 32x32 tile: row runs by warp ballot, union-find in shared
//                           memory with atomicMin root linking, per-root sizes by shared atomics
//   K3  ccl_merge_kernel    unions across tile borders (and periodic wrap) on the global table
//   K3b ccl_flatten_kernel  path compression to roots + reduction of tile-local sizes into roots
//   K4  summary / histogram, K5 spanning
#include <algorithm>
#include <vector>
#include "context.h"

namespace perc {

// ------------------------------------------------------------------------------------------
// union-find primitives (parents always point to a smaller index; the root is the minimum)
// ------------------------------------------------------------------------------------------
// find with path halving.  The plain store races benignly with atomicMin linking: it only ever
// writes an ancestor of `a`, and a failed atomicMin keeps uniting the displaced parent.
__device__ __forceinline__ int sm_find(volatile int* lab, int a)
{
    for (;;) {
        int p = lab[a];
        if (p == a) return a;
        int gp = lab[p];
        if (gp == p) return p;
        lab[a] = gp;
        a = gp;
    }
}

__device__ __forceinline__ void sm_unite(int* lab, int a, int b)
{
    for (;;) {
        a = sm_find(lab, a);
        b = sm_find(lab, b);
        if (a == b) return;
        if (a < b) { int tmp = a; a = b; b = tmp; }
        int old = atomicMin(&lab[a], b);
        if (old == a) return;
        a = old;
    }
}

// global table holds parent + 1 (0 = inactive site)
__device__ __forceinline__ int gl_find(const int32_t* label, int a)
{
    int p;
    while ((p = __ldcg(&label[a]) - 1) != a) a = p;
    return a;
}

__device__ __forceinline__ void gl_unite(int32_t* label, int a, int b)
{
    for (;;) {
        a = gl_find(label, a);
        b = gl_find(label, b);
        if (a == b) return;
        if (a < b) { int tmp = a; a = b; b = tmp; }
        int old = atomicMin(&label[a], b + 1);
        if (old == a + 1) return;
        a = old - 1;
    }
}

// ------------------------------------------------------------------------------------------
// K2: tile-local labeling
// ------------------------------------------------------------------------------------------
constexpr int TW = CCL_TW, TH = CCL_TH, HW = CCL_TW + 2;
constexpr int LOCAL_THREADS = 256;
constexpr int ROWS_PER_WARP = TH / (LOCAL_THREADS / 32);

template <int LAT>
__global__ void __launch_bounds__(LOCAL_THREADS)
ccl_local_kernel(Geom g, int kind, const uint8_t* __restrict__ mask, int32_t* __restrict__ label,
                 int32_t* __restrict__ size, Summary* __restrict__ sum)
{
    __shared__ int lab[TH * TW];
    __shared__ int cnt[TH * TW];
    __shared__ uint8_t msk[(TH + 2) * HW];
    __shared__ unsigned s_lone;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_lone = 0;

    // mask tile with a one-site halo (wrap-aware) -- the halo only feeds the size weights
    for (int k = tid; k < (TH + 2) * HW; k += LOCAL_THREADS) {
        int ly = k / HW - 1, lx = k % HW - 1;
        int gy = y0 + ly, gx = x0 + lx;
        if (g.pbc) { if (gx == -1) gx = g.m - 1; else if (gx == g.m) gx = 0; }
        uint8_t v = 0;
        if (gx >= 0 && gx < g.m && gy >= 0 && gy < g.n) v = mask[(int64_t)gy * g.m + gx];
        msk[k] = v;
    }
    __syncthreads();

    const int gx = x0 + lane;
    const bool colreal = gx < g.m;
    const bool up_type = LAT == LAT_TRIANGULAR && !(gx & 1);

    // 1. horizontal runs: initial label = first site of the run (warp ballot)
#pragma unroll
    for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
        int ly = warp + rr * (LOCAL_THREADS / 32);
        int gy = y0 + ly;
        const uint8_t* row = &msk[(ly + 1) * HW + 1];
        unsigned mk = row[lane];
        bool act = colreal && gy < g.n && (mk & MASK_SITE);
        unsigned left = lane > 0 ? row[lane - 1] : 0u;
        bool connl = act && lane > 0 && (left & MASK_SITE) && (left & MASK_E);
        unsigned L = __ballot_sync(0xffffffffu, connl);
        unsigned z = ~L & (0xffffffffu >> (31 - lane));
        int start = 31 - __clz(z);
        lab[ly * TW + lane] = act ? ly * TW + start : -1;
        cnt[ly * TW + lane] = 0;
    }
    __syncthreads();

    // 2. unions along N / NW / NE edges that stay inside the tile
#pragma unroll
    for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
        int ly = warp + rr * (LOCAL_THREADS / 32);
        int gy = y0 + ly;
        if (ly + 1 >= TH || gy + 1 >= g.n) continue;
        const uint8_t* row = &msk[(ly + 1) * HW + 1];
        const uint8_t* upr = row + HW;
        unsigned mk = row[lane];
        bool act = colreal && (mk & MASK_SITE);
        if (!act) continue;
        int idx = ly * TW + lane;
        if ((mk & MASK_N) && (upr[lane] & MASK_SITE)) {
            // skip when the left neighbour makes the same union (both rows continue a run there)
            bool redundant = false;
            if (lane > 0) {
                unsigned l0 = row[lane - 1], l1 = upr[lane - 1];
                redundant = (l0 & MASK_SITE) && (l0 & MASK_E) && (l0 & MASK_N) && (l1 & MASK_SITE) && (l1 & MASK_E);
            }
            if (!redundant) sm_unite(lab, idx, idx + TW);
        }
        if (up_type) {
            if ((mk & MASK_NW) && lane > 0 && (upr[lane - 1] & MASK_SITE)) sm_unite(lab, idx, idx + TW - 1);
            if ((mk & MASK_NE) && lane < TW - 1 && gx + 1 < g.m && (upr[lane + 1] & MASK_SITE)) sm_unite(lab, idx, idx + TW + 1);
        }
    }
    __syncthreads();

    // 3. roots + per-root size (weights by problem kind)
    int root[ROWS_PER_WARP];
    unsigned lone = 0;
#pragma unroll
    for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
        int ly = warp + rr * (LOCAL_THREADS / 32);
        int gy = y0 + ly;
        const uint8_t* row = &msk[(ly + 1) * HW + 1];
        unsigned mk = row[lane];
        bool real = colreal && gy < g.n;
        bool act = real && (mk & MASK_SITE);
        root[rr] = -1;
        int w = 0;
        if (act) {
            root[rr] = sm_find(lab, ly * TW + lane);
            if (kind == KIND_SITE) w = 1;
            else if (kind == KIND_BOND) w = __popc(mk & MASK_BONDS);
            else {
                // mixed (Sq/sitebond.f:231-305): the site, every owned occupied bond, and every
                // incoming occupied bond whose owner site is unoccupied (dangling onto this site)
                w = 1 + __popc(mk & MASK_BONDS);
                const uint8_t* dnr = row - HW;
                unsigned o;
                o = row[lane - 1]; if ((o & MASK_E) && !(o & MASK_SITE)) w++;          // W neighbour's E bond
                o = dnr[lane];     if ((o & MASK_N) && !(o & MASK_SITE)) w++;          // S neighbour's N bond
                if (LAT == LAT_TRIANGULAR && (gx & 1)) {
                    o = dnr[lane - 1]; if ((o & MASK_NE) && !(o & MASK_SITE)) w++;      // SW neighbour's NE bond
                    o = dnr[lane + 1]; if ((o & MASK_NW) && !(o & MASK_SITE)) w++;      // SE neighbour's NW bond
                }
            }
        } else if (real) {
            if (kind == KIND_BOND) w = __popc(mk & MASK_BONDS);   // cannot happen (site bit set by the mask builder)
            else if (kind == KIND_MIXED) {
                // owned occupied bonds of an unoccupied site: other end occupied -> counted there;
                // other end unoccupied -> a lone bond, its own size-1 cluster (Sq/sitebond.f:231-242)
                const uint8_t* upr = row + HW;
                if ((mk & MASK_E) && !(row[lane + 1] & MASK_SITE)) lone++;
                if ((mk & MASK_N) && !(upr[lane] & MASK_SITE)) lone++;
                if ((mk & MASK_NW) && !(upr[lane - 1] & MASK_SITE)) lone++;
                if ((mk & MASK_NE) && !(upr[lane + 1] & MASK_SITE)) lone++;
            }
        }
        if (w && root[rr] >= 0) atomicAdd(&cnt[root[rr]], w);
    }
    if (kind == KIND_MIXED) {
        lone = __reduce_add_sync(0xffffffffu, lone);
        if (lane == 0 && lone) atomicAdd(&s_lone, lone);
    }
    __syncthreads();

    // 4. write provisional labels (global index of the tile-local root + 1) and root sizes
#pragma unroll
    for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
        int ly = warp + rr * (LOCAL_THREADS / 32);
        int gy = y0 + ly;
        if (!colreal || gy >= g.n) continue;
        int64_t gi = (int64_t)gy * g.m + gx;
        int r = root[rr];
        int32_t lbl = 0, sz = 0;
        if (r >= 0) {
            lbl = (int32_t)((int64_t)(y0 + r / TW) * g.m + x0 + (r % TW)) + 1;
            if (r == ly * TW + lane) sz = cnt[r];
        }
        label[gi] = lbl;
        size[gi] = sz;
    }
    if (kind == KIND_MIXED && tid == 0 && s_lone) atomicAdd(&sum->nlone, (unsigned long long)s_lone);
}

// ------------------------------------------------------------------------------------------
// K3: unions across tile borders (part A: horizontal tile borders, all N/NW/NE edges of the
// last row of each tile row; part B: vertical tile borders and the periodic wrap column)
// ------------------------------------------------------------------------------------------
template <int LAT>
__global__ void __launch_bounds__(256)
ccl_merge_kernel(Geom g, const uint8_t* __restrict__ mask, int32_t* __restrict__ label, int nrowb, int ncolb)
{
    int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t na = (int64_t)nrowb * g.m;
    if (id < na) {
        int k = (int)(id / g.m), x = (int)(id % g.m);
        int y = (k + 1) * TH - 1;
        int64_t i = (int64_t)y * g.m + x;
        unsigned mk = mask[i];
        if (!(mk & MASK_SITE)) return;
        if ((mk & MASK_N) && (mask[i + g.m] & MASK_SITE)) gl_unite(label, (int)i, (int)(i + g.m));
        if (LAT == LAT_TRIANGULAR && !(x & 1)) {
            if (mk & MASK_NW) { int64_t j = bond_other_end(g, x, y, DIR_NW); if (mask[j] & MASK_SITE) gl_unite(label, (int)i, (int)j); }
            if (mk & MASK_NE) { int64_t j = bond_other_end(g, x, y, DIR_NE); if (mask[j] & MASK_SITE) gl_unite(label, (int)i, (int)j); }
        }
        return;
    }
    id -= na;
    int64_t nbcol = (int64_t)ncolb * g.n;
    if (id >= nbcol) return;
    int kb = (int)(id / g.n), y = (int)(id % g.n);
    // boundary kb: between column xl = (kb+1)*TW - 1 and xr = xl + 1; the last one may be the wrap
    int xl = (kb + 1) * TW - 1, xr = xl + 1;
    if (xl >= g.m - 1) { xl = g.m - 1; xr = 0; if (!g.pbc) return; }
    int64_t il = (int64_t)y * g.m + xl, ir = (int64_t)y * g.m + xr;
    unsigned ml = mask[il], mr = mask[ir];
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

// ------------------------------------------------------------------------------------------
// K3b: flatten to roots, fold tile-local sizes into the root's entry
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ccl_flatten_kernel(int64_t t, int32_t* __restrict__ label, int32_t* __restrict__ size)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t) return;
    int32_t l = label[i];
    if (l == 0) return;
    int r = gl_find(label, l - 1);
    label[i] = r + 1;
    int32_t s = size[i];
    if (s > 0 && r != (int)i) {
        atomicAdd(&size[r], s);
        size[i] = 0;
    }
}

// ------------------------------------------------------------------------------------------
// K4: number of clusters, largest cluster (size, min label)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ccl_summary_kernel(int64_t t, const int32_t* __restrict__ size, Summary* __restrict__ sum)
{
    unsigned cnt = 0;
    unsigned long long best = 0;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t; i += stride) {
        int32_t s = size[i];
        if (s > 0) {
            cnt++;
            unsigned long long pk = ((unsigned long long)(unsigned)s << 32) | (unsigned long long)(0xffffffffu - (unsigned)(i + 1));
            if (pk > best) best = pk;
        }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    for (int o = 16; o; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        if (other > best) best = other;
    }
    __shared__ unsigned s_cnt;
    __shared__ unsigned long long s_best;
    if (threadIdx.x == 0) { s_cnt = 0; s_best = 0; }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { if (cnt) atomicAdd(&s_cnt, cnt); if (best) atomicMax(&s_best, best); }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_cnt) atomicAdd(&sum->ncl, (unsigned long long)s_cnt);
        if (s_best) atomicMax(&sum->maxpack, s_best);
    }
}

// exact histogram: hist[s-1] for s < nbins, hist[nbins-1] = sizes >= nbins
__global__ void __launch_bounds__(256)
ccl_hist_kernel(int64_t t, const int32_t* __restrict__ size, int nbins, unsigned long long* __restrict__ hist)
{
    extern __shared__ unsigned sh_hist[];
    int nsh = nbins < 4096 ? nbins : 4096;
    for (int k = threadIdx.x; k < nsh; k += blockDim.x) sh_hist[k] = 0;
    __syncthreads();
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t; i += stride) {
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

// ------------------------------------------------------------------------------------------
// K5: spanning clusters = labels present in row 0 and in row n-1
// ------------------------------------------------------------------------------------------
__global__ void span_mark_bottom_kernel(int m, const int32_t* __restrict__ label, int32_t* __restrict__ mark)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= m) return;
    int32_t l = label[x];
    if (l) mark[l - 1] = 1;
}

__global__ void span_collect_top_kernel(int m, int64_t top0, const int32_t* __restrict__ label,
                                        int32_t* __restrict__ mark, int32_t* __restrict__ ids, Summary* __restrict__ sum)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= m) return;
    int32_t l = label[top0 + x];
    if (!l) return;
    if (mark[l - 1] == 1) {
        if (atomicExch(&mark[l - 1], 2) == 1) {
            int pos = atomicAdd(&sum->nspan, 1);
            if (pos < MAX_SPAN) ids[pos] = l; else sum->span_overflow = 1;
        }
    }
}

__global__ void span_clear_kernel(int m, int64_t top0, const int32_t* __restrict__ label, int32_t* __restrict__ mark)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= m) return;
    int32_t l = label[x];
    if (l) mark[l - 1] = 0;
    l = label[top0 + x];
    if (l) mark[l - 1] = 0;
}

// ------------------------------------------------------------------------------------------
// bond labels in reference row order: b(row,3)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
export_bond_labels_kernel(Geom g, int kind, const uint8_t* __restrict__ mask, const int32_t* __restrict__ label,
                          int32_t* __restrict__ b3)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= g.nb) return;
    int64_t a; int dir;
    ref_row_to_owner(g, r, &a, &dir);
    unsigned mk = mask[a];
    int32_t out = 0;
    if (kind != KIND_SITE && ((mk >> (dir + 1)) & 1u)) {
        if (kind == KIND_BOND) out = label[a];
        else {
            int64_t b = bond_other_end(g, (int)(a % g.m), (int)(a / g.m), dir);
            if (mk & MASK_SITE) out = label[a];
            else if (mask[b] & MASK_SITE) out = label[b];
            else out = (int32_t)(g.t + r + 1);
        }
    }
    b3[r] = out;
}

// ------------------------------------------------------------------------------------------
// host drivers
// ------------------------------------------------------------------------------------------
static unsigned nblk(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

int ccl_run(Ctx* c, int kind)
{
    const Geom& g = c->g;
    cudaStream_t st = c->stream;
    c->labeled = false;
    c->solved = false;
    PERC_CUDA(cudaMemsetAsync(c->d_sum, 0, sizeof(Summary), st));
    PERC_CUDA(cudaEventRecord(c->ev[0], st));
    int rc = occ_build_mask(c, kind);
    if (rc) return rc;
    PERC_CUDA(cudaEventRecord(c->ev[1], st));

    dim3 grid((g.m + TW - 1) / TW, (g.n + TH - 1) / TH);
    if (g.lattice == LAT_SQUARE)
        ccl_local_kernel<LAT_SQUARE><<<grid, LOCAL_THREADS, 0, st>>>(g, kind, c->mask, c->label, c->size, c->d_sum);
    else
        ccl_local_kernel<LAT_TRIANGULAR><<<grid, LOCAL_THREADS, 0, st>>>(g, kind, c->mask, c->label, c->size, c->d_sum);
    c->launches++;
    PERC_CUDA(cudaEventRecord(c->ev[2], st));

    int nrowb = (g.n - 1) / TH;                         // tile rows that have a tile above
    int ncolb = (g.m - 1) / TW;                         // vertical tile borders inside the lattice
    if (g.pbc) ncolb += 1;                              // plus the wrap column
    int64_t nthreads = (int64_t)nrowb * g.m + (int64_t)ncolb * g.n;
    if (nthreads > 0) {
        // without pbc the last boundary index must not alias the wrap column
        if (g.lattice == LAT_SQUARE)
            ccl_merge_kernel<LAT_SQUARE><<<nblk(nthreads), 256, 0, st>>>(g, c->mask, c->label, nrowb, ncolb);
        else
            ccl_merge_kernel<LAT_TRIANGULAR><<<nblk(nthreads), 256, 0, st>>>(g, c->mask, c->label, nrowb, ncolb);
        c->launches++;
    }
    PERC_CUDA(cudaEventRecord(c->ev[3], st));

    ccl_flatten_kernel<<<nblk(g.t), 256, 0, st>>>(g.t, c->label, c->size);
    c->launches++;
    ccl_summary_kernel<<<148 * 8, 256, 0, st>>>(g.t, c->size, c->d_sum);
    c->launches++;
    PERC_CUDA(cudaEventRecord(c->ev[4], st));

    int64_t top0 = (int64_t)(g.n - 1) * g.m;
    span_mark_bottom_kernel<<<nblk(g.m), 256, 0, st>>>(g.m, c->label, c->span_mark);
    span_collect_top_kernel<<<nblk(g.m), 256, 0, st>>>(g.m, top0, c->label, c->span_mark, c->span_ids, c->d_sum);
    span_clear_kernel<<<nblk(g.m), 256, 0, st>>>(g.m, top0, c->label, c->span_mark);
    c->launches += 3;
    PERC_CUDA(cudaEventRecord(c->ev[5], st));
    PERC_CUDA(cudaGetLastError());
    c->kind = kind;
    c->labeled = true;
    return ccl_fetch_summary(c);
}

int ccl_fetch_summary(Ctx* c)
{
    cudaStream_t st = c->stream;
    PERC_CUDA(cudaMemcpyAsync(&c->h_sum, c->d_sum, sizeof(Summary), cudaMemcpyDeviceToHost, st));
    PERC_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < 5; ++k) cudaEventElapsedTime(&c->phase_ms[k], c->ev[k], c->ev[k + 1]);
    int ns = c->h_sum.nspan < MAX_SPAN ? c->h_sum.nspan : MAX_SPAN;
    c->h_span_ids.assign((size_t)ns, 0);
    c->h_span_sizes.assign((size_t)ns, 0);
    if (ns > 0) {
        PERC_CUDA(cudaMemcpy(c->h_span_ids.data(), c->span_ids, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost));
        std::sort(c->h_span_ids.begin(), c->h_span_ids.end());
        for (int k = 0; k < ns; ++k)
            PERC_CUDA(cudaMemcpy(&c->h_span_sizes[k], c->size + (c->h_span_ids[k] - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
    }
    return 0;
}

int ccl_hist(Ctx* c, int nbins, int64_t* hist)
{
    if (nbins < 1) return -1;
    unsigned long long* d = (unsigned long long*)ctx_dev_stage(c, sizeof(unsigned long long) * nbins);
    if (!d) return (int)cudaErrorMemoryAllocation;
    PERC_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long) * nbins, c->stream));
    int nsh = nbins < 4096 ? nbins : 4096;
    ccl_hist_kernel<<<148 * 4, 256, sizeof(unsigned) * nsh, c->stream>>>(c->g.t, c->size, nbins, d);
    c->launches++;
    PERC_CUDA(cudaMemcpyAsync(hist, d, sizeof(unsigned long long) * nbins, cudaMemcpyDeviceToHost, c->stream));
    PERC_CUDA(cudaStreamSynchronize(c->stream));
    if (c->kind == KIND_MIXED) hist[0] += (int64_t)c->h_sum.nlone;     // lone bonds are size-1 clusters
    return 0;
}

int ccl_export_bond_labels(Ctx* c, int32_t* b3)
{
    int64_t nb = c->g.nb;
    int32_t* d = (int32_t*)ctx_dev_stage(c, sizeof(int32_t) * nb);
    if (!d) return (int)cudaErrorMemoryAllocation;
    export_bond_labels_kernel<<<nblk(nb), 256, 0, c->stream>>>(c->g, c->kind, c->mask, c->label, d);
    c->launches++;
    PERC_CUDA(cudaMemcpyAsync(b3, d, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, c->stream));
    PERC_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

}  // namespace perc
