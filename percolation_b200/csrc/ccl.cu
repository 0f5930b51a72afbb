// ccl.cu -- cluster labeling, sizes, spanning test, size histogram (kernels + host driver).
//
// Replaces the reference's incremental Newman-Ziff-style fills with their O(N) relabel scans
// (Sq/site.f:162-289, Sq/bond.f:165-369, Sq/sitebond.f:187-400, Sq/bondsite.f:182-354), the
// spanning scans (Sq/site.f:309-344, Sq/bond.f:389-432, Sq/sitebond.f:423-458) and the c()
// size table (Sq/site.f:260,278-287) by a static connected-component labeling of the
// occupancy mask.  The partition is order independent; labels are canonical:
// label = smallest member site id (1-based).
//
// The algorithm itself lives in ccl_tile.cuh (shared with the host emulation used by the CPU
// tests); this file wraps it into kernels:
//   K2  ccl_local_kernel    128x64 tile, one thread per 32-site word: bit planes, run-based
//                           union-find in shared memory, provisional labels, root list
//   K3  ccl_merge_kernel    unions across tile borders (and the periodic wrap) on the global table
//   K3b ccl_rootfix_kernel  tile-local roots -> global roots, sizes folded (K4: ncl, largest cluster)
//   K3c ccl_flatten_kernel  every site: one hop to its global root, 128-bit accesses
//   K5  ccl_span_kernel     labels present in row 0 and in row n-1
#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "context.h"

namespace perc {

// ------------------------------------------------------------------------------------------
// K2: tile-local labeling
// ------------------------------------------------------------------------------------------
template <int LAT, int KIND, int VAR>
__global__ void __launch_bounds__(CT_THREADS)
ccl_local_kernel(Geom g, const uint8_t* __restrict__ mask, int32_t* __restrict__ label, int32_t* __restrict__ size,
                 int32_t* __restrict__ rootlist, Summary* __restrict__ sum, int vec)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem& s = *reinterpret_cast<TileSmem*>(smem_raw);
    const int tid = threadIdx.x, x0 = blockIdx.x * CT_TW, y0 = blockIdx.y * CT_TH;
    TileRegs r;
    tile_phase0<LAT, KIND>(s, g, mask, x0, y0, tid, vec != 0);
    __syncthreads();
    tile_phase1(s, tid);
    __syncthreads();
    tile_phase2_words(s, tid);
    __syncthreads();
    tile_phase2_level<LAT, 1>(s, tid); __syncthreads();
    tile_phase2_level<LAT, 2>(s, tid); __syncthreads();
    tile_phase2_level<LAT, 3>(s, tid); __syncthreads();
    tile_phase2_level<LAT, 4>(s, tid); __syncthreads();
    tile_phase2_level<LAT, 5>(s, tid); __syncthreads();
    tile_phase2_level<LAT, 6>(s, tid); __syncthreads();
    tile_clear_ring(s, tid); __syncthreads();
    if (VAR == 2) tile_phase3_pair<LAT, KIND>(s, g, x0, y0, tid, r); else tile_phase3<LAT, KIND, VAR>(s, g, x0, y0, tid, r);
    __syncthreads();
    tile_phase4_fill(s, g, x0, y0, tid, r, size);
    __syncthreads();
    if (tid == 0) tile_phase4_reserve(s, sum);
    tile_phase4_labels<VAR>(s, g, x0, y0, tid, label, vec != 0);
    __syncthreads();
    tile_phase4_roots(s, g, x0, y0, tid, r, size, rootlist);
}

// ------------------------------------------------------------------------------------------
// K3: unions across tile borders
// ------------------------------------------------------------------------------------------
template <int LAT>
__global__ void __launch_bounds__(128)
ccl_merge_kernel(Geom g, const uint8_t* __restrict__ mask, int32_t* __restrict__ label, int nrowb, int nwords,
                 int ncolb, int vec)
{
    int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    merge_item<LAT>(g, mask, label, nrowb, nwords, ncolb, id, vec != 0);
}

// ------------------------------------------------------------------------------------------
// K3b: root list -> global roots; sizes; number of clusters; largest cluster
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ccl_rootfix_kernel(int32_t* __restrict__ label, int32_t* __restrict__ size, const int32_t* __restrict__ rootlist,
                   Summary* __restrict__ sum, int count)
{
    const unsigned nroots = sum->nroots;
    unsigned cnt = 0;
    unsigned long long best = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t kend = ((int64_t)nroots + 31) / 32 * 32;            // whole warps iterate together
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < kend; k += stride) {
        int isroot;
        unsigned long long pk = rootfix_item(label, size, rootlist, k, k < nroots, &isroot);
        cnt += isroot;
        if (pk > best) best = pk;
    }
    if (!count) return;                 // slab handles count afterwards (a root may own no site here)
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

// ------------------------------------------------------------------------------------------
// K3c: flatten.  After rootfix every tile-local root points at its global root, so one hop from
// the provisional label is enough.  Swept from the END of the array: the local kernel wrote the
// labels front to back, the tail is what is still resident in L2.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ccl_flatten4_kernel(int64_t nquad, int32_t* label)
{
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nquad) return;
    q = nquad - 1 - q;
    int4 l = reinterpret_cast<int4*>(label)[q];
    int4 o;
    o.x = flatten_one(label, l.x);
    o.y = l.y == l.x ? o.x : flatten_one(label, l.y);
    o.z = l.z == l.y ? o.y : flatten_one(label, l.z);
    o.w = l.w == l.z ? o.z : flatten_one(label, l.w);
    if (o.x != l.x || o.y != l.y || o.z != l.z || o.w != l.w) reinterpret_cast<int4*>(label)[q] = o;
}

__global__ void __launch_bounds__(256)
ccl_flatten1_kernel(int64_t first, int64_t t, int32_t* label)
{
    int64_t i = first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t) return;
    int32_t l = label[i];
    int32_t o = flatten_one(label, l);
    if (o != l) label[i] = o;
}

// ------------------------------------------------------------------------------------------
// K5: spanning clusters = labels present in row 0 and in row n-1.  Labels are the
// smallest member id, so a cluster touches row 0 iff its label is <= m: one pass over the top row,
// duplicates removed with a shared-memory bitmap over the m possible labels.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
ccl_span_kernel(int m, int64_t top0, const int32_t* __restrict__ label, const int32_t* __restrict__ size,
                Summary* __restrict__ sum)
{
    // the bitmap covers SPAN_CHUNK labels at a time (any m fits the 48 KB static limit); a label is examined in the pass
    // its chunk belongs to
    constexpr int SPAN_CHUNK = 1 << 18;
    __shared__ unsigned seen[SPAN_CHUNK / 32];
    for (int base = 0; base < m; base += SPAN_CHUNK) {
        const int nw = ((m - base < SPAN_CHUNK ? m - base : SPAN_CHUNK) + 31) / 32;
        __syncthreads();
        for (int k = threadIdx.x; k < nw; k += blockDim.x) seen[k] = 0;
        __syncthreads();
        for (int x = threadIdx.x; x < m; x += blockDim.x) {
            const int32_t l = label[top0 + x];
            if (l < 1 || l > m) continue;
            const int r = l - 1 - base;
            if (r < 0 || r >= SPAN_CHUNK) continue;
            const unsigned bit = 1u << (r & 31);
            if (atomicOr(&seen[r >> 5], bit) & bit) continue;
            const int32_t sz = size[l - 1];
            atomicMax(&sum->span_best, ((unsigned long long)(0xffffffffu - (unsigned)l) << 32) | (unsigned)sz);
            const int pos = atomicAdd(&sum->nspan, 1);
            if (pos < MAX_SPAN) { sum->span_ids[pos] = l; sum->span_sizes[pos] = sz; }
            else sum->span_overflow = 1;
        }
    }
}

// ------------------------------------------------------------------------------------------
// K4: exact histogram: hist[s-1] for s < nbins, hist[nbins-1] = sizes >= nbins
// ------------------------------------------------------------------------------------------
// logbin = 0: hist[s-1] for s < nbins, hist[nbins-1] = sizes >= nbins;  logbin = 1: hist[floor(log2 s)]
__global__ void __launch_bounds__(256)
ccl_hist_kernel(int64_t t, const int32_t* __restrict__ label, const int32_t* __restrict__ size, int nbins,
                unsigned long long* __restrict__ hist, int logbin)
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
            int b = logbin ? 31 - __clz(s) : s - 1;
            if (b >= nbins) b = nbins - 1;
            if (b < nsh) atomicAdd(&sh_hist[b], 1u);
            else atomicAdd(&hist[b], 1ull);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nsh; k += blockDim.x)
        if (sh_hist[k]) atomicAdd(&hist[k], (unsigned long long)sh_hist[k]);
}

// c(t) in the reference's shape: c(label) = size, 0 elsewhere
__global__ void __launch_bounds__(256)
export_sizes_kernel(int64_t t, const int32_t* __restrict__ label, const int32_t* __restrict__ size, int32_t* __restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < t) out[i] = label[i] == (int32_t)(i + 1) ? size[i] : 0;
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

// the tile kernel needs more than 48 KB of dynamic shared memory: the opt-in is a per-DEVICE function attribute, so it is
// tracked per handle (a process may hold handles on several GPUs), not per process
template <int LAT, int KIND>
static cudaError_t launch_local(Ctx* c, dim3 grid, int vec)
{
    constexpr int VAR = 2;          // per-site roots derived in the label phase, two runs per trip of the per-run loop
    const unsigned bit = 1u << ((LAT - 1) * 3 + (KIND - 1));
    if (!(c->ccl_attr & bit)) {
        cudaError_t e = cudaFuncSetAttribute(ccl_local_kernel<LAT, KIND, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)tile_smem_bytes<LAT>());
        if (e != cudaSuccess) return e;
        c->ccl_attr |= bit;
    }
    ccl_local_kernel<LAT, KIND, VAR><<<grid, CT_THREADS, tile_smem_bytes<LAT>(), c->stream>>>(c->g, c->mask, c->label, c->size,
                                                                                       c->rootlist, c->d_sum, vec);
    return cudaGetLastError();
}

int ccl_launch(Ctx* c, int kind)
{
    const Geom& g = c->g;
    cudaStream_t st = c->stream;
    c->labeled = false;
    c->solved = false;
    PERC_CUDA(cudaMemsetAsync(c->d_sum, 0, offsetof(Summary, span_ids), st));
    PERC_CUDA(cudaEventRecord(c->ev[0], st));
    int rc = occ_build_mask(c, kind);
    if (rc) return rc;
    PERC_CUDA(cudaEventRecord(c->ev[1], st));

    const int vec = (g.m % 16) == 0;          // rows of the byte mask are 16-byte aligned -> 128-bit accesses
    const bool sq = g.lattice == LAT_SQUARE;
    dim3 grid((g.m + CT_TW - 1) / CT_TW, (g.n + CT_TH - 1) / CT_TH);
    cudaError_t e;
    if (sq) e = kind == KIND_SITE ? launch_local<LAT_SQUARE, KIND_SITE>(c, grid, vec)
              : kind == KIND_BOND ? launch_local<LAT_SQUARE, KIND_BOND>(c, grid, vec)
                                  : launch_local<LAT_SQUARE, KIND_MIXED>(c, grid, vec);
    else    e = kind == KIND_SITE ? launch_local<LAT_TRIANGULAR, KIND_SITE>(c, grid, vec)
              : kind == KIND_BOND ? launch_local<LAT_TRIANGULAR, KIND_BOND>(c, grid, vec)
                                  : launch_local<LAT_TRIANGULAR, KIND_MIXED>(c, grid, vec);
    PERC_CUDA(e);
    c->launches++;
    PERC_CUDA(cudaEventRecord(c->ev[2], st));

    int nrowb = (g.n - 1) / CT_TH;                      // tile rows that have a tile above
    int nwords = (g.m + 31) / 32;
    int ncolb = (g.m - 1) / CT_TW;                      // vertical tile borders inside the lattice
    if (g.pbc) ncolb += 1;                              // plus the wrap column
    int64_t nitems = (int64_t)nrowb * nwords * 32 + (int64_t)ncolb * g.n;
    if (nitems > 0) {
        if (sq) ccl_merge_kernel<LAT_SQUARE><<<nblk(nitems, 128), 128, 0, st>>>(g, c->mask, c->label, nrowb, nwords, ncolb, vec);
        else    ccl_merge_kernel<LAT_TRIANGULAR><<<nblk(nitems, 128), 128, 0, st>>>(g, c->mask, c->label, nrowb, nwords, ncolb, vec);
        c->launches++;
    }
    PERC_CUDA(cudaEventRecord(c->ev[3], st));

    const bool slab = c->nranks > 1;
    ccl_rootfix_kernel<<<148 * 8, 256, 0, st>>>(c->label, c->size, c->rootlist, c->d_sum, slab ? 0 : 1);
    int64_t nquad = g.t / 4;
    if (nquad) ccl_flatten4_kernel<<<nblk(nquad), 256, 0, st>>>(nquad, c->label);
    if (g.t % 4) ccl_flatten1_kernel<<<1, 256, 0, st>>>(nquad * 4, g.t, c->label);
    c->launches += 2 + (g.t % 4 ? 1 : 0);
    PERC_CUDA(cudaEventRecord(c->ev[4], st));

    if (slab) {
        // slab of a decomposed lattice: exact rank-local counts, then the clusters are stitched across the
        // interfaces (all-gather of the interface rows + redundant union-find); spanning comes out of that
        rc = slab_count_roots(c);
        if (rc) return rc;
        rc = slab_stitch(c);
        if (rc) return rc;
    } else {
        ccl_span_kernel<<<1, 1024, 0, st>>>(g.m, (int64_t)(g.n - 1) * g.m, c->label, c->size, c->d_sum);
        c->launches++;
    }
    PERC_CUDA(cudaEventRecord(c->ev[5], st));
    PERC_CUDA(cudaGetLastError());
    c->kind = kind;
    return 0;
}

void ccl_span_launch(Ctx* c)
{
    ccl_span_kernel<<<1, 1024, 0, c->stream>>>(c->g.m, (int64_t)(c->g.n - 1) * c->g.m, c->label, c->size, c->d_sum);
}

void ccl_note_labeled(Ctx* c, int kind)
{
    c->lab_valid = c->nranks == 1;
    c->lab_kind = kind; c->lab_site_src = c->site_src; c->lab_bond_src = c->bond_src;
    c->lab_ks = c->ks; c->lab_kb = c->kb; c->lab_seed = c->seed; c->lab_stream = c->stream_id; c->lab_epoch = c->occ_epoch;
}

int ccl_run(Ctx* c, int kind)
{
    c->lab_valid = false;
    int rc = ccl_launch(c, kind);
    if (rc) return rc;
    const bool slab = c->nranks > 1;
    c->labeled = true;
    rc = ccl_fetch_summary(c);
    if (rc) return rc;
    if (slab) {
        c->h_span_gid = c->stitch.span_gid; c->h_span_total = c->stitch.span_size;
        c->h_sum.nspan = (int)c->h_span_gid.size();
        c->h_span_ids.clear(); c->h_span_sizes.clear();
    } else {
        c->h_span_gid.assign(c->h_span_ids.begin(), c->h_span_ids.end());
        c->h_span_total.assign(c->h_span_sizes.begin(), c->h_span_sizes.end());
    }
    ccl_note_labeled(c, kind);
    return 0;
}

int ccl_fetch_summary(Ctx* c)
{
    cudaStream_t st = c->stream;
    PERC_CUDA(cudaMemcpyAsync(c->h_sum_pin, c->d_sum, sizeof(Summary), cudaMemcpyDeviceToHost, st));
    PERC_CUDA(cudaStreamSynchronize(st));
    int ns = c->h_sum_pin->nspan < MAX_SPAN ? c->h_sum_pin->nspan : MAX_SPAN;
    memcpy(&c->h_sum, c->h_sum_pin, offsetof(Summary, span_ids));
    for (int k = 0; k < 5; ++k) cudaEventElapsedTime(&c->phase_ms[k], c->ev[k], c->ev[k + 1]);
    std::vector<std::pair<int32_t, int32_t>> v((size_t)ns);
    for (int k = 0; k < ns; ++k) v[k] = {c->h_sum_pin->span_ids[k], c->h_sum_pin->span_sizes[k]};
    std::sort(v.begin(), v.end());
    c->h_span_ids.assign((size_t)ns, 0);
    c->h_span_sizes.assign((size_t)ns, 0);
    for (int k = 0; k < ns; ++k) { c->h_span_ids[k] = v[k].first; c->h_span_sizes[k] = v[k].second; }
    // more spanning clusters than the list holds (strip lattices): which ones made it into the list depends on the order
    // of the atomics, the smallest id -- perccln, the default cluster of perc_conduct -- does not (span_best)
    if (c->h_sum.span_overflow && ns > 0) {
        const int32_t bid = (int32_t)(0xffffffffu - (unsigned)(c->h_sum.span_best >> 32)), bsz = (int32_t)(c->h_sum.span_best & 0xffffffffu);
        if (c->h_span_ids[0] != bid) {
            c->h_span_ids.insert(c->h_span_ids.begin(), bid); c->h_span_sizes.insert(c->h_span_sizes.begin(), bsz);
            c->h_span_ids.pop_back(); c->h_span_sizes.pop_back();
        }
    }
    return 0;
}

int ccl_hist(Ctx* c, int nbins, int64_t* hist, int logbin)
{
    if (nbins < 1) return -1;
    unsigned long long* d = (unsigned long long*)ctx_dev_stage(c, sizeof(unsigned long long) * nbins);
    if (!d) return (int)cudaErrorMemoryAllocation;
    PERC_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long) * nbins, c->stream));
    int nsh = nbins < 4096 ? nbins : 4096;
    ccl_hist_kernel<<<148 * 4, 256, sizeof(unsigned) * nsh, c->stream>>>(c->g.t, c->label, c->size, nbins, d, logbin);
    c->launches++;
    PERC_CUDA(cudaMemcpyAsync(hist, d, sizeof(unsigned long long) * nbins, cudaMemcpyDeviceToHost, c->stream));
    PERC_CUDA(cudaStreamSynchronize(c->stream));
    if (c->kind == KIND_MIXED) hist[0] += (int64_t)c->h_sum.nlone;     // lone bonds are size-1 clusters
    return 0;
}

int ccl_export_sizes(Ctx* c, int32_t* cs)
{
    int64_t t = c->g.t;
    int32_t* d = (int32_t*)ctx_dev_stage(c, sizeof(int32_t) * t);
    if (!d) return (int)cudaErrorMemoryAllocation;
    export_sizes_kernel<<<nblk(t), 256, 0, c->stream>>>(t, c->label, c->size, d);
    c->launches++;
    PERC_CUDA(cudaMemcpyAsync(cs, d, sizeof(int32_t) * t, cudaMemcpyDeviceToHost, c->stream));
    PERC_CUDA(cudaStreamSynchronize(c->stream));
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
