// ccl_incremental.cu -- re-labeling after elements were ADDED to a labeled lattice.
//
// The reference's sweep drivers add elements one by one and relabel as they go (Sq/site_perc.f:133-254,
// Sq/bond_cond.f:208-485, Sq/sb_perc.f, Sq/bs_perc.f): the labeling at one sweep point is the labeling at the
// previous one plus the unions the new elements cause (Newman-Ziff).  perc_label is a from-scratch pass (K1-K5); this
// file is the incremental one for a handle that already holds the labels of a SMALLER fill of the same order /
// generator stream: only the new elements are united on the global label table, the sizes of the roots that lost
// their independence are folded into their new roots, and the weights of the new elements are added.
//
//   K1   build_mask          the new occupancy mask into the second mask buffer (the old one is kept)
//   I1   inc_init            new cluster nodes become their own roots; roots of the old labeling are flagged
//   I2   inc_unite           every connection that exists now and did not before: atomicMin union (gl_unite)
//   I3   inc_fold            per site: root by read-only find, old root's size -> new root, weight difference of the
//                            site (site / owned bonds / dangling bonds, the size rules of ccl_tile.cuh phase 3) -> root,
//                            lone-bond difference, one-hop label written (canonical: the root is the smallest member)
//   I4   inc_stats           number of clusters, largest cluster (same packing as the full pass)
//   K5   ccl_span            unchanged
// When only BONDS were added (bond problem, or mixed problem at fixed sites: the p-sweeps of Sq/bond_cond.f and Sq/sb_perc.f)
// the kernels take four sites per thread (32-bit mask words), the weight of a new bond goes straight to its cluster's root,
// and the number of clusters / the largest cluster are carried over and corrected (new nodes, merges, returned values of the
// size atomics) instead of recounted: I1 + I2 + I3 = 0.026 + 0.082 + 0.096 ms at L = 4096 for a step of 0.5 % of the bonds,
// against 0.25 ms for the from-scratch pass (ccl_local + merge + rootfix + flatten) -- both behind the same 0.245 ms mask build.
// Labels, sizes, counts and spanning clusters are bit-identical to perc_label's on the same fill
// (tests/test_gpu_parity.py::test_incremental_labeling_equals_full_labeling).
#include <cstddef>
#include "context.h"

namespace perc {

namespace {

constexpr unsigned WASROOT = 0x80u;                 // flag bit stashed in the OLD mask byte by inc_init
constexpr unsigned MBITS = 0x1fu;                   // site bit + four owned-bond bits

// weight of site i in the cluster-size count and the lone bonds it owns, from a mask (the rules of tile_phase3:
// site problem = sites; bond problem = owned occupied bonds of a node; mixed = site + owned occupied bonds + occupied
// bonds of unoccupied owners that dangle onto it; a bond with no occupied end is a lone size-1 cluster)
__device__ __forceinline__ int inc_weight(const Geom& g, int kind, const uint8_t* __restrict__ mk, int64_t i, int x, int y, int* lone)
{
    const unsigned m = mk[i] & MBITS;
    const bool S = (m & MASK_SITE) != 0;
    *lone = 0;
    if (kind == KIND_SITE) return S ? 1 : 0;
    const int nown = __popc(m & (MASK_E | MASK_N | MASK_NW | MASK_NE));
    if (kind == KIND_BOND) return S ? nown : 0;
    if (!S) {
        int l = 0;
#pragma unroll
        for (int d = 0; d < 4; ++d)
            if (m & (2u << d)) { const int64_t j = bond_other_end(g, x, y, d); if (!(mk[j] & MASK_SITE)) ++l; }
        *lone = l;
        return 0;
    }
    int w = 1 + nown;
    const int xl = x > 0 ? x - 1 : (g.pbc ? g.m - 1 : -1), xr = x + 1 < g.m ? x + 1 : (g.pbc ? 0 : -1);
    const int64_t row = i - x;
    auto dangling = [&](int64_t j, unsigned bit) { const unsigned u = mk[j]; return (u & bit) && !(u & MASK_SITE); };
    if (xl >= 0 && dangling(row + xl, MASK_E)) ++w;
    if (y > 0 && dangling(i - g.m, MASK_N)) ++w;
    if (g.lattice == LAT_TRIANGULAR && y > 0) {
        if (xl >= 0 && dangling(row - g.m + xl, MASK_NE)) ++w;
        if (xr >= 0 && dangling(row - g.m + xr, MASK_NW)) ++w;
    }
    return w;
}

__global__ void __launch_bounds__(256)
inc_init_kernel(int64_t t, const uint8_t* __restrict__ mnew, uint8_t* __restrict__ mold, int32_t* __restrict__ label,
                int32_t* __restrict__ size)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t) return;
    const unsigned mo = mold[i], mn = mnew[i];
    if (mo & MASK_SITE) { if (label[i] == (int32_t)(i + 1)) mold[i] = (uint8_t)(mo | WASROOT); }
    else if (mn & MASK_SITE) { label[i] = (int32_t)(i + 1); size[i] = 0; }
}

__global__ void __launch_bounds__(256)
inc_unite_kernel(Geom g, const uint8_t* __restrict__ mnew, const uint8_t* __restrict__ mold, int32_t* __restrict__ label)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.t) return;
    const unsigned mn = mnew[i];
    if (!(mn & MASK_SITE) || !(mn & (MASK_E | MASK_N | MASK_NW | MASK_NE))) return;
    const unsigned mo = mold[i];
    const int x = (int)(i % g.m), y = (int)(i / g.m);
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        if (!(mn & (2u << d))) continue;
        const int64_t j = bond_other_end(g, x, y, d);
        if (!(mnew[j] & MASK_SITE)) continue;
        const bool before = (mo & MASK_SITE) && (mo & (2u << d)) && (mold[j] & MASK_SITE);
        if (!before) gl_unite(label, (int)i, (int)j);
    }
}

__global__ void __launch_bounds__(256)
inc_fold_kernel(Geom g, int kind, const uint8_t* __restrict__ mnew, const uint8_t* __restrict__ mold, int32_t* __restrict__ label,
                int32_t* __restrict__ size, Summary* __restrict__ sum)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    long long dl = 0;
    if (i < g.t) {
        const int x = (int)(i % g.m), y = (int)(i / g.m);
        int ln = 0, lo = 0;
        const int wn = inc_weight(g, kind, mnew, i, x, y, &ln), wo = inc_weight(g, kind, mold, i, x, y, &lo);
        dl = ln - lo;
        if (mnew[i] & MASK_SITE) {
            const int r = gl_find(label, (int)i);
            if ((mold[i] & WASROOT) && r != (int)i) atomicAdd(&size[r], size[i]);
            if (wn != wo) atomicAdd(&size[r], wn - wo);
            label[i] = r + 1;
        }
    }
    // lone bonds (mixed problem): the difference, summed over the block
    __shared__ long long s_dl;
    if (threadIdx.x == 0) s_dl = 0;
    __syncthreads();
    if (dl) atomicAdd((unsigned long long*)&s_dl, (unsigned long long)dl);
    __syncthreads();
    if (threadIdx.x == 0 && s_dl) atomicAdd(&sum->nlone, (unsigned long long)s_dl);
}

__global__ void __launch_bounds__(256)
inc_stats_kernel(int64_t t, const uint8_t* __restrict__ mnew, const int32_t* __restrict__ label, const int32_t* __restrict__ size,
                 Summary* __restrict__ sum)
{
    unsigned cnt = 0;
    unsigned long long best = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t; i += stride) {
        if (!(mnew[i] & MASK_SITE) || label[i] != (int32_t)(i + 1)) continue;
        const int32_t s = size[i];
        // (a node of the bond problem always owns or ends a bond; a root without weight cannot occur, but is not a cluster)
        if (s <= 0) continue;
        ++cnt;
        const unsigned long long pk = ((unsigned long long)s << 32) | (unsigned long long)(0xffffffffu - (unsigned)(i + 1));
        if (pk > best) best = pk;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    for (int o = 16; o; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
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

// ---- fast path: only BONDS were added (site bits of the mask unchanged, or the bond problem, where a site is a node iff
// one of its bonds is occupied): four sites per thread, 32-bit mask words; the weight of a new bond goes to whatever root
// its cluster has at that moment -- every root there can be is flagged (an old root or a new node), so inc_fold4 carries
// the sum on if that root loses its independence later in the same pass.
__global__ void __launch_bounds__(256)
inc_init4_kernel(int64_t nquad, const uint32_t* __restrict__ mnew4, uint32_t* __restrict__ mold4, int4* __restrict__ label4,
                 int32_t* __restrict__ size, Summary* __restrict__ sum)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nquad) return;
    const uint32_t mo = mold4[q], mn = mnew4[q];
    if (!((mo | mn) & 0x01010101u)) return;                      // no node among the four
    const int nnew = __popc(mn & ~mo & 0x01010101u);            // new nodes: clusters of their own until something unites them
    if (nnew) atomicAdd(&sum->ncl, (unsigned long long)nnew);
    int4 lab = label4[q];
    int l[4] = {lab.x, lab.y, lab.z, lab.w};
    uint32_t flags = 0;
    bool changed = false;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int64_t i = 4 * q + b;
        if ((mo >> (8 * b)) & 1u) { if (l[b] == (int32_t)(i + 1)) flags |= WASROOT << (8 * b); }
        else if ((mn >> (8 * b)) & 1u) { l[b] = (int32_t)(i + 1); size[i] = 0; flags |= WASROOT << (8 * b); changed = true; }
    }
    if (flags) mold4[q] = mo | flags;
    if (changed) label4[q] = make_int4(l[0], l[1], l[2], l[3]);
}

// Largest cluster without a recount: sizes only grow and a merged cluster is larger than its parts, so the previous maximum
// (carried in sum->maxpack) competes only with roots whose size entry received something in this pass; the LAST atomicAdd to
// an entry returns its final value minus the addend, so taking (returned + addend) of every add as a candidate includes the
// final size of every such root.  Candidates of a block are maximised in shared memory first.
__device__ __forceinline__ unsigned long long inc_pack(int size, int r)
{
    return ((unsigned long long)(unsigned)size << 32) | (unsigned long long)(0xffffffffu - (unsigned)(r + 1));
}

__global__ void __launch_bounds__(256)
inc_unite4_kernel(Geom g, int kind, int64_t nquad, const uint32_t* __restrict__ mnew4, const uint32_t* __restrict__ mold4,
                  const uint8_t* __restrict__ mnew, int32_t* __restrict__ label, int32_t* __restrict__ size, Summary* __restrict__ sum)
{
    // most new bonds of a block end up in ONE cluster (near the threshold: the spanning one): their weights are summed in
    // shared memory under the first root a thread of the block met, everything else goes to global memory directly
    __shared__ int s_root, s_cnt;
    __shared__ unsigned s_lone;
    __shared__ unsigned long long s_best;
    if (threadIdx.x == 0) { s_root = -1; s_cnt = 0; s_lone = 0; s_best = 0; }
    __syncthreads();
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t mn = 0, nb = 0;
    if (q < nquad) { mn = mnew4[q]; nb = mn & ~mold4[q] & 0x1e1e1e1eu; }      // bonds that are occupied now and were not
    unsigned lone = 0;
    while (nb) {
        const int bit = __ffs(nb) - 1;
        nb &= nb - 1;
        const int b = bit >> 3, d = (bit & 7) - 1;
        const int64_t i = 4 * q + b;
        const int x = (int)(i % g.m), y = (int)(i / g.m);
        const int64_t j = bond_other_end(g, x, y, d);
        const bool si = (mn >> (8 * b)) & 1u, sj = mnew[j] & MASK_SITE;
        if (si && sj) gl_unite(label, (int)i, (int)j);
        if (!si && !sj) { ++lone; continue; }                            // mixed problem: no occupied end
        // an owned bond of a node, or (mixed problem) a bond that dangles onto its other end
        const int r = gl_find(label, si ? (int)i : (int)j);
        const int first = atomicCAS(&s_root, -1, r);
        if (first == -1 || first == r) atomicAdd(&s_cnt, 1); else atomicMax(&s_best, inc_pack(atomicAdd(&size[r], 1) + 1, r));
    }
    if (lone) atomicAdd(&s_lone, lone);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long best = s_best;
        if (s_cnt) { const unsigned long long pk = inc_pack(atomicAdd(&size[s_root], s_cnt) + s_cnt, s_root); if (pk > best) best = pk; }
        if (best) atomicMax(&sum->maxpack, best);
        if (s_lone && kind == KIND_MIXED) atomicAdd(&sum->nlone, (unsigned long long)s_lone);
    }
}

__global__ void __launch_bounds__(256)
inc_fold4_kernel(int64_t nquad, const uint32_t* __restrict__ mnew4, const uint32_t* __restrict__ mold4, int32_t* __restrict__ label,
                 int32_t* __restrict__ size, Summary* __restrict__ sum)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned merged = 0;
    unsigned long long best = 0;
    if (q < nquad) {
        const uint32_t mn = mnew4[q];
        if (mn & 0x01010101u) {
            const uint32_t mo = mold4[q];
            const int4 lab = reinterpret_cast<const int4*>(label)[q];
            int l[4] = {lab.x, lab.y, lab.z, lab.w}, p[4];
            // first hop of the four sites together (after the last labeling every site points at its root: if that root is
            // still one, this is the only load; neighbours mostly share it; cached loads -- a value another SM has flattened
            // meanwhile is still an ancestor)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const bool act = (mn >> (8 * b)) & 1u;
                if (!act) p[b] = -1;
                else if (b > 0 && ((mn >> (8 * (b - 1))) & 1u) && l[b] == l[b - 1]) p[b] = p[b - 1];
                else p[b] = label[l[b] - 1] - 1;
            }
            bool changed = false;
            int rprev = -1, lprev = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                if (!((mn >> (8 * b)) & 1u)) { lprev = 0; continue; }
                const int i = (int)(4 * q + b);
                const int lb = l[b];
                const int r = lb == lprev ? rprev : (p[b] == lb - 1 ? p[b] : gl_find(label, p[b]));
                lprev = lb; rprev = r;
                if (((mo >> (8 * b)) & WASROOT) && r != i) {
                    const int si = size[i];
                    const unsigned long long pk = inc_pack(atomicAdd(&size[r], si) + si, r);
                    if (pk > best) best = pk;
                    ++merged;
                }
                if (lb != r + 1) { l[b] = r + 1; changed = true; }
            }
            if (changed) reinterpret_cast<int4*>(label)[q] = make_int4(l[0], l[1], l[2], l[3]);
        }
    }
    // every root that lost its independence is one cluster less
    merged = __reduce_add_sync(0xffffffffu, merged);
    for (int o = 16; o; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        if (other > best) best = other;
    }
    __shared__ unsigned s_m;
    __shared__ unsigned long long s_best;
    if (threadIdx.x == 0) { s_m = 0; s_best = 0; }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { if (merged) atomicAdd(&s_m, merged); if (best) atomicMax(&s_best, best); }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_m) atomicAdd(&sum->ncl, 0ull - (unsigned long long)s_m);
        if (s_best) atomicMax(&sum->maxpack, s_best);
    }
}

// number of clusters and the largest one: four sites per thread (t a multiple of 4)
__global__ void __launch_bounds__(256)
inc_stats4_kernel(int64_t nquad, const uint32_t* __restrict__ mnew4, const int4* __restrict__ label4, const int32_t* __restrict__ size,
                  Summary* __restrict__ sum)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned cnt = 0;
    unsigned long long best = 0;
    if (q < nquad) {
        const uint32_t mn = mnew4[q];
        if (mn & 0x01010101u) {
            const int4 lab = label4[q];
            const int l[4] = {lab.x, lab.y, lab.z, lab.w};
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int64_t i = 4 * q + b;
                if (!((mn >> (8 * b)) & 1u) || l[b] != (int32_t)(i + 1)) continue;
                const int32_t s = size[i];
                if (s <= 0) continue;
                ++cnt;
                const unsigned long long pk = ((unsigned long long)s << 32) | (unsigned long long)(0xffffffffu - (unsigned)(i + 1));
                if (pk > best) best = pk;
            }
        }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    for (int o = 16; o; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
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

// the fields of the summary the incremental pass recomputes (nlone is carried over and corrected by its difference)
// (fast path: the number of clusters and the largest cluster are carried over and corrected; else recounted)
__global__ void inc_reset_summary_kernel(Summary* sum, int keep_counts)
{
    if (!keep_counts) { sum->ncl = 0; sum->maxpack = 0; }
    sum->nocc_sites = 0; sum->nocc_bonds = 0; sum->span_best = 0;
    sum->nroots = 0; sum->nspan = 0; sum->span_overflow = 0;
}

unsigned nblk(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

}  // namespace

// can the labels on the handle be advanced to the current fill?  (same kind, same order / generator stream, fills not smaller)
bool ccl_incremental_applies(const Ctx* c, int kind)
{
    if (c->nranks != 1 || !c->lab_valid || c->lab_kind != kind || c->lab_epoch != c->occ_epoch) return false;
    if (c->site_src != c->lab_site_src || c->bond_src != c->lab_bond_src) return false;
    if ((c->site_src == SRC_PHILOX || c->bond_src == SRC_PHILOX) && (c->seed != c->lab_seed || c->stream_id != c->lab_stream)) return false;
    const bool sites = kind != KIND_BOND, bonds = kind != KIND_SITE;
    if (sites && c->ks < c->lab_ks) return false;
    if (bonds && c->kb < c->lab_kb) return false;
    return true;
}

int ccl_incremental_run(Ctx* c, int kind)
{
    const Geom& g = c->g;
    cudaStream_t st = c->stream;
    c->labeled = false;
    c->solved = false;
    c->lab_valid = false;                      // (set again by ccl_note_labeled when everything went through)
    if (!c->mask_prev) PERC_CUDA(cudaMalloc(&c->mask_prev, (size_t)g.t));
    { uint8_t* tmp = c->mask; c->mask = c->mask_prev; c->mask_prev = tmp; }      // mask_prev: the labeled fill; mask: rebuilt now
    // only bonds added (the usual sweep: Sq/bond_cond.f, Sq/sb_perc.f): four sites per thread; sites added: one site per thread
    const bool fast = (kind == KIND_BOND || (kind == KIND_MIXED && c->ks == c->lab_ks)) && g.m % 4 == 0;
    const int64_t nquad = g.t / 4;
    PERC_CUDA(cudaEventRecord(c->ev[0], st));
    inc_reset_summary_kernel<<<1, 1, 0, st>>>(c->d_sum, fast ? 1 : 0);
    int rc = occ_build_mask(c, kind);
    if (rc) return rc;
    PERC_CUDA(cudaEventRecord(c->ev[1], st));
    if (fast) {
        inc_init4_kernel<<<nblk(nquad), 256, 0, st>>>(nquad, (const uint32_t*)c->mask, (uint32_t*)c->mask_prev, (int4*)c->label, c->size, c->d_sum);
        inc_unite4_kernel<<<nblk(nquad), 256, 0, st>>>(g, kind, nquad, (const uint32_t*)c->mask, (const uint32_t*)c->mask_prev, c->mask,
                                                       c->label, c->size, c->d_sum);
    } else {
        inc_init_kernel<<<nblk(g.t), 256, 0, st>>>(g.t, c->mask, c->mask_prev, c->label, c->size);
        inc_unite_kernel<<<nblk(g.t), 256, 0, st>>>(g, c->mask, c->mask_prev, c->label);
    }
    PERC_CUDA(cudaEventRecord(c->ev[2], st));
    PERC_CUDA(cudaEventRecord(c->ev[3], st));
    if (fast) {
        inc_fold4_kernel<<<nblk(nquad), 256, 0, st>>>(nquad, (const uint32_t*)c->mask, (const uint32_t*)c->mask_prev, c->label, c->size, c->d_sum);
    } else {
        inc_fold_kernel<<<nblk(g.t), 256, 0, st>>>(g, kind, c->mask, c->mask_prev, c->label, c->size, c->d_sum);
        if (g.t % 4 == 0) inc_stats4_kernel<<<nblk(nquad), 256, 0, st>>>(nquad, (const uint32_t*)c->mask, (const int4*)c->label, c->size, c->d_sum);
        else inc_stats_kernel<<<148 * 8, 256, 0, st>>>(g.t, c->mask, c->label, c->size, c->d_sum);
    }
    PERC_CUDA(cudaEventRecord(c->ev[4], st));
    ccl_span_launch(c);
    c->launches += 6;
    PERC_CUDA(cudaEventRecord(c->ev[5], st));
    PERC_CUDA(cudaGetLastError());
    c->kind = kind;
    c->labeled = true;
    rc = ccl_fetch_summary(c);
    if (rc) return rc;
    c->h_span_gid.assign(c->h_span_ids.begin(), c->h_span_ids.end());
    c->h_span_total.assign(c->h_span_sizes.begin(), c->h_span_sizes.end());
    ccl_note_labeled(c, kind);
    return 0;
}

}  // namespace perc
