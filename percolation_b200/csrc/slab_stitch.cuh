// slab_stitch.cuh -- the union-find that stitches the rank-local labelings of a slab-decomposed lattice
// across the slab interfaces, written once for device and host (like ccl_tile.cuh).
//
// Input: the all-gathered interface blocks (slab.h, IfaceLayout).  Nodes are the distinct (rank, cluster)
// pairs seen on interface rows, kept in an open-addressing hash table; unions join the two labelings of
// every shared row; a class gets the smallest member id as lattice-wide label and the sum of the rank-local
// sizes.  Every rank runs the same stitch redundantly (no second exchange).  slab.cu wraps the item
// functions into kernels (one thread per item); perc_stitch_host runs the SAME functions item by item on the
// host, which is how the CPU tests check this logic against the oracle.
#pragma once
#include <stdint.h>
#include "geometry.cuh"
#include "slab.h"

#ifdef __CUDA_ARCH__
#define PERC_SDEV 1
#else
#define PERC_SDEV 0
#endif

namespace perc {

struct StitchTab {
    int64_t* keys;                 // [H] node key = (rank << 40) | lattice-wide id of the rank-local root; 0 = empty
    int64_t* hsize;                // [H] rank-local size of the node's cluster (owned rows only)
    int32_t* parent;               // [H] union-find over slots (the root is the smallest slot)
    unsigned long long* cgid;      // [H] at class roots: lattice-wide label = smallest member id
    long long* ctot;               // [H] at class roots: size of the class
    int32_t* cpos;                 // [H] at class roots: members with a non-zero rank-local size
    unsigned long long* rep;       // [H] at class roots: smallest member id that belongs to the CALLING rank
    int32_t* flag;                 // [H] at class roots: already reported as spanning
    long long* out;                // [8] 0 ncl adjustment  1 max size  2 its label  3 nspan  4 npairs  5 nclasses(mine)  6 error
    int64_t* span;                 // [2 * MAX_SPAN_CLASSES] (label, size) of the spanning classes
    int64_t* pairs;                // [4 * cap] (root id, representative id, class label, class size) of the calling rank's nodes
    int H, logH, cap, pad;
};
constexpr int MAX_SPAN_CLASSES = 4096;
constexpr unsigned long long STITCH_NONE = ~0ull;

PERC_HD int64_t stitch_key(int rank, int64_t gid) { return ((int64_t)rank << 40) | gid; }

// ---- atomics with a sequential host twin ------------------------------------------------------
PERC_HD int64_t st_cas64(int64_t* a, int64_t cmp, int64_t val)
{
#if PERC_SDEV
    return (int64_t)atomicCAS((unsigned long long*)a, (unsigned long long)cmp, (unsigned long long)val);
#else
    int64_t old = *a; if (old == cmp) *a = val; return old;
#endif
}
PERC_HD void st_min64(unsigned long long* a, unsigned long long v)
{
#if PERC_SDEV
    atomicMin(a, v);
#else
    if (v < *a) *a = v;
#endif
}
PERC_HD void st_max64(long long* a, long long v)
{
#if PERC_SDEV
    atomicMax(a, v);
#else
    if (v > *a) *a = v;
#endif
}
PERC_HD long long st_add64(long long* a, long long v)
{
#if PERC_SDEV
    return (long long)atomicAdd((unsigned long long*)a, (unsigned long long)v);
#else
    long long old = *a; *a += v; return old;
#endif
}
PERC_HD int st_add32(int32_t* a, int v)
{
#if PERC_SDEV
    return atomicAdd(a, v);
#else
    int old = *a; *a += v; return old;
#endif
}
PERC_HD int st_min32(int32_t* a, int v)
{
#if PERC_SDEV
    return atomicMin(a, v);
#else
    int old = *a; if (v < old) *a = v; return old;
#endif
}
PERC_HD int st_exch32(int32_t* a, int v)
{
#if PERC_SDEV
    return atomicExch(a, v);
#else
    int old = *a; *a = v; return old;
#endif
}
PERC_HD int32_t st_ld32(const int32_t* a)
{
#if PERC_SDEV
    return __ldcg(a);
#else
    return *a;
#endif
}

PERC_HD int stitch_hash(const StitchTab& T, int64_t key)
{
    return (int)(((unsigned long long)key * 0x9E3779B97F4A7C15ull) >> (64 - T.logH));
}

// slot of `key`; inserts it (with its size) when absent
PERC_HD int stitch_insert(const StitchTab& T, int64_t key, int64_t size)
{
    int s = stitch_hash(T, key);
    for (;;) {
        int64_t old = st_cas64(&T.keys[s], 0, key);
        if (old == 0) { T.hsize[s] = size; return s; }
        if (old == key) return s;
        s = (s + 1) & (T.H - 1);
    }
}

PERC_HD int stitch_lookup(const StitchTab& T, int64_t key)
{
    int s = stitch_hash(T, key);
    for (;;) {
        int64_t k = T.keys[s];
        if (k == key) return s;
        if (k == 0) return -1;
        s = (s + 1) & (T.H - 1);
    }
}

PERC_HD int stitch_find(const StitchTab& T, int a)
{
    int p;
    while ((p = st_ld32(&T.parent[a])) != a) a = p;
    return a;
}

PERC_HD void stitch_unite(const StitchTab& T, int a, int b)
{
    for (;;) {
        a = stitch_find(T, a);
        b = stitch_find(T, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        int old = st_min32(&T.parent[a], b);
        if (old == a) return;
        a = old;
    }
}

// ---- the phases; `e` / `s` is the item (thread) index --------------------------------------------
// phase 0: clear slot s
PERC_HD void stitch_clear(const StitchTab& T, int s)
{
    T.keys[s] = 0; T.hsize[s] = 0; T.parent[s] = s; T.cgid[s] = STITCH_NONE; T.ctot[s] = 0; T.cpos[s] = 0;
    T.rep[s] = STITCH_NONE; T.flag[s] = 0;
    if (s < 8) T.out[s] = 0;
}

// phase 1: interface entry e = (rank r, side, column x) -> node.  side 0 = first owned row (ranks >= 1),
// side 1 = top halo row (ranks < G-1).  Neighbouring columns mostly repeat the label: only changes insert.
PERC_HD void stitch_nodes(const StitchTab& T, int nranks, int m, const int64_t* __restrict__ gathered, int64_t e)
{
    IfaceLayout L{m};
    const int x = (int)(e % m), side = (int)((e / m) & 1), r = (int)(e / (2 * (int64_t)m));
    if (r >= nranks || (side == 0 && r == 0) || (side == 1 && r == nranks - 1)) return;
    const int64_t* blk = gathered + r * L.words();
    const int64_t* ids = blk + (side ? L.rowB() : L.rowA());
    const int64_t* sz = blk + (side ? L.sizeB() : L.sizeA());
    if (!ids[x] || (x > 0 && ids[x] == ids[x - 1])) return;
    stitch_insert(T, stitch_key(r, ids[x]), sz[x]);
}

// phase 2: row yb of interface r | r+1 carries two labelings: union them, column by column
PERC_HD void stitch_unions(const StitchTab& T, int nranks, int m, const int64_t* __restrict__ gathered, int64_t e)
{
    IfaceLayout L{m};
    const int x = (int)(e % m), r = (int)(e / m);
    if (r + 1 >= nranks) return;
    const int64_t* lo = gathered + r * L.words() + L.rowB();
    const int64_t* hi = gathered + (r + 1) * L.words() + L.rowA();
    if ((lo[x] != 0) != (hi[x] != 0)) { T.out[6] = 1; return; }        // the ranks disagree on the occupancy
    if (!lo[x]) return;
    if (x > 0 && lo[x] == lo[x - 1] && hi[x] == hi[x - 1]) return;       // same pair as the column before
    const int a = stitch_lookup(T, stitch_key(r, lo[x])), b = stitch_lookup(T, stitch_key(r + 1, hi[x]));
    if (a < 0 || b < 0) { T.out[6] = 2; return; }
    stitch_unite(T, a, b);
}

// phase 3: every node folds into its class root
PERC_HD void stitch_classes(const StitchTab& T, int rank, int s)
{
    const int64_t key = T.keys[s];
    if (!key) return;
    const int c = stitch_find(T, s);
    const unsigned long long gid = (unsigned long long)(key & (((int64_t)1 << 40) - 1));
    st_min64(&T.cgid[c], gid);
    st_add64(&T.ctot[c], T.hsize[s]);
    if (T.hsize[s] > 0) st_add32(&T.cpos[c], 1);
    if ((int)(key >> 40) == rank) st_min64(&T.rep[c], gid);
}

// phase 4a: class roots -> cluster-count adjustment (the members were counted once per rank), largest class
PERC_HD void stitch_summary_a(const StitchTab& T, int s)
{
    if (!T.keys[s] || T.parent[s] != s) return;
    st_add64(&T.out[0], -(long long)(T.cpos[s] - (T.ctot[s] > 0 ? 1 : 0)));
    st_max64(&T.out[1], T.ctot[s]);
}
// phase 4a': the ranks' own largest clusters (item r); an interface cluster is superseded by its class
PERC_HD void stitch_summary_rank(const StitchTab& T, int m, const int64_t* __restrict__ gathered, int r)
{
    IfaceLayout L{m};
    const int64_t* sc = gathered + r * L.words() + L.scalars();
    if (sc[3] && stitch_lookup(T, stitch_key(r, sc[3])) < 0) st_max64(&T.out[1], sc[2]);
}
// phase 4b: smallest label among the clusters of the largest size (out[2] starts at "none" = 0 -> use min over candidates)
PERC_HD void stitch_summary_b(const StitchTab& T, int s, unsigned long long* maxgid)
{
    if (!T.keys[s] || T.parent[s] != s) return;
    if (T.ctot[s] == T.out[1] && T.ctot[s] > 0) st_min64(maxgid, T.cgid[s]);
}
PERC_HD void stitch_summary_rank_b(const StitchTab& T, int m, const int64_t* __restrict__ gathered, int r, unsigned long long* maxgid)
{
    IfaceLayout L{m};
    const int64_t* sc = gathered + r * L.words() + L.scalars();
    if (sc[3] && sc[2] == T.out[1] && sc[2] > 0 && stitch_lookup(T, stitch_key(r, sc[3])) < 0) st_min64(maxgid, (unsigned long long)sc[3]);
}

// phase 5: spanning = the class reaches row 0 (label <= m) and a site of the lattice's top row (item x)
PERC_HD void stitch_span(const StitchTab& T, int nranks, int m, const int64_t* __restrict__ gathered, int x)
{
    IfaceLayout L{m};
    const int64_t* top = gathered + (nranks - 1) * L.words() + L.rowTop();
    if (!top[x] || (x > 0 && top[x] == top[x - 1])) return;
    const int k = stitch_lookup(T, stitch_key(nranks - 1, top[x]));
    if (k < 0) return;
    const int c = stitch_find(T, k);
    if (T.cgid[c] > (unsigned long long)m) return;
    if (st_exch32(&T.flag[c], 1)) return;
    const long long pos = st_add64(&T.out[3], 1);
    if (pos < MAX_SPAN_CLASSES) { T.span[2 * pos] = (int64_t)T.cgid[c]; T.span[2 * pos + 1] = T.ctot[c]; }
}

// phase 6: the calling rank's nodes -> (root id, representative id, class label, class size); on the device the
// root's label entry is redirected to the representative right here (label = rank-local table, off = y0 * m)
PERC_HD void stitch_mine(const StitchTab& T, int rank, int s, int32_t* label, int64_t off)
{
    const int64_t key = T.keys[s];
    if (!key || (int)(key >> 40) != rank) return;
    const int c = stitch_find(T, s);
    const int64_t gid = key & (((int64_t)1 << 40) - 1), rep = (int64_t)T.rep[c];
    const long long pos = st_add64(&T.out[4], 1);
    if (pos < T.cap) {
        T.pairs[4 * pos] = gid; T.pairs[4 * pos + 1] = rep; T.pairs[4 * pos + 2] = (int64_t)T.cgid[c]; T.pairs[4 * pos + 3] = T.ctot[c];
    }
    if (label && gid != rep) label[gid - off - 1] = (int32_t)(rep - off);      // root -> representative (+1 encoding)
}

}  // namespace perc
