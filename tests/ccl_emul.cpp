// ccl_emul.cpp -- host emulation of the CUDA labeling pipeline (TEST INFRASTRUCTURE).
//
// Compiles percolation_b200/csrc/ccl_tile.cuh -- the very source the kernels are built from -- with
// g++ and executes every phase of every "thread block" thread by thread (a __syncthreads() becomes
// the end of a loop over tid).  Lets the CPU test-suite check the bit-parallel labeling logic
// against the oracle without a GPU.  Never used by the product path.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../percolation_b200/csrc/ccl_tile.cuh"

using namespace perc;

// mirror of build_mask_kernel (occupancy.cu): socc[t] by site, bocc[nb] by reference bond row
static void build_mask(const Geom& g, int kind, const uint8_t* socc, const uint8_t* bocc, std::vector<uint8_t>& mask)
{
    mask.assign((size_t)g.t, 0);
    std::vector<uint8_t> bbits((size_t)g.t, 0), touched((size_t)g.t, 0);
    if (kind != KIND_SITE) {
        for (int64_t r = 0; r < g.nb; ++r) {
            if (!bocc[r]) continue;
            int64_t a; int dir;
            ref_row_to_owner(g, r, &a, &dir);
            bbits[a] |= (uint8_t)(2u << dir);
            touched[a] = 1;
            touched[bond_other_end(g, (int)(a % g.m), (int)(a / g.m), dir)] = 1;
        }
    }
    for (int64_t i = 0; i < g.t; ++i) {
        int x = (int)(i % g.m), y = (int)(i / g.m);
        unsigned bits; bool site;
        if (kind == KIND_SITE) { bits = owned_bond_bits(g, x, y); site = socc[i]; }
        else if (kind == KIND_BOND) { bits = bbits[i]; site = touched[i]; }
        else { bits = bbits[i]; site = socc[i]; }
        mask[i] = (uint8_t)(bits | (site ? 1u : 0u));
    }
}

static int g_var = 0;          // variant of the tile kernel (ccl_tile.cuh, VAR)
extern "C" void ccl_emul_set_variant(int v) { g_var = v; }

template <int LAT, int KIND, int VAR>
static void run_local_v(const Geom& g, const uint8_t* mask, int32_t* label, int32_t* size, int32_t* rootlist, Summary* sum, bool vec)
{
    TileSmem* s = new TileSmem;
    std::vector<TileRegs> regs(CT_THREADS);
    for (int y0 = 0; y0 < g.n; y0 += CT_TH)
        for (int x0 = 0; x0 < g.m; x0 += CT_TW) {
            memset(s, 0xAB, sizeof(TileSmem));          // shared memory starts as garbage
            for (int tid = 0; tid < CT_THREADS; ++tid) tile_phase0<LAT, KIND>(*s, g, mask, x0, y0, tid, vec);
            for (int tid = 0; tid < CT_THREADS; ++tid) tile_phase1(*s, tid);
            for (int tid = CT_THREADS - 1; tid >= 0; --tid) tile_phase2_words(*s, tid);    // any order must work
#define LEVEL(K) for (int tid = CT_THREADS - 1; tid >= 0; --tid) tile_phase2_level<LAT, K>(*s, tid)
            LEVEL(1); LEVEL(2); LEVEL(3); LEVEL(4); LEVEL(5); LEVEL(6);
#undef LEVEL
            for (int tid = 0; tid < CT_THREADS; ++tid) tile_clear_ring(*s, tid);
            for (int tid = 0; tid < CT_THREADS; ++tid) { if (VAR == 2) tile_phase3_pair<LAT, KIND>(*s, g, x0, y0, tid, regs[tid]); else tile_phase3<LAT, KIND, VAR>(*s, g, x0, y0, tid, regs[tid]); }
            for (int tid = 0; tid < CT_THREADS; ++tid) tile_phase4_fill(*s, g, x0, y0, tid, regs[tid], size);
            tile_phase4_reserve(*s, sum);
            for (int tid = 0; tid < CT_THREADS; ++tid) tile_phase4_labels<VAR>(*s, g, x0, y0, tid, label, vec);
            for (int tid = 0; tid < CT_THREADS; ++tid) tile_phase4_roots(*s, g, x0, y0, tid, regs[tid], size, rootlist);
        }
    delete s;
}

template <int LAT, int KIND>
static void run_local(const Geom& g, const uint8_t* mask, int32_t* label, int32_t* size, int32_t* rootlist, Summary* sum, bool vec)
{
    if (g_var == 2) run_local_v<LAT, KIND, 2>(g, mask, label, size, rootlist, sum, vec);
    else if (g_var == 1) run_local_v<LAT, KIND, 1>(g, mask, label, size, rootlist, sum, vec);
    else run_local_v<LAT, KIND, 0>(g, mask, label, size, rootlist, sum, vec);
}

extern "C" int ccl_emul(int lattice, int m, int n, int pbc, int kind, const uint8_t* socc, const uint8_t* bocc,
                        int32_t* label, int32_t* csize, int64_t* out)
{
    Geom g = make_geom(lattice, m, n, pbc);
    std::vector<uint8_t> mask;
    build_mask(g, kind, socc, bocc, mask);
    std::vector<int32_t> size((size_t)g.t, 0x5a5a5a5a), rootlist((size_t)g.t, -1);
    for (int64_t i = 0; i < g.t; ++i) label[i] = 0x7f7f7f7f;
    Summary* sum = new Summary;
    memset(sum, 0, sizeof(Summary));
    const bool vec = (m % 16) == 0;
    const bool sq = lattice == LAT_SQUARE;
#define RUN(L, K) run_local<L, K>(g, mask.data(), label, size.data(), rootlist.data(), sum, vec)
    if (sq) { if (kind == KIND_SITE) RUN(LAT_SQUARE, KIND_SITE); else if (kind == KIND_BOND) RUN(LAT_SQUARE, KIND_BOND); else RUN(LAT_SQUARE, KIND_MIXED); }
    else    { if (kind == KIND_SITE) RUN(LAT_TRIANGULAR, KIND_SITE); else if (kind == KIND_BOND) RUN(LAT_TRIANGULAR, KIND_BOND); else RUN(LAT_TRIANGULAR, KIND_MIXED); }
#undef RUN
    int nrowb = (g.n - 1) / CT_TH, nwords = (g.m + 31) / 32, ncolb = (g.m - 1) / CT_TW + (g.pbc ? 1 : 0);
    int64_t nitems = (int64_t)nrowb * nwords * 32 + (int64_t)ncolb * g.n;
    for (int64_t id = nitems - 1; id >= 0; --id) {
        if (sq) merge_item<LAT_SQUARE>(g, mask.data(), label, nrowb, nwords, ncolb, id, vec);
        else merge_item<LAT_TRIANGULAR>(g, mask.data(), label, nrowb, nwords, ncolb, id, vec);
    }
    unsigned long long best = sum->maxpack, ncl = sum->ncl;       // closed clusters were counted by the tiles
    for (int64_t k = 0; k < sum->nroots; ++k) {
        int isroot;
        unsigned long long pk = rootfix_item(label, size.data(), rootlist.data(), k, true, &isroot);
        ncl += isroot;
        if (pk > best) best = pk;
    }
    for (int64_t i = g.t - 1; i >= 0; --i) label[i] = flatten_one(label, label[i]);
    for (int64_t i = 0; i < g.t; ++i) csize[i] = label[i] == (int32_t)(i + 1) ? size[i] : 0;
    out[0] = (int64_t)ncl;
    out[1] = (int64_t)sum->nlone;
    out[2] = (int64_t)(best >> 32);
    out[3] = best ? (int64_t)(0xffffffffu - (unsigned)(best & 0xffffffffu)) : 0;
    out[4] = (int64_t)sum->nroots;
    delete sum;
    return 0;
}


// Slab of a decomposed lattice (rank `rank` of `nranks`): the rank-local labeling with sizes restricted to
// the owned rows, exactly what the GPU does before the stitch.  socc / bocc describe the WHOLE lattice.
// label / csize: g.t = m * (rows held) entries; out[0..1] = ncl (closed clusters only), nlone; out[5..8] =
// y0, rows held, own_lo, own_hi.
extern "C" int ccl_emul_slab(int lattice, int m, int ng, int pbc, int kind, int nranks, int rank,
                             const uint8_t* socc, const uint8_t* bocc, int32_t* label, int32_t* csize, int64_t* out)
{
    Geom gw = make_geom(lattice, m, ng, pbc);
    std::vector<uint8_t> wmask;
    build_mask(gw, kind, socc, bocc, wmask);
    Geom g = make_slab_geom(lattice, m, ng, pbc, nranks, rank);
    const uint8_t* mask = wmask.data() + (int64_t)g.y0 * m;            // bond bits follow the WHOLE lattice's geometry
    std::vector<int32_t> size((size_t)g.t, 0x5a5a5a5a), rootlist((size_t)g.t, -1);
    for (int64_t i = 0; i < g.t; ++i) label[i] = 0x7f7f7f7f;
    Summary* sum = new Summary;
    memset(sum, 0, sizeof(Summary));
    const bool vec = (m % 16) == 0;
    const bool sq = lattice == LAT_SQUARE;
#define RUN(L, K) run_local<L, K>(g, mask, label, size.data(), rootlist.data(), sum, vec)
    if (sq) { if (kind == KIND_SITE) RUN(LAT_SQUARE, KIND_SITE); else if (kind == KIND_BOND) RUN(LAT_SQUARE, KIND_BOND); else RUN(LAT_SQUARE, KIND_MIXED); }
    else    { if (kind == KIND_SITE) RUN(LAT_TRIANGULAR, KIND_SITE); else if (kind == KIND_BOND) RUN(LAT_TRIANGULAR, KIND_BOND); else RUN(LAT_TRIANGULAR, KIND_MIXED); }
#undef RUN
    int nrowb = (g.n - 1) / CT_TH, nwords = (g.m + 31) / 32, ncolb = (g.m - 1) / CT_TW + (g.pbc ? 1 : 0);
    int64_t nitems = (int64_t)nrowb * nwords * 32 + (int64_t)ncolb * g.n;
    for (int64_t id = 0; id < nitems; ++id) {
        if (sq) merge_item<LAT_SQUARE>(g, mask, label, nrowb, nwords, ncolb, id, vec);
        else merge_item<LAT_TRIANGULAR>(g, mask, label, nrowb, nwords, ncolb, id, vec);
    }
    for (int64_t k = 0; k < sum->nroots; ++k) { int isroot; rootfix_item(label, size.data(), rootlist.data(), k, true, &isroot); }
    for (int64_t i = g.t - 1; i >= 0; --i) label[i] = flatten_one(label, label[i]);
    for (int64_t i = 0; i < g.t; ++i) csize[i] = label[i] == (int32_t)(i + 1) ? size[i] : 0;
    out[0] = (int64_t)sum->ncl; out[1] = (int64_t)sum->nlone;
    out[5] = g.y0; out[6] = g.n; out[7] = g.own_lo; out[8] = g.own_hi;
    delete sum;
    return 0;
}
