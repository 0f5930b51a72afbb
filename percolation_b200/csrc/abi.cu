// abi.cu -- extern "C" entry points of libperc_b200.so (see include/perc_abi.h).
// Thin argument checking + dispatch; all compute is in the CUDA kernels of the sibling files.
#include <cstdio>
#include <cstring>
#include <string>
#include <mutex>
#include <unordered_map>
#include "../../include/perc_abi.h"
#include "context.h"

namespace perc {

static std::mutex g_mu;
static std::unordered_map<int64_t, Ctx*> g_ctx;
static int64_t g_next = 1;

static Ctx* lookup(const int64_t* h)
{
    if (!h) return nullptr;
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_ctx.find(*h);
    return it == g_ctx.end() ? nullptr : it->second;
}

void* ctx_dev_stage(Ctx* c, size_t bytes)
{
    if (bytes > c->d_stage_bytes) {
        if (c->d_stage) { cudaStreamSynchronize(c->stream); cudaFree(c->d_stage); c->d_stage = nullptr; c->d_stage_bytes = 0; }
        size_t want = bytes + (bytes >> 3) + 256;
        if (cudaMalloc(&c->d_stage, want) != cudaSuccess) return nullptr;
        c->d_stage_bytes = want;
    }
    return c->d_stage;
}

// rank tables exist only once a caller supplies an order / flags (the generator needs none)
int ctx_ensure_ranks(Ctx* c, bool sites, bool bonds)
{
    const int64_t t = c->g.t;
    if (c->nranks > 1) return PERC_E_STATE;              // slab handles take their occupancy from the generator
    if (sites && !c->srank) {
        PERC_CUDA(cudaMalloc(&c->srank, sizeof(int32_t) * t));
        PERC_CUDA(cudaMemsetAsync(c->srank, 0x7f, sizeof(int32_t) * t, c->stream));
    }
    if (bonds && !c->brank) {
        PERC_CUDA(cudaMalloc(&c->brank, sizeof(int32_t) * t * c->g.ndir));
        PERC_CUDA(cudaMemsetAsync(c->brank, 0x7f, sizeof(int32_t) * t * c->g.ndir, c->stream));
    }
    return 0;
}

void* ctx_host_stage(Ctx* c, size_t bytes)
{
    if (bytes > c->h_stage_bytes) {
        if (c->h_stage) { cudaFreeHost(c->h_stage); c->h_stage = nullptr; c->h_stage_bytes = 0; }
        if (cudaMallocHost(&c->h_stage, bytes) != cudaSuccess) return nullptr;
        c->h_stage_bytes = bytes;
    }
    return c->h_stage;
}

int ctx_alloc(Ctx* c)
{
    const int64_t t = c->g.t;
    PERC_CUDA(cudaSetDevice(c->device));
    PERC_CUDA(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, c->device));
    PERC_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    PERC_CUDA(cudaMalloc(&c->mask, t));
    PERC_CUDA(cudaMalloc(&c->label, sizeof(int32_t) * t));
    PERC_CUDA(cudaMalloc(&c->size, sizeof(int32_t) * t));
    PERC_CUDA(cudaMalloc(&c->rootlist, sizeof(int32_t) * t));
    PERC_CUDA(cudaMalloc(&c->d_sum, sizeof(Summary)));
    PERC_CUDA(cudaMallocHost(&c->h_sum_pin, sizeof(Summary)));
    PERC_CUDA(cudaMemsetAsync(c->d_sum, 0, sizeof(Summary), c->stream));
    PERC_CUDA(cudaMalloc(&c->d_pcg, sizeof(PcgState)));
    PERC_CUDA(cudaMallocHost(&c->h_pcg, sizeof(PcgState)));
    PERC_CUDA(cudaMalloc(&c->d_hist, sizeof(unsigned long long) * (4096 + 16)));
    PERC_CUDA(cudaMalloc(&c->d_thr, sizeof(PhiloxThreshold) * 2));
    c->cand_cap = 8192;
    PERC_CUDA(cudaMalloc(&c->d_cand, sizeof(unsigned long long) * 2 * c->cand_cap));
    for (auto& e : c->ev) PERC_CUDA(cudaEventCreate(&e));
    PERC_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

static int ensure_pcg(Ctx* c)
{
    if (c->vx) return 0;
    const int64_t t = c->g.t;
    PERC_CUDA(cudaMalloc(&c->cfull, t));
    PERC_CUDA(cudaMalloc(&c->vx, sizeof(double) * t));
    PERC_CUDA(cudaMalloc(&c->vr, sizeof(double) * t));
    PERC_CUDA(cudaMalloc(&c->vp, sizeof(double) * t));
    PERC_CUDA(cudaMalloc(&c->vp2, sizeof(double) * t));
    if (c->g.m % 16) PERC_CUDA(cudaMalloc(&c->vq, sizeof(double) * t));    // only the scalar fallback stores q = A p
    return 0;
}

void ctx_free(Ctx* c)
{
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (Ctx* k : c->batch_kids) { ctx_free(k); delete k; }
    c->batch_kids.clear();
    slab_comm_destroy(c);
    if (c->d_iface) cudaFree(c->d_iface);
    if (c->d_stitch) cudaFree(c->d_stitch);
    if (c->h_iface) cudaFreeHost(c->h_iface);
    void* ptrs[] = {c->srank, c->brank, c->mask, c->label, c->size, c->rootlist, c->d_sum, c->d_pcg,
                    c->d_hist, c->d_thr, c->d_cand, c->cfull, c->vx, c->vr, c->vp, c->vp2, c->vq, c->xprow, c->partial, c->d_stage, c->d_defl, c->bond_w, c->wplane, c->wdiag, c->d_sched, c->mask_prev};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (c->h_pcg) cudaFreeHost(c->h_pcg);
    if (c->h_sum_pin) cudaFreeHost(c->h_sum_pin);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->h_defl) cudaFreeHost(c->h_defl);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
}

static int check_geom(int lattice, int m, int n, int pbc)
{
    if (lattice != LAT_SQUARE && lattice != LAT_TRIANGULAR) return PERC_E_ARG;
    if (m < 2 || n < 2 || (pbc != 0 && pbc != 1)) return PERC_E_ARG;
    if (pbc && m < 3) return PERC_E_ARG;
    if (lattice == LAT_TRIANGULAR && (m & 1)) return PERC_E_ODD_M;
    if ((int64_t)m * n > 0x7ffffff0LL / 4) return PERC_E_SIZE;     // int32 labels and bond rows
    return 0;
}

static int set_fill(Ctx* c, int64_t ks, int64_t kb)
{
    if (ks >= 0) {
        if (ks > c->g.tg) return PERC_E_ARG;
        if (c->site_src == SRC_PHILOX) { int rc = occ_generate(c, c->seed, c->stream_id, ks, -1); if (rc) return rc; }
        c->ks = ks;
    }
    if (kb >= 0) {
        if (kb > c->g.nb) return PERC_E_ARG;
        if (c->bond_src == SRC_PHILOX) { int rc = occ_generate(c, c->seed, c->stream_id, -1, kb); if (rc) return rc; }
        c->kb = kb;
    }
    c->labeled = false;
    return 0;
}

static int download(Ctx* c, void* dst, const void* src, size_t bytes)
{
    PERC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    PERC_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

}  // namespace perc

using namespace perc;

#define GET_CTX(h)                                   \
    Ctx* c = lookup(h);                              \
    if (!c) return PERC_E_HANDLE;                    \
    { cudaError_t e__ = cudaSetDevice(c->device); if (e__ != cudaSuccess) return (int)e__; }

extern "C" {

int32_t perc_geom_nb(const int32_t* lattice, const int32_t* m, const int32_t* n, const int32_t* pbc, int32_t* nb)
{
    if (!lattice || !m || !n || !pbc || !nb) return PERC_E_ARG;
    int rc = check_geom(*lattice, *m, *n, *pbc);
    if (rc) return rc;
    *nb = (int32_t)make_geom(*lattice, *m, *n, *pbc).nb;
    return 0;
}

int32_t perc_geom_bondlist(const int32_t* lattice, const int32_t* m, const int32_t* n, const int32_t* pbc, int32_t* b)
{
    if (!lattice || !m || !n || !pbc || !b) return PERC_E_ARG;
    int rc = check_geom(*lattice, *m, *n, *pbc);
    if (rc) return rc;
    Geom g = make_geom(*lattice, *m, *n, *pbc);
    for (int64_t r = 0; r < g.nb; ++r) {
        int64_t a; int dir;
        ref_row_to_owner(g, r, &a, &dir);
        int64_t o = bond_other_end(g, (int)(a % g.m), (int)(a / g.m), dir);
        int64_t lo = a < o ? a : o, hi = a < o ? o : a;
        b[r] = (int32_t)lo + 1;
        b[g.nb + r] = (int32_t)hi + 1;
    }
    return 0;
}

int32_t perc_geom_nearestn(const int32_t* lattice, const int32_t* m, const int32_t* n, const int32_t* pbc,
                           const int32_t* rn, int32_t* nn)
{
    if (!lattice || !m || !n || !pbc || !rn || !nn) return PERC_E_ARG;
    int rc = check_geom(*lattice, *m, *n, *pbc);
    if (rc) return rc;
    Geom g = make_geom(*lattice, *m, *n, *pbc);
    if (*rn < 1 || *rn > g.t) return PERC_E_ARG;
    int64_t i = *rn - 1;
    int x = (int)(i % g.m), y = (int)(i / g.m);
    unsigned ex = neighbour_bits(g, x, y);
    int xl = x > 0 ? x - 1 : g.m - 1, xr = x + 1 < g.m ? x + 1 : 0;
    int64_t row = i - x;
    for (int k = 0; k < 6; ++k) nn[k] = 0;
    // nearestn order: ascending site number for in-lattice neighbours, periodic wraps appended last
    int64_t cand[8]; int cnt = 0;
    struct { unsigned bit; int64_t j; bool wrap; } dirs[8] = {
        {NB_SW, row - g.m + xl, false}, {NB_S, i - g.m, false}, {NB_SE, row - g.m + xr, x + 1 == g.m},
        {NB_W, row + xl, x == 0}, {NB_E, row + xr, x + 1 == g.m},
        {NB_NW, row + g.m + xl, x == 0}, {NB_N, i + g.m, false}, {NB_NE, row + g.m + xr, false}};
    for (int pass = 0; pass < 2; ++pass)
        for (int k = 0; k < 8; ++k)
            if ((ex & dirs[k].bit) && dirs[k].wrap == (pass == 1)) cand[cnt++] = dirs[k].j;
    // wrapped neighbours are listed in ascending site number too
    for (int k = 0; k < cnt && k < 6; ++k) nn[k] = (int32_t)cand[k] + 1;
    return 0;
}

int32_t perc_create(int64_t* h, const int32_t* lattice, const int32_t* m, const int32_t* n,
                    const int32_t* pbc, const int32_t* device)
{
    if (!h || !lattice || !m || !n || !pbc || !device) return PERC_E_ARG;
    int rc = check_geom(*lattice, *m, *n, *pbc);
    if (rc) return rc;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) return (int)e;                 // no CPU fallback: fail loudly
    if (*device < 0 || *device >= ndev) return PERC_E_ARG;
    Ctx* c = new Ctx();
    c->g = make_geom(*lattice, *m, *n, *pbc);
    c->device = *device;
    rc = ctx_alloc(c);
    if (rc) { ctx_free(c); delete c; return rc; }
    std::lock_guard<std::mutex> lk(g_mu);
    *h = g_next++;
    g_ctx[*h] = c;
    return 0;
}

int32_t perc_destroy(const int64_t* h)
{
    if (!h) return PERC_E_ARG;
    Ctx* c = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_ctx.find(*h);
        if (it == g_ctx.end()) return PERC_E_HANDLE;
        c = it->second;
        g_ctx.erase(it);
    }
    ctx_free(c);
    delete c;
    return 0;
}

int32_t perc_sync(const int64_t* h)
{
    GET_CTX(h);
    return (int)cudaStreamSynchronize(c->stream);
}

int32_t perc_set_site_order(const int64_t* h, const int32_t* order)
{
    GET_CTX(h);
    if (!order) return PERC_E_ARG;
    return occ_upload_site_order(c, order);
}

int32_t perc_set_bond_order(const int64_t* h, const int32_t* border)
{
    GET_CTX(h);
    if (!border) return PERC_E_ARG;
    return occ_upload_bond_order(c, border);
}

int32_t perc_set_fill(const int64_t* h, const int32_t* ks, const int32_t* kb)
{
    GET_CTX(h);
    if (!ks || !kb) return PERC_E_ARG;
    return set_fill(c, *ks, *kb);
}

int32_t perc_set_occupancy(const int64_t* h, const uint8_t* socc, const uint8_t* bocc)
{
    GET_CTX(h);
    return occ_upload_flags(c, socc, bocc);
}

int32_t perc_generate(const int64_t* h, const int64_t* seed, const int64_t* stream, const int32_t* ks, const int32_t* kb)
{
    GET_CTX(h);
    if (!seed || !stream || !ks || !kb) return PERC_E_ARG;
    return occ_generate(c, (unsigned long long)*seed, (unsigned long long)*stream, *ks, *kb);
}

int32_t perc_get_occupancy(const int64_t* h, uint8_t* socc, uint8_t* bocc)
{
    GET_CTX(h);
    return occ_export(c, socc, bocc);
}

int32_t perc_label(const int64_t* h, const int32_t* kind)
{
    GET_CTX(h);
    if (!kind || *kind < KIND_SITE || *kind > KIND_MIXED) return PERC_E_ARG;
    if ((*kind == KIND_SITE || *kind == KIND_MIXED) && c->site_src == SRC_NONE) return PERC_E_STATE;
    if ((*kind == KIND_BOND || *kind == KIND_MIXED) && c->bond_src == SRC_NONE) return PERC_E_STATE;
    return ccl_run(c, *kind);
}

// re-labeling along a sweep: when the handle holds the labels of a smaller fill of the same order / generator stream, only the
// added elements are united (csrc/ccl_incremental.cu); otherwise the full pass of perc_label runs.  *incremental tells which.
int32_t perc_label_incremental(const int64_t* h, const int32_t* kind, int32_t* incremental)
{
    GET_CTX(h);
    if (!kind || *kind < KIND_SITE || *kind > KIND_MIXED) return PERC_E_ARG;
    if ((*kind == KIND_SITE || *kind == KIND_MIXED) && c->site_src == SRC_NONE) return PERC_E_STATE;
    if ((*kind == KIND_BOND || *kind == KIND_MIXED) && c->bond_src == SRC_NONE) return PERC_E_STATE;
    const bool inc = ccl_incremental_applies(c, *kind);
    if (incremental) *incremental = inc ? 1 : 0;
    return inc ? ccl_incremental_run(c, *kind) : ccl_run(c, *kind);
}

int32_t perc_summary(const int64_t* h, int64_t* ncl, int32_t* maxcs, int32_t* maxcn, int32_t* nspan)
{
    GET_CTX(h);
    if (!c->labeled) return PERC_E_STATE;
    if (c->nranks > 1) return PERC_E_STATE;              // lattice-wide values need 64 bits: perc_summary_i8
    const Summary& s = c->h_sum;
    int32_t ms = (int32_t)(s.maxpack >> 32);
    int32_t mn = ms ? (int32_t)(0xffffffffu - (unsigned)(s.maxpack & 0xffffffffu)) : 0;
    if (ms == 0 && s.nlone > 0) { ms = 1; mn = 0; }     // only lone bonds: size-1 clusters
    if (ncl) *ncl = (int64_t)(s.ncl + s.nlone);
    if (maxcs) *maxcs = ms;
    if (maxcn) *maxcn = mn;
    if (nspan) *nspan = s.nspan;
    return 0;
}

int32_t perc_get_site_labels(const int64_t* h, int32_t* s)
{
    GET_CTX(h);
    if (!s) return PERC_E_ARG;
    if (!c->labeled) return PERC_E_STATE;
    if (c->kind == KIND_BOND) { std::memset(s, 0, sizeof(int32_t) * c->g.t); return 0; }   // bond problem has no s()
    return download(c, s, c->label, sizeof(int32_t) * c->g.t);
}

int32_t perc_get_bond_labels(const int64_t* h, int32_t* b3)
{
    GET_CTX(h);
    if (!b3) return PERC_E_ARG;
    if (!c->labeled) return PERC_E_STATE;
    return ccl_export_bond_labels(c, b3);
}

int32_t perc_get_sizes(const int64_t* h, int32_t* cs)
{
    GET_CTX(h);
    if (!cs) return PERC_E_ARG;
    if (!c->labeled) return PERC_E_STATE;
    return ccl_export_sizes(c, cs);
}

// ---- the reference programs' output files, in their own formats -----------------------------
// which = 1 `site.txt`   : j, s(j), c(j)                      format(i10,",",i10,",",i10)            Sq/site.f:354-359
//         2 `bond.txt`   : b(j,1), b(j,2), b(j,3), j, c(j)    format(i10,",",i10,",",i10,",",i10,",",i10)  Sq/bond.f:443-448
//         3 `sbsite.txt` : i, s(i), c(i)                                                               Sq/sitebond.f:469-471
//         4 `sbbond.txt` : b(i,1), b(i,2), b(i,3)                                                      Sq/sitebond.f:473-475
//         5 `bondlist.txt`: blist(i,1), blist(i,2)            format(i10,",",i10)                      Sq/site.f:106-120
// Labels are the library's canonical ones (smallest member site id), c(.) is indexed by them (a label beyond the table
// -- a lone bond of the mixed problem -- has size 1 by construction); MATLAB/ConductCalc.m and the *Plot.m scripts only
// compare labels with `perccln`, so they read these files unchanged.  path(pathlen): not NUL-terminated (Fortran character).
int32_t perc_write_txt(const int64_t* h, const int32_t* which, const char* path, const int32_t* pathlen)
{
    GET_CTX(h);
    if (!which || !path || !pathlen || *pathlen < 1 || *pathlen > 4096 || *which < 1 || *which > 5) return PERC_E_ARG;
    if (c->nranks > 1) return PERC_E_STATE;
    const Geom& g = c->g;
    const int w = *which;
    if (w != 5 && !c->labeled) return PERC_E_STATE;
    if ((w == 1 && c->kind != KIND_SITE) || (w == 2 && c->kind != KIND_BOND) || ((w == 3 || w == 4) && c->kind != KIND_MIXED)) return PERC_E_STATE;
    std::vector<int32_t> s, b3, cs, bl;
    int rc = 0;
    if (w == 1 || w == 3) { s.resize((size_t)g.t); rc = download(c, s.data(), c->label, sizeof(int32_t) * g.t); if (rc) return rc; }
    if (w == 2 || w == 4) { b3.resize((size_t)g.nb); rc = ccl_export_bond_labels(c, b3.data()); if (rc) return rc; }
    if (w <= 3) { cs.resize((size_t)g.t); rc = ccl_export_sizes(c, cs.data()); if (rc) return rc; }
    if (w == 2 || w == 4 || w == 5) {
        bl.resize(2 * (size_t)g.nb);
        const int32_t lat = g.lattice, m = g.m, n = g.n, pbc = g.pbc;
        rc = perc_geom_bondlist(&lat, &m, &n, &pbc, bl.data()); if (rc) return rc;
    }
    const std::string fn(path, (size_t)*pathlen);
    FILE* f = std::fopen(fn.c_str(), "w");
    if (!f) return PERC_E_ARG;
    if (w == 1 || w == 3)
        for (int64_t j = 0; j < g.t; ++j) std::fprintf(f, "%10d,%10d,%10d\n", (int)(j + 1), s[j], cs[j]);
    else if (w == 2)
        for (int64_t j = 0; j < g.nb; ++j)
            std::fprintf(f, "%10d,%10d,%10d,%10d,%10d\n", bl[j], bl[g.nb + j], b3[j], (int)(j + 1), j < g.t ? cs[j] : 0);
    else if (w == 4)
        for (int64_t j = 0; j < g.nb; ++j) std::fprintf(f, "%10d,%10d,%10d\n", bl[j], bl[g.nb + j], b3[j]);
    else
        for (int64_t j = 0; j < g.nb; ++j) std::fprintf(f, "%10d,%10d\n", bl[j], bl[g.nb + j]);
    const bool bad = std::ferror(f) != 0;
    if (std::fclose(f) != 0 || bad) return PERC_E_ARG;
    return 0;
}

int32_t perc_span(const int64_t* h, const int32_t* max_ids, int32_t* nspan, int32_t* ids, int32_t* sizes)
{
    GET_CTX(h);
    if (!c->labeled || !max_ids || !nspan) return PERC_E_STATE;
    *nspan = c->h_sum.nspan;
    int k = (int)c->h_span_ids.size();
    if (k > *max_ids) k = *max_ids;
    for (int j = 0; j < k; ++j) { if (ids) ids[j] = c->h_span_ids[j]; if (sizes) sizes[j] = c->h_span_sizes[j]; }
    return 0;
}

int32_t perc_hist(const int64_t* h, const int32_t* nbins, int64_t* hist)
{
    GET_CTX(h);
    if (!nbins || !hist) return PERC_E_ARG;
    if (!c->labeled || c->nranks > 1) return PERC_E_STATE;      // slab handles: rank-local labels, clusters not stitched here
    return ccl_hist(c, *nbins, hist);
}

int32_t perc_hist_log2(const int64_t* h, const int32_t* nbins, int64_t* hist)
{
    GET_CTX(h);
    if (!c->labeled || !nbins || !hist) return PERC_E_STATE;
    if (c->nranks > 1) return PERC_E_STATE;
    return ccl_hist(c, *nbins, hist, 1);
}

static int finish_label(Ctx* c, int32_t* maxcs, int32_t* perccln, int32_t* perccls)
{
    const Summary& s = c->h_sum;
    int32_t ms = (int32_t)(s.maxpack >> 32);
    if (ms == 0 && s.nlone > 0) ms = 1;
    if (maxcs) *maxcs = ms;
    if (perccln) *perccln = c->h_span_ids.empty() ? 0 : c->h_span_ids[0];
    if (perccls) *perccls = c->h_span_ids.empty() ? 0 : c->h_span_sizes[0];
    return 0;
}

int32_t perc_site(const int64_t* h, const int32_t* order, const int32_t* k,
                  int32_t* s, int32_t* cs, int32_t* maxcs, int32_t* perccln, int32_t* perccls)
{
    GET_CTX(h);
    if (!order || !k) return PERC_E_ARG;
    int rc = occ_upload_site_order(c, order);
    if (rc) return rc;
    rc = set_fill(c, *k, -1);
    if (rc) return rc;
    rc = ccl_run(c, KIND_SITE);
    if (rc) return rc;
    if (s) { rc = download(c, s, c->label, sizeof(int32_t) * c->g.t); if (rc) return rc; }
    if (cs) { rc = ccl_export_sizes(c, cs); if (rc) return rc; }
    return finish_label(c, maxcs, perccln, perccls);
}

int32_t perc_bond(const int64_t* h, const int32_t* border, const int32_t* k,
                  int32_t* b3, int32_t* cs, int32_t* maxcs, int32_t* perccln, int32_t* perccls)
{
    GET_CTX(h);
    if (!border || !k) return PERC_E_ARG;
    int rc = occ_upload_bond_order(c, border);
    if (rc) return rc;
    rc = set_fill(c, -1, *k);
    if (rc) return rc;
    rc = ccl_run(c, KIND_BOND);
    if (rc) return rc;
    if (b3) { rc = ccl_export_bond_labels(c, b3); if (rc) return rc; }
    if (cs) { rc = ccl_export_sizes(c, cs); if (rc) return rc; }
    return finish_label(c, maxcs, perccln, perccls);
}

int32_t perc_sitebond(const int64_t* h, const int32_t* sorder, const int32_t* ks,
                      const int32_t* border, const int32_t* kb,
                      int32_t* s, int32_t* b3, int32_t* cs, int32_t* maxcs, int32_t* perccln, int32_t* perccls)
{
    GET_CTX(h);
    if (!sorder || !ks || !border || !kb) return PERC_E_ARG;
    int rc = occ_upload_site_order(c, sorder);
    if (rc) return rc;
    rc = occ_upload_bond_order(c, border);
    if (rc) return rc;
    rc = set_fill(c, *ks, *kb);
    if (rc) return rc;
    rc = ccl_run(c, KIND_MIXED);
    if (rc) return rc;
    if (s) { rc = download(c, s, c->label, sizeof(int32_t) * c->g.t); if (rc) return rc; }
    if (b3) { rc = ccl_export_bond_labels(c, b3); if (rc) return rc; }
    if (cs) { rc = ccl_export_sizes(c, cs); if (rc) return rc; }
    return finish_label(c, maxcs, perccln, perccls);
}

int32_t perc_first_span(const int64_t* h, const int32_t* kind, const int32_t* which,
                        int32_t* kstar, float* f, int32_t* maxcs, int32_t* perccls)
{
    GET_CTX(h);
    if (!kind || !which || !kstar) return PERC_E_ARG;
    if (*kind < KIND_SITE || *kind > KIND_MIXED) return PERC_E_ARG;
    if (*which != KIND_SITE && *which != KIND_BOND) return PERC_E_ARG;
    if (*kind == KIND_SITE && *which != KIND_SITE) return PERC_E_ARG;
    if (*kind == KIND_BOND && *which != KIND_BOND) return PERC_E_ARG;
    const bool sites = *which == KIND_SITE;
    int64_t N = sites ? c->g.t : c->g.nb;
    auto probe = [&](int k, int* spans) -> int {
        int rc = sites ? set_fill(c, k, -1) : set_fill(c, -1, k);
        if (rc) return rc;
        rc = ccl_run(c, *kind);
        if (rc) return rc;
        *spans = c->h_sum.nspan > 0;
        return 0;
    };
    int spans = 0;
    int rc = probe((int)N, &spans);
    if (rc) return rc;
    int lo = 0, hi = (int)N;            // invariant: lo does not span, hi spans
    if (!spans) hi = 0;
    else {
        while (hi - lo > 1) {
            int mid = lo + (hi - lo) / 2;
            rc = probe(mid, &spans);
            if (rc) return rc;
            if (spans) hi = mid; else lo = mid;
        }
        rc = probe(hi, &spans);
        if (rc) return rc;
    }
    *kstar = hi;
    if (f) *f = (float)hi / (float)N;     // f = real(sf)/real(t), REAL*4 (Sq/site_perc.f:221)
    return finish_label(c, maxcs, nullptr, perccls);
}

static int conduct_common(Ctx* c, const int32_t* cluster_id, const double* Va, const double* g0,
                          const double* gleak, const double* tol, const int32_t* itmax, const double* read_thresh,
                          int keep_x, double* Gtop, double* Gbot, int32_t* iter, double* err, int warm = 0)
{
    if (!cluster_id || !Va || !g0 || !gleak || !tol || !itmax || !read_thresh || !Gtop || !Gbot || !iter || !err)
        return PERC_E_ARG;
    if (!c->labeled) return PERC_E_STATE;
    if (c->g.n < 3 || *Va == 0.0 || *itmax < 0) return PERC_E_ARG;
    int cid = *cluster_id;
    if (c->nranks > 1) {
        // slab of a decomposed lattice: the default spanning cluster (smallest lattice-wide label); every
        // rank solves with the label that cluster carries locally (0: no part of it lives here)
        if (cid != 0) return PERC_E_ARG;
        if (c->h_span_gid.empty()) return PERC_E_NOSPAN;
        cid = slab_local_label_of(c, c->h_span_gid[0]);
    } else if (cid == 0) {
        if (c->h_span_ids.empty()) return PERC_E_NOSPAN;
        cid = c->h_span_ids[0];
    } else {
        bool ok = false;
        for (int v : c->h_span_ids) ok |= (v == cid);
        if (!ok) return PERC_E_NOSPAN;
    }
    int rc = ensure_pcg(c);
    if (rc) return rc;
    int it = 0;
    rc = pcg_solve(c, cid, *Va, *g0, *gleak, *tol, *itmax, *read_thresh, keep_x, Gtop, Gbot, &it, err, warm);
    *iter = it;
    return rc;
}

int32_t perc_conduct(const int64_t* h, const int32_t* cluster_id, const double* Va, const double* g0,
                     const double* gleak, const double* tol, const int32_t* itmax, const double* read_thresh,
                     double* Gtop, double* Gbot, int32_t* iter, double* err)
{
    GET_CTX(h);
    return conduct_common(c, cluster_id, Va, g0, gleak, tol, itmax, read_thresh, 1, Gtop, Gbot, iter, err);
}

int32_t perc_conduct_warm(const int64_t* h, const int32_t* cluster_id, const double* Va, const double* g0,
                          const double* gleak, const double* tol, const int32_t* itmax, const double* read_thresh,
                          double* Gtop, double* Gbot, int32_t* iter, double* err)
{
    GET_CTX(h);
    // the labeling between two sweep points invalidates `solved`; the voltages themselves are still there
    // (pcg_solve checks have_x itself; `solved` is only set by a solve that succeeded)
    const bool have = c->have_x && c->vx != nullptr;
    return conduct_common(c, cluster_id, Va, g0, gleak, tol, itmax, read_thresh, 1, Gtop, Gbot, iter, err, have ? 1 : 0);
}

int32_t perc_conduct_g(const int64_t* h, const int32_t* cluster_id, const double* Va, const double* g0,
                       const double* gleak, const double* tol, const int32_t* itmax, const double* read_thresh,
                       double* Gtop, double* Gbot, int32_t* iter, double* err)
{
    GET_CTX(h);
    return conduct_common(c, cluster_id, Va, g0, gleak, tol, itmax, read_thresh, 0, Gtop, Gbot, iter, err);
}

int32_t perc_get_voltage(const int64_t* h, double* Vint)
{
    GET_CTX(h);
    if (!c->solved || !c->have_x || !Vint) return PERC_E_STATE;
    return download(c, Vint, c->vx + c->g.m, sizeof(double) * (c->g.t - 2 * (int64_t)c->g.m));
}

// ---- one lattice decomposed into row slabs over several GPUs -------------------------------------
int32_t perc_create_slab(int64_t* h, const int32_t* lattice, const int32_t* m, const int32_t* n,
                         const int32_t* pbc, const int32_t* device, const int32_t* nranks, const int32_t* rank)
{
    if (!h || !lattice || !m || !n || !pbc || !device || !nranks || !rank) return PERC_E_ARG;
    if (*lattice != LAT_SQUARE && *lattice != LAT_TRIANGULAR) return PERC_E_ARG;
    if (*m < 2 || *n < 2 || (*pbc != 0 && *pbc != 1) || (*pbc && *m < 3)) return PERC_E_ARG;
    if (*lattice == LAT_TRIANGULAR && (*m & 1)) return PERC_E_ODD_M;
    if (*nranks < 1 || *rank < 0 || *rank >= *nranks) return PERC_E_ARG;
    if (*n / *nranks < 2) return PERC_E_ARG;                         // every slab needs two rows of its own
    Geom g = make_slab_geom(*lattice, *m, *n, *pbc, *nranks, *rank);
    if (g.t > 0x7ffffff0LL) return PERC_E_SIZE;                      // rank-local labels are 32-bit
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) return (int)e;
    if (*device < 0 || *device >= ndev) return PERC_E_ARG;
    Ctx* c = new Ctx();
    c->g = g;
    c->device = *device;
    c->nranks = *nranks; c->rank = *rank;
    int rc = ctx_alloc(c);
    if (rc) { ctx_free(c); delete c; return rc; }
    std::lock_guard<std::mutex> lk(g_mu);
    *h = g_next++;
    g_ctx[*h] = c;
    return 0;
}

int32_t perc_comm_unique_id(uint8_t* id128)
{
    if (!id128) return PERC_E_ARG;
    return slab_unique_id(id128);
}

int32_t perc_comm_init(const int64_t* h, const uint8_t* id128)
{
    GET_CTX(h);
    if (!id128) return PERC_E_ARG;
    if (c->nranks == 1) return 0;
    return slab_comm_init(c, id128);
}

int32_t perc_slab_rows(const int64_t* h, int32_t* ya, int32_t* yb)
{
    GET_CTX(h);
    if (ya) *ya = c->g.y0 + c->g.own_lo;
    if (yb) *yb = c->g.y0 + c->g.own_hi;
    return 0;
}

int32_t perc_generate_i8(const int64_t* h, const int64_t* seed, const int64_t* stream, const int64_t* ks, const int64_t* kb)
{
    GET_CTX(h);
    if (!seed || !stream || !ks || !kb) return PERC_E_ARG;
    return occ_generate(c, (unsigned long long)*seed, (unsigned long long)*stream, *ks, *kb);
}

int32_t perc_summary_i8(const int64_t* h, int64_t* ncl, int64_t* maxcs, int64_t* maxcn, int64_t* nspan)
{
    GET_CTX(h);
    if (!c->labeled) return PERC_E_STATE;
    if (c->nranks > 1) {
        const StitchResult& R = c->stitch;
        if (ncl) *ncl = R.ncl + R.nlone;
        if (maxcs) *maxcs = R.maxcs ? R.maxcs : (R.nlone > 0 ? 1 : 0);
        if (maxcn) *maxcn = R.maxgid;
        if (nspan) *nspan = (int64_t)R.span_gid.size();
        return 0;
    }
    int32_t ms = 0, mn = 0, ns = 0;
    int rc = perc_summary(h, ncl, &ms, &mn, &ns);
    if (maxcs) *maxcs = ms;
    if (maxcn) *maxcn = mn;
    if (nspan) *nspan = ns;
    return rc;
}

int32_t perc_span_i8(const int64_t* h, const int32_t* max_ids, int32_t* nspan, int64_t* ids, int64_t* sizes)
{
    GET_CTX(h);
    if (!c->labeled || !max_ids || !nspan) return PERC_E_STATE;
    *nspan = (int32_t)c->h_span_gid.size();
    int k = (int)c->h_span_gid.size();
    if (k > *max_ids) k = *max_ids;
    for (int j = 0; j < k; ++j) { if (ids) ids[j] = c->h_span_gid[j]; if (sizes) sizes[j] = c->h_span_total[j]; }
    return 0;
}

int32_t perc_get_site_labels_i8(const int64_t* h, int64_t* s)
{
    GET_CTX(h);
    if (!s) return PERC_E_ARG;
    if (!c->labeled) return PERC_E_STATE;
    return slab_export_labels(c, s);
}

// the redundant host union-find of the stitch, on its own (no device): gathered = nranks blocks of
// 5*m + 8 int64 words (layout in csrc/slab.h).  out_summary[5] = ncl, nlone, maxcs, maxcn, nspan;
// pairs (root id, representative id, class label, class size) of the calling rank go to out_pairs[4*k..].
int32_t perc_stitch_host(const int32_t* nranks, const int32_t* rank, const int32_t* m, const int64_t* gathered,
                         int64_t* out_summary, const int32_t* max_span, int64_t* span_ids, int64_t* span_sizes,
                         const int32_t* max_pairs, int32_t* npairs, int64_t* out_pairs)
{
    if (!nranks || !rank || !m || !gathered || !out_summary || !max_span || !max_pairs || !npairs) return PERC_E_ARG;
    StitchResult R;
    stitch_host(*nranks, *rank, *m, gathered, &R);
    if (R.error) return PERC_E_ARG;
    out_summary[0] = R.ncl; out_summary[1] = R.nlone; out_summary[2] = R.maxcs; out_summary[3] = R.maxgid;
    out_summary[4] = (int64_t)R.span_gid.size();
    for (int k = 0; k < (int)R.span_gid.size() && k < *max_span; ++k) {
        if (span_ids) span_ids[k] = R.span_gid[k];
        if (span_sizes) span_sizes[k] = R.span_size[k];
    }
    *npairs = (int32_t)R.root_gid.size();
    for (int k = 0; k < *npairs && k < *max_pairs && out_pairs; ++k) {
        out_pairs[4 * k + 0] = R.root_gid[k]; out_pairs[4 * k + 1] = R.rep_gid[k];
        out_pairs[4 * k + 2] = R.class_gid[k]; out_pairs[4 * k + 3] = R.class_total[k];
    }
    return 0;
}

// ---- batches of independent realizations (trial loops of the *_perc / bond_cond drivers) ----------
int32_t perc_batch(const int64_t* h, const int32_t* kind, const int32_t* nreal, const int64_t* seed, const int64_t* stream0,
                   const int32_t* ks, const int32_t* kb, const int32_t* nbins, int64_t* hist, int64_t* stats)
{
    GET_CTX(h);
    if (!kind || !nreal || !seed || !stream0 || !ks || !kb || !nbins || !stats) return PERC_E_ARG;
    if (*kind < KIND_SITE || *kind > KIND_MIXED || *nreal < 0 || *nbins < 0) return PERC_E_ARG;
    if ((*kind != KIND_BOND && (*ks < 0 || *ks > c->g.t)) || (*kind != KIND_SITE && (*kb < 0 || *kb > c->g.nb))) return PERC_E_ARG;
    int rc = batch_run(c, *kind, *nreal, (unsigned long long)*seed, (unsigned long long)*stream0, *ks, *kb, *nbins, hist, stats);
    return rc == 0 && stats[6] != 0 ? PERC_E_SELECT : rc;
}

int32_t perc_batch_conduct(const int64_t* h, const int32_t* kind, const int32_t* nreal, const int64_t* seed, const int64_t* stream0,
                           const int32_t* ks, const int32_t* kb, const double* Va, const double* g0, const double* gleak,
                           const double* tol, const int32_t* itmax, const double* read_thresh,
                           double* G, int32_t* iters, int64_t* stats)
{
    GET_CTX(h);
    if (!kind || !nreal || !seed || !stream0 || !ks || !kb || !Va || !g0 || !gleak || !tol || !itmax || !read_thresh || !G || !iters || !stats)
        return PERC_E_ARG;
    if (*kind < KIND_SITE || *kind > KIND_MIXED || *nreal < 0 || c->g.n < 3 || *Va == 0.0 || *itmax < 0) return PERC_E_ARG;
    if ((*kind != KIND_BOND && (*ks < 0 || *ks > c->g.t)) || (*kind != KIND_SITE && (*kb < 0 || *kb > c->g.nb))) return PERC_E_ARG;
    if (!pcg_small_fits(c->g)) { int rc = ensure_pcg(c); if (rc) return rc; }
    int rc = batch_conduct_run(c, *kind, *nreal, (unsigned long long)*seed, (unsigned long long)*stream0, *ks, *kb,
                               *Va, *g0, *gleak, *tol, *itmax, *read_thresh, G, iters, stats);
    return rc == 0 && stats[6] != 0 ? PERC_E_SELECT : rc;
}

// communicator for handles that shard independent realizations (mode 1): same bootstrap as the slab mode
int32_t perc_comm_init_rank(const int64_t* h, const int32_t* nranks, const int32_t* rank, const uint8_t* id128)
{
    GET_CTX(h);
    if (!nranks || !rank || !id128 || *nranks < 1 || *rank < 0 || *rank >= *nranks) return PERC_E_ARG;
    if (c->nranks > 1) return PERC_E_STATE;                 // a slab handle already owns its communicator
    if (*nranks == 1) return 0;
    const int keep_n = c->nranks, keep_r = c->rank;
    c->nranks = *nranks; c->rank = *rank;
    int rc = slab_comm_init(c, id128);
    c->stat_nranks = *nranks;
    c->nranks = keep_n; c->rank = keep_r;
    return rc;
}

// the one NCCL reduction of the statistics at the end of a sharded run (sum of int64 / float64 arrays,
// host buffers in and out); a handle without a communicator leaves the arrays as they are
int32_t perc_allreduce_stats(const int64_t* h, const int32_t* ni, int64_t* ivals, const int32_t* nd, double* dvals)
{
    GET_CTX(h);
    if (!ni || !nd || *ni < 0 || *nd < 0) return PERC_E_ARG;
    if (!c->comm || c->stat_nranks <= 1) return 0;
    const size_t bytes = sizeof(int64_t) * (size_t)*ni + sizeof(double) * (size_t)*nd;
    if (!bytes) return 0;
    char* d = (char*)ctx_dev_stage(c, bytes + 64);
    if (!d) return (int)cudaErrorMemoryAllocation;
    int64_t* di = (int64_t*)d;
    double* dd = (double*)(d + sizeof(int64_t) * (size_t)*ni);
    if (*ni) PERC_CUDA(cudaMemcpyAsync(di, ivals, sizeof(int64_t) * *ni, cudaMemcpyHostToDevice, c->stream));
    if (*nd) PERC_CUDA(cudaMemcpyAsync(dd, dvals, sizeof(double) * *nd, cudaMemcpyHostToDevice, c->stream));
    const int keep = c->nranks;
    c->nranks = c->stat_nranks;                              // the reduction helpers skip single-rank handles
    int rc = 0;
    if (*ni) rc = slab_allreduce_i64(c, di, *ni);
    if (!rc && *nd) rc = slab_allreduce_f64(c, dd, *nd);
    c->nranks = keep;
    if (rc) return rc;
    if (*ni) PERC_CUDA(cudaMemcpyAsync(ivals, di, sizeof(int64_t) * *ni, cudaMemcpyDeviceToHost, c->stream));
    if (*nd) PERC_CUDA(cudaMemcpyAsync(dvals, dd, sizeof(double) * *nd, cudaMemcpyDeviceToHost, c->stream));
    PERC_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int32_t perc_launch_count(const int64_t* h, int64_t* count)
{
    GET_CTX(h);
    if (!count) return PERC_E_ARG;
    *count = c->launches;
    return 0;
}

int32_t perc_phase_ms(const int64_t* h, const int32_t* nphase, float* ms)
{
    GET_CTX(h);
    if (!nphase || !ms) return PERC_E_ARG;
    for (int k = 0; k < *nphase && k < 8; ++k) ms[k] = c->phase_ms[k];
    return 0;
}

int32_t perc_set_solver(const int64_t* h, const int32_t* mode)
{
    GET_CTX(h);
    if (!mode || *mode < 0 || (*mode > 2 && (*mode != 10 && *mode != 12 && *mode != 14))) return PERC_E_ARG;
    // 10 / 12 / 14 (diagnostic): the one-pass kernel, variant FtCfgA / FtCfgA3 / FtCfgD (pcg_fused_tile.cuh)
    const int pm = *mode >= 10 ? 0 : *mode, fc = *mode >= 10 ? *mode - 10 : -1;
    c->pcg_mode = pm; c->fused_cfg = fc;
    for (Ctx* k : c->batch_kids) { k->pcg_mode = pm; k->fused_cfg = fc; }
    return 0;
}

// per-bond conductances: w(nb), reference bond-row order; NULL = back to the uniform g0 (MATLAB/ConductCalc.m:38-47,94-96)
int32_t perc_set_bond_conductance(const int64_t* h, const double* w)
{
    GET_CTX(h);
    if (c->nranks > 1) return PERC_E_STATE;
    const int rc = pcg_set_bond_weights(c, w);
    c->solved = false;
    return rc < 0 ? PERC_E_ARG : rc;
}

// back to the uniform g0 on every conducting bond (a Fortran caller cannot pass a null array)
int32_t perc_clear_bond_conductance(const int64_t* h) { return perc_set_bond_conductance(h, nullptr); }

int32_t perc_solver_used(const int64_t* h, int32_t* fused)
{
    GET_CTX(h);
    if (!fused) return PERC_E_ARG;
    *fused = c->last_fused ? (c->last_fused_cfg == 4 ? 2 : 1) : 0;
    return 0;
}

int32_t perc_stream(const int64_t* h, uint64_t* stream)
{
    GET_CTX(h);
    if (!stream) return PERC_E_ARG;
    *stream = (uint64_t)(uintptr_t)c->stream;
    return 0;
}

}  // extern "C"
