// slab.cu -- slab decomposition of one lattice over several GPUs (see slab.h).
#include <dlfcn.h>
#include <algorithm>
#include <cstring>
#include "context.h"
#include "slab.h"
#include "slab_stitch.cuh"

namespace perc {

// ------------------------------------------------------------------------------------------
// the stitch (slab_stitch.cuh) run item by item on the host: CPU tests of the logic the kernels below execute
// ------------------------------------------------------------------------------------------
static int stitch_table_size(int nranks, int m, int* logH)
{
    int64_t entries = 2 * (int64_t)(nranks > 1 ? nranks - 1 : 1) * m;
    int lg = 10;
    while (((int64_t)1 << lg) < 2 * entries) ++lg;
    *logH = lg;
    return 1 << lg;
}

static void stitch_collect(const StitchTab& T, const long long* out, unsigned long long maxgid, const int64_t* span,
                           const int64_t* pairs, int m, int nranks, const int64_t* gathered, StitchResult* R)
{
    IfaceLayout L{m};
    *R = StitchResult();
    R->error = (int)out[6];
    if (R->error) return;
    for (int r = 0; r < nranks; ++r) {
        const int64_t* sc = gathered + r * L.words() + L.scalars();
        R->ncl += sc[0];
        R->nlone += sc[1];
    }
    R->ncl += out[0];
    R->maxcs = out[1];
    R->maxgid = out[1] > 0 && maxgid != STITCH_NONE ? (int64_t)maxgid : 0;
    int ns = (int)(out[3] < MAX_SPAN_CLASSES ? out[3] : MAX_SPAN_CLASSES);
    std::vector<std::pair<int64_t, int64_t>> sp;
    for (int k = 0; k < ns; ++k) sp.push_back({span[2 * k], span[2 * k + 1]});
    std::sort(sp.begin(), sp.end());
    for (auto& e : sp) { R->span_gid.push_back(e.first); R->span_size.push_back(e.second); }
    int np = (int)(out[4] < T.cap ? out[4] : T.cap);
    std::vector<int> order(np);
    for (int k = 0; k < np; ++k) order[k] = k;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return pairs[4 * a] < pairs[4 * b]; });
    for (int k : order) {
        R->root_gid.push_back(pairs[4 * k]); R->rep_gid.push_back(pairs[4 * k + 1]);
        R->class_gid.push_back(pairs[4 * k + 2]); R->class_total.push_back(pairs[4 * k + 3]);
    }
}

void stitch_host(int nranks, int rank, int m, const int64_t* gathered, StitchResult* out)
{
    StitchTab T{};
    T.H = stitch_table_size(nranks, m, &T.logH);
    T.cap = 2 * m;
    std::vector<int64_t> keys(T.H), hsize(T.H), span(2 * MAX_SPAN_CLASSES), pairs(4 * (size_t)T.cap);
    std::vector<int32_t> parent(T.H), cpos(T.H), flag(T.H);
    std::vector<unsigned long long> cgid(T.H), rep(T.H);
    std::vector<long long> ctot(T.H), scal(8);
    T.keys = keys.data(); T.hsize = hsize.data(); T.parent = parent.data(); T.cgid = cgid.data(); T.ctot = ctot.data();
    T.cpos = cpos.data(); T.rep = rep.data(); T.flag = flag.data(); T.out = scal.data(); T.span = span.data(); T.pairs = pairs.data();
    unsigned long long maxgid = STITCH_NONE;
    for (int s = 0; s < T.H; ++s) stitch_clear(T, s);
    for (int64_t e = 0; e < 2 * (int64_t)nranks * m; ++e) stitch_nodes(T, nranks, m, gathered, e);
    for (int64_t e = (int64_t)nranks * m - 1; e >= 0; --e) stitch_unions(T, nranks, m, gathered, e);       // any order must work
    for (int s = 0; s < T.H; ++s) stitch_classes(T, rank, s);
    for (int s = 0; s < T.H; ++s) stitch_summary_a(T, s);
    for (int r = 0; r < nranks; ++r) stitch_summary_rank(T, m, gathered, r);
    for (int s = 0; s < T.H; ++s) stitch_summary_b(T, s, &maxgid);
    for (int r = 0; r < nranks; ++r) stitch_summary_rank_b(T, m, gathered, r, &maxgid);
    for (int x = 0; x < m; ++x) stitch_span(T, nranks, m, gathered, x);
    for (int s = 0; s < T.H; ++s) stitch_mine(T, rank, s, nullptr, 0);
    stitch_collect(T, scal.data(), maxgid, span.data(), pairs.data(), m, nranks, gathered, out);
}

// ------------------------------------------------------------------------------------------
// NCCL, loaded at run time (single-GPU use of the library does not need it)
// ------------------------------------------------------------------------------------------
namespace {
typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
enum { NCCL_UINT8 = 1, NCCL_INT64 = 4, NCCL_FLOAT64 = 8, NCCL_SUM = 0 };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
} g_nccl;

int nccl_load()
{
    if (g_nccl.lib) return 0;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return -7;
#define SYM(field, name) *(void**)(&g_nccl.field) = dlsym(lib, name); if (!g_nccl.field) return -7;
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce") SYM(AllGather, "ncclAllGather") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
#undef SYM
    g_nccl.lib = lib;
    return 0;
}
#define PERC_NCCL(call) do { int e__ = (call); if (e__ != 0) return 1000 + e__; } while (0)
}  // namespace

int slab_unique_id(void* id128)
{
    int rc = nccl_load();
    if (rc) return rc;
    NcclUniqueId id;
    PERC_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return 0;
}

int slab_comm_init(Ctx* c, const void* id128)
{
    int rc = nccl_load();
    if (rc) return rc;
    NcclUniqueId id;
    memcpy(&id, id128, 128);
    NcclComm comm = nullptr;
    PERC_NCCL(g_nccl.CommInitRank(&comm, c->nranks, id, c->rank));
    c->comm = comm;
    return 0;
}

void slab_comm_destroy(Ctx* c)
{
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy((NcclComm)c->comm);
    c->comm = nullptr;
}

int slab_allreduce_f64(Ctx* c, double* buf, int count)
{
    if (c->nranks == 1) return 0;
    PERC_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, NCCL_FLOAT64, NCCL_SUM, (NcclComm)c->comm, c->stream));
    return 0;
}

int slab_allreduce_i64(Ctx* c, int64_t* buf, int count)
{
    if (c->nranks == 1) return 0;
    PERC_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, NCCL_INT64, NCCL_SUM, (NcclComm)c->comm, c->stream));
    return 0;
}

// halo exchange of one array (`bytes_per_site` bytes per site): my first owned row -> top halo of the
// rank below, my last owned row -> bottom halo of the rank above; both directions in one NCCL group
int slab_halo_exchange(Ctx* c, void* array, int bytes_per_site)
{
    if (c->nranks == 1) return 0;
    const Geom& g = c->g;
    const size_t rowb = (size_t)g.m * bytes_per_site;
    char* a = (char*)array;
    PERC_NCCL(g_nccl.GroupStart());
    if (c->rank > 0) {
        PERC_NCCL(g_nccl.Send(a + rowb * g.own_lo, rowb, NCCL_UINT8, c->rank - 1, (NcclComm)c->comm, c->stream));
        PERC_NCCL(g_nccl.Recv(a + rowb * (g.own_lo - 1), rowb, NCCL_UINT8, c->rank - 1, (NcclComm)c->comm, c->stream));
    }
    if (c->rank + 1 < c->nranks) {
        PERC_NCCL(g_nccl.Send(a + rowb * (g.own_hi - 1), rowb, NCCL_UINT8, c->rank + 1, (NcclComm)c->comm, c->stream));
        PERC_NCCL(g_nccl.Recv(a + rowb * g.own_hi, rowb, NCCL_UINT8, c->rank + 1, (NcclComm)c->comm, c->stream));
    }
    PERC_NCCL(g_nccl.GroupEnd());
    return 0;
}

// ------------------------------------------------------------------------------------------
// device side of the stitch
// ------------------------------------------------------------------------------------------
// interface rows -> lattice-wide ids (rank-local label + y0 * m) and rank-local sizes
__global__ void __launch_bounds__(256)
slab_iface_kernel(Geom g, int rank, int nranks, const int32_t* __restrict__ label, const int32_t* __restrict__ size,
                  int64_t* __restrict__ blk)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= g.m) return;
    const int64_t off = (int64_t)g.y0 * g.m;
    IfaceLayout L{g.m};
    int64_t a = 0, b = 0, sa = 0, sb = 0, top = 0;
    if (rank > 0) { int32_t l = label[(int64_t)g.own_lo * g.m + x]; if (l) { a = l + off; sa = size[l - 1]; } }
    if (rank + 1 < nranks) { int32_t l = label[(int64_t)g.own_hi * g.m + x]; if (l) { b = l + off; sb = size[l - 1]; } }
    if (rank + 1 == nranks) { int32_t l = label[(int64_t)(g.n - 1) * g.m + x]; if (l) top = l + off; }
    blk[L.rowA() + x] = a; blk[L.rowB() + x] = b; blk[L.sizeA() + x] = sa; blk[L.sizeB() + x] = sb; blk[L.rowTop() + x] = top;
}

__global__ void slab_scalars_kernel(Geom g, const Summary* __restrict__ sum, int64_t* __restrict__ blk)
{
    IfaceLayout L{g.m};
    int64_t* sc = blk + L.scalars();
    const unsigned long long mp = sum->maxpack;
    sc[0] = (int64_t)sum->ncl; sc[1] = (int64_t)sum->nlone;
    sc[2] = (int64_t)(mp >> 32);
    sc[3] = mp ? (int64_t)(0xffffffffu - (unsigned)(mp & 0xffffffffu)) + (int64_t)g.y0 * g.m : 0;
    for (int k = 4; k < 8; ++k) sc[k] = 0;
}

// exact count / largest cluster of the border roots (slab handles: a root may have no owned site at all)
__global__ void __launch_bounds__(256)
slab_rootcount_kernel(const int32_t* __restrict__ label, const int32_t* __restrict__ size, const int32_t* __restrict__ rootlist,
                      Summary* __restrict__ sum)
{
    const unsigned nroots = sum->nroots;
    unsigned cnt = 0;
    unsigned long long best = 0;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nroots; k += (int64_t)gridDim.x * blockDim.x) {
        const int j = rootlist[k];
        if (label[j] != j + 1) continue;
        const int s = size[j];
        if (s <= 0) continue;
        cnt++;
        unsigned long long pk = ((unsigned long long)(unsigned)s << 32) | (unsigned long long)(0xffffffffu - (unsigned)(j + 1));
        if (pk > best) best = pk;
    }
    if (cnt) atomicAdd(&sum->ncl, (unsigned long long)cnt);
    if (best) atomicMax(&sum->maxpack, best);
}

__global__ void __launch_bounds__(256)
slab_flatten_kernel(int64_t t, int32_t* label)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t) return;
    int32_t l = label[i];
    if (!l) return;
    int32_t o = __ldcg(&label[l - 1]);
    if (o != l) label[i] = o;
}

// owned rows -> lattice-wide canonical labels (int64): interface classes through the sorted table
__global__ void __launch_bounds__(256)
slab_export_labels_kernel(Geom g, const int32_t* __restrict__ label, int ntab, const int32_t* __restrict__ tab_rep,
                          const int64_t* __restrict__ tab_gid, int64_t* __restrict__ out)
{
    const int64_t nown = (int64_t)(g.own_hi - g.own_lo) * g.m;
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nown) return;
    const int32_t l = label[k + (int64_t)g.own_lo * g.m];
    int64_t o = 0;
    if (l) {
        o = (int64_t)l + (int64_t)g.y0 * g.m;
        int lo = 0, hi = ntab;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (tab_rep[mid] < l - 1) lo = mid + 1; else hi = mid; }
        if (lo < ntab && tab_rep[lo] == l - 1) o = tab_gid[lo];
    }
    out[k] = o;
}

static unsigned nblk(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

// after the rank-local labeling: exact counts of the border roots (no rootfix counting on slab handles)
int slab_count_roots(Ctx* c)
{
    slab_rootcount_kernel<<<148 * 4, 256, 0, c->stream>>>(c->label, c->size, c->rootlist, c->d_sum);
    c->launches++;
    return (int)cudaGetLastError();
}

// the stitch phases as kernels: one thread per item
__global__ void k_stitch_clear(StitchTab T) { int s = blockIdx.x * blockDim.x + threadIdx.x; if (s < T.H) stitch_clear(T, s); }
__global__ void k_stitch_nodes(StitchTab T, int nranks, int m, const int64_t* __restrict__ gathered)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < 2 * (int64_t)nranks * m) stitch_nodes(T, nranks, m, gathered, e);
}
__global__ void k_stitch_unions(StitchTab T, int nranks, int m, const int64_t* __restrict__ gathered)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < (int64_t)nranks * m) stitch_unions(T, nranks, m, gathered, e);
}
__global__ void k_stitch_classes(StitchTab T, int rank) { int s = blockIdx.x * blockDim.x + threadIdx.x; if (s < T.H) stitch_classes(T, rank, s); }
__global__ void k_stitch_summary_a(StitchTab T, int nranks, int m, const int64_t* __restrict__ gathered)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < T.H) stitch_summary_a(T, s);
    if (s < nranks) stitch_summary_rank(T, m, gathered, s);
}
__global__ void k_stitch_summary_b(StitchTab T, int nranks, int m, const int64_t* __restrict__ gathered, unsigned long long* maxgid)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < T.H) stitch_summary_b(T, s, maxgid);
    if (s < nranks) stitch_summary_rank_b(T, m, gathered, s, maxgid);
}
__global__ void k_stitch_span(StitchTab T, int nranks, int m, const int64_t* __restrict__ gathered)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x < m) stitch_span(T, nranks, m, gathered, x);
}
__global__ void k_stitch_mine(StitchTab T, int rank, int32_t* label, int64_t off)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < T.H) stitch_mine(T, rank, s, label, off);
}

// device scratch of the stitch: one allocation, carved into the table's arrays
static int stitch_device_table(Ctx* c, StitchTab* T, unsigned long long** maxgid)
{
    const Geom& g = c->g;
    T->H = stitch_table_size(c->nranks, g.m, &T->logH);
    T->cap = 2 * g.m;
    const size_t H = (size_t)T->H;
    const size_t bytes = H * (8 + 8 + 4 + 8 + 8 + 4 + 8 + 4) + 8 * 8 + 8 + sizeof(int64_t) * 2 * MAX_SPAN_CLASSES + sizeof(int64_t) * 4 * (size_t)T->cap + 256;
    if (bytes > c->d_stitch_bytes) {
        if (c->d_stitch) cudaFree(c->d_stitch);
        c->d_stitch = nullptr; c->d_stitch_bytes = 0;
        PERC_CUDA(cudaMalloc(&c->d_stitch, bytes));
        c->d_stitch_bytes = bytes;
    }
    char* p = (char*)c->d_stitch;
    T->keys = (int64_t*)p; p += 8 * H;
    T->hsize = (int64_t*)p; p += 8 * H;
    T->cgid = (unsigned long long*)p; p += 8 * H;
    T->ctot = (long long*)p; p += 8 * H;
    T->rep = (unsigned long long*)p; p += 8 * H;
    T->out = (long long*)p; p += 8 * 8;
    *maxgid = (unsigned long long*)p; p += 8;
    T->span = (int64_t*)p; p += sizeof(int64_t) * 2 * MAX_SPAN_CLASSES;
    T->pairs = (int64_t*)p; p += sizeof(int64_t) * 4 * (size_t)T->cap;
    T->parent = (int32_t*)p; p += 4 * H;
    T->cpos = (int32_t*)p; p += 4 * H;
    T->flag = (int32_t*)p; p += 4 * H;
    return 0;
}

// all-gather of the interface rows, then the union-find of slab_stitch.cuh on the device (every rank runs the
// same stitch redundantly: no second exchange) and the relabel of this rank's interface roots
int slab_stitch(Ctx* c)
{
    const Geom& g = c->g;
    IfaceLayout L{g.m};
    const int64_t W = L.words();
    cudaStream_t st = c->stream;
    if (!c->d_iface) {
        PERC_CUDA(cudaMalloc(&c->d_iface, sizeof(int64_t) * W * (c->nranks + 1)));
        PERC_CUDA(cudaMallocHost(&c->h_iface, sizeof(int64_t) * (W * c->nranks + 4 * (size_t)(2 * g.m) + 2 * MAX_SPAN_CLASSES + 16)));
    }
    int64_t* mine = c->d_iface + W * c->nranks;
    slab_iface_kernel<<<nblk(g.m), 256, 0, st>>>(g, c->rank, c->nranks, c->label, c->size, mine);
    slab_scalars_kernel<<<1, 1, 0, st>>>(g, c->d_sum, mine);
    c->launches += 2;
    if (c->nranks > 1) {
        if (!c->comm) return -4;
        PERC_NCCL(g_nccl.AllGather(mine, c->d_iface, (size_t)W, NCCL_INT64, (NcclComm)c->comm, st));
    } else PERC_CUDA(cudaMemcpyAsync(c->d_iface, mine, sizeof(int64_t) * W, cudaMemcpyDeviceToDevice, st));

    StitchTab T{};
    unsigned long long* d_maxgid = nullptr;
    int rc = stitch_device_table(c, &T, &d_maxgid);
    if (rc) return rc;
    const int64_t off = (int64_t)g.y0 * g.m;
    k_stitch_clear<<<nblk(T.H), 256, 0, st>>>(T);
    PERC_CUDA(cudaMemsetAsync(d_maxgid, 0xff, sizeof(unsigned long long), st));
    k_stitch_nodes<<<nblk(2 * (int64_t)c->nranks * g.m), 256, 0, st>>>(T, c->nranks, g.m, c->d_iface);
    k_stitch_unions<<<nblk((int64_t)c->nranks * g.m), 256, 0, st>>>(T, c->nranks, g.m, c->d_iface);
    k_stitch_classes<<<nblk(T.H), 256, 0, st>>>(T, c->rank);
    k_stitch_summary_a<<<nblk(T.H), 256, 0, st>>>(T, c->nranks, g.m, c->d_iface);
    k_stitch_summary_b<<<nblk(T.H), 256, 0, st>>>(T, c->nranks, g.m, c->d_iface, d_maxgid);
    k_stitch_span<<<nblk(g.m), 256, 0, st>>>(T, c->nranks, g.m, c->d_iface);
    k_stitch_mine<<<nblk(T.H), 256, 0, st>>>(T, c->rank, c->label, off);
    slab_flatten_kernel<<<nblk(g.t), 256, 0, st>>>(g.t, c->label);
    c->launches += 9;
    // results: scalars + largest label + spanning classes + this rank's (root, representative, label, size) list
    int64_t* h_sc = c->h_iface + W * c->nranks;            // [8] out, [1] maxgid, spans, pairs
    PERC_CUDA(cudaMemcpyAsync(h_sc, T.out, sizeof(long long) * 8 + 8 + sizeof(int64_t) * 2 * MAX_SPAN_CLASSES, cudaMemcpyDeviceToHost, st));
    PERC_CUDA(cudaMemcpyAsync(c->h_iface, c->d_iface, sizeof(int64_t) * W * c->nranks, cudaMemcpyDeviceToHost, st));
    PERC_CUDA(cudaStreamSynchronize(st));
    const long long* out = (const long long*)h_sc;
    const int np = (int)(out[4] < T.cap ? out[4] : T.cap);
    int64_t* h_pairs = h_sc + 9 + 2 * MAX_SPAN_CLASSES;
    if (np) {
        PERC_CUDA(cudaMemcpyAsync(h_pairs, T.pairs, sizeof(int64_t) * 4 * (size_t)np, cudaMemcpyDeviceToHost, st));
        PERC_CUDA(cudaStreamSynchronize(st));
    }
    StitchResult& R = c->stitch;
    stitch_collect(T, out, (unsigned long long)h_sc[8], h_sc + 9, h_pairs, g.m, c->nranks, c->h_iface, &R);
    if (R.error) return -8;
    // table representative -> (class label, class size), ascending representative
    c->tab_rep.clear(); c->tab_gid.clear(); c->tab_total.clear();
    std::vector<std::pair<int32_t, int>> reps;
    for (int k = 0; k < (int)R.root_gid.size(); ++k)
        if (R.root_gid[k] == R.rep_gid[k]) reps.push_back({(int32_t)(R.rep_gid[k] - off - 1), k});
    std::sort(reps.begin(), reps.end());
    for (auto& e : reps) { c->tab_rep.push_back(e.first); c->tab_gid.push_back(R.class_gid[e.second]); c->tab_total.push_back(R.class_total[e.second]); }
    return 0;
}

// rank-local representative (label value) of the lattice-wide cluster `gid`; 0 if this rank holds none of it
int32_t slab_local_label_of(const Ctx* c, int64_t gid)
{
    for (size_t k = 0; k < c->tab_gid.size(); ++k) if (c->tab_gid[k] == gid) return c->tab_rep[k] + 1;
    const int64_t off = (int64_t)c->g.y0 * c->g.m;
    if (c->nranks == 1 && gid - off >= 1 && gid - off <= c->g.t) return (int32_t)(gid - off);
    return 0;
}

int slab_export_labels(Ctx* c, int64_t* out)
{
    const Geom& g = c->g;
    const int64_t nown = (int64_t)(g.own_hi - g.own_lo) * g.m;
    const int ntab = (int)c->tab_rep.size();
    char* d = (char*)ctx_dev_stage(c, sizeof(int64_t) * nown + (sizeof(int32_t) + sizeof(int64_t)) * (size_t)(ntab + 2) + 64);
    if (!d) return (int)cudaErrorMemoryAllocation;
    int64_t* d_out = (int64_t*)d;
    int64_t* d_gid = d_out + nown;
    int32_t* d_rep = (int32_t*)(d_gid + ntab + 1);
    if (ntab) {
        PERC_CUDA(cudaMemcpyAsync(d_gid, c->tab_gid.data(), sizeof(int64_t) * ntab, cudaMemcpyHostToDevice, c->stream));
        PERC_CUDA(cudaMemcpyAsync(d_rep, c->tab_rep.data(), sizeof(int32_t) * ntab, cudaMemcpyHostToDevice, c->stream));
    }
    slab_export_labels_kernel<<<nblk(nown), 256, 0, c->stream>>>(g, c->label, ntab, d_rep, d_gid, d_out);
    c->launches++;
    PERC_CUDA(cudaMemcpyAsync(out, d_out, sizeof(int64_t) * nown, cudaMemcpyDeviceToHost, c->stream));
    PERC_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

}  // namespace perc
