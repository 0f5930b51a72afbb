// slab.cu -- slab decomposition of one lattice over several GPUs (see slab.h).
#include <dlfcn.h>
#include <algorithm>
#include <cstring>
#include <unordered_map>
#include "context.h"
#include "slab.h"

namespace perc {

// ------------------------------------------------------------------------------------------
// host: union-find over the interface labels (run redundantly by every rank)
// ------------------------------------------------------------------------------------------
namespace {
struct Node { int64_t key; int64_t size; };       // key = (rank << 40) | lattice-wide id of the rank-local root

inline int64_t node_key(int rank, int64_t gid) { return ((int64_t)rank << 40) | gid; }

int uf_find(std::vector<int>& p, int a)
{
    while (p[a] != a) { p[a] = p[p[a]]; a = p[a]; }
    return a;
}
}  // namespace

void stitch_host(int nranks, int rank, int m, const int64_t* gathered, StitchResult* out)
{
    IfaceLayout L{m};
    const int64_t W = L.words();
    *out = StitchResult();
    // 1. nodes: the distinct (rank, cluster) pairs seen on interface rows, with their rank-local sizes
    std::vector<Node> nodes;
    for (int r = 0; r < nranks; ++r) {
        const int64_t* blk = gathered + r * W;
        for (int side = 0; side < 2; ++side) {
            if ((side == 0 && r == 0) || (side == 1 && r == nranks - 1)) continue;
            const int64_t* ids = blk + (side ? L.rowB() : L.rowA());
            const int64_t* sz = blk + (side ? L.sizeB() : L.sizeA());
            for (int x = 0; x < m; ++x)                      // neighbouring sites mostly repeat the label: keep the changes
                if (ids[x] && (x == 0 || ids[x] != ids[x - 1])) nodes.push_back({node_key(r, ids[x]), sz[x]});
        }
    }
    std::sort(nodes.begin(), nodes.end(), [](const Node& a, const Node& b) { return a.key < b.key; });
    nodes.erase(std::unique(nodes.begin(), nodes.end(), [](const Node& a, const Node& b) { return a.key == b.key; }), nodes.end());
    auto index_of = [&](int64_t key) -> int {
        auto it = std::lower_bound(nodes.begin(), nodes.end(), key, [](const Node& a, int64_t k) { return a.key < k; });
        return (it != nodes.end() && it->key == key) ? (int)(it - nodes.begin()) : -1;
    };
    // 2. unions: row yb is the top halo row of rank r and the first owned row of rank r + 1
    std::vector<int> parent(nodes.size());
    for (size_t k = 0; k < nodes.size(); ++k) parent[k] = (int)k;
    for (int r = 0; r + 1 < nranks; ++r) {
        const int64_t* lo = gathered + r * W + L.rowB();
        const int64_t* hi = gathered + (r + 1) * W + L.rowA();
        for (int x = 0; x < m; ++x) {
            if ((lo[x] != 0) != (hi[x] != 0)) { out->error = 1; return; }
            if (!lo[x]) continue;
            if (x > 0 && lo[x] == lo[x - 1] && hi[x] == hi[x - 1]) continue;      // same pair as the column before
            int a = uf_find(parent, index_of(node_key(r, lo[x]))), b = uf_find(parent, index_of(node_key(r + 1, hi[x])));
            if (a != b) parent[a > b ? a : b] = a > b ? b : a;
        }
    }
    // 3. classes: canonical label = smallest member id; size = sum of the rank-local sizes
    const int64_t MASK40 = ((int64_t)1 << 40) - 1;
    std::vector<int64_t> cgid(nodes.size(), 0), ctot(nodes.size(), 0);
    std::vector<int> cpos(nodes.size(), 0);                      // members with a non-zero rank-local size
    for (size_t k = 0; k < nodes.size(); ++k) {
        int c = uf_find(parent, (int)k);
        int64_t gid = nodes[k].key & MASK40;
        if (cgid[c] == 0 || gid < cgid[c]) cgid[c] = gid;
        ctot[c] += nodes[k].size;
        cpos[c] += nodes[k].size > 0;
    }
    // 4. lattice-wide summary
    int64_t best_size = 0, best_gid = 0;
    auto consider = [&](int64_t size, int64_t gid) {
        if (size > best_size || (size == best_size && size > 0 && gid < best_gid)) { best_size = size; best_gid = gid; }
    };
    for (int r = 0; r < nranks; ++r) {
        const int64_t* sc = gathered + r * W + L.scalars();
        out->ncl += sc[0];
        out->nlone += sc[1];
        // the rank's largest cluster; if it is an interface cluster its class (below) supersedes it
        int k = index_of(node_key(r, sc[3]));
        if (k < 0) consider(sc[2], sc[3]);
    }
    for (size_t c = 0; c < nodes.size(); ++c) {
        if (parent[c] != (int)c) continue;
        out->ncl -= cpos[c] - (ctot[c] > 0 ? 1 : 0);              // the members were counted once per rank
        consider(ctot[c], cgid[c]);
    }
    out->maxcs = best_size; out->maxgid = best_gid;
    // 5. spanning: the class reaches row 0 (canonical label <= m) and a site of the top row
    {
        const int64_t* top = gathered + (nranks - 1) * W + L.rowTop();
        std::vector<int> cls;
        for (int x = 0; x < m; ++x) {
            if (!top[x] || (x > 0 && top[x] == top[x - 1])) continue;
            int k = index_of(node_key(nranks - 1, top[x]));
            if (k < 0) continue;
            int c = uf_find(parent, k);
            if (cgid[c] <= m) cls.push_back(c);
        }
        std::sort(cls.begin(), cls.end());
        cls.erase(std::unique(cls.begin(), cls.end()), cls.end());
        std::vector<std::pair<int64_t, int64_t>> sp;
        for (int c : cls) sp.push_back({cgid[c], ctot[c]});
        std::sort(sp.begin(), sp.end());
        for (auto& e : sp) { out->span_gid.push_back(e.first); out->span_size.push_back(e.second); }
    }
    // 6. this rank's interface clusters: representative = its smallest root of the class
    std::unordered_map<int, int64_t> rep;                        // class -> smallest root id of this rank
    const int64_t klo = node_key(rank, 0), khi = node_key(rank + 1, 0);
    for (size_t k = 0; k < nodes.size(); ++k) {
        if (nodes[k].key < klo || nodes[k].key >= khi) continue;
        int c = uf_find(parent, (int)k);
        int64_t gid = nodes[k].key & MASK40;
        auto it = rep.find(c);
        if (it == rep.end() || gid < it->second) rep[c] = gid;
    }
    for (size_t k = 0; k < nodes.size(); ++k) {
        if (nodes[k].key < klo || nodes[k].key >= khi) continue;
        int c = uf_find(parent, (int)k);
        out->root_gid.push_back(nodes[k].key & MASK40);
        out->rep_gid.push_back(rep[c]);
        out->class_gid.push_back(cgid[c]);
        out->class_total.push_back(ctot[c]);
    }
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

// point every interface root of a class at the class representative of this rank
__global__ void slab_relabel_kernel(int npairs, const int32_t* __restrict__ root, const int32_t* __restrict__ rep, int32_t* __restrict__ label)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < npairs && root[k] != rep[k]) label[root[k]] = rep[k] + 1;
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

// all-gather of the interface rows, redundant host union-find, relabel
int slab_stitch(Ctx* c)
{
    const Geom& g = c->g;
    IfaceLayout L{g.m};
    const int64_t W = L.words();
    cudaStream_t st = c->stream;
    if (!c->d_iface) {
        PERC_CUDA(cudaMalloc(&c->d_iface, sizeof(int64_t) * W * (c->nranks + 1)));
        PERC_CUDA(cudaMallocHost(&c->h_iface, sizeof(int64_t) * W * c->nranks));
    }
    int64_t* mine = c->d_iface + W * c->nranks;
    slab_iface_kernel<<<nblk(g.m), 256, 0, st>>>(g, c->rank, c->nranks, c->label, c->size, mine);
    slab_scalars_kernel<<<1, 1, 0, st>>>(g, c->d_sum, mine);
    c->launches += 2;
    if (c->nranks > 1) {
        if (!c->comm) return -4;
        PERC_NCCL(g_nccl.AllGather(mine, c->d_iface, (size_t)W, NCCL_INT64, (NcclComm)c->comm, st));
    } else PERC_CUDA(cudaMemcpyAsync(c->d_iface, mine, sizeof(int64_t) * W, cudaMemcpyDeviceToDevice, st));
    PERC_CUDA(cudaMemcpyAsync(c->h_iface, c->d_iface, sizeof(int64_t) * W * c->nranks, cudaMemcpyDeviceToHost, st));
    PERC_CUDA(cudaStreamSynchronize(st));
    StitchResult& R = c->stitch;
    stitch_host(c->nranks, c->rank, g.m, c->h_iface, &R);
    if (R.error) return -8;
    // relabel: interface roots of one class -> the class representative of this rank
    const int np = (int)R.root_gid.size();
    const int64_t off = (int64_t)g.y0 * g.m;
    c->tab_rep.clear(); c->tab_gid.clear(); c->tab_total.clear();
    if (np) {
        std::vector<int32_t> h(2 * (size_t)np);
        for (int k = 0; k < np; ++k) { h[k] = (int32_t)(R.root_gid[k] - off - 1); h[np + k] = (int32_t)(R.rep_gid[k] - off - 1); }
        int32_t* d = (int32_t*)ctx_dev_stage(c, sizeof(int32_t) * 2 * np + sizeof(int64_t) * np + 64);
        if (!d) return (int)cudaErrorMemoryAllocation;
        PERC_CUDA(cudaMemcpyAsync(d, h.data(), sizeof(int32_t) * 2 * np, cudaMemcpyHostToDevice, st));
        slab_relabel_kernel<<<nblk(np), 256, 0, st>>>(np, d, d + np, c->label);
        slab_flatten_kernel<<<nblk(g.t), 256, 0, st>>>(g.t, c->label);
        c->launches += 2;
        PERC_CUDA(cudaStreamSynchronize(st));
        // table rep -> (class label, class size), ascending rep
        std::vector<std::pair<int32_t, int>> reps;
        for (int k = 0; k < np; ++k) if (R.root_gid[k] == R.rep_gid[k]) reps.push_back({(int32_t)(R.rep_gid[k] - off - 1), k});
        std::sort(reps.begin(), reps.end());
        for (auto& e : reps) { c->tab_rep.push_back(e.first); c->tab_gid.push_back(R.class_gid[e.second]); c->tab_total.push_back(R.class_total[e.second]); }
    }
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
