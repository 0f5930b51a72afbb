// context.h -- per-handle state of libperc_b200 (host side, C++).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "geometry.cuh"
#include "ccl_tile.cuh"
#include "slab.h"

namespace perc {

constexpr int RANK_NONE = 0x7fffffff;

// PCG scalars living on the device (one solve at a time per handle)
struct PcgState {
    double bknum, bkden, akden, ak, bk, rr, bnrm, err, tol;
    double Itop, Ibot;
    double red[4];                     // slab mode: rank-local sums waiting for the all-reduce
    int iter, itmax, done;
    unsigned int ticket_a, ticket_b;   // last-block-done counters
};

struct PhiloxThreshold {
    unsigned long long key;   // occupied iff elem_key < key || (elem_key == key && id <= id)
    unsigned long long id;
    int enabled;              // 0: nothing occupied
    int all;                  // 1: everything occupied
    int failed;               // device-resident selection (batches) missed: the realization is skipped, not counted
    int pad;
};

enum OccSource { SRC_NONE = 0, SRC_RANK = 1, SRC_PHILOX = 2 };

struct Ctx {
    Geom g;
    int device = 0;
    int num_sms = 148;

    // ---- slab decomposition (SURVEY 8e mode 2); nranks = 1: the whole lattice lives here
    int nranks = 1, rank = 0;
    void* comm = nullptr;                // ncclComm_t
    int stat_nranks = 1;                 // ranks of a communicator used only to reduce statistics (mode 1)
    int64_t* d_iface = nullptr;          // [nranks + 1] interface blocks (all-gather target + own block)
    int64_t* h_iface = nullptr;          // pinned
    void* d_stitch = nullptr;            // hash table + outputs of the device-side stitch
    size_t d_stitch_bytes = 0;
    StitchResult stitch;
    std::vector<int32_t> tab_rep;        // interface classes held by this rank: representative root (ascending)
    std::vector<int64_t> tab_gid, tab_total;   //   -> lattice-wide label and size
    std::vector<int64_t> h_span_gid, h_span_total;
    cudaStream_t stream = nullptr;

    // ---- occupancy inputs
    int32_t* srank = nullptr;     // [t]        rank of the site in the fill order
    int32_t* brank = nullptr;     // [ndir][t]  rank of the bond owned by (site, dir)
    int site_src = SRC_NONE, bond_src = SRC_NONE;
    int64_t ks = 0, kb = 0;
    unsigned long long seed = 0, stream_id = 0;
    PhiloxThreshold thr_site{}, thr_bond{};
    PhiloxThreshold* d_thr = nullptr;    // [2] batch mode: thresholds selected on the device
    bool batch_thr = false;              // the mask builder reads d_thr instead of thr_site / thr_bond

    // ---- realization state
    uint8_t* mask = nullptr;      // [t]
    uint8_t* mask_prev = nullptr; // [t] incremental re-labeling: the mask of the fill the labels on the device belong to
    // what the labels on the device were made from (ccl_note_labeled): perc_label_incremental advances them when only elements were added
    bool lab_valid = false;
    int lab_kind = 0, lab_site_src = SRC_NONE, lab_bond_src = SRC_NONE;
    int64_t lab_ks = 0, lab_kb = 0;
    unsigned long long lab_seed = 0, lab_stream = 0, lab_epoch = 0;
    unsigned long long occ_epoch = 1;     // bumped whenever an order / occupancy array is uploaded or a batch consumed the inputs
    int32_t* label = nullptr;     // [t]  parent+1 during CCL, canonical label after flatten
    int32_t* size = nullptr;      // [t]  cluster size at the ROOT's index (label-1); other entries undefined
    int32_t* rootlist = nullptr;  // [t]  site indices of the tile-local roots of the last labeling
    Summary* d_sum = nullptr;
    Summary* h_sum_pin = nullptr; // pinned mirror
    Summary h_sum{};
    std::vector<int32_t> h_span_ids, h_span_sizes;
    int kind = 0;
    bool labeled = false;

    // ---- conductance
    uint8_t* cfull = nullptr;     // [t] conducting-neighbour bits (8 directions)
    double* vx = nullptr;         // [t] solution (rows 0 and n-1 unused)
    double* vr = nullptr;
    double* vp = nullptr;
    double* vp2 = nullptr;        // p is double-buffered (the fused SpMV reads neighbours' old p)
    double* vq = nullptr;
    double* xprow = nullptr;      // [4 m] one-pass solver: x and p of rows 1 and n-2 (the rows the read-out consumes)
    double* bond_w = nullptr;     // [nb] caller's per-bond conductances (perc_set_bond_conductance; pcg_weighted.cu)
    double* wplane = nullptr;     // [4][t] weight of the bond each site owns in direction E, N, NW, NE
    double* wdiag = nullptr;      // [t] diagonal of the weighted matrix
    bool have_bond_w = false;
    int pcg_mode = -1;            // -1 process default, 0 automatic (one-pass kernel when it applies), 1 two-kernel form
    int fused_cfg = -1;           // tile configuration of the one-pass kernel (-1: process default)
    bool last_fused = false;      // the last solve ran the one-pass kernel
    int last_fused_cfg = -1;      // ... in this tile configuration (4 = deflated)
    double* d_defl = nullptr;     // deflation: E^-1 [k*k], mu [KMAX], nu [KMAX], F [ntiles*8], W [ntiles*8]
    double* h_defl = nullptr;     // pinned: E^-1, W
    size_t defl_bytes = 0;
    int defl_k = 0;               // coarse dimension of the last deflated solve
    int* d_sched = nullptr;       // deflated sweep: which CTA walks which blocks (ft_defl_schedule)
    int sched_key[6] = {0, 0, 0, 0, 0, 0};
    double* partial = nullptr;    // per-block partial sums
    int partial_cap = 0;
    PcgState* d_pcg = nullptr;
    PcgState* h_pcg = nullptr;    // pinned
    bool solved = false;
    bool have_x = false;          // the last solve kept the interior voltages (perc_conduct, not perc_conduct_g)

    // ---- selection scratch
    unsigned long long* d_hist = nullptr;   // [4096 + 16]: window histogram, below-count, SelState
    unsigned long long* d_cand = nullptr;   // candidate (key,id) pairs
    int cand_cap = 0;

    // ---- host staging
    void* h_stage = nullptr;      // pinned staging buffer
    size_t h_stage_bytes = 0;
    void* d_stage = nullptr;
    size_t d_stage_bytes = 0;

    // ---- batches: child contexts (own arrays + stream) so that small lattices overlap on the GPU
    std::vector<Ctx*> batch_kids;

    // ---- per-device function attributes already set through this handle (bit sets)
    unsigned ccl_attr = 0, pcg_attr = 0;

    // ---- instrumentation
    int64_t launches = 0;
    float phase_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaEvent_t ev[12] = {};
};

// ---- implemented in the .cu files -----------------------------------------------------------
int ctx_alloc(Ctx* c);
void ctx_free(Ctx* c);
void* ctx_host_stage(Ctx* c, size_t bytes);
int ctx_ensure_ranks(Ctx* c, bool sites, bool bonds);
void* ctx_dev_stage(Ctx* c, size_t bytes);

int occ_upload_site_order(Ctx* c, const int32_t* order);
int occ_upload_bond_order(Ctx* c, const int32_t* border);
int occ_upload_flags(Ctx* c, const uint8_t* socc, const uint8_t* bocc);
int occ_generate(Ctx* c, unsigned long long seed, unsigned long long stream, int64_t ks, int64_t kb);
int occ_generate_dev(Ctx* c, unsigned long long seed, unsigned long long stream, int64_t ks, int64_t kb,
                     unsigned long long* d_nfail);
int occ_export(Ctx* c, uint8_t* socc, uint8_t* bocc);
int occ_build_mask(Ctx* c, int kind);

int ccl_run(Ctx* c, int kind);
int ccl_launch(Ctx* c, int kind);        // the labeling pipeline without the final host synchronisation
int batch_run(Ctx* c, int kind, int nreal, unsigned long long seed, unsigned long long stream0, int64_t ks, int64_t kb,
              int nbins, int64_t* hist, int64_t* stats);
int ccl_fetch_summary(Ctx* c);
void ccl_span_launch(Ctx* c);            // K5 on the handle's stream
void ccl_note_labeled(Ctx* c, int kind);  // remember what the labels on the device were made from
bool ccl_incremental_applies(const Ctx* c, int kind);
int ccl_incremental_run(Ctx* c, int kind);
int ccl_hist(Ctx* c, int nbins, int64_t* hist, int logbin = 0);
int ccl_export_bond_labels(Ctx* c, int32_t* b3);
int ccl_export_sizes(Ctx* c, int32_t* cs);

int slab_unique_id(void* id128);
int slab_comm_init(Ctx* c, const void* id128);
void slab_comm_destroy(Ctx* c);
int slab_allreduce_f64(Ctx* c, double* buf, int count);
int slab_allreduce_i64(Ctx* c, int64_t* buf, int count);
int slab_halo_exchange(Ctx* c, void* array, int bytes_per_site);
int slab_count_roots(Ctx* c);
int slab_stitch(Ctx* c);
int32_t slab_local_label_of(const Ctx* c, int64_t gid);
int slab_export_labels(Ctx* c, int64_t* out);

bool pcg_small_fits(const Geom& g);
int pcg_small_stage(Ctx* c, uint8_t* cfbatch, int slot);
int pcg_small_solve(Ctx* c, const uint8_t* cfbatch, int nreal, double Va, double g0, double gleak, double tol, int itmax,
                    double read_thresh, double* d_G, int* d_iters, double* d_errs);
int batch_conduct_run(Ctx* c, int kind, int nreal, unsigned long long seed, unsigned long long stream0, int64_t ks, int64_t kb,
                      double Va, double g0, double gleak, double tol, int itmax, double read_thresh,
                      double* G, int32_t* iters, int64_t* stats);
bool pcg_fused_applies(const Ctx* c, int keep_x, int warm);
int pcg_set_bond_weights(Ctx* c, const double* w);
int pcg_solve_weighted(Ctx* c, double Va, double gleak, double tol, int itmax, double read_thresh);
int pcg_solve(Ctx* c, int cluster_id, double Va, double g0, double gleak, double tol, int itmax,
              double read_thresh, int keep_x, double* Gtop, double* Gbot, int* iter, double* err, int warm = 0);

#define PERC_CUDA(call)                                              \
    do {                                                             \
        cudaError_t e__ = (call);                                    \
        if (e__ != cudaSuccess) return (int)e__;                     \
    } while (0)

}  // namespace perc
