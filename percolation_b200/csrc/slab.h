// slab.h -- one lattice decomposed into row slabs over several GPUs (mode 2).
//
// Every rank labels its slab (+ one halo row on each inner side) on its own; the clusters are then
// stitched across the G-1 interfaces: one NCCL all-gather of each rank's interface rows (the label of
// every site as a lattice-wide id, and the rank-local size of that cluster), after which EVERY rank
// runs the same small union-find over the interface labels on its GPU (slab_stitch.cuh: hash table of
// the (rank, cluster) nodes, atomicMin unions; redundantly on every rank: no second exchange, no
// iteration until quiescence) and relabels its own sites.  The Kirchhoff solve exchanges one halo row
// of the residual per iteration with the slab below / above (ncclSend / ncclRecv over NVLink) and
// all-reduces the two dot products.
#pragma once
#include <stdint.h>
#include <vector>
#include "geometry.cuh"

namespace perc {

// layout of the block a rank contributes to the all-gather (int64 words)
struct IfaceLayout {
    int m;
    PERC_HD int64_t rowA() const { return 0; }                // ids of the FIRST OWNED row (row ya); used for ranks >= 1
    PERC_HD int64_t rowB() const { return m; }                // ids of the top halo row (row yb); used for ranks < G-1
    PERC_HD int64_t sizeA() const { return 2 * (int64_t)m; }  // rank-local sizes (owned rows only) of those clusters
    PERC_HD int64_t sizeB() const { return 3 * (int64_t)m; }
    PERC_HD int64_t rowTop() const { return 4 * (int64_t)m; } // ids of the lattice's top row (last rank only)
    PERC_HD int64_t scalars() const { return 5 * (int64_t)m; }   // [0] ncl  [1] nlone  [2] max size  [3] id of the largest
    PERC_HD int64_t words() const { return 5 * (int64_t)m + 8; }
};

struct StitchResult {
    // for the calling rank: every interface cluster it holds
    std::vector<int64_t> root_gid;     // lattice-wide id of the rank-local root (ascending)
    std::vector<int64_t> rep_gid;      // id of the rank-local root that represents the class on this rank
    std::vector<int64_t> class_gid;    // canonical label of the class
    std::vector<int64_t> class_total;  // size of the class
    // lattice-wide summary (identical on every rank)
    int64_t ncl = 0, nlone = 0, maxcs = 0, maxgid = 0;
    std::vector<int64_t> span_gid, span_size;       // spanning clusters, ascending id
    int error = 0;                                   // != 0: the ranks disagree on the occupancy of an interface row
};

// the stitch's item functions (slab_stitch.cuh) run one by one on the host: exported through the C-ABI as
// perc_stitch_host so that the CPU tests exercise the very code the kernels execute
void stitch_host(int nranks, int rank, int m, const int64_t* gathered, StitchResult* out);

}  // namespace perc
