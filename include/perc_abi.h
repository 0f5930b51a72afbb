/*
 * perc_abi.h -- C-ABI of libperc_b200.so, the B200 (sm_100a) replacement for the
 * percolation-realization hot path of IsaiahSteinke/Percolation.
 *
 * The reference has no FFI: it is 20 monolithic Fortran-77 PROGRAMs whose hot path is
 * inline code blocks.  Each entry point below replaces one of those blocks (cited as
 * file:line into the reference's Fortran/ tree, Sq = Square, Tri = Triangular) so the
 * Fortran drivers keep their structure and call this library through ISO_C_BINDING
 * (fortran/perc_iface.f90; binding stub in INTEGRATION.md).
 *
 * Conventions (Fortran-callable, bind(C) without VALUE):
 *   - every scalar is passed BY REFERENCE; INTEGER <-> int32_t, DOUBLE PRECISION <-> double;
 *   - sites are 1-based, row-major rn = y*m + x (x = 1..m fastest; row y = 0 is the grounded
 *     bottom edge, y = n-1 the top edge at Va)                      Sq/site.f:371-469
 *   - 2-D arrays are column-major: border(nb,2) = nb "lo" ends followed by nb "hi" ends
 *     Sq/bond.f:40,112-129
 *   - bond row numbering is the reference's own (do i=1,t-1; do j=1,scn; if nn(j)>i)
 *   - labels are CANONICAL: label = smallest member id (site rn; a bond whose two ends are
 *     both unoccupied in a mixed problem is its own cluster, label t + row).  The reference's
 *     labels are creation counters; its partition and c() sizes are identical.
 *   - caller owns every host array; the library owns device memory behind the handle;
 *   - return value: 0 ok; < 0 bad argument (PERC_E_*); > 0 a cudaError_t code.  Never aborts
 *     (the reference's convention is `pause`, Sq/bondc.f:737,777).
 *   - one handle per host thread / GPU; calls on a handle are serialised by the caller.
 * There is no CPU fallback: without a CUDA device perc_create fails with the CUDA error.
 */
#ifndef PERC_ABI_H
#define PERC_ABI_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define PERC_SQUARE      1   /* lattype, MATLAB/ConductCalc.m:30 */
#define PERC_TRIANGULAR  2
#define PERC_SITE        1   /* perctype, MATLAB/ConductCalc.m:29 */
#define PERC_BOND        2
#define PERC_MIXED       3

#define PERC_E_ARG       (-1)  /* bad argument                                  */
#define PERC_E_ODD_M     (-2)  /* triangular lattice needs even m  */
#define PERC_E_HANDLE    (-3)  /* unknown handle                                */
#define PERC_E_STATE     (-4)  /* call order (e.g. conduct before label)        */
#define PERC_E_SIZE      (-5)  /* lattice too large for 32-bit labels           */
#define PERC_E_NOSPAN    (-6)  /* requested cluster does not span               */
#define PERC_E_NCCL      (-7)  /* libnccl.so.2 could not be loaded (slab mode)    */
#define PERC_E_IFACE     (-8)  /* ranks disagree on an interface row (slab mode) */
#define PERC_E_SELECT    (-9)  /* batch: the device-side exact-count selection missed for stats(7) realizations;
                                   they are excluded from every statistic (re-run them singly with perc_generate) */

/* ---- geometry (host-only index arithmetic; no device needed) ------------------------------ */
/* nb formulas Sq/site.f:89-93, Tri/site.f:91-95 */
int32_t perc_geom_nb(const int32_t *lattice, const int32_t *m, const int32_t *n, const int32_t *pbc, int32_t *nb);
/* bond list b(nb,2) in the reference's numbering, Sq/bond.f:112-129 (= bondlist.txt, Sq/site.f:106-124) */
int32_t perc_geom_bondlist(const int32_t *lattice, const int32_t *m, const int32_t *n, const int32_t *pbc, int32_t *b);
/* neighbours of site rn in nearestn order, 0 = none; nn has 6 entries. Sq/site.f:371-469, Tri/site.f:373-558 */
int32_t perc_geom_nearestn(const int32_t *lattice, const int32_t *m, const int32_t *n, const int32_t *pbc,
                           const int32_t *rn, int32_t *nn);

/* ---- handle: the parameter block m, n, pbc, lattice (Sq/site.f:51-67) --------------------- */
int32_t perc_create(int64_t *h, const int32_t *lattice, const int32_t *m, const int32_t *n,
                    const int32_t *pbc, const int32_t *device);
int32_t perc_destroy(const int64_t *h);
int32_t perc_sync(const int64_t *h);

/* ---- occupancy inputs ------------------------------------------------------------------------ */
/* order(t): the shuffled site order of Sq/site.f:131-147; sites order(1..k) get occupied (:164-176).
 * Uploads the order once (as a rank table) -- later calls may re-threshold with perc_set_fill. */
int32_t perc_set_site_order(const int64_t *h, const int32_t *order);
/* border(nb,2): the shuffled bond order of Sq/bond.f:137-150 */
int32_t perc_set_bond_order(const int64_t *h, const int32_t *border);
/* fill counts ks = int(ps*t), kb = int(pb*nb) (Sq/site.f:164, Sq/bond.f:167); -1 keeps the current one */
int32_t perc_set_fill(const int64_t *h, const int32_t *ks, const int32_t *kb);
/* direct occupancy flags (0/1 bytes): socc(t) by site, bocc(nb) by reference bond row; either may be NULL */
int32_t perc_set_occupancy(const int64_t *h, const uint8_t *socc, const uint8_t *bocc);
/* B200 occupancy generator replacing permute.f + rand (Fortran/permute.f:39-44): counter-based
 * Philox-4x32-10 keys per element, EXACTLY ks (kb) smallest keys occupied (ties by element id).
 * stream: realization index.  -1 for ks / kb skips that element type. */
int32_t perc_generate(const int64_t *h, const int64_t *seed, const int64_t *stream,
                      const int32_t *ks, const int32_t *kb);
/* read back the occupancy in reference layout: socc(t) bytes, bocc(nb) bytes (either may be NULL) */
int32_t perc_get_occupancy(const int64_t *h, uint8_t *socc, uint8_t *bocc);

/* ---- labeling + sizes + spanning (Sq/site.f:162-344, bond.f:165-432, sitebond.f:187-458) ---- */
/* kind = PERC_SITE | PERC_BOND | PERC_MIXED.  Results stay on the device until fetched. */
int32_t perc_label(const int64_t *h, const int32_t *kind);
/* summary: ncl = number of clusters, maxcs / maxcn = largest cluster size and its (canonical) label
 * (Sq/site.f:278-287), nspan = number of spanning clusters */
int32_t perc_summary(const int64_t *h, int64_t *ncl, int32_t *maxcs, int32_t *maxcn, int32_t *nspan);
int32_t perc_get_site_labels(const int64_t *h, int32_t *s);           /* s(t)            */
int32_t perc_get_bond_labels(const int64_t *h, int32_t *b3);          /* b(nb,3) column 3 */
int32_t perc_get_sizes(const int64_t *h, int32_t *c);                 /* c(t): c(label) = size, else 0 */
/* spanning clusters, ascending canonical id; ids[0] is the default choice for perc_conduct.
 * (reference: lowest historical label that spans, Sq/site.f:309-344 -- SURVEY F7) */
int32_t perc_span(const int64_t *h, const int32_t *max_ids, int32_t *nspan, int32_t *ids, int32_t *sizes);
/* exact cluster-size histogram: hist(s) = number of clusters of size s for s = 1..nbins-1,
 * hist(nbins) = clusters of size >= nbins.  (multiset of non-zero c()) */
int32_t perc_hist(const int64_t *h, const int32_t *nbins, int64_t *hist);
/* log-binned: hist(b+1) = clusters with 2^b <= size < 2^(b+1), the last bin takes the rest (n_s plots) */
int32_t perc_hist_log2(const int64_t *h, const int32_t *nbins, int64_t *hist);

/* reference-shaped one-call wrappers (upload order, label, download) */
int32_t perc_site(const int64_t *h, const int32_t *order, const int32_t *k,
                  int32_t *s, int32_t *c, int32_t *maxcs, int32_t *perccln, int32_t *perccls);
int32_t perc_bond(const int64_t *h, const int32_t *border, const int32_t *k,
                  int32_t *b3, int32_t *c, int32_t *maxcs, int32_t *perccln, int32_t *perccls);
int32_t perc_sitebond(const int64_t *h, const int32_t *sorder, const int32_t *ks,
                      const int32_t *border, const int32_t *kb,
                      int32_t *s, int32_t *b3, int32_t *c, int32_t *maxcs, int32_t *perccln, int32_t *perccls);

/* first-spanning search of the *_perc drivers (Sq/site_perc.f:133-254, bond_perc.f, sb_perc.f,
 * bs_perc.f): smallest fill count of `which` (PERC_SITE or PERC_BOND elements) at which a spanning
 * cluster exists, the other count held fixed; kstar = 0 if it never spans.  f = real(kstar)/real(N). */
int32_t perc_first_span(const int64_t *h, const int32_t *kind, const int32_t *which,
                        int32_t *kstar, float *f, int32_t *maxcs, int32_t *perccls);

/* ---- Kirchhoff conductance (Sq/bondc.f:465-595 + linbcg :750-838) ----------------------------- */
/* cluster_id: canonical id of a spanning cluster (0 = ids[0]).  Reference defaults: Va = g0 = 1
 * (:85-86), gleak = 1e-12 (:487), tol = 1e-8, itmax = 2500 (:545), read_thresh = 1e-10 (:576). */
int32_t perc_conduct(const int64_t *h, const int32_t *cluster_id, const double *Va, const double *g0,
                     const double *gleak, const double *tol, const int32_t *itmax, const double *read_thresh,
                     double *Gtop, double *Gbot, int32_t *iter, double *err);
/* perc_conduct with the voltages of the handle's previous perc_conduct / perc_conduct_warm as the initial
 * guess -- linbcg's own `x` is in/out (Sq/bondc.f:759-763, r = b - A x); the reference driver zeroes it before
 * every sweep point (Sq/bond_cond.f:400-405), successive points of a p-sweep differ by 0.5 % of the bonds.
 * Same converged answer (to tol); falls back to x = 0 when there is no previous solution. */
int32_t perc_conduct_warm(const int64_t *h, const int32_t *cluster_id, const double *Va, const double *g0,
                          const double *gleak, const double *tol, const int32_t *itmax, const double *read_thresh,
                          double *Gtop, double *Gbot, int32_t *iter, double *err);
/* same solve for the p-sweep drivers (Sq/bond_cond.f:392-485), which only write pb, Gbot, Gtop, avg
 * (:481-482): the iterate x is kept on rows 1 and n-2 only -- the rows the read-out G~.V (:576-592)
 * consumes -- so Gtop / Gbot / iter / err are bit-identical to perc_conduct while the interior
 * voltages are not formed (16 B per site and iteration less HBM traffic).  perc_get_voltage is not
 * available after this call (PERC_E_STATE). */
int32_t perc_conduct_g(const int64_t *h, const int32_t *cluster_id, const double *Va, const double *g0,
                       const double *gleak, const double *tol, const int32_t *itmax, const double *read_thresh,
                       double *Gtop, double *Gbot, int32_t *iter, double *err);
int32_t perc_get_voltage(const int64_t *h, double *Vint);             /* Vint(t-2m), Sq/bondc.f:545 */

/* ---- one lattice decomposed into row slabs over several GPUs (north star: L = 65536 over 8 x B200) ----
 * The reference has no counterpart (serial program, t <= 10^6, Sq/site.f:37); the entry points mirror
 * the single-GPU ones.  One process per GPU; rank r of nranks holds rows [n*r/nranks, n*(r+1)/nranks) of
 * the m x n lattice plus one halo row on each inner side.  The occupancy comes from the generator
 * (element keys are functions of the lattice-wide element id, so the realization is the one the
 * single-GPU handle draws); perc_label labels the slab, all-gathers the interface rows over NCCL and
 * stitches the clusters with a union-find every rank runs redundantly on its GPU; perc_conduct_g exchanges one halo
 * row of the residual per iteration (ncclSend/ncclRecv) and all-reduces the two dot products.
 * Lattice-wide labels / sizes / counts are 64-bit (t = 2^32 at L = 65536). */
int32_t perc_create_slab(int64_t *h, const int32_t *lattice, const int32_t *m, const int32_t *n,
                         const int32_t *pbc, const int32_t *device, const int32_t *nranks, const int32_t *rank);
/* NCCL bootstrap: rank 0 calls perc_comm_unique_id, the host distributes the 128 bytes (MPI_Bcast in a
 * Fortran driver), every rank calls perc_comm_init.  NCCL is loaded at run time (libnccl.so.2). */
int32_t perc_comm_unique_id(uint8_t *id128);
int32_t perc_comm_init(const int64_t *h, const uint8_t *id128);
int32_t perc_slab_rows(const int64_t *h, int32_t *ya, int32_t *yb);     /* owned rows [ya, yb), 0-based */
int32_t perc_generate_i8(const int64_t *h, const int64_t *seed, const int64_t *stream,
                         const int64_t *ks, const int64_t *kb);         /* perc_generate with 64-bit fill counts */
int32_t perc_summary_i8(const int64_t *h, int64_t *ncl, int64_t *maxcs, int64_t *maxcn, int64_t *nspan);
int32_t perc_span_i8(const int64_t *h, const int32_t *max_ids, int32_t *nspan, int64_t *ids, int64_t *sizes);
/* s((yb-ya)*m): lattice-wide canonical labels of the rows this rank owns */
int32_t perc_get_site_labels_i8(const int64_t *h, int64_t *s);
/* the stitch on its own, its device code executed item by item on the host (no device needed; CPU tests):
 * gathered = nranks blocks of 5*m+8 words
 * (layout: percolation_b200/csrc/slab.h); out_summary(5) = ncl, nlone, maxcs, maxcn, nspan; out_pairs(4,k) =
 * (root id, representative id, class label, class size) of every interface cluster the calling rank holds */
int32_t perc_stitch_host(const int32_t *nranks, const int32_t *rank, const int32_t *m, const int64_t *gathered,
                         int64_t *out_summary, const int32_t *max_span, int64_t *span_ids, int64_t *span_sizes,
                         const int32_t *max_pairs, int32_t *npairs, int64_t *out_pairs);

/* ---- batches of independent realizations: the trial loops `do ii = 1, numtrials` of the *_perc and
 * bond_cond drivers (Sq/site_perc.f:87, Sq/bond_perc.f, Sq/sb_perc.f:104, Sq/bond_cond.f:123) ------------
 * Realization i = 0..nreal-1 draws stream = stream0 + i from the generator with exact fill counts ks / kb,
 * is labeled, and is folded into device-resident statistics; nothing is synchronised inside the loop.
 * hist(nbins): cluster-size histogram summed over the batch (layout of perc_hist); stats(16), int64:
 *  (1) realizations (2) sum ncl (3) sum maxcs (4) realizations that span (5) sum nspan (6) sum perccls
 *  (7) failed selections -- must be 0 -- (8) sum maxcs^2 (9) sum occupied sites (10) sum occupied bonds.
 * The handle's occupancy input is consumed: call perc_generate / perc_set_*_order again before perc_label. */
int32_t perc_batch(const int64_t *h, const int32_t *kind, const int32_t *nreal, const int64_t *seed,
                   const int64_t *stream0, const int32_t *ks, const int32_t *kb, const int32_t *nbins,
                   int64_t *hist, int64_t *stats);
/* the same batch with the Kirchhoff conductance of every realization's default spanning cluster (trial loop of
 * Sq/bond_cond.f:123-498 at one fill; parameters as perc_conduct).  G(2,nreal) = Gtop, Gbot; iters(nreal) (-1: the
 * realization does not span, G = 0).  Lattices of t <= 13312 sites are solved one CTA per realization in a single
 * launch (the whole linbcg loop inside the kernel); larger ones realization by realization. */
int32_t perc_batch_conduct(const int64_t *h, const int32_t *kind, const int32_t *nreal, const int64_t *seed,
                           const int64_t *stream0, const int32_t *ks, const int32_t *kb, const double *Va,
                           const double *g0, const double *gleak, const double *tol, const int32_t *itmax,
                           const double *read_thresh, double *G, int32_t *iters, int64_t *stats);
/* independent realizations sharded over GPUs (one process per GPU, no data-path collective): communicator
 * for the single reduction of the statistics at the end, and that reduction (integer sums are order
 * independent: the result is the same for any number of GPUs) */
int32_t perc_comm_init_rank(const int64_t *h, const int32_t *nranks, const int32_t *rank, const uint8_t *id128);
int32_t perc_allreduce_stats(const int64_t *h, const int32_t *ni, int64_t *ivals, const int32_t *nd, double *dvals);

/* ---- re-labeling along a sweep -----------------------------------------------------------
 * The reference's sweep drivers add elements one by one to a labeled lattice (Sq/site_perc.f:133-254, Sq/bond_cond.f:208-485,
 * Sq/sb_perc.f, Sq/bs_perc.f).  After perc_set_fill raised the fill counts of a handle that holds the labels of a SMALLER
 * fill of the same order / generator stream (same kind, one GPU), perc_label_incremental unites only the added elements on the
 * label table, folds the sizes of the clusters that merged and adds the new elements' weights -- labels, sizes, counts and
 * spanning clusters are identical to perc_label's.  In every other case it runs perc_label's full pass.
 * incremental (may be NULL): 1 = the incremental pass ran, 0 = the full one. */
int32_t perc_label_incremental(const int64_t *h, const int32_t *kind, int32_t *incremental);

/* ---- the reference programs' output files --------------------------------------
 * Writes one of the text files the reference's programs leave behind, in their own record formats, from the handle's
 * current labeling, so that MATLAB/SitePlot.m, BondPlot.m, SiteBondPlot.m and ConductCalc.m read GPU results unchanged:
 *   which = 1 site.txt     j, s(j), c(j)                     (i10,",",i10,",",i10)                    Sq/site.f:354-359
 *           2 bond.txt     b(j,1), b(j,2), b(j,3), j, c(j)   (i10,",",i10,",",i10,",",i10,",",i10)    Sq/bond.f:443-448
 *           3 sbsite.txt   i, s(i), c(i)                                                               Sq/sitebond.f:469-471
 *           4 sbbond.txt   b(i,1), b(i,2), b(i,3)                                                      Sq/sitebond.f:473-475
 *           5 bondlist.txt blist(i,1), blist(i,2)            (i10,",",i10)                             Sq/site.f:106-120
 * 1 needs a site labeling, 2 a bond labeling, 3 / 4 a mixed one (PERC_E_STATE otherwise).  Labels are canonical
 * (smallest member site id), not the reference's creation counters: the partition, the sizes and the label of the
 * spanning cluster (`perccln` of perc_site / perc_bond / perc_sitebond) are consistent within the files, which is all the
 * MATLAB scripts use.  path(pathlen) is a Fortran character variable (no terminating NUL needed). */
int32_t perc_write_txt(const int64_t *h, const int32_t *which, const char *path, const int32_t *pathlen);

/* ---- per-bond conductances ----------------------------------------------------------
 * The reference's MATLAB post-processor can give every bond of the spanning cluster its own conductance
 * (MATLAB/ConductCalc.m:38-47 `condtype = 2`; :94-96, :117-119, :139-141: G(i,j) = -g0*rand for a conducting bond, -1e-12
 * for every other lattice bond).  w(nb): conductance of bond row i of the reference's bond list (perc_geom_bondlist)
 * WHEN that bond conducts -- which bonds conduct is decided by the labels as before (bond / site / mixed rule), the
 * others keep gleak; g0 of the perc_conduct* calls is then ignored.  The array is copied; it stays in force for
 * every following perc_conduct / perc_conduct_g on the handle (voltages are always formed) until w = NULL
 * (Fortran: c_null_ptr) drops it.  Not available on slab handles and for warm starts.  The solve runs plain two-kernel
 * Jacobi-PCG on four weight planes + the diagonal (112 B per site and iteration on the square lattice, 144 B
 * triangular, against 50 B for the uniform-g0 kernels: csrc/pcg_weighted.cu). */
int32_t perc_set_bond_conductance(const int64_t *h, const double *w);
int32_t perc_clear_bond_conductance(const int64_t *h);                     /* = perc_set_bond_conductance(h, NULL) */

/* ---- solver selection --------------------------------------------------------------------------- */
/* Iteration kernels of the conductance solve (the body of linbcg's loop, Sq/bondc.f:780-833; Jacobi asolve :855-864):
 *   mode 0 (default): automatic -- perc_conduct_g on one GPU (pbc = 0, or pbc = 1 with m a multiple of 128) runs the
 *           DEFLATED ONE-PASS kernel:
 *           the Chronopoulos-Gear arrangement of the Jacobi-PCG recurrences (one reduction per iteration, 33 B per
 *           site and iteration) on top of a block-constant deflation space (Saad et al. 2000: x0 = Z E^-1 Z^T b,
 *           search directions A-orthogonal to the blocks).  x, err = |r| / |D^-1 b| and the stopping rule are
 *           linbcg's (r is the true residual of x); the iteration COUNT is several times smaller than linbcg's.
 *           Every other call runs the two-kernel form;
 *   mode 1: always the two-kernel form (50 B per site and iteration): plain Jacobi-PCG, linbcg's own sequence of
 *           operations -- iter is linbcg's count exactly;
 *   mode 2: the one-pass kernel without deflation (linbcg's iterates in exact arithmetic, iter within +-1);
 *   modes 10 / 12 / 14 (diagnostic): the one-pass kernel in one of its variants (pcg_fused_tile.cuh): FtCfgA
 *           (per-tile partial sums folded in tile order: the result does not depend on the number of SMs),
 *           FtCfgA3 (= mode 2), FtCfgD (= mode 0).
 * The process-wide default can be set with the environment variable PERC_PCG_SOLVER=classic|fused|deflated.
 * perc_solver_used reports which one the handle's last solve ran (0 = two-kernel form, 1 = one-pass kernel,
 * 2 = deflated one-pass kernel). */
int32_t perc_set_solver(const int64_t *h, const int32_t *mode);
int32_t perc_solver_used(const int64_t *h, int32_t *fused);

/* ---- instrumentation ---------------------------------------------------------------------------- */
/* kernels launched by this handle since creation (bench.py's gpu_launches) */
int32_t perc_launch_count(const int64_t *h, int64_t *count);
/* device time (ms, CUDA events on the handle's stream) of the last call's phases:
 * 0 mask build, 1 CCL local, 2 CCL merge, 3 CCL flatten+sizes, 4 spanning, 5 PCG total,
 * 6 PCG SpMV kernel avg, 7 PCG update kernel avg (one-pass solver: 6 = the iteration kernel, 7 = 0) */
int32_t perc_phase_ms(const int64_t *h, const int32_t *nphase, float *ms);
/* raw stream handle (cudaStream_t) so a host can order its own work against the library's */
int32_t perc_stream(const int64_t *h, uint64_t *stream);

#ifdef __cplusplus
}
#endif
#endif
