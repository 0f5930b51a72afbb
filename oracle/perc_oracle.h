/*
 * perc_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the percolation-realization hot path of
 * IsaiahSteinke/Percolation (Fortran 77).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product path (percolation_b200/, libperc_b200.so) never links or calls it.
 *
 * Pinning status: the reference ships no numeric test vectors.  The oracle is
 * pinned against (i) the reference's SampleOutput PNG renderings decoded into
 * tests/golden/ (occupancy counts, nesting, largest-cluster membership),
 * (ii) the closed forms the reference states (bond counts, p=1 conductance),
 * (iii) the libgfortran runtime RNG known-answer vectors.
 * The Fortran itself cannot be compiled in this image (no Fortran compiler),
 * so conductance VALUES and label tables are "parity unpinned" beyond those.
 *
 * All site / bond / label integers are 1-based exactly as in the Fortran
 * (C arrays hold 1-based values at 0-based positions).
 * File:line citations are into /root/reference/Fortran (Sq = Square, Tri = Triangular).
 */
#ifndef PERC_ORACLE_H
#define PERC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_SQUARE = 1, ORC_TRIANGULAR = 2 };           /* lattype, MATLAB/ConductCalc.m:30 */
enum { ORC_SITE = 1, ORC_BOND = 2, ORC_MIXED = 3 };    /* perctype, MATLAB/ConductCalc.m:29 */

typedef struct {
    int cln;      /* lowest unused cluster number                      Sq/site.f:165   */
    int maxcs;    /* largest overall cluster size                      Sq/site.f:278   */
    int maxcn;    /* its (historical) cluster number                                   */
    int perccln;  /* lowest historical label that spans, 0 if none     Sq/site.f:309   */
    int perccls;  /* its size (0 if none; App. B quirk)                                */
    int filled;   /* number of elements filled when the loop ended                     */
} orc_result;

/* ---- geometry ---------------------------------------------------------- */
int  orc_scn(int lattice);                              /* Sq/site.f:66, Tri/site.f:68 */
int  orc_bcn(int lattice);                              /* Sq/site.f:67, Tri/site.f:69 */
int  orc_nb(int lattice, int m, int n, int pbc);        /* Sq/site.f:89-93, Tri/site.f:91-95 */
void orc_nearestn(int lattice, int m, int n, int pbc, int rn, int nn[10]);
int  orc_bondlist(int lattice, int m, int n, int pbc, int *b1, int *b2);

/* ---- libgfortran srand/rand + shuffles ---------------------------------- */
void  orc_srand(int seed);
int   orc_irand(void);
float orc_rand(void);
void  orc_shuffle_sites(int seed, int t, int *order);
void  orc_shuffle_bonds(int seed, int nb, int *bo1, int *bo2);
void  orc_seed_table(int master, int count, int mult, int *out);
void  orc_sb_seed_tables(int master, int npseed, int which, int iter, int *pseed, int *sseed, int *bseed);
int   orc_fill_count(double p, int N);
int   orc_sweep_table(double p0, double dp, int npts, int nb, double *pbarr, int *nbarr);
float orc_fraction(int filled, int total);

/* ---- literal (history-dependent) labelers -------------------------------- */
int orc_site_literal(int lattice, int m, int n, int pbc, const int *order, int k,
                     int stop_at_span, int *s, int *c, orc_result *res);
int orc_bond_literal(int lattice, int m, int n, int pbc, int nb, const int *b1, const int *b2,
                     const int *bo1, const int *bo2, int k, int stop_at_span,
                     int *b3, int *c, orc_result *res);
int orc_sitebond_literal(int lattice, int m, int n, int pbc, int nb, const int *b1, const int *b2,
                         const int *sorder, int ks, const int *bo1, const int *bo2, int kb,
                         int stop_at_span, int *s, int *b3, int *c, orc_result *res);
int orc_bondsite_literal(int lattice, int m, int n, int pbc, int nb, const int *b1, const int *b2,
                         const int *bo1, const int *bo2, int kb, const int *sorder, int ks,
                         int stop_at_span, int *s, int *b3, int *c, orc_result *res);

/* ---- canonical labels ----------------------------------------------------- */
/* efficient union-find labeler on occupancy flags; canonical min-id labels.
 * site_occ[t] / bond_occ[nb] are 0/1 bytes (NULL where the kind has none).
 * s_can[t], b3_can[nb] canonical labels, csize[t+1] size by canonical label
 * (labels <= t); returns number of clusters (mixed: incl. lone bonds, size 1,
 * label t+row) via *ncl. */
int orc_label_uf(int kind, int lattice, int m, int n, int pbc, int nb, const int *b1, const int *b2,
                 const uint8_t *site_occ, const uint8_t *bond_occ,
                 int *s_can, int *b3_can, int *csize, int64_t *ncl, int *maxcs);
/* map literal (reference) labels to canonical ones; returns 0 if the mapping
 * is a bijection between clusters and sizes agree, else a negative code. */
int orc_canonicalise(int kind, int t, int nb, const int *b1, const int *b2,
                     const int *s_ref, const int *b3_ref, const int *c_ref, int cln,
                     int *s_can, int *b3_can, int *csize);
/* spanning clusters (canonical ids ascending) touching row 0 and row n-1 */
int orc_spanning(int kind, int m, int n, int nb, const int *b1, const int *b2,
                 const int *s_can, const int *b3_can, int *ids, int max_ids);
/* exact size histogram: hist[sz] for sz in 1..maxsize (hist has maxsize+1 entries) */
void orc_size_hist(int kind, int t, int nb, const int *csize, const int *b3_can, int64_t *hist, int maxsize);

/* ---- conductance ----------------------------------------------------------- */
void orc_weights(int kind, int nb, const int *b1, const int *b2, const int *s, const int *b3,
                 int perccln, double g0, double gleak, double *w);
int orc_conduct_literal(int m, int n, int nb, const int *b1, const int *b2, const double *w,
                        double Va, double tol, int itmax, double read_thresh,
                        double *Vint, double *Gtop, double *Gbot, int *iter, double *err);
int orc_conduct_cg(int m, int n, int nb, const int *b1, const int *b2, const double *w,
                   double Va, double tol, int itmax, double read_thresh,
                   double *Vint, double *Gtop, double *Gbot, int *iter, double *err);
/* independent check: ||b - A*Vint||_2 / ||D^-1 b||_2 (the linbcg itol=2 measure) and read-out for given Vint */
int orc_conduct_check(int m, int n, int nb, const int *b1, const int *b2, const double *w,
                      double Va, double read_thresh, const double *Vint,
                      double *Gtop, double *Gbot, double *err);
/* time `iters` CG iterations (no convergence test) -- CPU baseline sampling */
double orc_cg_time_iters(int m, int n, int nb, const int *b1, const int *b2, const double *w,
                         double Va, int iters);

#ifdef __cplusplus
}
#endif
#endif
