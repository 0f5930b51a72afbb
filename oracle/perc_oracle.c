/*
 * perc_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE).
 * See perc_oracle.h for the scope / pinning statement.
 *
 * Restates, in plain C99, the algorithms of /root/reference/Fortran (Sq = Square,
 * Tri = Triangular).  Compile with -O2 -ffp-contract=off (float32 shuffle
 * arithmetic must not be contracted or widened).
 */
#include "perc_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>

#define ORC_OMP_MIN 2000000      /* systems above this size use OpenMP in the plain CG (golden-fixture runs) */

/* ======================================================================== */
/* geometry                                                                  */
/* ======================================================================== */

int orc_scn(int lattice) { return lattice == ORC_SQUARE ? 4 : 6; }
int orc_bcn(int lattice) { return lattice == ORC_SQUARE ? 6 : 10; }

int orc_nb(int lattice, int m, int n, int pbc)
{
    if (lattice == ORC_SQUARE)                       /* Sq/site.f:89-93 */
        return pbc == 1 ? m * (2 * n - 1) : 2 * m * n - m - n;
    return pbc == 1 ? m * (3 * n - 2) : 3 * m * n - 2 * m - 2 * n + 1;   /* Tri/site.f:91-95 */
}

/* Sq/site.f:371-469 */
static void nearestn_sq(int m, int n, int pbc, int rn, int *nn)
{
    int t = m * n;
    (void)n;
    if (rn == 1)            { nn[0] = rn + 1; nn[1] = rn + m; if (pbc == 1) nn[2] = m; return; }
    if (rn == m)            { nn[0] = rn - 1; nn[1] = rn + m; if (pbc == 1) nn[2] = 1; return; }
    if (rn == t - (m - 1))  { nn[0] = rn - m; nn[1] = rn + 1; if (pbc == 1) nn[2] = t; return; }
    if (rn == t)            { nn[0] = rn - m; nn[1] = rn - 1; if (pbc == 1) nn[2] = rn - (m - 1); return; }
    if (rn < m)             { nn[0] = rn - 1; nn[1] = rn + 1; nn[2] = rn + m; return; }
    if (rn > t - m)         { nn[0] = rn - m; nn[1] = rn - 1; nn[2] = rn + 1; return; }
    if ((rn - 1) % m == 0)  { nn[0] = rn - m; nn[1] = rn + 1; nn[2] = rn + m; if (pbc == 1) nn[3] = rn + (m - 1); return; }
    if (rn % m == 0)        { nn[0] = rn - m; nn[1] = rn - 1; nn[2] = rn + m; if (pbc == 1) nn[3] = rn - (m - 1); return; }
    nn[0] = rn - m; nn[1] = rn - 1; nn[2] = rn + 1; nn[3] = rn + m;
}

/* Tri/site.f:373-558 (including the odd-m branches, literally) */
static void nearestn_tri(int m, int n, int pbc, int rn, int *nn)
{
    int t = m * n;
    int up;   /* 1: neighbours rn-m, rn-1, rn+1, rn+m-1, rn+m, rn+m+1 */
    (void)n;
    if (rn == 1) {
        nn[0] = rn + 1; nn[1] = rn + m; nn[2] = rn + (m + 1);
        if (pbc == 1) { nn[3] = rn + (m - 1); nn[4] = rn + (2 * m - 1); }
        return;
    }
    if (rn == m) {
        nn[0] = rn - 1; nn[1] = rn + m;
        if (m % 2 == 1) { nn[2] = rn + (m - 1); return; }
        if (pbc == 1) nn[2] = 1;
        return;
    }
    if (rn == t - (m - 1)) {
        nn[0] = rn - m; nn[1] = rn + 1;
        if (pbc == 1) nn[2] = t;
        return;
    }
    if (rn == t) {
        if (m % 2 == 1) { nn[0] = rn - m; nn[1] = rn - 1; return; }
        nn[0] = rn - (m + 1); nn[1] = rn - m; nn[2] = rn - 1;
        if (pbc == 1) { nn[3] = rn - (2 * m - 1); nn[4] = rn - (m - 1); }
        return;
    }
    if (rn < m) {
        if (rn % 2 == 0) { nn[0] = rn - 1; nn[1] = rn + 1; nn[2] = rn + m; }
        else { nn[0] = rn - 1; nn[1] = rn + 1; nn[2] = rn + (m - 1); nn[3] = rn + m; nn[4] = rn + (m + 1); }
        return;
    }
    if (rn > t - m) {
        if (rn % 2 == 0) { nn[0] = rn - (m + 1); nn[1] = rn - m; nn[2] = rn - (m - 1); nn[3] = rn - 1; nn[4] = rn + 1; }
        else { nn[0] = rn - m; nn[1] = rn - 1; nn[2] = rn + 1; }
        return;
    }
    if ((rn - 1) % m == 0) {
        nn[0] = rn - m; nn[1] = rn + 1; nn[2] = rn + m; nn[3] = rn + (m + 1);
        if (pbc == 1) { nn[4] = rn + (m - 1); nn[5] = rn + (2 * m - 1); }
        return;
    }
    if (rn % m == 0) {
        if (m % 2 == 1) { nn[0] = rn - m; nn[1] = rn - 1; nn[2] = rn + (m - 1); nn[3] = rn + m; return; }
        nn[0] = rn - (m + 1); nn[1] = rn - m; nn[2] = rn - 1; nn[3] = rn + m;
        if (pbc == 1) { nn[4] = rn - (2 * m - 1); nn[5] = rn - (m - 1); }
        return;
    }
    if (m % 2 == 1) {
        if ((rn / m) % 2 == 0) up = (rn % 2 != 0);
        else                   up = (rn % 2 == 0);
    } else {
        up = (rn % 2 != 0);
    }
    if (up) { nn[0] = rn - m; nn[1] = rn - 1; nn[2] = rn + 1; nn[3] = rn + (m - 1); nn[4] = rn + m; nn[5] = rn + (m + 1); }
    else    { nn[0] = rn - (m + 1); nn[1] = rn - m; nn[2] = rn - (m - 1); nn[3] = rn - 1; nn[4] = rn + 1; nn[5] = rn + m; }
}

void orc_nearestn(int lattice, int m, int n, int pbc, int rn, int nn[10])
{
    int z;
    for (z = 0; z < 10; ++z) nn[z] = 0;
    if (lattice == ORC_SQUARE) nearestn_sq(m, n, pbc, rn, nn);
    else nearestn_tri(m, n, pbc, rn, nn);
}

/* Sq/bond.f:112-129 (== Sq/site.f:106-120): rows (i, nn(j)) for nn(j) > i, in nn order */
int orc_bondlist(int lattice, int m, int n, int pbc, int *b1, int *b2)
{
    int t = m * n, scn = orc_scn(lattice), rc = 0, i, j, nn[10];
    for (i = 1; i <= t - 1; ++i) {
        orc_nearestn(lattice, m, n, pbc, i, nn);
        for (j = 0; j < scn; ++j)
            if (nn[j] > i) {
                if (b1) { b1[rc] = i; b2[rc] = nn[j]; }
                rc++;
            }
    }
    return rc;
}

/* ======================================================================== */
/* libgfortran srand / irand / rand  (GCC libgfortran/intrinsics/rand.c):     */
/* Park-Miller x <- 16807 x mod (2^31-1); srand(0) -> 123459876;              */
/* rand = float32((x-1) & 0xFFFFFE00) / float32(2^31-2).                      */
/* Call sites: Fortran/permute.f:25,40; Sq/site.f:132,143; Sq/site_perc.f:71-74 */
/* ======================================================================== */

static uint64_t g_rand_seed = 1;

void orc_srand(int seed) { g_rand_seed = seed ? (uint64_t)(uint32_t)seed : 123459876u; }

int orc_irand(void)
{
    g_rand_seed = (16807u * g_rand_seed) % 2147483647u;
    return (int)g_rand_seed;
}

float orc_rand(void)
{
    uint32_t mask = ~(uint32_t)0u << (32 - 24 + 1);
    uint32_t v = ((uint32_t)orc_irand() - 1u) & mask;
    return (float)v / (float)2147483646;
}

/* Sq/site.f:131-147: j = i + (t-i+1)*rand(0), mixed INTEGER/REAL*4 -> REAL*4, truncated */
static int shuffle_target(int i, int N)
{
    volatile float r = orc_rand();
    volatile float prod = (float)(N - i + 1) * r;
    volatile float sum = (float)i + prod;
    return (int)sum;
}

void orc_shuffle_sites(int seed, int t, int *order)
{
    int i, j, tmp;
    orc_srand(seed);
    for (i = 1; i <= t; ++i) order[i - 1] = i;
    for (i = 1; i <= t; ++i) {
        j = shuffle_target(i, t);
        tmp = order[i - 1]; order[i - 1] = order[j - 1]; order[j - 1] = tmp;
    }
}

/* Sq/bond.f:137-150: both columns swap; bo1/bo2 arrive holding the bond list */
void orc_shuffle_bonds(int seed, int nb, int *bo1, int *bo2)
{
    int i, j, t1, t2;
    orc_srand(seed);
    for (i = 1; i <= nb; ++i) {
        j = shuffle_target(i, nb);
        t1 = bo1[i - 1]; t2 = bo2[i - 1];
        bo1[i - 1] = bo1[j - 1]; bo2[i - 1] = bo2[j - 1];
        bo1[j - 1] = t1; bo2[j - 1] = t2;
    }
}

/* Sq/site_perc.f:70-75 (mult 1000000), Sq/bond_cond.f:65-70 (mult 10000000): REAL*4 multiply */
static int seed_draw(int mult)
{
    volatile float r = orc_rand();
    volatile float p = r * (float)mult;
    return (int)p + 1;
}

void orc_seed_table(int master, int count, int mult, int *out)
{
    int i;
    orc_srand(master);
    for (i = 0; i < count; ++i) out[i] = seed_draw(mult);
}

/* Sq/sb_perc.f:94-97,108-112: pseed(1..100) from master; for ps point `which`
 * (1-based) srand(pseed(which)) then sseed/bseed drawn interleaved. */
void orc_sb_seed_tables(int master, int npseed, int which, int iter, int *pseed, int *sseed, int *bseed)
{
    int i;
    orc_srand(master);
    for (i = 0; i < npseed; ++i) pseed[i] = seed_draw(10000000);
    if (which >= 1 && which <= npseed && sseed && bseed) {
        orc_srand(pseed[which - 1]);
        for (i = 0; i < iter; ++i) { sseed[i] = seed_draw(10000000); bseed[i] = seed_draw(10000000); }
    }
}

/* Sq/site.f:164 tsites = ps*t: DOUBLE PRECISION product truncated on assignment */
int orc_fill_count(double p, int N)
{
    volatile double v = p * (double)N;
    return (int)v;
}

/* Sq/bond_cond.f:89-94: pbarr(1)=p0; pbarr(i)=pbarr(i-1)+dp; nbarr(i)=pbarr(i)*nb */
int orc_sweep_table(double p0, double dp, int npts, int nb, double *pbarr, int *nbarr)
{
    int i;
    volatile double acc = p0;
    for (i = 0; i < npts; ++i) {
        if (i > 0) acc = acc + dp;
        pbarr[i] = acc;
        nbarr[i] = orc_fill_count(acc, nb);
    }
    return npts;
}

/* f = real(sf)/real(t): REAL*4 division (Sq/site_perc.f:221) */
float orc_fraction(int filled, int total)
{
    volatile float a = (float)filled, b = (float)total;
    return a / b;
}

/* ======================================================================== */
/* literal labelers                                                           */
/* ======================================================================== */

/* spanning scan through site labels, Sq/site.f:309-344 */
static void span_sites_literal(int m, int n, const int *s, const int *c, int cln, int thr, orc_result *res)
{
    int t = m * n, i, j, bot, top;
    res->perccln = 0; res->perccls = 0;
    for (i = 1; i <= cln - 1; ++i) {
        if (c[i] < thr) continue;
        bot = top = 0;
        for (j = 1; j <= m; ++j) if (s[j - 1] == i) { bot = 1; break; }
        if (!bot) continue;
        for (j = t - m + 1; j <= t; ++j) if (s[j - 1] == i) { top = 1; break; }
        if (top + bot == 2) { res->perccln = i; res->perccls = c[i]; return; }
    }
}

/* Sq/site.f:162-289 (stop_at_span: Sq/site_perc.f:133-254).  c has t+2 entries, c[0] == 0. */
int orc_site_literal(int lattice, int m, int n, int pbc, const int *order, int k,
                     int stop_at_span, int *s, int *c, orc_result *res)
{
    int t = m * n, scn = orc_scn(lattice);
    int cln = 1, maxcs = 0, maxcn = 0, i, j, kk, l, nn[10];
    int lcn, lcs, nnlc, clsum, oldcn = 0, sn, dup;
    memset(s, 0, sizeof(int) * (size_t)t);
    memset(c, 0, sizeof(int) * (size_t)(t + 2));
    memset(res, 0, sizeof(*res));
#define S_(x) ((x) == 0 ? 0 : s[(x) - 1])
    for (i = 1; i <= k; ++i) {
        sn = order[i - 1];
        orc_nearestn(lattice, m, n, pbc, sn, nn);
        lcn = S_(nn[0]); lcs = c[lcn]; nnlc = nn[0];
        for (kk = 1; kk < scn; ++kk)
            if (nn[kk] != 0 && S_(nn[kk]) != 0 && c[S_(nn[kk])] > lcs) {
                lcn = S_(nn[kk]); lcs = c[lcn]; nnlc = nn[kk];
            }
        if (lcs == 0) {                                   /* case 1, :211-220 */
            s[sn - 1] = cln; c[cln] = 1; cln++;
        } else {                                          /* case 2, :224-260 */
            clsum = lcs;
            for (kk = 0; kk < scn; ++kk) {
                if (nn[kk] == 0 || S_(nn[kk]) == 0) continue;
                if (S_(nn[kk]) == S_(nnlc)) continue;
                dup = 0;
                for (l = 0; l < kk; ++l) if (S_(nn[l]) == S_(nn[kk])) { dup = 1; break; }
                if (!dup) {
                    clsum += c[S_(nn[kk])];
                    oldcn = S_(nn[kk]);
                    for (j = 0; j < t; ++j) if (s[j] == oldcn) s[j] = lcn;     /* :245-249 */
                }
                c[oldcn] = 0;                             /* label 260 */
            }
            s[sn - 1] = lcn; clsum++; c[lcn] = clsum;
        }
        if (c[lcn] > maxcs) { maxcs = c[lcn]; maxcn = lcn; }                    /* :278-287 */
        else if (lcs == 0 && maxcs == 0) { maxcs = 1; maxcn = 1; }
        res->filled = i;
        if (stop_at_span && i >= n) {                     /* Sq/site_perc.f:227-252 */
            span_sites_literal(m, n, s, c, cln, n, res);
            if (res->perccln) break;
        }
    }
#undef S_
    res->cln = cln; res->maxcs = maxcs; res->maxcn = maxcn;
    if (!stop_at_span) span_sites_literal(m, n, s, c, cln, n, res);
    return 0;
}

/* row lookup of bond (a,b), a<b: rows are sorted by b1, so search a's block.
 * (The reference scans all nb rows, Sq/bond.f:226-235; the result is the same.) */
typedef struct { int *start; int t; } rowindex;

static rowindex rowindex_build(int t, int nb, const int *b1)
{
    rowindex ri; int r, i;
    ri.t = t;
    ri.start = (int *)calloc((size_t)t + 2, sizeof(int));
    for (r = 0; r < nb; ++r) ri.start[b1[r] + 1]++;
    for (i = 1; i <= t + 1; ++i) ri.start[i] += ri.start[i - 1];
    /* start[a] .. start[a+1]-1 are the 0-based rows with b1 == a */
    return ri;
}

static int row_of(const rowindex *ri, const int *b2, int a, int b)
{
    int r;
    for (r = ri->start[a]; r < ri->start[a + 1]; ++r) if (b2[r] == b) return r;
    return -1;
}

/* Sq/bond.f:389-432: bond with b1 in bottom row and one with b2 in top row */
static void span_bonds_literal(int m, int n, int nb, const int *b1, const int *b2, const int *b3,
                               const int *c, int cln, orc_result *res)
{
    int t = m * n, i, r;
    unsigned char *bot = (unsigned char *)calloc((size_t)cln + 1, 1);
    unsigned char *top = (unsigned char *)calloc((size_t)cln + 1, 1);
    res->perccln = 0; res->perccls = 0;
    for (r = 0; r < nb; ++r) {
        if (b3[r] == 0) continue;
        if (b1[r] <= m) bot[b3[r]] = 1;
        if (b2[r] >= t - m + 1) top[b3[r]] = 1;
    }
    for (i = 1; i <= cln - 1; ++i)
        if (c[i] >= n - 1 && bot[i] && top[i]) { res->perccln = i; res->perccls = c[i]; break; }
    free(bot); free(top);
}

/* Sq/bond.f:165-369 (stop_at_span: Sq/bond_perc.f:326-359).  c has nb+2 entries. */
int orc_bond_literal(int lattice, int m, int n, int pbc, int nb, const int *b1, const int *b2,
                     const int *bo1, const int *bo2, int k, int stop_at_span,
                     int *b3, int *c, orc_result *res)
{
    int t = m * n, scn = orc_scn(lattice), bcn = orc_bcn(lattice);
    int cln = 1, maxcs = 0, maxcn = 0, i, j, kk, l, rc, nn[10], row, dup;
    int nnb[25][4], lcn, lcs, clsum, e, a, other;
    rowindex ri = rowindex_build(t, nb, b1);
    memset(b3, 0, sizeof(int) * (size_t)nb);
    memset(c, 0, sizeof(int) * (size_t)(nb + 2));
    memset(res, 0, sizeof(*res));
    for (i = 1; i <= k; ++i) {
        memset(nnb, 0, sizeof(nnb));
        rc = 0;
        for (e = 0; e < 2; ++e) {                         /* :192-224 */
            a = e == 0 ? bo1[i - 1] : bo2[i - 1];
            other = e == 0 ? bo2[i - 1] : bo1[i - 1];
            orc_nearestn(lattice, m, n, pbc, a, nn);
            for (j = 0; j < scn; ++j) {
                if (nn[j] == 0 || nn[j] == other) continue;
                if (nn[j] > a) { nnb[rc][0] = a; nnb[rc][1] = nn[j]; }
                else           { nnb[rc][0] = nn[j]; nnb[rc][1] = a; }
                rc++;
            }
        }
        for (kk = 0; kk < bcn; ++kk) {                    /* :226-235 snapshot */
            if (nnb[kk][0] == 0) continue;
            row = row_of(&ri, b2, nnb[kk][0], nnb[kk][1]);
            if (row >= 0) { nnb[kk][2] = b3[row]; nnb[kk][3] = c[b3[row]]; }
        }
        lcn = nnb[0][2]; lcs = nnb[0][3];                 /* :251-265 */
        for (kk = 1; kk < bcn; ++kk)
            if (nnb[kk][0] != 0 && nnb[kk][2] != 0 && nnb[kk][3] > lcs) { lcn = nnb[kk][2]; lcs = nnb[kk][3]; }
        row = row_of(&ri, b2, bo1[i - 1], bo2[i - 1]);
        if (lcs == 0) {                                   /* case 1, :280-295 */
            b3[row] = cln; c[cln] = 1; cln++;
        } else {                                          /* case 2, :299-340 */
            clsum = lcs;
            for (kk = 0; kk < bcn; ++kk) {
                if (nnb[kk][0] == 0 || nnb[kk][2] == 0 || nnb[kk][2] == lcn) continue;
                dup = 0;
                for (l = 0; l < kk; ++l) if (nnb[l][2] == nnb[kk][2]) { dup = 1; break; }
                if (!dup) {
                    clsum += nnb[kk][3];
                    for (j = 0; j < nb; ++j) if (b3[j] == nnb[kk][2]) b3[j] = lcn;    /* :319-323 */
                }
                c[nnb[kk][2]] = 0;                        /* label 330 */
            }
            b3[row] = lcn; clsum++; c[lcn] = clsum;
        }
        if (c[lcn] > maxcs) { maxcs = c[lcn]; maxcn = lcn; }
        else if (lcs == 0 && maxcs == 0) { maxcs = 1; maxcn = 1; }
        res->filled = i;
        if (stop_at_span && i >= n - 1) {
            span_bonds_literal(m, n, nb, b1, b2, b3, c, cln, res);
            if (res->perccln) break;
        }
    }
    res->cln = cln; res->maxcs = maxcs; res->maxcn = maxcn;
    if (!stop_at_span) span_bonds_literal(m, n, nb, b1, b2, b3, c, cln, res);
    free(ri.start);
    return 0;
}

/* Sq/sitebond.f:187-400 (stop_at_span: Sq/sb_perc.f:232-357).  c has t+nb+2 entries. */
int orc_sitebond_literal(int lattice, int m, int n, int pbc, int nb, const int *b1, const int *b2,
                         const int *sorder, int ks, const int *bo1, const int *bo2, int kb,
                         int stop_at_span, int *s, int *b3, int *c, orc_result *res)
{
    int t = m * n, cln = 1, maxcs = 1, maxcn = 1, i, kk, row, lcn = 0, lcs = 0, oldcn, clsum, sa, sb;
    rowindex ri = rowindex_build(t, nb, b1);
    (void)lattice; (void)pbc;
    memset(s, 0, sizeof(int) * (size_t)t);
    memset(b3, 0, sizeof(int) * (size_t)nb);
    memset(c, 0, sizeof(int) * (size_t)(t + nb + 2));
    memset(res, 0, sizeof(*res));
    for (i = 1; i <= ks; ++i) { s[sorder[i - 1] - 1] = cln; c[cln] = 1; cln++; }      /* :187-196 */
    for (i = 1; i <= kb; ++i) {
        row = row_of(&ri, b2, bo1[i - 1], bo2[i - 1]);
        sa = s[b1[row] - 1]; sb = s[b2[row] - 1];
        if (sa == 0 && sb == 0) {                         /* case 1, :231-242 */
            b3[row] = cln; c[cln] = 1; cln++;
        } else if (sa > 0 && sb == 0) {                   /* case 2 */
            lcn = sa; lcs = c[sa]; b3[row] = lcn; c[sa] = lcs + 1;
        } else if (sa == 0 && sb > 0) {
            lcn = sb; lcs = c[sb]; b3[row] = lcn; c[sb] = lcs + 1;
        } else if (sa == sb) {                            /* case 3 same cluster, :293-305 */
            lcn = sa; lcs = c[sa]; b3[row] = lcn; c[sa] = lcs + 1;
        } else {
            if (c[sa] > c[sb]) { lcn = sa; oldcn = sb; }  /* :307 strict > keeps the b1 side */
            else               { lcn = sb; oldcn = sa; }
            lcs = c[lcn]; b3[row] = lcn; clsum = lcs + c[oldcn] + 1;
            for (kk = 0; kk < t; ++kk) if (s[kk] == oldcn) s[kk] = lcn;
            for (kk = 0; kk < nb; ++kk) if (b3[kk] == oldcn) b3[kk] = lcn;
            c[oldcn] = 0; c[lcn] = clsum;
        }
        if (c[lcn] > maxcs) { maxcs = c[lcn]; maxcn = lcn; }                            /* 310 */
        res->filled = i;
        if (stop_at_span && i >= n - 1) {
            span_sites_literal(m, n, s, c, cln, 2 * n - 1, res);
            if (res->perccln) break;
        }
    }
    res->cln = cln; res->maxcs = maxcs; res->maxcn = maxcn;
    if (!stop_at_span) span_sites_literal(m, n, s, c, cln, 2 * n - 1, res);
    free(ri.start);
    return 0;
}

/* Sq/bondsite.f:182-354 (stop_at_span: Sq/bs_perc.f:237-380).  c has t+nb+2 entries. */
int orc_bondsite_literal(int lattice, int m, int n, int pbc, int nb, const int *b1, const int *b2,
                         const int *bo1, const int *bo2, int kb, const int *sorder, int ks,
                         int stop_at_span, int *s, int *b3, int *c, orc_result *res)
{
    int t = m * n, scn = orc_scn(lattice), cln = 1, maxcs = 1, maxcn = 1;
    int i, j, kk, l, rc, row, nn[10], nnb[12][4], lcn = 0, lcs = 0, clsum, sn, dup;
    rowindex ri = rowindex_build(t, nb, b1);
    memset(s, 0, sizeof(int) * (size_t)t);
    memset(b3, 0, sizeof(int) * (size_t)nb);
    memset(c, 0, sizeof(int) * (size_t)(t + nb + 2));
    memset(res, 0, sizeof(*res));
    for (i = 1; i <= kb; ++i) {                           /* :182-197 */
        row = row_of(&ri, b2, bo1[i - 1], bo2[i - 1]);
        b3[row] = cln; c[cln] = 1; cln++;
    }
    for (i = 1; i <= ks; ++i) {
        sn = sorder[i - 1];
        memset(nnb, 0, sizeof(nnb));
        rc = 0;
        orc_nearestn(lattice, m, n, pbc, sn, nn);         /* :232-246 */
        for (j = 0; j < scn; ++j) {
            if (nn[j] == 0) continue;
            if (nn[j] > sn) { nnb[rc][0] = sn; nnb[rc][1] = nn[j]; }
            else            { nnb[rc][0] = nn[j]; nnb[rc][1] = sn; }
            rc++;
        }
        for (kk = 0; kk < scn; ++kk) {
            if (nnb[kk][0] == 0) continue;
            row = row_of(&ri, b2, nnb[kk][0], nnb[kk][1]);
            if (row >= 0) { nnb[kk][2] = b3[row]; nnb[kk][3] = c[b3[row]]; }
        }
        lcn = nnb[0][2]; lcs = nnb[0][3];
        for (kk = 1; kk < scn; ++kk)
            if (nnb[kk][0] != 0 && nnb[kk][2] != 0 && nnb[kk][3] > lcs) { lcn = nnb[kk][2]; lcs = nnb[kk][3]; }
        if (lcs == 0) {                                   /* case 1 */
            s[sn - 1] = cln; c[cln] = 1; cln++;
        } else {
            clsum = lcs;
            for (kk = 0; kk < scn; ++kk) {
                if (nnb[kk][0] == 0 || nnb[kk][2] == 0 || nnb[kk][2] == lcn) continue;
                dup = 0;
                for (l = 0; l < kk; ++l) if (nnb[l][2] == nnb[kk][2]) { dup = 1; break; }
                if (!dup) {
                    clsum += nnb[kk][3];
                    for (j = 0; j < nb; ++j) if (b3[j] == nnb[kk][2]) b3[j] = lcn;
                    for (j = 0; j < t; ++j) if (s[j] == nnb[kk][2]) s[j] = lcn;
                }
                c[nnb[kk][2]] = 0;
            }
            s[sn - 1] = lcn; clsum++; c[lcn] = clsum;
        }
        if (c[lcn] > maxcs) { maxcs = c[lcn]; maxcn = lcn; }
        res->filled = i;
        if (stop_at_span && i >= n) {
            span_sites_literal(m, n, s, c, cln, 2 * n - 1, res);
            if (res->perccln) break;
        }
    }
    res->cln = cln; res->maxcs = maxcs; res->maxcn = maxcn;
    if (!stop_at_span) span_sites_literal(m, n, s, c, cln, 2 * n - 1, res);
    free(ri.start);
    return 0;
}

/* ======================================================================== */
/* canonical (order-independent) labels                                       */
/* ======================================================================== */

static int uf_find(int *p, int x)
{
    int r = x, nx;
    while (p[r] != r) r = p[r];
    while (p[x] != r) { nx = p[x]; p[x] = r; x = nx; }
    return r;
}

static void uf_union_min(int *p, int a, int b)
{
    a = uf_find(p, a); b = uf_find(p, b);
    if (a == b) return;
    if (a < b) p[b] = a; else p[a] = b;
}

/* Order-independent result of the four fills:
 * site  -> components of occupied sites, size = #sites, label = min site id;
 * bond  -> components of occupied bonds linked through shared sites, size =
 *          #bonds, label = min end-point site id;
 * mixed -> components of {occupied sites} u {occupied bonds}, bond-site edge
 *          iff that end site is occupied, size = #sites + #bonds, label = min
 *          element id (sites rn, bonds t+row). */
int orc_label_uf(int kind, int lattice, int m, int n, int pbc, int nb, const int *b1, const int *b2,
                 const uint8_t *site_occ, const uint8_t *bond_occ,
                 int *s_can, int *b3_can, int *csize, int64_t *ncl, int *maxcs)
{
    int t = m * n, i, r, a, b, root;
    int *p = (int *)malloc(sizeof(int) * (size_t)(t + 1));
    int64_t ncl_ = 0; int mx = 0;
    (void)lattice; (void)pbc;
    for (i = 0; i <= t; ++i) p[i] = i;
    memset(csize, 0, sizeof(int) * (size_t)(t + 1));
    if (s_can) memset(s_can, 0, sizeof(int) * (size_t)t);
    if (b3_can) memset(b3_can, 0, sizeof(int) * (size_t)nb);
    for (r = 0; r < nb; ++r) {
        a = b1[r]; b = b2[r];
        if (kind == ORC_SITE)  { if (site_occ[a - 1] && site_occ[b - 1]) uf_union_min(p, a, b); }
        else if (kind == ORC_BOND) { if (bond_occ[r]) uf_union_min(p, a, b); }
        else { if (bond_occ[r] && site_occ[a - 1] && site_occ[b - 1]) uf_union_min(p, a, b); }
    }
    if (kind == ORC_SITE || kind == ORC_MIXED)
        for (i = 1; i <= t; ++i)
            if (site_occ[i - 1]) { root = uf_find(p, i); s_can[i - 1] = root; csize[root]++; }
    if (kind == ORC_BOND)
        for (r = 0; r < nb; ++r)
            if (bond_occ[r]) { root = uf_find(p, b1[r]); b3_can[r] = root; csize[root]++; }
    if (kind == ORC_MIXED)
        for (r = 0; r < nb; ++r) {
            if (!bond_occ[r]) continue;
            a = b1[r]; b = b2[r];
            if (site_occ[a - 1])      { root = uf_find(p, a); b3_can[r] = root; csize[root]++; }
            else if (site_occ[b - 1]) { root = uf_find(p, b); b3_can[r] = root; csize[root]++; }
            else { b3_can[r] = t + r + 1; ncl_++; if (mx < 1) mx = 1; }
        }
    for (i = 1; i <= t; ++i) if (csize[i] > 0) { ncl_++; if (csize[i] > mx) mx = csize[i]; }
    if (ncl) *ncl = ncl_;
    if (maxcs) *maxcs = mx;
    free(p);
    return 0;
}

/* reference label -> canonical label (smallest member element id).  c_ref is
 * indexed by reference label (c_ref[0] unused), cln = lowest unused label. */
int orc_canonicalise(int kind, int t, int nb, const int *b1, const int *b2,
                     const int *s_ref, const int *b3_ref, const int *c_ref, int cln,
                     int *s_can, int *b3_can, int *csize)
{
    int i, r, L, rc = 0;
    int *canon = (int *)malloc(sizeof(int) * (size_t)(cln + 1));
    int *cnt = (int *)calloc((size_t)cln + 1, sizeof(int));
    (void)b2;
    for (i = 0; i <= cln; ++i) canon[i] = 0x7fffffff;
    memset(csize, 0, sizeof(int) * (size_t)(t + 1));
    if (kind == ORC_SITE || kind == ORC_MIXED)
        for (i = 1; i <= t; ++i) { L = s_ref[i - 1]; if (L) { if (i < canon[L]) canon[L] = i; cnt[L]++; } }
    if (kind == ORC_BOND)
        for (r = 0; r < nb; ++r) { L = b3_ref[r]; if (L) { if (b1[r] < canon[L]) canon[L] = b1[r]; cnt[L]++; } }
    if (kind == ORC_MIXED)
        for (r = 0; r < nb; ++r) { L = b3_ref[r]; if (L) { if (t + r + 1 < canon[L]) canon[L] = t + r + 1; cnt[L]++; } }
    /* sizes the reference keeps in c() must equal the member counts */
    for (L = 1; L < cln; ++L) {
        if (cnt[L] != c_ref[L]) { rc = -1; break; }
        if (cnt[L] && canon[L] <= t) {
            if (csize[canon[L]] != 0) { rc = -2; break; }     /* two ref labels -> one canonical */
            csize[canon[L]] = cnt[L];
        }
    }
    if (s_can && s_ref) for (i = 0; i < t; ++i) s_can[i] = s_ref[i] ? canon[s_ref[i]] : 0;
    if (b3_can && b3_ref) for (r = 0; r < nb; ++r) b3_can[r] = b3_ref[r] ? canon[b3_ref[r]] : 0;
    free(canon); free(cnt);
    return rc;
}

static int cmp_int(const void *a, const void *b)
{
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

/* site/mixed via s on rows 0 and n-1; bond via b1 in bottom row and b2 in top row */
int orc_spanning(int kind, int m, int n, int nb, const int *b1, const int *b2,
                 const int *s_can, const int *b3_can, int *ids, int max_ids)
{
    int t = m * n, i, r, cnt = 0, nbot = 0, ntop = 0, j;
    int *bot = (int *)malloc(sizeof(int) * (size_t)(kind == ORC_BOND ? nb : m) + 4);
    int *top = (int *)malloc(sizeof(int) * (size_t)(kind == ORC_BOND ? nb : m) + 4);
    if (kind == ORC_BOND) {
        for (r = 0; r < nb; ++r) {
            if (!b3_can[r]) continue;
            if (b1[r] <= m) bot[nbot++] = b3_can[r];
            if (b2[r] >= t - m + 1) top[ntop++] = b3_can[r];
        }
    } else {
        for (i = 1; i <= m; ++i) if (s_can[i - 1]) bot[nbot++] = s_can[i - 1];
        for (i = t - m + 1; i <= t; ++i) if (s_can[i - 1]) top[ntop++] = s_can[i - 1];
    }
    qsort(bot, (size_t)nbot, sizeof(int), cmp_int);
    qsort(top, (size_t)ntop, sizeof(int), cmp_int);
    i = j = 0;
    while (i < nbot && j < ntop) {
        if (bot[i] < top[j]) i++;
        else if (bot[i] > top[j]) j++;
        else {
            if (cnt == 0 || ids[cnt - 1] != bot[i]) { if (cnt < max_ids) ids[cnt] = bot[i]; cnt++; }
            i++; j++;
        }
    }
    free(bot); free(top);
    return cnt;
}

void orc_size_hist(int kind, int t, int nb, const int *csize, const int *b3_can, int64_t *hist, int maxsize)
{
    int i, r;
    memset(hist, 0, sizeof(int64_t) * (size_t)(maxsize + 1));
    for (i = 1; i <= t; ++i) if (csize[i] > 0 && csize[i] <= maxsize) hist[csize[i]]++;
    if (kind == ORC_MIXED && b3_can && maxsize >= 1)
        for (r = 0; r < nb; ++r) if (b3_can[r] > t) hist[1]++;
}

/* ======================================================================== */
/* conductance                                                                */
/* ======================================================================== */

/* which bonds conduct: bond Sq/bondc.f:482-489; site MATLAB/ConductCalc.m:88-109;
 * mixed MATLAB/ConductCalc.m:134-160.  Labels may be reference or canonical. */
void orc_weights(int kind, int nb, const int *b1, const int *b2, const int *s, const int *b3,
                 int perccln, double g0, double gleak, double *w)
{
    int r, on;
    for (r = 0; r < nb; ++r) {
        if (kind == ORC_BOND) on = (b3[r] == perccln);
        else if (kind == ORC_SITE) on = (s[b1[r] - 1] == perccln && s[b2[r] - 1] == perccln);
        else on = (b3[r] == perccln && s[b1[r] - 1] == perccln && s[b2[r] - 1] == perccln);
        w[r] = on ? g0 : gleak;
    }
}

/* CSR adjacency of the full t x t matrix G (off-diagonals -w, neighbour ids ascending,
 * i.e. the column order of the dense row scans in Sq/bondc.f:499-505 and sprsin :723-746). */
typedef struct { int t; int64_t *ptr; int *col; double *val; double *diag; } csr;

static csr csr_build(int t, int nb, const int *b1, const int *b2, const double *w)
{
    csr A; int r, i; int64_t k, kk; int64_t *fill;
    A.t = t;
    A.ptr = (int64_t *)calloc((size_t)t + 2, sizeof(int64_t));
    for (r = 0; r < nb; ++r) { A.ptr[b1[r] + 1]++; A.ptr[b2[r] + 1]++; }
    for (i = 1; i <= t + 1; ++i) A.ptr[i] += A.ptr[i - 1];
    A.col = (int *)malloc(sizeof(int) * (size_t)(2 * (int64_t)nb + 1));
    A.val = (double *)malloc(sizeof(double) * (size_t)(2 * (int64_t)nb + 1));
    A.diag = (double *)calloc((size_t)t + 1, sizeof(double));
    fill = (int64_t *)malloc(sizeof(int64_t) * (size_t)(t + 2));
    memcpy(fill, A.ptr, sizeof(int64_t) * (size_t)(t + 2));
    for (r = 0; r < nb; ++r) {
        k = fill[b1[r]]++; A.col[k] = b2[r]; A.val[k] = -w[r];
        k = fill[b2[r]]++; A.col[k] = b1[r]; A.val[k] = -w[r];
    }
    free(fill);
    for (i = 1; i <= t; ++i) {                     /* insertion sort by column (<= 6 entries) */
        for (k = A.ptr[i] + 1; k < A.ptr[i + 1]; ++k) {
            int cc = A.col[k]; double vv = A.val[k];
            kk = k - 1;
            while (kk >= A.ptr[i] && A.col[kk] > cc) { A.col[kk + 1] = A.col[kk]; A.val[kk + 1] = A.val[kk]; kk--; }
            A.col[kk + 1] = cc; A.val[kk + 1] = vv;
        }
        {   /* G(i,i) = -rowsum, rowsum accumulated over ascending j (Sq/bondc.f:499-505) */
            double rowsum = 0.0;
            for (k = A.ptr[i]; k < A.ptr[i + 1]; ++k) rowsum = rowsum + A.val[k];
            A.diag[i] = -rowsum;
        }
    }
    return A;
}

static void csr_free(csr *A) { free(A->ptr); free(A->col); free(A->val); free(A->diag); }

/* interior operator: unknown j (1..N) is site j+m; neighbours outside rows 1..n-2 drop out
 * (Gtemp = G(m+1..t-m, m+1..t-m), Sq/bondc.f:520-524; sprsin thresh 1e-16 keeps every bond). */
static void interior_ax(const csr *A, int m, int N, const double *x, double *y, double thresh)
{
    int i, j; int64_t k;
    /* (rows are independent: large systems are spread over the host threads, OMP_NUM_THREADS; each row's sum keeps its order) */
#pragma omp parallel for private(j, k) schedule(static) if (N > ORC_OMP_MIN)
    for (i = 1; i <= N; ++i) {
        int site = i + m;
        double acc = A->diag[site] * x[i - 1];
        for (k = A->ptr[site]; k < A->ptr[site + 1]; ++k) {
            j = A->col[k] - m;
            if (j >= 1 && j <= N && fabs(A->val[k]) >= thresh) acc = acc + A->val[k] * x[j - 1];
        }
        y[i - 1] = acc;
    }
}

/* dsprstx on the same matrix (Sq/bondc.f:902-917): scatter form of A^T x */
static void interior_atx(const csr *A, int m, int N, const double *x, double *y, double thresh)
{
    int i, j; int64_t k;
    for (i = 1; i <= N; ++i) y[i - 1] = A->diag[i + m] * x[i - 1];
    for (i = 1; i <= N; ++i) {
        int site = i + m;
        for (k = A->ptr[site]; k < A->ptr[site + 1]; ++k) {
            j = A->col[k] - m;
            if (j >= 1 && j <= N && fabs(A->val[k]) >= thresh) y[j - 1] = y[j - 1] + A->val[k] * x[i - 1];
        }
    }
}

static void build_rhs(int m, int n, int nb, const int *b1, const int *b2, const double *w, double Va, double *rhs)
{
    int t = m * n, r;
    memset(rhs, 0, sizeof(double) * (size_t)(t - 2 * m));
    for (r = 0; r < nb; ++r)                         /* Sq/bondc.f:490-497 */
        if (b1[r] > t - 2 * m && b1[r] <= t - m && b2[r] > t - m)
            rhs[b1[r] - m - 1] = rhs[b1[r] - m - 1] - ((-w[r]) * Va);
}

static double snrm2(int N, const double *x)         /* Sq/bondc.f:867-884, itol <= 3 */
{
    double s = 0.0; int i;
    for (i = 0; i < N; ++i) s = s + x[i] * x[i];
    return sqrt(s);
}

/* Sq/bondc.f:554-592: V = [0 | Vint | Va], Iout = G~ V with G~ = sprsin(G, thresh) */
static void readout(const csr *A, int m, int n, double Va, double thresh, const double *Vint,
                    double *Gtop, double *Gbot)
{
    int t = m * n, i, l; int64_t k;
    double Ibot = 0.0, Itop = 0.0;
#define V_(s) ((s) <= m ? 0.0 : ((s) > t - m ? Va : Vint[(s) - m - 1]))
    for (l = 1; l <= m; ++l) {
        for (i = 0; i < 2; ++i) {
            int site = i == 0 ? l : l + t - m;
            double acc = A->diag[site] * V_(site);
            for (k = A->ptr[site]; k < A->ptr[site + 1]; ++k)
                if (fabs(A->val[k]) >= thresh) acc = acc + A->val[k] * V_(A->col[k]);
            if (i == 0) Ibot = Ibot + acc; else Itop = Itop + acc;
        }
    }
#undef V_
    *Gtop = Itop / Va;
    *Gbot = fabs(Ibot) / Va;
}

/* literal linbcg (Sq/bondc.f:750-838), itol = 2, x0 = Vint on entry (reference: zeros) */
int orc_conduct_literal(int m, int n, int nb, const int *b1, const int *b2, const double *w,
                        double Va, double tol, int itmax, double read_thresh,
                        double *Vint, double *Gtop, double *Gbot, int *iter, double *err)
{
    int t = m * n, N = t - 2 * m, j, it = 0;
    csr A = csr_build(t, nb, b1, b2, w);
    double *b = (double *)malloc(sizeof(double) * (size_t)N * 7);
    double *p = b + N, *pp = p + N, *r = pp + N, *rr = r + N, *z = rr + N, *zz = z + N;
    double ak, akden, bk, bkden = 1.0, bknum, bnrm, e = 0.0;
    const double sp_thresh = 1.0e-16;                /* Sq/bondc.f:538 */
    build_rhs(m, n, nb, b1, b2, w, Va, b);
    interior_ax(&A, m, N, Vint, r, sp_thresh);
    for (j = 0; j < N; ++j) { r[j] = b[j] - r[j]; rr[j] = r[j]; }
    for (j = 0; j < N; ++j) z[j] = b[j] / A.diag[j + 1 + m];
    bnrm = snrm2(N, z);
    for (j = 0; j < N; ++j) z[j] = r[j] / A.diag[j + 1 + m];
    while (it <= itmax) {
        it++;
        for (j = 0; j < N; ++j) zz[j] = rr[j] / A.diag[j + 1 + m];
        bknum = 0.0;
        for (j = 0; j < N; ++j) bknum = bknum + z[j] * rr[j];
        if (it == 1) { for (j = 0; j < N; ++j) { p[j] = z[j]; pp[j] = zz[j]; } }
        else {
            bk = bknum / bkden;
            for (j = 0; j < N; ++j) { p[j] = bk * p[j] + z[j]; pp[j] = bk * pp[j] + zz[j]; }
        }
        bkden = bknum;
        interior_ax(&A, m, N, p, z, sp_thresh);
        akden = 0.0;
        for (j = 0; j < N; ++j) akden = akden + z[j] * pp[j];
        ak = bknum / akden;
        interior_atx(&A, m, N, pp, zz, sp_thresh);
        for (j = 0; j < N; ++j) { Vint[j] = Vint[j] + ak * p[j]; r[j] = r[j] - ak * z[j]; rr[j] = rr[j] - ak * zz[j]; }
        for (j = 0; j < N; ++j) z[j] = r[j] / A.diag[j + 1 + m];
        e = snrm2(N, r) / bnrm;
        if (!(e > tol)) break;
    }
    readout(&A, m, n, Va, read_thresh, Vint, Gtop, Gbot);
    if (iter) *iter = it;
    if (err) *err = e;
    free(b); csr_free(&A);
    return 0;
}

/* same system, plain Jacobi-PCG (what linbcg reduces to on a symmetric matrix with rr = r):
 * one SpMV per iteration, same stopping measure. */
int orc_conduct_cg(int m, int n, int nb, const int *b1, const int *b2, const double *w,
                   double Va, double tol, int itmax, double read_thresh,
                   double *Vint, double *Gtop, double *Gbot, int *iter, double *err)
{
    int t = m * n, N = t - 2 * m, j, it = 0;
    csr A = csr_build(t, nb, b1, b2, w);
    double *b = (double *)malloc(sizeof(double) * (size_t)N * 4);
    double *p = b + N, *r = p + N, *q = r + N;
    double ak, akden, bk, bkden = 1.0, bknum, bnrm, e = 0.0, s;
    build_rhs(m, n, nb, b1, b2, w, Va, b);
    interior_ax(&A, m, N, Vint, r, 0.0);
    for (j = 0; j < N; ++j) r[j] = b[j] - r[j];
    s = 0.0;
    for (j = 0; j < N; ++j) { double zb = b[j] / A.diag[j + 1 + m]; s += zb * zb; }
    bnrm = sqrt(s);
    /* (large systems, N > ORC_OMP_MIN: the vector loops run on the host threads -- sums are then folded per thread, which
     * changes their rounding, not the algorithm; small systems run serially and are bit-reproducible) */
    while (it <= itmax) {
        it++;
        bknum = 0.0;
#pragma omp parallel for reduction(+:bknum) schedule(static) if (N > ORC_OMP_MIN)
        for (j = 0; j < N; ++j) bknum += r[j] * r[j] / A.diag[j + 1 + m];
        if (it == 1) {
#pragma omp parallel for schedule(static) if (N > ORC_OMP_MIN)
            for (j = 0; j < N; ++j) p[j] = r[j] / A.diag[j + 1 + m];
        } else {
            bk = bknum / bkden;
#pragma omp parallel for schedule(static) if (N > ORC_OMP_MIN)
            for (j = 0; j < N; ++j) p[j] = bk * p[j] + r[j] / A.diag[j + 1 + m];
        }
        bkden = bknum;
        interior_ax(&A, m, N, p, q, 0.0);
        akden = 0.0;
#pragma omp parallel for reduction(+:akden) schedule(static) if (N > ORC_OMP_MIN)
        for (j = 0; j < N; ++j) akden += q[j] * p[j];
        ak = bknum / akden;
        s = 0.0;
#pragma omp parallel for reduction(+:s) schedule(static) if (N > ORC_OMP_MIN)
        for (j = 0; j < N; ++j) { Vint[j] += ak * p[j]; r[j] -= ak * q[j]; s += r[j] * r[j]; }
        e = sqrt(s) / bnrm;
        if (!(e > tol)) break;
    }
    readout(&A, m, n, Va, read_thresh, Vint, Gtop, Gbot);
    if (iter) *iter = it;
    if (err) *err = e;
    free(b); csr_free(&A);
    return 0;
}

int orc_conduct_check(int m, int n, int nb, const int *b1, const int *b2, const double *w,
                      double Va, double read_thresh, const double *Vint,
                      double *Gtop, double *Gbot, double *err)
{
    int t = m * n, N = t - 2 * m, j;
    csr A = csr_build(t, nb, b1, b2, w);
    double *b = (double *)malloc(sizeof(double) * (size_t)N * 2);
    double *r = b + N, s = 0.0, sr = 0.0;
    build_rhs(m, n, nb, b1, b2, w, Va, b);
    interior_ax(&A, m, N, Vint, r, 0.0);
    for (j = 0; j < N; ++j) {
        double zb = b[j] / A.diag[j + 1 + m], rj = b[j] - r[j];
        s += zb * zb; sr += rj * rj;
    }
    if (err) *err = sqrt(sr) / sqrt(s);
    readout(&A, m, n, Va, read_thresh, Vint, Gtop, Gbot);
    free(b); csr_free(&A);
    return 0;
}

double orc_cg_time_iters(int m, int n, int nb, const int *b1, const int *b2, const double *w,
                         double Va, int iters)
{
    int t = m * n, N = t - 2 * m, j, it;
    csr A = csr_build(t, nb, b1, b2, w);
    double *b = (double *)calloc((size_t)N * 5, sizeof(double));
    double *p = b + N, *r = p + N, *q = r + N, *x = q + N;
    double ak, akden, bk, bkden = 1.0, bknum, s = 0.0, sec;
    struct timespec t0, t1;
    build_rhs(m, n, nb, b1, b2, w, Va, b);
    for (j = 0; j < N; ++j) r[j] = b[j];
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (it = 1; it <= iters; ++it) {
        bknum = 0.0;
        for (j = 0; j < N; ++j) bknum += r[j] * r[j] / A.diag[j + 1 + m];
        if (it == 1) for (j = 0; j < N; ++j) p[j] = r[j] / A.diag[j + 1 + m];
        else { bk = bknum / bkden; for (j = 0; j < N; ++j) p[j] = bk * p[j] + r[j] / A.diag[j + 1 + m]; }
        bkden = bknum;
        interior_ax(&A, m, N, p, q, 0.0);
        akden = 0.0;
        for (j = 0; j < N; ++j) akden += q[j] * p[j];
        ak = bknum / akden;
        for (j = 0; j < N; ++j) { x[j] += ak * p[j]; r[j] -= ak * q[j]; s += r[j] * r[j]; }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    sec = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    if (s < 0) sec = -sec;
    free(b); csr_free(&A);
    return sec;
}
