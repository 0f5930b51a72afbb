/* abi_c_driver.c -- calls libperc_b200.so exactly as a Fortran 77/ISO_C_BINDING driver would:
 * every scalar by reference, 1-based site ids, column-major border(nb,2).  Restates the driver
 * part of PROGRAM bondc (reference Fortran/Square/bondc.f:60-160): enumerate bonds, shuffle with
 * a caller-side RNG, fill tbonds = pb*nb bonds, label, test spanning, conductance.
 * usage: abi_c_driver [geometry-only]   (geometry-only needs no GPU)
 * prints one line "OK ..." on success; exit code != 0 on any error. */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include "../include/perc_abi.h"

static uint64_t rng = 88172645463325252ULL;
static uint32_t xorshift(void) { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (uint32_t)(rng >> 32); }

int main(int argc, char **argv)
{
    int32_t lattice = PERC_SQUARE, m = 40, n = 36, pbc = 0, device = 0, nb = 0, rc;
    int32_t i, j, t = m * n;
    rc = perc_geom_nb(&lattice, &m, &n, &pbc, &nb);
    if (rc || nb != 2 * m * n - m - n) { printf("FAIL geom_nb %d %d\n", rc, nb); return 1; }
    int32_t *b = malloc(sizeof(int32_t) * 2 * nb), *border = malloc(sizeof(int32_t) * 2 * nb);
    rc = perc_geom_bondlist(&lattice, &m, &n, &pbc, b);
    if (rc || b[0] != 1 || b[nb] != 2) { printf("FAIL bondlist %d\n", rc); return 1; }
    int32_t rn = 1, nn[6];
    rc = perc_geom_nearestn(&lattice, &m, &n, &pbc, &rn, nn);
    if (rc || nn[0] != 2 || nn[1] != 1 + m) { printf("FAIL nearestn %d\n", rc); return 1; }
    if (argc > 1) { printf("OK geometry nb=%d\n", nb); return 0; }

    int64_t h = 0;
    rc = perc_create(&h, &lattice, &m, &n, &pbc, &device);
    if (rc) { printf("FAIL perc_create rc=%d (no CUDA device? there is no CPU fallback)\n", rc); return 2; }
    for (i = 0; i < 2 * nb; ++i) border[i] = b[i];
    for (i = 0; i < nb; ++i) {                       /* Fisher-Yates on both columns, Sq/bond.f:142-150 */
        j = i + (int32_t)(xorshift() % (uint32_t)(nb - i));
        int32_t t1 = border[i], t2 = border[nb + i];
        border[i] = border[j]; border[nb + i] = border[nb + j];
        border[j] = t1; border[nb + j] = t2;
    }
    double pb = 0.56;
    int32_t tbonds = (int32_t)(pb * nb);              /* tbonds = pb*nb, Sq/bond.f:167 */
    int32_t *b3 = malloc(sizeof(int32_t) * nb), *c = malloc(sizeof(int32_t) * t);
    int32_t maxcs = 0, perccln = 0, perccls = 0;
    rc = perc_bond(&h, border, &tbonds, b3, c, &maxcs, &perccln, &perccls);
    if (rc) { printf("FAIL perc_bond rc=%d\n", rc); return 3; }
    int64_t occ = 0, csum = 0;
    for (i = 0; i < nb; ++i) occ += b3[i] != 0;
    for (i = 0; i < t; ++i) csum += c[i];
    if (occ != tbonds || csum != tbonds) { printf("FAIL counts %ld %ld %d\n", (long)occ, (long)csum, tbonds); return 4; }
    double Va = 1.0, g0 = 1.0, gleak = 1e-12, tol = 1e-12, thr = 1e-10, Gtop = 0, Gbot = 0, err = 0;
    int32_t itmax = 100000, iter = 0;
    if (perccln) {
        rc = perc_conduct(&h, &perccln, &Va, &g0, &gleak, &tol, &itmax, &thr, &Gtop, &Gbot, &iter, &err);
        if (rc) { printf("FAIL perc_conduct rc=%d\n", rc); return 5; }
        if (!(Gtop > 0) || !(Gbot > 0) || (Gtop - Gbot) > 1e-9 || (Gbot - Gtop) > 1e-9) {
            printf("FAIL conductance Gtop=%.15g Gbot=%.15g\n", Gtop, Gbot); return 6; }
    }
    rc = perc_destroy(&h);
    if (rc) { printf("FAIL destroy %d\n", rc); return 7; }
    printf("OK nb=%d tbonds=%d maxcs=%d perccln=%d perccls=%d Gtop=%.12f Gbot=%.12f iter=%d err=%.3e\n",
           nb, tbonds, maxcs, perccln, perccls, Gtop, Gbot, iter, err);
    return 0;
}
