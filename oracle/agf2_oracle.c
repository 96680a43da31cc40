/*
 * TEST INFRASTRUCTURE ONLY.  Plain-C restatement of the reference's DF-AGF2(None,0) moment loop.
 * Nothing in the product path (auxgf_b200/) links, imports or calls this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it, and only as the checker.
 *
 * What it restates (citations relative to /root/reference):
 *   oracle_build_part_loop_rhf  <- auxgf/lib/libagf2/optragf2.c:117-295
 *       per occupied pole i in [istart, iend):
 *         xja[x,j,a] = sum_Q ixq[i,x,Q] qja[Q,j,a]      (optragf2.c:175-188)
 *         xia[x,j,a] = sum_Q ixq[j,x,Q] qja[Q,i,a]      (optragf2.c:121-159)
 *         eja[j,a]   = ei[i] + ei[j] - ea[a]            (optragf2.c:195-199)
 *         vv [x,y] += sum_ja xja[x,ja] (2 xja[y,ja] - xia[y,ja])            (optragf2.c:206-239)
 *         vev[x,y] += sum_ja xja[x,ja] eja[ja] (2 xja[y,ja] - xia[y,ja])    (optragf2.c:246-290)
 *   oracle_build_part_loop_uhf  <- auxgf/lib/libagf2/optuagf2.c:125-315
 *       same with aa Coulomb + ab Coulomb - aa exchange, unit coefficients.
 * The `os`/`ss` generalisation (Coulomb coefficient os+ss, exchange coefficient ss; UHF: ss on the
 * same-spin terms, os on the opposite-spin term) follows auxgf/aux/dfrmp2.py:82-84,107-109 and
 * auxgf/aux/dfump2.py:99-100,131-132; os = ss = 1 reproduces the C loops above exactly.
 *
 * Deliberately naive (triple loops, no BLAS): an independent statement of the arithmetic, not a
 * fast implementation.  Sizes and indices are 64-bit (the reference's uint32 products overflow at
 * the large configs).
 *
 * Parity pin: tests/test_oracle.py checks this file against oracle/_ref/agf2.so (the reference's
 * own C loop compiled from /root/reference by oracle/Makefile) and against the committed golden
 * vectors in tests/golden/ that were generated from that same binary (oracle/gen_golden.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* x[r, c] = sum_q a[r, q] * b[q, c * cstride_b ... ]   with explicit strides, naive */
static void form_block(const double *ixq_i, /* (nphys, naux) rows of pole i            */
                       const double *q_j,   /* &qja[0, j, 0]; element (Q, a) at Q*ldq + a */
                       int64_t nphys, int64_t nvir, int64_t naux, int64_t ldq,
                       double *out /* (nphys, nvir) */)
{
    for (int64_t x = 0; x < nphys; x++) {
        double *o = out + x * nvir;
        for (int64_t a = 0; a < nvir; a++) o[a] = 0.0;
        for (int64_t q = 0; q < naux; q++) {
            const double l = ixq_i[x * naux + q];
            const double *r = q_j + q * ldq;
            for (int64_t a = 0; a < nvir; a++) o[a] += l * r[a];
        }
    }
}

/* vv[x,y] += sum_a u[x,a] * w[y,a];  vev[x,y] += sum_a u[x,a] * e[a] * w[y,a] */
static void accumulate(const double *u, const double *w, const double *e,
                       int64_t nphys, int64_t nvir, double *vv, double *vev)
{
    for (int64_t x = 0; x < nphys; x++)
        for (int64_t y = 0; y < nphys; y++) {
            double s0 = 0.0, s1 = 0.0;
            const double *ux = u + x * nvir, *wy = w + y * nvir;
            for (int64_t a = 0; a < nvir; a++) {
                const double t = ux[a] * wy[a];
                s0 += t;
                s1 += t * e[a];
            }
            vv[x * nphys + y] += s0;
            vev[x * nphys + y] += s1;
        }
}

int oracle_build_part_loop_rhf(const double *ixq, const double *qja, const double *ei,
                               const double *ea, int64_t nphys, int64_t nocc, int64_t nvir,
                               int64_t naux, int64_t istart, int64_t iend, double os_factor,
                               double ss_factor, double *vv, double *vev)
{
    const double coul = os_factor + ss_factor, exch = ss_factor;
    const int64_t ldq = nocc * nvir;
    double *xja = malloc(sizeof(double) * nphys * nvir);
    double *xia = malloc(sizeof(double) * nphys * nvir);
    double *w = malloc(sizeof(double) * nphys * nvir);
    double *e = malloc(sizeof(double) * nvir);
    if (!xja || !xia || !w || !e) return 1;
    for (int64_t i = istart; i < iend; i++)
        for (int64_t j = 0; j < nocc; j++) {
            form_block(ixq + i * nphys * naux, qja + j * nvir, nphys, nvir, naux, ldq, xja);
            form_block(ixq + j * nphys * naux, qja + i * nvir, nphys, nvir, naux, ldq, xia);
            for (int64_t k = 0; k < nphys * nvir; k++) w[k] = coul * xja[k] - exch * xia[k];
            for (int64_t a = 0; a < nvir; a++) e[a] = ei[i] + ei[j] - ea[a];
            accumulate(xja, w, e, nphys, nvir, vv, vev);
        }
    free(xja); free(xia); free(w); free(e);
    return 0;
}

int oracle_build_part_loop_uhf(const double *ixq_a, const double *qja_a, const double *qja_b,
                               const double *ei_a, const double *ei_b, const double *ea_a,
                               const double *ea_b, int64_t nphys, int64_t nocc_a, int64_t nocc_b,
                               int64_t nvir_a, int64_t nvir_b, int64_t naux, int64_t istart,
                               int64_t iend, double os_factor, double ss_factor, double *vv,
                               double *vev)
{
    const int64_t nvmax = nvir_a > nvir_b ? nvir_a : nvir_b;
    double *xja = malloc(sizeof(double) * nphys * nvmax);
    double *xia = malloc(sizeof(double) * nphys * nvmax);
    double *w = malloc(sizeof(double) * nphys * nvmax);
    double *e = malloc(sizeof(double) * nvmax);
    if (!xja || !xia || !w || !e) return 1;
    for (int64_t i = istart; i < iend; i++) {
        /* same-spin: ss * xja_aa (xja_aa - xia_aa)   (optuagf2.c:211-225,242-256) */
        for (int64_t j = 0; j < nocc_a; j++) {
            form_block(ixq_a + i * nphys * naux, qja_a + j * nvir_a, nphys, nvir_a, naux,
                       nocc_a * nvir_a, xja);
            form_block(ixq_a + j * nphys * naux, qja_a + i * nvir_a, nphys, nvir_a, naux,
                       nocc_a * nvir_a, xia);
            for (int64_t k = 0; k < nphys * nvir_a; k++) w[k] = ss_factor * (xja[k] - xia[k]);
            for (int64_t a = 0; a < nvir_a; a++) e[a] = ei_a[i] + ei_a[j] - ea_a[a];
            accumulate(xja, w, e, nphys, nvir_a, vv, vev);
        }
        /* opposite-spin: os * xja_ab xja_ab          (optuagf2.c:226-240) */
        for (int64_t j = 0; j < nocc_b; j++) {
            form_block(ixq_a + i * nphys * naux, qja_b + j * nvir_b, nphys, nvir_b, naux,
                       nocc_b * nvir_b, xja);
            for (int64_t k = 0; k < nphys * nvir_b; k++) w[k] = os_factor * xja[k];
            for (int64_t a = 0; a < nvir_b; a++) e[a] = ei_a[i] + ei_b[j] - ea_b[a];
            accumulate(xja, w, e, nphys, nvir_b, vv, vev);
        }
    }
    free(xja); free(xia); free(w); free(e);
    return 0;
}
