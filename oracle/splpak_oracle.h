/*
 * splpak_oracle.h -- CPU restatement of jacobwilliams/splpak's fit-and-evaluate path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle: it restates the reference's
 * algorithm (src/splpak.F90) in plain C so the CUDA path can be checked against it.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product library (libsplpak_b200.so) never links, loads or
 * falls back to anything in this directory.
 *
 * Pinning status (see oracle/README.md): the reference cannot be compiled here (no
 * Fortran compiler in the image), so the oracle is pinned against everything the
 * reference's own tests assert (test/splpak_test_linear.f90:56-89, test/splpak_test.f90:68-84)
 * plus analytic known answers and an independent numpy least-squares cross-check.
 * It is NOT pinned against outputs of the reference binary itself.
 *
 * Precision: real == double, or float when compiled with -DREAL32, mirroring
 * src/splpak.F90:33-41.  Compile with -O2 -ffp-contract=off so no FMA contraction changes
 * roundoff relative to an unfused Fortran build.
 */
#ifndef SPLPAK_ORACLE_H
#define SPLPAK_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#ifdef REAL32
typedef float oreal;
#else
typedef double oreal;
#endif

#define ORACLE_MAXDIM 8

/* splpak_type (src/splpak.F90:45-127): former COMMON block + former SAVE state of suprls. */
typedef struct oracle_splpak {
    int mdim;                       /* :95  */
    oreal dx[ORACLE_MAXDIM];        /* :96  */
    oreal dxin[ORACLE_MAXDIM];      /* :97  */
    int ib[ORACLE_MAXDIM];          /* :98  */
    int ibmn[ORACLE_MAXDIM];        /* :99  */
    int ibmx[ORACLE_MAXDIM];        /* :100 */
    long long ilast, isav, iold, np1, l, il1, k, k1; /* :103-110 */
    oreal errsum;                   /* :111 */
    /* oracle-only instrumentation (not in the reference) */
    int quiet;                      /* suppress cfaerr printing */
    long long nmsg;                 /* number of cfaerr calls */
    int last_suprls_ier;            /* last non-zero ier seen from suprls (32..35) */
    /* optional row sink: when set, splcw hands each row here INSTEAD of suprls
       (used by tests to get B, r for G = B^T B parity and by lstsq cross-checks). */
    void (*row_sink)(void *ctx, long long irow, const oreal *row, long long ncol, oreal rhs);
    void *row_sink_ctx;
} oracle_splpak;

int oracle_sizeof_real(void);
void oracle_init(oracle_splpak *me);

/* destroy_splpak, src/splpak.F90:136-165 */
void oracle_destroy(oracle_splpak *me, int ndim_present, int ndim);

/* bascmp, src/splpak.F90:206-389 */
void oracle_bascmp(oracle_splpak *me, const oreal *x, const int *nderiv, const oreal *xmin,
                   const int *nodes, long long *icol, oreal *basm);

/* splcw, src/splpak.F90:512-1060.  wdata has ndata entries, or 1 entry < 0 (all weights 1). */
void oracle_splcw(oracle_splpak *me, int ndim, const oreal *xdata, int l1xdat, const oreal *ydata,
                  const oreal *wdata, long long ndata, const oreal *xmin, const oreal *xmax,
                  const int *nodes, oreal xtrap, oreal *coef, long long ncf, oreal *work,
                  long long nwrk, int *ierror);

/* splcc, src/splpak.F90:421-446 */
void oracle_splcc(oracle_splpak *me, int ndim, const oreal *xdata, int l1xdat, const oreal *ydata,
                  long long ndata, const oreal *xmin, const oreal *xmax, const int *nodes,
                  oreal xtrap, oreal *coef, long long ncf, oreal *work, long long nwrk, int *ierror);

/* splde, src/splpak.F90:1089-1240 */
oreal oracle_splde(oracle_splpak *me, int ndim, const oreal *x, const int *nderiv, const oreal *coef,
                   const oreal *xmin, const oreal *xmax, const int *nodes, int *ierror);

/* splfe, src/splpak.F90:1258-1275 */
oreal oracle_splfe(oracle_splpak *me, int ndim, const oreal *x, const oreal *coef, const oreal *xmin,
                   const oreal *xmax, const int *nodes, int *ierror);

/* suprls, src/splpak.F90:1375-1695 */
void oracle_suprls(oracle_splpak *me, long long i, const oreal *rowi, long long n, oreal bi, oreal *a,
                   long long nn, oreal *soln, oreal *err, int *ier);

/* ---- conveniences built only from the functions above (no new arithmetic) ---- */

/* Scalar splde/splfe in a loop over nq points (x is (l1x, nq) column-major).  nderiv may be NULL. */
int oracle_eval_batch(int ndim, const oreal *x, int l1x, long long nq, const int *nderiv,
                      const oreal *coef, const oreal *xmin, const oreal *xmax, const int *nodes,
                      oreal *out);

/* Run splcw's row generation (data rows + constraint rows) into dense row storage.
   rows: (maxrows, ncol) row-major, rhs: (maxrows).  Returns number of rows, or -ierror. */
long long oracle_rows(int ndim, const oreal *xdata, int l1xdat, const oreal *ydata, const oreal *wdata,
                      long long ndata, const oreal *xmin, const oreal *xmax, const int *nodes,
                      oreal xtrap, oreal *rows, oreal *rhs, long long maxrows);

/* Steady-state suprls sample for the CPU baseline: prime a full n x n triangle (synthetic,
   well conditioned), then push m real data rows of the given problem through suprls and
   return the seconds spent in those m calls.  See bench.py (cpu_baseline.sample). */
double oracle_suprls_steady_sample(int ndim, const oreal *xdata, int l1xdat, const oreal *ydata,
                                   const oreal *wdata, long long m, const oreal *xmin,
                                   const oreal *xmax, const int *nodes);

#ifdef __cplusplus
}
#endif
#endif
