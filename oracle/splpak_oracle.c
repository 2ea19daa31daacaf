/*
 * splpak_oracle.c -- CPU restatement of the reference's fit-and-evaluate path.
 * TEST INFRASTRUCTURE ONLY (see splpak_oracle.h).  Each function follows the cited
 * lines of /root/reference/src/splpak.F90 statement by statement, keeping the
 * reference's evaluation order so that roundoff matches an unfused Fortran build
 * (compile with -O2 -ffp-contract=off).
 *
 * Index conventions: the reference is 1-based.  Arrays that the reference indexes
 * from 1 are accessed through the F1() macro so the index arithmetic reads exactly as
 * in the source.  Integer state that can exceed 2^31 at the large configs is held
 * in long long; for every size the reference can represent the values are identical.
 */
#include "splpak_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define F1(arr, i) ((arr)[(i) - 1])

static oreal o_abs(oreal v) { return v < (oreal)0 ? -v : v; }
static oreal o_sqrt(oreal v) { return (oreal)sqrt((double)v); }
static long long ll_min(long long a, long long b) { return a < b ? a : b; }
static int i_min(int a, int b) { return a < b ? a : b; }
static int i_max(int a, int b) { return a > b ? a : b; }

int oracle_sizeof_real(void) { return (int)sizeof(oreal); }

void oracle_init(oracle_splpak *me) {
    memset(me, 0, sizeof(*me));
}

/* cfaerr, src/splpak.F90:399-407: ' IERR=' I5 then the message, on output_unit. */
static void cfaerr(oracle_splpak *me, int ierr, const char *mess) {
    me->nmsg++;
    if (me->quiet) return;
    if (ierr != 0) printf(" IERR=%5d\n", ierr);
    printf("%s\n", mess);
}

/* destroy_splpak, src/splpak.F90:136-165 */
void oracle_destroy(oracle_splpak *me, int ndim_present, int ndim) {
    int d;
    (void)ndim;
    if (ndim_present) {
        for (d = 0; d < ORACLE_MAXDIM; ++d) {
            me->dx[d] = 0; me->dxin[d] = 0; me->ib[d] = 0; me->ibmn[d] = 0; me->ibmx[d] = 0;
        }
    }
    me->mdim = 0;
    me->ilast = 0; me->isav = 0; me->iold = 0; me->np1 = 0;
    me->l = 0; me->il1 = 0; me->k = 0; me->k1 = 0;
    me->errsum = 0;
}

/* bascmp, src/splpak.F90:206-389 */
void oracle_bascmp(oracle_splpak *me, const oreal *x, const int *nderiv, const oreal *xmin,
                   const int *nodes, long long *icol_out, oreal *basm_out) {
    oreal xb, bas1, z, fact = 0, z1, basm;
    int idim, mdmid, ntyp, ngo;
    long long icol;

    icol = 0;                                   /* :220 */
    basm = (oreal)1.0;                          /* :223 */
    for (idim = 1; idim <= me->mdim; ++idim) {  /* :224 */
        /* Horner, :227-228 */
        mdmid = me->mdim + 1 - idim;
        icol = (long long)F1(nodes, mdmid) * icol + F1(me->ib, mdmid);

        /* node type, :231-240 */
        ntyp = 1;
        if (F1(me->ib, idim) > 1) {
            ntyp = 2;
            if (F1(me->ib, idim) >= F1(nodes, idim) - 2) ntyp = 3;
        }
        ngo = 3 * ntyp + F1(nderiv, idim) - 2;  /* :243 */

        xb = F1(xmin, idim) + (oreal)F1(me->ib, idim) * F1(me->dx, idim);  /* :246 */
        bas1 = (oreal)0.0;                      /* :249 */

        switch (ngo) {                          /* :251 */
        case 4:                                 /* :253-270 chapeau value */
            z = o_abs(F1(me->dxin, idim) * (F1(x, idim) - xb)) - (oreal)2.0;
            if (z < (oreal)0.0) {
                bas1 = (oreal)-0.25 * (z * z * z);
                z = z + (oreal)1.0;
                if (z < (oreal)0.0) bas1 = bas1 + z * z * z;
            }
            break;
        case 5:                                 /* :272-286 chapeau 1st derivative */
            z = F1(x, idim) - xb;
            fact = F1(me->dxin, idim);
            if (z < (oreal)0.0) fact = -fact;
            z = fact * z - (oreal)2.0;
            if (z < (oreal)0.0) {
                bas1 = (oreal)-0.75 * (z * z);
                z = z + (oreal)1.0;
                if (z < (oreal)0.0) bas1 = bas1 + (oreal)3.0 * (z * z);
                bas1 = fact * bas1;
            }
            break;
        case 6:                                 /* :288-300 chapeau 2nd derivative */
            fact = F1(me->dxin, idim);
            z = fact * o_abs(F1(x, idim) - xb) - (oreal)2.0;
            if (z < (oreal)0.0) {
                bas1 = (oreal)-1.5 * z;
                z = z + (oreal)1.0;
                if (z < (oreal)0.0) bas1 = bas1 + (oreal)6.0 * z;
                bas1 = (fact * fact) * bas1;
            }
            break;
        case 2:
        case 8:                                 /* :302-322 edge 1st derivative */
            if (ngo == 2) fact = -F1(me->dxin, idim);
            else          fact =  F1(me->dxin, idim);
            z = fact * (F1(x, idim) - xb) + (oreal)2.0;
            if (z > (oreal)0.0) {
                if (z < (oreal)2.0) {
                    bas1 = (oreal)1.5 * (z * z);
                    z = z - (oreal)1.0;
                    if (z > (oreal)0.0) bas1 = bas1 - (oreal)3.0 * (z * z);
                    bas1 = fact * bas1;
                } else {
                    bas1 = (oreal)3.0 * fact;
                }
            }
            break;
        case 3:
        case 9:                                 /* :324-340 edge 2nd derivative */
            if (ngo == 3) fact = -F1(me->dxin, idim);
            else          fact =  F1(me->dxin, idim);
            z = fact * (F1(x, idim) - xb) + (oreal)2.0;
            z1 = z - (oreal)1.0;
            if (o_abs(z1) < (oreal)1.0) {
                bas1 = (oreal)3.0 * z;
                if (z1 > (oreal)0.0) bas1 = bas1 - (oreal)6.0 * z1;
                bas1 = (fact * fact) * bas1;
            }
            break;
        default:                                /* :342-379 edge value (1 left, 7 right, or anything else) */
            if (ngo != 7) z = F1(me->dxin, idim) * (xb - F1(x, idim)) + (oreal)2.0;
            else          z = F1(me->dxin, idim) * (F1(x, idim) - xb) + (oreal)2.0;
            if (z > (oreal)0.0) {
                if (z < (oreal)2.0) {
                    bas1 = (oreal)0.5 * (z * z * z);
                    z = z - (oreal)1.0;
                    if (z > (oreal)0.0) bas1 = bas1 - z * z * z;
                } else {
                    bas1 = (oreal)3.0 * z - (oreal)3.0;
                }
            }
            break;
        }
        basm = basm * bas1;                     /* :383 */
    }
    *icol_out = icol + 1;                       /* :387 */
    *basm_out = basm;
}

/* Hand one finished row to the solver (or to the test sink).  src/splpak.F90:849-854 / :1025-1031 */
static void emit_row(oracle_splpak *me, long long irow, oreal *coef, long long ncol, oreal rhs,
                     oreal *work_nwrk1, long long nwlft, int *ierror) {
    oreal reserr = 0;
    int lserr = 0;
    if (me->row_sink) {
        me->row_sink(me->row_sink_ctx, irow, coef, ncol, rhs);
        return;
    }
    oracle_suprls(me, irow, coef, ncol, rhs, work_nwrk1, nwlft, coef, &reserr, &lserr);
    if (lserr != 0) {
        me->last_suprls_ier = lserr;
        *ierror = 107;
        cfaerr(me, *ierror, " splcc or splcw - suprls failure (this usually indicates insufficient input data)");
    }
}

/* splcw, src/splpak.F90:512-1060 */
void oracle_splcw(oracle_splpak *me, int ndim, const oreal *xdata, int l1xdat, const oreal *ydata,
                  const oreal *wdata, long long ndata, const oreal *xmin, const oreal *xmax,
                  const int *nodes, oreal xtrap, oreal *coef, long long ncf, oreal *work,
                  long long nwrk, int *ierror) {
    oreal x[ORACLE_MAXDIM];
    int nderiv[ORACLE_MAXDIM], in[ORACLE_MAXDIM], inmx[ORACLE_MAXDIM];
    oreal xrng, swght, rowwt, rhs, basm, totlwt, bump, wtprrc, expect, dcwght;
    long long ncol, nwrk1, mdata, nwlft, irow, idata, icol, iin, nrect;
    int idim, nod, it, idimc, idm, jdm, inidim, boundary, carry;
    const oreal spcrit = (oreal)0.75;           /* :696 */
    rhs = (oreal)0.0;                           /* (undefined in the reference until the first row) */

    oracle_destroy(me, 1, ndim);                /* :710 */

    *ierror = 0;                                /* :716 */
    me->mdim = ndim;
    if (me->mdim < 1) {                         /* :718 */
        *ierror = 101;
        cfaerr(me, *ierror, " splcc or splcw - NDIM is less than 1");
        return;
    }

    ncol = 1;                                   /* :725 */
    for (idim = 1; idim <= me->mdim; ++idim) {
        nod = F1(nodes, idim);
        if (nod < 4) {                          /* :728 */
            *ierror = 102;
            cfaerr(me, *ierror, " splcc or splcw - NODES(IDIM) is less than 4 for some IDIM");
            return;
        }
        ncol = ncol * nod;                      /* :737 */
        xrng = F1(xmax, idim) - F1(xmin, idim);
        if (xrng == (oreal)0.0) {               /* :739 */
            *ierror = 103;
            cfaerr(me, *ierror, " splcc or splcw - XMIN(IDIM) equals XMAX(IDIM) for some IDIM");
            return;
        }
        F1(me->dx, idim) = xrng / (oreal)(nod - 1);           /* :747 */
        F1(me->dxin, idim) = (oreal)1.0 / F1(me->dx, idim);   /* :748 */
        F1(nderiv, idim) = 0;
    }
    if (ncol > ncf) {                           /* :751 */
        *ierror = 104;
        cfaerr(me, *ierror, " splcc or splcw - NCF (size of COEF) is too small");
        return;
    }
    nwrk1 = 1;                                  /* :757 */
    mdata = ndata;
    if (mdata < 1) {                            /* :759 */
        *ierror = 105;
        cfaerr(me, *ierror, " splcc or splcw - Ndata Is less than 1");
        return;
    }

    swght = xtrap;                              /* :769 */
    if (swght != (oreal)0.0) nwrk1 = ncol + 1;  /* :772 */
    nwlft = nwrk - nwrk1 + 1;                   /* :775 */
    if (nwlft < 1) {
        *ierror = 106;
        cfaerr(me, *ierror, " splcc or splcw - NWRK (size of WORK) is too small");
        return;
    }
    irow = 0;                                   /* :782 */
    rowwt = (oreal)1.0;                         /* :785 */

    /* ---- data rows, :788-855 ---- */
    for (idata = 1; idata <= mdata; ++idata) {
        if (F1(wdata, 1) >= (oreal)0.0) {       /* :796 */
            rowwt = F1(wdata, idata);
            if (rowwt == (oreal)0.0) continue;  /* :799 */
        }
        irow = irow + 1;
        rhs = rowwt * F1(ydata, idata);         /* :806 */
        for (idim = 1; idim <= me->mdim; ++idim)
            F1(x, idim) = xdata[(idim - 1) + (long long)l1xdat * (idata - 1)];   /* :808 */

        for (icol = 1; icol <= ncol; ++icol) F1(coef, icol) = (oreal)0.0;        /* :814-816 */

        for (idim = 1; idim <= me->mdim; ++idim) {                               /* :821-827 */
            nod = F1(nodes, idim);
            it = (int)(F1(me->dxin, idim) * (F1(x, idim) - F1(xmin, idim)));     /* truncation toward zero */
            F1(me->ibmn, idim) = i_min(i_max(it - 1, 0), nod - 2);
            F1(me->ib, idim) = F1(me->ibmn, idim);
            F1(me->ibmx, idim) = i_max(i_min(it + 2, nod - 1), 1);
        }

        for (;;) {                              /* basis_index, :829-846 */
            oracle_bascmp(me, x, nderiv, xmin, nodes, &icol, &basm);
            F1(coef, icol) = rowwt * basm;      /* :837 */
            carry = 1;
            for (idim = 1; idim <= me->mdim; ++idim) {
                F1(me->ib, idim) = F1(me->ib, idim) + 1;
                if (F1(me->ib, idim) <= F1(me->ibmx, idim)) { carry = 0; break; }
                F1(me->ib, idim) = F1(me->ibmn, idim);
            }
            if (carry) break;
        }

        emit_row(me, irow, coef, ncol, rhs, work + (nwrk1 - 1), nwlft, ierror);  /* :849-854 */
    }

    /* ---- smoothing rows for data sparse areas, :862-1048 ---- */
    if (swght != (oreal)0.0) {
        rhs = (oreal)0.0;                       /* :866 */
        nrect = 1;
        for (idim = 1; idim <= me->mdim; ++idim) {     /* :871-875 */
            F1(in, idim) = 0;
            F1(inmx, idim) = F1(nodes, idim) - 1;
            nrect = nrect * F1(inmx, idim);
        }
        for (iin = 1; iin <= ncol; ++iin) F1(work, iin) = (oreal)0.0;   /* :879-881 */

        totlwt = (oreal)0.0;                    /* :885 */
        for (idata = 1; idata <= mdata; ++idata) {
            bump = (oreal)1.0;
            if (F1(wdata, 1) >= (oreal)0.0) bump = F1(wdata, idata);    /* :890 */
            if (bump == (oreal)0.0) continue;
            iin = 0;
            for (idimc = 1; idimc <= me->mdim; ++idimc) {               /* :895-902 */
                idim = me->mdim + 1 - idimc;
                inidim = (int)(F1(me->dxin, idim) *
                               (xdata[(idim - 1) + (long long)l1xdat * (idata - 1)] - F1(xmin, idim)) +
                               (oreal)0.5);
                /* :899 -- the bare `cycle` only skips this dimension (the quirk in SURVEY 8.0) */
                if (inidim < 0 || inidim > F1(inmx, idim)) continue;
                iin = (long long)(F1(inmx, idim) + 1) * iin + inidim;
            }
            F1(work, iin + 1) = F1(work, iin + 1) + bump;               /* :905 */
            totlwt = totlwt + bump;
        }

        wtprrc = totlwt / (oreal)nrect;         /* :910 */
        iin = 0;

        for (;;) {                              /* node_index, :921-1046 */
            iin = iin + 1;
            expect = wtprrc;
            for (idim = 1; idim <= me->mdim; ++idim)                    /* :927-929 */
                if (F1(in, idim) == 0 || F1(in, idim) == F1(inmx, idim)) expect = (oreal)0.5 * expect;

            if (F1(work, iin) < spcrit * expect) {                      /* :936 */
                dcwght = expect - F1(work, iin);
                for (idim = 1; idim <= me->mdim; ++idim) {              /* :939-956 */
                    inidim = F1(in, idim);
                    F1(x, idim) = F1(xmin, idim) + (oreal)inidim * F1(me->dx, idim);
                    F1(me->ibmn, idim) = inidim - 1;
                    F1(me->ibmx, idim) = inidim + 1;
                    if (inidim == 0) F1(me->ibmn, idim) = 0;
                    if (inidim == F1(inmx, idim)) F1(me->ibmx, idim) = F1(inmx, idim);
                    F1(me->ib, idim) = F1(me->ibmn, idim);
                }
                dcwght = swght * dcwght;        /* :960 */
                for (icol = 1; icol <= ncol; ++icol) F1(coef, icol) = (oreal)0.0;   /* :965-967 */

                for (idm = 1; idm <= me->mdim; ++idm) {                 /* :974 */
                    for (jdm = idm; jdm <= me->mdim; ++jdm) {
                        for (idim = 1; idim <= me->mdim; ++idim) F1(nderiv, idim) = 0;
                        boundary = 1;
                        rowwt = (oreal)2.0 * dcwght;                    /* :983 */
                        if (jdm == idm) {
                            rowwt = dcwght;
                            F1(nderiv, jdm) = 2;
                            if (F1(in, idm) != 0 && F1(in, idm) != F1(inmx, idm)) boundary = 0;
                        }
                        if (boundary) {                                 /* :992-1000 */
                            F1(nderiv, idm) = 1;
                            F1(nderiv, jdm) = 1;
                        }
                        irow = irow + 1;

                        for (;;) {                                      /* basis, :1003-1022 */
                            oracle_bascmp(me, x, nderiv, xmin, nodes, &icol, &basm);
                            F1(coef, icol) = rowwt * basm;
                            carry = 1;
                            for (idim = 1; idim <= me->mdim; ++idim) {
                                F1(me->ib, idim) = F1(me->ib, idim) + 1;
                                if (F1(me->ib, idim) <= F1(me->ibmx, idim)) { carry = 0; break; }
                                F1(me->ib, idim) = F1(me->ibmn, idim);
                            }
                            if (carry) break;
                        }
                        emit_row(me, irow, coef, ncol, rhs, work + (nwrk1 - 1), nwlft, ierror);  /* :1025 */
                    }
                }
            }

            carry = 1;                          /* :1038-1044 */
            for (idim = 1; idim <= me->mdim; ++idim) {
                F1(in, idim) = F1(in, idim) + 1;
                if (F1(in, idim) <= F1(inmx, idim)) { carry = 0; break; }
                F1(in, idim) = 0;
            }
            if (carry) break;
        }
    }

    /* final solve, :1051-1058 */
    if (!me->row_sink) {
        oreal reserr = 0;
        int lserr = 0;
        irow = 0;
        oracle_suprls(me, irow, coef, ncol, rhs, work + (nwrk1 - 1), nwlft, coef, &reserr, &lserr);
        if (lserr != 0) {
            me->last_suprls_ier = lserr;
            *ierror = 107;
            cfaerr(me, *ierror, " splcc or splcw - suprls failure (this usually indicates insufficient input data)");
        }
    }
}

/* splcc, src/splpak.F90:421-446 */
void oracle_splcc(oracle_splpak *me, int ndim, const oreal *xdata, int l1xdat, const oreal *ydata,
                  long long ndata, const oreal *xmin, const oreal *xmax, const int *nodes,
                  oreal xtrap, oreal *coef, long long ncf, oreal *work, long long nwrk, int *ierror) {
    const oreal wdata[1] = {(oreal)-1.0};       /* :440 */
    oracle_splcw(me, ndim, xdata, l1xdat, ydata, wdata, ndata, xmin, xmax, nodes, xtrap, coef, ncf,
                 work, nwrk, ierror);
}

/* splde, src/splpak.F90:1089-1240 */
oreal oracle_splde(oracle_splpak *me, int ndim, const oreal *x, const int *nderiv, const oreal *coef,
                   const oreal *xmin, const oreal *xmax, const int *nodes, int *ierror) {
    oreal xrng, sum, basm;
    long long iibmx, iib, icof;
    int idim, nod, it, carry;

    *ierror = 0;                                /* :1166 */
    me->mdim = ndim;
    if (me->mdim < 1) {
        *ierror = 101;
        cfaerr(me, *ierror, " splfe or splde - NDIM is less than 1");
        return (oreal)0.0;                      /* function result undefined in the reference */
    }
    iibmx = 1;
    for (idim = 1; idim <= me->mdim; ++idim) {  /* :1175 */
        nod = F1(nodes, idim);
        if (nod < 4) {
            *ierror = 102;
            cfaerr(me, *ierror, " splfe or splde - NODES(IDIM) is less than  4for some IDIM");
            return (oreal)0.0;
        }
        xrng = F1(xmax, idim) - F1(xmin, idim);
        if (xrng == (oreal)0.0) {
            *ierror = 103;
            cfaerr(me, *ierror, " splfe or splde - XMIN(IDIM) = XMAX(IDIM) for some IDIM");
            return (oreal)0.0;
        }
        if (F1(nderiv, idim) < 0 || F1(nderiv, idim) > 2) {    /* :1190 -- sets 104, does NOT return */
            *ierror = 104;
            cfaerr(me, *ierror, " splde - NDERIV(IDIM) IS less than 0 or greater than 2 for some IDIM");
        }
        F1(me->dx, idim) = xrng / (oreal)(nod - 1);             /* :1197 */
        F1(me->dxin, idim) = (oreal)1.0 / F1(me->dx, idim);
        it = (int)(F1(me->dxin, idim) * (F1(x, idim) - F1(xmin, idim)));   /* :1201 */
        F1(me->ibmn, idim) = i_min(i_max(it - 1, 0), nod - 2);
        F1(me->ibmx, idim) = i_max(i_min(it + 2, nod - 1), 1);
        iibmx = iibmx * (F1(me->ibmx, idim) - F1(me->ibmn, idim) + 1);
        F1(me->ib, idim) = F1(me->ibmn, idim);
    }

    sum = (oreal)0.0;                           /* :1212 */
    iib = 0;
    for (;;) {                                  /* basis_index, :1215-1236 */
        iib = iib + 1;
        oracle_bascmp(me, x, nderiv, xmin, nodes, &icof, &basm);
        sum = sum + F1(coef, icof) * basm;      /* :1225 */
        carry = 1;
        if (iib < iibmx) {
            for (idim = 1; idim <= me->mdim; ++idim) {
                F1(me->ib, idim) = F1(me->ib, idim) + 1;
                if (F1(me->ib, idim) <= F1(me->ibmx, idim)) { carry = 0; break; }
                F1(me->ib, idim) = F1(me->ibmn, idim);
            }
        }
        if (carry) break;
    }
    return sum;                                 /* :1238 */
}

/* splfe, src/splpak.F90:1258-1275 */
oreal oracle_splfe(oracle_splpak *me, int ndim, const oreal *x, const oreal *coef, const oreal *xmin,
                   const oreal *xmax, const int *nodes, int *ierror) {
    int nderiv[ORACLE_MAXDIM];
    int d;
    for (d = 0; d < ORACLE_MAXDIM; ++d) nderiv[d] = 0;         /* :1272 */
    return oracle_splde(me, ndim, x, nderiv, coef, xmin, xmax, nodes, ierror);
}

/* suprls, src/splpak.F90:1375-1695 */
void oracle_suprls(oracle_splpak *me, long long i, const oreal *rowi, long long n, oreal bi, oreal *a,
                   long long nn, oreal *soln, oreal *err, int *ier) {
    oreal s, temp, temp1, cn, sn;
    long long j, ilj, ilnp, nreq, k, idiag, i1, i2 = 0, ii, jp1, lmkm1, j1, jdel, idj, iijd, i1jd, k11,
        k1m1, i11, np1mk, lmk, imov, iii, iiim, iim1, ilk, npk, ilii, npii;
    int complete_reduction;
    const oreal tol = (oreal)1.0e-18;           /* :1423 */

    *ier = 0;                                   /* :1425 */
    complete_reduction = (i <= 0);

    if (!complete_reduction) {
        if (i <= 1) {                           /* :1430 first call set-up */
            me->iold = 0;
            me->np1 = n + 1;
            me->l = nn / me->np1;               /* :1437 */
            me->ilast = 0;
            me->il1 = 0;
            me->k = 0;
            me->k1 = 0;
            me->errsum = (oreal)0.0;
            nreq = ((n + 5) * n + 2) / 2;       /* :1443 */
            if (nn < nreq) {
                *ier = 32;
                if (!me->quiet) {
                    printf(" nn   =  %lld\n", nn);
                    printf(" nreq =  %lld\n", nreq);
                }
                cfaerr(me, *ier, " suprls - insufficient scratch storage provided. at least ((N+5)*N+2)/2 locations needed");
                return;
            }
        }
        if ((i - me->iold) != 1) {              /* :1459 */
            *ier = 35;
            if (!me->quiet) {
                printf(" i    = %lld\n", i);
                printf(" me%%iold = %lld\n", me->iold);
            }
            cfaerr(me, *ier, " suprls - values of I not in sequence");
            return;
        }
        me->iold = i;                           /* :1468 store the row */
        for (j = 1; j <= n; ++j) {
            ilj = me->ilast + j;
            F1(a, ilj) = F1(rowi, j);
        }
        ilnp = me->ilast + me->np1;
        F1(a, ilnp) = bi;
        me->ilast = me->ilast + me->np1;
        me->isav = i;
        if (i < me->l) return;                  /* :1477 */
    }

    for (;;) {                                  /* main, :1481 */
        if (!complete_reduction) {
            if (me->k != 0) {                   /* :1485 */
                me->k1 = ll_min(me->k, n);
                idiag = -me->np1;
                if (me->l - me->k == 1) {
                    /* rotations for a single new row, :1488-1515 */
                    for (j = 1; j <= me->k1; ++j) {
                        idiag = idiag + (me->np1 - j + 2);
                        i1 = me->il1 + j;
                        if (o_abs(F1(a, i1)) <= tol)          s = o_sqrt(F1(a, idiag) * F1(a, idiag));
                        else if (o_abs(F1(a, idiag)) < tol)   s = o_sqrt(F1(a, i1) * F1(a, i1));
                        else s = o_sqrt(F1(a, idiag) * F1(a, idiag) + F1(a, i1) * F1(a, i1));
                        if (s == (oreal)0.0) continue;
                        temp = F1(a, idiag);
                        F1(a, idiag) = s;
                        s = (oreal)1.0 / s;
                        cn = temp * s;
                        sn = F1(a, i1) * s;
                        jp1 = j + 1;
                        for (j1 = jp1; j1 <= me->np1; ++j1) {
                            jdel = j1 - j;
                            idj = idiag + jdel;
                            temp = F1(a, idj);
                            i1jd = i1 + jdel;
                            F1(a, idj) = cn * temp + sn * F1(a, i1jd);
                            F1(a, i1jd) = -sn * temp + cn * F1(a, i1jd);
                        }
                    }
                } else {
                    /* Householder against the triangle, :1516-1549 */
                    for (j = 1; j <= me->k1; ++j) {
                        idiag = idiag + (me->np1 - j + 2);
                        i1 = me->il1 + j;
                        i2 = i1 + me->np1 * (me->l - me->k - 1);
                        s = F1(a, idiag) * F1(a, idiag);
                        for (ii = i1; ii <= i2; ii += me->np1) s = s + F1(a, ii) * F1(a, ii);
                        if (s == (oreal)0.0) continue;
                        temp = F1(a, idiag);
                        F1(a, idiag) = o_sqrt(s);
                        if (temp > (oreal)0.0) F1(a, idiag) = -F1(a, idiag);
                        temp = temp - F1(a, idiag);
                        temp1 = (oreal)1.0 / (temp * F1(a, idiag));
                        jp1 = j + 1;
                        for (j1 = jp1; j1 <= me->np1; ++j1) {
                            jdel = j1 - j;
                            idj = idiag + jdel;
                            s = temp * F1(a, idj);
                            for (ii = i1; ii <= i2; ii += me->np1) {
                                iijd = ii + jdel;
                                s = s + F1(a, ii) * F1(a, iijd);
                            }
                            s = s * temp1;
                            F1(a, idj) = F1(a, idj) + s * temp;
                            for (ii = i1; ii <= i2; ii += me->np1) {
                                iijd = ii + jdel;
                                F1(a, iijd) = F1(a, iijd) + s * F1(a, ii);
                            }
                        }
                    }
                }

                if (me->k >= n) {               /* :1551 triangle already complete */
                    lmkm1 = me->l - me->k;
                    for (ii = 1; ii <= lmkm1; ++ii) {
                        ilnp = me->il1 + ii * me->np1;
                        me->errsum = me->errsum + F1(a, ilnp) * F1(a, ilnp);
                    }
                    if (i <= 0) break;          /* exit main */
                    me->k = me->l;
                    me->ilast = me->il1;
                    me->l = me->k + (nn - me->ilast) / me->np1;   /* :1564 */
                    return;
                }
            }

            k11 = me->k1 + 1;                   /* :1569 */
            me->k1 = ll_min(me->l, n);
            if (me->l - me->k != 1) {
                k1m1 = me->k1 - 1;
                if (me->l > n) k1m1 = n;
                i1 = me->il1 + k11 - me->np1 - 1;
                /* Householder among the new rows, :1578-1609 */
                for (j = k11; j <= k1m1; ++j) {
                    i1 = i1 + (me->np1 + 1);
                    i2 = i1 + (me->l - j) * me->np1;
                    s = (oreal)0.0;
                    for (ii = i1; ii <= i2; ii += me->np1) s = s + F1(a, ii) * F1(a, ii);
                    if (s == (oreal)0.0) continue;
                    temp = F1(a, i1);
                    F1(a, i1) = o_sqrt(s);
                    if (temp > (oreal)0.0) F1(a, i1) = -F1(a, i1);
                    temp = temp - F1(a, i1);
                    temp1 = (oreal)1.0 / (temp * F1(a, i1));
                    jp1 = j + 1;
                    i11 = i1 + me->np1;
                    for (j1 = jp1; j1 <= me->np1; ++j1) {
                        jdel = j1 - j;
                        i1jd = i1 + jdel;
                        s = temp * F1(a, i1jd);
                        for (ii = i11; ii <= i2; ii += me->np1) {
                            iijd = ii + jdel;
                            s = s + F1(a, ii) * F1(a, iijd);
                        }
                        s = s * temp1;
                        i1jd = i1 + jdel;
                        F1(a, i1jd) = F1(a, i1jd) + s * temp;
                        for (ii = i11; ii <= i2; ii += me->np1) {
                            iijd = ii + jdel;
                            F1(a, iijd) = F1(a, iijd) + s * F1(a, ii);
                        }
                    }
                }
                if (me->l > n) {                /* :1610-1618 */
                    np1mk = me->np1 - me->k;
                    lmk = me->l - me->k;
                    for (ii = np1mk; ii <= lmk; ++ii) {
                        ilnp = me->il1 + ii * me->np1;
                        me->errsum = me->errsum + F1(a, ilnp) * F1(a, ilnp);
                    }
                }
            }
            imov = 0;                           /* :1620 squeeze */
            i1 = me->il1 + k11 - me->np1 - 1;
            for (ii = k11; ii <= me->k1; ++ii) {
                imov = imov + (ii - 1);
                i1 = i1 + me->np1 + 1;
                i2 = i1 + me->np1 - ii;
                for (iii = i1; iii <= i2; ++iii) {
                    iiim = iii - imov;
                    F1(a, iiim) = F1(a, iii);
                }
            }
            me->ilast = i2 - imov;              /* :1634 */
            me->il1 = me->ilast;
            if (i <= 0) break;                  /* exit main */
            me->k = me->l;
            me->l = me->k + (nn - me->ilast) / me->np1;   /* :1640 */
            return;
        }

        /* complete the reduction, :1645-1657 */
        complete_reduction = 0;
        me->l = me->isav;
        if (me->l < n) {
            *ier = 33;
            cfaerr(me, *ier, " suprls - array has too few rows.");
            return;
        }
        if (me->k == me->isav) break;           /* exit main */
    }

    me->ilast = (me->np1 * (me->np1 + 1)) / 2 - 1;   /* :1661 */
    if (F1(a, me->ilast - 1) == (oreal)0.0) {
        *ier = 34;
        cfaerr(me, *ier, " suprls - system is singular.");
        return;
    }

    F1(soln, n) = F1(a, me->ilast) / F1(a, me->ilast - 1);   /* :1670 */
    for (ii = 2; ii <= n; ++ii) {
        iim1 = ii - 1;
        me->ilast = me->ilast - ii;
        s = F1(a, me->ilast);
        for (k = 1; k <= iim1; ++k) {
            ilk = me->ilast - k;
            npk = me->np1 - k;
            s = s - F1(a, ilk) * F1(soln, npk);
        }
        me->k = k;                              /* :1680 */
        ilii = me->ilast - ii;
        if (F1(a, ilii) == (oreal)0.0) {
            *ier = 34;
            cfaerr(me, *ier, " suprls - system is singular.");
            return;
        }
        npii = me->np1 - ii;
        F1(soln, npii) = s / F1(a, ilii);
    }
    *err = o_sqrt(me->errsum);                  /* :1693 */
}

/* ------------------------------------------------------------------------------------ */
/* conveniences (loops over the functions above; no new arithmetic)                      */
/* ------------------------------------------------------------------------------------ */

int oracle_eval_batch(int ndim, const oreal *x, int l1x, long long nq, const int *nderiv,
                      const oreal *coef, const oreal *xmin, const oreal *xmax, const int *nodes,
                      oreal *out) {
    oracle_splpak me;
    long long q;
    int ierror = 0, worst = 0;
    oracle_init(&me);
    me.quiet = 1;
    for (q = 0; q < nq; ++q) {
        const oreal *xq = x + (long long)l1x * q;
        if (nderiv) out[q] = oracle_splde(&me, ndim, xq, nderiv, coef, xmin, xmax, nodes, &ierror);
        else        out[q] = oracle_splfe(&me, ndim, xq, coef, xmin, xmax, nodes, &ierror);
        if (ierror != 0) worst = ierror;
    }
    return worst;
}

typedef struct {
    oreal *rows, *rhs;
    long long ncol, maxrows, nrows;
} row_collect;

static void collect_sink(void *ctx, long long irow, const oreal *row, long long ncol, oreal rhs) {
    row_collect *c = (row_collect *)ctx;
    if (irow <= c->maxrows) {
        memcpy(c->rows + (irow - 1) * c->ncol, row, (size_t)ncol * sizeof(oreal));
        c->rhs[irow - 1] = rhs;
    }
    if (irow > c->nrows) c->nrows = irow;
}

long long oracle_rows(int ndim, const oreal *xdata, int l1xdat, const oreal *ydata, const oreal *wdata,
                      long long ndata, const oreal *xmin, const oreal *xmax, const int *nodes,
                      oreal xtrap, oreal *rows, oreal *rhs, long long maxrows) {
    oracle_splpak me;
    row_collect c;
    long long ncol = 1;
    oreal *coef, *work;
    int ierror = 0, d;
    for (d = 0; d < ndim; ++d) ncol *= nodes[d];
    if (ndim < 1 || ncol < 1) ncol = 1;
    coef = (oreal *)calloc((size_t)ncol, sizeof(oreal));
    work = (oreal *)calloc((size_t)ncol + 2, sizeof(oreal));
    oracle_init(&me);
    me.quiet = 1;
    c.rows = rows; c.rhs = rhs; c.ncol = ncol; c.maxrows = maxrows; c.nrows = 0;
    me.row_sink = collect_sink;
    me.row_sink_ctx = &c;
    oracle_splcw(&me, ndim, xdata, l1xdat, ydata, wdata, ndata, xmin, xmax, nodes, xtrap, coef, ncol,
                 work, ncol + 2, &ierror);
    free(coef);
    free(work);
    if (ierror != 0) return -(long long)ierror;
    return c.nrows;
}

typedef struct {
    oracle_splpak *solver;
    oreal *a;
    long long nn, base;
} steady_fwd;

static void steady_sink(void *ctx, long long irow, const oreal *row, long long ncol, oreal rhs) {
    steady_fwd *f = (steady_fwd *)ctx;
    oreal err = 0;
    int ier = 0;
    /* soln is only written on the final (i=0) call, which this sample never makes. */
    oracle_suprls(f->solver, f->base + irow, row, ncol, rhs, f->a, f->nn, (oreal *)row, &err, &ier);
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

double oracle_suprls_steady_sample(int ndim, const oreal *xdata, int l1xdat, const oreal *ydata,
                                   const oreal *wdata, long long m, const oreal *xmin,
                                   const oreal *xmax, const int *nodes) {
    oracle_splpak gen, solver;
    steady_fwd f;
    long long n = 1, tri, nn, j, c, p;
    oreal *a, *coef, *work;
    unsigned long long s = 0x9E3779B97F4A7C15ull;
    double t0, t1;
    int ierror = 0, d;

    for (d = 0; d < ndim; ++d) n *= nodes[d];
    tri = n * (n + 3) / 2;                      /* packed triangle incl. rhs column */
    nn = tri + m * (n + 1);                     /* room for exactly m new rows per reduction */
    a = (oreal *)malloc((size_t)nn * sizeof(oreal));
    coef = (oreal *)calloc((size_t)n, sizeof(oreal));
    work = (oreal *)calloc((size_t)n + 2, sizeof(oreal));
    if (!a || !coef || !work) { free(a); free(coef); free(work); return -1.0; }

    /* synthetic, well-conditioned full triangle: unit diagonal, small off-diagonals */
    p = 0;
    for (j = 1; j <= n; ++j) {
        for (c = j; c <= n + 1; ++c) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            a[p++] = (c == j) ? (oreal)1.0 : (oreal)(1e-3 * ((double)(s >> 11) / 9007199254740992.0 - 0.5));
        }
    }

    oracle_init(&solver);
    solver.quiet = 1;
    solver.np1 = n + 1;
    solver.k = n; solver.k1 = n;
    solver.il1 = tri; solver.ilast = tri;
    solver.iold = n; solver.isav = n;
    solver.l = solver.k + (nn - solver.ilast) / solver.np1;   /* = n + m */
    solver.errsum = 0;

    oracle_init(&gen);
    gen.quiet = 1;
    f.solver = &solver; f.a = a; f.nn = nn; f.base = n;
    gen.row_sink = steady_sink;
    gen.row_sink_ctx = &f;

    t0 = now_s();
    /* xtrap = 0: data rows only; each row = zero-fill + bascmp odometer + suprls, as in :788-855 */
    oracle_splcw(&gen, ndim, xdata, l1xdat, ydata, wdata, m, xmin, xmax, nodes, (oreal)0.0, coef, n,
                 work, n + 2, &ierror);
    t1 = now_s();

    free(a); free(coef); free(work);
    if (ierror != 0) return -(double)ierror;
    return t1 - t0;
}
