// basis.cuh -- per-point 1-D basis evaluation and index box, device side.
//
// Follows bascmp (src/splpak.F90:206-389) for the arithmetic of ONE dimension; the N-D basis
// value is the product over dimensions (:383), formed by the callers.  Every floating operation
// goes through the *_rn intrinsics so nvcc cannot contract a*b+c into an FMA: the 1-D values are
// then bit-identical to an unfused CPU evaluation of the same formulas (the oracle is built with
// -ffp-contract=off).  FMA is used deliberately only in the accumulation loops of the callers.
#pragma once

#include "common.cuh"

__device__ __forceinline__ double spl_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double spl_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double spl_sub(double a, double b) { return __dsub_rn(a, b); }

// One-dimensional basis function (or its 1st/2nd derivative) of node ib evaluated at x.
//   node type (:231-240): ib <= 1 left-linear, 2 <= ib < nod-2 chapeau, ib >= nod-2 right-linear
//   ngo = 3*ntyp + nder - 2 (:243) selects the formula exactly as the reference's select case,
//   including the `case default` catch-all, so out-of-range nder behaves as in the reference.
__device__ __forceinline__ double spl_bas1(int ib, int nod, int nder, double x, double xmin,
                                           double dx, double dxin) {
    int ntyp = 1;
    if (ib > 1) {
        ntyp = 2;
        if (ib >= nod - 2) ntyp = 3;
    }
    const int ngo = 3 * ntyp + nder - 2;
    const double xb = spl_add(xmin, spl_mul((double)ib, dx));   // :246
    double bas1 = 0.0, z, fact, z1;
    switch (ngo) {
    case 4:   // chapeau value, :253-270
        z = spl_sub(fabs(spl_mul(dxin, spl_sub(x, xb))), 2.0);
        if (z < 0.0) {
            bas1 = spl_mul(-0.25, spl_mul(spl_mul(z, z), z));
            z = spl_add(z, 1.0);
            if (z < 0.0) bas1 = spl_add(bas1, spl_mul(spl_mul(z, z), z));
        }
        break;
    case 5:   // chapeau 1st derivative, :272-286
        z = spl_sub(x, xb);
        fact = dxin;
        if (z < 0.0) fact = -fact;
        z = spl_sub(spl_mul(fact, z), 2.0);
        if (z < 0.0) {
            bas1 = spl_mul(-0.75, spl_mul(z, z));
            z = spl_add(z, 1.0);
            if (z < 0.0) bas1 = spl_add(bas1, spl_mul(3.0, spl_mul(z, z)));
            bas1 = spl_mul(fact, bas1);
        }
        break;
    case 6:   // chapeau 2nd derivative, :288-300
        fact = dxin;
        z = spl_sub(spl_mul(fact, fabs(spl_sub(x, xb))), 2.0);
        if (z < 0.0) {
            bas1 = spl_mul(-1.5, z);
            z = spl_add(z, 1.0);
            if (z < 0.0) bas1 = spl_add(bas1, spl_mul(6.0, z));
            bas1 = spl_mul(spl_mul(fact, fact), bas1);
        }
        break;
    case 2:
    case 8:   // edge 1st derivative, :302-322
        fact = (ngo == 2) ? -dxin : dxin;
        z = spl_add(spl_mul(fact, spl_sub(x, xb)), 2.0);
        if (z > 0.0) {
            if (z < 2.0) {
                bas1 = spl_mul(1.5, spl_mul(z, z));
                z = spl_sub(z, 1.0);
                if (z > 0.0) bas1 = spl_sub(bas1, spl_mul(3.0, spl_mul(z, z)));
                bas1 = spl_mul(fact, bas1);
            } else {
                bas1 = spl_mul(3.0, fact);
            }
        }
        break;
    case 3:
    case 9:   // edge 2nd derivative, :324-340
        fact = (ngo == 3) ? -dxin : dxin;
        z = spl_add(spl_mul(fact, spl_sub(x, xb)), 2.0);
        z1 = spl_sub(z, 1.0);
        if (fabs(z1) < 1.0) {
            bas1 = spl_mul(3.0, z);
            if (z1 > 0.0) bas1 = spl_sub(bas1, spl_mul(6.0, z1));
            bas1 = spl_mul(spl_mul(fact, fact), bas1);
        }
        break;
    default:  // edge value: ngo 1 (left), 7 (right) or anything else, :342-379
        if (ngo != 7) z = spl_add(spl_mul(dxin, spl_sub(xb, x)), 2.0);
        else          z = spl_add(spl_mul(dxin, spl_sub(x, xb)), 2.0);
        if (z > 0.0) {
            if (z < 2.0) {
                bas1 = spl_mul(0.5, spl_mul(spl_mul(z, z), z));
                z = spl_sub(z, 1.0);
                if (z > 0.0) bas1 = spl_sub(bas1, spl_mul(spl_mul(z, z), z));
            } else {
                bas1 = spl_sub(spl_mul(3.0, z), 3.0);
            }
        }
        break;
    }
    return bas1;
}

// Value-only (nder = 0) 1-D basis, branch-free.  With s = 2-|u| (chapeau), 2-u (left edge) or 2+u
// (right edge), u = dxin*(x-xb), all three value formulas of bascmp (:253-270, :342-379) read
//   v = alpha*max(s,0)^3 - max(s-1,0)^3,  alpha = 1/4 (chapeau) or 1/2 (edge);  edge and s >= 2: 3s-3.
// Every intermediate equals the reference's up to an exact sign flip (z = -s) or an exact
// power-of-two scaling, so the result is bit-identical to spl_bas1(.., nder = 0, ..) without the
// divergent select-case.  Used by the hot kernels (evaluation with nderiv = 0, assembly).
__device__ __forceinline__ double spl_bas1_value(int ib, int nod, double x, double xmin, double dx,
                                                 double dxin) {
    const double xb = spl_add(xmin, spl_mul((double)ib, dx));       // :246
    const double u = spl_mul(dxin, spl_sub(x, xb));
    const bool is_l = ib <= 1;
    const bool edge = is_l || ib >= nod - 2;
    const double w = edge ? (is_l ? -u : u) : -fabs(u);
    const double s = spl_add(2.0, w);
    const double sp = fmax(s, 0.0);
    const double s3 = spl_mul(spl_mul(sp, sp), sp);
    const double m = fmax(spl_sub(s, 1.0), 0.0);
    const double m3 = spl_mul(spl_mul(m, m), m);
    double v = spl_sub(spl_mul(edge ? 0.5 : 0.25, s3), m3);
    if (edge && s >= 2.0) v = spl_sub(spl_mul(3.0, s), 3.0);
    return v;
}

// Index box of one dimension (:821-827 and :1201-1207):
//   it = trunc(dxin*(x-xmin)); ibmn = min(max(it-1,0),nod-2); ibmx = max(min(it+2,nod-1),1).
// The fixed 4-wide window ws = clamp(it-1, 0, nod-4) always covers [ibmn, ibmx]; the callers
// evaluate the four window nodes and zero the ones outside the reference's box, so the set of
// terms is exactly the reference's.
__device__ __forceinline__ void spl_box(double x, double xmin, double dxin, int nod, int &ws,
                                        int &ibmn, int &ibmx) {
    double t = spl_mul(dxin, spl_sub(x, xmin));
    // Fortran real->integer assignment truncates toward zero; far-away points are clamped first
    // (any it <= -2 or >= nod+2 gives the same box), which also defines the NaN case (it = 0).
    t = fmin(fmax(t, -4.0), (double)nod + 4.0);
    const int it = __double2int_rz(t);
    ibmn = min(max(it - 1, 0), nod - 2);
    ibmx = max(min(it + 2, nod - 1), 1);
    ws = min(max(it - 1, 0), nod - 4);
}

// Same as spl_window_weights below for nder = 0, through the branch-free value formula.
__device__ __forceinline__ void spl_window_weights_value(double x, double xmin, double dx, double dxin,
                                                         int nod, int &ws, double b[4]) {
    int ibmn, ibmx;
    spl_box(x, xmin, dxin, nod, ws, ibmn, ibmx);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int ib = ws + k;
        const double v = spl_bas1_value(ib, nod, x, xmin, dx, dxin);
        b[k] = (ib >= ibmn && ib <= ibmx) ? v : 0.0;
    }
}

// The four window weights of one dimension: b[k] = basis of node ws+k at x (0 outside the box).
__device__ __forceinline__ void spl_window_weights(double x, double xmin, double dx, double dxin,
                                                   int nod, int nder, int &ws, double b[4]) {
    int ibmn, ibmx;
    spl_box(x, xmin, dxin, nod, ws, ibmn, ibmx);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int ib = ws + k;
        const double v = spl_bas1(ib, nod, nder, x, xmin, dx, dxin);
        b[k] = (ib >= ibmn && ib <= ibmx) ? v : 0.0;
    }
}
