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
    const double t = spl_mul(dxin, spl_sub(x, xmin));
    // Fortran real->integer assignment truncates toward zero.  The conversion saturates (and maps NaN
    // to 0); far-away points are then clamped (any it <= -2 or >= nod+2 gives the same box).
    const int it = min(max(__double2int_rz(t), -4), nod + 4);
    ibmn = min(max(it - 1, 0), nod - 2);
    ibmx = max(min(it + 2, nod - 1), 1);
    ws = min(max(it - 1, 0), nod - 4);
}

// s < 0 ? 0 : s through the sign bit (three integer-pipe instructions instead of the DSETP/SEL/FSEL/LOP3
// sequence fmax() compiles to; -0 and NaN pass through, both harmless below).
__device__ __forceinline__ double spl_clamp0(double s) {
    const int hi = __double2hiint(s);
    const int keep = ~(hi >> 31);
    return __hiloint2double(hi & keep, __double2loint(s) & keep);
}

// The four window weights of one dimension for nder = 0 -- the hot-path form used by evaluation (splfe) and
// assembly.  Every value is bit-identical to the reference's bascmp (same operations on the same operands, every
// product that is not rounded in the reference is an exact power-of-two scaling here); what differs is the plumbing,
// which the instruction mix of the evaluation kernel showed to dominate (613 instructions per 3-D query, 267 FP64):
//   * it = trunc(t) through the saturating F2I (NaN -> 0);
//   * the sign of u is STATIC in the window position k: a chapeau node at k <= 1 has x >= xb (it >= ws + 1 >= ib
//     whenever the window is not clamped from below -- and then the node is a left-edge node), one at k >= 2 has
//     x < xb; left-edge nodes (ib <= 1) only ever sit at k <= 1, right-edge nodes (ib >= nod-2) at k >= 2.  So
//     s = 2 - u for k < 2 and s = 2 + u for k >= 2 is bascmp's -z / z in all three node types (:253, :352, :358);
//     a rounding disagreement between t and u can flip the sign only at |u| ~ 1e-15 at the peak of the node's
//     function, where the value changes by O(u^2) < 1e-28;
//   * max(s, 0) as (s + |s|) / 2 with the division deferred: sq = s + |s| is exact, sq^3 = 8 max(s,0)^3 with the
//     reference's two roundings, and alpha * sq^3 - tq^3 is ONE rounding in the reference too (alpha is a power of
//     two), so a single FMA gives 8 x the reference value bit for bit; no integer-pipe clamps, no FSEL pairs;
//   * no box mask: a window node outside the reference's box [ibmn, ibmx] is >= 2 cells away from x, so s <= 0 (up
//     to one ulp of u, i.e. a term <= 1e-47 instead of an exact 0) and the formula returns 0 for it.
// SCALE8: return 8 x the values (the caller multiplies its final sum by 8^-ndim, exact) and save one multiply each.
template <bool SCALE8 = false>
__device__ __forceinline__ void spl_window_weights_value(double x, double xmin, double dx, double dxin,
                                                         int nod, int &ws, double b[4]) {
    const double t = spl_mul(dxin, spl_sub(x, xmin));
    const int it = max(__double2int_rz(t), -4);              // Fortran int(): truncation toward zero
    ws = min(max(it - 1, 0), nod - 4);
    const double wsf = (double)ws;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double xb = spl_add(xmin, spl_mul(wsf + (double)k, dx));     // :246 (wsf + k is exact)
        const double u = spl_mul(dxin, spl_sub(x, xb));
        const double s = (k < 2) ? spl_sub(2.0, u) : spl_add(2.0, u);
        const bool edge = (k < 2) ? (ws + k <= 1) : (ws + k >= nod - 2);
        const double sq = spl_add(s, fabs(s));                             // 2 max(s, 0), exact
        const double sq3 = spl_mul(spl_mul(sq, sq), sq);                   // 8 max(s, 0)^3
        const double s1 = spl_sub(s, 1.0);
        const double tq = spl_add(s1, fabs(s1));                           // 2 max(s - 1, 0), exact
        const double tq3 = spl_mul(spl_mul(tq, tq), tq);                   // 8 max(s - 1, 0)^3
        const double alpha = __hiloint2double(edge ? 0x3fe00000 : 0x3fd00000, 0);   // 1/2 : 1/4
        double v = __fma_rn(sq3, alpha, -tq3);                             // product exact: one rounding, as :262/:371
        const double lin = spl_sub(spl_mul(24.0, s), 24.0);                // 8 (3 s - 3), :376
        if (edge && __double2hiint(s) >= 0x40000000) v = lin;              // s >= 2 (s < 0 has the sign bit set)
        b[k] = SCALE8 ? v : spl_mul(v, 0.125);
    }
}

// The same in WORKING PRECISION real32 (the reference built with -DREAL32, src/splpak.F90:33-34, evaluates bascmp in
// float): identical formulas with unfused float operations (__fmul_rn / __fadd_rn), so the 1-D values are bit-identical
// to a float evaluation of bascmp.  Used by the REAL32 library's splfe path: FP32 pipe instead of the FP64 pipe.
__device__ __forceinline__ float spl_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float spl_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float spl_sub(float a, float b) { return __fsub_rn(a, b); }
template <bool SCALE8 = false>
__device__ __forceinline__ void spl_window_weights_value_f32(float x, float xmin, float dx, float dxin, int nod, int &ws,
                                                             float b[4]) {
    const float t = spl_mul(dxin, spl_sub(x, xmin));
    const int it = max(__float2int_rz(t), -4);               // saturating; NaN -> 0
    ws = min(max(it - 1, 0), nod - 4);
    const float wsf = (float)ws;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float xb = spl_add(xmin, spl_mul(wsf + (float)k, dx));
        const float u = spl_mul(dxin, spl_sub(x, xb));
        const float s = (k < 2) ? spl_sub(2.0f, u) : spl_add(2.0f, u);
        const bool edge = (k < 2) ? (ws + k <= 1) : (ws + k >= nod - 2);
        const float sq = spl_add(s, fabsf(s));
        const float sq3 = spl_mul(spl_mul(sq, sq), sq);
        const float s1 = spl_sub(s, 1.0f);
        const float tq = spl_add(s1, fabsf(s1));
        const float tq3 = spl_mul(spl_mul(tq, tq), tq);
        float v = __fmaf_rn(sq3, edge ? 0.5f : 0.25f, -tq3);
        const float lin = spl_sub(spl_mul(24.0f, s), 24.0f);
        if (edge && __float_as_int(s) >= 0x40000000) v = lin;                 // s >= 2
        b[k] = SCALE8 ? v : spl_mul(v, 0.125f);
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

// ------------------------------------------------------------------------------------------
// Uniform ("phantom node") form of the same spline, used by batched evaluation when the extended table fits in
// shared memory.  The reference's natural spline with its linear edge functions (bascmp :302-379) is a plain uniform
// cubic B-spline series on the grid EXTENDED by one phantom node per side:
//     s(x) = sum_{j=-1}^{n} a_j C((x - xmin)/dx - j),     C = the chapeau function of bascmp (:253-270),
//     a_0 = 2 c_0 + 4 c_1,  a_1 = 2 c_1,  a_j = c_j (2 <= j <= n-3),  a_{n-2} = 2 c_{n-2},  a_{n-1} = 2 c_{n-1} + 4 c_{n-2},
//     a_{-1} = 2 a_0 - a_1,  a_n = 2 a_{n-1} - a_{n-2}            (zero second derivative at both ends),
// because L_0 = 2 (C_0 + 2 C_{-1}) and L_1 = 6 C_{-1} + 4 C_0 + 2 C_1 on x >= xmin (mirror image on the right), for every
// n >= 4.  Outside [xmin, xmax] the reference's functions continue linearly, so the four weights there are the
// first-order Taylor expansion of the boundary cell's weights.  With f = t - cell in [0, 1], g = 1 - f the weights of
// the nodes cell-1 .. cell+2 are  C(f+1), C(f), C(f-1), C(f-2) = g^3/4, 1 - 3/2 f^2 + 3/4 f^3, 1 - 3/2 g^2 + 3/4 g^3, f^3/4:
// 9 FP64 operations per dimension instead of the 56 of the node-by-node formulas.  Not bit-identical to bascmp: the
// node positions of the reference carry their own rounding of eps*node index in u (:246), this form does not; the
// difference is bounded by ~nod eps |phi'| (tests/test_gpu_eval.py states the tolerance).
// All weights carry a factor 4 (value), 4*dx (1st) or 4*dx^2 (2nd derivative); the caller scales its final sum.
// ------------------------------------------------------------------------------------------
#define SPL_UNI_OOB 0x40000000u     // flag bit: the coordinate lies outside [xmin, xmax] (linear continuation)

// cell index (0 .. nod-2), fractional coordinate f = t - cell and the out-of-range flag of one dimension
__device__ __forceinline__ void spl_uni_cell(double x, double xmin, double dxin, int nod, int &cell, double &f, bool &oob) {
    const double t = (x - xmin) * dxin;
    const int it = __double2int_rd(t);                        // saturating; NaN -> 0
    oob = (unsigned)it > (unsigned)(nod - 2);
    cell = min(max(it, 0), nod - 2);
    f = t - (double)cell;
}

// weights (x 4 / dxin^nder) of the extended-table entries cell .. cell+3 (nodes cell-1 .. cell+2); nder in 0..2
template <bool VALUE>
__device__ __forceinline__ void spl_uni_weights(double f, bool oob, int nder, double b[4]) {
    const double g = 1.0 - f;
    if (VALUE || nder == 0) {
        const double f2 = f * f, g2 = g * g;
        b[0] = g2 * g;
        b[1] = fma(f2, fma(3.0, f, -6.0), 4.0);
        b[2] = fma(g2, fma(3.0, g, -6.0), 4.0);
        b[3] = f2 * f;
        if (oob) {
            if (f < 0.0) {                                    // left of xmin: value + slope at f = 0
                b[0] = fma(-3.0, f, 1.0); b[1] = 4.0; b[2] = fma(3.0, f, 1.0); b[3] = 0.0;
            } else if (f > 1.0) {                             // right of xmax: value + slope at f = 1
                const double e = f - 1.0;
                b[0] = 0.0; b[1] = fma(-3.0, e, 1.0); b[2] = 4.0; b[3] = fma(3.0, e, 1.0);
            }
        }
    } else if (nder == 1) {
        b[0] = -3.0 * g * g;
        b[1] = f * fma(9.0, f, -12.0);
        b[2] = -g * fma(9.0, g, -12.0);
        b[3] = 3.0 * f * f;
        if (oob) {
            if (f < 0.0) { b[0] = -3.0; b[1] = 0.0; b[2] = 3.0; b[3] = 0.0; }
            else if (f > 1.0) { b[0] = 0.0; b[1] = -3.0; b[2] = 0.0; b[3] = 3.0; }
        }
    } else {
        b[0] = 6.0 * g;
        b[1] = fma(18.0, f, -12.0);
        b[2] = fma(18.0, g, -12.0);
        b[3] = 6.0 * f;
        if (oob && (f < 0.0 || f > 1.0)) { b[0] = 0.0; b[1] = 0.0; b[2] = 0.0; b[3] = 0.0; }
    }
}

// value weights of the uniform form in float (REAL32 library); f is the float64-formed fractional coordinate rounded once
__device__ __forceinline__ void spl_uni_weights_f32(float f, bool oob, float b[4]) {
    const float g = 1.0f - f;
    const float f2 = f * f, g2 = g * g;
    b[0] = g2 * g;
    b[1] = fmaf(f2, fmaf(3.0f, f, -6.0f), 4.0f);
    b[2] = fmaf(g2, fmaf(3.0f, g, -6.0f), 4.0f);
    b[3] = f2 * f;
    if (oob) {
        if (f < 0.0f) {
            b[0] = fmaf(-3.0f, f, 1.0f); b[1] = 4.0f; b[2] = fmaf(3.0f, f, 1.0f); b[3] = 0.0f;
        } else if (f > 1.0f) {
            const float e = f - 1.0f;
            b[0] = 0.0f; b[1] = fmaf(-3.0f, e, 1.0f); b[2] = 4.0f; b[3] = fmaf(3.0f, e, 1.0f);
        }
    }
}
