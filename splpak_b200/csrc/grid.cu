// grid.cu -- evaluation of the spline (or a partial derivative) on a REGULAR OUTPUT GRID for sm_100a.
//
// New entry point (the main upstream use of splpak is fit-then-grid, README.md:60):
// the reference would call splfe/splde (src/splpak.F90:1089-1275) once per grid point.  On a tensor grid
// x(i1, .., iN) = (a1[i1], .., aN[iN]) the spline separates,
//     s = sum_{k1..kN} coef[k1..kN] * B1[i1][k1] * ... * BN[iN][kN],
// with B_d[i][k] = the 1-D basis (or derivative) of node k at a_d[i]: four non-zeros per row, the same
// values and the same 4-wide window as the point-wise path (basis.cuh).  So instead of one 4^N gather +
// contraction per output point, the table is contracted one dimension at a time ("mode products"):
//     T1[kN..k2][i1]      = sum_a coef[kN..k2][s1(i1)+a] * w1[i1][a]
//     T2[kN..k3][i2][i1]  = sum_a T1[kN..k3][s2(i2)+a][i1] * w2[i2][a]
//     ...
//     out[iN..i1]         = sum_a T(N-1)[sN(iN)+a][i(N-1)..i1] * wN[iN][a]
// Every stage is 4 FMAs per element it WRITES; the last stage writes the output, reads coalesced rows of a
// tensor nodes(N)/4 times smaller per output plane (L2 resident) -- so the whole evaluation is bound by
// streaming the output to HBM (8 bytes per point) instead of by the shared-memory gather.
#include "basis.cuh"

// per-axis window start and the four weights of every axis point
__global__ void __launch_bounds__(256)
spl_axis_weights_kernel(const real_t *__restrict__ axis, long long n, double xmin, double dx, double dxin, int nod,
                        int nder, int *__restrict__ ws, double *__restrict__ w4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double b[4];
    int s;
    const double x = (double)axis[i];
    if (nder == 0) {
        spl_window_weights_value(x, xmin, dx, dxin, nod, s, b);
        if (x != x) b[0] = b[1] = b[2] = b[3] = 0.0;       // NaN fails every comparison of bascmp: all terms 0
    } else {
        spl_window_weights(x, xmin, dx, dxin, nod, nder, s, b);
    }
    ws[i] = s;
#pragma unroll
    for (int a = 0; a < 4; ++a) w4[4 * i + a] = b[a];
}

// out[o][i][in] = sum_a A[o][ws[i] + a][in] * w4[i][a]     A: (outer, nin, inner), out: (outer, nout, inner)
template <typename OutT>
__global__ void __launch_bounds__(256)
spl_mode_product_kernel(const double *__restrict__ A, long long outer, int nin, long long nout, long long inner,
                        const int *__restrict__ ws, const double *__restrict__ w4, OutT *__restrict__ out) {
    const long long total = outer * nout * inner;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const long long in = e % inner;
        const long long t = e / inner;
        const long long i = t % nout;
        const long long o = t / nout;
        const double *p = A + (o * nin + ws[i]) * inner + in;
        const double *w = w4 + 4 * i;
        double s = p[0] * w[0];
        s = fma(p[inner], w[1], s);
        s = fma(p[2 * inner], w[2], s);
        s = fma(p[3 * inner], w[3], s);
        out[e] = (OutT)s;
    }
}

// Same contraction for the stages with a long contiguous inner extent (every stage but the first): a thread
// owns one inner position and walks ICH consecutive output indices i, keeping the four input values of the
// current window in registers and reloading them only when the window start changes.  On a sorted axis that is
// once per node interval, so the input tensor is read ~once and the stage streams its output at HBM speed
// instead of fetching 4 values from L2 per output.  No integer division per element: blockIdx.y = (o, i-chunk).
#define GRID_ICH 32
template <typename OutT>
__global__ void __launch_bounds__(256)
spl_mode_product_rows_kernel(const double *__restrict__ A, int nin, long long nout, long long inner,
                             const int *__restrict__ ws, const double *__restrict__ w4, OutT *__restrict__ out,
                             long long nchunk) {
    const long long o = blockIdx.y / nchunk;
    const long long i0 = (blockIdx.y % nchunk) * GRID_ICH;
    const long long in = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (in >= inner) return;
    const double *base = A + o * nin * inner + in;
    OutT *dst = out + o * nout * inner + in;
    double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
    int cur = -1;
    const long long i1 = (i0 + GRID_ICH < nout) ? i0 + GRID_ICH : nout;
    for (long long i = i0; i < i1; ++i) {
        const int s = ws[i];                                   // uniform across the block: broadcast loads
        if (s != cur) {
            const double *p = base + (long long)s * inner;
            p0 = p[0];
            p1 = p[inner];
            p2 = p[2 * inner];
            p3 = p[3 * inner];
            cur = s;
        }
        const double *w = w4 + 4 * i;
        double v = p0 * w[0];
        v = fma(p1, w[1], v);
        v = fma(p2, w[2], v);
        v = fma(p3, w[3], v);
        dst[i * inner] = (OutT)v;
    }
}

// d_axis[d]: device pointer of axis d (naxis[d] points); d_coef64: float64 table; d_out: prod(naxis)
// values, dimension 1 fastest.  d_tmp: 2 * tmp_elems doubles of scratch (tmp_elems from spl_grid_tmp_elems),
// d_iws / d_w4: sum(naxis) ints / 4*sum(naxis) doubles.
long long spl_grid_tmp_elems(const GridParams &gp, const long long *naxis) {
    long long mx = 0, inner = 1;
    for (int d = 0; d + 1 < gp.ndim; ++d) {
        inner *= naxis[d];
        long long outer = 1;
        for (int e = d + 1; e < gp.ndim; ++e) outer *= gp.nodes[e];
        if (outer * inner > mx) mx = outer * inner;
    }
    return mx;
}

int spl_eval_grid_launch(const GridParams &gp, const int *nderiv, const real_t *const *d_axis, const long long *naxis,
                         const double *d_coef64, real_t *d_out, double *d_tmp, long long tmp_elems, int *d_iws,
                         double *d_w4, cudaStream_t st, int nsm) {
    long long off = 0;
    for (int d = 0; d < gp.ndim; ++d) {
        const long long n = naxis[d];
        if (n <= 0) return SPLPAK_OK;
        spl_axis_weights_kernel<<<spl_div_up(n, 256), 256, 0, st>>>(d_axis[d], n, gp.xmin[d], gp.dx[d], gp.dxin[d],
                                                                     gp.nodes[d], nderiv ? nderiv[d] : 0, d_iws + off,
                                                                     d_w4 + 4 * off);
        ++g_spl_launches;
        off += n;
    }
    const double *src = d_coef64;
    long long inner = 1;
    off = 0;
    for (int d = 0; d < gp.ndim; ++d) {
        long long outer = 1;
        for (int e = d + 1; e < gp.ndim; ++e) outer *= gp.nodes[e];
        const long long total = outer * naxis[d] * inner;
        long long blocks = (total + 255) / 256;
        const long long cap = (long long)nsm * 32;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        const long long nchunk = (naxis[d] + GRID_ICH - 1) / GRID_ICH;
        const bool rows = inner >= 256 && outer * nchunk <= 65535;
        const dim3 rgrid((unsigned)((inner + 255) / 256), (unsigned)(rows ? outer * nchunk : 1));
        if (d == gp.ndim - 1) {
            if (rows)
                spl_mode_product_rows_kernel<real_t><<<rgrid, 256, 0, st>>>(src, gp.nodes[d], naxis[d], inner, d_iws + off,
                                                                           d_w4 + 4 * off, d_out, nchunk);
            else
                spl_mode_product_kernel<real_t><<<(unsigned)blocks, 256, 0, st>>>(src, outer, gp.nodes[d], naxis[d], inner,
                                                                                 d_iws + off, d_w4 + 4 * off, d_out);
        } else {
            double *dst = d_tmp + (d & 1) * tmp_elems;
            if (rows)
                spl_mode_product_rows_kernel<double><<<rgrid, 256, 0, st>>>(src, gp.nodes[d], naxis[d], inner, d_iws + off,
                                                                           d_w4 + 4 * off, dst, nchunk);
            else
                spl_mode_product_kernel<double><<<(unsigned)blocks, 256, 0, st>>>(src, outer, gp.nodes[d], naxis[d], inner,
                                                                                 d_iws + off, d_w4 + 4 * off, dst);
            src = dst;
        }
        ++g_spl_launches;
        inner *= naxis[d];
        off += naxis[d];
    }
    SPL_CUDA_TRY(cudaGetLastError());
    return SPLPAK_OK;
}
