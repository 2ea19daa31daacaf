// solve.cu -- derivative-constraint rows, band expansion and the blocked FP64 Cholesky solve.
//
// Replaces, for the normal equations G c = g assembled by assemble.cu:
//   * the data-sparse smoothing rows of splcw (src/splpak.F90:862-1048), added to G as rank-1
//     updates (their right-hand side is zero, :866);
//   * the triangular reduction and back-substitution of suprls (:1481-1693).  suprls reduces the
//     dense m x n row stream with Householder/Givens transforms to R with R^T R = G; here G is
//     factored directly, G = L L^T, in LOWER BAND storage (half bandwidth
//     b = 3 * sum_d prod_{d'<d} nodes(d')), which is the dense algorithm when b = n-1.
//
// Band storage: element (i, j), 0 <= i-j < lda, lives at AB[i + j*lda] -- LAPACK band storage with
// ldab = lda + 1 viewed as a dense column-major matrix with leading dimension lda = b + NB, so every
// block kernel below is an ordinary dense column-major kernel on a sub-block (ncol*(lda+1) doubles).
//
// Right-looking blocked Cholesky, panel width NB = 64, two launches per panel:
//   spl_panel_kernel   every CTA factors the NB x NB diagonal block AND inverts the factor in
//                      registers (redundantly: it is a latency chain, and this saves a launch and a
//                      dependency per panel), forward-solves the right-hand side of the block, then
//                      forms its 64 rows of the sub-diagonal panel L21 = A21 L11^-T as a tensor-core
//                      GEMM against the inverse and updates the right-hand side below: the forward
//                      solve L y = g rides along with the factorization.
//   spl_syrk_kernel    trailing update A22 -= L21 L21^T on the lower-triangular 64 x 64 tiles of
//                      the (<= b) x (<= b) window, with FP64 tensor-core MMA
//                      (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4), operands staged k-major in
//                      shared memory.
// Back-substitution L^T c = y runs block by block from the end (spl_backsolve_kernel): every CTA
// forms c_k = L11^-T y_k from the stored block inverse, then eliminates c_k from its slice of the b
// preceding entries.
#include <stdlib.h>

#include "basis.cuh"

#define SOLVE_NB 64

// ------------------------------------------------------------------------------------------
// constraint rows (:862-1048).  One warp per node.
// ------------------------------------------------------------------------------------------
template <int NDIM>
__global__ void __launch_bounds__(128)
spl_constraints_kernel(const __grid_constant__ GridParams gp, double xtrap,
                       const double *__restrict__ cnt, const double *__restrict__ totals_in,
                       double *__restrict__ S, double *__restrict__ totals_out) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const double spcrit = 0.75;                                  // :696
    long long nrect = 1;
#pragma unroll
    for (int d = 0; d < NDIM; ++d) nrect *= (gp.nodes[d] - 1);
    const double wtprrc = __ddiv_rn(totals_in[0], (double)nrect);   // :910
    constexpr int NPAIR = NDIM * (NDIM + 1) / 2;
    constexpr int NCOMBO = spl_ipow(6, NDIM);

    for (long long node = warp_global; node < gp.ncol; node += nwarps) {
        int in[NDIM];
        {
            long long k = node;
#pragma unroll
            for (int d = 0; d < NDIM; ++d) {
                in[d] = (int)(k % gp.nodes[d]);
                k /= gp.nodes[d];
            }
        }
        double expect = wtprrc;
#pragma unroll
        for (int d = 0; d < NDIM; ++d)
            if (in[d] == 0 || in[d] == gp.nodes[d] - 1) expect = spl_mul(0.5, expect);   // :927-929
        const double have = cnt[node];
        if (!(have < spl_mul(spcrit, expect))) continue;                                // :936
        const double dcwght = spl_mul(xtrap, spl_sub(expect, have));                     // :938, :960

        int ibmn[NDIM], nbox[NDIM];
        double xn[NDIM];
#pragma unroll
        for (int d = 0; d < NDIM; ++d) {
            xn[d] = spl_add(gp.xmin[d], spl_mul((double)in[d], gp.dx[d]));               // :943
            int lo = in[d] - 1, hi = in[d] + 1;
            if (in[d] == 0) lo = 0;
            if (in[d] == gp.nodes[d] - 1) hi = gp.nodes[d] - 1;
            ibmn[d] = lo;
            nbox[d] = hi - lo + 1;
        }

        for (int idm = 0; idm < NDIM; ++idm) {
            for (int jdm = idm; jdm < NDIM; ++jdm) {
                int nder[NDIM];
#pragma unroll
                for (int d = 0; d < NDIM; ++d) nder[d] = 0;
                bool boundary = true;
                double rowwt = spl_mul(2.0, dcwght);                                     // :983
                if (jdm == idm) {
                    rowwt = dcwght;
                    nder[jdm] = 2;
                    if (in[idm] != 0 && in[idm] != gp.nodes[idm] - 1) boundary = false;
                }
                if (boundary) {
                    nder[idm] = 1;
                    nder[jdm] = 1;
                }
                // phi[d][k]: 1-D factor of box node k in dimension d (same in every lane)
                double phi[NDIM][3];
#pragma unroll
                for (int d = 0; d < NDIM; ++d)
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        phi[d][k] = (k < nbox[d]) ? spl_bas1(ibmn[d] + k, gp.nodes[d], nder[d], xn[d],
                                                             gp.xmin[d], gp.dx[d], gp.dxin[d])
                                                  : 0.0;
                const double rw2 = rowwt * rowwt;
                // all per-dimension pairs (k <= k') of box nodes: 6 per dimension
                for (int c = lane; c < NCOMBO; c += 32) {
                    int cc = c;
                    double v = rw2;
                    long long nd = 0, nstride = 1;
                    int sten = 0, sstride = 1;
                    bool ok = true;
#pragma unroll
                    for (int d = 0; d < NDIM; ++d) {
                        const int pr = cc % 6;
                        cc /= 6;
                        // pairs of {0,1,2}: (0,0)(0,1)(0,2)(1,1)(1,2)(2,2)
                        const int k0 = (pr >= 3) + (pr >= 5);
                        const int k1 = (pr < 3) ? pr : (pr < 5 ? pr - 2 : 2);
                        if (k1 >= nbox[d]) ok = false;
                        v *= phi[d][k0] * phi[d][k1];
                        nd += (long long)(ibmn[d] + k0) * nstride;
                        sten += (k1 - k0) * sstride;
                        nstride *= gp.nodes[d];
                        sstride *= 4;
                    }
                    if (ok && v != 0.0) atomicAdd(S + nd * gp.nsten + sten, v);
                }
            }
        }
        if (lane == 0) atomicAdd(totals_out + 1, (double)NPAIR);   // constraint rows count as rows
    }
}

int spl_constraints_launch(const GridParams &gp, double xtrap, const double *d_cnt,
                           const double *d_totals_in, double *d_S, double *d_totals_out,
                           cudaStream_t st, int nsm) {
    long long blocks = (gp.ncol + 3) / 4;
    if (blocks > (long long)nsm * 16) blocks = (long long)nsm * 16;
    switch (gp.ndim) {
    case 1: spl_constraints_kernel<1><<<(unsigned)blocks, 128, 0, st>>>(gp, xtrap, d_cnt, d_totals_in, d_S, d_totals_out); break;
    case 2: spl_constraints_kernel<2><<<(unsigned)blocks, 128, 0, st>>>(gp, xtrap, d_cnt, d_totals_in, d_S, d_totals_out); break;
    case 3: spl_constraints_kernel<3><<<(unsigned)blocks, 128, 0, st>>>(gp, xtrap, d_cnt, d_totals_in, d_S, d_totals_out); break;
    case 4: spl_constraints_kernel<4><<<(unsigned)blocks, 128, 0, st>>>(gp, xtrap, d_cnt, d_totals_in, d_S, d_totals_out); break;
    default: return SPLPAK_ERR_NDIM;
    }
    ++g_spl_launches;
    SPL_CUDA_TRY(cudaGetLastError());
    return SPLPAK_OK;
}

// ------------------------------------------------------------------------------------------
// S (orthant stencil) -> lower band storage
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
spl_expand_band_kernel(const __grid_constant__ GridParams gp, const double *__restrict__ S,
                       double *__restrict__ AB, long long lda, int bw) {
    // one thread per (column j, band offset off in 0..bw)
    const long long total = gp.ncol * (long long)(bw + 1);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const long long j = e / (bw + 1);
        const int off = (int)(e - j * (bw + 1));
        const long long i = j + off;
        if (i >= gp.ncol) continue;
        long long ki = i, kj = j, node = 0, nstride = 1;
        int sten = 0, sstride = 1;
        bool inside = true;
        for (int d = 0; d < gp.ndim; ++d) {
            const int id = (int)(ki % gp.nodes[d]);
            const int jd = (int)(kj % gp.nodes[d]);
            ki /= gp.nodes[d];
            kj /= gp.nodes[d];
            const int del = id > jd ? id - jd : jd - id;
            if (del > 3) inside = false;
            node += (long long)(id < jd ? id : jd) * nstride;
            sten += del * sstride;
            nstride *= gp.nodes[d];
            sstride *= 4;
        }
        if (inside) AB[i + j * lda] = S[node * gp.nsten + sten];
    }
}

// ------------------------------------------------------------------------------------------
// panel: diagonal-block Cholesky + inverse (registers), rhs forward solve, TRSM as a DMMA GEMM
// ------------------------------------------------------------------------------------------
#define PANEL_THREADS 256
#define TILE_LD 68      // 64 + 4: (t4*68 + g) hits 16 distinct 8-byte banks per half-warp (conflict-free LDS.64)

__device__ __forceinline__ void spl_dmma_8x8x4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void spl_cp_async8(double *smem_dst, const double *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void spl_cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// Every CTA factors the nb x nb diagonal block A11 = L11 L11^T and inverts L11 in registers (thread
// (ti, tj) of a 16 x 16 grid owns a 4 x 4 sub-block of A and of X = L11^-1; one barrier per pivot,
// pivot column / row broadcast through double-buffered shared memory).  Doing this redundantly in
// every CTA costs no time (it is a latency chain) and saves a launch + dependency per panel.
// Then  y1 = L11^-1 g1,  L21 = A21 L11^-T  for this CTA's 64 rows (FP64 tensor-core MMA against the
// inverse, so no per-row substitution chain) and  g2 -= L21 y1.
// CTA 0 stores L11^-1 (for the back-substitution) and y1.
__global__ void __launch_bounds__(PANEL_THREADS)
spl_panel_kernel(double *__restrict__ AB, long long lda, long long j0, int nb, int m,
                 double *__restrict__ g, double *__restrict__ ysol, double *__restrict__ linv_blk,
                 int *__restrict__ fail) {
    extern __shared__ __align__(16) double s_pan[];
    double *sA = s_pan;                          // [k][row]  A21 tile, 64 x TILE_LD
    double *sB = s_pan + 64 * TILE_LD;           // [k][n]    L11^-1 [n][k]
    double *s_col = sB + 64 * TILE_LD;           // 2 x 64   pivot column (double buffered)
    double *s_row = s_col + 128;                 // 2 x 64   pivot row of X
    double *s_g = s_row + 128;                   // 64       g1, then y1
    const int tid = threadIdx.x;
    if (*fail) return;                           // an earlier panel failed (uniform across the grid)
    const int R0 = blockIdx.x * 64;              // first row of this CTA's tile, relative to j0 + nb
    const long long r0 = j0 + nb;

    // ---- start streaming this CTA's A21 tile into shared memory; it lands under the factorization ----
    for (int e = tid; e < 64 * 64; e += PANEL_THREADS) {
        const int k = e >> 6, r = e & 63;
        double *dst = sA + k * TILE_LD + r;
        if (k < nb && R0 + r < m) spl_cp_async8(dst, AB + (r0 + R0 + r) + (j0 + k) * lda);
        else *dst = 0.0;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");

    // ---- load A11 (4 x 4 per thread) ----
    const int ti = tid >> 4, tj = tid & 15;
    double A[4][4], X[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int i = 4 * ti + a, j = 4 * tj + b;
            double v = (i == j) ? 1.0 : 0.0;                 // identity padding when nb < 64
            if (i < nb && j < nb && i >= j) v = AB[(j0 + i) + (j0 + j) * lda];
            A[a][b] = v;
            X[a][b] = (i == j) ? 1.0 : 0.0;
        }
    if (tid < 64) s_g[tid] = (tid < nb) ? g[j0 + tid] : 0.0;

    bool bad = false;
#pragma unroll 1
    for (int kb = 0; kb < 16; ++kb) {
        const bool rows_live = (ti >= kb);      // this thread still owns rows >= the pivot block
        const bool cols_live = (tj >= kb);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int k = 4 * kb + kk;
            double *col = s_col + (k & 1) * 64;
            double *row = s_row + (k & 1) * 64;
            if (tj == kb) {
                *reinterpret_cast<double2 *>(col + 4 * ti) = make_double2(A[0][kk], A[1][kk]);
                *reinterpret_cast<double2 *>(col + 4 * ti + 2) = make_double2(A[2][kk], A[3][kk]);
            }
            if (ti == kb) {
                *reinterpret_cast<double2 *>(row + 4 * tj) = make_double2(X[kk][0], X[kk][1]);
                *reinterpret_cast<double2 *>(row + 4 * tj + 2) = make_double2(X[kk][2], X[kk][3]);
            }
            __syncthreads();
            const double d = col[k];
            if (!(d > 0.0)) bad = true;                       // non-positive (or NaN) pivot -> 107
            // 1/sqrt(d): single-precision seed + two Newton steps in double (full precision, and a
            // much shorter dependent chain than the library rsqrt on this latency-bound path)
            double rinv;
            {
                const float df = (float)d;
                if (df > 1e-30f && df < 1e30f) {
                    double r = (double)rsqrtf(df);
                    const double hd = 0.5 * d;
                    r = r * fma(-hd * r, r, 1.5);
                    r = r * fma(-hd * r, r, 1.5);
                    rinv = r;
                } else {
                    rinv = bad ? 0.0 : rsqrt(d);
                }
            }
            if (rows_live) {
                const double2 c01 = *reinterpret_cast<const double2 *>(col + 4 * ti);
                const double2 c23 = *reinterpret_cast<const double2 *>(col + 4 * ti + 2);
                const double2 r01 = *reinterpret_cast<const double2 *>(row + 4 * tj);
                const double2 r23 = *reinterpret_cast<const double2 *>(row + 4 * tj + 2);
                const double li[4] = {c01.x * rinv, c01.y * rinv, c23.x * rinv, c23.y * rinv};
                const double xr[4] = {r01.x * rinv, r01.y * rinv, r23.x * rinv, r23.y * rinv};
                if (tj == kb) {
#pragma unroll
                    for (int a = 0; a < 4; ++a) A[a][kk] = li[a];          // column k of L11 (rows >= k valid)
                }
                if (ti == kb) {
#pragma unroll
                    for (int b = 0; b < 4; ++b) X[kk][b] = xr[b];          // row k of L11^-1
                }
                double lj[4] = {0.0, 0.0, 0.0, 0.0};
                if (cols_live) {
                    const double2 d01 = *reinterpret_cast<const double2 *>(col + 4 * tj);
                    const double2 d23 = *reinterpret_cast<const double2 *>(col + 4 * tj + 2);
                    lj[0] = d01.x * rinv;
                    lj[1] = d01.y * rinv;
                    lj[2] = d23.x * rinv;
                    lj[3] = d23.y * rinv;
                }
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const bool row_gt = (ti > kb) || (a > kk);             // global row > k (ti >= kb here)
                    if (row_gt) {
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const bool col_gt = (tj > kb) || (tj == kb && b > kk);
                            if (col_gt) A[a][b] = fma(-li[a], lj[b], A[a][b]);
                            X[a][b] = fma(-li[a], xr[b], X[a][b]);
                        }
                    }
                }
            }
        }
    }
    if (bad) {
        if (blockIdx.x == 0 && tid == 0) *fail = 1;
        spl_cp_async_wait_all();
        return;
    }

    // ---- stage L11^-1 as the B operand: sB[k][n] = Linv[n][k];  CTA 0 also stores it for the backsolve ----
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int n = 4 * ti + a, k = 4 * tj + b;
            const double v = (n >= k) ? X[a][b] : 0.0;
            sB[k * TILE_LD + n] = v;
            if (blockIdx.x == 0) linv_blk[n * 64 + k] = v;
        }
    spl_cp_async_wait_all();
    __syncthreads();
    // y1[c] = sum_k Linv[c][k] g1[k]
    double y1c = 0.0;
    if (tid < 64) {
#pragma unroll 8
        for (int k = 0; k < 64; ++k) y1c = fma(sB[k * TILE_LD + tid], s_g[k], y1c);
    }
    __syncthreads();
    if (tid < 64) {
        s_g[tid] = y1c;
        if (blockIdx.x == 0 && tid < nb) ysol[j0 + tid] = y1c;
    }
    __syncthreads();
    if (m <= 0) return;

    // ---- L21 tile = A21 tile * Linv^T: out[r][c] = sum_k A21[r][k] Linv[c][k] ----
    const int warp = tid >> 5, lane = tid & 31;
    const int gq = lane >> 2, t4 = lane & 3;
    const int wy = (warp >> 1) * 16, wx = (warp & 1) * 32;
    double acc[2][4][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
#pragma unroll 4
    for (int k0 = 0; k0 < 64; k0 += 4) {
        double af[2], bf[4];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) af[mi] = sA[(k0 + t4) * TILE_LD + wy + mi * 8 + gq];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) bf[ni] = sB[(k0 + t4) * TILE_LD + wx + ni * 8 + gq];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) spl_dmma_8x8x4(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
    }
    // store L21 and update the right-hand side below the block: g2 -= L21 y1
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
        const int r = R0 + wy + mi * 8 + gq;
        double part = 0.0;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = wx + ni * 8 + 2 * t4 + h;
                const double v = acc[mi][ni][h];
                if (r < m && c < nb) AB[(r0 + r) + (j0 + c) * lda] = v;
                part = fma(v, s_g[c], part);
            }
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        if (t4 == 0 && r < m && part != 0.0) atomicAdd(g + r0 + r, -part);
    }
}

// ------------------------------------------------------------------------------------------
// trailing update with FP64 tensor-core MMA
// ------------------------------------------------------------------------------------------
#define SYRK_TILE 64
#define SYRK_THREADS 128

// C[I,J] -= P[I,:] P[J,:]^T over the lower-triangular tiles of the m x m window whose first
// row/column is global index r0; P = panel rows r0.., columns j0..j0+nb-1.  Operand tiles are
// streamed with cp.async (k-major, conflict-free fragment reads); the C tile is prefetched into the
// accumulators while the operands are in flight, so the kernel is one global round trip long.
__global__ void __launch_bounds__(SYRK_THREADS)
spl_syrk_kernel(double *__restrict__ AB, long long lda, long long r0, long long j0, int nb, int m,
                const int *__restrict__ fail) {
    extern __shared__ __align__(16) double s_ab[];
    double *sA = s_ab;                          // [k][row]
    double *sB = s_ab + 64 * TILE_LD;
    if (*fail) return;
    const int tile = blockIdx.x;
    int ti = (int)((sqrt(8.0 * (double)tile + 1.0) - 1.0) * 0.5);
    while ((long long)(ti + 1) * (ti + 2) / 2 <= tile) ++ti;
    while ((long long)ti * (ti + 1) / 2 > tile) --ti;
    const int tj = tile - ti * (ti + 1) / 2;
    const int I0 = ti * SYRK_TILE, J0 = tj * SYRK_TILE;
    const int t = threadIdx.x;
    const bool diag = (ti == tj);

    for (int e = t; e < 64 * 64; e += SYRK_THREADS) {
        const int k = e >> 6, r = e & 63;
        double *da = sA + k * TILE_LD + r;
        if (k < nb && I0 + r < m) spl_cp_async8(da, AB + (r0 + I0 + r) + (j0 + k) * lda);
        else *da = 0.0;
        if (!diag) {
            double *db = sB + k * TILE_LD + r;
            if (k < nb && J0 + r < m) spl_cp_async8(db, AB + (r0 + J0 + r) + (j0 + k) * lda);
            else *db = 0.0;
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");

    const int warp = t >> 5, lane = t & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int wy = (warp >> 1) * 32, wx = (warp & 1) * 32;
    // prefetch the C tile into the accumulators (lower triangle, inside the window)
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int li = I0 + wy + mi * 8 + g;
                const int lj = J0 + wx + ni * 8 + 2 * t4 + h;
                acc[mi][ni][h] = (li < m && lj < m && li >= lj) ? AB[(r0 + li) + (r0 + lj) * lda] : 0.0;
            }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const double *pB = diag ? sA : sB;

#pragma unroll 4
    for (int k0 = 0; k0 < 64; k0 += 4) {
        double a[4], b[4];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) a[mi] = -sA[(k0 + t4) * TILE_LD + wy + mi * 8 + g];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) b[ni] = pB[(k0 + t4) * TILE_LD + wx + ni * 8 + g];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) spl_dmma_8x8x4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int li = I0 + wy + mi * 8 + g;
                const int lj = J0 + wx + ni * 8 + 2 * t4 + h;
                if (li < m && lj < m && li >= lj) AB[(r0 + li) + (r0 + lj) * lda] = acc[mi][ni][h];
            }
}

// ------------------------------------------------------------------------------------------
// back-substitution L^T c = y, block by block from the end, using the stored block inverses
// ------------------------------------------------------------------------------------------
#define BACK_THREADS 128
__global__ void __launch_bounds__(BACK_THREADS)
spl_backsolve_kernel(const double *__restrict__ AB, long long lda, long long j0, int nb, int bw,
                     const double *__restrict__ linv_blk, double *__restrict__ ysol,
                     double *__restrict__ csol, const int *__restrict__ fail) {
    __shared__ double s_li[64 * 65];
    __shared__ double s_y[64];
    __shared__ double s_c[64];
    const int t = threadIdx.x;
    if (*fail) return;
    // the column this thread eliminates from, prefetched while the block solve runs
    const long long jlo = (j0 - bw > 0) ? j0 - bw : 0;
    const long long j = jlo + (long long)blockIdx.x * BACK_THREADS + t;
    double colv[64];
    if (j < j0) {
        const double *col = AB + j0 + j * lda;     // rows j0.. of column j, contiguous
#pragma unroll
        for (int i = 0; i < 64; ++i) colv[i] = (i < nb) ? col[i] : 0.0;
    }
    // stage L11^-1 (row-major [r][c]) and y_k
#pragma unroll
    for (int e = 0; e < 32; ++e) {
        const int idx = t + e * BACK_THREADS;
        s_li[(idx >> 6) * 65 + (idx & 63)] = linv_blk[idx];
    }
    if (t < 64) s_y[t] = (t < nb) ? ysol[j0 + t] : 0.0;
    __syncthreads();
    // c_k = L11^-T y_k:  c[i] = sum_{r >= i} Linv[r][i] y[r]
    if (t < 64) {
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int r = 0; r < 64; r += 2) {
            c0 = fma(s_li[r * 65 + t], s_y[r], c0);
            c1 = fma(s_li[(r + 1) * 65 + t], s_y[r + 1], c1);
        }
        const double c = c0 + c1;
        s_c[t] = c;
        if (blockIdx.x == 0 && t < nb) csol[j0 + t] = c;
    }
    __syncthreads();
    // eliminate c_k from the bw preceding unknowns: y[j] -= sum_i L[i][j] c[i], i in the block
    if (j < j0) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
            s0 = fma(colv[i], s_c[i], s0);
            s1 = fma(colv[i + 1], s_c[i + 1], s1);
        }
        ysol[j] -= (s0 + s1);
    }
}

// ------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------
long long spl_band_lda(int bw) { return (long long)bw + SOLVE_NB; }

int spl_half_bandwidth(const GridParams &gp) {
    long long b = 0, stride = 1;
    for (int d = 0; d < gp.ndim; ++d) {
        b += 3 * stride;
        stride *= gp.nodes[d];
    }
    if (b > gp.ncol - 1) b = gp.ncol - 1;
    return (int)b;
}

// Workspace (doubles) next to the band matrix: block inverses (64*64 per panel) + ysol (n) + csol (n).
long long spl_solve_workspace(const GridParams &gp) {
    const long long nblk = (gp.ncol + SOLVE_NB - 1) / SOLVE_NB;
    return nblk * 64 * 64 + 2 * (gp.ncol + 64);
}

// AB must hold ncol*(lda+1) doubles and be zero-filled on entry.  g enters as the right-hand side
// (destroyed); the solution is left in d_work + nblk*4096 + (ncol+64)  (returned through *d_coef_out).
int spl_solve_launch(const GridParams &gp, const double *d_S, double *d_AB, double *d_g, double *d_work,
                     double **d_coef_out, int *d_fail, cudaStream_t st, int nsm, cudaEvent_t *ev) {
    const long long n = gp.ncol;
    const int bw = spl_half_bandwidth(gp);
    const long long lda = spl_band_lda(bw);
    const long long nblk = (n + SOLVE_NB - 1) / SOLVE_NB;
    double *d_linv = d_work;
    double *d_ysol = d_work + nblk * 4096;
    double *d_csol = d_ysol + (n + 64);
    *d_coef_out = d_csol;
    (void)nsm;
    if (ev) cudaEventRecord(ev[0], st);
    {
        const long long total = n * (long long)(bw + 1);
        long long blocks = (total + 255) / 256;
        if (blocks > 148LL * 32) blocks = 148LL * 32;
        spl_expand_band_kernel<<<(unsigned)blocks, 256, 0, st>>>(gp, d_S, d_AB, lda, bw);
        ++g_spl_launches;
    }
    if (ev) cudaEventRecord(ev[1], st);
    const size_t syrk_smem = sizeof(double) * 2 * 64 * TILE_LD;
    const size_t panel_smem = sizeof(double) * (2 * 64 * TILE_LD + 128 + 128 + 64);
    SPL_CUDA_TRY(cudaFuncSetAttribute(spl_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)syrk_smem));
    SPL_CUDA_TRY(cudaFuncSetAttribute(spl_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)panel_smem));
    for (long long kb = 0; kb < nblk; ++kb) {
        const long long j0 = kb * SOLVE_NB;
        const int nb = (int)((n - j0 < SOLVE_NB) ? n - j0 : SOLVE_NB);
        const long long r0 = j0 + nb;
        long long mm = n - r0;
        if (mm > bw) mm = bw;
        const int m = (int)mm;
        int pblocks = (m + 63) / 64;
        if (pblocks < 1) pblocks = 1;
        spl_panel_kernel<<<pblocks, PANEL_THREADS, panel_smem, st>>>(d_AB, lda, j0, nb, m, d_g, d_ysol,
                                                                     d_linv + kb * 4096, d_fail);
        ++g_spl_launches;
        if (m > 0) {
            const int T = (m + SYRK_TILE - 1) / SYRK_TILE;
            const int tiles = T * (T + 1) / 2;
            spl_syrk_kernel<<<tiles, SYRK_THREADS, syrk_smem, st>>>(d_AB, lda, r0, j0, nb, m, d_fail);
            ++g_spl_launches;
        }
        if (getenv("SPLPAK_B200_DEBUG")) {
            int f = 0;
            cudaError_t e = cudaStreamSynchronize(st);
            cudaMemcpy(&f, d_fail, sizeof(int), cudaMemcpyDeviceToHost);
            double dg[4] = {0, 0, 0, 0};
            const long long jn = (j0 + nb < n) ? j0 + nb : j0;
            cudaMemcpy(dg, d_AB + jn + jn * lda, sizeof(double), cudaMemcpyDeviceToHost);
            cudaMemcpy(dg + 1, d_AB + (j0 + nb - 1) + (j0) * lda, sizeof(double), cudaMemcpyDeviceToHost);
            // host Cholesky of the NEXT diagonal block (as the next panel will see it)
            double minpiv = 0.0;
            int badk = -1;
            if (j0 + nb < n) {
                const long long jj = j0 + nb;
                const int nn2 = (int)((n - jj < 64) ? n - jj : 64);
                static double blk[64 * 64];
                for (int c = 0; c < nn2; ++c)
                    cudaMemcpy(blk + c * 64, d_AB + (jj) + (jj + c) * lda, sizeof(double) * nn2, cudaMemcpyDeviceToHost);
                // blk[c*64 + r] = A[jj + r][jj + c] for r >= c (column c starts at row jj, so shift)
                double Lh[64][64];
                for (int r = 0; r < nn2; ++r) for (int c = 0; c <= r; ++c) Lh[r][c] = blk[c * 64 + r];
                minpiv = 1e300;
                for (int k2 = 0; k2 < nn2 && badk < 0; ++k2) {
                    double d2 = Lh[k2][k2];
                    for (int c = 0; c < k2; ++c) d2 -= Lh[k2][c] * Lh[k2][c];
                    if (d2 < minpiv) minpiv = d2;
                    if (!(d2 > 0)) { badk = k2; break; }
                    double pv = sqrt(d2);
                    Lh[k2][k2] = pv;
                    for (int r = k2 + 1; r < nn2; ++r) {
                        double v = Lh[r][k2];
                        for (int c = 0; c < k2; ++c) v -= Lh[r][c] * Lh[k2][c];
                        Lh[r][k2] = v / pv;
                    }
                }
            }
            fprintf(stderr, "panel kb=%lld j0=%lld nb=%d m=%d fail=%d err=%s nextdiag=%g lastrow0=%g next-block host chol: minpiv=%g badk=%d\n",
                    kb, j0, nb, m, f, cudaGetErrorString(e), dg[0], dg[1], minpiv, badk);
            if (f) break;
        }
    }
    if (ev) cudaEventRecord(ev[2], st);
    for (long long kb = nblk - 1; kb >= 0; --kb) {
        const long long j0 = kb * SOLVE_NB;
        const int nb = (int)((n - j0 < SOLVE_NB) ? n - j0 : SOLVE_NB);
        const long long span = (j0 < bw) ? j0 : bw;
        int blocks = (int)((span + BACK_THREADS - 1) / BACK_THREADS);
        if (blocks < 1) blocks = 1;
        spl_backsolve_kernel<<<blocks, BACK_THREADS, 0, st>>>(d_AB, lda, j0, nb, bw, d_linv + kb * 4096, d_ysol,
                                                              d_csol, d_fail);
        ++g_spl_launches;
    }
    if (ev) cudaEventRecord(ev[3], st);
    SPL_CUDA_TRY(cudaGetLastError());
    return SPLPAK_OK;
}
