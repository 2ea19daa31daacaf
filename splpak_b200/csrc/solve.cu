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
// Band storage: element (i, j), 0 <= i-j <= lda, lives at AB[i + j*lda] -- a dense column-major
// matrix with leading dimension lda = b + NB, so every block kernel below is an ordinary dense
// column-major kernel on a sub-block.
//
// Right-looking blocked Cholesky, panel width NB = 64, two launches per panel:
//   spl_panel_kernel   every CTA factors the NB x NB diagonal block in shared memory (redundantly,
//                      which saves a launch and a dependency per panel), forward-substitutes the
//                      right-hand side of the block, then solves its 128 rows of the sub-diagonal
//                      panel L21 = A21 L11^-T and updates the right-hand side below: the forward
//                      solve L y = g rides along with the factorization.
//   spl_syrk_kernel    trailing update A22 -= L21 L21^T on the lower-triangular 64 x 64 tiles of
//                      the (<= b) x (<= b) window, with FP64 tensor-core MMA
//                      (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4), operands staged k-major in
//                      shared memory.
// Back-substitution L^T c = y runs block by block from the end (spl_backsolve_kernel): every CTA
// solves the diagonal block for c_k, then eliminates c_k from its slice of the b preceding entries.
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
// panel: diagonal-block Cholesky + forward substitution + TRSM of the sub-diagonal panel
// ------------------------------------------------------------------------------------------
#define PANEL_THREADS 128
#define PANEL_LD (SOLVE_NB + 1)

// Factor the nb x nb lower block held in s_L (row-major, leading dim PANEL_LD) in place.
// Left-looking by columns; all PANEL_THREADS threads cooperate.  Returns false on a bad pivot.
__device__ __forceinline__ bool spl_block_cholesky(double *s_L, int nb, int *s_flag) {
    const int t = threadIdx.x;
    for (int k = 0; k < nb; ++k) {
        // column k: rows i >= k get  A[i][k] - sum_{c<k} L[i][c] L[k][c]
        if (t >= k && t < nb) {
            double s0 = 0.0, s1 = 0.0;
            int c = 0;
            for (; c + 1 < k; c += 2) {
                s0 = fma(s_L[t * PANEL_LD + c], s_L[k * PANEL_LD + c], s0);
                s1 = fma(s_L[t * PANEL_LD + c + 1], s_L[k * PANEL_LD + c + 1], s1);
            }
            if (c < k) s0 = fma(s_L[t * PANEL_LD + c], s_L[k * PANEL_LD + c], s0);
            s_L[t * PANEL_LD + k] -= (s0 + s1);
        }
        __syncthreads();
        const double d = s_L[k * PANEL_LD + k];
        if (!(d > 0.0)) {          // non-positive (or NaN) pivot -> solver failure (107)
            if (t == 0) *s_flag = 1;
            __syncthreads();
            return false;
        }
        const double piv = sqrt(d);
        __syncthreads();
        if (t == k) s_L[k * PANEL_LD + k] = piv;
        else if (t > k && t < nb) s_L[t * PANEL_LD + k] = s_L[t * PANEL_LD + k] / piv;
        __syncthreads();
    }
    return true;
}

__global__ void __launch_bounds__(PANEL_THREADS)
spl_panel_kernel(double *__restrict__ AB, long long lda, long long n, long long j0, int nb, int m,
                 double *__restrict__ y, int *__restrict__ fail) {
    __shared__ double s_L[SOLVE_NB * PANEL_LD];
    __shared__ double s_y[SOLVE_NB];
    __shared__ double s_rd[SOLVE_NB];
    __shared__ int s_flag;
    const int t = threadIdx.x;
    if (t == 0) s_flag = 0;
    if (*fail) return;   // an earlier panel already failed (uniform across the grid)
    // load the diagonal block (lower triangle; upper part zeroed)
    for (int e = t; e < nb * nb; e += PANEL_THREADS) {
        const int c = e / nb, r = e - c * nb;     // column-major walk: coalesced over r
        s_L[r * PANEL_LD + c] = (r >= c) ? AB[(j0 + r) + (j0 + c) * lda] : 0.0;
    }
    if (t < nb) s_y[t] = y[j0 + t];
    __syncthreads();
    if (!spl_block_cholesky(s_L, nb, &s_flag)) {
        if (blockIdx.x == 0 && t == 0) *fail = 1;
        return;
    }
    // forward substitution on the block: y1 = L11^-1 g1 (warp 0, sequential over columns)
    if (t < 32) {
        for (int k = 0; k < nb; ++k) {
            const double yk = s_y[k] / s_L[k * PANEL_LD + k];
            __syncwarp();
            if (t == 0) s_y[k] = yk;
            for (int i = k + 1 + t; i < nb; i += 32) s_y[i] = fma(-s_L[i * PANEL_LD + k], yk, s_y[i]);
            __syncwarp();
        }
    }
    if (t < nb) s_rd[t] = 1.0 / s_L[t * PANEL_LD + t];
    __syncthreads();
    if (blockIdx.x == 0) {
        for (int e = t; e < nb * nb; e += PANEL_THREADS) {
            const int c = e / nb, r = e - c * nb;
            if (r >= c) AB[(j0 + r) + (j0 + c) * lda] = s_L[r * PANEL_LD + c];
        }
        if (t < nb) y[j0 + t] = s_y[t];
    }
    // TRSM: row r of A21 (global row j0+nb+r):  x L11^T = a, forward over the nb columns
    const int r = blockIdx.x * PANEL_THREADS + t;
    if (r < m) {
        const long long gi = j0 + nb + r;
        double xr[SOLVE_NB];
        double dot = 0.0;
#pragma unroll
        for (int c = 0; c < SOLVE_NB; ++c) {
            if (c < nb) {
                double s = AB[gi + (j0 + c) * lda];
#pragma unroll
                for (int cp = 0; cp < c; ++cp) s = fma(-xr[cp], s_L[c * PANEL_LD + cp], s);
                s *= s_rd[c];
                xr[c] = s;
                AB[gi + (j0 + c) * lda] = s;
                dot = fma(s, s_y[c], dot);
            } else {
                xr[c] = 0.0;
            }
        }
        y[gi] -= dot;     // right-hand side below the block: g2 -= L21 y1
    }
}

// ------------------------------------------------------------------------------------------
// trailing update with FP64 tensor-core MMA
// ------------------------------------------------------------------------------------------
#define SYRK_TILE 64
#define SYRK_LD 68      // 64 + 4: t4*68 + g hits 16 distinct 8-byte banks per half-warp (conflict-free LDS.64)
#define SYRK_THREADS 128

__device__ __forceinline__ void spl_dmma_8x8x4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// C[I,J] -= P[I,:] P[J,:]^T over the lower-triangular tiles of the m x m window whose first
// row/column is global index r0; P = panel rows r0.., columns j0..j0+nb-1.
__global__ void __launch_bounds__(SYRK_THREADS)
spl_syrk_kernel(double *__restrict__ AB, long long lda, long long r0, long long j0, int nb, int m,
                const int *__restrict__ fail) {
    extern __shared__ __align__(16) double s_ab[];
    double *sA = s_ab;                          // [k][row], SOLVE_NB x SYRK_LD
    double *sB = s_ab + SOLVE_NB * SYRK_LD;
    if (*fail) return;
    // linear tile id -> (ti >= tj)
    const int tile = blockIdx.x;
    int ti = (int)((sqrt(8.0 * (double)tile + 1.0) - 1.0) * 0.5);
    while ((long long)(ti + 1) * (ti + 2) / 2 <= tile) ++ti;
    while ((long long)ti * (ti + 1) / 2 > tile) --ti;
    const int tj = tile - ti * (ti + 1) / 2;
    const int I0 = ti * SYRK_TILE, J0 = tj * SYRK_TILE;
    const int t = threadIdx.x;

    for (int e = t; e < SOLVE_NB * SYRK_TILE; e += SYRK_THREADS) {
        const int k = e / SYRK_TILE, r = e - k * SYRK_TILE;   // coalesced over r
        double va = 0.0, vb = 0.0;
        if (k < nb) {
            if (I0 + r < m) va = AB[(r0 + I0 + r) + (j0 + k) * lda];
            if (J0 + r < m) vb = AB[(r0 + J0 + r) + (j0 + k) * lda];
        }
        sA[k * SYRK_LD + r] = va;
        sB[k * SYRK_LD + r] = vb;
    }
    __syncthreads();

    const int warp = t >> 5, lane = t & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int wy = (warp >> 1) * 32, wx = (warp & 1) * 32;
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

#pragma unroll 4
    for (int k0 = 0; k0 < SOLVE_NB; k0 += 4) {
        double a[4], b[4];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) a[mi] = sA[(k0 + t4) * SYRK_LD + wy + mi * 8 + g];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) b[ni] = sB[(k0 + t4) * SYRK_LD + wx + ni * 8 + g];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) spl_dmma_8x8x4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }

    // epilogue: C -= acc on the lower triangle, inside the window
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int li = I0 + wy + mi * 8 + g;
                const int lj = J0 + wx + ni * 8 + 2 * t4 + h;
                if (li < m && lj < m && li >= lj) {
                    double *p = AB + (r0 + li) + (r0 + lj) * lda;
                    *p -= acc[mi][ni][h];
                }
            }
}

// ------------------------------------------------------------------------------------------
// back-substitution L^T c = y, block by block from the end
// ------------------------------------------------------------------------------------------
#define BACK_THREADS 128
__global__ void __launch_bounds__(BACK_THREADS)
spl_backsolve_kernel(const double *__restrict__ AB, long long lda, long long j0, int nb, int bw,
                     double *__restrict__ y, const int *__restrict__ fail) {
    __shared__ double s_L[SOLVE_NB * PANEL_LD];
    __shared__ double s_c[SOLVE_NB];
    const int t = threadIdx.x;
    if (*fail) return;
    for (int e = t; e < nb * nb; e += BACK_THREADS) {
        const int c = e / nb, r = e - c * nb;
        s_L[r * PANEL_LD + c] = (r >= c) ? AB[(j0 + r) + (j0 + c) * lda] : 0.0;
    }
    if (t < nb) s_c[t] = y[j0 + t];
    __syncthreads();
    // solve L11^T c = y_k (warp 0): from the last row up
    if (t < 32) {
        for (int k = nb - 1; k >= 0; --k) {
            const double ck = s_c[k] / s_L[k * PANEL_LD + k];
            __syncwarp();
            if (t == 0) s_c[k] = ck;
            for (int i = t; i < k; i += 32) s_c[i] = fma(-s_L[k * PANEL_LD + i], ck, s_c[i]);
            __syncwarp();
        }
    }
    __syncthreads();
    if (blockIdx.x == 0 && t < nb) y[j0 + t] = s_c[t];
    // eliminate c_k from the bw preceding unknowns: y[j] -= sum_i L[i][j] c[i], i in the block
    const long long jlo = (j0 - bw > 0) ? j0 - bw : 0;
    const long long j = jlo + (long long)blockIdx.x * BACK_THREADS + t;
    if (j < j0) {
        const double *col = AB + j0 + j * lda;     // rows j0.. of column j, contiguous
        double s = 0.0;
#pragma unroll 8
        for (int i = 0; i < nb; ++i) s = fma(col[i], s_c[i], s);
        y[j] -= s;
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

// AB must hold ncol*lda + lda doubles and be zero-filled on entry.  y enters as g, leaves as coef.
int spl_solve_launch(const GridParams &gp, const double *d_S, double *d_AB, double *d_y, int *d_fail,
                     cudaStream_t st, int nsm, cudaEvent_t *ev) {
    const long long n = gp.ncol;
    const int bw = spl_half_bandwidth(gp);
    const long long lda = spl_band_lda(bw);
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
    const size_t syrk_smem = sizeof(double) * 2 * SOLVE_NB * SYRK_LD;
    SPL_CUDA_TRY(cudaFuncSetAttribute(spl_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)syrk_smem));
    for (long long j0 = 0; j0 < n; j0 += SOLVE_NB) {
        const int nb = (int)((n - j0 < SOLVE_NB) ? n - j0 : SOLVE_NB);
        const long long r0 = j0 + nb;
        long long mm = n - r0;
        if (mm > bw) mm = bw;
        const int m = (int)mm;
        int pblocks = (m + PANEL_THREADS - 1) / PANEL_THREADS;
        if (pblocks < 1) pblocks = 1;
        spl_panel_kernel<<<pblocks, PANEL_THREADS, 0, st>>>(d_AB, lda, n, j0, nb, m, d_y, d_fail);
        ++g_spl_launches;
        if (m > 0) {
            const int T = (m + SYRK_TILE - 1) / SYRK_TILE;
            const int tiles = T * (T + 1) / 2;
            spl_syrk_kernel<<<tiles, SYRK_THREADS, syrk_smem, st>>>(d_AB, lda, r0, j0, nb, m, d_fail);
            ++g_spl_launches;
        }
    }
    if (ev) cudaEventRecord(ev[2], st);
    const long long nblk = (n + SOLVE_NB - 1) / SOLVE_NB;
    for (long long kb = nblk - 1; kb >= 0; --kb) {
        const long long j0 = kb * SOLVE_NB;
        const int nb = (int)((n - j0 < SOLVE_NB) ? n - j0 : SOLVE_NB);
        const long long span = (j0 < bw) ? j0 : bw;
        int blocks = (int)((span + BACK_THREADS - 1) / BACK_THREADS);
        if (blocks < 1) blocks = 1;
        spl_backsolve_kernel<<<blocks, BACK_THREADS, 0, st>>>(d_AB, lda, j0, nb, bw, d_y, d_fail);
        ++g_spl_launches;
    }
    if (ev) cudaEventRecord(ev[3], st);
    SPL_CUDA_TRY(cudaGetLastError());
    return SPLPAK_OK;
}
