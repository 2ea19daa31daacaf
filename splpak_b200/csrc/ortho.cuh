// ortho.cuh -- orthogonal-transformation fit path (included at the end of assemble.cu; 1-D..3-D).
//
// The reference solves the row stream with Householder / Givens transformations (suprls, src/splpak.F90:1375-1695:
// Householder against the triangle :1516-1549, triangularisation of new rows :1569-1609), i.e. at cond(A).  The default
// GPU path factors the normal equations (cond(A)^2, repaired by corrected-semi-normal-equation steps while
// eps*cond(A)^2 < 1).  This file is the variant the north star keeps for ill-conditioned, constraint-dominated fits: the
// same least-squares rows, reduced by Householder reflections only -- a two-level TSQR that exploits what the rows
// look like:
//
//   stage 1  (spl_window_qr_kernel, one CTA per window, windows in parallel)
//            every data row touches only the 4^ndim columns of its window (SURVEY B), so the rows of a window are
//            reduced on the spot to a 4^ndim x 4^ndim triangle R_w and the transformed right-hand side z_w:
//            points arrive window-sorted (the counting sort of the direct assembly), 64 rows at a time are generated
//            in shared memory (w * phi, w * y, :806, :837) and annihilated against R_w column by column with the
//            reflector of :1527-1547 (sign opposite to the old diagonal).  The derivative-constraint rows of a
//            data-sparse node (:921-1046) live in the 3^ndim box around the node, which lies inside ONE window, and are
//            absorbed the same way (second pass of the same kernel).  1e8 rows become 9,261 x 64 rows.
//   stage 2  (spl_band_qr_kernel, persistent, software-pipelined over the SMs)
//            the stacked R_w are swept, in window order (= order of their first column), into the global upper BAND
//            factor Rb (half bandwidth b as in solve.cu): a block of <= NRB rows lives in shared memory as a dense
//            strip W over the columns [c0, c0 + b], and for every column the reflector built from (Rb[c][c], W[:, c])
//            updates row c of Rb and the strip.  Rows of Rb to the right of the strip have not been touched by any
//            earlier block (blocks come in order of their first column), so the strip never grows.  Block k+1 may
//            work on column c as soon as block k has released it: the CTAs form a pipeline, each one column behind
//            its predecessor (release/acquire flags in global memory, all CTAs co-resident).
//   solve    R c = Q^T r by back-substitution in the band (one CTA; exactly-zero pivot -> 107 like :1662).
//
// Cost at cfg3: stage 1 ~8.4 kflop per point, stage 2 ~2 * 65 * b^2 per window = 3.9e12 flops in total (the Cholesky
// path: 4.5e10) -- this is the accurate path, not the fast one.  4-D (256-column windows) is not supported here.
#pragma once

#define ORTHO_THREADS 128
#define ORTHO_BR 64            // rows generated / absorbed per batch in stage 1
#define ORTHO_NRB_MAX 16       // rows of one stage-2 block
#define ORTHO_S2_THREADS 256
#define ORTHO_TINY2 1e-280     // squared magnitude below which a column of new rows counts as zero (see ortho_absorb)


// ------------------------------------------------------------------------------------------
// one Householder step shared by the stage-1 passes: annihilate column j of the nb batch rows in s_B against
// s_R[j][j]; columns j+1..ncw (ncw = rhs) are updated by one thread each.  Uniform control flow.
// ------------------------------------------------------------------------------------------
template <int NCW>
__device__ __forceinline__ void ortho_absorb(double *s_R, double *s_B, int nb, double *s_red) {
    constexpr int LD = NCW + 2;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int j = 0; j < NCW; ++j) {
        double sq = 0.0;
        if (t < ORTHO_BR && t < nb) {
            const double v = s_B[t * LD + j];
            sq = v * v;
        }
        if (t < ORTHO_BR) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            if (lane == 0) s_red[warp] = sq;
        }
        __syncthreads();
        const double sigma = s_red[0] + s_red[1];
        // :1527 `if (s==0) cycle`.  A column whose batch entries are all below 1e-140 holds nothing but the rounding residue
        // of rows that were already absorbed (each step shrinks it by ~eps, down into the subnormals, where 1 / (v0 beta)
        // overflows): it is skipped like an exactly zero one.
        if (sigma > ORTHO_TINY2) {
            const double alpha = s_R[j * LD + j];
            double beta = sqrt(alpha * alpha + sigma);
            if (alpha > 0.0) beta = -beta;                              // :1530
            const double v0 = alpha - beta;                             // :1531
            const double temp1 = 1.0 / (v0 * beta);                     // :1532
            const int k = j + 1 + t;
            if (k <= NCW) {
                double s = v0 * s_R[j * LD + k];
                for (int r = 0; r < nb; ++r) s = fma(s_B[r * LD + j], s_B[r * LD + k], s);
                s *= temp1;
                s_R[j * LD + k] = fma(s, v0, s_R[j * LD + k]);
                for (int r = 0; r < nb; ++r) s_B[r * LD + k] = fma(s, s_B[r * LD + j], s_B[r * LD + k]);
            }
            __syncthreads();
            if (t == 0) s_R[j * LD + j] = beta;
        }
        __syncthreads();
    }
}

// pass 0: data rows of the chunk (window-sorted through perm); pass 1: derivative-constraint rows of the sparse nodes.
template <int NDIM, int PASS>
__global__ void __launch_bounds__(ORTHO_THREADS)
spl_window_qr_kernel(const __grid_constant__ GridParams gp, const real_t *__restrict__ x, int l1x,
                     const real_t *__restrict__ y, const real_t *__restrict__ w, int weighted,
                     const unsigned *__restrict__ perm, const unsigned *__restrict__ wincount,
                     const unsigned *__restrict__ winstart, double xtrap, const double *__restrict__ cnt,
                     double *__restrict__ totals, double *__restrict__ Rw, unsigned *__restrict__ work) {
    constexpr int NCW = spl_ipow(4, NDIM), LD = NCW + 2;
    extern __shared__ __align__(16) double s_o[];
    double *s_R = s_o;                       // NCW x LD
    double *s_B = s_R + NCW * LD;            // ORTHO_BR x LD
    double *s_red = s_B + ORTHO_BR * LD;     // 4
    __shared__ unsigned s_w;
    const int t = threadIdx.x;
    for (;;) {
        __syncthreads();
        if (t == 0) s_w = atomicAdd(work, 1u);
        __syncthreads();
        const long long win = s_w;
        if (win >= gp.nwindows) break;
        int ws[NDIM];
        {
            long long k = win;
#pragma unroll
            for (int d = 0; d < NDIM; ++d) {
                ws[d] = (int)(k % gp.nwin[d]);
                k /= gp.nwin[d];
            }
        }
        double *Rg = Rw + win * (long long)(NCW * (NCW + 1));
        bool loaded = false;
        auto load_R = [&]() {
            for (int e = t; e < NCW * (NCW + 1); e += ORTHO_THREADS) s_R[(e / (NCW + 1)) * LD + e % (NCW + 1)] = Rg[e];
            loaded = true;
        };
        if (PASS == 0) {
            const unsigned m = wincount[win];
            if (m == 0u) continue;
            load_R();
            const long long first = winstart[win];
            for (unsigned b0 = 0; b0 < m; b0 += ORTHO_BR) {
                const int nb = (int)min((unsigned)ORTHO_BR, m - b0);
                __syncthreads();
                if (t < nb) {
                    const long long i = perm[first + b0 + t];
                    const double rowwt = weighted ? (double)w[i] : 1.0;
                    double bb[NDIM][4];
#pragma unroll
                    for (int d = 0; d < NDIM; ++d) {
                        int wsd;
                        spl_window_weights_value((double)x[i * (long long)l1x + d], gp.xmin[d], gp.dx[d], gp.dxin[d],
                                                 gp.nodes[d], wsd, bb[d]);
                    }
                    for (int q = 0; q < NCW; ++q) {
                        double basm = 1.0;
#pragma unroll
                        for (int d = 0; d < NDIM; ++d) basm = spl_mul(basm, bb[d][(q >> (2 * d)) & 3]);    // :383
                        s_B[t * LD + q] = spl_mul(rowwt, basm);                                            // :837
                    }
                    s_B[t * LD + NCW] = spl_mul(rowwt, (double)y[i]);                                      // :806
                }
                __syncthreads();
                ortho_absorb<NCW>(s_R, s_B, nb, s_red);
            }
        } else {
            // nodes whose 3^ndim constraint box lies in this window: clamp(in - 1, 0, nod - 4) == ws
            int lo[NDIM], hi[NDIM];
            long long ncand = 1;
#pragma unroll
            for (int d = 0; d < NDIM; ++d) {
                lo[d] = (ws[d] == 0) ? 0 : ws[d] + 1;
                hi[d] = (ws[d] == gp.nodes[d] - 4) ? gp.nodes[d] - 1 : ws[d] + 1;
                ncand *= hi[d] - lo[d] + 1;
            }
            long long nrect = 1;
#pragma unroll
            for (int d = 0; d < NDIM; ++d) nrect *= (gp.nodes[d] - 1);
            const double wtprrc = __ddiv_rn(totals[0], (double)nrect);                                      // :910
            constexpr int NPAIR = NDIM * (NDIM + 1) / 2;
            int nb = 0;                                                                                     // rows waiting in s_B
            for (long long cnd = 0; cnd < ncand; ++cnd) {
                int in[NDIM];
                long long k = cnd, node = 0, nstride = 1;
#pragma unroll
                for (int d = 0; d < NDIM; ++d) {
                    in[d] = lo[d] + (int)(k % (hi[d] - lo[d] + 1));
                    k /= (hi[d] - lo[d] + 1);
                    node += (long long)in[d] * nstride;
                    nstride *= gp.nodes[d];
                }
                double expect = wtprrc;
#pragma unroll
                for (int d = 0; d < NDIM; ++d)
                    if (in[d] == 0 || in[d] == gp.nodes[d] - 1) expect = spl_mul(0.5, expect);             // :927-929
                const double have = cnt[node];
                if (!(have < spl_mul(0.75, expect))) continue;                                              // :936 (uniform)
                const double dcwght = spl_mul(xtrap, spl_sub(expect, have));                                // :938, :960
                if (!loaded) {
                    __syncthreads();
                    load_R();
                }
                if (t == 0) atomicAdd(totals + 1, (double)NPAIR);
                int ibmn[NDIM], ibmx[NDIM];
                double xn[NDIM];
#pragma unroll
                for (int d = 0; d < NDIM; ++d) {
                    xn[d] = spl_add(gp.xmin[d], spl_mul((double)in[d], gp.dx[d]));                          // :943
                    ibmn[d] = (in[d] == 0) ? 0 : in[d] - 1;
                    ibmx[d] = (in[d] == gp.nodes[d] - 1) ? in[d] : in[d] + 1;
                }
                for (int idm = 0; idm < NDIM; ++idm)
                    for (int jdm = idm; jdm < NDIM; ++jdm) {
                        int nder[NDIM];
#pragma unroll
                        for (int d = 0; d < NDIM; ++d) nder[d] = 0;
                        bool boundary = true;
                        double rowwt = spl_mul(2.0, dcwght);                                                // :983
                        if (jdm == idm) {
                            rowwt = dcwght;
                            nder[jdm] = 2;
                            if (in[idm] != 0 && in[idm] != gp.nodes[idm] - 1) boundary = false;
                        }
                        if (boundary) {
                            nder[idm] = 1;
                            nder[jdm] = 1;
                        }
                        if (nb == ORTHO_BR) {
                            __syncthreads();
                            ortho_absorb<NCW>(s_R, s_B, nb, s_red);
                            nb = 0;
                        }
                        __syncthreads();
                        for (int q = t; q <= NCW; q += ORTHO_THREADS) {
                            double val = 0.0;
                            if (q < NCW) {
                                double basm = 1.0;
                                bool inside = true;
#pragma unroll
                                for (int d = 0; d < NDIM; ++d) {
                                    const int ib = ws[d] + ((q >> (2 * d)) & 3);
                                    if (ib < ibmn[d] || ib > ibmx[d]) inside = false;
                                    basm = spl_mul(basm, spl_bas1(ib, gp.nodes[d], nder[d], xn[d], gp.xmin[d], gp.dx[d], gp.dxin[d]));
                                }
                                val = inside ? spl_mul(rowwt, basm) : 0.0;
                            }
                            s_B[nb * LD + q] = val;                                                         // rhs = 0 (:866)
                        }
                        ++nb;
                    }
            }
            if (nb > 0) {
                __syncthreads();
                ortho_absorb<NCW>(s_R, s_B, nb, s_red);
            }
            if (!loaded) continue;
        }
        __syncthreads();
        for (int e = t; e < NCW * (NCW + 1); e += ORTHO_THREADS) Rg[e] = s_R[(e / (NCW + 1)) * LD + e % (NCW + 1)];
    }
}

// ------------------------------------------------------------------------------------------
// stage 2
// ------------------------------------------------------------------------------------------
// compacted, ordered list of the non-zero row blocks (window, rb): one warp per candidate, then a single-CTA scan
__global__ void __launch_bounds__(256)
spl_ortho_flag_kernel(const double *__restrict__ Rw, long long nwindows, int ncw, int nrb, int nblk_win,
                      unsigned *__restrict__ flags) {
    const long long gwarp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gwarp >= nwindows * nblk_win) return;
    const long long win = gwarp / nblk_win;
    const int rb = (int)(gwarp % nblk_win);
    const double *R = Rw + win * (long long)(ncw * (ncw + 1));
    const int r0 = rb * nrb, r1 = min(ncw, r0 + nrb);
    bool nz = false;
    for (int e = r0 * (ncw + 1) + lane; e < r1 * (ncw + 1); e += 32) nz |= (R[e] != 0.0);
    nz = __any_sync(0xffffffffu, nz);
    if (lane == 0) flags[gwarp] = nz ? 1u : 0u;
}
__global__ void __launch_bounds__(1024)
spl_ortho_compact_kernel(const unsigned *__restrict__ flags, long long n, unsigned *__restrict__ blist,
                         unsigned *__restrict__ meta2) {
    __shared__ unsigned s_a[1024];
    const int t = threadIdx.x;
    const long long per = (n + 1023) / 1024, lo = (long long)t * per, hi = (lo + per < n) ? lo + per : n;
    unsigned c = 0;
    for (long long k = lo; k < hi; ++k) c += flags[k];
    s_a[t] = c;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        unsigned v = 0;
        if (t >= off) v = s_a[t - off];
        __syncthreads();
        s_a[t] += v;
        __syncthreads();
    }
    unsigned pos = s_a[t] - c;
    for (long long k = lo; k < hi; ++k)
        if (flags[k]) blist[pos++] = (unsigned)k;
    if (t == 1023) meta2[0] = s_a[1023];
}

__device__ __forceinline__ long long ortho_ld_acquire(const long long *p) {
    long long v;
    asm volatile("ld.acquire.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ortho_st_release(long long *p, long long v) {
    asm volatile("st.release.gpu.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Persistent: CTA k takes the blocks k, k + G, ... (G = gridDim.x co-resident CTAs).  progress[b] = number of global
// columns block b has released (ncol + 1 when it is done); block b touches column c only after progress[b-1] > c.
template <int NDIM>
__global__ void __launch_bounds__(ORTHO_S2_THREADS, 1)
spl_band_qr_kernel(const __grid_constant__ GridParams gp, const double *__restrict__ Rw,
                   const unsigned *__restrict__ blist, const unsigned *__restrict__ meta2, int nrb, int nblk_win, int bw,
                   double *Rb, long long *progress) {
    constexpr int NCW = spl_ipow(4, NDIM);
    extern __shared__ __align__(16) double s_W[];          // nrb x (bw + 2)
    __shared__ int s_off[NCW];
    const int t = threadIdx.x;
    const long long n = gp.ncol;
    const int ldw = bw + 2;
    const unsigned nb = meta2[0];
    if (t < NCW) {
        int o = 0, stride = 1;
#pragma unroll
        for (int d = 0; d < NDIM; ++d) {
            o += ((t >> (2 * d)) & 3) * stride;
            stride *= gp.nodes[d];
        }
        s_off[t] = o;
    }
    __syncthreads();
    for (unsigned b = blockIdx.x; b < nb; b += gridDim.x) {
        const unsigned id = blist[b];
        const long long win = id / (unsigned)nblk_win;
        const int rb = (int)(id % (unsigned)nblk_win);
        const int r0 = rb * nrb, nr = min(NCW, r0 + nrb) - r0;
        long long c0 = 0;
        {
            long long k = win, stride = 1;
#pragma unroll
            for (int d = 0; d < NDIM; ++d) {
                c0 += (k % gp.nwin[d]) * stride;
                k /= gp.nwin[d];
                stride *= gp.nodes[d];
            }
        }
        // strip: W[r][off[q]] = R_w[r0 + r][q] (q >= r0 + r), rhs in column bw + 1
        for (int e = t; e < nr * ldw; e += ORTHO_S2_THREADS) s_W[e] = 0.0;
        __syncthreads();
        const double *R = Rw + win * (long long)(NCW * (NCW + 1));
        for (int e = t; e < nr * (NCW + 1); e += ORTHO_S2_THREADS) {
            const int r = e / (NCW + 1), q = e % (NCW + 1);
            if (q == NCW) s_W[r * ldw + bw + 1] = R[(r0 + r) * (NCW + 1) + NCW];
            else if (q >= r0 + r) s_W[r * ldw + s_off[q]] = R[(r0 + r) * (NCW + 1) + q];
        }
        long long seen = (b == 0) ? (n + 1) : 0;             // last value read from progress[b - 1]
        const int j0 = s_off[r0];
        const long long jend = (c0 + bw < n - 1) ? bw : (n - 1 - c0);
        // The first wait.  This block touches no column below cf = c0 + j0, so while it waits for its predecessor to
        // release cf it FORWARDS the predecessor's progress: without that, a block that starts far to the right (the last
        // rows of a window triangle start ~b columns in) would hold back every later block, and the pipeline would run
        // one block at a time (measured: 67 s instead of < 1 s at cfg3).
        if (t == 0) {
            const long long cf = c0 + j0;
            if (b > 0) {
                long long published = 0;
                for (;;) {
                    seen = ortho_ld_acquire(progress + b - 1);
                    if (seen > cf) break;
                    if (seen > published) {
                        ortho_st_release(progress + b, seen);
                        published = seen;
                    }
                }
            }
            ortho_st_release(progress + b, cf);
        }
        __syncthreads();
        for (long long j = j0; j <= jend; ++j) {
            const long long c = c0 + j;
            double *rc = Rb + c * (long long)ldw;
            double wj[ORTHO_NRB_MAX];
            double sigma = 0.0;
#pragma unroll
            for (int r = 0; r < ORTHO_NRB_MAX; ++r) {
                wj[r] = (r < nr) ? s_W[r * ldw + j] : 0.0;
                sigma = fma(wj[r], wj[r], sigma);
            }
            if (sigma > ORTHO_TINY2) {                        // uniform: every thread read the same values
                const double alpha = __ldcg(rc);                 // rows of Rb are written by other CTAs: L2 loads
                double beta = sqrt(alpha * alpha + sigma);
                if (alpha > 0.0) beta = -beta;
                const double v0 = alpha - beta;
                const double temp1 = 1.0 / (v0 * beta);
                // columns c + 1 .. c0 + bw of row c and of the strip, then the right-hand side
                const int ncols = (int)(bw - j);
                for (int tt = t; tt <= ncols; tt += ORTHO_S2_THREADS) {
                    const bool rhs = (tt == ncols);
                    const int kr = rhs ? bw + 1 : tt + 1;             // index inside row c of Rb
                    const int kw = rhs ? bw + 1 : (int)j + tt + 1;     // index inside the strip
                    if (!rhs && c + tt + 1 >= n) continue;
                    const double rv = __ldcg(rc + kr);
                    double s = v0 * rv;
#pragma unroll
                    for (int r = 0; r < ORTHO_NRB_MAX; ++r)
                        if (r < nr) s = fma(wj[r], s_W[r * ldw + kw], s);
                    s *= temp1;
                    rc[kr] = fma(s, v0, rv);
#pragma unroll
                    for (int r = 0; r < ORTHO_NRB_MAX; ++r)
                        if (r < nr) s_W[r * ldw + kw] = fma(s, wj[r], s_W[r * ldw + kw]);
                }
                __syncthreads();                              // every thread has read rc[0] before it changes
                if (t == 0) rc[0] = beta;
            }
            __syncthreads();
            if (t == 0) {
                ortho_st_release(progress + b, c + 1);        // cumulative: covers the CTA's writes ordered by the barrier
                if (b > 0 && j < jend)
                    while (seen <= c + 1) seen = ortho_ld_acquire(progress + b - 1);
            }
            __syncthreads();
        }
        if (t == 0) ortho_st_release(progress + b, n + 1);
        __syncthreads();
    }
}

// R c = z in the band, from the last row up (:1661-1690); a pivot that is exactly zero -> fail (suprls error 34)
__global__ void __launch_bounds__(1024)
spl_band_backsub_kernel(const double *__restrict__ Rb, long long n, int bw, double *csol, int *fail) {
    __shared__ double s_part[32];
    __shared__ int s_bad;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int ldw = bw + 2;
    if (t == 0) s_bad = 0;
    __syncthreads();
    for (long long i = n - 1; i >= 0; --i) {
        const double *ri = Rb + i * (long long)ldw;
        const long long kmax = (n - 1 - i < bw) ? n - 1 - i : bw;
        double s = 0.0;
        for (long long k = 1 + t; k <= kmax; k += 1024) s = fma(ri[k], csol[i + k], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) s_part[warp] = s;
        __syncthreads();
        if (t == 0) {
            double tot = 0.0;
            for (int q = 0; q < 32; ++q) tot += s_part[q];
            const double piv = ri[0];
            if (piv == 0.0) {
                s_bad = 1;
                csol[i] = 0.0;
            } else {
                csol[i] = (ri[bw + 1] - tot) / piv;
            }
        }
        __syncthreads();
    }
    if (t == 0 && s_bad) *fail = 1;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
int spl_ortho_supported(const GridParams &gp) { return gp.ndim >= 1 && gp.ndim <= 3; }

void spl_ortho_free(OrthoScratch &os) {
    void *p[] = {os.wincount, os.winstart, os.wincursor, os.itemstart, os.meta, os.perm, os.Rw, os.blist, os.meta2,
                 os.Rb, os.progress, os.csol};
    for (void *q : p)
        if (q) cudaFree(q);
    spl_hist_scratch_free(os.hist);
    os = OrthoScratch();
}

int spl_ortho_init(const GridParams &gp, OrthoScratch &os, cudaStream_t st, size_t smem_optin) {
    if (os.ready) return SPLPAK_OK;
    if (!spl_ortho_supported(gp)) return SPLPAK_ERR_HANDLE;
    os.ncw = 1;
    for (int d = 0; d < gp.ndim; ++d) os.ncw *= 4;
    long long b = 0, stride = 1;
    for (int d = 0; d < gp.ndim; ++d) {
        b += 3 * stride;
        stride *= gp.nodes[d];
    }
    if (b > gp.ncol - 1) b = gp.ncol - 1;
    os.bw = (int)b;
    long long nrb = (long long)((smem_optin > 16384 ? smem_optin - 16384 : 0) / (sizeof(double) * (size_t)(os.bw + 2)));
    if (nrb > ORTHO_NRB_MAX) nrb = ORTHO_NRB_MAX;
    if (nrb > os.ncw) nrb = os.ncw;
    if (nrb < 1) return SPLPAK_ERR_HANDLE;                  // the strip of even one row does not fit shared memory
    os.nrb = (int)nrb;
    os.nblk_win = (os.ncw + os.nrb - 1) / os.nrb;
    os.max_blocks = gp.nwindows * os.nblk_win;
    if (os.max_blocks >= (1LL << 32)) return SPLPAK_ERR_HANDLE;
    const long long nb = gp.nwindows;
    SPL_CUDA_TRY(cudaMalloc((void **)&os.wincount, sizeof(unsigned) * (size_t)nb));
    SPL_CUDA_TRY(cudaMalloc((void **)&os.winstart, sizeof(unsigned) * (size_t)nb));
    SPL_CUDA_TRY(cudaMalloc((void **)&os.wincursor, sizeof(unsigned) * (size_t)nb));
    SPL_CUDA_TRY(cudaMalloc((void **)&os.itemstart, sizeof(unsigned) * (size_t)nb));
    SPL_CUDA_TRY(cudaMalloc((void **)&os.meta, sizeof(unsigned) * 8));
    SPL_CUDA_TRY(cudaMalloc((void **)&os.Rw, sizeof(double) * (size_t)nb * os.ncw * (os.ncw + 1)));
    SPL_CUDA_TRY(cudaMalloc((void **)&os.blist, sizeof(unsigned) * (size_t)os.max_blocks * 2));   // flags behind the list
    SPL_CUDA_TRY(cudaMalloc((void **)&os.meta2, sizeof(unsigned) * 4));
    SPL_CUDA_TRY(cudaMalloc((void **)&os.Rb, sizeof(double) * (size_t)gp.ncol * (os.bw + 2)));
    SPL_CUDA_TRY(cudaMalloc((void **)&os.progress, sizeof(long long) * (size_t)os.max_blocks));
    SPL_CUDA_TRY(cudaMalloc((void **)&os.csol, sizeof(double) * (size_t)(gp.ncol + 2)));
    SPL_CUDA_TRY(cudaMemsetAsync(os.Rw, 0, sizeof(double) * (size_t)nb * os.ncw * (os.ncw + 1), st));
    os.ready = 1;
    return SPLPAK_OK;
}

int spl_ortho_reset(const GridParams &gp, OrthoScratch &os, cudaStream_t st) {
    if (!os.ready) return SPLPAK_OK;
    SPL_CUDA_TRY(cudaMemsetAsync(os.Rw, 0, sizeof(double) * (size_t)gp.nwindows * os.ncw * (os.ncw + 1), st));
    return SPLPAK_OK;
}

template <int NDIM>
static int ortho_add_chunk_t(const GridParams &gp, const real_t *d_x, int l1x, const real_t *d_y, const real_t *d_w,
                             int weighted, long long n, int do_hist, OrthoScratch &os, double *d_cnt, double *d_totals,
                             cudaStream_t st, int nsm) {
    constexpr int NCW = spl_ipow(4, NDIM);
    const long long nbins = gp.nwindows;
    if (n > os.perm_cap) {
        if (os.perm) cudaFree(os.perm);
        os.perm = nullptr;
        os.perm_cap = 0;
        SPL_CUDA_TRY(cudaMalloc((void **)&os.perm, sizeof(unsigned) * (size_t)n));
        os.perm_cap = n;
    }
    SPL_CUDA_TRY(cudaMemsetAsync(os.wincount, 0, sizeof(unsigned) * nbins, st));
    SPL_CUDA_TRY(cudaMemsetAsync(os.wincursor, 0, sizeof(unsigned) * nbins, st));
    SPL_CUDA_TRY(cudaMemsetAsync(os.meta, 0, sizeof(unsigned) * 8, st));
    long long cb = (n + 512LL * BIN_U - 1) / (512LL * BIN_U);
    const long long ccap = (long long)nsm * 4;
    const int cgrid = (int)(cb < ccap ? (cb < 1 ? 1 : cb) : ccap);
    if (do_hist) {
        int rh = spl_hist_scratch_init(gp, os.hist, st);
        if (rh == SPLPAK_OK) rh = spl_hist_prepare(os.hist, d_w, weighted, n, 1, st, nsm);
        if (rh != SPLPAK_OK) return rh;
    }
    spl_classify_kernel<NDIM, false, false><<<cgrid, 512, 0, st>>>(gp, d_x, l1x, d_w, weighted, n, (int)nbins, os.wincount,
                                                                   do_hist, os.hist.hq, os.hist.qparams, d_totals, d_y, nullptr);
    if (do_hist) spl_hist_finalize(os.hist, gp, d_cnt, d_totals, st);
    spl_scan_kernel<<<1, 1024, 0, st>>>(os.wincount, nbins, 1u << 30, os.winstart, os.itemstart, os.meta);
    long long nb = (n + 255) / 256;
    const long long cap = (long long)nsm * 8;
    spl_perm_kernel<NDIM, false><<<(int)(nb < cap ? nb : cap), 256, 0, st>>>(gp, d_x, l1x, d_w, weighted, n, os.winstart,
                                                                             os.wincursor, 1, os.perm);
    const size_t smem = sizeof(double) * (size_t)((NCW + ORTHO_BR) * (NCW + 2) + 8);
    auto kern = spl_window_qr_kernel<NDIM, 0>;
    SPL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    SPL_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ORTHO_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)nsm * per_sm;
    if (grid > nbins) grid = nbins;
    // meta[2] is the scan kernel's work counter (zero after it); the window loop claims from it
    kern<<<(unsigned)grid, ORTHO_THREADS, smem, st>>>(gp, d_x, l1x, d_y, d_w, weighted, os.perm, os.wincount, os.winstart, 0.0,
                                                      d_cnt, d_totals, os.Rw, os.meta + 2);
    g_spl_launches += 4;
    SPL_CUDA_TRY(cudaGetLastError());
    return SPLPAK_OK;
}

int spl_ortho_add_chunk(const GridParams &gp, const real_t *d_x, int l1x, const real_t *d_y, const real_t *d_w,
                        int weighted, long long n, int do_hist, OrthoScratch &os, double *d_cnt, double *d_totals,
                        cudaStream_t st, int nsm) {
    switch (gp.ndim) {
    case 1: return ortho_add_chunk_t<1>(gp, d_x, l1x, d_y, d_w, weighted, n, do_hist, os, d_cnt, d_totals, st, nsm);
    case 2: return ortho_add_chunk_t<2>(gp, d_x, l1x, d_y, d_w, weighted, n, do_hist, os, d_cnt, d_totals, st, nsm);
    case 3: return ortho_add_chunk_t<3>(gp, d_x, l1x, d_y, d_w, weighted, n, do_hist, os, d_cnt, d_totals, st, nsm);
    }
    return SPLPAK_ERR_HANDLE;
}

template <int NDIM>
static int ortho_compute_t(const GridParams &gp, double xtrap, OrthoScratch &os, const double *d_cnt, double *d_totals,
                           int *d_fail, cudaStream_t st, int nsm) {
    constexpr int NCW = spl_ipow(4, NDIM);
    if (xtrap != 0.0) {
        // derivative-constraint rows of the data-sparse nodes, absorbed into the triangle of the window around each node
        SPL_CUDA_TRY(cudaMemsetAsync(os.meta, 0, sizeof(unsigned) * 8, st));
        const size_t smem = sizeof(double) * (size_t)((NCW + ORTHO_BR) * (NCW + 2) + 8);
        auto kern = spl_window_qr_kernel<NDIM, 1>;
        SPL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 1;
        SPL_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ORTHO_THREADS, smem));
        if (per_sm < 1) per_sm = 1;
        long long grid = (long long)nsm * per_sm;
        if (grid > gp.nwindows) grid = gp.nwindows;
        kern<<<(unsigned)grid, ORTHO_THREADS, smem, st>>>(gp, nullptr, 0, nullptr, nullptr, 0, nullptr, os.wincount, os.winstart,
                                                          xtrap, d_cnt, d_totals, os.Rw, os.meta + 2);
        ++g_spl_launches;
    }
    unsigned *flags = os.blist + os.max_blocks;
    const long long nwarps = os.max_blocks;
    spl_ortho_flag_kernel<<<(unsigned)((nwarps * 32 + 255) / 256), 256, 0, st>>>(os.Rw, gp.nwindows, os.ncw, os.nrb,
                                                                                os.nblk_win, flags);
    spl_ortho_compact_kernel<<<1, 1024, 0, st>>>(flags, os.max_blocks, os.blist, os.meta2);
    SPL_CUDA_TRY(cudaMemsetAsync(os.Rb, 0, sizeof(double) * (size_t)gp.ncol * (os.bw + 2), st));
    SPL_CUDA_TRY(cudaMemsetAsync(os.progress, 0, sizeof(long long) * (size_t)os.max_blocks, st));
    SPL_CUDA_TRY(cudaMemsetAsync(os.csol, 0, sizeof(double) * (size_t)(gp.ncol + 2), st));
    {
        const size_t smem = sizeof(double) * (size_t)os.nrb * (os.bw + 2);
        auto kern = spl_band_qr_kernel<NDIM>;
        SPL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // the pipeline spins on flags of lower-numbered blocks: every CTA of the grid must be resident
        int per_sm = 0;
        SPL_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ORTHO_S2_THREADS, smem));
        if (per_sm < 1) return SPLPAK_ERR_HANDLE;
        long long grid = nsm;
        if (grid > os.max_blocks) grid = os.max_blocks;
        const GridParams a_gp = gp;
        const double *a_Rw = os.Rw;
        const unsigned *a_bl = os.blist, *a_m2 = os.meta2;
        int a_nrb = os.nrb, a_nbw = os.nblk_win, a_bw = os.bw;
        double *a_Rb = os.Rb;
        long long *a_pr = os.progress;
        void *args[] = {(void *)&a_gp, &a_Rw, &a_bl, &a_m2, &a_nrb, &a_nbw, &a_bw, &a_Rb, &a_pr};
        SPL_CUDA_TRY(cudaLaunchCooperativeKernel((const void *)kern, dim3((unsigned)grid), dim3(ORTHO_S2_THREADS), args, smem, st));
    }
    spl_band_backsub_kernel<<<1, 1024, 0, st>>>(os.Rb, gp.ncol, os.bw, os.csol, d_fail);
    g_spl_launches += 4;
    SPL_CUDA_TRY(cudaGetLastError());
    return SPLPAK_OK;
}

// solution left in os.csol (ncol doubles)
int spl_ortho_compute(const GridParams &gp, double xtrap, OrthoScratch &os, const double *d_cnt, double *d_totals,
                      int *d_fail, cudaStream_t st, int nsm) {
    switch (gp.ndim) {
    case 1: return ortho_compute_t<1>(gp, xtrap, os, d_cnt, d_totals, d_fail, st, nsm);
    case 2: return ortho_compute_t<2>(gp, xtrap, os, d_cnt, d_totals, d_fail, st, nsm);
    case 3: return ortho_compute_t<3>(gp, xtrap, os, d_cnt, d_totals, d_fail, st, nsm);
    }
    return SPLPAK_ERR_HANDLE;
}
