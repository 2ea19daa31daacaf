// eval.cu -- batched splfe / splde (src/splpak.F90:1089-1275) for sm_100a.
//
// One query per thread.  The whole coefficient table (<= 166 KB at the named configs) is staged
// once per CTA into shared memory with a 1-D bulk async copy (cp.async.bulk -> SASS UBLKCP)
// signalled through an mbarrier; CTAs are persistent (grid = k * SM count) and walk the query
// stream with a grid stride, so the table is staged k*148 times per launch, not once per tile.
// Tables that do not fit in shared memory are gathered through the read-only path from L2.
//
// Per query: four 1-D weights per dimension (basis.cuh, same formulas and box as the reference),
// then the nested contraction  sum_k b3[k] sum_j b2[j] sum_i coef[..]*b1[i]  with dimension 1
// innermost (4 contiguous coefficients).  Algorithmic HBM traffic: (ndim + 1) reals per query.
#include "basis.cuh"

struct DerivParams {
    int nd[SPL_MAXDIM];
};

#define EVAL_THREADS 512

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// Stage `bytes` (multiple of 16, 16-byte aligned src/dst) global -> shared with cp.async.bulk.
__device__ __forceinline__ void bulk_stage(double *dst, const double *src, uint32_t bytes,
                                           uint64_t *mbar) {
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)),
                     "r"(bytes)
                     : "memory");
        const uint32_t chunk = 32768;
        for (uint32_t off = 0; off < bytes; off += chunk) {
            const uint32_t nb = min(chunk, bytes - off);
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    smem_u32((const char *)dst + off)),
                "l"((const char *)src + off), "r"(nb), "r"(smem_u32(mbar))
                : "memory");
        }
    }
    // all threads wait for phase 0 to complete
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(mbar))
            : "memory");
    }
}

template <int NDIM, bool SMEM>
__global__ void __launch_bounds__(EVAL_THREADS)
spl_eval_kernel(const __grid_constant__ GridParams gp, const DerivParams dp,
                const real_t *__restrict__ x, int l1x, long long nq,
                const double *__restrict__ coef, long long ncol_padded, real_t *__restrict__ out) {
    extern __shared__ __align__(128) double s_coef[];
    __shared__ __align__(8) uint64_t mbar;
    const double *cf = coef;
    if (SMEM) {
        bulk_stage(s_coef, coef, (uint32_t)(ncol_padded * sizeof(double)), &mbar);
        cf = s_coef;
    }

    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride) {
        double b[NDIM][4];
        int ws[NDIM];
        const real_t *xq = x + q * (long long)l1x;
#pragma unroll
        for (int d = 0; d < NDIM; ++d) {
            const double xd = (double)xq[d];
            spl_window_weights(xd, gp.xmin[d], gp.dx[d], gp.dxin[d], gp.nodes[d], dp.nd[d], ws[d],
                               b[d]);
        }
        double sum = 0.0;
        if (NDIM == 1) {
            const double *p = cf + ws[0];
#pragma unroll
            for (int i = 0; i < 4; ++i) sum = fma(p[i], b[0][i], sum);
        } else if (NDIM == 2) {
            const int n0 = gp.nodes[0];
            const double *p0 = cf + ws[0] + (long long)n0 * ws[1];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double *p = p0 + (long long)n0 * j;
                double sj = 0.0;
#pragma unroll
                for (int i = 0; i < 4; ++i) sj = fma(p[i], b[0][i], sj);
                sum = fma(sj, b[1 % NDIM][j], sum);
            }
        } else if (NDIM == 3) {
            const int n0 = gp.nodes[0];
            const long long n01 = (long long)n0 * gp.nodes[1];
            const double *p0 = cf + ws[0] + (long long)n0 * ws[1 % NDIM] + n01 * ws[2 % NDIM];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                double sk = 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double *p = p0 + (long long)n0 * j + n01 * k;
                    double sj = 0.0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) sj = fma(p[i], b[0][i], sj);
                    sk = fma(sj, b[1 % NDIM][j], sk);
                }
                sum = fma(sk, b[2 % NDIM][k], sum);
            }
        } else {
            const int n0 = gp.nodes[0];
            const long long n01 = (long long)n0 * gp.nodes[1];
            const long long n012 = n01 * gp.nodes[2];
            const double *p0 = cf + ws[0] + (long long)n0 * ws[1 % NDIM] + n01 * ws[2 % NDIM] +
                               n012 * ws[3 % NDIM];
#pragma unroll 1
            for (int l = 0; l < 4; ++l) {
                double sl = 0.0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    double sk = 0.0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double *p = p0 + (long long)n0 * j + n01 * k + n012 * l;
                        double sj = 0.0;
#pragma unroll
                        for (int i = 0; i < 4; ++i) sj = fma(p[i], b[0][i], sj);
                        sk = fma(sj, b[1 % NDIM][j], sk);
                    }
                    sl = fma(sk, b[2 % NDIM][k], sl);
                }
                sum = fma(sl, b[3 % NDIM][l], sum);
            }
        }
        out[q] = (real_t)sum;
    }
}

template <int NDIM>
static int launch_eval(const GridParams &gp, const DerivParams &dp, const real_t *d_x, int l1x,
                       long long nq, const double *d_coef, long long ncol_padded, real_t *d_out,
                       cudaStream_t stream, int nsm, size_t smem_optin) {
    const size_t coef_bytes = (size_t)ncol_padded * sizeof(double);
    const bool use_smem = coef_bytes + 1024 <= smem_optin && coef_bytes < (1u << 20);
    long long blocks_needed = (nq + EVAL_THREADS - 1) / EVAL_THREADS;
    if (blocks_needed < 1) blocks_needed = 1;
    if (use_smem) {
        auto kern = spl_eval_kernel<NDIM, true>;
        SPL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)coef_bytes));
        int per_sm = 1;
        SPL_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, EVAL_THREADS,
                                                                   coef_bytes));
        if (per_sm < 1) per_sm = 1;
        long long grid = (long long)nsm * per_sm;
        if (grid > blocks_needed) grid = blocks_needed;
        kern<<<(unsigned)grid, EVAL_THREADS, coef_bytes, stream>>>(gp, dp, d_x, l1x, nq, d_coef,
                                                                   ncol_padded, d_out);
    } else {
        auto kern = spl_eval_kernel<NDIM, false>;
        long long grid = (long long)nsm * 4;
        if (grid > blocks_needed) grid = blocks_needed;
        kern<<<(unsigned)grid, EVAL_THREADS, 0, stream>>>(gp, dp, d_x, l1x, nq, d_coef, ncol_padded,
                                                          d_out);
    }
    ++g_spl_launches;
    SPL_CUDA_TRY(cudaGetLastError());
    return SPLPAK_OK;
}

// d_coef: float64 device table with ncol_padded (even, >= ncol) entries.
int spl_eval_launch(const GridParams &gp, const int *nderiv, const real_t *d_x, int l1x, long long nq,
                    const double *d_coef, long long ncol_padded, real_t *d_out, cudaStream_t stream,
                    int nsm, size_t smem_optin) {
    DerivParams dp;
    for (int d = 0; d < SPL_MAXDIM; ++d) dp.nd[d] = (nderiv && d < gp.ndim) ? nderiv[d] : 0;
    if (nq <= 0) return SPLPAK_OK;
    switch (gp.ndim) {
    case 1: return launch_eval<1>(gp, dp, d_x, l1x, nq, d_coef, ncol_padded, d_out, stream, nsm, smem_optin);
    case 2: return launch_eval<2>(gp, dp, d_x, l1x, nq, d_coef, ncol_padded, d_out, stream, nsm, smem_optin);
    case 3: return launch_eval<3>(gp, dp, d_x, l1x, nq, d_coef, ncol_padded, d_out, stream, nsm, smem_optin);
    case 4: return launch_eval<4>(gp, dp, d_x, l1x, nq, d_coef, ncol_padded, d_out, stream, nsm, smem_optin);
    }
    return SPLPAK_ERR_NDIM;
}
