// eval.cu -- batched splfe / splde (src/splpak.F90:1089-1275) for sm_100a.
//
// One persistent CTA per SM.  The whole coefficient table (<= 166 KB at the named configs) is staged
// once per CTA into shared memory with 1-D bulk async copies (cp.async.bulk -> SASS UBLKCP) signalled
// through an mbarrier.  After that there is no CTA-wide synchronisation: every warp claims chunks of
// 256 queries from a global atomic counter (dynamic scheduling), one query per lane, the coordinates
// of the next 32 queries in flight while the current 32 are evaluated.  Tables that do not fit in
// shared memory are gathered through the read-only path from L2 (same kernel, SMEM = false).
//
// Per query: four 1-D weights per dimension (basis.cuh: same formulas and index box as the
// reference), then the nested contraction  sum_k b3[k] sum_j b2[j] sum_i coef[..]*b1[i]  with
// dimension 1 innermost (4 contiguous coefficients).
//
// What binds this kernel for scattered queries is the GATHER of the 4^ndim coefficients: 64 8-byte
// shared-memory loads per 3-D query at random addresses, ~3 bank-conflict wavefronts per half-warp
// each (measured: 12 crossbar cycles per query per SM; 4 would be the conflict-free floor).  Queries
// that arrive in coherent runs (raster order of an output grid) broadcast and are bound by
// instruction issue and the FP64 pipe instead.  Two regrouping schemes that make half-warps
// conflict-free by construction (a per-tile counting sort by bank class, then a persistent ring of
// per-bank-class FIFOs) were built and measured in round 1; both lost more to CTA-wide barriers and
// bookkeeping than they saved in wavefronts (experiments/eval_ring_r01.cu.txt,
// profiles/r01_eval_ring_experiment.md).
//
// Algorithmic HBM traffic: (ndim + 1) reals per query.
#include <stdlib.h>

#include "basis.cuh"

struct DerivParams {
    int nd[SPL_MAXDIM];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *mbar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *mbar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(mbar)), "r"(parity)
            : "memory");
    }
}
// global -> shared bulk copy (bytes: multiple of 16; src/dst 16-byte aligned), completion on mbar.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *mbar) {
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

// Strides (in doubles) of the coefficient table the kernel reads: the caller's natural layout
// (s1 = nodes(1), s2 = nodes(1)*nodes(2), ...) or the padded shared-memory copy of the ring kernel.
struct TableLayout {
    int s1, s2;
    long long s3;
};

// Nested contraction over the 4^NDIM window, dimension 1 innermost.
template <int NDIM>
__device__ __forceinline__ double spl_contract(const TableLayout &tl, const double *__restrict__ cf,
                                               const int *ws, const double (*b)[4]) {
    double sum = 0.0;
    if constexpr (NDIM == 1) {
        const double *p = cf + ws[0];
#pragma unroll
        for (int i = 0; i < 4; ++i) sum = fma(p[i], b[0][i], sum);
    } else if constexpr (NDIM == 2) {
        const int n0 = tl.s1;
        const double *p0 = cf + ws[0] + n0 * ws[1];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double *p = p0 + n0 * j;
            double sj = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) sj = fma(p[i], b[0][i], sj);
            sum = fma(sj, b[1][j], sum);
        }
    } else if constexpr (NDIM == 3) {
        const int n0 = tl.s1;
        const int n01 = tl.s2;
        const double *p0 = cf + ws[0] + n0 * ws[1] + (long long)n01 * ws[2];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double sk = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double *p = p0 + n0 * j + n01 * k;
                double sj = 0.0;
#pragma unroll
                for (int i = 0; i < 4; ++i) sj = fma(p[i], b[0][i], sj);
                sk = fma(sj, b[1][j], sk);
            }
            sum = fma(sk, b[2][k], sum);
        }
    } else {
        const int n0 = tl.s1;
        const int n01 = tl.s2;
        const long long n012 = tl.s3;
        const double *p0 = cf + ws[0] + n0 * ws[1] + (long long)n01 * ws[2] + n012 * ws[3];
#pragma unroll 1
        for (int l = 0; l < 4; ++l) {
            double sl = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                double sk = 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double *p = p0 + n0 * j + n01 * k + n012 * l;
                    double sj = 0.0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) sj = fma(p[i], b[0][i], sj);
                    sk = fma(sj, b[1][j], sk);
                }
                sl = fma(sk, b[2][k], sl);
            }
            sum = fma(sl, b[3][l], sum);
        }
    }
    return sum;
}

// One query: weights of every dimension, then the contraction.  VALUE: every nderiv is 0 (splfe).
template <int NDIM, bool VALUE>
__device__ __forceinline__ double spl_eval_point(const GridParams &gp, const DerivParams &dp,
                                                 const TableLayout &tl, const double *__restrict__ cf,
                                                 const double *xv) {
    double b[NDIM][4];
    int ws[NDIM];
#pragma unroll
    for (int d = 0; d < NDIM; ++d) {
        if (VALUE)
            spl_window_weights_value(xv[d], gp.xmin[d], gp.dx[d], gp.dxin[d], gp.nodes[d], ws[d], b[d]);
        else
            spl_window_weights(xv[d], gp.xmin[d], gp.dx[d], gp.dxin[d], gp.nodes[d], dp.nd[d], ws[d], b[d]);
    }
    double sum = spl_contract<NDIM>(tl, cf, ws, b);
    if (VALUE) {
        // a NaN coordinate fails every comparison of bascmp, so every basis value stays 0 (:253-379)
        bool isnan_q = false;
#pragma unroll
        for (int d = 0; d < NDIM; ++d) isnan_q |= (xv[d] != xv[d]);
        if (isnan_q) sum = 0.0;
    }
    return sum;
}

template <int NDIM> struct EvalCfg {
    static constexpr int THREADS = (NDIM <= 3) ? 1024 : 512;
};

// ------------------------------------------------------------------------------------------
// kernel: one query per lane; every WARP claims its own chunks of queries from the global counter, so
// after the table is staged there is no CTA-wide barrier and the 32 warps of an SM drift out of phase
// (their FP64, shared-memory and issue demands interleave instead of peaking together).
// ------------------------------------------------------------------------------------------
#define EVAL_WCHUNK 256   // queries a warp claims per atomic (8 per lane)

template <int NDIM, bool SMEM, bool VALUE>
__global__ void __launch_bounds__(EvalCfg<NDIM>::THREADS, 1)
spl_eval_kernel(const __grid_constant__ GridParams gp, const DerivParams dp,
                const real_t *__restrict__ x, int l1x, long long nq,
                const double *__restrict__ coef, long long ncol_padded, real_t *__restrict__ out,
                unsigned long long *__restrict__ chunk_counter) {
    extern __shared__ __align__(128) double s_dyn[];
    __shared__ __align__(8) uint64_t mbar;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const double *cf = coef;
    TableLayout tl;
    tl.s1 = gp.nodes[0];
    tl.s2 = gp.nodes[0] * gp.nodes[1];
    tl.s3 = (long long)tl.s2 * gp.nodes[2];
    if (SMEM) {
        if (tid == 0) {
            mbar_init(&mbar, 1);
            mbar_fence_init();
            mbar_expect_tx(&mbar, (uint32_t)(ncol_padded * sizeof(double)));
            bulk_g2s(s_dyn, coef, (uint32_t)(ncol_padded * sizeof(double)), &mbar);
        }
        __syncthreads();
        mbar_wait(&mbar, 0);
        cf = s_dyn;
    }
    for (;;) {
        unsigned long long c = 0;
        if (lane == 0) c = atomicAdd(chunk_counter, 1ULL);
        c = __shfl_sync(0xffffffffu, c, 0);
        const long long base = (long long)c * EVAL_WCHUNK;
        if (base >= nq) break;
        // software pipeline: the coordinates of the next sub-step are in flight while this one is evaluated
        double xn[NDIM];
        {
            const long long q = base + lane;
#pragma unroll
            for (int d = 0; d < NDIM; ++d) xn[d] = (q < nq) ? (double)x[q * (long long)l1x + d] : 0.0;
        }
#pragma unroll 1
        for (int sub = 0; sub < EVAL_WCHUNK / 32; ++sub) {
            const long long q = base + sub * 32 + lane;
            double xv[NDIM];
#pragma unroll
            for (int d = 0; d < NDIM; ++d) xv[d] = xn[d];
            const long long q2 = q + 32;
            if (sub + 1 < EVAL_WCHUNK / 32) {
#pragma unroll
                for (int d = 0; d < NDIM; ++d) xn[d] = (q2 < nq) ? (double)x[q2 * (long long)l1x + d] : 0.0;
            }
            if (q < nq) out[q] = (real_t)spl_eval_point<NDIM, VALUE>(gp, dp, tl, cf, xv);
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <int NDIM, bool VALUE>
static int launch_eval(const GridParams &gp, const DerivParams &dp, const real_t *d_x, int l1x,
                       long long nq, const double *d_coef, long long ncol_padded, real_t *d_out,
                       cudaStream_t stream, int nsm, size_t smem_optin, unsigned long long *d_counter) {
    constexpr int THREADS = EvalCfg<NDIM>::THREADS;
    const size_t coef_bytes = (size_t)ncol_padded * sizeof(double);
    const size_t static_reserve = 2048;
    SPL_CUDA_TRY(cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), stream));
    const bool use_smem = coef_bytes + static_reserve <= smem_optin && coef_bytes < (1u << 20);
    long long chunks = (nq + EVAL_WCHUNK - 1) / EVAL_WCHUNK;
    long long ctas = (chunks + THREADS / 32 - 1) / (THREADS / 32);
    if (ctas < 1) ctas = 1;
    if (use_smem) {
        auto kern = spl_eval_kernel<NDIM, true, VALUE>;
        SPL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)coef_bytes));
        long long grid = nsm;
        if (grid > ctas) grid = ctas;
        kern<<<(unsigned)grid, THREADS, coef_bytes, stream>>>(gp, dp, d_x, l1x, nq, d_coef, ncol_padded, d_out,
                                                              d_counter);
    } else {
        auto kern = spl_eval_kernel<NDIM, false, VALUE>;
        long long grid = (long long)nsm * (2048 / THREADS);
        if (grid > ctas) grid = ctas;
        kern<<<(unsigned)grid, THREADS, 0, stream>>>(gp, dp, d_x, l1x, nq, d_coef, ncol_padded, d_out, d_counter);
    }
    ++g_spl_launches;
    SPL_CUDA_TRY(cudaGetLastError());
    return SPLPAK_OK;
}

// d_coef: float64 device table with ncol_padded (even, >= ncol) entries; d_counter: 8-byte scratch.
int spl_eval_launch(const GridParams &gp, const int *nderiv, const real_t *d_x, int l1x, long long nq,
                    const double *d_coef, long long ncol_padded, real_t *d_out, cudaStream_t stream,
                    int nsm, size_t smem_optin, unsigned long long *d_counter) {
    DerivParams dp;
    bool value = true;
    for (int d = 0; d < SPL_MAXDIM; ++d) {
        dp.nd[d] = (nderiv && d < gp.ndim) ? nderiv[d] : 0;
        if (dp.nd[d] != 0) value = false;
    }
    if (nq <= 0) return SPLPAK_OK;
#define SPL_EVAL_CASE(N)                                                                                  \
    case N:                                                                                               \
        return value ? launch_eval<N, true>(gp, dp, d_x, l1x, nq, d_coef, ncol_padded, d_out, stream, nsm, \
                                            smem_optin, d_counter)                                        \
                     : launch_eval<N, false>(gp, dp, d_x, l1x, nq, d_coef, ncol_padded, d_out, stream, nsm, \
                                             smem_optin, d_counter);
    switch (gp.ndim) {
        SPL_EVAL_CASE(1)
        SPL_EVAL_CASE(2)
        SPL_EVAL_CASE(3)
        SPL_EVAL_CASE(4)
    }
#undef SPL_EVAL_CASE
    return SPLPAK_ERR_NDIM;
}
