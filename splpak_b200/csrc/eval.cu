// eval.cu -- batched splfe / splde (src/splpak.F90:1089-1275) for sm_100a.
//
// One query per thread, one persistent CTA per SM.  The whole coefficient table (<= 166 KB at the
// named configs) is staged once per CTA into shared memory with 1-D bulk async copies
// (cp.async.bulk -> SASS UBLKCP) signalled through an mbarrier.  CTAs pull tiles of queries from a
// global atomic counter (dynamic scheduling: the launch never waits for one slow SM).  Tables that
// do not fit in shared memory are gathered through the read-only path from L2.
//
// Per query: four 1-D weights per dimension (basis.cuh: same formulas and index box as the
// reference), then the nested contraction  sum_k b3[k] sum_j b2[j] sum_i coef[..]*b1[i]  with
// dimension 1 innermost (4 contiguous coefficients).
//
// The gather of the 4^ndim coefficients is what binds this kernel for scattered queries: 64
// 8-byte shared-memory loads per 3-D query at random addresses cost ~6 wavefronts each through
// bank conflicts (measured: 12.8 wavefronts per query, 69 % of them conflict replays).  All 4^ndim
// loads of a lane use the SAME offsets from the lane's window base, so a half-warp is conflict-free
// for every one of them as soon as its 16 window bases fall into 16 different 8-byte banks.  Each
// CTA therefore counting-sorts its tile by (window base mod 16) and deals the sorted queries round
// robin to the half-warps before evaluating; results go back to the queries' original slots.
//
// Algorithmic HBM traffic: (ndim + 1) reals per query.
#include "basis.cuh"

struct DerivParams {
    int nd[SPL_MAXDIM];
};

#define EVAL_SUBTILES 4   // a CTA claims THREADS*EVAL_SUBTILES queries per atomic

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

// Nested contraction over the 4^NDIM window, dimension 1 innermost.
template <int NDIM>
__device__ __forceinline__ double spl_contract(const GridParams &gp, const double *__restrict__ cf,
                                               const int *ws, const double (*b)[4]) {
    double sum = 0.0;
    if constexpr (NDIM == 1) {
        const double *p = cf + ws[0];
#pragma unroll
        for (int i = 0; i < 4; ++i) sum = fma(p[i], b[0][i], sum);
    } else if constexpr (NDIM == 2) {
        const int n0 = gp.nodes[0];
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
        const int n0 = gp.nodes[0];
        const int n01 = n0 * gp.nodes[1];
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
        const int n0 = gp.nodes[0];
        const int n01 = n0 * gp.nodes[1];
        const long long n012 = (long long)n01 * gp.nodes[2];
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

template <int NDIM> struct EvalCfg {
    static constexpr int THREADS = (NDIM <= 3) ? 1024 : 512;
};

// VALUE: every nderiv is 0 (splfe) -> branch-free value basis.  SMEM: table staged in shared memory.
template <int NDIM, bool SMEM, bool VALUE>
__global__ void __launch_bounds__(EvalCfg<NDIM>::THREADS, 1)
spl_eval_kernel(const __grid_constant__ GridParams gp, const DerivParams dp,
                const real_t *__restrict__ x, int l1x, long long nq,
                const double *__restrict__ coef, long long ncol_padded, real_t *__restrict__ out,
                unsigned long long *__restrict__ tile_counter) {
    constexpr int THREADS = EvalCfg<NDIM>::THREADS;
    constexpr int NW = THREADS / 32;          // warps
    constexpr int NHW = THREADS / 16;         // half-warps
    constexpr bool PERMUTE = SMEM && NDIM >= 2;
    extern __shared__ __align__(128) double s_dyn[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ unsigned long long s_tile[2];
    __shared__ int s_cnt[PERMUTE ? 16 * NW : 1];
    double *s_coef = s_dyn;
    double *s_xq = s_dyn + (SMEM ? ncol_padded : 0);                  // THREADS * NDIM
    unsigned short *s_src = reinterpret_cast<unsigned short *>(s_xq + THREADS * NDIM);

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const double *cf = coef;
    if (tid == 0) s_tile[0] = atomicAdd(tile_counter, 1ULL);
    if (SMEM) {
        bulk_stage(s_coef, coef, (uint32_t)(ncol_padded * sizeof(double)), &mbar);
        cf = s_coef;
    }
    __syncthreads();

    const long long tile_q = (long long)THREADS * EVAL_SUBTILES;
    for (int it = 0;; ++it) {
        const unsigned long long tile = s_tile[it & 1];
        const long long base0 = (long long)tile * tile_q;
        if (base0 >= nq) break;
        unsigned long long next_tile = 0;
        if (tid == 0) next_tile = atomicAdd(tile_counter, 1ULL);      // prefetch; stored at the end

#pragma unroll 1
        for (int sub = 0; sub < EVAL_SUBTILES; ++sub) {
            const long long base = base0 + (long long)sub * THREADS;
            if (base >= nq) break;
            const long long q = base + tid;
            const bool valid = q < nq;
            double xv[NDIM];
#pragma unroll
            for (int d = 0; d < NDIM; ++d) xv[d] = valid ? (double)x[q * (long long)l1x + d] : gp.xmin[d];
            int src = tid;

            if (PERMUTE) {
                // bank key of the window base: (linear index of the window's first node) mod 16
                int key;
                {
                    int lin = 0, stride = 1;
#pragma unroll
                    for (int d = 0; d < NDIM; ++d) {
                        int wsd, ibmn, ibmx;
                        spl_box(xv[d], gp.xmin[d], gp.dxin[d], gp.nodes[d], wsd, ibmn, ibmx);
                        lin += wsd * stride;
                        stride *= gp.nodes[d];
                    }
                    key = lin & 15;
                }
                for (int e = tid; e < 16 * NW; e += THREADS) s_cnt[e] = 0;
                __syncthreads();
                const unsigned peers = __match_any_sync(0xffffffffu, key);
                const int rank = __popc(peers & ((1u << lane) - 1u));
                if (rank == 0) s_cnt[key * NW + warp] = __popc(peers);
                __syncthreads();
                if (warp == 0) {
                    // exclusive scan of the 16*NW counts (key-major): lane owns NW/2 consecutive entries
                    constexpr int PER = 16 * NW / 32;
                    int v[PER], tot = 0;
#pragma unroll
                    for (int e = 0; e < PER; ++e) {
                        v[e] = s_cnt[lane * PER + e];
                        tot += v[e];
                    }
                    int incl = tot;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int nbr = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += nbr;
                    }
                    int run = incl - tot;
#pragma unroll
                    for (int e = 0; e < PER; ++e) {
                        s_cnt[lane * PER + e] = run;
                        run += v[e];
                    }
                }
                __syncthreads();
                const int p = s_cnt[key * NW + warp] + rank;          // position in key-sorted order
                const int slot = (p % NHW) * 16 + p / NHW;            // deal round robin to half-warps
#pragma unroll
                for (int d = 0; d < NDIM; ++d) s_xq[slot * NDIM + d] = xv[d];
                s_src[slot] = (unsigned short)tid;
                __syncthreads();
#pragma unroll
                for (int d = 0; d < NDIM; ++d) xv[d] = s_xq[tid * NDIM + d];
                src = s_src[tid];
            }

            double b[NDIM][4];
            int ws[NDIM];
#pragma unroll
            for (int d = 0; d < NDIM; ++d) {
                if (VALUE)
                    spl_window_weights_value(xv[d], gp.xmin[d], gp.dx[d], gp.dxin[d], gp.nodes[d], ws[d], b[d]);
                else
                    spl_window_weights(xv[d], gp.xmin[d], gp.dx[d], gp.dxin[d], gp.nodes[d], dp.nd[d], ws[d], b[d]);
            }
            const double sum = spl_contract<NDIM>(gp, cf, ws, b);
            const long long qo = base + src;
            if (qo < nq) out[qo] = (real_t)sum;
        }
        __syncthreads();
        if (tid == 0) s_tile[(it + 1) & 1] = next_tile;
        __syncthreads();
    }
}

template <int NDIM, bool VALUE>
static int launch_eval(const GridParams &gp, const DerivParams &dp, const real_t *d_x, int l1x,
                       long long nq, const double *d_coef, long long ncol_padded, real_t *d_out,
                       cudaStream_t stream, int nsm, size_t smem_optin, unsigned long long *d_counter) {
    constexpr int THREADS = EvalCfg<NDIM>::THREADS;
    const size_t coef_bytes = (size_t)ncol_padded * sizeof(double);
    const size_t xchg_bytes = (size_t)THREADS * NDIM * sizeof(double) + (size_t)THREADS * sizeof(unsigned short);
    const bool use_smem = coef_bytes + xchg_bytes + 2048 <= smem_optin && coef_bytes < (1u << 20);
    const long long tile_q = (long long)THREADS * EVAL_SUBTILES;
    long long tiles = (nq + tile_q - 1) / tile_q;
    if (tiles < 1) tiles = 1;
    SPL_CUDA_TRY(cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), stream));
    if (use_smem) {
        auto kern = spl_eval_kernel<NDIM, true, VALUE>;
        const size_t smem = coef_bytes + xchg_bytes;
        SPL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SPL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          cudaSharedmemCarveoutMaxShared));
        long long grid = nsm;
        if (grid > tiles) grid = tiles;
        kern<<<(unsigned)grid, THREADS, smem, stream>>>(gp, dp, d_x, l1x, nq, d_coef, ncol_padded, d_out, d_counter);
    } else {
        auto kern = spl_eval_kernel<NDIM, false, VALUE>;
        long long grid = (long long)nsm * (2048 / THREADS);
        if (grid > tiles) grid = tiles;
        kern<<<(unsigned)grid, THREADS, xchg_bytes, stream>>>(gp, dp, d_x, l1x, nq, d_coef, ncol_padded, d_out,
                                                             d_counter);
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
