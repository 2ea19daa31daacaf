// common.cuh -- shared definitions for the sm_100a kernels of splpak_b200.
//
// Data layout in HBM:
//   points      AoS as Fortran passes them: x(l1x, n) + y(n) [+ w(n)], splpak_real
//   records     window-sorted AoS scratch: (ndim+2) float64 per point = x[0..ndim-1], y, w
//   S           "orthant stencil" storage of the Gram matrix G = sum (w phi)(w phi)^T:
//               S[node * 4^ndim + sum_d delta_d 4^d], delta_d = |col_d - row_d| in 0..3, node_d =
//               min(row_d, col_d).  G is symmetric under swapping row_d <-> col_d in every
//               dimension separately (tensor-product basis), so 4^ndim values per node hold all of it.
//   AB          lower band storage of G for the Cholesky: element (i, j), i >= j, at AB[i + j*lda]
//   coef        float64 on the device, dimension 1 fastest (src/splpak.F90:661-666)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/splpak_b200.h"

typedef splpak_real real_t;

#define SPL_MAXDIM 4
#define SPL_NSM_DEFAULT 148
#define SPL_FAIL_WORDS 256      // ints behind a handle's failure flag: [0] flag, [1] barrier counter, [32..] data-flow flags

struct GridParams {
    int ndim;
    int nodes[SPL_MAXDIM];
    int nwin[SPL_MAXDIM];          // windows per dimension = nodes - 3
    double xmin[SPL_MAXDIM];
    double dx[SPL_MAXDIM];         // (xmax-xmin)/(nodes-1), src/splpak.F90:747
    double dxin[SPL_MAXDIM];       // 1/dx, :748
    long long ncol;                // product of nodes
    long long nwindows;            // product of nwin
    int nsten;                     // 4^ndim
};

// 10 symmetric pairs (i <= j) of the 4 window-local node indices of one dimension.
__host__ __device__ __forceinline__ void spl_pair(int a, int &i, int &j) {
    // a: 0..9 -> (0,0)(0,1)(0,2)(0,3)(1,1)(1,2)(1,3)(2,2)(2,3)(3,3)
    const int pi = (a >= 4) + (a >= 7) + (a >= 9);
    const int start = (pi == 0) ? 0 : (pi == 1) ? 4 : (pi == 2) ? 7 : 9;
    i = pi;
    j = pi + (a - start);
}

__host__ __device__ constexpr int spl_ipow(int b, int e) { return e <= 0 ? 1 : b * spl_ipow(b, e - 1); }

// Fixed-point nearest-node histogram (assemble.cu: spl_classify_kernel): two int64 limbs per node (+ one pair for
// totlwt), the chunk's scale (qscale, lsb exponent) and the bit pattern of its largest |w|.
struct HistScratch {
    unsigned long long *hq = nullptr;
    double *qparams = nullptr;
    unsigned long long *wmax = nullptr;
};

// Scratch of the chunk pipeline of assemble.cu (owned by a fit handle, sized by spl_assemble_scratch_*).
struct AssembleScratch {
    unsigned *wincount, *winstart, *wincursor, *itemstart, *item_win, *item_seg, *meta;
    unsigned *perm;
    long long max_items;
    long long nbins;      // bins of the counting sort: windows (direct path) or cells (moment path)
    int moments;          // 1: 3-D assembly by cell moments (moments.cuh)
    double *celltab;      // moment path: per-(dimension, cell) coefficient tables
    double *cellmom;      // moment path: per-cell moment sums of the chunk in flight (kept zero between chunks)
    int cursor_stride;    // 4-byte words between the per-bin cursors of the second binning pass
    HistScratch hist;     // fixed-point histogram scratch
    double *yw;           // moment path: interleaved (y, w) copy of the chunk, 2 doubles per point (sized with perm)
};

// Scratch of the orthogonal fit path (ortho.cuh), owned by a fit handle.
struct OrthoScratch {
    unsigned *wincount = nullptr, *winstart = nullptr, *wincursor = nullptr, *itemstart = nullptr, *meta = nullptr;
    unsigned *perm = nullptr;
    long long perm_cap = 0;
    HistScratch hist;              // fixed-point histogram scratch
    double *Rw = nullptr;          // nwindows x ncw x (ncw + 1)
    unsigned *blist = nullptr;     // non-zero stage-2 blocks, in order; meta2[0] = their number
    unsigned *meta2 = nullptr;
    double *Rb = nullptr;          // ncol x (bw + 2): band row + transformed rhs
    long long *progress = nullptr; // pipeline flags, one per block
    double *csol = nullptr;        // ncol
    int ncw = 0, nrb = 0, nblk_win = 0, bw = 0;
    long long max_blocks = 0;
    int ready = 0;
};

extern unsigned long long g_spl_launches;   // host-side launch counter (capi.cu)

#define SPL_CUDA_TRY(expr)                                                              \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            fprintf(stderr, "splpak_b200: CUDA error %s at %s:%d (%s)\n",               \
                    cudaGetErrorString(_e), __FILE__, __LINE__, #expr);                 \
            return SPLPAK_ERR_CUDA;                                                     \
        }                                                                               \
    } while (0)

static inline int spl_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
