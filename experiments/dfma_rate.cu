// Micro-benchmark: achievable DFMA rate for the register-tiled outer-product pattern of the accumulate
// kernel (acc[r][a] = fma(P[r], in[a], acc[r][a]), 40 accumulators) with 8 warps per SM:
//   mode 0  operands in registers
//   mode 1  operands of every point loaded from a shared-memory record ring exactly as the kernel does
//           (5 LDS.128 broadcast + 2 LDS.128 + 5 LDS.64, 4 DMUL for P, rhs FMAs), unroll 2
//   mode 2  same, with a hand-written software pipeline (next point's loads before this point's DFMAs)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_rate dfma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#define RS 46
#define PB 256
struct Ops { double in[10]; double2 ta, tb; double ha, hb, rh, rt, ri[2]; };
__device__ __forceinline__ void ld(const double *rec, int offT, int offHa, int offHb, int offRH, int offRT, int offRI, Ops &o) {
#pragma unroll
    for (int a = 0; a < 10; a += 2) { const double2 v = *reinterpret_cast<const double2 *>(rec + a); o.in[a] = v.x; o.in[a + 1] = v.y; }
    o.ta = *reinterpret_cast<const double2 *>(rec + offT);
    o.tb = *reinterpret_cast<const double2 *>(rec + offT + 2);
    o.ha = rec[offHa]; o.hb = rec[offHb]; o.rh = rec[offRH]; o.rt = rec[offRT]; o.ri[0] = rec[offRI]; o.ri[1] = rec[offRI + 1];
}
template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, const double *in_g, int iters) {
    extern __shared__ __align__(16) double s_pts[];
    for (int e = threadIdx.x; e < PB * RS; e += blockDim.x) s_pts[e] = in_g[e % 4096] + 1e-3 * e;
    __syncthreads();
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5, u = lane;
    const int uo = (u < 25 ? u : 24) * 4;
    const int offT = 16 + (uo % 10), offHa = 32 + uo / 10, offHb = 32 + (uo + 2) / 10;
    const int e0 = u * 2;
    const int offRI = 10 + (e0 & 3), offRT = 16 + 12 + ((e0 >> 2) & 3), offRH = 32 + 10 + (e0 >> 4);
    double acc[4][10], racc[2] = {0.0, 0.0};
    for (int r = 0; r < 4; ++r) for (int a = 0; a < 10; ++a) acc[r][a] = 0.0;
    if (MODE == 0) {
        double P[4], in[10];
        for (int r = 0; r < 4; ++r) P[r] = in_g[threadIdx.x + r];
        for (int a = 0; a < 10; ++a) in[a] = in_g[threadIdx.x + 4 + a];
#pragma unroll 1
        for (int it = 0; it < iters * 32; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int a = 0; a < 10; ++a) acc[r][a] = fma(P[r], in[a], acc[r][a]);
        }
    } else if (MODE == 1) {
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll 2
            for (int p = grp; p < PB; p += 8) {
                Ops o; ld(s_pts + p * RS, offT, offHa, offHb, offRH, offRT, offRI, o);
                double P[4] = {o.ha * o.ta.x, o.ha * o.ta.y, o.hb * o.tb.x, o.hb * o.tb.y};
                const double t = o.rh * o.rt;
                racc[0] = fma(t, o.ri[0], racc[0]); racc[1] = fma(t, o.ri[1], racc[1]);
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int a = 0; a < 10; ++a) acc[r][a] = fma(P[r], o.in[a], acc[r][a]);
            }
        }
    } else {
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
            Ops cur; ld(s_pts + grp * RS, offT, offHa, offHb, offRH, offRT, offRI, cur);
#pragma unroll 2
            for (int p = grp; p < PB; p += 8) {
                Ops nxt; ld(s_pts + ((p + 8) & (PB - 1)) * RS, offT, offHa, offHb, offRH, offRT, offRI, nxt);
                double P[4] = {cur.ha * cur.ta.x, cur.ha * cur.ta.y, cur.hb * cur.tb.x, cur.hb * cur.tb.y};
                const double t = cur.rh * cur.rt;
                racc[0] = fma(t, cur.ri[0], racc[0]); racc[1] = fma(t, cur.ri[1], racc[1]);
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int a = 0; a < 10; ++a) acc[r][a] = fma(P[r], cur.in[a], acc[r][a]);
                cur = nxt;
            }
        }
    }
    double s = racc[0] + racc[1];
    for (int r = 0; r < 4; ++r) for (int a = 0; a < 10; ++a) s += acc[r][a];
    if (s == 123.456) out[0] = s;
}
template <int MODE> void run(const char *name, double *d, double *in, int nsm) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 600;
    const size_t smem = PB * RS * 8;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<nsm, 256, smem>>>(d, in, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    const double pts = (double)iters * PB * nsm;     // points per launch (each warp takes PB/8 per iter)
    printf("%-28s %.3f ms  %.2f TFLOP/s (1064 FMA/point)  %.1f cycles/point/SM @1.965GHz  err=%s\n", name, best,
           2.0 * 1064 * pts / best / 1e9, best * 1e-3 * 1.965e9 / (iters * PB), cudaGetErrorString(cudaGetLastError()));
}
int main() {
    double *d, *in; cudaMalloc(&d, 64); cudaMalloc(&in, 8192 * 8); cudaMemset(in, 0, 8192 * 8);
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    run<0>("registers only", d, in, nsm);
    run<1>("smem operands, unroll 2", d, in, nsm);
    run<2>("smem operands, sw pipeline", d, in, nsm);
    return 0;
}
