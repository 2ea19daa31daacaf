// DMMA.8x8x4 issue rate with (a) constant operand registers (the peaks.cu measurement), (b) the update tile's operand
// pattern: 2 A x 4 B distinct registers per k-step, rotating over 4 k-steps, (c) the same with the operands loaded from
// shared memory every k-step (the tile's inner loop without its global traffic and barriers).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_rate dmma_rate.cu && ./dmma_rate
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, double a0, double b0) {
    __shared__ double s[2 * 32 * 68];
    for (int i = threadIdx.x; i < 2 * 32 * 68; i += 256) s[i] = a0 + i * 1e-9;
    __syncthreads();
    double c[2][4][2];
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 4; ++j) c[i][j][0] = c[i][j][1] = 0.0;
    const int lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3, warp = threadIdx.x >> 5;
    const int wI = (warp & 1) * 32, wJ = (warp >> 1) * 16;
    double A[4][2], B[4][4];
    for (int q = 0; q < 4; ++q) { for (int i = 0; i < 2; ++i) A[q][i] = a0 + threadIdx.x + q + 4 * i; for (int j = 0; j < 4; ++j) B[q][j] = b0 + threadIdx.x * 3 + q + 4 * j; }
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double a[2], b[4];
            if (MODE == 0) { a[0] = a[1] = A[0][0]; b[0] = b[1] = b[2] = b[3] = B[0][0]; }
            if (MODE == 1) { a[0] = A[q][0]; a[1] = A[q][1]; for (int j = 0; j < 4; ++j) b[j] = B[q][j]; }
            if (MODE == 2) {
                const int k0 = ((it & 1) * 4 + q) * 4;
                const double *sA = s, *sB = s + 32 * 68;
                for (int i = 0; i < 2; ++i) a[i] = -sB[(k0 + t4) * 68 + wJ + i * 8 + g];
                for (int j = 0; j < 4; ++j) b[j] = sA[(k0 + t4) * 68 + wI + j * 8 + g];
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(c[i][j][0], c[i][j][1], a[i], b[j]);
        }
    }
    double sum = 0;
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 4; ++j) sum += c[i][j][0] + c[i][j][1];
    if (sum == 123.456) out[0] = sum;
}
template <int MODE> void run(const char *name, int ctas_per_sm) {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    double *d; cudaMalloc(&d, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0); k<MODE><<<nsm * ctas_per_sm, 256>>>(d, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    const double flops = 2.0 * 256 * 32.0 * ITERS * 8 * nsm * ctas_per_sm;
    printf("%-28s %d CTAs/SM: %.3f ms  %.2f TFLOP/s\n", name, ctas_per_sm, best, flops / best / 1e9);
    cudaFree(d);
}
int main() {
    for (int c : {1, 2, 3, 6}) { run<0>("constant operands", c); run<1>("2A x 4B rotating registers", c); run<2>("operands from shared memory", c); }
    return 0;
}
