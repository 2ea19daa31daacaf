// peaks.cu -- FP64 roofline denominators measured on the device the library runs on.
// MEASURED_PEAKS.json (driver-written) has HBM and bf16 only; the assembly and solve kernels are
// bound by the FP64 pipe, so bench.py reports their fraction against these two numbers:
//   DFMA  : register-resident fused multiply-add chains, 8 independent chains per thread
//   DMMA  : mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4), 8 independent accumulators per warp
// plus a device copy bandwidth cross-check of hbm_gbs.
#include "common.cuh"

#define PEAK_ITERS 4096

__global__ void __launch_bounds__(256) spl_peak_dfma_kernel(double *out, double a, double b) {
    double r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = (double)(threadIdx.x + k);
#pragma unroll 1
    for (int it = 0; it < PEAK_ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = fma(r[k], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += r[k];
    if (s == 123.456) out[0] = s;   // keep the chains alive
}

__global__ void __launch_bounds__(256) spl_peak_dmma_kernel(double *out, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int k = 0; k < 8; ++k) c[k][0] = c[k][1] = 0.0;
    const double av = a + threadIdx.x, bv = b + threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < PEAK_ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[k][0]), "+d"(c[k][1])
                         : "d"(av), "d"(bv));
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) spl_peak_copy_kernel(const double2 *__restrict__ in,
                                                            double2 *__restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i];
}

int spl_measure_peaks_impl(double *out, int n) {
    int dev = 0, count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) {
        cudaGetLastError();
        return SPLPAK_ERR_CUDA;
    }
    SPL_CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp p;
    SPL_CUDA_TRY(cudaGetDeviceProperties(&p, dev));
    const int nsm = p.multiProcessorCount;
    double *d = nullptr;
    SPL_CUDA_TRY(cudaMalloc((void **)&d, 64));
    cudaEvent_t e0, e1;
    SPL_CUDA_TRY(cudaEventCreate(&e0));
    SPL_CUDA_TRY(cudaEventCreate(&e1));
    const int grid = nsm * 8;
    float best_fma = 1e30f, best_mma = 1e30f, best_cp = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        spl_peak_dfma_kernel<<<grid, 256>>>(d, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best_fma) best_fma = ms;
        cudaEventRecord(e0);
        spl_peak_dmma_kernel<<<grid, 256>>>(d, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best_mma) best_mma = ms;
        g_spl_launches += 2;
    }
    const double fma_flops = 2.0 * 8.0 * PEAK_ITERS * 256.0 * grid;
    const double mma_flops = 2.0 * 256.0 * 8.0 * PEAK_ITERS * 8.0 * grid;   // 8 warps/CTA, 256 FMA per MMA
    // copy: 1 GiB read + 1 GiB write
    const long long nelem = (1LL << 30) / sizeof(double2);
    double2 *src = nullptr, *dst = nullptr;
    if (cudaMalloc((void **)&src, 1LL << 30) == cudaSuccess && cudaMalloc((void **)&dst, 1LL << 30) == cudaSuccess) {
        cudaMemset(src, 1, 1LL << 30);
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            spl_peak_copy_kernel<<<nsm * 16, 256>>>(src, dst, nelem);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best_cp) best_cp = ms;
            ++g_spl_launches;
        }
    } else {
        cudaGetLastError();
    }
    if (src) cudaFree(src);
    if (dst) cudaFree(dst);
    cudaFree(d);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    SPL_CUDA_TRY(cudaGetLastError());
    if (n > 0) out[0] = fma_flops / (best_fma * 1e-3) * 1e-12;
    if (n > 1) out[1] = mma_flops / (best_mma * 1e-3) * 1e-12;
    if (n > 2) out[2] = (best_cp < 1e29f) ? 2.0 * (double)(1LL << 30) / (best_cp * 1e-3) * 1e-9 : 0.0;
    return SPLPAK_OK;
}
