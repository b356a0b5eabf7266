// peak_fp64.cu — FP64 roofline denominators measured on the device (MEASURED_PEAKS.json has no FP64 entry).
//   DFMA: independent fused multiply-add chains on the CUDA cores (8 chains per thread, 1024 threads per SM).
//   DMMA: mma.sync.aligned.m8n8k4.row.col.f64 chains (tcgen05 has no FP64 kind; on sm_100a FP64 tensor work is mma.sync).
#include "../../include/lexls_b200.h"

#include <cuda_runtime.h>
#include <string>

__global__ void __launch_bounds__(256) dfma_chain_kernel(double *out, int iters, double a, double b)
{
    double r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
    for (int i = 0; i < iters; i++)
    {
#pragma unroll
        for (int u = 0; u < 8; u++)
        {
            r0 = fma(r0, a, b);
            r1 = fma(r1, a, b);
            r2 = fma(r2, a, b);
            r3 = fma(r3, a, b);
            r4 = fma(r4, a, b);
            r5 = fma(r5, a, b);
            r6 = fma(r6, a, b);
            r7 = fma(r7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
}

__global__ void __launch_bounds__(256) dmma_chain_kernel(double *out, int iters, double a, double b)
{
    double c0[2] = {0.0, 0.0}, c1[2] = {1.0, 0.0}, c2[2] = {2.0, 0.0}, c3[2] = {3.0, 0.0};
    const double fa = a + threadIdx.x * 1e-9, fb = b;
    for (int i = 0; i < iters; i++)
    {
#pragma unroll
        for (int u = 0; u < 4; u++)
        {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(fa), "d"(fb));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(fa), "d"(fb));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(fa), "d"(fb));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(fa), "d"(fb));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (c0[0] + c0[1]) + (c1[0] + c1[1]) + (c2[0] + c2[1]) + (c3[0] + c3[1]);
}

extern "C" int lexls_b200_measure_fp64_peak(int device, double *dfma_tflops, double *dmma_tflops, double *sm_clock_mhz)
{
    if (cudaSetDevice(device) != cudaSuccess) return LEXLS_B200_ECUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return LEXLS_B200_ECUDA;
    const int sms = prop.multiProcessorCount, blocks = sms * 4, threads = 256;
    double *out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return LEXLS_B200_ECUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best_f = 1e30f, best_m = 1e30f;
    const int iters = 4096;
    for (int rep = 0; rep < 6; rep++)
    {
        float ms;
        cudaEventRecord(e0);
        dfma_chain_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best_f) best_f = ms;
        cudaEventRecord(e0);
        dmma_chain_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best_m) best_m = ms;
    }
    cudaError_t err = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (err != cudaSuccess) return LEXLS_B200_ECUDA;
    const double fl_f = 2.0 * 64.0 * iters * (double)blocks * threads;                  // 64 FMAs per thread per iteration
    const double fl_m = 2.0 * 8 * 8 * 4 * 16.0 * iters * (double)blocks * (threads / 32); // 16 mma (8x8x4) per warp per iteration
    if (dfma_tflops) *dfma_tflops = fl_f / (best_f * 1e-3) / 1e12;
    if (dmma_tflops) *dmma_tflops = fl_m / (best_m * 1e-3) / 1e12;
    if (sm_clock_mhz) *sm_clock_mhz = prop.clockRate / 1000.0;
    return 0;
}
