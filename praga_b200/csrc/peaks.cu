// peaks.cu — live measurement of the two pipe ceilings the K1 roofline is reported against (bench.py):
// FP32 FMA throughput (FFMA2 chain) and MUFU.RSQ throughput. MEASURED_PEAKS.json only carries HBM and bf16 GEMM
// numbers, so the FP32 denominator is measured on the box the bench runs on.
#include "common.cuh"

namespace pb {

constexpr int kIters = 4096;
constexpr int kChains = 8;

__global__ void __launch_bounds__(256) ffma2_chain_kernel(float* out, float b, float c)
{
    f32x2 a[kChains];
    const f32x2 bb = pack2(b, b), cc = pack2(c, c);
#pragma unroll
    for (int i = 0; i < kChains; i++) a[i] = pack2(threadIdx.x * 1e-3f + i, float(i));
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < kChains; i++) a[i] = fma2(a[i], bb, cc);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; i++) { float x, y; unpack2(a[i], x, y); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) rsq_chain_kernel(float* out)
{
    float a[kChains];
#pragma unroll
    for (int i = 0; i < kChains; i++) a[i] = threadIdx.x * 1e-3f + i + 1.f;
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < kChains; i++) a[i] = rsqrt_approx(a[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dfma_chain_kernel(double* out, double b, double c)
{
    double a[kChains];
#pragma unroll
    for (int i = 0; i < kChains; i++) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < kIters / 4; it++) {
#pragma unroll
        for (int i = 0; i < kChains; i++) a[i] = fma(a[i], b, c);
    }
    double s = 0.;
#pragma unroll
    for (int i = 0; i < kChains; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FP64 FMA ceiling (TFLOP/s, DFMA chain): the denominator of the K2 (local detrending) and K5-estimate rooflines
cudaError_t measureFp64Peak(cudaStream_t s, int sms, float* fp64Tflops)
{
    const int blocks = sms * 8, threads = 256;
    double* out = nullptr;
    cudaError_t e = cudaMalloc(&out, sizeof(double) * blocks * threads);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0, s);
        dfma_chain_kernel<<<blocks, threads, 0, s>>>(out, 1.0000001, 1e-9);
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    const double lanes = double(blocks) * threads * double(kIters / 4) * kChains;
    *fp64Tflops = float(lanes * 2.0 / (best * 1e-3) * 1e-12);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return cudaGetLastError();
}

// returns best-of-`reps` TFLOP/s (FMA = 2 flop) and G MUFU.RSQ/s
cudaError_t measurePipePeaks(cudaStream_t s, int sms, float* fp32Tflops, float* mufuGops)
{
    const int blocks = sms * 8, threads = 256;
    float* out = nullptr;
    cudaError_t e = cudaMalloc(&out, sizeof(float) * blocks * threads);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float bestF = 1e30f, bestR = 1e30f;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(e0, s);
        ffma2_chain_kernel<<<blocks, threads, 0, s>>>(out, 1.0001f, 1e-7f);
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < bestF) bestF = ms;
        cudaEventRecord(e0, s);
        rsq_chain_kernel<<<blocks, threads, 0, s>>>(out);
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < bestR) bestR = ms;
    }
    const double lanes = double(blocks) * threads * double(kIters) * kChains;
    *fp32Tflops = float(lanes * 4.0 / (bestF * 1e-3) * 1e-12);
    *mufuGops = float(lanes / (bestR * 1e-3) * 1e-9);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return cudaGetLastError();
}

}  // namespace pb
