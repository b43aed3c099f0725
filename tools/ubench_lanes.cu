// ubench_lanes.cu — latency of a small dependent chain (H2D, kernel, D2H, synchronise) issued by N host threads on N streams of
// one context: how much do the lanes of the batch drivers slow each other down on the host side, with pageable and with
// page-locked host buffers?   nvcc -O2 -arch=sm_100a -o tools/ubench_lanes tools/ubench_lanes.cu && tools/ubench_lanes
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

__global__ void spin_kernel(float* p, int iters)
{
    float v = p[threadIdx.x];
    for (int i = 0; i < iters; i++) v = v * 1.0001f + 0.5f;
    p[threadIdx.x] = v;
}

static double run(int nThreads, bool pinned, int chain, int iters)
{
    std::vector<double> us(nThreads, 0.0);
    std::vector<std::thread> th;
    for (int t = 0; t < nThreads; t++)
        th.emplace_back([&, t] {
            cudaSetDevice(0);
            cudaStream_t s;
            cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
            float *d, *h;
            cudaMalloc(&d, 4096);
            if (pinned) cudaMallocHost(&h, 4096);
            else h = (float*)malloc(4096);
            for (int i = 0; i < 1024; i++) h[i] = 1.f;
            const int reps = 200;
            for (int w = 0; w < 2; w++) {
                auto t0 = std::chrono::steady_clock::now();
                for (int r = 0; r < reps; r++)
                    for (int c = 0; c < chain; c++) {
                        cudaMemcpyAsync(d, h, 2048, cudaMemcpyHostToDevice, s);
                        spin_kernel<<<1, 256, 0, s>>>(d, iters);
                        cudaMemcpyAsync(h, d, 2048, cudaMemcpyDeviceToHost, s);
                        cudaStreamSynchronize(s);
                    }
                us[t] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / (reps * chain);
            }
            cudaStreamDestroy(s);
        });
    for (auto& x : th) x.join();
    double m = 0;
    for (double v : us) m += v;
    return m / nThreads;
}

int main()
{
    cudaFree(0);
    for (int iters : {2000, 40000})
        for (int pinned = 0; pinned < 2; pinned++)
            for (int n : {1, 2, 4, 8}) {
                const double v = run(n, pinned != 0, 4, iters);
                printf("{\"kernel_iters\": %d, \"host_buffers\": \"%s\", \"threads\": %d, \"us_per_link\": %.1f}\n", iters, pinned ? "pinned" : "pageable", n, v);
            }
    return 0;
}
