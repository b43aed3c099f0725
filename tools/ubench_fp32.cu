// Microbenchmark: FP32 / packed-FP32x2 / MUFU.RSQ issue throughput on sm_100a, plus
// prototypes of the IDW station-loop body. Standalone (no torch); build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/ubench_fp32 tools/ubench_fp32.cu
// The FP32 "peak" used as the K1 roofline denominator is measured by this program on the box
// (MEASURED_PEAKS.json has no FP32 figure).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float a, float b) {
    u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ float rsq(float x) {
    float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
}

constexpr int ITERS = 4096;
constexpr int CH = 8;   // independent chains per thread

__global__ void k_ffma(float* out, float b, float c) {
    float a[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) a[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) a[i] = fmaf(a[i], b, c);
    }
    float s = 0; for (int i = 0; i < CH; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fadd(float* out, float b, float c) {
    float a[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) a[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) a[i] = __fadd_rn(a[i], b);
    }
    float s = 0; for (int i = 0; i < CH; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fmul(float* out, float b, float c) {
    float a[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) a[i] = threadIdx.x * 1e-3f + i + 1.f;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) a[i] = __fmul_rn(a[i], b);
    }
    float s = 0; for (int i = 0; i < CH; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float b, float c) {
    u64 a[CH]; u64 bb = pack2(b, b), cc = pack2(c, c);
#pragma unroll
    for (int i = 0; i < CH; i++) a[i] = pack2(threadIdx.x * 1e-3f + i, i);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) a[i] = fma2(a[i], bb, cc);
    }
    float s = 0; for (int i = 0; i < CH; i++) { float x, y; unpack2(a[i], x, y); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fadd2(float* out, float b, float c) {
    u64 a[CH]; u64 bb = pack2(b, c);
#pragma unroll
    for (int i = 0; i < CH; i++) a[i] = pack2(threadIdx.x * 1e-3f + i, i);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) a[i] = add2(a[i], bb);
    }
    float s = 0; for (int i = 0; i < CH; i++) { float x, y; unpack2(a[i], x, y); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fmul2(float* out, float b, float c) {
    u64 a[CH]; u64 bb = pack2(b, c);
#pragma unroll
    for (int i = 0; i < CH; i++) a[i] = pack2(threadIdx.x * 1e-3f + i + 1.f, i + 1.f);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) a[i] = mul2(a[i], bb);
    }
    float s = 0; for (int i = 0; i < CH; i++) { float x, y; unpack2(a[i], x, y); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_rsq(float* out, float b, float c) {
    float a[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) a[i] = threadIdx.x * 1e-3f + i + 1.f;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) a[i] = rsq(a[i]);
    }
    float s = 0; for (int i = 0; i < CH; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// rsq interleaved with FFMA: does MUFU issue overlap with the FMA pipe?
__global__ void k_rsq_ffma(float* out, float b, float c) {
    float a[CH], m[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { a[i] = threadIdx.x * 1e-3f + i + 1.f; m[i] = a[i]; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            m[i] = rsq(m[i]);
            a[i] = fmaf(a[i], b, c); a[i] = fmaf(a[i], b, c); a[i] = fmaf(a[i], b, c); a[i] = fmaf(a[i], b, c);
        }
    }
    float s = 0; for (int i = 0; i < CH; i++) s += a[i] + m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------- IDW station-loop prototypes ------------------------------------------------
// Stations in shared memory; each thread owns C cells. Result: sum w, sum w*v per cell.

// scalar: smem float4 {sx, sy, sv, 0}
template <int C>
__global__ void __launch_bounds__(256) k_idw_scalar(const float4* __restrict__ st, int S,
                                                    const float* __restrict__ cx, const float* __restrict__ cy,
                                                    float* __restrict__ out, int ncell) {
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < S; i += blockDim.x) sm[i] = st[i];
    __syncthreads();
    int base = (blockIdx.x * blockDim.x + threadIdx.x);
    int stride = gridDim.x * blockDim.x;
    float x[C], y[C], sw[C], swv[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
        int id = base + c * stride; id = id < ncell ? id : ncell - 1;
        x[c] = cx[id]; y[c] = cy[id]; sw[c] = 0.f; swv[c] = 0.f;
    }
#pragma unroll 4
    for (int s = 0; s < S; s++) {
        float4 p = sm[s];
#pragma unroll
        for (int c = 0; c < C; c++) {
            float dx = p.x - x[c], dy = p.y - y[c];
            float d2 = fmaf(dx, dx, dy * dy);
            float r = rsq(d2);
            float w = r * r * r;
            sw[c] += w;
            swv[c] = fmaf(w, p.z, swv[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < C; c++) {
        int id = base + c * stride;
        if (id < ncell) out[id] = swv[c] / sw[c];
    }
}

// packed: smem layout per station: float4 {sx, sx, sy, sy} and float2 {sv, sv}; thread owns 2*P cells
template <int P>
__global__ void __launch_bounds__(256) k_idw_packed(const float4* __restrict__ stxy, const float2* __restrict__ stv, int S,
                                                    const float* __restrict__ cx, const float* __restrict__ cy,
                                                    float* __restrict__ out, int ncell) {
    extern __shared__ float4 sm4[];
    float4* sxy = sm4;
    float2* sv = reinterpret_cast<float2*>(sm4 + S);
    for (int i = threadIdx.x; i < S; i += blockDim.x) { sxy[i] = stxy[i]; sv[i] = stv[i]; }
    __syncthreads();
    int base = (blockIdx.x * blockDim.x + threadIdx.x);
    int stride = gridDim.x * blockDim.x;
    u64 nx[P], ny[P], sw[P], swv[P];
#pragma unroll
    for (int c = 0; c < P; c++) {
        int id0 = base + (2 * c) * stride, id1 = base + (2 * c + 1) * stride;
        id0 = id0 < ncell ? id0 : ncell - 1; id1 = id1 < ncell ? id1 : ncell - 1;
        nx[c] = pack2(-cx[id0], -cx[id1]); ny[c] = pack2(-cy[id0], -cy[id1]);
        sw[c] = pack2(0.f, 0.f); swv[c] = pack2(0.f, 0.f);
    }
#pragma unroll 2
    for (int s = 0; s < S; s++) {
        float4 p = sxy[s]; float2 q = sv[s];
        u64 px = pack2(p.x, p.y), py = pack2(p.z, p.w), pv = pack2(q.x, q.y);
#pragma unroll
        for (int c = 0; c < P; c++) {
            u64 dx = add2(px, nx[c]), dy = add2(py, ny[c]);
            u64 d2 = fma2(dx, dx, mul2(dy, dy));
            float a, b; unpack2(d2, a, b);
            float ra = rsq(a), rb = rsq(b);
            u64 r = pack2(ra, rb);
            u64 w = mul2(mul2(r, r), r);
            sw[c] = add2(sw[c], w);
            swv[c] = fma2(w, pv, swv[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < P; c++) {
        int id0 = base + (2 * c) * stride, id1 = base + (2 * c + 1) * stride;
        float a, b, va, vb; unpack2(sw[c], a, b); unpack2(swv[c], va, vb);
        if (id0 < ncell) out[id0] = va / a;
        if (id1 < ncell) out[id1] = vb / b;
    }
}

template <typename F>
static float time_ms(F f, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) f();
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms / reps;
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    int sms = pr.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", pr.name, sms, pr.clockRate);
    int blocks = sms * 8, threads = 256;
    float* out; CK(cudaMalloc(&out, sizeof(float) * blocks * threads));
    double lanes = double(blocks) * threads * double(ITERS) * CH;   // lane-instructions per launch
    struct { const char* name; void (*k)(float*, float, float); double flop_per_inst; double extra; } tests[] = {
        {"ffma", k_ffma, 2, 1}, {"fadd", k_fadd, 1, 1}, {"fmul", k_fmul, 1, 1},
        {"ffma2", k_ffma2, 4, 1}, {"fadd2", k_fadd2, 2, 1}, {"fmul2", k_fmul2, 2, 1},
        {"mufu_rsq", k_rsq, 1, 1}, {"rsq+4ffma", k_rsq_ffma, 9, 1},
    };
    for (auto& t : tests) {
        float ms = time_ms([&] { t.k<<<blocks, threads>>>(out, 1.0001f, 1e-7f); }, 10);
        CK(cudaGetLastError());
        double inst_per_s = lanes / (ms * 1e-3);
        printf("{\"ubench\": \"%s\", \"ms\": %.4f, \"Glane_inst_per_s\": %.1f, \"lane_inst_per_clk_per_sm_at_1965\": %.2f, \"TFLOPs\": %.2f}\n",
               t.name, ms, inst_per_s * 1e-9, inst_per_s / (sms * 1.965e9), inst_per_s * t.flop_per_inst * 1e-12);
    }

    // IDW prototypes
    const int S = 512; const int ncell = 1 << 22;
    std::vector<float4> st(S), stxy(S); std::vector<float2> stv(S);
    std::vector<float> cx(ncell), cy(ncell);
    srand(1);
    for (int i = 0; i < S; i++) {
        float x = 500000.f + 200000.f * (rand() / (float)RAND_MAX), y = 4800000.f + 200000.f * (rand() / (float)RAND_MAX);
        float v = -3.f + 6.f * (rand() / (float)RAND_MAX);
        st[i] = make_float4(x, y, v, 0.f); stxy[i] = make_float4(x, x, y, y); stv[i] = make_float2(v, v);
    }
    for (int i = 0; i < ncell; i++) { cx[i] = 500000.5f + 100.f * (i % 2048); cy[i] = 4800000.f + 100.f * (i / 2048) - 0.5f; }
    float4 *dst, *dstxy; float2* dstv; float *dcx, *dcy, *dout, *dout2;
    CK(cudaMalloc(&dst, S * 16)); CK(cudaMalloc(&dstxy, S * 16)); CK(cudaMalloc(&dstv, S * 8));
    CK(cudaMalloc(&dcx, ncell * 4)); CK(cudaMalloc(&dcy, ncell * 4)); CK(cudaMalloc(&dout, ncell * 4)); CK(cudaMalloc(&dout2, ncell * 4));
    CK(cudaMemcpy(dst, st.data(), S * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dstxy, stxy.data(), S * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dstv, stv.data(), S * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dcx, cx.data(), ncell * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dcy, cy.data(), ncell * 4, cudaMemcpyHostToDevice));
    double pairs = double(ncell) * S;
    auto report = [&](const char* name, float ms) {
        printf("{\"idw_proto\": \"%s\", \"ms\": %.4f, \"Gpairs_per_s\": %.1f, \"frac_of_4650\": %.3f}\n", name, ms, pairs / (ms * 1e-3) * 1e-9,
               pairs / (ms * 1e-3) / 4.65e12);
    };
#define RUN_SCALAR(C) { int g = (ncell + 256 * C - 1) / (256 * C); \
        float ms = time_ms([&] { k_idw_scalar<C><<<g, 256, S * 16>>>(dst, S, dcx, dcy, dout, ncell); }, 5); CK(cudaGetLastError()); report("scalar_C" #C, ms); }
#define RUN_PACKED(P) { int g = (ncell + 256 * 2 * P - 1) / (256 * 2 * P); \
        float ms = time_ms([&] { k_idw_packed<P><<<g, 256, S * 24>>>(dstxy, dstv, S, dcx, dcy, dout2, ncell); }, 5); CK(cudaGetLastError()); report("packed_P" #P, ms); }
    RUN_SCALAR(1) RUN_SCALAR(2) RUN_SCALAR(4) RUN_SCALAR(8)
    RUN_PACKED(1) RUN_PACKED(2) RUN_PACKED(4)
    // cross-check scalar vs packed vs host double on a few cells
    std::vector<float> h1(ncell), h2(ncell);
    { int g = (ncell + 256 * 4 - 1) / (256 * 4); k_idw_scalar<4><<<g, 256, S * 16>>>(dst, S, dcx, dcy, dout, ncell); }
    { int g = (ncell + 256 * 4 - 1) / (256 * 4); k_idw_packed<2><<<g, 256, S * 24>>>(dstxy, dstv, S, dcx, dcy, dout2, ncell); }
    CK(cudaMemcpy(h1.data(), dout, ncell * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h2.data(), dout2, ncell * 4, cudaMemcpyDeviceToHost));
    double e1 = 0, e2 = 0;
    for (int k = 0; k < 2000; k++) {
        int i = (int)((long long)k * 2097 % ncell);
        double sw = 0, swv = 0;
        for (int s = 0; s < S; s++) {
            float dx = st[s].x - cx[i], dy = st[s].y - cy[i];
            float d = sqrtf(dx * dx + dy * dy);
            if (d > 0) { double dk = double(d) / 10000.; double w = 1. / (dk * dk * dk); sw += w; swv += w * st[s].z; }
        }
        double ref = swv / sw;
        e1 = fmax(e1, fabs(h1[i] - ref)); e2 = fmax(e2, fabs(h2[i] - ref));
    }
    printf("{\"idw_proto_maxabs_vs_f64\": {\"scalar\": %.3g, \"packed\": %.3g}}\n", e1, e2);
    return 0;
}
