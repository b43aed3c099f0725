// common.cuh — shared device/host declarations of the praga_b200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/praga_b200.h"

// Device-side capacity checks (tile / scratch / ring sizes), compiled in with -DPB_DEBUG_ASSERTS (PB_DEBUG=1 python -m
// praga_b200.build --force): compute-sanitizer is not available on the GPU pool, so the engine carries its own bounds checks.
// A failed check prints the site and traps: the launch ends with an error that the C ABI reports as PB_ERR_CUDA.
#ifdef PB_DEBUG_ASSERTS
#include <stdio.h>
#define PB_DEV_ASSERT(cond)                                                                                          \
    do {                                                                                                             \
        if (!(cond)) {                                                                                               \
            printf("praga_b200 device assert: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__, blockIdx.x, threadIdx.x); \
            __trap();                                                                                                \
        }                                                                                                            \
    } while (0)
#else
#define PB_DEV_ASSERT(cond) ((void)0)
#endif

namespace pb {

constexpr float kNoData = -9999.f;           // NODATA, agrolib/mathFunctions/commonConstants.h:31
constexpr double kEpsilon = 0.00001;         // EPSILON, commonConstants.h:252

// isEqual(float,float): |double(a)-double(b)| < EPSILON  (agrolib/mathFunctions/basicMath.h:25-26)
__host__ __device__ __forceinline__ bool isEqualF(float a, float b)
{
    return fabs(double(a) - double(b)) < kEpsilon;
}

// ---- packed FP32x2 (Blackwell FADD2/FMUL2/FFMA2): two cells per lane, half the issue slots ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ float rsqrt_approx(float x)
{
    float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, size multiple of 16 B, both addresses 16 B aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- descriptors passed by value to kernels ----
struct DevGrid {              // a raster resident in HBM, row-major, row 0 = north
    const float* data;
    int nrRows, nrCols;
    double llx, lly, cellSize, invCellSize;
    float flag;
};

enum ClampMode { CLAMP_NONE = 0, CLAMP_PREC = 1, CLAMP_RH = 2, CLAMP_POS = 3 };

struct DevProxy {
    int kind;                 // PB_PROXY_*
    int active, significant;  // currentCombination
    int gridSlot;             // index into DevModel::grid, -1 = none
    float slope, r2;
    float H0, H1, invLapseRate;
    int inversionIsSignificative;
};

// everything the per-cell epilogue needs: getProxyValuesXY + retrend + clamp
// (agrolib/interpolation/interpolation.cpp:2620-2647, 1288-1351, 2532-2558)
struct DevModel {
    int nProxy;
    int useDetrendingVar;     // getUseDetrendingVar(variable), interpolation.cpp:1205
    int useMultipleDetrending, useThermalInversion, useDoNotRetrend, useRetrendOnly;
    int clampMode;
    float rainfallThreshold;
    float minRegressionR2;
    double valueOffset;       // c: station values were centred (v - c); added back before the f32 rounding
    DevProxy proxy[PB_MAX_PROXY];
    DevGrid grid[PB_MAX_PROXY];
    // multiple detrending: functions/params in evaluation order, proxy position each applies to
    int nFit;                 // functions evaluated by functionSum (== fittingFunction.size())
    int nActiveSig;           // active && significant proxies, elevation first (getMultipleDetrendingValues order)
    int activeSigPos[PB_MAX_PROXY];   // proxy position of the k-th pushed value; function k applies to it
    int fitFunction[PB_MAX_PROXY];
    double fitParams[PB_MAX_PROXY][PB_MAX_FIT_PARAMS];
};

// packed station tile layout: xy[i] = {x, x, y, y}, v[i] = {v - c, v - c}; padded with far-away zero-weight stations
struct DevStations {
    const float4* xy;
    const float2* v;
    const float* x;           // plain SoA copies for the exact slow path (cell coincides with a station)
    const float* y;
    const float* val;
    const float* z;           // float(point->z), topographic distance only
    int n;                    // real stations
    int nPadded;              // multiple of the tile size
};

}  // namespace pb
