// kriging_tc.cu — K5: the kriging variance contraction on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Reference arithmetic being replaced (agrolib/interpolation/kriging.cpp): krigingSetWeight :256-261
// (weight_i = sum_j Vinv[i][j] D_j, 2 n (n+1) flop per target) followed by krigingRMSE :287-295 (sqrt|sum_i weight_i D_i|).
//
// Formulation (kriging.cu header): y = M c with M = L^-1 (lower triangular, C = sill - Gamma = L L^T) and
// c_j = sill - gamma(|p - s_j|);  sum_i weight_i D_i = sill - (|y|^2 - (z.y)(z.y - 1) / |z|^2),  z = M 1.
//
// Two kernels share the formulation, the operand layouts and the generator arithmetic:
//   kriging_variance_pair_kernel  (default, further down): a cluster of two CTAs per tile, tcgen05 cta_group::2, one N block per
//                                  pass, two pairs resident per TPC, K stages out of the variogram range skipped;
//   kriging_variance_tc_kernel    (round 1, PB_KRIGING_SINGLE_CTA=1 or systems too large for the pair kernel's stage list): one
//                                  CTA per 256 cells, described first.
//
// GEMM mapping: D[cell][i] = sum_j A[cell][j] * B[i][j]
//   A (2 M tiles x 128 cells x K): c values GENERATED on the fly by 8 generator warps (thread = cell; FP32 variogram of
//     centred FP32 distances), split into FP16 hi + lo, written to shared memory in the canonical K-major no-swizzle UMMA
//     layout;
//   B (N = 128 stations i x K): M pre-split into FP16 hi + lo and pre-tiled on the host in the same canonical layout,
//     one 1-D bulk async copy (TMA engine) per 16 KiB (hi + lo) tile, completion on an mbarrier; only tiles on or below the
//     diagonal exist (M is lower triangular) so half of the K range is skipped; every tile that is loaded serves 256 cells
//     (the L2 -> SM stream of L^-1 is what a 128-cell CTA would be bound by: ~6300 B/clk for the whole chip);
//   D: FP32 accumulators in TMEM: 2 M tiles x 2 N tiles x 128 columns = 512 columns = one pass over 256 stations i; three
//     MMAs per K step (hi*hi + hi*lo + lo*hi) give ~22 mantissa bits on each product;
//   pipeline: kTcStages shared-memory stages of K = 32; warp 8 = loader (station coordinates of the stage + the bulk copies),
//     warp 9 = MMA issuer (one elected lane, tcgen05.mma + tcgen05.commit on the stage's mbarrier), warps 0-7 generate the A
//     tiles of stage s+1.. while the tensor core works on stage s; nobody waits for an MMA except to reuse its buffers;
//   epilogue per pass: tcgen05.ld, thread = cell (TMEM lane), per-thread sums of y^2 and z*y, no cross-lane reduction.
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <string.h>

#include "kriging.cuh"

namespace pb {

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers (tcgen05 / TMEM); mbarrier + bulk copy wrappers are in common.cuh
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dstSmem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dstSmem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmemD),
        "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
          "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
          "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleave"): core matrix = 8 rows x 16 B contiguous;
// LBO = 128 B between the two core matrices of one K = 16 step, SBO = kTcBK * 16 B between 8-row groups; version 1 (Blackwell)
constexpr uint32_t kTcSbo = kTcBK * 16;
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    return uint64_t((saddr & 0x3FFFFu) >> 4) | (uint64_t(128 >> 4) << 16) | (uint64_t(kTcSbo >> 4) << 32) | (uint64_t(1) << 46);
}
// instruction descriptor: D = F32, A = B = F16, both K-major, N = 128, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (uint32_t(kTcNT >> 3) << 17) | (uint32_t(kTcM >> 4) << 24);

__host__ __device__ inline size_t tileOffset(int r, int k) { return size_t(r >> 3) * (kTcBK * 16) + size_t(k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2; }

struct TcParams {
    int n, nPad, nKb, mode;
    int skip;         // bounded model: K stages out of every cell's range are skipped (pair kernel)
    int dbg;          // PB_KRIGING_DBG (timing experiments only: 1 = generators idle, 2 = no MMAs, 4 = no epilogue reads)
    float range, invRange, sn;  // sn = sill - nugget
    double sill, zz;
};

// covariance c = sill - gamma(h) in FP32 (kriging.cpp:160-192 rewritten for c); the model is a template parameter of the
// kernel so that the generator loop carries no switch, and h / range is a multiplication by the precomputed reciprocal
template <int MODE>
__device__ __forceinline__ float covariance(const TcParams& p, float h)
{
    const float t = h * p.invRange;
    if (MODE == PB_KRIGING_SPHERICAL) return (h < p.range) ? p.sn * fmaf(fmaf(0.5f * t, t, -1.5f), t, 1.f) : 0.f;
    if (MODE == PB_KRIGING_EXPONENTIAL) return p.sn * __expf(-3.f * t);
    return p.sn * __expf(-4.f * t * t);
}
// run-time form for the CUDA-core check kernel
__device__ __forceinline__ float covarianceRt(const TcParams& p, float h)
{
    switch (p.mode) {
    case PB_KRIGING_SPHERICAL: return covariance<PB_KRIGING_SPHERICAL>(p, h);
    case PB_KRIGING_EXPONENTIAL: return covariance<PB_KRIGING_EXPONENTIAL>(p, h);
    default: return covariance<PB_KRIGING_GAUSSIAN>(p, h);
    }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void gen_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }      // the 16 generator warps

constexpr int kTcGenThreads = 2 * kTcCells;             // 512: two threads per cell, each half of a stage's K range / of the columns
constexpr int kTcGenWarps = kTcGenThreads / 32;
constexpr int kTcThreads = kTcGenThreads + 64;          // + loader warp + MMA warp
constexpr int kTcStageA = kTcMTiles * 2 * kTcTileBytes; // 2 M tiles x (hi, lo)
constexpr int kTcStageB = (kTcPass / kTcNT) * 2 * kTcTileBytes;
constexpr int kTcStageBytes = kTcStageA + kTcStageB;
constexpr int kTcSmem = kTcStages * kTcStageBytes + kTcStages * kTcBK * 8 + kTcPass * 4 + 2 * kTcCells * 8 + 1024;

template <int MODE>
__global__ void __launch_bounds__(kTcThreads, 1)
kriging_variance_tc_kernel(TcParams p, const double* __restrict__ pos, const unsigned char* __restrict__ Mt,
                           const float* __restrict__ z, KrigingTargets t, long long c0, int m, double* __restrict__ q)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    // stage s: [A: M tile 0 hi, lo, M tile 1 hi, lo][B: N tile 0 hi, lo, N tile 1 hi, lo]
    float2* sRel = reinterpret_cast<float2*>(smem + kTcStages * kTcStageBytes);       // [stage][kTcBK]
    float* sZ = reinterpret_cast<float*>(sRel + kTcStages * kTcBK);                   // [kTcPass]
    double* sRed = reinterpret_cast<double*>(sZ + kTcPass);                           // [2 sums][kTcCells]
    __shared__ __align__(8) uint64_t relFull[kTcStages], aFull[kTcStages], bFull[kTcStages], mmaDone[kTcStages], passDone,
        tmemFree;
    __shared__ uint32_t tmemBase;
    __shared__ double sCenter[2];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long tile0 = (long long)blockIdx.x * kTcCells;
    if (tid == 0) {
        for (int s = 0; s < kTcStages; s++) {
            mbar_init(&relFull[s], 32);
            mbar_init(&aFull[s], kTcGenThreads);
            mbar_init(&bFull[s], 1);
            mbar_init(&mmaDone[s], 1);
        }
        mbar_init(&passDone, 1);
        mbar_init(&tmemFree, kTcGenThreads);
        mbar_fence_init();
        double cx, cy;
        // centre of the tile: its first target (the tile is spatially compact, station offsets reach ~1e6 m)
        if (t.xy) { cx = t.xy[2 * (c0 + tile0)]; cy = t.xy[2 * (c0 + tile0) + 1]; }
        else {
            const int cell = t.validIdx[c0 + tile0];
            const int row = cell / t.grid.nrCols, col = cell - row * t.grid.nrCols;
            cx = t.grid.llx + t.grid.cellSize * (col + 0.5);
            cy = t.grid.lly + t.grid.cellSize * (t.grid.nrRows - row - 0.5);
        }
        sCenter[0] = cx;
        sCenter[1] = cy;
    }
    if (warp == 0) {
        tmem_alloc(&tmemBase, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmemBase;
    const double cx = sCenter[0], cy = sCenter[1];
    const int nPasses = (p.nPad + kTcPass - 1) / kTcPass;

    if (warp == kTcGenWarps) {
        // ------------------------------------------------------------ loader: station coordinates + L^-1 tiles of each stage
        uint32_t it = 0;
        for (int pass = 0; pass < nPasses; pass++) {
            const int i0p = pass * kTcPass;
            const int nT = min(kTcPass / kTcNT, (p.nPad - i0p) / kTcNT);
            const int kbEnd = min(p.nKb, (i0p + nT * kTcNT) / kTcBK);
            for (int kb = 0; kb < kbEnd; kb++, it++) {
                const int s = it % kTcStages;
                const uint32_t u = it / kTcStages;
                if (u > 0) mbar_wait(&mmaDone[s], (u - 1) & 1);            // the MMAs that read this stage's buffers are done
                const int j0 = kb * kTcBK, j = j0 + lane;
                // pairs of stations as {x0, x1, y0, y1}: the generators load a pair's x (y) as one packed FP32x2 operand;
                // padding stations sit far away: covariance 0 against the zero columns of M
                float* relS = reinterpret_cast<float*>(sRel + s * kTcBK) + (lane >> 1) * 4 + (lane & 1);
                relS[0] = (j < p.n) ? float(pos[2 * j] - cx) : 1e18f;
                relS[2] = (j < p.n) ? float(pos[2 * j + 1] - cy) : 1e18f;
                mbar_arrive(&relFull[s]);
                if (lane == 0) {
                    unsigned char* sB = smem + s * kTcStageBytes + kTcStageA;
                    int cnt = 0;
                    for (int tt = 0; tt < nT; tt++) cnt += (j0 <= i0p + tt * kTcNT + kTcNT - 1);
                    mbar_expect_tx(&bFull[s], uint32_t(cnt) * 2u * kTcTileBytes);
                    for (int tt = 0; tt < nT; tt++)
                        if (j0 <= i0p + tt * kTcNT + kTcNT - 1) {
                            const size_t itile = size_t(i0p / kTcNT + tt);
                            bulk_g2s(sB + size_t(tt) * 2 * kTcTileBytes, Mt + (itile * p.nKb + kb) * 2 * kTcTileBytes,
                                     2 * kTcTileBytes, &bFull[s]);
                        }
                }
                __syncwarp();
            }
        }
    } else if (warp == kTcGenWarps + 1) {
        // ------------------------------------------------------------ MMA issuer (one elected lane)
        if (lane == 0) {
            uint32_t it = 0;
            for (int pass = 0; pass < nPasses; pass++) {
                const int i0p = pass * kTcPass;
                const int nT = min(kTcPass / kTcNT, (p.nPad - i0p) / kTcNT);
                const int kbEnd = min(p.nKb, (i0p + nT * kTcNT) / kTcBK);
                if (pass > 0) {
                    mbar_wait(&tmemFree, (pass - 1) & 1);                  // the epilogue has read the accumulators
                    tc_fence_after();
                }
                for (int kb = 0; kb < kbEnd; kb++, it++) {
                    const int s = it % kTcStages;
                    const uint32_t ph = (it / kTcStages) & 1;
                    const int j0 = kb * kTcBK;
                    mbar_wait(&aFull[s], ph);
                    mbar_wait(&bFull[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * kTcStageBytes), sb = sa + kTcStageA;
                    for (int mt = 0; mt < kTcMTiles; mt++) {
                        const uint32_t aHi = sa + uint32_t(mt) * 2 * kTcTileBytes, aLo = aHi + kTcTileBytes;
                        for (int tt = 0; tt < nT; tt++) {
                            if (!(j0 <= i0p + tt * kTcNT + kTcNT - 1)) continue;
                            const uint32_t bHi = sb + uint32_t(tt) * 2 * kTcTileBytes, bLo = bHi + kTcTileBytes;
                            const uint32_t d = tmem + uint32_t(mt * kTcPass + tt * kTcNT);
#pragma unroll
                            for (int ks = 0; ks < kTcBK / 16; ks++) {
                                const uint32_t o = ks * 256;
                                umma_f16(d, umma_desc(aHi + o), umma_desc(bHi + o), kIdesc, (kb > 0 || ks > 0) ? 1u : 0u);
                                umma_f16(d, umma_desc(aHi + o), umma_desc(bLo + o), kIdesc, 1u);
                                umma_f16(d, umma_desc(aLo + o), umma_desc(bHi + o), kIdesc, 1u);
                            }
                        }
                    }
                    umma_commit(&mmaDone[s]);                               // frees the stage once these MMAs have run
                    if (kb == kbEnd - 1) umma_commit(&passDone);
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ generators: two threads per cell (= MMA row = TMEM lane)
        const int cellInCta = tid & (kTcCells - 1), kh = tid >> 8;        // kh: which half of the stage / of the columns
        const int mt = cellInCta >> 7, r = cellInCta & (kTcM - 1);
        const long long cidx = tile0 + cellInCta;
        const bool valid = cidx < m;
        float xr = 0.f, yr = 0.f;
        if (valid) {
            double x, y;
            if (t.xy) { x = t.xy[2 * (c0 + cidx)]; y = t.xy[2 * (c0 + cidx) + 1]; }
            else {
                const int cell = t.validIdx[c0 + cidx];
                const int row = cell / t.grid.nrCols, col = cell - row * t.grid.nrCols;
                x = t.grid.llx + t.grid.cellSize * (col + 0.5);
                y = t.grid.lly + t.grid.cellSize * (t.grid.nrRows - row - 0.5);
            }
            xr = float(x - cx);
            yr = float(y - cy);
        }
        double accY2 = 0., accZY = 0.;
        uint32_t it = 0;
        const uint32_t laneBase = uint32_t((warp & 3) * 32) << 16;
        const f32x2 nxr2 = pack2(-xr, -xr), nyr2 = pack2(-yr, -yr);
        const float sphA = 0.5f * p.sn * p.invRange * p.invRange * p.invRange, sphNB = -1.5f * p.sn * p.invRange;
        const f32x2 sphA2 = pack2(sphA, sphA), sphNB2 = pack2(sphNB, sphNB), sn2 = pack2(p.sn, p.sn);
        const float range2 = p.range * p.range;
        for (int pass = 0; pass < nPasses; pass++) {
            const int i0p = pass * kTcPass;
            const int nT = min(kTcPass / kTcNT, (p.nPad - i0p) / kTcNT);
            const int kbEnd = min(p.nKb, (i0p + nT * kTcNT) / kTcBK);
            for (int kb = 0; kb < kbEnd; kb++, it++) {
                const int s = it % kTcStages;
                const uint32_t u = it / kTcStages;
                if (u > 0) mbar_wait(&mmaDone[s], (u - 1) & 1);            // A[s] is free again
                mbar_wait(&relFull[s], u & 1);
                unsigned char* sA = smem + s * kTcStageBytes + size_t(mt) * 2 * kTcTileBytes;
                const ulonglong2* rel = reinterpret_cast<const ulonglong2*>(sRel + s * kTcBK);     // [pair] = {x0 x1, y0 y1}
#pragma unroll
                for (int g = 0; g < kTcBK / 16; g++) {
                    const int k0 = kh * (kTcBK / 2) + g * 8;
                    __half2 hi[4], lo[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const ulonglong2 sxy = rel[(k0 >> 1) + e];
                        const f32x2 dx = add2(sxy.x, nxr2), dy = add2(sxy.y, nyr2);
                        const f32x2 d2p = fma2(dx, dx, mul2(dy, dy));
                        float d2a, d2b, c0v, c1v;
                        unpack2(d2p, d2a, d2b);
                        if (MODE == PB_KRIGING_SPHERICAL) {
                            // sn (1 - 1.5 t + 0.5 t^3), t = h / range, as h (h^2 A - B) + sn with h^2 = d2 already at hand
                            const f32x2 rs = pack2(rsqrt_approx(fmaxf(d2a, 1e-30f)), rsqrt_approx(fmaxf(d2b, 1e-30f)));
                            const f32x2 h = mul2(d2p, rs);
                            const f32x2 c = fma2(h, fma2(d2p, sphA2, sphNB2), sn2);
                            unpack2(c, c0v, c1v);
                            c0v = (d2a < range2) ? c0v : 0.f;
                            c1v = (d2b < range2) ? c1v : 0.f;
                        } else {
                            c0v = covariance<MODE>(p, d2a * rsqrt_approx(fmaxf(d2a, 1e-30f)));
                            c1v = covariance<MODE>(p, d2b * rsqrt_approx(fmaxf(d2b, 1e-30f)));
                        }
                        const __half2 h2 = __floats2half2_rn(c0v, c1v);
                        const float2 back = __half22float2(h2);
                        hi[e] = h2;
                        lo[e] = __floats2half2_rn(c0v - back.x, c1v - back.y);
                    }
                    const size_t off = tileOffset(r, k0);
                    *reinterpret_cast<uint4*>(sA + off) = *reinterpret_cast<const uint4*>(hi);
                    *reinterpret_cast<uint4*>(sA + kTcTileBytes + off) = *reinterpret_cast<const uint4*>(lo);
                }
                fence_proxy_async_smem();       // generic-proxy stores -> visible to the tensor core (async proxy)
                mbar_arrive(&aFull[s]);
            }
            // ---- epilogue of the pass: columns = stations i of this pass, split between the two threads of a cell ----
            if (tid < kTcPass) sZ[tid] = (i0p + tid < p.nPad) ? z[i0p + tid] : 0.f;
            mbar_wait(&passDone, pass & 1);
            tc_fence_after();
            gen_bar_sync();
            const int nCb = nT * kTcNT / 32;
            for (int cb = kh; cb < nCb; cb += 2) {
                uint32_t v[32];
                tmem_ld32(tmem + laneBase + uint32_t(mt * kTcPass + cb * 32), v);
                float y2 = 0.f, zy = 0.f;
#pragma unroll
                for (int e = 0; e < 32; e++) {
                    const float y = __uint_as_float(v[e]);
                    y2 = fmaf(y, y, y2);
                    zy = fmaf(sZ[cb * 32 + e], y, zy);
                }
                accY2 += double(y2);
                accZY += double(zy);
            }
            tc_fence_before();
            gen_bar_sync();                     // sZ is rewritten for the next pass
            mbar_arrive(&tmemFree);
        }
        if (kh == 1) { sRed[cellInCta] = accY2; sRed[kTcCells + cellInCta] = accZY; }
        gen_bar_sync();
        if (kh == 0 && valid) {
            const double y2 = accY2 + sRed[cellInCta], zy = accZY + sRed[kTcCells + cellInCta];
            q[cidx] = p.sill - (y2 - zy * (zy - 1.) / p.zz);      // = sum_i weight_i D_i of krigingRMSE
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------
// CTA-pair form (tcgen05 cta_group::2): the default for the variance contraction
// ---------------------------------------------------------------------------------------------------------
// The single-CTA kernel above regenerates the covariance tile A 4.5 times on average (once per pass of 256 stations i:
// 512 TMEM columns = 2 M tiles x 256) and its M = N = 128 MMAs read 8 KiB of operands from shared memory per 64 tensor
// clocks, i.e. the whole shared-memory bandwidth, while 16 generator warps store A through the same port: ncu showed the
// tensor pipe 56 % active. Here two CTAs of a cluster (one TPC) run ONE MMA of M = 256 (128 cells from each CTA), N = 256:
//   * each CTA generates the A rows of its own 128 cells only and keeps HALF of the B rows (stations i) of every N block:
//     per 128 tensor clocks a CTA's shared memory feeds 4 KiB of A + 4 KiB of B, half of the single-CTA rate;
//   * the L2 -> SM stream of L^-1 stays at 16 KiB per (stage, N block) per SM (every loaded tile still serves 256 cells);
//   * a pass is ONE N block (256 stations i = 256 TMEM columns per CTA) and the ring has three 32 KiB stages, so that TWO
//     pairs are resident on a TPC (2 x 256 TMEM columns, 2 x 105 KiB of shared memory, 2 x 352 threads): while one pair
//     drains a pass, reads its accumulators (epilogue) or starts up, the other one keeps the tensor cores busy. With the
//     512-column passes of the first pair kernel (one pair per TPC, A regenerated 2.5 instead of 4.5 times) the dense
//     contraction took 7.8 ms on the 925k-cell probe and the skipping one 3.2; two resident pairs: 7.7 and 3.0 ms, although A
//     is regenerated as often as in the single-CTA kernel (the generators are not what bounds it any more).
// Roles per CTA: warps 0-7 generators + epilogue (thread = (cell, half of the stage's K range / of the pass's columns)),
// warp 8 B loader (this CTA's halves of the L^-1 tiles by bulk async copies), warp 9: in the leader CTA (cluster rank 0)
// the MMA issuer, in the peer a relay that forwards "my B halves have landed" to the leader, warp 10 the stations'
// coordinates of each stage (loaded ahead of the wait for the stage). The MMA issuer is lane 0 of its warp; a
// warp-convergent form with the issuing lane picked by elect.sync (no ELECT / R2UR.BROADCAST waterfall around every
// UTCHMMA) measured 10 % slower (PB_KRIGING_DBG=8 selects it).
// A ring of three A stages + ten separate B tiles was tried and is slower (10.6 vs 7.9 ms without stage skipping): the
// issuer then waits on and commits to twice as many barriers per stage.
// Cross-CTA signalling: generator warps of both CTAs arrive on the LEADER's aFull (fence.proxy.async by every writer, then
// one arrive per warp), tcgen05.commit multicasts mmaDone / passDone to both CTAs, the epilogue warps of both
// CTAs arrive on the leader's tmemFree.
constexpr int kPrCells = 128;                                   // cells per CTA = one M tile; the pair's MMA has M = 256
constexpr int kPrN = 256;                                       // MMA N: stations i per MMA, 128 rows of B in each CTA
constexpr int kPrPass = 256;                                    // stations i per pass = TMEM columns a CTA allocates: one N block, so that
                                                                // TWO pairs fit a TPC (2 x 256 of the 512 columns, 2 x 104 KiB of shared memory)
constexpr int kPrStages = 3;                                    // ring of K = 32 stages
constexpr int kPrStageA = 2 * kTcTileBytes;                     // hi, lo of this CTA's 128 cells: 16 KiB
constexpr int kPrBTile = 2 * kTcTileBytes;                      // this CTA's 128 rows of one N block, hi + lo: 16 KiB
constexpr int kPrStageBytes = kPrStageA + (kPrPass / kPrN) * kPrBTile;     // 48 KiB
constexpr int kPrGenWarps = 8;
constexpr int kPrGenThreads = kPrGenWarps * 32;
constexpr int kPrKSplit = kPrGenThreads / kPrCells;             // threads per cell: each a slice of the stage's K range / of the pass's columns
constexpr int kPrGenK = kTcBK / kPrKSplit;                      // K columns a generator thread writes per stage (groups of 8)
constexpr int kPrThreads = kPrGenThreads + 96;                  // + B loader, MMA issuer / relay, station-coordinate warp
constexpr int kPrSmem = kPrStages * kPrStageBytes + kPrStages * kTcBK * 8 + kPrPass * 4 + 2 * (kPrKSplit - 1) * kPrCells * 8 + 1024;
static_assert(kPrGenThreads == kPrPass, "the generators reload z with one element per thread");
static_assert((kPrPass & (kPrPass - 1)) == 0 && kPrPass >= 32 && kPrPass <= 512, "TMEM allocations are powers of two");
constexpr int kPrMaxPasses = 128;
constexpr int kPrMaxKb = 1024;                                  // K stages the active list can hold (n <= 32768)
constexpr uint32_t kIdescPair = (1u << 4) | (uint32_t(kPrN >> 3) << 17) | (uint32_t(256 >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    // relaxed arrive (as CUTLASS after fence_barrier_init): .release would put a MEMBAR.ALL.GPU in front of both syncs
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local, uint32_t rank)
{
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t clusterAddr)
{
    // default semantics (release at CTA scope) as in CUTLASS' ClusterBarrier::arrive(cta_id): the data the leader's MMA reads
    // from this CTA was made visible to the async proxy by fence.proxy.async in the writing threads; .release.cluster would
    // put a MEMBAR.ALL.GPU in front of every arrive (measured: half of the kernel's stall samples)
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(clusterAddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity)
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
__device__ __forceinline__ void tmem_alloc2(uint32_t* dstSmem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dstSmem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmemD),
        "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the mbarrier at this shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void umma2_commit(uint64_t* bar)
{
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
// warp-convergent forms: every lane executes them, one elected lane issues
__device__ __forceinline__ void umma2_f16_elect(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p, e;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmemD),
        "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_elect(uint64_t* bar)
{
    const uint16_t mask = 3;
    asm volatile(
        "{\n"
        ".reg .pred e;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
        "}\n" ::"r"(smem_u32(bar)),
        "h"(mask)
        : "memory");
}
__device__ __forceinline__ float sqrt_approx(float x)
{
    float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
}
__device__ __forceinline__ float ex2_approx(float x)
{
    float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
}
__device__ __forceinline__ void pair_gen_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kPrGenThreads) : "memory"); }

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPrThreads, 2)
kriging_variance_pair_kernel(TcParams p, const double* __restrict__ pos, const unsigned char* __restrict__ Mt,
                             const float* __restrict__ z, const double* __restrict__ kbBox, KrigingTargets t, long long c0, int m,
                             double* __restrict__ q, unsigned long long* __restrict__ mmaCount)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    // stage s: [A: hi, lo of this CTA's 128 cells][B: N block 0 hi, lo, N block 1 hi, lo of this CTA's 128 stations i]
    float2* sRel = reinterpret_cast<float2*>(smem + kPrStages * kPrStageBytes);       // [stage][kTcBK]
    float* sZ = reinterpret_cast<float*>(sRel + kPrStages * kTcBK);                   // [kPrPass]
    double* sRed = reinterpret_cast<double*>(sZ + kPrPass);                           // [2 sums][kPrKSplit - 1 partials][kPrCells]
    __shared__ __align__(8) uint64_t relFull[kPrStages], aFull[kPrStages], bFull[kPrStages], bPeer[kPrStages], mmaDone[kPrStages],
        passDone, tmemFree;
    __shared__ uint32_t tmemBase;
    __shared__ double sCenter[2];
    __shared__ double sBoxPart[8][4];
    __shared__ unsigned short sList[kPrMaxKb];      // K stages (32 stations j) that can be inside the range of a cell of the pair
    __shared__ int sNAct;
    __shared__ unsigned short sPassStages[kPrMaxPasses];   // active stages of each pass = list entries below the end of its K range

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const long long tile0 = (long long)blockIdx.x * kPrCells;
    if (tid == 0) {
        for (int s = 0; s < kPrStages; s++) {
            mbar_init(&relFull[s], 32);
            mbar_init(&aFull[s], 2 * kPrGenWarps);      // leader only: one arrival per generator warp of either CTA
            mbar_init(&bFull[s], 1);
            mbar_init(&bPeer[s], 1);                    // leader only: the peer's relay
            mbar_init(&mmaDone[s], 1);
        }
        mbar_init(&passDone, 1);
        mbar_init(&tmemFree, 2 * kPrGenWarps);          // leader only
        mbar_fence_init();
        // centre of the CTA's cells: its first target (an empty peer CTA of the last pair takes the last target)
        const long long cc = c0 + (tile0 < m ? tile0 : (long long)m - 1);
        double cx, cy;
        if (t.xy) { cx = t.xy[2 * cc]; cy = t.xy[2 * cc + 1]; }
        else {
            const int cell = t.validIdx[cc];
            const int row = cell / t.grid.nrCols, col = cell - row * t.grid.nrCols;
            cx = t.grid.llx + t.grid.cellSize * (col + 0.5);
            cy = t.grid.lly + t.grid.cellSize * (t.grid.nrRows - row - 0.5);
        }
        sCenter[0] = cx;
        sCenter[1] = cy;
    }
    // ---- K stages of this PAIR of CTAs: with a bounded model (p.skip) c_j = 0 beyond the range, so a stage whose 32 stations
    // (spatially compact after the host's k-d ordering) all lie farther than the range from the pair's 256 cells contributes
    // nothing to y = L^-1 c: it is neither generated, loaded nor multiplied. Both CTAs build the same ascending list.
    {
        const long long pairBase = (long long)(blockIdx.x & ~1u) * kPrCells;
        if (tid < 2 * kPrCells) {
            const long long ci = pairBase + tid;
            double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
            if (p.skip && ci < m) {
                double x, y;
                if (t.xy) { x = t.xy[2 * (c0 + ci)]; y = t.xy[2 * (c0 + ci) + 1]; }
                else {
                    const int cell = t.validIdx[c0 + ci];
                    const int row = cell / t.grid.nrCols, col = cell - row * t.grid.nrCols;
                    x = t.grid.llx + t.grid.cellSize * (col + 0.5);
                    y = t.grid.lly + t.grid.cellSize * (t.grid.nrRows - row - 0.5);
                }
                x0 = x1 = x;
                y0 = y1 = y;
            }
            for (int o = 16; o > 0; o >>= 1) {
                x0 = fmin(x0, __shfl_xor_sync(~0u, x0, o));
                x1 = fmax(x1, __shfl_xor_sync(~0u, x1, o));
                y0 = fmin(y0, __shfl_xor_sync(~0u, y0, o));
                y1 = fmax(y1, __shfl_xor_sync(~0u, y1, o));
            }
            if (lane == 0) { sBoxPart[warp][0] = x0; sBoxPart[warp][1] = x1; sBoxPart[warp][2] = y0; sBoxPart[warp][3] = y1; }
        }
        __syncthreads();
        if (warp == 0) {
            double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
            for (int w = 0; w < 8; w++) {
                x0 = fmin(x0, sBoxPart[w][0]); x1 = fmax(x1, sBoxPart[w][1]);
                y0 = fmin(y0, sBoxPart[w][2]); y1 = fmax(y1, sBoxPart[w][3]);
            }
            const double reach = double(p.range) * (1. + 1e-6), reach2 = reach * reach;
            int cnt = 0;
            for (int base = 0; base < p.nKb; base += 32) {
                const int kb = base + lane;
                bool act = false;
                if (kb < p.nKb) {
                    if (!p.skip) act = true;
                    else {
                        const double bx0 = kbBox[4 * kb], bx1 = kbBox[4 * kb + 1], by0 = kbBox[4 * kb + 2], by1 = kbBox[4 * kb + 3];
                        const double ex = fmax(fmax(x0 - bx1, bx0 - x1), 0.), ey = fmax(fmax(y0 - by1, by0 - y1), 0.);
                        act = ex * ex + ey * ey < reach2;          // an empty (padding) stage has bx0 = +inf: never active
                    }
                }
                const unsigned b = __ballot_sync(~0u, act);
                if (act) sList[cnt + __popc(b & ((1u << lane) - 1u))] = (unsigned short)kb;
                cnt += __popc(b);
            }
            PB_DEV_ASSERT(cnt <= kPrMaxKb && (p.nPad + kPrPass - 1) / kPrPass <= kPrMaxPasses && p.nPad % kPrN == 0);
            if (lane == 0) sNAct = cnt;
            __syncwarp();
            for (int pass = lane; pass < (p.nPad + kPrPass - 1) / kPrPass; pass += 32) {
                const int i0p = pass * kPrPass;
                const int nNb = min(kPrPass / kPrN, (p.nPad - i0p) / kPrN);
                const int kbEnd = min(p.nKb, (i0p + nNb * kPrN) / kTcBK);
                int c = 0;
                while (c < cnt && sList[c] < kbEnd) c++;
                sPassStages[pass] = (unsigned short)c;
            }
        }
    }
    if (warp == 0) {
        tmem_alloc2(&tmemBase, kPrPass);
        tmem_relinquish2();
    }
    tc_fence_before();
    cluster_sync_all();                                  // barriers of both CTAs initialised before any remote arrive; sList visible
    tc_fence_after();
    const uint32_t tmem = tmemBase;
    const double cx = sCenter[0], cy = sCenter[1];
    const double invRangeD = 1.0 / double(p.range);
    const int nPasses = (p.nPad + kPrPass - 1) / kPrPass;
    const int nAct = sNAct;

    if (warp == kPrGenWarps) {
        // ------------------------------------------------------------ B loader: this CTA's halves of the L^-1 tiles
        if (lane == 0) {
            uint32_t it = 0;
            int idxEnd = 0;
            for (int pass = 0; pass < nPasses; pass++) {
                const int i0p = pass * kPrPass;
                const int nNb = min(kPrPass / kPrN, (p.nPad - i0p) / kPrN);
                const int kbEnd = min(p.nKb, (i0p + nNb * kPrN) / kTcBK);
                while (idxEnd < nAct && sList[idxEnd] < kbEnd) idxEnd++;
                for (int idx = 0; idx < idxEnd; idx++, it++) {
                    const int kb = sList[idx], j0 = kb * kTcBK;
                    const int s = it % kPrStages;
                    const uint32_t u = it / kPrStages;
                    if (u > 0) mbar_wait(&mmaDone[s], (u - 1) & 1);        // the MMAs that read this stage's buffers are done
                    unsigned char* sB = smem + s * kPrStageBytes + kPrStageA;
                    int cnt = 0;
                    for (int nb = 0; nb < nNb; nb++) cnt += (j0 <= i0p + nb * kPrN + kPrN - 1);
                    mbar_expect_tx(&bFull[s], uint32_t(cnt) * kPrBTile);
                    for (int nb = 0; nb < nNb; nb++)
                        if (j0 <= i0p + nb * kPrN + kPrN - 1) {
                            const size_t itile = size_t(i0p / kTcNT + 2 * nb) + rank;      // this CTA's 128 rows of the N block
                            PB_DEV_ASSERT(itile < size_t(p.nPad / kTcNT) && kb < p.nKb);
                            bulk_g2s(sB + size_t(nb) * kPrBTile, Mt + (itile * p.nKb + kb) * kPrBTile, kPrBTile, &bFull[s]);
                        }
                }
            }
        }
        __syncwarp();
    } else if (warp == kPrGenWarps + 2) {
        // ------------------------------------------------------------ station coordinates of each stage
        uint32_t it = 0;
        int idxEnd = 0;
        for (int pass = 0; pass < nPasses; pass++) {
            const int i0p = pass * kPrPass;
            const int nNb = min(kPrPass / kPrN, (p.nPad - i0p) / kPrN);
            const int kbEnd = min(p.nKb, (i0p + nNb * kPrN) / kTcBK);
            while (idxEnd < nAct && sList[idxEnd] < kbEnd) idxEnd++;
            for (int idx = 0; idx < idxEnd; idx++, it++) {
                const int kb = sList[idx];
                const int s = it % kPrStages;
                const uint32_t u = it / kPrStages;
                const int j = kb * kTcBK + lane;
                // coordinates in units of the range, pairs of stations as {x0, x1, y0, y1}; padding stations sit far away
                const float rx = (j < p.n) ? float((pos[2 * j] - cx) * invRangeD) : 1e12f;
                const float ry = (j < p.n) ? float((pos[2 * j + 1] - cy) * invRangeD) : 1e12f;
                if (u > 0) mbar_wait(&mmaDone[s], (u - 1) & 1);            // the generators have read the stage's previous use
                float* relS = reinterpret_cast<float*>(sRel + s * kTcBK) + (lane >> 1) * 4 + (lane & 1);
                relS[0] = rx;
                relS[2] = ry;
                mbar_arrive(&relFull[s]);
            }
        }
    } else if (warp == kPrGenWarps + 1) {
        uint32_t it = 0, ep = 0, nMma = 0;
        int idxEnd = 0;
        if (rank != 0) {
            // -------------------------------------------------------- relay (peer CTA): my B halves of the stage are in place
            if (lane == 0) {
                const uint32_t bPeerLeader = mapa_shared(smem_u32(&bPeer[0]), 0);
                for (int pass = 0; pass < nPasses; pass++) {
                    const int i0p = pass * kPrPass;
                    const int nNb = min(kPrPass / kPrN, (p.nPad - i0p) / kPrN);
                    const int kbEnd = min(p.nKb, (i0p + nNb * kPrN) / kTcBK);
                    while (idxEnd < nAct && sList[idxEnd] < kbEnd) idxEnd++;
                    for (int idx = 0; idx < idxEnd; idx++, it++) {
                        const int s = it % kPrStages;
                        mbar_wait(&bFull[s], (it / kPrStages) & 1);
                        mbar_arrive_cluster(bPeerLeader + uint32_t(s) * 8u);
                    }
                }
            }
            __syncwarp();
        } else {
            // -------------------------------------------------------- MMA issuer (leader CTA), warp-convergent
            for (int pass = 0; pass < nPasses; pass++) {
                const int i0p = pass * kPrPass;
                const int nNb = min(kPrPass / kPrN, (p.nPad - i0p) / kPrN);
                const int kbEnd = min(p.nKb, (i0p + nNb * kPrN) / kTcBK);
                while (idxEnd < nAct && sList[idxEnd] < kbEnd) idxEnd++;
                if (idxEnd == 0) continue;                                  // no station of this pass's K range is near: y = 0
                if (ep > 0) {
                    if (lane == 0) mbar_wait(&tmemFree, (ep - 1) & 1);     // both CTAs' epilogues have read the accumulators
                    __syncwarp();
                    tc_fence_after();
                }
                ep++;
                for (int idx = 0; idx < idxEnd; idx++, it++) {
                    const int kb = sList[idx], j0 = kb * kTcBK;
                    const int s = it % kPrStages;
                    const uint32_t ph = (it / kPrStages) & 1;
                    if (lane == 0) {
                        mbar_wait(&aFull[s], ph);
                        mbar_wait(&bFull[s], ph);
                        mbar_wait(&bPeer[s], ph);
                    }
                    __syncwarp();
                    tc_fence_after();
                    if (!(p.dbg & 8)) {
                        if (lane == 0) {
                            const uint32_t sa = smem_u32(smem + s * kPrStageBytes), sb = sa + kPrStageA;
                            const uint32_t aHi = sa, aLo = sa + kTcTileBytes;
                            for (int nb = 0; nb < nNb; nb++) {
                                if (!(j0 <= i0p + nb * kPrN + kPrN - 1)) continue;
                                nMma += 3 * (kTcBK / 16);
                                const uint32_t bHi = sb + uint32_t(nb) * kPrBTile, bLo = bHi + kTcTileBytes;
                                const uint32_t d = tmem + uint32_t(nb * kPrN);
#pragma unroll
                                for (int ks = 0; ks < kTcBK / 16; ks++) {
                                    const uint32_t o = ks * 256;
                                    umma2_f16(d, umma_desc(aHi + o), umma_desc(bHi + o), kIdescPair, (idx > 0 || ks > 0) ? 1u : 0u);
                                    umma2_f16(d, umma_desc(aHi + o), umma_desc(bLo + o), kIdescPair, 1u);
                                    umma2_f16(d, umma_desc(aLo + o), umma_desc(bHi + o), kIdescPair, 1u);
                                }
                            }
                            umma2_commit(&mmaDone[s]);
                            if (idx == idxEnd - 1) umma2_commit(&passDone);
                        }
                        __syncwarp();
                        continue;
                    }
                    if (!(p.dbg & 2)) {
                        const uint32_t sa = smem_u32(smem + s * kPrStageBytes), sb = sa + kPrStageA;
                        const uint32_t aHi = sa, aLo = sa + kTcTileBytes;
                        for (int nb = 0; nb < nNb; nb++) {
                            if (!(j0 <= i0p + nb * kPrN + kPrN - 1)) continue;
                            const uint32_t bHi = sb + uint32_t(nb) * kPrBTile, bLo = bHi + kTcTileBytes;
                            const uint32_t d = tmem + uint32_t(nb * kPrN);
#pragma unroll
                            for (int ks = 0; ks < kTcBK / 16; ks++) {
                                const uint32_t o = ks * 256;
                                umma2_f16_elect(d, umma_desc(aHi + o), umma_desc(bHi + o), kIdescPair, (idx > 0 || ks > 0) ? 1u : 0u);
                                umma2_f16_elect(d, umma_desc(aHi + o), umma_desc(bLo + o), kIdescPair, 1u);
                                umma2_f16_elect(d, umma_desc(aLo + o), umma_desc(bHi + o), kIdescPair, 1u);
                            }
                        }
                    }
                    umma2_commit_elect(&mmaDone[s]);                        // frees the stage in both CTAs
                    if (idx == idxEnd - 1) umma2_commit_elect(&passDone);
                }
            }
            // issued M = 256, N = 256, K = 16 MMAs of this pair (bench.py: executed tensor flops next to the algorithmic ones)
            if (lane == 0 && mmaCount && nMma) atomicAdd(mmaCount, (unsigned long long)nMma);
        }
    } else {
        // ------------------------------------------------------------ generators: kPrKSplit threads per cell (= MMA row = TMEM lane)
        const int r = (warp & 3) * 32 + lane;             // a warp reads the TMEM lanes 32 (warp % 4) ..
        const int kq = warp >> 2;                         // slice of the stage's K range / of the pass's columns
        const long long cidx = tile0 + r;
        const bool valid = cidx < m;
        float xr = 0.f, yr = 0.f;
        if (valid) {
            double x, y;
            if (t.xy) { x = t.xy[2 * (c0 + cidx)]; y = t.xy[2 * (c0 + cidx) + 1]; }
            else {
                const int cell = t.validIdx[c0 + cidx];
                const int row = cell / t.grid.nrCols, col = cell - row * t.grid.nrCols;
                x = t.grid.llx + t.grid.cellSize * (col + 0.5);
                y = t.grid.lly + t.grid.cellSize * (t.grid.nrRows - row - 0.5);
            }
            xr = float((x - cx) * invRangeD);
            yr = float((y - cy) * invRangeD);
        }
        double accY2 = 0., accZY = 0.;
        uint32_t it = 0;
        const uint32_t laneBase = uint32_t((warp & 3) * 32) << 16;
        const uint32_t aFullLeader = mapa_shared(smem_u32(&aFull[0]), 0), tmemFreeLeader = mapa_shared(smem_u32(&tmemFree), 0);
        const f32x2 nxr2 = pack2(-xr, -xr), nyr2 = pack2(-yr, -yr);
        const f32x2 one2 = pack2(1.f, 1.f), negOne2 = pack2(-1.f, -1.f), sn2 = pack2(p.sn, p.sn), hsn2 = pack2(0.5f * p.sn, 0.5f * p.sn);
        const int k0 = kq * kPrGenK;
        const size_t offA = tileOffset(r, k0);         // + g * 128 for the g-th group of 8 K columns
        // one stage of A: the generators do not need to know which stations it holds (their coordinates arrive in sRel)
        auto genStage = [&]() {
            {
                const int s = it % kPrStages;
                const uint32_t u = it / kPrStages;
                if (u > 0) mbar_wait(&mmaDone[s], (u - 1) & 1);            // A[s] is free again
                mbar_wait(&relFull[s], u & 1);
                unsigned char* sA = smem + s * kPrStageBytes;
                const ulonglong2* rel = reinterpret_cast<const ulonglong2*>(sRel + s * kTcBK) + (k0 >> 1);     // [pair] = {x0 x1, y0 y1}
                if (!(p.dbg & 1)) {
#pragma unroll
                for (int g = 0; g < kPrGenK / 8; g++) {
                __half2 hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const ulonglong2 sxy = rel[4 * g + e];
                    const f32x2 dx = add2(sxy.x, nxr2), dy = add2(sxy.y, nyr2);
                    const f32x2 t2p = fma2(dx, dx, mul2(dy, dy));          // (h / range)^2
                    float ta, tb, c0v, c1v;
                    unpack2(t2p, ta, tb);
                    if (MODE == PB_KRIGING_SPHERICAL) {
                        // sn (1 - 1.5 t + 0.5 t^3) = sn (1 - t)^2 (1 + t / 2), t clamped to 1: exactly 0 beyond the range
                        const f32x2 tt = pack2(fminf(sqrt_approx(ta), 1.f), fminf(sqrt_approx(tb), 1.f));
                        const f32x2 um = fma2(tt, negOne2, one2);
                        const f32x2 c = mul2(mul2(um, um), fma2(tt, hsn2, sn2));
                        unpack2(c, c0v, c1v);
                    } else if (MODE == PB_KRIGING_EXPONENTIAL) {
                        c0v = p.sn * ex2_approx(-4.328085123f * sqrt_approx(ta));      // exp(-3 t)
                        c1v = p.sn * ex2_approx(-4.328085123f * sqrt_approx(tb));
                    } else {
                        c0v = p.sn * ex2_approx(-5.770780164f * ta);                   // exp(-4 t^2)
                        c1v = p.sn * ex2_approx(-5.770780164f * tb);
                    }
                    const __half2 h2 = __floats2half2_rn(c0v, c1v);
                    const float2 back = __half22float2(h2);
                    hi[e] = h2;
                    lo[e] = __floats2half2_rn(c0v - back.x, c1v - back.y);
                }
                *reinterpret_cast<uint4*>(sA + offA + g * 128) = *reinterpret_cast<const uint4*>(hi);
                *reinterpret_cast<uint4*>(sA + kTcTileBytes + offA + g * 128) = *reinterpret_cast<const uint4*>(lo);
                }
                }
                fence_proxy_async_smem();       // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(aFullLeader + uint32_t(s) * 8u);
                it++;
            }
        };
        // active stages of a pass = list entries below the end of its K range
        auto stagesOf = [&](int pass) { return pass < nPasses ? int(sPassStages[pass]) : 0; };
        uint32_t ep = 0;
        int carried = 0;                        // stages of this pass already generated while the previous pass drained
        for (int pass = 0; pass < nPasses; pass++) {
            const int i0p = pass * kPrPass;
            const int nNb = min(kPrPass / kPrN, (p.nPad - i0p) / kPrN);
            const int nSt = stagesOf(pass);
            if (nSt == 0) continue;                                         // nothing near in this pass's K range: y = 0
            for (int k = carried; k < nSt; k++) genStage();
            // run ahead: the first stages of the next pass are generated while the tensor core drains this one (their ring
            // slots are freed by this pass's MMAs; the issuer holds the next pass's MMAs back until tmemFree), so the
            // pipeline is full again as soon as the epilogue below has released the accumulators
            carried = min(kPrStages - 1, stagesOf(pass + 1));
            for (int k = 0; k < carried; k++) genStage();
            // ---- epilogue of the pass: columns = stations i of this pass, a slice for each of the threads of a cell ----
            sZ[tid] = (i0p + tid < p.nPad) ? z[i0p + tid] : 0.f;          // kPrGenThreads == kPrPass
            mbar_wait(&passDone, ep & 1);
            ep++;
            // the first N block has columns only if the pass's first active stage reaches it (its K range ends 256 earlier)
            const bool hasNb0 = int(sList[0]) * kTcBK <= i0p + kPrN - 1;
            tc_fence_after();
            pair_gen_bar_sync();
            const int nCb = nNb * kPrN / 32;
            for (int cb = kq; cb < nCb && !(p.dbg & 4); cb += kPrKSplit) {
                if (cb < kPrN / 32 && !hasNb0) continue;
                uint32_t v[32];
                tmem_ld32(tmem + laneBase + uint32_t(cb * 32), v);
                float y2 = 0.f, zy = 0.f;
                const float4* z4 = reinterpret_cast<const float4*>(sZ + cb * 32);
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const float4 zz4 = z4[e];
                    const float y0 = __uint_as_float(v[4 * e]), y1 = __uint_as_float(v[4 * e + 1]);
                    const float y2_ = __uint_as_float(v[4 * e + 2]), y3 = __uint_as_float(v[4 * e + 3]);
                    y2 = fmaf(y0, y0, y2); zy = fmaf(zz4.x, y0, zy);
                    y2 = fmaf(y1, y1, y2); zy = fmaf(zz4.y, y1, zy);
                    y2 = fmaf(y2_, y2_, y2); zy = fmaf(zz4.z, y2_, zy);
                    y2 = fmaf(y3, y3, y2); zy = fmaf(zz4.w, y3, zy);
                }
                accY2 += double(y2);
                accZY += double(zy);
            }
            tc_fence_before();
            pair_gen_bar_sync();                // sZ is rewritten for the next pass; every warp's TMEM reads are complete
            if (lane == 0) mbar_arrive_cluster(tmemFreeLeader);
        }
        if (kq > 0) { sRed[(kq - 1) * kPrCells + r] = accY2; sRed[(kPrKSplit - 1 + kq - 1) * kPrCells + r] = accZY; }
        pair_gen_bar_sync();
        if (kq == 0 && valid) {
            double y2 = accY2, zy = accZY;
#pragma unroll
            for (int h = 0; h < kPrKSplit - 1; h++) { y2 += sRed[h * kPrCells + r]; zy += sRed[(kPrKSplit - 1 + h) * kPrCells + r]; }
            q[cidx] = p.sill - (y2 - zy * (zy - 1.) / p.zz);      // = sum_i weight_i D_i of krigingRMSE
        }
    }
    tc_fence_before();
    cluster_sync_all();                         // nobody leaves while the peer can still signal or read this CTA
    if (warp == 0) tmem_dealloc2(tmem, kPrPass);
}

// CUDA-core check of the same whitened formulation from the same pre-tiled FP16 pieces (tests / debugging only:
// PB_KRIGING_SIMPLE=1). One thread per target, O(n^2) covariance evaluations: small systems only.
__global__ void kriging_variance_simple_kernel(TcParams p, const double* __restrict__ pos, const unsigned char* __restrict__ Mt,
                                               const float* __restrict__ z, KrigingTargets t, long long c0, int m,
                                               double* __restrict__ q)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    double x, y;
    if (t.xy) { x = t.xy[2 * (c0 + c)]; y = t.xy[2 * (c0 + c) + 1]; }
    else {
        const int cell = t.validIdx[c0 + c];
        const int row = cell / t.grid.nrCols, col = cell - row * t.grid.nrCols;
        x = t.grid.llx + t.grid.cellSize * (col + 0.5);
        y = t.grid.lly + t.grid.cellSize * (t.grid.nrRows - row - 0.5);
    }
    double y2 = 0, zy = 0;
    for (int i = 0; i < p.n; i++) {
        float yi = 0.f;
        const int itile = i / kTcNT, rr = i % kTcNT;
        for (int j = 0; j <= i; j++) {
            const int kb = j / kTcBK, kk = j % kTcBK;
            const unsigned char* tile = Mt + (size_t(itile) * p.nKb + kb) * 2 * kTcTileBytes;
            const size_t off = tileOffset(rr, kk);
            const float mh = __half2float(*reinterpret_cast<const __half*>(tile + off));
            const float ml = __half2float(*reinterpret_cast<const __half*>(tile + kTcTileBytes + off));
            const float dx = float(pos[2 * j] - x), dy = float(pos[2 * j + 1] - y);
            const float cj = covarianceRt(p, sqrtf(dx * dx + dy * dy));
            yi += (mh + ml) * cj;
        }
        y2 += double(yi) * yi;
        zy += double(z[i]) * yi;
    }
    q[c] = p.sill - (y2 - zy * (zy - 1.) / p.zz);
}

// ---------------------------------------------------------------------------------------------------------
// host: split M into FP16 hi/lo pieces, pre-tiled in the canonical UMMA layout; upload with z
// ---------------------------------------------------------------------------------------------------------
int krigingUploadWhitened(KrigingSystem& k, const std::vector<double>& M, const std::vector<double>& z, double zz,
                          const double* pos)
{
    const int n = k.n;
    const int nPad = (n + 2 * kTcNT - 1) / (2 * kTcNT) * (2 * kTcNT);     // whole N blocks of the CTA-pair kernel
    const int nIt = nPad / kTcNT, nKb = nPad / kTcBK;
    double maxAbs = 0;
    for (double v : M) maxAbs = std::max(maxAbs, fabs(v));
    if (!(maxAbs < 6.0e4) || !(k.params.sillNugget < 6.0e4)) return 1;      // outside FP16 range: exact path instead
    const size_t bytes = size_t(nIt) * nKb * 2 * kTcTileBytes;
    std::vector<unsigned char> blob(bytes, 0);
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (int it = 0; it < nIt; it++)
        for (int kb = 0; kb < nKb; kb++) {
            if (kb * kTcBK > it * kTcNT + kTcNT - 1) continue;          // above the diagonal: never loaded
            unsigned char* tile = blob.data() + (size_t(it) * nKb + kb) * 2 * kTcTileBytes;
            for (int r = 0; r < kTcNT; r++) {
                const int i = it * kTcNT + r;
                if (i >= n) continue;
                for (int kk = 0; kk < kTcBK; kk++) {
                    const int j = kb * kTcBK + kk;
                    if (j > i) break;
                    const float v = float(M[size_t(i) * n + j]);
                    const __half h = __float2half_rn(v);
                    const __half l = __float2half_rn(float(M[size_t(i) * n + j] - double(__half2float(h))));
                    const size_t off = tileOffset(r, kk);
                    memcpy(tile + off, &h, 2);
                    memcpy(tile + kTcTileBytes + off, &l, 2);
                    (void)v;
                }
            }
        }
    std::vector<float> zf(nPad, 0.f);
    for (int i = 0; i < n; i++) zf[i] = float(z[i]);
    // bounding box of every K stage's stations (kriging.cu kdOrder makes them compact); padding stages are empty
    std::vector<double> box(size_t(4) * nKb);
    for (int kb = 0; kb < nKb; kb++) {
        double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
        for (int j = kb * kTcBK; j < std::min(n, (kb + 1) * kTcBK); j++) {
            x0 = std::min(x0, pos[2 * j]); x1 = std::max(x1, pos[2 * j]);
            y0 = std::min(y0, pos[2 * j + 1]); y1 = std::max(y1, pos[2 * j + 1]);
        }
        box[4 * kb] = x0; box[4 * kb + 1] = x1; box[4 * kb + 2] = y0; box[4 * kb + 3] = y1;
    }
    if (cudaMalloc(&k.dKbBox, sizeof(double) * box.size()) != cudaSuccess) return 2;
    if (cudaMemcpy(k.dKbBox, box.data(), sizeof(double) * box.size(), cudaMemcpyHostToDevice) != cudaSuccess) return 2;
    if (cudaMalloc(&k.dMmaCount, sizeof(unsigned long long)) != cudaSuccess) return 2;
    if (cudaMemset(k.dMmaCount, 0, sizeof(unsigned long long)) != cudaSuccess) return 2;
    if (cudaMalloc(&k.dMhi, bytes) != cudaSuccess) return 2;
    if (cudaMalloc(&k.dZ, sizeof(float) * nPad) != cudaSuccess) return 2;
    if (cudaMemcpy(k.dMhi, blob.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) return 2;
    if (cudaMemcpy(k.dZ, zf.data(), sizeof(float) * nPad, cudaMemcpyHostToDevice) != cudaSuccess) return 2;
    k.nPad = nPad;
    k.nIt = nIt;
    k.nKb = nKb;
    k.zz = zz;
    return 0;
}

cudaError_t launchKrigingVarianceTc(const KrigingSystem& k, const KrigingTargets& t, long long c0, int m, double* q,
                                    cudaStream_t s)
{
    TcParams p{};
    p.n = k.n;
    p.nPad = k.nPad;
    p.nKb = k.nKb;
    p.mode = k.params.mode;
    p.range = float(k.params.range);
    p.invRange = 1.f / p.range;
    p.sn = float(k.params.sillNugget);
    p.sill = k.params.sill;
    p.zz = k.zz;
    static const int dbg = getenv("PB_KRIGING_DBG") ? atoi(getenv("PB_KRIGING_DBG")) : 0;
    p.dbg = dbg;
    static const bool simple = getenv("PB_KRIGING_SIMPLE") && atoi(getenv("PB_KRIGING_SIMPLE")) != 0;
    if (simple) {
        kriging_variance_simple_kernel<<<(m + 127) / 128, 128, 0, s>>>(p, k.dPos, k.dMhi, k.dZ, t, c0, m, q);
        return cudaGetLastError();
    }
    static const bool envSingle = getenv("PB_KRIGING_SINGLE_CTA") && atoi(getenv("PB_KRIGING_SINGLE_CTA")) != 0;
    static const bool noSkip = getenv("PB_KRIGING_NOSKIP") && atoi(getenv("PB_KRIGING_NOSKIP")) != 0;
    const bool single = envSingle || k.nKb > kPrMaxKb || (k.nPad + kPrPass - 1) / kPrPass > kPrMaxPasses;
    p.skip = (p.mode == PB_KRIGING_SPHERICAL && !noSkip) ? 1 : 0;
    cudaError_t e;
    if (single) {
        const int grid = (m + kTcCells - 1) / kTcCells;
        auto launch = [&](auto kernel) -> cudaError_t {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem);
            if (e != cudaSuccess) return e;
            kernel<<<grid, kTcThreads, kTcSmem, s>>>(p, k.dPos, k.dMhi, k.dZ, t, c0, m, q);
            return cudaSuccess;
        };
        switch (p.mode) {
        case PB_KRIGING_SPHERICAL: e = launch(kriging_variance_tc_kernel<PB_KRIGING_SPHERICAL>); break;
        case PB_KRIGING_EXPONENTIAL: e = launch(kriging_variance_tc_kernel<PB_KRIGING_EXPONENTIAL>); break;
        default: e = launch(kriging_variance_tc_kernel<PB_KRIGING_GAUSSIAN>); break;
        }
    } else {
        const int grid = 2 * ((m + 2 * kPrCells - 1) / (2 * kPrCells));      // whole CTA pairs (__cluster_dims__(2, 1, 1))
        auto launch = [&](auto kernel) -> cudaError_t {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPrSmem);
            if (e != cudaSuccess) return e;
            // two CTAs per SM (two pairs per TPC): the whole L1 / shared-memory array as shared memory
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (e != cudaSuccess) return e;
            // PB_KRIGING_PAD_SMEM=90000 pads the dynamic shared memory so that only ONE CTA fits an SM (A/B of the co-residency:
            // 3.0 -> 3.85 ms on tools/bench_kriging.py; cudaOccupancyMaxActiveClusters reports 74 clusters for this kernel either
            // way, the hardware places two pairs per TPC when they fit)
            static const int padSmem = getenv("PB_KRIGING_PAD_SMEM") ? atoi(getenv("PB_KRIGING_PAD_SMEM")) : 0;
            if (padSmem) {
                cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPrSmem + padSmem);
                kernel<<<grid, kPrThreads, kPrSmem + padSmem, s>>>(p, k.dPos, k.dMhi, k.dZ, k.dKbBox, t, c0, m, q, k.dMmaCount);
                return cudaSuccess;
            }
            kernel<<<grid, kPrThreads, kPrSmem, s>>>(p, k.dPos, k.dMhi, k.dZ, k.dKbBox, t, c0, m, q, k.dMmaCount);
            return cudaSuccess;
        };
        switch (p.mode) {
        case PB_KRIGING_SPHERICAL: e = launch(kriging_variance_pair_kernel<PB_KRIGING_SPHERICAL>); break;
        case PB_KRIGING_EXPONENTIAL: e = launch(kriging_variance_pair_kernel<PB_KRIGING_EXPONENTIAL>); break;
        default: e = launch(kriging_variance_pair_kernel<PB_KRIGING_GAUSSIAN>); break;
        }
    }
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace pb
