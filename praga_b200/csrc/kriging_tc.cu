// kriging_tc.cu — K5: the kriging variance contraction on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Reference arithmetic being replaced (agrolib/interpolation/kriging.cpp): krigingSetWeight :256-261
// (weight_i = sum_j Vinv[i][j] D_j, 2 n (n+1) flop per target) followed by krigingRMSE :287-295 (sqrt|sum_i weight_i D_i|).
//
// Formulation (kriging.cu header): y = M c with M = L^-1 (lower triangular, C = sill - Gamma = L L^T) and
// c_j = sill - gamma(|p - s_j|);  sum_i weight_i D_i = sill - (|y|^2 - (z.y)(z.y - 1) / |z|^2),  z = M 1.
//
// GEMM mapping: D[cell][i] = sum_j A[cell][j] * B[i][j]
//   A (MMA M = 128 cells x K): c values GENERATED on the fly by the CTA's threads (FP32 variogram of centred FP32
//     distances), split into FP16 hi + lo, written to shared memory in the canonical K-major no-swizzle UMMA layout;
//   B (MMA N = 128 stations i x K): M pre-split into FP16 hi + lo and pre-tiled on the host in the same canonical layout,
//     one 1-D bulk async copy (TMA engine) per 16 KiB tile, completion on an mbarrier; only tiles on or below the
//     diagonal exist (M is lower triangular) so half of the K range is skipped;
//   D: FP32 accumulators in TMEM, 4 N-tiles = 512 columns = one pass over 512 stations i; three MMAs per K step
//     (hi*hi + hi*lo + lo*hi) give ~22 mantissa bits on each product.
//   Epilogue: tcgen05.ld, thread = cell (TMEM lane), per-thread sums of y^2 and z*y, no cross-lane reduction.
#include <stdlib.h>
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
// LBO = 128 B between the two core matrices of one K = 16 step, SBO = 1024 B between 8-row groups; version 1 (Blackwell)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    return uint64_t((saddr & 0x3FFFFu) >> 4) | (uint64_t(128 >> 4) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46);
}
// instruction descriptor: D = F32, A = B = F16, both K-major, N = 128, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (uint32_t(kTcNT >> 3) << 17) | (uint32_t(kTcCells >> 4) << 24);

__host__ __device__ inline size_t tileOffset(int r, int k) { return size_t(r >> 3) * 1024 + size_t(k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2; }

struct TcParams {
    int n, nPad, nKb, mode;
    float range, sn;            // sn = sill - nugget
    double sill, zz;
};

// covariance c = sill - gamma(h) in FP32 (kriging.cpp:160-192 rewritten for c)
__device__ __forceinline__ float covariance(const TcParams& p, float h)
{
    const float t = h / p.range;
    switch (p.mode) {
    case PB_KRIGING_SPHERICAL: return (h < p.range) ? p.sn * (1.f - 1.5f * t + 0.5f * t * t * t) : 0.f;
    case PB_KRIGING_EXPONENTIAL: return p.sn * __expf(-3.f * t);
    default: return p.sn * __expf(-4.f * t * t);
    }
}

constexpr int kTcThreads = 256;
constexpr int kTcSmem = 2 * kTcTileBytes + 8 * kTcTileBytes + 64 * 8 + 512 * 4 + 2 * 2 * kTcCells * 8;

__global__ void __launch_bounds__(kTcThreads, 1)
kriging_variance_tc_kernel(TcParams p, const double* __restrict__ pos, const unsigned char* __restrict__ Mt,
                           const float* __restrict__ z, KrigingTargets t, long long c0, int m, double* __restrict__ q)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem;                                   // [2 pieces][16 KiB]
    unsigned char* sB = smem + 2 * kTcTileBytes;                // [4 N tiles][2 pieces][16 KiB]
    float2* sRel = reinterpret_cast<float2*>(sB + 8 * kTcTileBytes);
    float* sZ = reinterpret_cast<float*>(sRel + 64);
    double* sRed = reinterpret_cast<double*>(sZ + 512);         // [2 halves][2 sums][128 cells]
    __shared__ __align__(8) uint64_t bBar, mmaBar;
    __shared__ uint32_t tmemBase;
    __shared__ double sCenter[2];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long tile0 = (long long)blockIdx.x * kTcCells;
    if (tid == 0) {
        mbar_init(&bBar, 1);
        mbar_init(&mmaBar, 1);
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

    // this thread's cell (= MMA row = TMEM lane) and its half of the K block / of the accumulator columns
    const int r = tid & (kTcCells - 1), half = tid >> 7;
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
        xr = float(x - cx);
        yr = float(y - cy);
    }
    double accY2 = 0., accZY = 0.;
    uint32_t it = 0;
    const int nPasses = (p.nPad + kTcPass - 1) / kTcPass;

    for (int pass = 0; pass < nPasses; pass++) {
        const int i0p = pass * kTcPass;
        const int nT = min(kTcPass / kTcNT, (p.nPad - i0p) / kTcNT);
        const int kbEnd = min(p.nKb, (i0p + nT * kTcNT) / kTcBK);
        for (int i = tid; i < kTcPass; i += kTcThreads) sZ[i] = (i0p + i < p.nPad) ? z[i0p + i] : 0.f;

        for (int kb = 0; kb < kbEnd; kb++) {
            const int j0 = kb * kTcBK;
            if (tid < kTcBK) {
                const int j = j0 + tid;
                sRel[tid] = (j < p.n) ? make_float2(float(pos[2 * j] - cx), float(pos[2 * j + 1] - cy))
                                      : make_float2(1e18f, 1e18f);      // padding: covariance 0 against zero columns of M
            }
            if (tid == 0) {
                int cnt = 0;
                for (int tt = 0; tt < nT; tt++) cnt += (j0 <= i0p + tt * kTcNT + kTcNT - 1);
                mbar_expect_tx(&bBar, uint32_t(cnt) * 2u * kTcTileBytes);
                for (int tt = 0; tt < nT; tt++)
                    if (j0 <= i0p + tt * kTcNT + kTcNT - 1) {
                        const size_t itile = size_t(i0p / kTcNT + tt);
                        bulk_g2s(sB + size_t(tt) * 2 * kTcTileBytes, Mt + (itile * p.nKb + kb) * 2 * kTcTileBytes,
                                 2 * kTcTileBytes, &bBar);
                    }
            }
            __syncthreads();
            // ---- generate the A tile: 128 cells x 64 stations, this thread: row r, 32 stations of its half ----
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const int k0 = half * 32 + g * 8;
                __half2 hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    float c2[2];
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const float2 s = sRel[k0 + 2 * e + u];
                        const float dx = s.x - xr, dy = s.y - yr;
                        c2[u] = covariance(p, sqrtf(dx * dx + dy * dy));
                    }
                    const __half2 h2 = __floats2half2_rn(c2[0], c2[1]);
                    const float2 back = __half22float2(h2);
                    hi[e] = h2;
                    lo[e] = __floats2half2_rn(c2[0] - back.x, c2[1] - back.y);
                }
                const size_t off = tileOffset(r, k0);
                *reinterpret_cast<uint4*>(sA + off) = *reinterpret_cast<const uint4*>(hi);
                *reinterpret_cast<uint4*>(sA + kTcTileBytes + off) = *reinterpret_cast<const uint4*>(lo);
            }
            fence_proxy_async_smem();       // generic-proxy stores -> visible to the tensor core (async proxy)
            __syncthreads();
            if (tid == 0) {
                mbar_wait(&bBar, it & 1);
                tc_fence_after();
                const uint32_t aHi = smem_u32(sA), aLo = aHi + kTcTileBytes;
                for (int tt = 0; tt < nT; tt++) {
                    if (!(j0 <= i0p + tt * kTcNT + kTcNT - 1)) continue;
                    const uint32_t bHi = smem_u32(sB) + uint32_t(tt) * 2 * kTcTileBytes, bLo = bHi + kTcTileBytes;
                    const uint32_t d = tmem + uint32_t(tt * kTcNT);
#pragma unroll
                    for (int ks = 0; ks < kTcBK / 16; ks++) {
                        const uint32_t o = ks * 256;
                        umma_f16(d, umma_desc(aHi + o), umma_desc(bHi + o), kIdesc, (kb > 0 || ks > 0) ? 1u : 0u);
                        umma_f16(d, umma_desc(aHi + o), umma_desc(bLo + o), kIdesc, 1u);
                        umma_f16(d, umma_desc(aLo + o), umma_desc(bHi + o), kIdesc, 1u);
                    }
                }
                umma_commit(&mmaBar);
            }
            mbar_wait(&mmaBar, it & 1);     // every thread: the MMAs have consumed this block's shared memory
            tc_fence_after();
            it++;
        }

        // ---- epilogue of the pass: thread = cell (TMEM lane), columns = stations i of this pass ----
        const uint32_t laneBase = uint32_t((warp & 3) * 32) << 16;
        for (int cb = 0; cb < (kTcPass / 2) / 32; cb++) {
            const int col = half * (kTcPass / 2) + cb * 32;
            if (col >= nT * kTcNT) break;
            uint32_t v[32];
            tmem_ld32(tmem + laneBase + uint32_t(col), v);
            float y2 = 0.f, zy = 0.f;
#pragma unroll
            for (int e = 0; e < 32; e++) {
                const float y = __uint_as_float(v[e]);
                y2 = fmaf(y, y, y2);
                zy = fmaf(sZ[col + e], y, zy);
            }
            accY2 += double(y2);
            accZY += double(zy);
        }
        tc_fence_before();
        __syncthreads();        // TMEM and sZ are rewritten by the next pass
    }

    sRed[(half * 2 + 0) * kTcCells + r] = accY2;
    sRed[(half * 2 + 1) * kTcCells + r] = accZY;
    __syncthreads();
    if (half == 0 && valid) {
        const double y2 = sRed[0 * kTcCells + r] + sRed[2 * kTcCells + r];
        const double zy = sRed[1 * kTcCells + r] + sRed[3 * kTcCells + r];
        q[cidx] = p.sill - (y2 - zy * (zy - 1.) / p.zz);      // = sum_i weight_i D_i of krigingRMSE
    }
    if (warp == 0) tmem_dealloc(tmem, 512);
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
            const float cj = covariance(p, sqrtf(dx * dx + dy * dy));
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
int krigingUploadWhitened(KrigingSystem& k, const std::vector<double>& M, const std::vector<double>& z, double zz)
{
    const int n = k.n;
    const int nPad = (n + kTcNT - 1) / kTcNT * kTcNT;
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
    p.sn = float(k.params.sillNugget);
    p.sill = k.params.sill;
    p.zz = k.zz;
    static const bool simple = getenv("PB_KRIGING_SIMPLE") && atoi(getenv("PB_KRIGING_SIMPLE")) != 0;
    if (simple) {
        kriging_variance_simple_kernel<<<(m + 127) / 128, 128, 0, s>>>(p, k.dPos, k.dMhi, k.dZ, t, c0, m, q);
        return cudaGetLastError();
    }
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(kriging_variance_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    const int grid = (m + kTcCells - 1) / kTcCells;
    kriging_variance_tc_kernel<<<grid, kTcThreads, kTcSmem, s>>>(p, k.dPos, k.dMhi, k.dZ, t, c0, m, q);
    return cudaGetLastError();
}

}  // namespace pb
