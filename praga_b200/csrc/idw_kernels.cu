// idw_kernels.cu — K1: detrended inverse-distance-weighting estimator for raster cells and for point targets.
//
// Replaces, per target, the reference chain (paths relative to the reference tree, agrolib/):
//   gis::getUtmXYFromRowColSinglePrecision      gis/gis.cpp:799-804
//   getProxyValuesXY / getValueFromXY           interpolation/interpolation.cpp:2620-2647, gis/gis.cpp:533-546
//   interpolate                                 interpolation/interpolation.cpp:2502-2560
//     computeDistances                          interpolation.cpp:208-256   (no topographic distance here)
//     inverseDistanceWeighted                   interpolation.cpp:1031-1051
//     retrend                                   interpolation.cpp:1288-1351 (+ getMultipleDetrendingValues :2563-2617,
//                                               functionSum / lapseRatePiecewise_* mathFunctions/furtherMathFunctions.cpp:115-201)
//     per-variable clamp                        interpolation.cpp:2540-2558
//   gis::updateMinMaxRasterGrid                 gis/gis.cpp:569-614 (fused min/max)
//
// B200 mapping (measured with tools/ubench_fp32.cu, profiles/r01_ubench_fp32.jsonl): the station loop costs
// 8 FP32 ops + 1 MUFU.RSQ per (cell, station) pair; 128 FP32 lanes/clk/SM and 16 MUFU/clk/SM make both pipes
// limit at 0.0625 clk/pair/SM.  Packed FFMA2/FADD2/FMUL2 (two cells per lane) halve the issue slots so that the
// LDS and MUFU issue fit beside them.  Station tiles arrive in shared memory through 1-D bulk async copies
// (TMA engine) double-buffered on mbarriers; every lane reads the same station (LDS broadcast).
// Accumulation: FP32 inside a 512-station tile, FP64 across tiles; station values are centred on the host
// (IDW is affine invariant), so the FP32 error scales with the spread of the residuals, not their level.
#include <algorithm>
#include "epilogue.cuh"

namespace pb {

constexpr int kTileStations = 512;   // stations per shared-memory tile (12 KiB)
constexpr int kThreads = 256;
constexpr int kChunk = 32;           // stations per f32 accumulation chain
constexpr int kTileBytes = kTileStations * (int)(sizeof(float4) + sizeof(float2));

// ---------------------------------------------------------------------------------------------------------
// valid-cell compaction: ascending list of linear indices of cells with !isEqual(z, flag)
// (interpolationCmd.cpp:377).  Built once per DEM upload; deterministic order.
// ---------------------------------------------------------------------------------------------------------
constexpr int kCompactChunk = 4096;  // cells per block

__global__ void __launch_bounds__(256) count_valid_kernel(const float* __restrict__ dem, long long n, float flag,
                                                          int* __restrict__ blockCount)
{
    long long base = (long long)blockIdx.x * kCompactChunk;
    int cnt = 0;
    for (int i = threadIdx.x; i < kCompactChunk; i += 256) {
        long long k = base + i;
        if (k < n && !isEqualF(dem[k], flag)) cnt++;
    }
    __shared__ int red[8];
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < 8; w++) s += red[w];
        blockCount[blockIdx.x] = s;
    }
}

// single-block exclusive scan of blockCount (nBlocks <= a few 10^5); total written to *total
__global__ void __launch_bounds__(1024) scan_counts_kernel(int* __restrict__ blockCount, int nBlocks,
                                                           long long* __restrict__ total)
{
    __shared__ long long carry;
    __shared__ int warpSum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nBlocks; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < nBlocks ? blockCount[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) warpSum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = warpSum[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += y;
            }
            warpSum[threadIdx.x] = w;
        }
        __syncthreads();
        int prefix = (threadIdx.x >> 5) ? warpSum[(threadIdx.x >> 5) - 1] : 0;
        long long excl = carry + prefix + x - v;
        if (i < nBlocks) blockCount[i] = (int)excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry += prefix + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(256) write_valid_kernel(const float* __restrict__ dem, long long n, float flag,
                                                          const int* __restrict__ blockOffset,
                                                          int* __restrict__ validIdx, int indexBase)
{
    // one warp handles 32 consecutive cells per step so the output order stays ascending
    long long base = (long long)blockIdx.x * kCompactChunk;
    __shared__ int warpBase[8];
    __shared__ int warpCnt[8];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // each warp owns a contiguous 512-cell slice of the chunk
    long long wbase = base + warp * (kCompactChunk / 8);
    int cnt = 0;
    for (int s = 0; s < kCompactChunk / 8; s += 32) {
        long long k = wbase + s + lane;
        bool ok = k < n && !isEqualF(dem[k], flag);
        cnt += __popc(__ballot_sync(0xffffffffu, ok));
    }
    if (lane == 0) warpCnt[warp] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = blockOffset[blockIdx.x];
        for (int w = 0; w < 8; w++) { warpBase[w] = acc; acc += warpCnt[w]; }
    }
    __syncthreads();
    int pos = warpBase[warp];
    for (int s = 0; s < kCompactChunk / 8; s += 32) {
        long long k = wbase + s + lane;
        bool ok = k < n && !isEqualF(dem[k], flag);
        unsigned b = __ballot_sync(0xffffffffu, ok);
        if (ok) validIdx[pos + __popc(b & ((1u << lane) - 1u))] = indexBase + (int)k;
        pos += __popc(b);
    }
}

// Band pipeline of the drop-in call: ONE pass over a band of the DEM writes the flag into the band of the output grid
// and appends the band's valid cells to its list (slots reserved with one atomicAdd per warp: the list is ascending inside
// every 512-cell slice, the slices land in arrival order, which the kernels do not care about). *count must be zero.
__global__ void __launch_bounds__(256) compact_fill_kernel(const float* __restrict__ dem, long long n, float flag,
                                                           int* __restrict__ validIdx, int indexBase,
                                                           unsigned long long* __restrict__ count, float* __restrict__ outBand)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long wbase = (long long)blockIdx.x * kCompactChunk + warp * (kCompactChunk / 8);
    unsigned masks[kCompactChunk / 8 / 32];
    int cnt = 0;
#pragma unroll
    for (int s = 0; s < kCompactChunk / 8 / 32; s++) {
        const long long k = wbase + s * 32 + lane;
        bool ok = false;
        if (k < n) {
            ok = !isEqualF(dem[k], flag);
            outBand[k] = flag;
        }
        masks[s] = __ballot_sync(0xffffffffu, ok);
        cnt += __popc(masks[s]);
    }
    if (cnt == 0) return;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(count, (unsigned long long)cnt);
    base = __shfl_sync(0xffffffffu, base, 0);
    int pos = 0;
#pragma unroll
    for (int s = 0; s < kCompactChunk / 8 / 32; s++) {
        const unsigned b = masks[s];
        if (b & (1u << lane)) validIdx[base + pos + __popc(b & ((1u << lane) - 1u))] = indexBase + int(wbase + s * 32 + lane);
        pos += __popc(b);
    }
}

// exact evaluation of computeDistances + inverseDistanceWeighted for ONE target, reference operation order
// (f32 distance, f64 weights). Used when the fast path saw a zero distance (station on the target).
__device__ __noinline__ void idwExact(const DevStations& st, float x, float y, double& sw, double& swv)
{
    sw = 0.;
    swv = 0.;
    for (int i = 0; i < st.n; i++) {
        const float dx = __fsub_rn(st.x[i], x), dy = __fsub_rn(st.y[i], y);
        const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
        if (d > 0.f) {
            const double dk = double(d) / 10000.;
            const double w = 1.0 / (dk * dk * dk);
            sw += w;
            swv += double(st.val[i]) * w;
        }
    }
}

__device__ __forceinline__ unsigned orderedKey(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ---------------------------------------------------------------------------------------------------------
// K1 kernel. POINTS = false: targets are DEM cells listed in validIdx; POINTS = true: explicit targets.
// Each thread owns 2*P targets (P packed pairs).
// ---------------------------------------------------------------------------------------------------------
template <int P, bool POINTS>
__global__ void __launch_bounds__(kThreads)
idw_kernel(DevStations st, DevModel m, DevGrid dem, const int* __restrict__ validIdx, int nTargets,
           const float* __restrict__ tx, const float* __restrict__ ty, const double* __restrict__ tproxy,
           float* __restrict__ out, unsigned* __restrict__ minmax /* [0]=min key, [1]=max key */)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar[2];
    float4* sxy0 = reinterpret_cast<float4*>(smem);
    float2* sv0 = reinterpret_cast<float2*>(smem + kTileStations * sizeof(float4));
    float4* sxy1 = reinterpret_cast<float4*>(smem + kTileBytes);
    float2* sv1 = reinterpret_cast<float2*>(smem + kTileBytes + kTileStations * sizeof(float4));

    const int tid = threadIdx.x;
    const int nTiles = st.nPadded / kTileStations;
    const int usedPadded = (st.n + kChunk - 1) / kChunk * kChunk;     // stations rounded up to whole accumulation chunks
    constexpr int C = 2 * P;
    PB_DEV_ASSERT(st.nPadded % kTileStations == 0 && st.n <= st.nPadded);

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int t = 0; t < 2 && t < nTiles; t++) {
            mbar_expect_tx(&bar[t], kTileBytes);
            bulk_g2s(t ? sxy1 : sxy0, st.xy + (size_t)t * kTileStations, kTileStations * sizeof(float4), &bar[t]);
            bulk_g2s(t ? sv1 : sv0, st.v + (size_t)t * kTileStations, kTileStations * sizeof(float2), &bar[t]);
        }
    }

    // ---- this thread's targets ----
    const long long t0 = (long long)blockIdx.x * (kThreads * C) + tid;
    float x[C], y[C];
    int cell[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
        const long long t = t0 + (long long)c * kThreads;
        if (t < nTargets) {
            if (POINTS) {
                cell[c] = (int)t;
                x[c] = tx[t];
                y[c] = ty[t];
            } else {
                const int idx = validIdx[t];
                const int row = idx / dem.nrCols, col = idx - row * dem.nrCols;
                cell[c] = idx;
                // getUtmXYFromRowColSinglePrecision(header,...) gis.cpp:799-804: corner + (0.5 m, -0.5 m), f64 -> f32
                x[c] = float(__dadd_rn(__dadd_rn(dem.llx, __dmul_rn(dem.cellSize, double(col))), 0.5));
                y[c] = float(__dsub_rn(__dadd_rn(dem.lly, __dmul_rn(dem.cellSize, double(dem.nrRows - row))), 0.5));
            }
        } else {
            cell[c] = -1;
            x[c] = 0.f;
            y[c] = 0.f;
        }
    }
    f32x2 nx[P], ny[P];
#pragma unroll
    for (int p = 0; p < P; p++) {
        nx[p] = pack2(-x[2 * p], -x[2 * p + 1]);
        ny[p] = pack2(-y[2 * p], -y[2 * p + 1]);
    }

    double SW[C], SWV[C];
#pragma unroll
    for (int c = 0; c < C; c++) { SW[c] = 0.; SWV[c] = 0.; }

    // ---- station loop ----
    for (int t = 0; t < nTiles; t++) {
        const int b = t & 1;
        mbar_wait(&bar[b], (t >> 1) & 1);
        const float4* __restrict__ sxy = b ? sxy1 : sxy0;
        const float2* __restrict__ sv = b ? sv1 : sv0;
        // three-level accumulation (measured against the reference's f64 sums, tests/test_accumulation.py):
        // f32 over a 32-station chunk, f32 over the 16 chunks of a tile, f64 across tiles -> max error ~1.5e-5
        // where one f32 chain over 512 stations reaches 1.5e-4.
        f32x2 tsw[P], tswv[P];
#pragma unroll
        for (int p = 0; p < P; p++) { tsw[p] = pack2(0.f, 0.f); tswv[p] = pack2(0.f, 0.f); }
        // chunks that hold only padding (zero-weight stations) add +0 to every sum: skipped, same bits
        const int tileLim = min(kTileStations, usedPadded - t * kTileStations);
#pragma unroll 1
        for (int c0 = 0; c0 < tileLim; c0 += kChunk) {
            f32x2 sw[P], swv[P];
#pragma unroll
            for (int s = 0; s < kChunk; s++) {
                const float4 pxy = sxy[c0 + s];
                const float2 pv = sv[c0 + s];
                const f32x2 px = pack2(pxy.x, pxy.y), py = pack2(pxy.z, pxy.w), pvv = pack2(pv.x, pv.y);
#pragma unroll
                for (int p = 0; p < P; p++) {
                    const f32x2 dx = add2(px, nx[p]), dy = add2(py, ny[p]);
                    const f32x2 d2 = fma2(dx, dx, mul2(dy, dy));
                    float a, bq;
                    unpack2(d2, a, bq);
                    if (POINTS) {   // leave-one-out: the target IS a station, its zero distance must weigh nothing (:1039)
                        a = a > 0.f ? a : __int_as_float(0x7f800000);
                        bq = bq > 0.f ? bq : __int_as_float(0x7f800000);
                    }
                    const f32x2 r = pack2(rsqrt_approx(a), rsqrt_approx(bq));
                    const f32x2 w = mul2(mul2(r, r), r);       // 1/d^3 (the 1/10000 scale cancels in the ratio)
                    if (s == 0) {
                        sw[p] = w;
                        swv[p] = mul2(w, pvv);
                    } else {
                        sw[p] = add2(sw[p], w);
                        swv[p] = fma2(w, pvv, swv[p]);
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < P; p++) { tsw[p] = add2(tsw[p], sw[p]); tswv[p] = add2(tswv[p], swv[p]); }
        }
#pragma unroll
        for (int p = 0; p < P; p++) {
            float a, bq, va, vb;
            unpack2(tsw[p], a, bq);
            unpack2(tswv[p], va, vb);
            SW[2 * p] += double(a);
            SW[2 * p + 1] += double(bq);
            SWV[2 * p] += double(va);
            SWV[2 * p + 1] += double(vb);
        }
        __syncthreads();   // everyone is done with buffer b
        if (tid == 0 && t + 2 < nTiles) {
            mbar_expect_tx(&bar[b], kTileBytes);
            bulk_g2s(b ? sxy1 : sxy0, st.xy + (size_t)(t + 2) * kTileStations, kTileStations * sizeof(float4), &bar[b]);
            bulk_g2s(b ? sv1 : sv0, st.v + (size_t)(t + 2) * kTileStations, kTileStations * sizeof(float2), &bar[b]);
        }
    }

    // ---- epilogue ----
    float vmin = INFINITY, vmax = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < C; c++) {
        if (cell[c] < 0) continue;
        double sw = SW[c], swv = SWV[c];
        if (!(sw <= 1.0e300) || !(fabs(swv) <= 1.0e300)) idwExact(st, x[c], y[c], sw, swv);   // inf/NaN: zero distance seen

        double pv[PB_MAX_PROXY];
        if (POINTS) {
            for (int i = 0; i < m.nProxy; i++) pv[i] = tproxy[(size_t)cell[c] * m.nProxy + i];
        } else if (m.useDetrendingVar) {
            // getProxyValuesXY (interpolation.cpp:2620-2647)
            for (int i = 0; i < m.nProxy; i++) {
                pv[i] = double(kNoData);
                const DevProxy& p = m.proxy[i];
                if (p.active && p.gridSlot >= 0) {
                    const DevGrid& g = m.grid[p.gridSlot];
                    const float v = gridValueFromXY(g, double(x[c]), double(y[c]));
                    if (v != g.flag) pv[i] = double(v);
                }
            }
        } else {
            for (int i = 0; i < m.nProxy; i++) pv[i] = double(kNoData);
        }

        float result;
        if (m.useRetrendOnly) result = 0.f;
        else result = (sw > 0.0) ? float(swv / sw + m.valueOffset) : kNoData;
        const float o = finishEstimate(m, result, pv);
        out[cell[c]] = o;
        if (!isEqualF(o, dem.flag) && !isEqualF(o, kNoData)) {
            vmin = fminf(vmin, o);
            vmax = fmaxf(vmax, o);
        }
    }
    if (minmax) {
        for (int o = 16; o; o >>= 1) {
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        }
        if ((tid & 31) == 0) {
            if (vmin != INFINITY) atomicMin(&minmax[0], orderedKey(vmin));
            if (vmax != -INFINITY) atomicMax(&minmax[1], orderedKey(vmax));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// K1, persistent form (the raster path whenever the whole padded station set fits in shared memory, i.e. up to
// kResidentMaxStations stations): ONE CTA per SM stays resident, loads every station tile once (1-D bulk copies
// on one mbarrier) and then each WARP pulls work items of 32 x 2P consecutive valid cells from a global counter.
// No CTA-wide barrier in the loop: warps drift apart, so one warp's epilogue (dependent DEM / proxy-grid loads,
// f64 retrend) overlaps the other warps' station loops, and the tail costs one 128-cell item instead of a wave
// of 1024-cell CTAs. Arithmetic and accumulation order are those of idw_kernel (bit-identical results).
// ---------------------------------------------------------------------------------------------------------
constexpr int kResidentMaxStations = 9216;     // 18 tiles x 12 KiB = 216 KiB of the 227 KiB a CTA may own

template <int P, int kPersistThreads>
__global__ void __launch_bounds__(kPersistThreads, 1)
idw_persistent_kernel(DevStations st, DevModel m, DevGrid dem, const int* __restrict__ validIdx, int nTargetsHost,
                      float* __restrict__ out, unsigned* __restrict__ minmax, unsigned* __restrict__ workCounter,
                      const long long* __restrict__ nTargetsDev)
{
    // band pipeline of the drop-in call: the number of valid cells of a band is only known on the device
    const int nTargets = nTargetsDev ? int(*nTargetsDev) : nTargetsHost;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    const float4* __restrict__ sxyAll = reinterpret_cast<const float4*>(smem);
    const float2* __restrict__ svAll = reinterpret_cast<const float2*>(smem + (size_t)st.nPadded * sizeof(float4));
    const int tid = threadIdx.x, lane = tid & 31;
    const int nTiles = st.nPadded / kTileStations;
    const int usedPadded = (st.n + kChunk - 1) / kChunk * kChunk;
    constexpr int C = 2 * P;
    // every station tile is resident: the launch sized the dynamic shared memory for nPadded stations
    PB_DEV_ASSERT(st.nPadded % kTileStations == 0 && st.nPadded <= kResidentMaxStations && st.n <= st.nPadded);

    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, (uint32_t)st.nPadded * (uint32_t)(sizeof(float4) + sizeof(float2)));
        for (int t = 0; t < nTiles; t++) {
            bulk_g2s(smem + (size_t)t * kTileStations * sizeof(float4), st.xy + (size_t)t * kTileStations,
                     kTileStations * sizeof(float4), &bar);
            bulk_g2s(smem + (size_t)st.nPadded * sizeof(float4) + (size_t)t * kTileStations * sizeof(float2),
                     st.v + (size_t)t * kTileStations, kTileStations * sizeof(float2), &bar);
        }
    }
    mbar_wait(&bar, 0);

    float vmin = INFINITY, vmax = -INFINITY;
    for (;;) {
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(workCounter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        const long long base = (long long)item * (32 * C);
        if (base >= nTargets) break;

        float x[C], y[C];
        int cell[C];
#pragma unroll
        for (int c = 0; c < C; c++) {
            const long long t = base + c * 32 + lane;
            if (t < nTargets) {
                const int idx = validIdx[t];
                const int row = idx / dem.nrCols, col = idx - row * dem.nrCols;
                cell[c] = idx;
                // getUtmXYFromRowColSinglePrecision(header,...) gis.cpp:799-804: corner + (0.5 m, -0.5 m), f64 -> f32
                x[c] = float(__dadd_rn(__dadd_rn(dem.llx, __dmul_rn(dem.cellSize, double(col))), 0.5));
                y[c] = float(__dsub_rn(__dadd_rn(dem.lly, __dmul_rn(dem.cellSize, double(dem.nrRows - row))), 0.5));
            } else {
                cell[c] = -1;
                x[c] = 0.f;
                y[c] = 0.f;
            }
        }
        f32x2 nx[P], ny[P];
#pragma unroll
        for (int p = 0; p < P; p++) {
            nx[p] = pack2(-x[2 * p], -x[2 * p + 1]);
            ny[p] = pack2(-y[2 * p], -y[2 * p + 1]);
        }
        double SW[C], SWV[C];
#pragma unroll
        for (int c = 0; c < C; c++) { SW[c] = 0.; SWV[c] = 0.; }

#pragma unroll 1
        for (int t = 0; t < nTiles; t++) {
            const float4* __restrict__ sxy = sxyAll + (size_t)t * kTileStations;
            const float2* __restrict__ sv = svAll + (size_t)t * kTileStations;
            f32x2 tsw[P], tswv[P];
#pragma unroll
            for (int p = 0; p < P; p++) { tsw[p] = pack2(0.f, 0.f); tswv[p] = pack2(0.f, 0.f); }
            const int tileLim = min(kTileStations, usedPadded - t * kTileStations);      // padding-only chunks add +0: skipped
#pragma unroll 1
            for (int c0 = 0; c0 < tileLim; c0 += kChunk) {
                f32x2 sw[P], swv[P];
#pragma unroll
                for (int s = 0; s < kChunk; s++) {
                    const float4 pxy = sxy[c0 + s];
                    const float2 pv = sv[c0 + s];
                    const f32x2 px = pack2(pxy.x, pxy.y), py = pack2(pxy.z, pxy.w), pvv = pack2(pv.x, pv.y);
#pragma unroll
                    for (int p = 0; p < P; p++) {
                        const f32x2 dx = add2(px, nx[p]), dy = add2(py, ny[p]);
                        const f32x2 d2 = fma2(dx, dx, mul2(dy, dy));
                        float a, bq;
                        unpack2(d2, a, bq);
                        const f32x2 r = pack2(rsqrt_approx(a), rsqrt_approx(bq));
                        const f32x2 w = mul2(mul2(r, r), r);       // 1/d^3 (the 1/10000 scale cancels in the ratio)
                        if (s == 0) {
                            sw[p] = w;
                            swv[p] = mul2(w, pvv);
                        } else {
                            sw[p] = add2(sw[p], w);
                            swv[p] = fma2(w, pvv, swv[p]);
                        }
                    }
                }
#pragma unroll
                for (int p = 0; p < P; p++) { tsw[p] = add2(tsw[p], sw[p]); tswv[p] = add2(tswv[p], swv[p]); }
            }
#pragma unroll
            for (int p = 0; p < P; p++) {
                float a, bq, va, vb;
                unpack2(tsw[p], a, bq);
                unpack2(tswv[p], va, vb);
                SW[2 * p] += double(a);
                SW[2 * p + 1] += double(bq);
                SWV[2 * p] += double(va);
                SWV[2 * p + 1] += double(vb);
            }
        }

#pragma unroll 1
        for (int c = 0; c < C; c++) {
            if (cell[c] < 0) continue;
            double sw = SW[c], swv = SWV[c];
            if (!(sw <= 1.0e300) || !(fabs(swv) <= 1.0e300)) idwExact(st, x[c], y[c], sw, swv);   // inf/NaN: zero distance seen
            double pv[PB_MAX_PROXY];
            if (m.useDetrendingVar) {
                // getProxyValuesXY (interpolation.cpp:2620-2647)
                for (int i = 0; i < m.nProxy; i++) {
                    pv[i] = double(kNoData);
                    const DevProxy& p = m.proxy[i];
                    if (p.active && p.gridSlot >= 0) {
                        const DevGrid& g = m.grid[p.gridSlot];
                        const float v = gridValueFromXY(g, double(x[c]), double(y[c]));
                        if (v != g.flag) pv[i] = double(v);
                    }
                }
            } else {
                for (int i = 0; i < m.nProxy; i++) pv[i] = double(kNoData);
            }
            float result;
            if (m.useRetrendOnly) result = 0.f;
            else result = (sw > 0.0) ? float(swv / sw + m.valueOffset) : kNoData;
            const float o = finishEstimate(m, result, pv);
            out[cell[c]] = o;
            if (!isEqualF(o, dem.flag) && !isEqualF(o, kNoData)) {
                vmin = fminf(vmin, o);
                vmax = fmaxf(vmax, o);
            }
        }
    }
    if (minmax) {
        for (int o = 16; o; o >>= 1) {
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        }
        if (lane == 0) {
            if (vmin != INFINITY) atomicMin(&minmax[0], orderedKey(vmin));
            if (vmax != -INFINITY) atomicMax(&minmax[1], orderedKey(vmax));
        }
    }
}

// constant fill of valid cells (precipitationAllZero short-circuit, interpolation.cpp:2506-2507) and
// flag initialisation of the output (Crit3DRasterGrid::initializeGrid(raster), gis.cpp:357-363)
__global__ void fill_kernel(float* __restrict__ out, long long n, float v)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = v;
}
__global__ void fill_valid_kernel(float* __restrict__ out, const int* __restrict__ validIdx, int n, float v)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[validIdx[i]] = v;
}

// first position in the ascending valid list whose cell lies in row >= `row` (lower_bound), for every row boundary
__global__ void row_start_kernel(const int* __restrict__ validIdx, long long total, int nrRows, int nrCols,
                                 long long* __restrict__ rowStart)
{
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row > nrRows) return;
    const long long first = (long long)row * nrCols;
    long long lo = 0, hi = total;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if ((long long)validIdx[mid] < first) lo = mid + 1;
        else hi = mid;
    }
    rowStart[row] = lo;
}

// ---- host-side launchers (called from engine.cu) ----
static int currentSmCount()
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

cudaError_t launchRowStart(const int* validIdx, long long total, int nrRows, int nrCols, long long* rowStart, cudaStream_t s)
{
    row_start_kernel<<<(nrRows + 1 + 255) / 256, 256, 0, s>>>(validIdx, total, nrRows, nrCols, rowStart);
    return cudaGetLastError();
}

cudaError_t launchCompactValid(const float* dem, long long n, float flag, int* blockCount, long long* dTotal,
                               int* validIdx, int phase, cudaStream_t s, int indexBase)
{
    const int nBlocks = (int)((n + kCompactChunk - 1) / kCompactChunk);
    if (phase == 0) {
        count_valid_kernel<<<nBlocks, 256, 0, s>>>(dem, n, flag, blockCount);
        scan_counts_kernel<<<1, 1024, 0, s>>>(blockCount, nBlocks, dTotal);
    } else {
        write_valid_kernel<<<nBlocks, 256, 0, s>>>(dem, n, flag, blockCount, validIdx, indexBase);
    }
    return cudaGetLastError();
}

cudaError_t launchCompactFill(const float* dem, long long n, float flag, int* validIdx, int indexBase, long long* dCount,
                              float* outBand, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    const int nBlocks = (int)((n + kCompactChunk - 1) / kCompactChunk);
    compact_fill_kernel<<<nBlocks, 256, 0, s>>>(dem, n, flag, validIdx, indexBase, reinterpret_cast<unsigned long long*>(dCount), outBand);
    return cudaGetLastError();
}

int compactBlocks(long long n) { return (int)((n + kCompactChunk - 1) / kCompactChunk); }
bool idwResidentFits(int nPaddedStations) { return nPaddedStations <= 9216; }
int tileStations() { return kTileStations; }

cudaError_t launchIdwRaster(const DevStations& st, const DevModel& m, const DevGrid& dem, const int* validIdx,
                            int nTargets, float* out, unsigned* minmax, cudaStream_t s, const long long* nTargetsDev,
                            unsigned* workCounter, int reserveSms)
{
    if (nTargets <= 0) return cudaSuccess;
    if (nTargetsDev && st.nPadded > kResidentMaxStations) return cudaErrorInvalidValue;    // callers check idwResidentFits()
    constexpr int P = 2;
    if (st.nPadded <= kResidentMaxStations) {
        // persistent form: minmax[2] is the work-item counter (zeroed by the caller together with min/max)
        constexpr int kPersistThreads = 1024;     // 32 warps/SM: measured best on B200 (512: -5 % at C2, equal at S = 5000)
        auto k = idw_persistent_kernel<P, kPersistThreads>;
        // the attribute is per device and this launcher runs on several lane threads / devices of one process: set on every
        // launch (a microsecond against a kernel of hundreds), SM count of the CURRENT device
        const int sms = std::max(1, currentSmCount() - std::max(0, reserveSms));      // batch lanes leave SMs to each other's small kernels
        cudaError_t ea = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kResidentMaxStations * 24);
        if (ea != cudaSuccess) return ea;
        const size_t smem = (size_t)st.nPadded * (sizeof(float4) + sizeof(float2));
        const int items = (nTargets + 64 * P - 1) / (64 * P);
        const int grid = std::min(sms, (items + kPersistThreads / 32 - 1) / (kPersistThreads / 32));
        k<<<grid, kPersistThreads, smem, s>>>(st, m, dem, validIdx, nTargets, out, minmax, workCounter ? workCounter : minmax + 2,
                                              nTargetsDev);
        return cudaGetLastError();
    }
    const int per = kThreads * 2 * P;
    const int grid = (nTargets + per - 1) / per;
    auto k = idw_kernel<P, false>;           // 2 * kTileBytes = 24 KiB of dynamic shared memory: no opt-in attribute needed
    k<<<grid, kThreads, 2 * kTileBytes, s>>>(st, m, dem, validIdx, nTargets, nullptr, nullptr, nullptr, out, minmax);
    return cudaGetLastError();
}

template <int P>
static cudaError_t launchIdwPointsT(const DevStations& st, const DevModel& m, const float* tx, const float* ty,
                                    const double* tproxy, int nTargets, float* out, cudaStream_t s)
{
    const int per = kThreads * 2 * P;
    const int grid = (nTargets + per - 1) / per;
    auto k = idw_kernel<P, true>;
    DevGrid none{};
    none.flag = kNoData;
    k<<<grid, kThreads, 2 * kTileBytes, s>>>(st, m, none, nullptr, nTargets, tx, ty, tproxy, out, nullptr);
    return cudaGetLastError();
}

cudaError_t launchIdwPoints(const DevStations& st, const DevModel& m, const float* tx, const float* ty,
                            const double* tproxy, int nTargets, float* out, cudaStream_t s)
{
    if (nTargets <= 0) return cudaSuccess;
    // leave-one-out sets (a few thousand targets) spread over more CTAs with 2 targets per thread; raster-sized target
    // lists (glocal areas) amortise each station load over 4 targets like the raster kernel. Same arithmetic per target.
    if (nTargets >= currentSmCount() * kThreads * 8) return launchIdwPointsT<2>(st, m, tx, ty, tproxy, nTargets, out, s);
    return launchIdwPointsT<1>(st, m, tx, ty, tproxy, nTargets, out, s);
}

cudaError_t launchFill(float* out, long long n, float v, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    fill_kernel<<<currentSmCount() * 8, 256, 0, s>>>(out, n, v);
    return cudaGetLastError();
}
cudaError_t launchFillValid(float* out, const int* validIdx, int n, float v, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    fill_valid_kernel<<<(n + 255) / 256, 256, 0, s>>>(out, validIdx, n, v);
    return cudaGetLastError();
}

}  // namespace pb
