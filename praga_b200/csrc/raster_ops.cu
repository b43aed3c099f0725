// raster_ops.cu — K9: the raster map kernels on either side of the gridding path. All of them are
// one pass over a grid (HBM bound: a few bytes read and 4 written per cell) except the topographic-distance map, which
// is the DEM ray march of K6 from one point to every cell.
//
//   computeRelativeHumidityMap   agrolib/project/meteoMaps.cpp:301-321 with relHumFromTdew, agrolib/meteo/meteo.cpp:301-315
//   fixDailyThermalConsistency   agrolib/project/meteoMaps.cpp:103-126
//   gis::topographicDistanceMap  agrolib/gis/gis.cpp:1648-1672 (gis::topographicDistance :1595-1647)
//   gis::resampleGrid            agrolib/gis/gis.cpp:1722-1811 (statistics::mean, sorting::percentile basicMath.cpp:327-366,
//                                gis::prevailingValue gis.cpp:1513-1554)
//
// Built with -fmad=false: the ray march and the sampling coordinates follow the reference's operation order.
#include <math.h>
#include <string.h>
#include <string>

#include "context.cuh"
#include "nvtx.cuh"

namespace pb {

// gis::topographicDistance (gis.cpp:1595-1647), same code as K6 (td_kernels.cu) for one pair
__device__ __forceinline__ float topoDistancePair(const DevGrid& dem, float x1, float y1, float z1, float x2, float y2, float z2,
                                                  float distance)
{
    const float stepMeter = float(dem.cellSize);
    if (distance < stepMeter) return 0.f;
    const int nrStep = int(distance / stepMeter);
    float Xi, Yi, Zi, Xf, Yf;
    if (z1 < z2) { Xi = x1; Yi = y1; Zi = z1; Xf = x2; Yf = y2; }
    else { Xi = x2; Yi = y2; Zi = z2; Xf = x1; Yf = y1; }
    const float dx = (Xf - Xi) / float(nrStep);
    const float dy = (Yf - Yi) / float(nrStep);
    float x = Xi, y = Yi, maxDeltaZ = 0.f;
    for (int i = 1; i <= nrStep; i++) {
        x = x + dx;
        y = y + dy;
        const int r = int((double(y) - dem.lly) * dem.invCellSize);
        const int row = (dem.nrRows - 1) - r;
        const int col = int((double(x) - dem.llx) * dem.invCellSize);
        float demValue = dem.flag;
        if (unsigned(row) < unsigned(dem.nrRows) && unsigned(col) < unsigned(dem.nrCols))
            demValue = __ldg(dem.data + (size_t)row * dem.nrCols + col);
        if (!isEqualF(demValue, dem.flag))
            if (demValue > Zi) maxDeltaZ = (maxDeltaZ > demValue - Zi) ? maxDeltaZ : demValue - Zi;
    }
    return maxDeltaZ;
}

__device__ __forceinline__ unsigned orderedKeyR(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// block-wide min/max of the valid outputs -> two atomics per block (gis::updateMinMaxRasterGrid, gis.cpp:569-614)
__device__ __forceinline__ void blockMinMax(float v, bool valid, unsigned* minmax)
{
    float vmin = valid ? v : INFINITY, vmax = valid ? v : -INFINITY;
    for (int o = 16; o; o >>= 1) {
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (vmin != INFINITY) atomicMin(&minmax[0], orderedKeyR(vmin));
        if (vmax != -INFINITY) atomicMax(&minmax[1], orderedKeyR(vmax));
    }
}

// isEqual(float, float) = fabs(double(a) - double(b)) < EPSILON (basicMath.cpp) without FP64 on the fast side: the float
// difference is the correctly rounded exact difference, so it only disagrees with the double test inside a hair of EPSILON;
// there (never in practice) the double test decides. The HBM-bound map kernels below run six of these per cell.
__device__ __forceinline__ bool isEqualQ(float a, float b)
{
    const float diff = fabsf(a - b);
    if (diff < 0.99e-5f) return true;
    if (diff > 1.01e-5f) return false;
    return isEqualF(a, b);
}

// relHumFromTdew(float Td, float T), meteo.cpp:301-315: rh = 100 pow(exp(c Td - (c T / (T + d)) (Td + d)), 1 / (T + d)).
// The exponent collapses algebraically to a = c d (Td - T) / (T + d)^2 (the two products cancel term by term), which has no
// cancellation left: a is taken in FP64 (a handful of instructions, the reciprocal from the FP64 seed instruction + one Newton
// step), rounded to float once and the exponential is FP32. Error against the reference's FP64 pow(exp()): the rounding of a
// (6e-8 |a|, i.e. at most 2e-6 RH units at rh = 37) + expf (1-2 ulp of a value <= 100: 1e-5); the parity bar is 1e-4.
// Conversions and special functions share one quarter-rate pipe: this form needs five of them per cell.
__device__ __forceinline__ float relHumFromTdewDev(float Td, float T)
{
    if (isEqualQ(Td, kNoData) || isEqualQ(T, kNoData)) return kNoData;
    const double d = 237.3;
    const double c = 17.2693882;
    const double td = double(T) + d;
    const double q = td * td;
    double x0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(q));           // ~20 bits
    const double x1 = fma(x0, fma(-q, x0, 1.0), x0);                  // ~40 bits
    const double a = ((c * d) * (double(Td) - double(T))) * x1;
    const float rh = expf(float(a)) * 100.f;
    if (rh > 100.f) return 100.f;
    return fmaxf(1.f, rh);       // (NaN inputs: fmaxf returns 1 like std::max(1.f, NaN))
}

// four cells per thread and iteration (128-bit loads / stores), grid-stride over the map, scalar tail; min / max leave the CTA
// as ONE atomic pair (a pair per warp on the same two words serialises in L2: 1.5 M atomics cost more than the 1.2 GB of traffic)
__global__ void __launch_bounds__(256) relative_humidity_map_kernel(const float* __restrict__ airT, float flagT,
                                                                    const float* __restrict__ dewT, float flagTd, long long n,
                                                                    float outFlag, float* __restrict__ out,
                                                                    unsigned* __restrict__ minmax)
{
    float vmin = INFINITY, vmax = -INFINITY;
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long k0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; k0 < n; k0 += stride) {
        float tv[4], dv[4], ov[4];
        const bool full = k0 + 3 < n;
        if (full) {
            *reinterpret_cast<float4*>(tv) = __ldcs(reinterpret_cast<const float4*>(airT + k0));
            *reinterpret_cast<float4*>(dv) = __ldcs(reinterpret_cast<const float4*>(dewT + k0));
        } else {
            for (int q = 0; q < 4; q++) { tv[q] = (k0 + q < n) ? airT[k0 + q] : flagT; dv[q] = (k0 + q < n) ? dewT[k0 + q] : flagTd; }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            float o = outFlag;
            if (!isEqualQ(tv[q], flagT) && !isEqualQ(dv[q], flagTd)) o = relHumFromTdewDev(dv[q], tv[q]);
            ov[q] = o;
            if (k0 + q < n && !isEqualQ(o, outFlag) && !isEqualQ(o, kNoData)) { vmin = fminf(vmin, o); vmax = fmaxf(vmax, o); }
        }
        if (full) *reinterpret_cast<float4*>(out + k0) = *reinterpret_cast<const float4*>(ov);
        else for (int q = 0; q < 4; q++) if (k0 + q < n) out[k0 + q] = ov[q];
    }
    for (int o = 16; o; o >>= 1) {
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    }
    __shared__ float smin[8], smax[8];
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = vmin; smax[threadIdx.x >> 5] = vmax; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) { vmin = fminf(vmin, smin[w]); vmax = fmaxf(vmax, smax[w]); }
        if (vmin != INFINITY) atomicMin(&minmax[0], orderedKeyR(vmin));
        if (vmax != -INFINITY) atomicMax(&minmax[1], orderedKeyR(vmax));
    }
}

__global__ void __launch_bounds__(256) thermal_consistency_kernel(const float* __restrict__ tmax, float flagMax,
                                                                  float* __restrict__ tmin, float flagMin, long long n,
                                                                  unsigned long long* __restrict__ nFixed)
{
    const long long k0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    float a[4], b[4];
    const bool full = k0 + 3 < n;
    if (full) {
        *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(tmax + k0);
        *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(tmin + k0);
    } else {
        for (int q = 0; q < 4; q++) { a[q] = (k0 + q < n) ? tmax[k0 + q] : flagMax; b[q] = (k0 + q < n) ? tmin[k0 + q] : flagMin; }
    }
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < 4; q++)
        if (!isEqualQ(a[q], flagMax) && !isEqualQ(b[q], flagMin) && b[q] > a[q]) {
            b[q] = a[q] - 0.1f;
            cnt++;
        }
    if (cnt) {                                    // only the quads that changed are written back
        if (full) *reinterpret_cast<float4*>(tmin + k0) = *reinterpret_cast<const float4*>(b);
        else for (int q = 0; q < 4; q++) if (k0 + q < n) tmin[k0 + q] = b[q];
    }
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(nFixed, (unsigned long long)cnt);
}

__global__ void __launch_bounds__(128) topographic_distance_map_kernel(DevGrid dem, float px, float py, float pz,
                                                                       float* __restrict__ out)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)dem.nrRows * dem.nrCols;
    if (k >= n) return;
    const float demValue = dem.data[k];
    if (isEqualF(demValue, dem.flag)) { out[k] = dem.flag; return; }
    const int row = int(k / dem.nrCols), col = int(k - (long long)row * dem.nrCols);
    // Crit3DRasterGrid::getXY (gis.cpp:473-477), narrowed to float at the calls
    const double gx = dem.llx + dem.cellSize * (double(col) + 0.5);
    const double gy = dem.lly + dem.cellSize * (double(dem.nrRows - row) - 0.5);
    const float x = float(gx), y = float(gy);
    const float dx = x - px, dy = y - py;                      // gis::computeDistance(x1, y1, x2, y2), gis.cpp:685-691
    const float distance = sqrtf(dx * dx + dy * dy);
    out[k] = topoDistancePair(dem, x, y, demValue, px, py, pz, distance);
}

constexpr int kResampleMax = 256;       // samples per new cell held for the median (nrStep <= 16)

// one thread per cell of the new grid; METHOD: PB_AGGR_AVERAGE / _MEDIAN / _CENTER
template <int METHOD>
__global__ void __launch_bounds__(128) resample_grid_kernel(DevGrid oldGrid, DevGrid newGrid, float nodataRatioThreshold,
                                                            float* __restrict__ out, unsigned* __restrict__ minmax)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)newGrid.nrRows * newGrid.nrCols;
    float o = newGrid.flag;
    bool valid = false;
    if (k < n) {
        const int row = int(k / newGrid.nrCols), col = int(k - (long long)row * newGrid.nrCols);
        const double resampleFactor = newGrid.cellSize / oldGrid.cellSize;
        float value = kNoData;
        // Crit3DRasterGrid::getXY (gis.cpp:473-477)
        const double x0 = newGrid.llx + newGrid.cellSize * (double(col) + 0.5);
        const double y0 = newGrid.lly + newGrid.cellSize * (double(newGrid.nrRows - row) - 0.5);
        if (resampleFactor <= 1. || METHOD == PB_AGGR_CENTER) {
            // getRowCol (gis.cpp:486-493) + isOutOfGridRowCol
            const int r = int((y0 - oldGrid.lly) * oldGrid.invCellSize);
            const int tmpRow = (oldGrid.nrRows - 1) - r;
            const int tmpCol = int((x0 - oldGrid.llx) * oldGrid.invCellSize);
            if (unsigned(tmpRow) < unsigned(oldGrid.nrRows) && unsigned(tmpCol) < unsigned(oldGrid.nrCols)) {
                const float tmpValue = oldGrid.data[(size_t)tmpRow * oldGrid.nrCols + tmpCol];
                if (!isEqualF(tmpValue, oldGrid.flag)) value = tmpValue;
            }
        } else {
            const int nrStep = int(floor(resampleFactor)) + 1;
            const double step = newGrid.cellSize / nrStep;
            const double halfStep = step * 0.5;
            const double llxS = x0 - (newGrid.cellSize / 2) + halfStep;
            const double llyS = y0 - (newGrid.cellSize / 2) + halfStep;
            float values[(METHOD == PB_AGGR_MEDIAN || METHOD == PB_AGGR_PREVAILING) ? kResampleMax : 1];
            int nrValues = 0, maxNrValues = 0, nrMean = 0;
            double sum = 0.;
            for (int i = 0; i < nrStep; i++) {
                const double x = llxS + step * i;
                for (int j = 0; j < nrStep; j++) {
                    const double y = llyS + step * j;
                    // gis::getValueFromXY (gis.cpp:533-546)
                    const int r = int((y - oldGrid.lly) * oldGrid.invCellSize);
                    const int sr = (oldGrid.nrRows - 1) - r;
                    const int sc = int((x - oldGrid.llx) * oldGrid.invCellSize);
                    float tmpValue = oldGrid.flag;
                    if (unsigned(sr) < unsigned(oldGrid.nrRows) && unsigned(sc) < unsigned(oldGrid.nrCols))
                        tmpValue = __ldg(oldGrid.data + (size_t)sr * oldGrid.nrCols + sc);
                    if (!isEqualF(tmpValue, oldGrid.flag)) {
                        if (METHOD == PB_AGGR_MEDIAN || METHOD == PB_AGGR_PREVAILING) values[nrValues] = tmpValue;
                        nrValues++;
                        if (!isEqualF(tmpValue, kNoData)) { sum += double(tmpValue); nrMean++; }     // statistics::mean skips NODATA
                    }
                    maxNrValues++;
                }
            }
            if (maxNrValues > 0 && (float(nrValues) / float(maxNrValues)) > nodataRatioThreshold) {
                if (METHOD == PB_AGGR_AVERAGE) {
                    // statistics::mean(std::vector<float>) (statistics.cpp:1300-1321): double sum of the non-NODATA values
                    value = (nrValues < 1 || nrMean == 0) ? kNoData : float(sum / double(nrMean));
                } else if (METHOD == PB_AGGR_PREVAILING) {
                    // gis::prevailingValue (gis.cpp:1513-1554) when fewer samples are missing than present (:1794-1799): the
                    // first value (in sample order) with the largest number of isEqual matches; a value counts for the FIRST
                    // distinct entry it is isEqual to, like the reference's linear search
                    if (maxNrValues - nrValues < nrValues) {
                        float distinct[kResampleMax];
                        unsigned short counters[kResampleMax];
                        int nd = 0;
                        for (int q = 0; q < nrValues; q++) {
                            bool found = false;
                            for (int d = 0; d < nd && !found; d++)
                                if (isEqualF(values[q], distinct[d])) { found = true; counters[d]++; }
                            if (!found) { distinct[nd] = values[q]; counters[nd] = 1; nd++; }
                        }
                        unsigned maxCount = counters[0];
                        value = distinct[0];
                        for (int d = 1; d < nd; d++)
                            if (counters[d] > maxCount) { maxCount = counters[d]; value = distinct[d]; }
                    }
                } else if (METHOD == PB_AGGR_MEDIAN) {
                    // sorting::percentile(values, n, 50, true) (basicMath.cpp:327-366): < 3 values -> NODATA; drop NODATA, sort
                    // ascending, rank = n * 0.5 - 1, linear interpolation
                    int m = 0;
                    for (int q = 0; q < nrValues; q++) if (!isEqualF(values[q], kNoData)) values[m++] = values[q];
                    if (nrValues >= 3 && m >= 3) {
                        for (int a = 1; a < m; a++) {                     // insertion sort (a few dozen values)
                            const float v = values[a];
                            int b = a - 1;
                            while (b >= 0 && values[b] > v) { values[b + 1] = values[b]; b--; }
                            values[b + 1] = v;
                        }
                        const double rank = m * double(50.0 / 100.f) - 1.0;
                        if (rank < 0.0) value = values[0];
                        else {
                            const int low = int(rank);
                            if (low >= m - 1) value = values[m - 1];
                            else {
                                const double frac = rank - low;
                                value = float(values[low] + frac * (values[low + 1] - values[low]));
                            }
                        }
                    }
                }
            }
        }
        if (!isEqualF(value, kNoData)) o = value;
        out[k] = o;
        valid = !isEqualF(o, newGrid.flag) && !isEqualF(o, kNoData);
    }
    blockMinMax(o, valid, minmax);
}

// ---- Crit3DMeteoGrid::spatialAggregateMeteoGrid (meteo/meteoGrid.cpp:1135-1190) + spatialAggregateMeteoGridPoint (:1193-1240):
// one WARP per meteo-grid cell. The lanes sample the raster at the cell's aggregation points (gis::getValueFromXY, nearest
// cell); lane 0 then reduces them in the list's order with the reference's FLOAT arithmetic (statistics::mean / variance,
// mathFunctions/statistics.cpp:1274-1297, 1160-1186) or sorts them for sorting::percentile (basicMath.cpp:327-366).
__device__ void heapSortAscending(float* v, int n)
{
    auto sift = [&](int root, int end) {
        const float t = v[root];
        int child;
        while ((child = 2 * root + 1) < end) {
            if (child + 1 < end && v[child + 1] > v[child]) child++;
            if (!(v[child] > t)) break;
            v[root] = v[child];
            root = child;
        }
        v[root] = t;
    };
    for (int i = n / 2 - 1; i >= 0; i--) sift(i, n);
    for (int end = n - 1; end > 0; end--) {
        const float t = v[0]; v[0] = v[end]; v[end] = t;
        sift(0, end);
    }
}

// rank-k value (0-based, ascending) of the valid entries of z[0..n) and the value of rank k + 1 (= the same value when it
// repeats, else the smallest larger one; the rank-k value again when k is the last rank), by a 4-pass radix select over the
// order-preserving keys: the two neighbours sorting::percentile (mathFunctions/statistics.cpp) interpolates between, without
// sorting. One warp, all lanes return the same pair. hist = 256 counters of this warp in shared memory.
__device__ __forceinline__ unsigned orderedKeyA(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float keyToFloatA(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ void warpSelectPair(const float* __restrict__ z, int n, int k, unsigned* hist, float& vk, float& vk1)
{
    const int lane = threadIdx.x & 31;
    unsigned prefix = 0u, mask = 0u;
    int cntEq = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int b = lane; b < 256; b += 32) hist[b] = 0u;
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            const float v = z[i];
            if (v != kNoData) {
                const unsigned key = orderedKeyA(v);
                if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
            }
        }
        __syncwarp();
        // lane l owns bins 8l .. 8l+7
        unsigned own[8], sum = 0u;
#pragma unroll
        for (int j = 0; j < 8; j++) { own[j] = hist[lane * 8 + j]; sum += own[j]; }
        unsigned incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const unsigned excl = incl - sum;
        const bool mine = unsigned(k) >= excl && unsigned(k) < incl;
        int bin = 0, before = 0, inBin = 0;
        if (mine) {
            unsigned run = excl;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (unsigned(k) >= run && unsigned(k) < run + own[j]) { bin = lane * 8 + j; before = int(run); inBin = int(own[j]); }
                run += own[j];
            }
        }
        const unsigned who = __ballot_sync(0xffffffffu, mine);
        const int src = __ffs(who) - 1;
        bin = __shfl_sync(0xffffffffu, bin, src);
        before = __shfl_sync(0xffffffffu, before, src);
        inBin = __shfl_sync(0xffffffffu, inBin, src);
        k -= before;
        prefix |= unsigned(bin) << shift;
        mask |= 255u << shift;
        cntEq = inBin;
        __syncwarp();
    }
    vk = keyToFloatA(prefix);
    if (k + 1 < cntEq) { vk1 = vk; return; }
    unsigned best = 0xffffffffu;
    for (int i = lane; i < n; i += 32) {
        const float v = z[i];
        if (v != kNoData) {
            const unsigned key = orderedKeyA(v);
            if (key > prefix && key < best) best = key;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    vk1 = best == 0xffffffffu ? vk : keyToFloatA(best);
}

__global__ void __launch_bounds__(256)
spatial_aggregate_kernel(DevGrid raster, const double* __restrict__ px, const double* __restrict__ py,
                         const long long* __restrict__ first, const int* __restrict__ maxNr, int nCells, int method,
                         float* __restrict__ scratch, float* __restrict__ value)
{
    const int lane = threadIdx.x & 31;
    const int cell = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (cell >= nCells) return;
    const long long p0 = first[cell];
    const int n = int(first[cell + 1] - p0);
    PB_DEV_ASSERT(n >= 0 && p0 >= 0);
    float* z = scratch + p0;
    int valid = 0;
#pragma unroll 4
    for (int i = lane; i < n; i += 32) {           // independent iterations: four samples' loads in flight per lane
        const double x = px[p0 + i], y = py[p0 + i];
        const int r = int((y - raster.lly) * raster.invCellSize);
        const int row = (raster.nrRows - 1) - r;
        const int col = int((x - raster.llx) * raster.invCellSize);
        float v = raster.flag;
        if (unsigned(row) < unsigned(raster.nrRows) && unsigned(col) < unsigned(raster.nrCols))
            v = __ldg(raster.data + (size_t)row * raster.nrCols + col);
        const bool ok = !isEqualF(v, raster.flag);
        z[i] = ok ? v : kNoData;          // a fresh Crit3DPoint::z is NODATA (meteoGrid.cpp:703)
        valid += ok;
    }
    for (int o = 16; o; o >>= 1) valid += __shfl_xor_sync(0xffffffffu, valid, o);
    __syncwarp();
    if (method == PB_AGGR_MEDIAN) {
        // sorting::percentile(validValues, 50) (mathFunctions/statistics.cpp): the whole warp selects the two neighbours of the
        // rank instead of sorting the list
        __shared__ unsigned histAll[8][256];
        unsigned* hist = histAll[threadIdx.x >> 5];
        float result = kNoData;
        int m = 0;
        for (int i = lane; i < n; i += 32) m += z[i] != kNoData;
        for (int o = 16; o; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
        if (n > 0 && (double(valid) / double(maxNr[cell])) > 0 && m >= 3) {       // MINIMUM_PERCENTILE_DATA
            const double rank = m * double(50.0 / 100.f) - 1.0;
            float lo, hi;
            if (rank < 0.0) { warpSelectPair(z, n, 0, hist, lo, hi); result = lo; }
            else {
                const int low = int(rank);
                if (low >= m - 1) { warpSelectPair(z, n, m - 1, hist, lo, hi); result = lo; }
                else {
                    warpSelectPair(z, n, low, hist, lo, hi);
                    const double frac = rank - low;
                    result = float(lo + frac * (hi - lo));
                }
            }
        }
        if (lane == 0) value[cell] = result;
        return;
    }
    // average / standard deviation (meteoGrid.cpp:1201-1233; statistics::mean / standardDeviation): float sums over the samples
    // in LIST ORDER. The warp reads 32 samples at a time (every lane the ones it wrote above: coalesced, independent loads)
    // and takes them into the running sum in list order through shuffles; a dropped sample adds +0, which leaves a float
    // sum unchanged. All lanes carry the same sum. (One lane walking the list alone re-read every sample from L2, one
    // dependent load per add: 0.28 ms for the 1600 cells of the hourly batch.)
    float result = kNoData;
    // (validValues / aggregationPointsMaxNr) > (GRID_MIN_COVERAGE / 100), GRID_MIN_COVERAGE = 0 (meteoGrid.h:8)
    if (n > 0 && (double(valid) / double(maxNr[cell])) > 0) {
        float sum = 0.f;
        int m = 0;
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            const float v = i < n ? z[i] : kNoData;
            const bool ok = v != kNoData;
            m += __popc(__ballot_sync(0xffffffffu, ok));
            const float a = ok ? v : 0.f;
#pragma unroll
            for (int k = 0; k < 32; k++) sum += __shfl_sync(0xffffffffu, a, k);
        }
        if (method == PB_AGGR_AVERAGE) {
            if (m > 0) result = sum / float(m);
        } else if (method == PB_AGGR_STDDEV && m > 1) {
            const float mean = sum / float(m);
            float squareDiff = 0.f;
            for (int base = 0; base < n; base += 32) {
                const int i = base + lane;
                const float v = i < n ? z[i] : kNoData;
                const float d = v - mean;
                const float a = v != kNoData ? d * d : 0.f;
#pragma unroll
                for (int k = 0; k < 32; k++) squareDiff += __shfl_sync(0xffffffffu, a, k);
            }
            result = sqrtf(squareDiff / (m - 1));
        }
    }
    if (lane == 0) value[cell] = result;
}

static cudaError_t launchAggregatePair(const DevGrid& raster, const double* px, const double* py, const long long* first,
                                       const int* maxNr, int nCells, int method, float* scratch, float* value, cudaStream_t s)
{
    const unsigned nb = (unsigned)((size_t(nCells) * 32 + 255) / 256);
    spatial_aggregate_kernel<<<nb, 256, 0, s>>>(raster, px, py, first, maxNr, nCells, method, scratch, value);
    return cudaGetLastError();
}

// values of a resident raster at explicit points: Crit3DRasterGrid::getRowCol (gis.cpp:485-492) + gis::isOutOfGridRowCol
// (:771-776) + value[row][col], the lookup of Project::passInterpolatedTemperatureToHumidityPoints (project.cpp:2826-2848)
__global__ void __launch_bounds__(256)
sample_points_kernel(DevGrid g, const double* __restrict__ px, const double* __restrict__ py, int n, float* __restrict__ value,
                     unsigned char* __restrict__ inGrid)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = int((py[i] - g.lly) * g.invCellSize);
    const int row = (g.nrRows - 1) - r;
    const int col = int((px[i] - g.llx) * g.invCellSize);
    const bool in = unsigned(row) < unsigned(g.nrRows) && unsigned(col) < unsigned(g.nrCols);
    value[i] = in ? __ldg(g.data + (size_t)row * g.nrCols + col) : kNoData;
    inGrid[i] = in;
}

__global__ void fill_flag_kernel(float* __restrict__ out, long long n, float v)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = v;
}

// gis::updateMinMaxRasterGrid (gis.cpp:569-614) as a separate pass over a finished grid
__global__ void __launch_bounds__(256) raster_minmax_ops_kernel(const float* __restrict__ v, long long n, float flag, unsigned* __restrict__ minmax)
{
    float x = flag;
    bool any = false;
    float vmin = INFINITY, vmax = -INFINITY;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        x = v[k];
        if (!isEqualF(x, flag) && !isEqualF(x, kNoData)) { vmin = fminf(vmin, x); vmax = fmaxf(vmax, x); any = true; }
    }
    (void)any;
    for (int o = 16; o; o >>= 1) {
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (vmin != INFINITY) atomicMin(&minmax[0], orderedKeyR(vmin));
        if (vmax != -INFINITY) atomicMax(&minmax[1], orderedKeyR(vmax));
    }
}

// ---- topographicIndex (project/interpolationCmd.cpp:397-482): the orography-index proxy. Per window width: over the disc of
// cellNr cells around a cell — its own row and column excluded (:438) — the weights 1 - delta / cellNr of the lower, higher and
// equal (|dz| <= EPSILON) neighbours; index = (lower - higher - equal / 2) / total; the widths add up. One thread per cell walks
// its window in the reference's row-major order with FLOAT accumulators (same sums, same bits); the weights only depend on
// the offset and come from a shared-memory table built once per CTA (IEEE sqrtf, float division like the reference).
__global__ void __launch_bounds__(256)
topographic_index_kernel(DevGrid dem, int k /* cellNr */, int useTable, float* __restrict__ out)
{
    extern __shared__ float wTable[];          // (k + 1) x (k + 1): weight of offset (|dr|, |dc|), < 0 outside the disc
    const float cellNr = float(k);
    if (useTable) {
        for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < (k + 1) * (k + 1); i += blockDim.x * blockDim.y) {
            const float dx = float(i / (k + 1)), dy = float(i % (k + 1));
            const float cellDelta = sqrtf(dx * dx + dy * dy);
            wTable[i] = cellDelta <= cellNr ? 1 - (cellDelta / cellNr) : -1.f;
        }
        __syncthreads();
    }
    const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y * blockDim.y + threadIdx.y;
    if (row >= dem.nrRows || col >= dem.nrCols) return;
    const float z = __ldg(dem.data + (size_t)row * dem.nrCols + col);
    if (isEqualF(z, dem.flag)) return;
    const float threshold = float(kEpsilon);
    float higherSum = 0, lowerSum = 0, equalSum = 0, weightSum = 0;
    const int r1 = max(row - k, 0), r2 = min(row + k, dem.nrRows - 1), c1 = max(col - k, 0), c2 = min(col + k, dem.nrCols - 1);
    for (int wr = r1; wr <= r2; wr++) {
        if (wr == row) continue;
        const float* __restrict__ line = dem.data + (size_t)wr * dem.nrCols;
        const int adr = abs(wr - row);
        for (int wc = c1; wc <= c2; wc++) {
            if (wc == col) continue;
            float weight;
            if (useTable) {
                weight = wTable[adr * (k + 1) + abs(wc - col)];
                if (weight < 0.f) continue;
            } else {
                const float dx = float(row - wr), dy = float(col - wc);
                const float cellDelta = sqrtf(dx * dx + dy * dy);
                if (!(cellDelta <= cellNr)) continue;
                weight = 1 - (cellDelta / cellNr);
            }
            const float value = __ldg(line + wc);
            if (isEqualF(value, dem.flag)) continue;
            if (value - z > threshold) higherSum += weight;
            else if (value - z < -threshold) lowerSum += weight;
            else equalSum += weight;
            weightSum += weight;
        }
    }
    if (weightSum > 0) {
        float* o = out + (size_t)row * dem.nrCols + col;
        const double v = (double(lowerSum - higherSum) - double(equalSum) * 0.5) / double(weightSum);
        *o = isEqualF(*o, dem.flag) ? float(v) : float(double(*o) + v);
    }
}

}  // namespace pb

using namespace pb;

namespace {

int fail(pb_context* ctx, int code, const std::string& msg) { return pb::failCtx(ctx, code, msg); }

RasterEntry* entryOf(pb_context* ctx, pb_raster_id id)
{
    if (id < 0 || id >= (int)ctx->rasters.size() || !ctx->rasters[id].used) return nullptr;
    return &ctx->rasters[id];
}

DevGrid gridOf(const RasterEntry& r)
{
    DevGrid g{};
    g.data = r.d;
    g.nrRows = r.h.nrRows; g.nrCols = r.h.nrCols;
    g.llx = r.h.llx; g.lly = r.h.lly; g.cellSize = r.h.cellSize; g.invCellSize = r.h.invCellSize;
    g.flag = r.h.flag;
    return g;
}

float keyToFloatR(unsigned k)
{
    unsigned u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// D2H of a device grid into a pb_raster_out (contiguous data or row pointers) + min / max / nValid bookkeeping
int finishGrid(pb_context* ctx, const float* dGrid, int rows, int cols, bool wantMinMax, pb_raster_out* out)
{
    if (out && (out->data || out->rows)) {
        if (out->data)
            CUDA_TRY(ctx, cudaMemcpyAsync(out->data, dGrid, sizeof(float) * size_t(rows) * cols, cudaMemcpyDeviceToHost, ctx->stream));
        else
            for (int r = 0; r < rows; r++)
                CUDA_TRY(ctx, cudaMemcpyAsync(out->rows[r], dGrid + size_t(r) * cols, sizeof(float) * cols, cudaMemcpyDeviceToHost,
                                              ctx->stream));
        ctx->timing.d2h_bytes += (int64_t)(sizeof(float) * size_t(rows) * cols);
    }
    if (wantMinMax) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hMinMax, ctx->dMinMax, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    cudaEventRecord(ctx->ev[3], ctx->stream);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->timing.kernel_ms, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->timing.total_ms, ctx->ev[1], ctx->ev[3]);
    if (out && wantMinMax) {
        if (ctx->hMinMax[0] == 0xffffffffu) {
            out->minimum = kNoData;
            out->maximum = kNoData;
            return fail(ctx, PB_ERR_NO_VALID_CELL, "no valid cell in the output raster");
        }
        out->minimum = keyToFloatR(ctx->hMinMax[0]);
        out->maximum = keyToFloatR(ctx->hMinMax[1]);
    }
    return PB_OK;
}

int resetMinMax(pb_context* ctx)
{
    ctx->hMinMax[0] = 0xffffffffu;
    ctx->hMinMax[1] = 0u;
    ctx->hMinMax[2] = 0u;
    ctx->hMinMax[3] = 0u;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dMinMax, ctx->hMinMax, 4 * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
    return PB_OK;
}

// a new resident raster (device buffer from the pool) with the given header; returns its id
int newResident(pb_context* ctx, pb_raster_header h /* by value: the caller's may live in the vector that grows here */, pb_raster_id* id)
{
    const size_t n = size_t(h.nrRows) * h.nrCols;
    if (h.nrRows <= 0 || h.nrCols <= 0) return fail(ctx, PB_ERR_INVALID, "raster has no cells");
    if (n > size_t(0x7fffffff)) return fail(ctx, PB_ERR_UNSUPPORTED, "raster larger than 2^31-1 cells");
    int slot = -1;
    for (size_t i = 0; i < ctx->rasters.size(); i++) if (!ctx->rasters[i].used) { slot = (int)i; break; }
    if (slot < 0) { ctx->rasters.emplace_back(); slot = (int)ctx->rasters.size() - 1; }
    void* d = nullptr;
    CUDA_TRY(ctx, ctx->pool.acquire(n * sizeof(float), &d));
    RasterEntry& e = ctx->rasters[slot];
    e = RasterEntry();
    e.d = static_cast<float*>(d);
    e.h = h;
    e.n = n;
    e.used = true;
    if (ctx->outInitRaster == slot) ctx->outInitRaster = -1;
    *id = slot;
    return PB_OK;
}

}  // namespace

namespace pb {
// device-pointer forms of three map operators for the batch drivers (station_space.cu: pb_meteo_grid_step_batch), which chain
// them on a lane's stream without a host round trip; scratch / value: the caller's (one float per aggregation point / cell)
cudaError_t launchSpatialAggregateDev(const DevGrid& raster, const AggregationGeometry& g, int method, float* scratch, float* value,
                                      cudaStream_t s)
{
    return launchAggregatePair(raster, g.x, g.y, g.first, g.maxNr, g.nCells, method, scratch, value, s);
}
// grid of the grid-stride map kernels: 8 CTAs of 256 threads per SM (a full SM), never more than the map needs
static unsigned mapGridBlocks(long long n)
{
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return (unsigned)std::max<long long>(1, std::min<long long>((n + 1023) / 1024, (long long)sms * 8));
}
cudaError_t launchRelativeHumidityMapDev(const float* airT, float flagT, const float* dewT, float flagTd, long long n, float outFlag,
                                         float* out, unsigned* minmax, cudaStream_t s)
{
    relative_humidity_map_kernel<<<mapGridBlocks(n), 256, 0, s>>>(airT, flagT, dewT, flagTd, n, outFlag, out, minmax);
    return cudaGetLastError();
}
cudaError_t launchSamplePointsDev(const DevGrid& g, const double* x, const double* y, int n, float* value, unsigned char* inGrid,
                                  cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    sample_points_kernel<<<(n + 255) / 256, 256, 0, s>>>(g, x, y, n, value, inGrid);
    return cudaGetLastError();
}
}  // namespace pb

extern "C" {

int pb_raster_download(pb_context* ctx, pb_raster_id id, pb_raster_out* out)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    RasterEntry* e = entryOf(ctx, id);
    if (!e || !out) return fail(ctx, PB_ERR_INVALID, "bad raster id / null output");
    cudaSetDevice(ctx->device);
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    cudaEventRecord(ctx->ev[1], ctx->stream);
    cudaEventRecord(ctx->ev[2], ctx->stream);
    return finishGrid(ctx, e->d, e->h.nrRows, e->h.nrCols, false, out);
}

int pb_relative_humidity_map(pb_context* ctx, pb_raster_id airT, pb_raster_id dewT, pb_raster_out* out, pb_raster_id* outId)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    RasterEntry *t = entryOf(ctx, airT), *td = entryOf(ctx, dewT);
    if (!t || !td) return fail(ctx, PB_ERR_INVALID, "bad raster id");
    if (t->h.nrRows != td->h.nrRows || t->h.nrCols != td->h.nrCols)
        return fail(ctx, PB_ERR_INVALID, "temperature and dew point maps differ in size");
    cudaSetDevice(ctx->device);
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    const long long n = (long long)t->n;
    float* dOut = nullptr;
    pb_raster_id nid = -1;
    if (outId) {
        int rc = newResident(ctx, t->h, &nid);
        if (rc) return rc;
        t = entryOf(ctx, airT);            // the raster vector may have grown
        td = entryOf(ctx, dewT);
        dOut = ctx->rasters[nid].d;
        *outId = nid;
    } else {
        CUDA_TRY(ctx, ctx->dOut.reserve(size_t(n)));
        ctx->outInitRaster = -1;
        dOut = ctx->dOut.p;
    }
    int rc = resetMinMax(ctx);
    if (rc) return rc;
    cudaEventRecord(ctx->ev[1], ctx->stream);
    // emptyGrid(): every cell = the output's own flag; the output header is the air temperature map's
    CUDA_TRY(ctx, pb::launchRelativeHumidityMapDev(t->d, t->h.flag, td->d, td->h.flag, (long long)n, t->h.flag, dOut, ctx->dMinMax,
                                                   ctx->stream));
    ctx->timing.launches++;
    cudaEventRecord(ctx->ev[2], ctx->stream);
    if (out) out->nValid = n;
    return finishGrid(ctx, dOut, t->h.nrRows, t->h.nrCols, true, out);
}

int pb_fix_daily_thermal_consistency(pb_context* ctx, pb_raster_id tmax, pb_raster_id tmin, int64_t* nFixed)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    RasterEntry *a = entryOf(ctx, tmax), *b = entryOf(ctx, tmin);
    if (!a || !b) return fail(ctx, PB_ERR_INVALID, "bad raster id");
    if (a->h.nrRows != b->h.nrRows || a->h.nrCols != b->h.nrCols) return fail(ctx, PB_ERR_INVALID, "maps differ in size");
    cudaSetDevice(ctx->device);
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    if (!ctx->dTotal) {
        CUDA_TRY(ctx, cudaMalloc(&ctx->dTotal, sizeof(long long)));
        CUDA_TRY(ctx, cudaMallocHost(&ctx->hTotal, sizeof(long long)));
    }
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->dTotal, 0, sizeof(long long), ctx->stream));
    const long long n = (long long)a->n;
    cudaEventRecord(ctx->ev[1], ctx->stream);
    thermal_consistency_kernel<<<(unsigned)((n + 1023) / 1024), 256, 0, ctx->stream>>>(a->d, a->h.flag, b->d, b->h.flag, n,
                                                                                   reinterpret_cast<unsigned long long*>(ctx->dTotal));
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.launches++;
    cudaEventRecord(ctx->ev[2], ctx->stream);
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hTotal, ctx->dTotal, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->timing.kernel_ms, ctx->ev[1], ctx->ev[2]);
    if (nFixed) *nFixed = *ctx->hTotal;
    b->nValid = -1;        // the list of valid cells does not change (flags are untouched), values do
    return PB_OK;
}

int pb_topographic_distance_map(pb_context* ctx, pb_raster_id demId, double x, double y, double z, pb_raster_out* out,
                                pb_raster_id* outId)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    RasterEntry* dem = entryOf(ctx, demId);
    if (!dem) return fail(ctx, PB_ERR_INVALID, "dem id is not a live raster");
    cudaSetDevice(ctx->device);
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    const long long n = (long long)dem->n;
    float* dOut = nullptr;
    if (outId) {
        pb_raster_id nid = -1;
        int rc = newResident(ctx, dem->h, &nid);
        if (rc) return rc;
        dem = entryOf(ctx, demId);
        dOut = ctx->rasters[nid].d;
        *outId = nid;
    } else {
        CUDA_TRY(ctx, ctx->dOut.reserve(size_t(n)));
        ctx->outInitRaster = -1;
        dOut = ctx->dOut.p;
    }
    cudaEventRecord(ctx->ev[1], ctx->stream);
    topographic_distance_map_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(gridOf(*dem), float(x), float(y), float(z), dOut);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.launches++;
    cudaEventRecord(ctx->ev[2], ctx->stream);
    if (out) out->nValid = n;
    return finishGrid(ctx, dOut, dem->h.nrRows, dem->h.nrCols, false, out);
}

int pb_resample_grid(pb_context* ctx, pb_raster_id srcId, const pb_raster_header* newHeader, int method, float nodataRatioThreshold,
                     pb_raster_out* out, pb_raster_id* outId)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    RasterEntry* src = entryOf(ctx, srcId);
    if (!src || !newHeader) return fail(ctx, PB_ERR_INVALID, "bad raster id / null header");
    if (method != PB_AGGR_AVERAGE && method != PB_AGGR_MEDIAN && method != PB_AGGR_CENTER && method != PB_AGGR_PREVAILING)
        return fail(ctx, PB_ERR_UNSUPPORTED, "resampleGrid: aggrAverage, aggrMedian, aggrPrevailing and aggrCenter run on the GPU");
    const double factor = newHeader->cellSize / src->h.cellSize;
    if (factor > 1. && (method == PB_AGGR_MEDIAN || method == PB_AGGR_PREVAILING) && int(floor(factor)) + 1 > 16)
        return fail(ctx, PB_ERR_UNSUPPORTED, "median / prevailing resampling above factor 15 (more than 256 samples per cell)");
    cudaSetDevice(ctx->device);
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    pb_raster_header h = *newHeader;
    const long long n = (long long)h.nrRows * h.nrCols;
    if (n <= 0) return fail(ctx, PB_ERR_INVALID, "new grid has no cells");
    float* dOut = nullptr;
    if (outId) {
        pb_raster_id nid = -1;
        int rc = newResident(ctx, h, &nid);
        if (rc) return rc;
        src = entryOf(ctx, srcId);
        dOut = ctx->rasters[nid].d;
        *outId = nid;
    } else {
        CUDA_TRY(ctx, ctx->dOut.reserve(size_t(n)));
        ctx->outInitRaster = -1;
        dOut = ctx->dOut.p;
    }
    int rc = resetMinMax(ctx);
    if (rc) return rc;
    DevGrid ng{};
    ng.data = nullptr;
    ng.nrRows = h.nrRows; ng.nrCols = h.nrCols;
    ng.llx = h.llx; ng.lly = h.lly; ng.cellSize = h.cellSize; ng.invCellSize = h.invCellSize;
    ng.flag = h.flag;
    cudaEventRecord(ctx->ev[1], ctx->stream);
    const unsigned nb = (unsigned)((n + 127) / 128);
    if (method == PB_AGGR_AVERAGE)
        resample_grid_kernel<PB_AGGR_AVERAGE><<<nb, 128, 0, ctx->stream>>>(gridOf(*src), ng, nodataRatioThreshold, dOut, ctx->dMinMax);
    else if (method == PB_AGGR_MEDIAN)
        resample_grid_kernel<PB_AGGR_MEDIAN><<<nb, 128, 0, ctx->stream>>>(gridOf(*src), ng, nodataRatioThreshold, dOut, ctx->dMinMax);
    else if (method == PB_AGGR_PREVAILING)
        resample_grid_kernel<PB_AGGR_PREVAILING><<<nb, 128, 0, ctx->stream>>>(gridOf(*src), ng, nodataRatioThreshold, dOut, ctx->dMinMax);
    else
        resample_grid_kernel<PB_AGGR_CENTER><<<nb, 128, 0, ctx->stream>>>(gridOf(*src), ng, nodataRatioThreshold, dOut, ctx->dMinMax);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.launches++;
    cudaEventRecord(ctx->ev[2], ctx->stream);
    if (out) out->nValid = n;
    rc = finishGrid(ctx, dOut, h.nrRows, h.nrCols, true, out);
    return rc == PB_ERR_NO_VALID_CELL ? PB_OK : rc;      // resampleGrid ignores updateMinMaxRasterGrid's result
}

int pb_aggregation_upload(pb_context* ctx, const double* x, const double* y, const int64_t* first, const int32_t* maxNr, int nCells,
                          pb_aggregation_id* id)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (!x || !y || !first || !maxNr || nCells <= 0 || !id) return fail(ctx, PB_ERR_INVALID, "null argument");
    if (first[0] != 0) return fail(ctx, PB_ERR_INVALID, "first[0] must be 0");
    for (int c = 0; c < nCells; c++) if (first[c + 1] < first[c]) return fail(ctx, PB_ERR_INVALID, "first[] must not decrease");
    cudaSetDevice(ctx->device);
    const long long nPoints = first[nCells];
    int slot = -1;
    for (size_t i = 0; i < ctx->aggregations.size(); i++) if (!ctx->aggregations[i].used) { slot = int(i); break; }
    if (slot < 0) { ctx->aggregations.emplace_back(); slot = int(ctx->aggregations.size()) - 1; }
    AggregationGeometry g;
    g.nCells = nCells;
    g.nPoints = nPoints;
    const size_t np = size_t(std::max<long long>(nPoints, 1));
    CUDA_TRY(ctx, cudaMalloc(&g.x, np * 8));
    CUDA_TRY(ctx, cudaMalloc(&g.y, np * 8));
    CUDA_TRY(ctx, cudaMalloc(&g.first, size_t(nCells + 1) * 8));
    CUDA_TRY(ctx, cudaMalloc(&g.maxNr, size_t(nCells) * 4));
    CUDA_TRY(ctx, cudaMalloc(&g.scratch, np * 4));
    CUDA_TRY(ctx, cudaMalloc(&g.value, size_t(nCells) * 4));
    CUDA_TRY(ctx, cudaMemcpyAsync(g.x, x, size_t(nPoints) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(g.y, y, size_t(nPoints) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(g.first, first, size_t(nCells + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(g.maxNr, maxNr, size_t(nCells) * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    g.used = true;
    ctx->aggregations[slot] = g;
    *id = slot;
    return PB_OK;
}

int pb_aggregation_free(pb_context* ctx, pb_aggregation_id id)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (id < 0 || id >= (int)ctx->aggregations.size() || !ctx->aggregations[id].used) return fail(ctx, PB_ERR_INVALID, "bad aggregation id");
    cudaSetDevice(ctx->device);
    AggregationGeometry& g = ctx->aggregations[id];
    cudaFree(g.x); cudaFree(g.y); cudaFree(g.first); cudaFree(g.maxNr); cudaFree(g.scratch); cudaFree(g.value);
    g = AggregationGeometry();
    return PB_OK;
}

int pb_spatial_aggregate_meteo_grid(pb_context* ctx, pb_raster_id rasterId, pb_aggregation_id aggId, int method, float* value)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    RasterEntry* src = entryOf(ctx, rasterId);
    if (!src || !value) return fail(ctx, PB_ERR_INVALID, "bad raster id / null output");
    if (aggId < 0 || aggId >= (int)ctx->aggregations.size() || !ctx->aggregations[aggId].used)
        return fail(ctx, PB_ERR_INVALID, "bad aggregation id");
    if (method != PB_AGGR_AVERAGE && method != PB_AGGR_MEDIAN && method != PB_AGGR_STDDEV) {
        // spatialAggregateMeteoGridPoint returns NODATA for every other method (meteoGrid.cpp:1234-1237)
        for (int c = 0; c < ctx->aggregations[aggId].nCells; c++) value[c] = kNoData;
        return PB_OK;
    }
    cudaSetDevice(ctx->device);
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    AggregationGeometry& g = ctx->aggregations[aggId];
    cudaEventRecord(ctx->ev[1], ctx->stream);
    CUDA_TRY(ctx, launchAggregatePair(gridOf(*src), g.x, g.y, g.first, g.maxNr, g.nCells, method, g.scratch, g.value, ctx->stream));
    ctx->timing.launches++;
    cudaEventRecord(ctx->ev[2], ctx->stream);
    CUDA_TRY(ctx, cudaMemcpyAsync(value, g.value, size_t(g.nCells) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    cudaEventRecord(ctx->ev[3], ctx->stream);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->timing.d2h_bytes = (int64_t)(size_t(g.nCells) * 4);
    cudaEventElapsedTime(&ctx->timing.kernel_ms, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->timing.total_ms, ctx->ev[1], ctx->ev[3]);
    return PB_OK;
}

int pb_raster_sample_points(pb_context* ctx, pb_raster_id rasterId, const double* x, const double* y, int n, float* value,
                            uint8_t* inGrid)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    RasterEntry* src = entryOf(ctx, rasterId);
    if (!src || !x || !y || !value || n < 0) return fail(ctx, PB_ERR_INVALID, "bad raster id / null argument");
    if (n == 0) return PB_OK;
    cudaSetDevice(ctx->device);
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    const size_t nAl = (size_t(n) + 1) / 2 * 2;
    CUDA_TRY(ctx, ctx->dScratch.reserve(nAl * 8 * 2 + nAl * 4 + nAl + 64));
    unsigned char* d = ctx->dScratch.p;
    double* dX = reinterpret_cast<double*>(d);
    double* dY = dX + nAl;
    float* dV = reinterpret_cast<float*>(dY + nAl);
    unsigned char* dIn = reinterpret_cast<unsigned char*>(dV + nAl);
    CUDA_TRY(ctx, cudaMemcpyAsync(dX, x, size_t(n) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(dY, y, size_t(n) * 8, cudaMemcpyHostToDevice, ctx->stream));
    sample_points_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(gridOf(*src), dX, dY, n, dV, dIn);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.launches++;
    CUDA_TRY(ctx, cudaMemcpyAsync(value, dV, size_t(n) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<uint8_t> tmp;
    if (!inGrid) { tmp.resize(n); inGrid = tmp.data(); }
    CUDA_TRY(ctx, cudaMemcpyAsync(inGrid, dIn, size_t(n), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->timing.h2d_bytes = int64_t(n) * 16;
    ctx->timing.d2h_bytes = int64_t(n) * 5;
    return PB_OK;
}

int pb_topographic_index(pb_context* ctx, pb_raster_id demId, const float* windowWidths, int nWidths, pb_raster_out* out,
                         pb_raster_id* outId)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    RasterEntry* src = entryOf(ctx, demId);
    if (!src || !windowWidths) return fail(ctx, PB_ERR_INVALID, "bad raster id / null widths");
    if (nWidths <= 0) return fail(ctx, PB_ERR_INVALID, "no window width (interpolationCmd.cpp:403)");
    cudaSetDevice(ctx->device);
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    pb_raster_header h = src->h;
    const long long n = (long long)h.nrRows * h.nrCols;
    float* dOut = nullptr;
    if (outId) {
        pb_raster_id nid = -1;
        int rc = newResident(ctx, h, &nid);
        if (rc) return rc;
        src = entryOf(ctx, demId);
        dOut = ctx->rasters[nid].d;
        *outId = nid;
    } else {
        CUDA_TRY(ctx, ctx->dOut.reserve(size_t(n)));
        ctx->outInitRaster = -1;
        dOut = ctx->dOut.p;
    }
    cudaEventRecord(ctx->ev[1], ctx->stream);
    {   // initializeGrid(DEM): every cell = flag
        const unsigned nb = (unsigned)std::min<long long>((n + 255) / 256, 148 * 32);
        fill_flag_kernel<<<nb, 256, 0, ctx->stream>>>(dOut, n, h.flag);
        CUDA_TRY(ctx, cudaGetLastError());
        ctx->timing.launches++;
    }
    const dim3 block(32, 8), grid((h.nrCols + 31) / 32, (h.nrRows + 7) / 8);
    for (int w = 0; w < nWidths; w++) {
        const float cellNr = float(round(windowWidths[w] / h.cellSize));        // :413
        if (!(cellNr >= 0.f) || cellNr > 1.0e6f) return fail(ctx, PB_ERR_INVALID, "window width out of range");
        const int k = int(cellNr);
        const size_t table = size_t(k + 1) * (k + 1) * sizeof(float);
        const int useTable = table <= 200 * 1024;
        if (useTable && table > 48 * 1024)
            CUDA_TRY(ctx, cudaFuncSetAttribute(topographic_index_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)table));
        topographic_index_kernel<<<grid, block, useTable ? table : 0, ctx->stream>>>(gridOf(*src), k, useTable, dOut);
        CUDA_TRY(ctx, cudaGetLastError());
        ctx->timing.launches++;
    }
    int rc = resetMinMax(ctx);
    if (rc) return rc;
    raster_minmax_ops_kernel<<<148 * 8, 256, 0, ctx->stream>>>(dOut, n, h.flag, ctx->dMinMax);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.launches++;
    cudaEventRecord(ctx->ev[2], ctx->stream);
    if (out) out->nValid = n;
    rc = finishGrid(ctx, dOut, h.nrRows, h.nrCols, true, out);
    return rc == PB_ERR_NO_VALID_CELL ? PB_OK : rc;      // topographicIndex ignores updateMinMaxRasterGrid's result
}

}  // extern "C"
