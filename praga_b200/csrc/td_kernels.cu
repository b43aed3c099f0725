// td_kernels.cu — K6: IDW with topographic distance (computeDistances with getUseTD(), interpolation.cpp:226-251;
// gis::topographicDistance, gis/gis.cpp:1595-1647). For every (target, station) pair the reference marches
// nrStep = int(d / cellSize) equal steps from the lower to the higher endpoint through the DEM and adds
// kh * max(DEM - z_low) to the planar distance. The march is a chain of f32 additions whose rounding decides which
// DEM cell each step samples, so it is reproduced operation by operation (this TU is built with -fmad=false) and the
// weights are summed in f64 in station order like inverseDistanceWeighted (:1031-1051): the result is bit-identical.
// One thread per target; the DEM stays in L2 (the march of neighbouring targets touches neighbouring cells).
// Per-station precomputed TD rasters (Crit3DInterpolationDataPoint::topographicDistance, pb_set_topographic_distance_maps)
// are read in front of the march like computeDistances does (td_device.cuh).
#include "epilogue.cuh"
#include "td_device.cuh"

namespace pb {

__device__ __forceinline__ unsigned orderedKeyTd(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

template <bool POINTS>
__global__ void __launch_bounds__(128)
idw_td_kernel(DevStations st, const float* __restrict__ sz, DevModel m, DevGrid dem, int kh, const int* __restrict__ validIdx,
              int nTargets, const float* __restrict__ tx, const float* __restrict__ ty, const float* __restrict__ tz,
              const double* __restrict__ tproxy, float* __restrict__ out, unsigned* __restrict__ minmax,
              const DevGrid* __restrict__ tdMaps)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    float vmin = INFINITY, vmax = -INFINITY;
    if (t < nTargets) {
        float x, y, z;
        int cell;
        if (POINTS) { cell = t; x = tx[t]; y = ty[t]; z = tz[t]; }
        else {
            cell = validIdx[t];
            const int row = cell / dem.nrCols, col = cell - row * dem.nrCols;
            // getUtmXYFromRowColSinglePrecision(header, ...), gis.cpp:799-804
            x = float(dem.llx + dem.cellSize * double(col) + 0.5);
            y = float(dem.lly + dem.cellSize * double(dem.nrRows - row) - 0.5);
            z = dem.data[cell];
        }
        double sum = 0, sumWeights = 0;
        for (int i = 0; i < st.n; i++) {
            const float px = st.x[i], py = st.y[i];
            const float dx = px - x, dy = py - y;
            float d = sqrtf(dx * dx + dy * dy);
            const float topo = topographicDistanceCachedDev(tdMaps, i, dem, x, y, z, px, py, sz[i], d);
            d += (kh * topo);
            if (d > 0.f) {
                const double dk = double(d) / 10000.;
                const double w = 1.0 / (dk * dk * dk);
                sumWeights += w;
                sum += double(st.val[i]) * w;
            }
        }
        double pv[PB_MAX_PROXY];
        if (POINTS) {
            for (int i = 0; i < m.nProxy; i++) pv[i] = tproxy[(size_t)cell * m.nProxy + i];
        } else {
            for (int i = 0; i < m.nProxy; i++) {
                pv[i] = double(kNoData);
                const DevProxy& p = m.proxy[i];
                if (m.useDetrendingVar && p.active && p.gridSlot >= 0) {
                    const DevGrid& g = m.grid[p.gridSlot];
                    const float v = gridValueFromXY(g, double(x), double(y));
                    if (v != g.flag) pv[i] = double(v);
                }
            }
        }
        float result;
        if (m.useRetrendOnly) result = 0.f;
        else result = (sumWeights > 0.0) ? float(sum / sumWeights + m.valueOffset) : kNoData;
        const float o = finishEstimate(m, result, pv);
        out[cell] = o;
        if (!isEqualF(o, dem.flag) && !isEqualF(o, kNoData)) { vmin = o; vmax = o; }
    }
    if (minmax) {
        for (int o = 16; o; o >>= 1) {
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        }
        if ((threadIdx.x & 31) == 0) {
            if (vmin != INFINITY) atomicMin(&minmax[0], orderedKeyTd(vmin));
            if (vmax != -INFINITY) atomicMax(&minmax[1], orderedKeyTd(vmax));
        }
    }
}

cudaError_t launchIdwTdRaster(const DevStations& st, const float* sz, const DevModel& m, const DevGrid& dem, int kh,
                              const int* validIdx, int nTargets, float* out, unsigned* minmax, cudaStream_t s,
                              const DevGrid* tdMaps)
{
    if (nTargets <= 0) return cudaSuccess;
    idw_td_kernel<false><<<(nTargets + 127) / 128, 128, 0, s>>>(st, sz, m, dem, kh, validIdx, nTargets, nullptr, nullptr, nullptr,
                                                               nullptr, out, minmax, tdMaps);
    return cudaGetLastError();
}

cudaError_t launchIdwTdPoints(const DevStations& st, const float* sz, const DevModel& m, const DevGrid& dem, int kh,
                              const float* tx, const float* ty, const float* tz, const double* tproxy, int nTargets, float* out,
                              cudaStream_t s, const DevGrid* tdMaps)
{
    if (nTargets <= 0) return cudaSuccess;
    idw_td_kernel<true><<<(nTargets + 127) / 128, 128, 0, s>>>(st, sz, m, dem, kh, nullptr, nTargets, tx, ty, tz, tproxy, out,
                                                              nullptr, tdMaps);
    return cudaGetLastError();
}

}  // namespace pb
