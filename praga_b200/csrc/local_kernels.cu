// local_kernels.cu — K2: local-detrending raster. One GROUP of lanes (half a warp for the 16-first-guess two-piece
// lapse rate, a full warp for the three-piece ones) owns one DEM cell and does, for that cell, what the loop of
// Project::interpolationDemLocalDetrending does (agrolib/project/project.cpp:3220-3250):
//
//   gis::getUtmXYFromRowCol                   gis/gis.cpp:806-810   (true cell centre in f64, narrowed to f32 at the calls)
//   getProxyValuesXY                          interpolation/interpolation.cpp:2620-2647
//   localSelection                            :1087-1170  ring search, selection order = ring-major then station index
//   preInterpolation -> multipleDetrendingMain :2444 / :1832 (weighted LM fits, detrend_device.cuh)
//   interpolate (idw) + retrend + clamp       :2502-2560, :1031-1051, :1288-1351
//   clearFitting / setCurrentCombination      (state is per cell here, nothing to reset)
//
// Built with -fmad=false: the FP64 fits follow the reference's operation order without FMA contraction, so the
// Levenberg-Marquardt descents take the same accept/reject path as on the CPU. Distances use IEEE f32 ops.
#include <stdlib.h>
#include "local_device.cuh"

namespace pb {

// one cell of the loop of Project::interpolationDemLocalDetrending: selection, multipleDetrendingMain (:1832-1859), estimate
template <int G, int F>
__device__ float localCell(const LocalModel& m, const LocalStations& st, const float2* __restrict__ xy, float xf, float yf,
                           const double* pv, const WorkPtrs& gw, int gcap, unsigned char* tile, int tileCap, int* errFlag,
                           bool excludeSupplemental, double* fitResults)
{
    WorkPtrs w;
    float localRadius = 0.f;
    const int n = localSelect<G>(m, st, xy, xf, yf, gw, gcap, tile, tileCap, errFlag, w, localRadius);
    MdFit<Fit<F>::N> fit;
    multipleDetrendingGroup<G, F>(m, st, w, n, true, fit, false, fitResults);
    return localFinish<G, F>(m, st, w, n, pv, fit, excludeSupplemental, xf, yf, localRadius);
}

template <int G, int F, int MINB>
__global__ void __launch_bounds__(kLocalThreads, MINB)
local_detrending_kernel(const LocalModel* __restrict__ mp, LocalStations st, DevGrid dem, const int* __restrict__ validIdx,
                        int nTargetsHost, LocalScratch scr, int xyInSmem, float* __restrict__ out, unsigned* __restrict__ minmax,
                        LocalPointTargets pt, const int* __restrict__ nTargetsDev)
{
    // list form behind the pipeline (local_pipeline.cu): the number of cells left to this kernel is only known on the device
    const int nTargets = nTargetsDev ? *nTargetsDev : nTargetsHost;
    if (nTargets <= 0) return;
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int groupsPerCta = kLocalThreads / G;
    const LocalModel& m = *mp;
    const int grp = threadIdx.x / G;
    const int lane = grpLane<G>();
    constexpr int tileCap = localTileCap(G);
    const size_t tileBytes = workBytes(tileCap);
    unsigned char* tile = smem + grp * tileBytes;
    float2* sxy = reinterpret_cast<float2*>(smem + groupsPerCta * tileBytes);
    if (xyInSmem) {
        for (int i = threadIdx.x; i < st.n; i += kLocalThreads) sxy[i] = st.xy[i];
        __syncthreads();
    }
    const float2* xy = xyInSmem ? sxy : st.xy;
    const long long groupsTotal = (long long)gridDim.x * groupsPerCta;
    const long long gid = (long long)blockIdx.x * groupsPerCta + grp;
    const WorkPtrs gw = carve(scr.base + gid * scr.perGroup, scr.cap);
    // per-guess results of the elevation fit when a group has fewer lanes than first guesses (behind the group's work arrays)
    double* fitResults = reinterpret_cast<double*>(scr.base + gid * scr.perGroup + workBytes(scr.cap));

    float vmin = INFINITY, vmax = -INFINITY;
    for (long long t = gid; t < nTargets; t += groupsTotal) {
        if (pt.x) {       // point form: the target's own coordinates and proxy values, result t
            double pv[PB_MAX_PROXY];
            for (int i = 0; i < m.nProxy; i++) pv[i] = pt.pv[t * m.nProxy + i];
            const float o = localCell<G, F>(m, st, xy, pt.x[t], pt.y[t], pv, gw, scr.cap, tile, tileCap, scr.error, pt.excludeSupplemental != 0, fitResults);
            if (lane == 0) out[t] = o;
            continue;
        }
        const int cell = validIdx[t];
        const int row = cell / dem.nrCols, col = cell - row * dem.nrCols;
        // gis::getUtmXYFromRowCol (gis.cpp:806-810), then narrowed to float by the callee signatures (interpolation.h:77-95,162)
        const double x = dem.llx + dem.cellSize * (col + 0.5);
        const double y = dem.lly + dem.cellSize * (dem.nrRows - row - 0.5);
        const float xf = float(x), yf = float(y);
        double pv[PB_MAX_PROXY];
        for (int i = 0; i < m.nProxy; i++) {
            pv[i] = double(kNoData);
            if (m.selectedActive[i] && m.gridSlot[i] >= 0) {
                const DevGrid& g = m.grid[m.gridSlot[i]];
                const float v = gridValueXY(g, double(xf), double(yf));
                if (v != g.flag) pv[i] = double(v);
            }
        }
        const float o = localCell<G, F>(m, st, xy, xf, yf, pv, gw, scr.cap, tile, tileCap, scr.error, true, fitResults);
        if (lane == 0) {
            out[cell] = o;
            if (!isEqualF(o, dem.flag) && !isEqualF(o, kNoData)) {
                vmin = fminf(vmin, o);
                vmax = fmaxf(vmax, o);
            }
        }
    }
    if (lane == 0 && minmax) {
        if (vmin != INFINITY) atomicMin(&minmax[0], orderedKeyL(vmin));
        if (vmax != -INFINITY) atomicMax(&minmax[1], orderedKeyL(vmax));
    }
}

size_t localScratchBytesPerGroup(int cap) { return workBytes(cap) + sizeof(double) * kFitResultDoubles; }

// persistent grid: every SM filled to the occupancy the fit kernel's registers allow (at most 2 CTAs per SM)
static int localCtas()
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms * 4;
}
int localGroupsTotal(int groupLanes) { return localCtas() * (kLocalThreads / groupLanes); }
// Lanes per cell. The descents of a cell end after 2-60 iterations: with one lane per first guess the group waits for its
// longest descent (12 of 32 lanes busy on average); with fewer lanes each lane walks several guesses back to back and the
// totals even out, the group-uniform parts (validity checks, the other proxies' fit, the estimate) are replicated on fewer
// lanes, and a warp holds more cells. Measured on B200 at config-3 density (profiles/r01_local_groups.json).
int localGroupLanes(int nFirstGuess)
{
    static const int forced = getenv("PB_LOCAL_G") ? atoi(getenv("PB_LOCAL_G")) : 0;
    if (forced == 4 || forced == 8 || forced == 16) return forced;
    return nFirstGuess <= 16 ? 4 : 8;       // 4 lanes x 4 guesses (two-piece), 8 lanes x 4 guesses (three-piece)
}

template <int G, int F, int MINB>
static cudaError_t launchT(const LocalModel* dModel, const LocalStations& st, const DevGrid& dem, const int* validIdx,
                           int nTargets, const LocalScratch& scratch, float* out, unsigned* minmax, cudaStream_t s,
                           const LocalPointTargets& pt, const int* nTargetsDev)
{
    constexpr int groupsPerCta = kLocalThreads / G;
    size_t smem = groupsPerCta * workBytes(localTileCap(G));
    int xyInSmem = 0;
    if (smem + size_t(st.n) * sizeof(float2) <= (G >= 16 ? 100 : 200) * 1024) {
        xyInSmem = 1;
        smem += size_t(st.n) * sizeof(float2);
    }
    auto k = local_detrending_kernel<G, F, MINB>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int perSm = 1, dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k, kLocalThreads, smem);
    if (e != cudaSuccess) return e;
    perSm = perSm < 1 ? 1 : (perSm > 4 ? 4 : perSm);
    const long long need = (nTargets + groupsPerCta - 1) / groupsPerCta;
    const int grid = (int)(need < (long long)sms * perSm ? need : (long long)sms * perSm);
    k<<<grid, kLocalThreads, smem, s>>>(dModel, st, dem, validIdx, nTargets, scratch, xyInSmem, out, minmax, pt, nTargetsDev);
    return cudaGetLastError();
}

cudaError_t launchLocalDetrendingRaster(const LocalModel* dModel, const LocalStations& st, const DevGrid& dem,
                                        const int* validIdx, int nTargets, const LocalScratch& scratch, int groupLanes,
                                        int elevFunction, float* out, unsigned* minmax, cudaStream_t s,
                                        const LocalPointTargets* points, const int* nTargetsDev)
{
    if (nTargets <= 0) return cudaSuccess;
    const LocalPointTargets pt = points ? *points : LocalPointTargets{nullptr, nullptr, nullptr, 1};
    // 1 resident CTA per SM (255 registers, no spills in the Levenberg-Marquardt loop) measured faster than 2 CTAs at 128
    // registers (155 vs 164 ms on 925k cells / 5000 stations): the loop is latency bound and spills sit on its critical path
#define PB_LOCAL_CASE(G_, F_) return launchT<G_, F_, 1>(dModel, st, dem, validIdx, nTargets, scratch, out, minmax, s, pt, nTargetsDev)
    if (groupLanes == 4) {
        switch (elevFunction) {
        case PB_FIT_PIECEWISE_THREE: PB_LOCAL_CASE(4, PB_FIT_PIECEWISE_THREE);
        case PB_FIT_PIECEWISE_THREE_FREE: PB_LOCAL_CASE(4, PB_FIT_PIECEWISE_THREE_FREE);
        default: PB_LOCAL_CASE(4, PB_FIT_PIECEWISE_TWO);
        }
    }
    if (groupLanes == 8) {
        switch (elevFunction) {
        case PB_FIT_PIECEWISE_THREE: PB_LOCAL_CASE(8, PB_FIT_PIECEWISE_THREE);
        case PB_FIT_PIECEWISE_THREE_FREE: PB_LOCAL_CASE(8, PB_FIT_PIECEWISE_THREE_FREE);
        default: PB_LOCAL_CASE(8, PB_FIT_PIECEWISE_TWO);
        }
    }
    switch (elevFunction) {     // 16 lanes: one lane per first guess of the two-piece function (the form measured before)
    case PB_FIT_PIECEWISE_THREE: PB_LOCAL_CASE(16, PB_FIT_PIECEWISE_THREE);
    case PB_FIT_PIECEWISE_THREE_FREE: PB_LOCAL_CASE(16, PB_FIT_PIECEWISE_THREE_FREE);
    default: PB_LOCAL_CASE(16, PB_FIT_PIECEWISE_TWO);
    }
#undef PB_LOCAL_CASE
}

}  // namespace pb
