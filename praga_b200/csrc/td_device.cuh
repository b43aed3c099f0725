// td_device.cuh — gis::topographicDistance (agrolib/gis/gis.cpp:1595-1647) on the device, shared by the IDW (td_kernels.cu) and
// Shepard (shepard_kernels.cu) estimators: computeDistances adds kh * topographicDistance to the planar distance before ANY
// estimator runs (interpolation.cpp:208-256, 2513-2530). Include only from TUs built with -fmad=false.
#pragma once
#include "common.cuh"

namespace pb {

__device__ __forceinline__ float topographicDistanceDev(const DevGrid& dem, float x1, float y1, float z1, float x2, float y2,
                                                        float z2, float distance)
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
        // Crit3DRasterGrid::getValueFromXY(double, double), gis.cpp:533-546
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

// computeDistances (interpolation.cpp:226-251) in front of the march: the point's precomputed map
// (Crit3DInterpolationDataPoint::topographicDistance, written by gis::topographicDistanceMap gis.cpp:1650-1675 and loaded
// by Project::loadTopographicDistanceMaps project.cpp:2368-2430) where the target lies inside it; NODATA there (or no map)
// => topographicDistance() on the current DEM. maps: one descriptor per staged station, data == nullptr = no map.
__device__ __forceinline__ float topographicDistanceCachedDev(const DevGrid* __restrict__ maps, int i, const DevGrid& dem, float x1,
                                                              float y1, float z1, float x2, float y2, float z2, float distance)
{
    if (maps) {
        const DevGrid& g = maps[i];
        if (g.data) {
            const double x = double(x1), y = double(y1);
            // isOutOfGridXY (gis.cpp:840-845), getRowColFromXY (:735-739)
            const bool out = (x < g.llx) || (y < g.lly) || (x >= (g.llx + g.nrCols * g.cellSize)) || (y >= (g.lly + g.nrRows * g.cellSize));
            if (!out) {
                const int row = (g.nrRows - 1) - int(floor((y - g.lly) * g.invCellSize));
                const int col = int(floor((x - g.llx) * g.invCellSize));
                if (unsigned(row) < unsigned(g.nrRows) && unsigned(col) < unsigned(g.nrCols)) {
                    const float topo = __ldg(g.data + (size_t)row * g.nrCols + col);
                    if (!isEqualF(topo, kNoData)) return topo;
                }
            }
        }
    }
    return topographicDistanceDev(dem, x1, y1, z1, x2, y2, z2, distance);
}

}  // namespace pb
