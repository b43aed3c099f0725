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

}  // namespace pb
