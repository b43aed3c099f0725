// shepard.cuh — descriptors of the Shepard estimators (K8), shared by engine.cu and shepard_kernels.cu
#pragma once
#include "common.cuh"

namespace pb {

// the stations computeDistances leaves with a non-zero distance, in their original order: f32 coordinates for the
// distances (gis::computeDistance, gis.cpp:685-691), f64 coordinates for the direction cosines (point->utm), raw values
struct ShepardStations {
    const float* x;
    const float* y;
    const double* xd;
    const double* yd;
    const float* value;
    const float* z;                // float(point->z): topographic distance only
    int n;
};

// topographic distance of computeDistances (interpolation.cpp:226-251) in front of the Shepard estimators: kh != 0 and the DEM
// the ray march samples; tz = the targets' elevation in the point form (raster form: the DEM value of the cell)
struct ShepardTd {
    int kh;                        // 0: planar distances only
    DevGrid dem;
    const float* tz;
    const DevGrid* maps;           // the stations' precomputed maps (td_device.cuh), nullptr = none
};

cudaError_t launchShepardRaster(const ShepardStations& st, const DevModel& m, const DevGrid& dem, float R0, bool modified,
                                const int* validIdx, int nTargets, float* out, unsigned* minmax, cudaStream_t s,
                                const ShepardTd* td = nullptr);
cudaError_t launchShepardPoints(const ShepardStations& st, const DevModel& m, float R0, bool modified, const float* tx,
                                const float* ty, const double* tproxy, int nTargets, float* out, cudaStream_t s,
                                const ShepardTd* td = nullptr);

}  // namespace pb
