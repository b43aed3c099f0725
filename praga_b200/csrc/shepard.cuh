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
    int n;
};

cudaError_t launchShepardRaster(const ShepardStations& st, const DevModel& m, const DevGrid& dem, float R0, bool modified,
                                const int* validIdx, int nTargets, float* out, unsigned* minmax, cudaStream_t s);
cudaError_t launchShepardPoints(const ShepardStations& st, const DevModel& m, float R0, bool modified, const float* tx,
                                const float* ty, const double* tproxy, int nTargets, float* out, cudaStream_t s);

}  // namespace pb
