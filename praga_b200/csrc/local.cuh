// local.cuh — descriptors of the local-detrending raster kernel (K2), shared by engine.cu and local_kernels.cu
#pragma once
#include "common.cuh"

namespace pb {

// everything multipleDetrendingMain + localSelection + interpolate + retrend read from Crit3DInterpolationSettings
// (interpolation/interpolationSettings.h:185-391), flattened once per call on the host. Lives in global memory.
struct LocalModel {
    int nProxy;
    int elevationPos;              // last proxy whose name is elevation (multipleDetrendingMain :1838-1842), -1 = none
    int useLapseRateCode;
    int minPointsLocal;            // getMinPointsLocalDetrending()
    unsigned minPoints;            // unsigned(minPointsLocal * 1.2f)   (localSelection :1092-1093)
    unsigned thr80;                // unsigned(minPoints * 0.8)         (:1152)
    float candRadius;              // radius of the per-cell candidate list of the ring search (0 = scan all stations)
    float shepardRadius;           // computeShepardInitialRadius(bboxArea, S, minPoints) when S <= minPoints (:1098)
    float bboxArea;                // getPointsBoundingBoxArea(): shepardSearchNeighbour over the subset (:812)
    int clampMode;
    float rainfallThreshold;
    int useDoNotRetrend, useRetrendOnly;
    int method;                    // PB_METHOD_IDW / _SHEPARD / _SHEPARD_MODIFIED (interpolate :2513-2530; modified = the demo's default)
    uint8_t selectedActive[PB_MAX_PROXY];
    int kind[PB_MAX_PROXY];
    float stdDevThreshold[PB_MAX_PROXY];
    int gridSlot[PB_MAX_PROXY];
    // elevation fit (setFittingParameters_elevation :1625-1677 after setMultipleDetrendingHeightTemperatureRange :1506-1553)
    int elevFunction, elevNPar, nFirstGuess;
    double elevMin[PB_MAX_FIT_PARAMS], elevMax[PB_MAX_FIT_PARAMS], elevDelta[PB_MAX_FIT_PARAMS];
    double firstGuess[PB_MAX_FIRST_GUESS][PB_MAX_FIT_PARAMS];
    // other proxies (setFittingParameters_otherProxies :1680-1728): linear, 2 parameters each
    double otherMin[PB_MAX_PROXY][2], otherMax[PB_MAX_PROXY][2], otherDelta[PB_MAX_PROXY][2], otherStart[PB_MAX_PROXY][2];
    DevGrid grid[PB_MAX_PROXY];
};

// all stations in index order (the selection order of localSelection depends on it)
struct LocalStations {
    const float2* xy;              // float(point->utm.x), float(point->utm.y)
    const float* value;
    const float* weight;           // regressionWeight as passed in (only used when S <= minPoints)
    const float* proxy;            // n * nProxy
    const int* code;               // lapseRateCode
    const uint8_t* active;
    const double* xd;              // point->utm.x / y in f64: the direction term of modifiedShepardIdw (:994-1003)
    const double* yd;
    int n;
};

// per-group scratch in global memory (capacity `cap` entries per group), used for the selection list and, when a
// selection is larger than the shared-memory tile, for the working arrays
struct LocalScratch {
    unsigned char* base;
    size_t perGroup;               // bytes
    int cap;
    int* error;                    // != 0: a selection exceeded cap
};

// point form (computeResidualsLocalDetrending, spatialControl.cpp:158-228): explicit targets instead of DEM cells
struct LocalPointTargets {
    const float* x;                // float(point.utm.x); nullptr = raster form
    const float* y;
    const double* pv;              // n * nProxy: the targets' own proxy values (getProxyValues())
    int excludeSupplemental;       // the flag interpolate() receives (raster loop: true; leave-one-out: false)
};

size_t localScratchBytesPerGroup(int cap);
int localGroupsTotal(int groupLanes);
int localGroupLanes(int nFirstGuess);      // lanes per cell chosen for a model (PB_LOCAL_G overrides)
cudaError_t launchLocalDetrendingRaster(const LocalModel* dModel, const LocalStations& st, const DevGrid& dem,
                                        const int* validIdx, int nTargets, const LocalScratch& scratch, int groupLanes,
                                        int elevFunction, float* out, unsigned* minmax, cudaStream_t s,
                                        const LocalPointTargets* points = nullptr, const int* nTargetsDev = nullptr);

// the raster form as a three-kernel pipeline over one chunk of cells (local_pipeline.cu): selection -> flat
// Levenberg-Marquardt descents -> pick + estimate; cells that do not fit a record end up in overflowList (count in counters[1])
size_t localRecordStride(int nGuess);
size_t localPipelineWorkInts(int nCells);
cudaError_t launchLocalPipelineChunk(const LocalModel* dModel, const LocalStations& st, const DevGrid& dem, const int* validIdx,
                                     int nCells, int elevFunction, float candRadius, unsigned char* records, size_t recStride,
                                     int* work, int* overflowList, int* overflowCount, float* out, unsigned* minmax, cudaStream_t s,
                                     int* launches);

}  // namespace pb
