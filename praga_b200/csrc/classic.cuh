// classic.cuh — descriptors of the station-space kernels (K7): classic detrending, exact leave-one-out, CV error tables,
// station-to-station topographic distances. Shared by engine.cu and classic_kernels.cu.
#pragma once
#include "common.cuh"

namespace pb {

// Crit3DProxy regression fields (interpolationSettings.h:18-102) as one problem leaves them; `written` has one bit per
// field the reference code path actually assigned (optimalDetrending walks its combinations on ONE settings object, so
// fields a later combination does not touch keep the earlier values: the host replays the writes in the reference order)
enum ClassicField {
    CF_SLOPE = 1, CF_INTERCEPT = 2, CF_R2 = 4, CF_H0 = 8, CF_H1 = 16, CF_T0 = 32, CF_T1 = 64, CF_INVLAPSE = 128, CF_INVSIG = 256
};
struct ClassicProxyState {
    float slope, intercept, r2, H0, H1, T0, T1, invLapse;
    int invSig;
    unsigned written;
};

// one detrending() call (interpolation.cpp:1380-1415): a time step's values under one proxy combination
struct ClassicProblem {
    uint8_t active[PB_MAX_PROXY];          // inCombination.isProxyActive
    uint8_t significant[PB_MAX_PROXY];     // out: setSignificantCurrentCombination
    int useThermalInversion;               // inCombination.getUseThermalInversion
    int error;                             // out: != 0 => more than kMaxIntervals height intervals
    ClassicProxyState proxy[PB_MAX_PROXY]; // in: state before the call; out: state after it
};

struct ClassicSetup {
    int n, nProxy, nProblems;
    int isThermalVar;                      // isThermal(myVar), interpolation.cpp:1190
    int isElaborationVar;
    int useLapseRateCode;
    int indexHeight;                       // getIndexHeight(): last proxy named elevation, -1 = none
    float minRegressionR2, maxHeightInversion, climateLapseRate;
    int kind[PB_MAX_PROXY];
};

// station constants shared by every problem
struct ClassicStations {
    const double* z;                       // point->z
    const float* proxy;                    // n * nProxy
    const int* code;
    const uint8_t* active;                 // isActive (NULL = all)
};

// exact leave-one-out (computeResiduals, spatialControl.cpp:97-155): everything interpolate() reads besides the points.
// Sources = the interpolation points (nS), targets = the meteo points (nT >= nS: multiple detrending may have erased
// interpolation points, the meteo points stay).
struct LooSetup {
    int nS, nT, nProxy, nProblems, nKh;
    int useLapseRateCode, excludeSupplemental, excludeOutsideDem;
    int useDetrendingVar, useThermalInversion, useDoNotRetrend, useRetrendOnly;
    int clampMode, isPrecVar;
    float rainfallThreshold;
    int kind[PB_MAX_PROXY];
    int useMultipleModel;                  // retrend with the fitted multiple-detrending model instead of a problem's proxies
    DevModel multiple;
};
struct LooTargets {
    const float* x;
    const float* y;
    const float* proxy;                    // nT * nProxy
    const int* code;
    const uint8_t* insideDem;
    const float* observed;                 // Crit3DMeteoPoint::currentValue
};

constexpr int kLooProblemsPerThread = 8;

cudaError_t launchClassicInit(const float* values /* T x n */, int T, int nCombos, int n, float* work /* n x NP */, cudaStream_t s);
cudaError_t launchClassicDetrend(const ClassicSetup& cs, const ClassicStations& st, ClassicProblem* problems, float* work,
                                 cudaStream_t s);
cudaError_t launchClassicGather(const float* work, int n, int NP, int problem, float* out, cudaStream_t s);
cudaError_t launchTopoMatrix(const float* tx, const float* ty, const float* tz, int nT, const float* sx, const float* sy,
                             const float* sz, int nS, const DevGrid& dem, float* topo, cudaStream_t s,
                             const DevGrid* tdMaps = nullptr);
cudaError_t launchLooExact(const LooSetup& ls, const LooTargets& t, const float* sx, const float* sy, const float* work,
                           const ClassicProblem* problems, const float* topo, const int* khList, float* residual, cudaStream_t s,
                           float* estimate = nullptr);

// neighbourhoodVariability (interpolation.cpp:259-301) for nT targets over the nS interpolation points
struct NeighbourOut {
    int ok;                                // nrValidPoints > 1
    float devSt, avgDeltaZ, minDistance;
};
cudaError_t launchNeighbourhood(const float* tx, const float* ty, const float* tz, int nT, const float* sx, const float* sy,
                                const double* sz, const int* scode, const float* svalue, int nS, int useLapseRateCode,
                                const float* topo, int kh, int maxNrPoints, NeighbourOut* out, cudaStream_t s);
cudaError_t launchCvMae(const float* residual, const float* observed, int n, int nJobs, float* mae, cudaStream_t s);

}  // namespace pb
