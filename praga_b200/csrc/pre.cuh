// pre.cuh — batched preInterpolation (K3): descriptors shared by engine.cu and pre_kernels.cu
#pragma once
#include "local.cuh"

namespace pb {

// what multipleDetrendingMain leaves behind for one time step
struct PreFit {
    int elevSig;
    float elevR2;
    int nPred;
    int sigPos[PB_MAX_PROXY];
    double elevPar[PB_MAX_FIT_PARAMS];
    double othPar[PB_MAX_PROXY][2];
};

// station subsets of the units (glocal detrending: one unit per macro area); idx == nullptr: every unit = all stations
struct PreSubsets {
    const int* idx;                // concatenated station indices, in each area's order
    const int* off;                // nUnits + 1 offsets into idx
    int unweightedElevation;       // glocalDetrendingFitting: bestFittingMarquardt_nDimension without weights
};

int preGroupsTotal(int groupLanes, int T);
size_t preScratchBytesPerGroup(int cap);
cudaError_t launchPreInterpolationBatch(const LocalModel* models, const LocalStations& st, const float* values, int T,
                                        const LocalScratch& scr, int groupLanes, int elevFunction, float* detrended,
                                        unsigned char* keep, PreFit* fits, cudaStream_t s, const PreSubsets* subsets = nullptr);

}  // namespace pb
