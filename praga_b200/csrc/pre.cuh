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

int preGroupsTotal(int groupLanes, int T);
size_t preScratchBytesPerGroup(int cap);
cudaError_t launchPreInterpolationBatch(const LocalModel* models, const LocalStations& st, const float* values, int T,
                                        const LocalScratch& scr, int groupLanes, int elevFunction, float* detrended,
                                        unsigned char* keep, PreFit* fits, cudaStream_t s);

}  // namespace pb
