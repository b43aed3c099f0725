// pre_kernels.cu — K3: preInterpolation for a BATCH of time steps (agrolib/interpolation/interpolation.cpp:2444-2499),
// multiple-detrending branch: multipleDetrendingMain :1832-1859 on all stations of each step.
//
// The station-space fits are latency-bound scalar work (a few ms per step on a CPU core); one time step alone cannot fill
// a GPU. The batch drivers of the reference (interpolationMeteoGridPeriod, pragaProject.cpp:3048; 365 days x 24 h x
// variables) make thousands of INDEPENDENT steps, so the mapping is: one group of 16/32 lanes per time step, lane g =
// first guess g of the elevation fit (multiple_detrend.cuh), thousands of steps in flight. Same arithmetic and operation
// order as the reference (-fmad=false), so the detrended values and fitted parameters are bit-identical.
#include "multiple_detrend.cuh"
#include "pre.cuh"

namespace pb {

// Units: time steps of one station set (sub.idx == nullptr: unit t = all stations with values[t * n + k]) or, for glocal
// detrending, MACRO AREAS of one time step (unit t = the stations sub.idx[sub.off[t] .. sub.off[t + 1]) with the shared
// values[k], one model for all units, the unweighted elevation fit of glocalDetrendingFitting :2239-2294).
template <int G, int F>
__global__ void __launch_bounds__(kLocalThreads, 2)
pre_interpolation_batch_kernel(const LocalModel* __restrict__ models, LocalStations st, const float* __restrict__ values, int T,
                               LocalScratch scr, PreSubsets sub, float* __restrict__ detrended, unsigned char* __restrict__ keep,
                               PreFit* __restrict__ fits)
{
    constexpr int groupsPerCta = kLocalThreads / G;
    constexpr int N = Fit<F>::N;
    const int grp = threadIdx.x / G, lane = grpLane<G>();
    const long long groupsTotal = (long long)gridDim.x * groupsPerCta;
    const long long gid = (long long)blockIdx.x * groupsPerCta + grp;
    const WorkPtrs w = carve(scr.base + gid * scr.perGroup, scr.cap);
    for (long long t = gid; t < T; t += groupsTotal) {
        const LocalModel& m = models[sub.idx ? 0 : t];
        const int first = sub.idx ? sub.off[t] : 0;
        const int n = sub.idx ? sub.off[t + 1] - first : st.n;
        const float* val = sub.idx ? values : values + (size_t)t * st.n;
        for (int k = lane; k < n; k += G) {
            const int i = sub.idx ? sub.idx[first + k] : k;
            w.idx[k] = i;
            w.w[k] = st.weight ? st.weight[i] : kNoData;
            w.val[k] = val[i];
            w.flg[k] = 1;
        }
        grpSync<G>();
        MdFit<N> fit;
        multipleDetrendingGroup<G, F>(m, st, w, n, false, fit, sub.unweightedElevation != 0);
        grpSync<G>();
        if (detrended)
            for (int k = lane; k < n; k += G) {
                detrended[(size_t)t * st.n + k] = w.val[k];
                keep[(size_t)t * st.n + k] = w.flg[k] & 1;
            }
        if (lane == 0) {
            PreFit& o = fits[t];
            o.elevSig = fit.elevSig;
            o.elevR2 = fit.elevR2;
            o.nPred = fit.nPred;
            for (int j = 0; j < PB_MAX_FIT_PARAMS; j++) o.elevPar[j] = j < N ? fit.elevPar[j] : 0.;
            for (int i = 0; i < PB_MAX_PROXY; i++) {
                o.sigPos[i] = i < fit.nPred ? fit.sigPos[i] : -1;
                o.othPar[i][0] = i < fit.nPred ? fit.othPar[i][0] : 0.;
                o.othPar[i][1] = i < fit.nPred ? fit.othPar[i][1] : 0.;
            }
        }
        grpSync<G>();
    }
}

int preGroupsTotal(int groupLanes, int T)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int perCta = kLocalThreads / groupLanes;
    const int ctas = std::min(sms * 2, (T + perCta - 1) / perCta);
    return ctas * perCta;
}

size_t preScratchBytesPerGroup(int cap) { return workBytes(cap); }

template <int G, int F>
static cudaError_t launchPreT(const LocalModel* models, const LocalStations& st, const float* values, int T,
                              const LocalScratch& scr, const PreSubsets& sub, float* detrended, unsigned char* keep, PreFit* fits,
                              cudaStream_t s)
{
    const int perCta = kLocalThreads / G;
    const int grid = preGroupsTotal(G, T) / perCta;
    pre_interpolation_batch_kernel<G, F><<<grid, kLocalThreads, 0, s>>>(models, st, values, T, scr, sub, detrended, keep, fits);
    return cudaGetLastError();
}

cudaError_t launchPreInterpolationBatch(const LocalModel* models, const LocalStations& st, const float* values, int T,
                                        const LocalScratch& scr, int groupLanes, int elevFunction, float* detrended,
                                        unsigned char* keep, PreFit* fits, cudaStream_t s, const PreSubsets* subsets)
{
    if (T <= 0) return cudaSuccess;
    PreSubsets sub{};
    if (subsets) sub = *subsets;
#define PB_PRE_CASE(G_, F_) return launchPreT<G_, F_>(models, st, values, T, scr, sub, detrended, keep, fits, s)
    if (groupLanes == 16) {
        switch (elevFunction) {
        case PB_FIT_PIECEWISE_THREE: PB_PRE_CASE(16, PB_FIT_PIECEWISE_THREE);
        case PB_FIT_PIECEWISE_THREE_FREE: PB_PRE_CASE(16, PB_FIT_PIECEWISE_THREE_FREE);
        default: PB_PRE_CASE(16, PB_FIT_PIECEWISE_TWO);
        }
    }
    switch (elevFunction) {
    case PB_FIT_PIECEWISE_THREE: PB_PRE_CASE(32, PB_FIT_PIECEWISE_THREE);
    case PB_FIT_PIECEWISE_THREE_FREE: PB_PRE_CASE(32, PB_FIT_PIECEWISE_THREE_FREE);
    default: PB_PRE_CASE(32, PB_FIT_PIECEWISE_TWO);
    }
#undef PB_PRE_CASE
}

}  // namespace pb
