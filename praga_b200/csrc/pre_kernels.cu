// pre_kernels.cu — K3: preInterpolation for a BATCH of time steps (agrolib/interpolation/interpolation.cpp:2444-2499),
// multiple-detrending branch: multipleDetrendingMain :1832-1859 on all stations of each step.
//
// The station-space fits are latency-bound scalar work (a few ms per step on a CPU core); one time step alone cannot fill
// a GPU. The batch drivers of the reference (interpolationMeteoGridPeriod, pragaProject.cpp:3048; 365 days x 24 h x
// variables) make thousands of INDEPENDENT steps, so the mapping is: one group of 16/32 lanes per time step, lane g =
// first guess g of the elevation fit (multiple_detrend.cuh), thousands of steps in flight. Same arithmetic and operation
// order as the reference (-fmad=false), so the detrended values and fitted parameters are bit-identical.
#include <stdlib.h>
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

// ---------------------------------------------------------------------------------------------------------
// Few units (one time step of the drop-in preInterpolation, the spatial quality control's two fits, a handful of macro
// areas): the GPU is otherwise idle and the LATENCY of the longest descent is the cost. One CTA per unit, one WARP per first
// guess: lmElevationWarp spreads the data points of each iteration over the lanes and keeps every sum in the reference's
// order (detrend_device.cuh), so parameters and detrended values stay bit-identical to the lane-per-descent form above.
// Warp 0 stages the points, applies the pick rule of bestFittingMarquardt_nDimension and runs the rest of
// multipleDetrendingMain (mdFinish) as a 32-lane group.
// ---------------------------------------------------------------------------------------------------------
constexpr int kCoopWarps = 16;
constexpr int kCoopMaxUnits = 64;

template <int N>
struct CoopResult {
    double q[N];
    double R2, RMSE;
    int inRange, secondBelow;
};

template <int F>
__global__ void __launch_bounds__(kCoopWarps * 32, 1)
pre_interpolation_coop_kernel(const LocalModel* __restrict__ models, LocalStations st, const float* __restrict__ values, int T,
                              LocalScratch scr, PreSubsets sub, float* __restrict__ detrended, unsigned char* __restrict__ keep,
                              PreFit* __restrict__ fits)
{
    constexpr int N = Fit<F>::N;
    extern __shared__ __align__(16) unsigned char coopSmem[];
    double* smWarp = reinterpret_cast<double*>(coopSmem) + (size_t)(threadIdx.x >> 5) * (kLmWarpTerms * kLmWarpStride);
    CoopResult<N>* results = reinterpret_cast<CoopResult<N>*>(coopSmem + sizeof(double) * kCoopWarps * kLmWarpTerms * kLmWarpStride);
    __shared__ int sNE, sValid;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x;
    if (t >= T) return;
    const WorkPtrs w = carve(scr.base + (size_t)t * scr.perGroup, scr.cap);
    const LocalModel& m = models[sub.idx ? 0 : t];
    const int first = sub.idx ? sub.off[t] : 0;
    const int n = sub.idx ? sub.off[t + 1] - first : st.n;
    const float* val = sub.idx ? values : values + (size_t)t * st.n;
    const int ePos = m.elevationPos;
    const bool wantElev = ePos >= 0 && m.selectedActive[ePos];
    if (warp == 0) {
        for (int k = lane; k < n; k += 32) {
            const int i = sub.idx ? sub.idx[first + k] : k;
            w.idx[k] = i;
            w.w[k] = st.weight ? st.weight[i] : kNoData;
            w.val[k] = val[i];
            w.flg[k] = 1;
        }
        __syncwarp();
        bool isValid = false;
        int nE = 0;
        if (wantElev) {
            nE = mdStageElevation<32>(m, st, w, n, false, isValid);
            if (isValid && sub.unweightedElevation) {
                for (int j = lane; j < nE; j += 32) w.W[j] = 1.;
                __syncwarp();
            }
        }
        if (lane == 0) { sNE = nE; sValid = isValid ? 1 : 0; }
    }
    __syncthreads();
    const int nE = sNE;
    const bool fitRan = sValid != 0;
    const bool unweighted = sub.unweightedElevation != 0;
    if (fitRan) {
        double maxZ = 0;
        if (unweighted && nE > 0) {
            maxZ = w.X[0];
            for (int i = 0; i < nE; i++) if (w.X[i] > maxZ) maxZ = w.X[i];
        }
        for (int k = warp; k < m.nFirstGuess; k += kCoopWarps) {
            double q[N], R2 = 0, RMSE = 0;
            bool inRange = true, secondBelow = true;
#pragma unroll
            for (int j = 0; j < N; j++) q[j] = m.firstGuess[k][j];
            lmElevationWarp<F>(m.elevMin, m.elevMax, m.elevDelta, q, 1000, 0.002, w.X, w.Y, w.W, nE, smWarp);
            if (unweighted) {
#pragma unroll
                for (int j = 0; j < N; j++) inRange = inRange && !(q[j] < m.elevMin[j] || q[j] > m.elevMax[j]);
                plainScores<F>(q, w.X, w.Y, nE, R2, RMSE);
                if (N > 4) secondBelow = (q[0] + q[2]) < maxZ;
            } else {
                weightedScores<F>(q, w.X, w.Y, w.W, nE, R2, RMSE);
            }
            if (lane == 0) {
                CoopResult<N>& r = results[k];
#pragma unroll
                for (int j = 0; j < N; j++) r.q[j] = q[j];
                r.R2 = R2; r.RMSE = RMSE; r.inRange = inRange; r.secondBelow = secondBelow;
            }
        }
    }
    __syncthreads();
    if (warp != 0) return;
    // the pick rule of bestFittingMarquardt_nDimension (:1360-1427 weighted, :1433-1529 unweighted), guesses in order
    double elevPar[N], R2best = 0;
#pragma unroll
    for (int j = 0; j < N; j++) elevPar[j] = 0;
    if (fitRan) {
        double bestR2 = -9999, bestRMSE = -9999;
        int rmseIndex = -9999;
        for (int k = 0; k < m.nFirstGuess; k++) {
            const CoopResult<N>& r = results[k];
            if (!r.inRange) continue;
            if (isEqualD(bestRMSE, -9999) || r.RMSE < bestRMSE) { bestRMSE = r.RMSE; rmseIndex = k; }
            if (isEqualD(bestR2, -9999) || (r.R2 > (bestR2 + 0.005) && r.secondBelow)) {
#pragma unroll
                for (int j = 0; j < N; j++) elevPar[j] = r.q[j];
                bestR2 = r.R2;
            }
        }
        if (bestR2 < 0 && rmseIndex >= 0) {
#pragma unroll
            for (int j = 0; j < N; j++) elevPar[j] = m.firstGuess[rmseIndex][j];
            lmElevationWarp<F>(m.elevMin, m.elevMax, m.elevDelta, elevPar, 1000, 0.002, w.X, w.Y, w.W, nE, smWarp);
        }
        R2best = bestR2;
    }
    MdFit<N> fit;
    mdFinish<32, F>(m, st, w, n, false, fitRan, R2best, elevPar, fit, smWarp);
    __syncwarp();
    if (detrended)
        for (int k = lane; k < n; k += 32) {
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
}

template <int F>
static cudaError_t launchCoopT(const LocalModel* models, const LocalStations& st, const float* values, int T, const LocalScratch& scr,
                               const PreSubsets& sub, float* detrended, unsigned char* keep, PreFit* fits, cudaStream_t s)
{
    constexpr int N = Fit<F>::N;
    const size_t smem = sizeof(double) * kCoopWarps * kLmWarpTerms * kLmWarpStride + sizeof(CoopResult<N>) * PB_MAX_FIRST_GUESS;
    auto k = pre_interpolation_coop_kernel<F>;
    // set on every launch: the attribute is per device and a host may drive several devices from one process
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<T, kCoopWarps * 32, smem, s>>>(models, st, values, T, scr, sub, detrended, keep, fits);
    return cudaGetLastError();
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
    static const bool noCoop = getenv("PB_PRE_NO_COOP") != nullptr;
    if (T <= kCoopMaxUnits && !noCoop) {
        switch (elevFunction) {
        case PB_FIT_PIECEWISE_THREE: return launchCoopT<PB_FIT_PIECEWISE_THREE>(models, st, values, T, scr, sub, detrended, keep, fits, s);
        case PB_FIT_PIECEWISE_THREE_FREE:
            return launchCoopT<PB_FIT_PIECEWISE_THREE_FREE>(models, st, values, T, scr, sub, detrended, keep, fits, s);
        default: return launchCoopT<PB_FIT_PIECEWISE_TWO>(models, st, values, T, scr, sub, detrended, keep, fits, s);
        }
    }
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
