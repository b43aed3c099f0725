// multiple_detrend.cuh — multipleDetrendingMain (agrolib/interpolation/interpolation.cpp:1832-1859) for ONE point set, run
// by a group of G lanes: elevation fit (first guesses across lanes), elevation detrend, other-proxies fit, other-proxies
// detrend. Shared by the local-detrending raster kernel (K2: point set = the stations selected around a cell) and the
// batched preInterpolation kernel (K3: point set = all stations of a time step). Include only from TUs built with
// -fmad=false.
#pragma once
#include "detrend_device.cuh"
#include "local.cuh"

namespace pb {

constexpr int kLocalThreads = 256;
constexpr int kCap = 64;      // stations per group held in shared memory; larger selections work from global scratch
// smaller groups put more cells in flight per CTA: their shared-memory tiles shrink so that all of them (and the stations'
// coordinates) still fit; config-3 density selects 24-35 stations per cell
__host__ __device__ constexpr int localTileCap(int G) { return G >= 16 ? kCap : (G == 8 ? 56 : 40); }

struct WorkPtrs {
    double *X, *Y, *W;        // LM data (compact)
    int* idx;                 // station index of selected point k
    float *d, *w, *val;       // distance to the cell, regressionWeight, (detrended) value
    int* oi;                  // compact list of point slots used by the other-proxies fit
    unsigned char* flg;       // bit0: still in myPoints, bit1: in othersPoints
    unsigned char* act;       // isActive of compact elevation entry
};

__host__ __device__ constexpr size_t workBytes(int cap) { return (size_t(cap) * (3 * 8 + 5 * 4 + 2) + 15) / 16 * 16; }

__device__ __forceinline__ WorkPtrs carve(unsigned char* base, int cap)
{
    WorkPtrs p;
    p.X = reinterpret_cast<double*>(base);
    p.Y = p.X + cap;
    p.W = p.Y + cap;
    p.idx = reinterpret_cast<int*>(p.W + cap);
    p.d = reinterpret_cast<float*>(p.idx + cap);
    p.w = p.d + cap;
    p.val = p.w + cap;
    p.oi = reinterpret_cast<int*>(p.val + cap);
    p.flg = reinterpret_cast<unsigned char*>(p.oi + cap);
    p.act = p.flg + cap;
    return p;
}

// Crit3DRasterGrid::getValueFromXY (gis.cpp:533-546)
__device__ __forceinline__ float gridValueXY(const DevGrid& g, double x, double y)
{
    const int r = int((y - g.lly) * g.invCellSize);
    const int row = (g.nrRows - 1) - r;
    const int col = int((x - g.llx) * g.invCellSize);
    if (unsigned(row) >= unsigned(g.nrRows) || unsigned(col) >= unsigned(g.nrCols)) return g.flag;
    return __ldg(g.data + (size_t)row * g.nrCols + col);
}

// proxyValidity (interpolation.cpp:1418-1457) over the points whose flag has `bit` set; group-uniform
static __device__ bool proxyValidityFlagged(const LocalStations& st, const WorkPtrs& w, int n, int nProxy, int pos, int bit,
                                     float stdDevThreshold)
{
    double sum = 0;
    unsigned cnt = 0;
    for (int k = 0; k < n; k++) {
        if (!(w.flg[k] & bit)) continue;
        const int i = w.idx[k];
        if (!st.active[i]) continue;
        const float v = st.proxy[(size_t)i * nProxy + pos];
        if (v != kNoData) { sum += v; cnt++; }
    }
    if (cnt < 5) return false;
    const double avg = sum / cnt;
    sum = 0;
    for (int k = 0; k < n; k++) {
        if (!(w.flg[k] & bit)) continue;
        const int i = w.idx[k];
        if (!st.active[i]) continue;
        const float v = st.proxy[(size_t)i * nProxy + pos];
        if (v != kNoData) sum += (v - avg) * (v - avg);
    }
    const double stdDev = sqrt(sum / (cnt - 1));
    if (stdDevThreshold != kNoData) return stdDev > stdDevThreshold;
    return true;
}


// result of multipleDetrendingMain for one point set: what the reference leaves in Crit3DInterpolationSettings
// (currentCombination significance, fittingFunction / fittingParameters)
template <int N>
struct MdFit {
    bool elevSig;
    float elevR2;            // proxy.regressionR2 of the elevation proxy: NODATA until a fit ran, then float(bestR2)
    double elevPar[N];
    int nPred;
    int sigPos[PB_MAX_PROXY];
    double othPar[PB_MAX_PROXY][2];
};

// multipleDetrendingElevationFitting (:1862-1953), first half: the points with a usable lapse-rate code and an elevation
// value go to X (height), Y (value), W (regression weight), act; then proxyValidity (:1418-1457) and the local
// minimum-points rule (:1885). Returns nE; isValid = the elevation function is to be fitted.
template <int G>
__device__ int mdStageElevation(const LocalModel& m, const LocalStations& st, const WorkPtrs& w, int n, bool isLocal, bool& isValid)
{
    const int lane = grpLane<G>();
    const int nProxy = m.nProxy;
    const bool useCode = m.useLapseRateCode != 0;
    const unsigned ltMask = (1u << lane) - 1u;
    const int ePos = m.elevationPos;
    int nE = 0;
    for (int base = 0; base < n; base += G) {
        const int k = base + lane;
        bool ok = false;
        float h = 0.f;
        int i = 0;
        if (k < n) {
            i = w.idx[k];
            h = st.proxy[(size_t)i * nProxy + ePos];
            ok = checkLapseRateCodeDev(st.code[i], useCode, true) && h != kNoData;
        }
        const unsigned b = grpBallot<G>(ok);
        if (ok) {
            const int pos = nE + __popc(b & ltMask);
            w.X[pos] = double(h);
            w.Y[pos] = double(w.val[k]);
            w.W[pos] = (w.w[k] != kNoData) ? double(w.w[k]) : 1.;
            w.act[pos] = st.active[i];
        }
        nE += __popc(b);
    }
    grpSync<G>();
    isValid = false;
    {
        double sum = 0;
        unsigned cnt = 0;
        for (int j = 0; j < nE; j++)
            if (w.act[j]) { sum += w.X[j]; cnt++; }
        if (cnt >= 5) {
            const double avg = sum / cnt;
            sum = 0;
            for (int j = 0; j < nE; j++)
                if (w.act[j]) sum += (w.X[j] - avg) * (w.X[j] - avg);
            const double stdDev = sqrt(sum / (cnt - 1));
            const float thr = m.stdDevThreshold[ePos];
            isValid = (thr != kNoData) ? (stdDev > thr) : true;
        }
    }
    if (isLocal) isValid = isValid && unsigned(nE) >= unsigned(m.minPointsLocal);          // local detrending (:1885)
    return nE;
}

template <int G, int F>
__device__ void mdFinish(const LocalModel& m, const LocalStations& st, const WorkPtrs& w, int n, bool isLocal, bool fitRan,
                         double R2, const double* elevParIn, MdFit<Fit<F>::N>& fit, double* smWarp = nullptr);

// w: idx (station index), w (regressionWeight), val (value, detrended in place), flg (bit0 = still in myPoints) of the
// n points; X, Y, W, oi, act are scratch. isLocal adds the minPointsLocalDetrending conditions (:1885, :2100).
// unweightedElevation: the elevation fit of glocalDetrendingFitting (:2273: isWeighted = false).
template <int G, int F>
__device__ void multipleDetrendingGroup(const LocalModel& m, const LocalStations& st, const WorkPtrs& w, int n, bool isLocal,
                                        MdFit<Fit<F>::N>& fit, bool unweightedElevation = false, double* fitResults = nullptr)
{
    constexpr int N = Fit<F>::N;
    const int ePos = m.elevationPos;
    double elevPar[N];
#pragma unroll
    for (int j = 0; j < N; j++) elevPar[j] = 0;
    bool fitRan = false;
    double R2 = 0;
    if (ePos >= 0 && m.selectedActive[ePos]) {
        bool isValid;
        const int nE = mdStageElevation<G>(m, st, w, n, isLocal, isValid);
        if (isValid) {
            if (unweightedElevation) {
                for (int j = grpLane<G>(); j < nE; j += G) w.W[j] = 1.;
                grpSync<G>();
            }
            R2 = bestFitElevation<G, F>(m.elevMin, m.elevMax, m.elevDelta, &m.firstGuess[0][0], m.nFirstGuess, elevPar, 1000,
                                        0.002, 0.005, w.X, w.Y, w.W, nE, unweightedElevation, fitResults);
            fitRan = true;
        }
    }
    mdFinish<G, F>(m, st, w, n, isLocal, fitRan, R2, elevPar, fit);
}

// the rest of multipleDetrendingMain once the elevation parameters are chosen (fitRan: a fit ran and gave R2, elevParIn):
// significance of the elevation proxy, detrendingElevation, the other proxies' fit and detrendingOtherProxies
// smWarp (G == 32 only): shared memory of lmOthersWarp; with it the other-proxies fit runs in its warp-cooperative form.
template <int G, int F>
__device__ void mdFinish(const LocalModel& m, const LocalStations& st, const WorkPtrs& w, int n, bool isLocal, bool fitRan,
                         double R2, const double* elevParIn, MdFit<Fit<F>::N>& fit, double* smWarp)
{
    constexpr int N = Fit<F>::N;
    const int lane = grpLane<G>();
    const int nProxy = m.nProxy;
    const bool useCode = m.useLapseRateCode != 0;
    const unsigned ltMask = (1u << lane) - 1u;
    const int ePos = m.elevationPos;
    bool elevSig = false;
    double elevPar[N];
#pragma unroll
    for (int j = 0; j < N; j++) elevPar[j] = elevParIn[j];
    fit.elevR2 = kNoData;
    if (fitRan) {
        fit.elevR2 = float(R2);
        bool nodataOrZero = false;                                         // isVectorNodataOrZero (:2696-2705)
#pragma unroll
        for (int j = 0; j < N; j++) nodataOrZero |= isEqualD(elevPar[j], -9999) || isEqualD(elevPar[j], 0);
        elevSig = !isEqualD(R2, -9999) || !nodataOrZero;
    }
    grpSync<G>();
    if (elevSig) {
        // detrendingElevation (:1956-1972): every point of myPoints, including those without an elevation value
        double pc[N];
#pragma unroll
        for (int j = 0; j < N; j++) pc[j] = elevPar[j];
        Fit<F>::clamp(pc);
        for (int k = lane; k < n; k += G) {
            const float h = st.proxy[(size_t)w.idx[k] * nProxy + ePos];
            const float detrendValue = float(Fit<F>::eval(double(h), pc));
            w.val[k] -= detrendValue;
        }
        grpSync<G>();
    }

    // multipleDetrendingOtherProxiesFitting (:1975-2171)
    bool othSig[PB_MAX_PROXY];
    int sigPos[PB_MAX_PROXY];
    double othPar[PB_MAX_PROXY][2];
    int nPred = 0;
#pragma unroll
    for (int i = 0; i < PB_MAX_PROXY; i++) othSig[i] = false;
    {
        int nrPredictors = 0;
        for (int pos = 0; pos < nProxy; pos++)
            if (pos != ePos && m.selectedActive[pos]) nrPredictors++;
        if (nrPredictors > 0) {
            for (int k = lane; k < n; k += G)
                w.flg[k] = 1 | (checkLapseRateCodeDev(st.code[w.idx[k]], useCode, false) ? 2 : 0);
            grpSync<G>();
            int validNr = 0;
            for (int pos = 0; pos < nProxy; pos++)
                if (pos != ePos && m.selectedActive[pos]) {
                    othSig[pos] = proxyValidityFlagged(st, w, n, nProxy, pos, 2, m.stdDevThreshold[pos]);
                    if (othSig[pos]) validNr++;
                }
            if (validNr > 0) {
                // points with incomplete (or all-zero) proxies leave myPoints; incomplete ones leave othersPoints (:2040-2093)
                grpSync<G>();
                for (int k = lane; k < n; k += G) {
                    const size_t row = (size_t)w.idx[k] * nProxy;
                    bool isValid = true, atLeastOne = false, othValid = true;
                    for (int pos = 0; pos < nProxy; pos++)
                        if (pos != ePos && m.selectedActive[pos] && othSig[pos]) {
                            const float v = st.proxy[row + pos];
                            if (v == kNoData) othValid = false;
                            if (isValid) {
                                if (isEqualF(v, kNoData)) isValid = false;
                                else if (!isEqualF(v, 0.f)) atLeastOne = true;
                            }
                        }
                    unsigned char f = w.flg[k];
                    if (!(isValid && atLeastOne)) f &= ~1;
                    if (!othValid) f &= ~2;
                    w.flg[k] = f;
                }
                grpSync<G>();
                validNr = 0;
                for (int pos = 0; pos < nProxy; pos++)
                    if (pos != ePos && m.selectedActive[pos]) {
                        othSig[pos] = proxyValidityFlagged(st, w, n, nProxy, pos, 2, m.stdDevThreshold[pos]);
                        if (othSig[pos]) validNr++;
                    }
                if (validNr > 0) {
                    // compact predictands / weights of othersPoints, in order. With a single significant proxy (the common
                    // case) its values go to X as well (free once the elevation fit is over): the fit below then reads one
                    // array instead of chasing oi -> proxy for every model evaluation (same values)
                    int onlyPos = -1;
                    if (validNr == 1)
                        for (int i = 0; i < nProxy; i++)
                            if (i != ePos && m.selectedActive[i] && othSig[i]) onlyPos = i;
                    int nO = 0;
                    for (int base = 0; base < n; base += G) {
                        const int k = base + lane;
                        const bool ok = k < n && (w.flg[k] & 2);
                        const unsigned b = grpBallot<G>(ok);
                        if (ok) {
                            const int pos = nO + __popc(b & ltMask);
                            w.oi[pos] = w.idx[k];
                            w.Y[pos] = double(w.val[k]);
                            w.W[pos] = !isEqualF(w.w[k], kNoData) ? double(w.w[k]) : 1.;
                            if (onlyPos >= 0) w.X[pos] = double(st.proxy[(size_t)w.idx[k] * nProxy + onlyPos]);
                        }
                        nO += __popc(b);
                    }
                    grpSync<G>();
                    if (isLocal && nO < m.minPointsLocal) {                        // local detrending (:2100-2105)
#pragma unroll
                        for (int i = 0; i < PB_MAX_PROXY; i++) othSig[i] = false;
                    } else {
                        double pmin[PB_MAX_PROXY][2], pmax[PB_MAX_PROXY][2], delta[PB_MAX_PROXY][2];
                        for (int i = 0; i < nProxy; i++)
                            if (i != ePos && m.selectedActive[i] && othSig[i]) {
                                for (int j = 0; j < 2; j++) {
                                    pmin[nPred][j] = m.otherMin[i][j];
                                    pmax[nPred][j] = m.otherMax[i][j];
                                    delta[nPred][j] = m.otherDelta[i][j];
                                    othPar[nPred][j] = m.otherStart[i][j];
                                }
                                sigPos[nPred++] = i;
                            }
                        const float* proxy = st.proxy;
                        const int* oi = w.oi;
                        const bool coop = G == 32 && smWarp != nullptr;
                        if (nPred == 1) {
                            const double* px = w.X;
                            auto pred = [=](int j, int) { return px[j]; };
                            if (coop) lmOthersWarp<1>(pmin, pmax, othPar, delta, 100, 0.005, pred, w.Y, w.W, nO, smWarp);
                            else lmOthers<1>(1, pmin, pmax, othPar, delta, 100, 0.005, pred, w.Y, w.W, nO);
                        } else if (nPred == 2) {
                            const int p0 = sigPos[0], p1 = sigPos[1];
                            auto pred = [=](int j, int c) { return double(proxy[(size_t)oi[j] * nProxy + (c == 0 ? p0 : p1)]); };
                            if (coop) lmOthersWarp<2>(pmin, pmax, othPar, delta, 100, 0.005, pred, w.Y, w.W, nO, smWarp);
                            else lmOthers<2>(2, pmin, pmax, othPar, delta, 100, 0.005, pred, w.Y, w.W, nO);
                        } else if (nPred == 3 && coop) {
                            const int* sp = sigPos;
                            auto pred = [=](int j, int c) { return double(proxy[(size_t)oi[j] * nProxy + sp[c]]); };
                            lmOthersWarp<3>(pmin, pmax, othPar, delta, 100, 0.005, pred, w.Y, w.W, nO, smWarp);
                        } else {
                            const int* sp = sigPos;
                            auto pred = [=](int j, int c) { return double(proxy[(size_t)oi[j] * nProxy + sp[c]]); };
                            lmOthers<0>(nPred, pmin, pmax, othPar, delta, 100, 0.005, pred, w.Y, w.W, nO);
                        }
                    }
                }
            }
        }
    }
    grpSync<G>();

    // detrendingOtherProxies (:2173-2236)
    if (nPred > 0) {
        for (int k = lane; k < n; k += G) {
            if (!(w.flg[k] & 1)) continue;
            const size_t row = (size_t)w.idx[k] * nProxy;
            bool complete = true;
            double r = 0.0;
            for (int c = 0; c < nPred; c++) {
                const double v = double(st.proxy[row + sigPos[c]]);
                if (isEqualD(v, -9999)) { complete = false; break; }
                r += othPar[c][0] * v + othPar[c][1];
            }
            if (complete) w.val[k] -= float(r);
            else w.flg[k] &= ~1;
        }
        grpSync<G>();
    }


    fit.elevSig = elevSig;
#pragma unroll
    for (int j = 0; j < N; j++) fit.elevPar[j] = elevPar[j];
    fit.nPred = nPred;
#pragma unroll
    for (int i = 0; i < PB_MAX_PROXY; i++) {
        fit.sigPos[i] = sigPos[i];
        fit.othPar[i][0] = othPar[i][0];
        fit.othPar[i][1] = othPar[i][1];
    }
}

}  // namespace pb
