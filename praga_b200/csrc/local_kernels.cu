// local_kernels.cu — K2: local-detrending raster. One GROUP of lanes (half a warp for the 16-first-guess two-piece
// lapse rate, a full warp for the three-piece ones) owns one DEM cell and does, for that cell, what the loop of
// Project::interpolationDemLocalDetrending does (agrolib/project/project.cpp:3220-3250):
//
//   gis::getUtmXYFromRowCol                   gis/gis.cpp:806-810   (true cell centre in f64, narrowed to f32 at the calls)
//   getProxyValuesXY                          interpolation/interpolation.cpp:2620-2647
//   localSelection                            :1087-1170  ring search, selection order = ring-major then station index
//   preInterpolation -> multipleDetrendingMain :2444 / :1832 (weighted LM fits, detrend_device.cuh)
//   interpolate (idw) + retrend + clamp       :2502-2560, :1031-1051, :1288-1351
//   clearFitting / setCurrentCombination      (state is per cell here, nothing to reset)
//
// Built with -fmad=false: the FP64 fits follow the reference's operation order without FMA contraction, so the
// Levenberg-Marquardt descents take the same accept/reject path as on the CPU. Distances use IEEE f32 ops.
#include <stdlib.h>
#include "detrend_device.cuh"
#include "local.cuh"

namespace pb {

constexpr int kLocalThreads = 256;
constexpr int kCap = 64;      // stations per group held in shared memory; larger selections work from global scratch

struct WorkPtrs {
    double *X, *Y, *W;        // LM data (compact)
    int* idx;                 // station index of selected point k
    float *d, *w, *val;       // distance to the cell, regressionWeight, (detrended) value
    int* oi;                  // compact list of point slots used by the other-proxies fit
    unsigned char* flg;       // bit0: still in myPoints, bit1: in othersPoints
    unsigned char* act;       // isActive of compact elevation entry
};

__host__ __device__ inline size_t workBytes(int cap) { return (size_t(cap) * (3 * 8 + 5 * 4 + 2) + 15) / 16 * 16; }

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
__device__ bool proxyValidityFlagged(const LocalStations& st, const WorkPtrs& w, int n, int nProxy, int pos, int bit,
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

template <int G, int F>
__device__ float localCell(const LocalModel& m, const LocalStations& st, const float2* __restrict__ xy, float xf, float yf,
                           const double* pv, const WorkPtrs& gw, int gcap, unsigned char* tile, int* errFlag)
{
    constexpr int N = Fit<F>::N;
    const int lane = grpLane<G>();
    const int S = st.n;
    const int nProxy = m.nProxy;
    const bool useCode = m.useLapseRateCode != 0;
    const unsigned ltMask = (1u << lane) - 1u;

    // ------------------------------------------------------------------ localSelection (:1087-1170)
    int n = 0;
    float maxDistance = 0.f;
    const bool takeAll = unsigned(S) <= m.minPoints;
    if (takeAll) {
        for (int i = lane; i < S && i < gcap; i += G) {
            const float2 p = xy[i];
            const float dx = p.x - xf, dy = p.y - yf;
            gw.idx[i] = i;
            gw.d[i] = sqrtf(dx * dx + dy * dy);
        }
        n = S < gcap ? S : gcap;
        if (S > gcap && lane == 0) *errFlag = 1;
    } else {
        unsigned nrValid = 0, nrPrimaries = 0;
        float stepRadius = 7500.f, r0 = 0.f, r1 = 7500.f;
        bool beyondLastPoint = false;
        while (((!useCode && nrValid < m.minPoints) || (useCode && nrPrimaries < m.minPoints)) && !beyondLastPoint) {
            bool farther = false;
            for (int base = 0; base < S; base += G) {
                const int i = base + lane;
                bool in = false, prim = false;
                float d = 0.f;
                if (i < S) {
                    const float2 p = xy[i];
                    const float dx = p.x - xf, dy = p.y - yf;       // gis::computeDistance(x, y, px, py), gis.cpp:685-691
                    d = sqrtf(dx * dx + dy * dy);
                    in = d > r0 && d <= r1;
                    farther |= d > r1;
                    if (in) prim = checkLapseRateCodeDev(st.code[i], useCode, true);
                }
                const unsigned b = grpBallot<G>(in);
                if (b) {
                    const unsigned bp = grpBallot<G>(prim);
                    if (in) {
                        const int pos = n + __popc(b & ltMask);
                        if (pos < gcap) { gw.idx[pos] = i; gw.d[pos] = d; }
                        maxDistance = fmaxf(maxDistance, d);
                    }
                    n += __popc(b);
                    nrPrimaries += __popc(bp);
                }
            }
            nrValid = unsigned(n);
            beyondLastPoint = grpBallot<G>(farther) == 0u;
            if (nrValid > m.thr80) stepRadius = 1000.f;
            r0 = r1;
            r1 += stepRadius;
        }
        // group max of maxDistance
#pragma unroll
        for (int o = G / 2; o; o >>= 1) maxDistance = fmaxf(maxDistance, __shfl_xor_sync(grpMask<G>(), maxDistance, o, G));
        if (n > gcap) {
            if (lane == 0) *errFlag = 1;
            n = gcap;
        }
    }
    grpSync<G>();

    WorkPtrs w = gw;
    if (n <= kCap) {
        w = carve(tile, kCap);
        for (int k = lane; k < n; k += G) { w.idx[k] = gw.idx[k]; w.d[k] = gw.d[k]; }
    }
    for (int k = lane; k < n; k += G) {
        const int i = w.idx[k];
        float rw = st.weight ? st.weight[i] : kNoData;
        if (!takeAll && !isEqualF(maxDistance, 0.f)) {
            const float a = 1.f - w.d[k] / maxDistance;
            rw = (a > float(kEpsilon)) ? a : float(kEpsilon);                 // MAXVALUE(1 - d / maxDistance, EPSILON)
        }
        w.w[k] = rw;
        w.val[k] = st.value[i];
        w.flg[k] = 1;
    }
    grpSync<G>();

    // ------------------------------------------------------------------ multipleDetrendingMain (:1832-1859)
    const int ePos = m.elevationPos;
    bool elevSig = false;
    double elevPar[N];
#pragma unroll
    for (int j = 0; j < N; j++) elevPar[j] = 0;

    if (ePos >= 0 && m.selectedActive[ePos]) {
        // multipleDetrendingElevationFitting (:1862-1953): points with a usable lapse-rate code and an elevation value
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
        // proxyValidity (:1418-1457) on those points
        bool isValid = false;
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
        isValid = isValid && unsigned(nE) >= unsigned(m.minPointsLocal);          // local detrending (:1885)
        if (isValid) {
            const double R2 = bestFitElevation<G, F>(m.elevMin, m.elevMax, m.elevDelta, &m.firstGuess[0][0], m.nFirstGuess,
                                                     elevPar, 1000, 0.002, 0.005, w.X, w.Y, w.W, nE);
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
                    // compact predictands / weights of othersPoints, in order
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
                        }
                        nO += __popc(b);
                    }
                    grpSync<G>();
                    if (nO < m.minPointsLocal) {                                   // local detrending (:2100-2105)
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
                        if (nPred == 1) {
                            const int p0 = sigPos[0];
                            auto pred = [=](int j, int) { return double(proxy[(size_t)oi[j] * nProxy + p0]); };
                            lmOthers<1>(1, pmin, pmax, othPar, delta, 100, 0.005, pred, w.Y, w.W, nO);
                        } else if (nPred == 2) {
                            const int p0 = sigPos[0], p1 = sigPos[1];
                            auto pred = [=](int j, int c) { return double(proxy[(size_t)oi[j] * nProxy + (c == 0 ? p0 : p1)]); };
                            lmOthers<2>(2, pmin, pmax, othPar, delta, 100, 0.005, pred, w.Y, w.W, nO);
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

    // ------------------------------------------------------------------ interpolate (:2502-2560), idw over the subset
    float result;
    if (m.useRetrendOnly) result = 0.f;
    else {
        double sum = 0, sumWeights = 0;
        for (int k = 0; k < n; k++) {
            if (!(w.flg[k] & 1)) continue;
            if (!checkLapseRateCodeDev(st.code[w.idx[k]], useCode, false)) continue;      // excludeSupplemental = true
            const float d = w.d[k];
            if (d > 0.f) {
                const double dk = double(d) / 10000.;
                const double wt = 1.0 / (dk * dk * dk);
                sumWeights += wt;
                sum += double(w.val[k]) * wt;
            }
        }
        result = (sumWeights > 0.0) ? float(sum / sumWeights) : kNoData;
    }
    grpSync<G>();     // the tile is reused by the next cell
    if (isEqualF(result, kNoData)) return kNoData;

    if (!m.useDoNotRetrend) {
        // retrend, multiple detrending (:1298-1313) with getMultipleDetrendingValues (:2563-2617) and functionSum
        float retrendValue = 0.f;
        bool ok = true;
        double total = 0.;
        int nAct = 0;
        if (ePos >= 0 && m.selectedActive[ePos] && elevSig) {
            if (pv[ePos] == double(kNoData)) ok = false;
            else {
                double pc[N];
#pragma unroll
                for (int j = 0; j < N; j++) pc[j] = elevPar[j];
                Fit<F>::clamp(pc);
                total += Fit<F>::eval(pv[ePos], pc);
            }
            nAct++;
        }
        for (int c = 0; c < nPred && ok; c++) {
            const double v = pv[sigPos[c]];
            if (v == double(kNoData)) ok = false;
            else total += othPar[c][0] * v + othPar[c][1];
            nAct++;
        }
        if (ok && nAct > 0) retrendValue = float(total);
        result += retrendValue;
    }
    switch (m.clampMode) {
    case CLAMP_PREC: return (result < m.rainfallThreshold) ? 0.f : result;
    case CLAMP_RH: return fmaxf(0.f, fminf(result, 100.f));
    case CLAMP_POS: return fmaxf(result, 0.f);
    default: return result;
    }
}

__device__ __forceinline__ unsigned orderedKeyL(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

template <int G, int F, int MINB>
__global__ void __launch_bounds__(kLocalThreads, MINB)
local_detrending_kernel(const LocalModel* __restrict__ mp, LocalStations st, DevGrid dem, const int* __restrict__ validIdx,
                        int nTargets, LocalScratch scr, int xyInSmem, float* __restrict__ out, unsigned* __restrict__ minmax)
{
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int groupsPerCta = kLocalThreads / G;
    const LocalModel& m = *mp;
    const int grp = threadIdx.x / G;
    const int lane = grpLane<G>();
    const size_t tileBytes = workBytes(kCap);
    unsigned char* tile = smem + grp * tileBytes;
    float2* sxy = reinterpret_cast<float2*>(smem + groupsPerCta * tileBytes);
    if (xyInSmem) {
        for (int i = threadIdx.x; i < st.n; i += kLocalThreads) sxy[i] = st.xy[i];
        __syncthreads();
    }
    const float2* xy = xyInSmem ? sxy : st.xy;
    const long long groupsTotal = (long long)gridDim.x * groupsPerCta;
    const long long gid = (long long)blockIdx.x * groupsPerCta + grp;
    const WorkPtrs gw = carve(scr.base + gid * scr.perGroup, scr.cap);

    float vmin = INFINITY, vmax = -INFINITY;
    for (long long t = gid; t < nTargets; t += groupsTotal) {
        const int cell = validIdx[t];
        const int row = cell / dem.nrCols, col = cell - row * dem.nrCols;
        // gis::getUtmXYFromRowCol (gis.cpp:806-810), then narrowed to float by the callee signatures (interpolation.h:77-95,162)
        const double x = dem.llx + dem.cellSize * (col + 0.5);
        const double y = dem.lly + dem.cellSize * (dem.nrRows - row - 0.5);
        const float xf = float(x), yf = float(y);
        double pv[PB_MAX_PROXY];
        for (int i = 0; i < m.nProxy; i++) {
            pv[i] = double(kNoData);
            if (m.selectedActive[i] && m.gridSlot[i] >= 0) {
                const DevGrid& g = m.grid[m.gridSlot[i]];
                const float v = gridValueXY(g, double(xf), double(yf));
                if (v != g.flag) pv[i] = double(v);
            }
        }
        const float o = localCell<G, F>(m, st, xy, xf, yf, pv, gw, scr.cap, tile, scr.error);
        if (lane == 0) {
            out[cell] = o;
            if (!isEqualF(o, dem.flag) && !isEqualF(o, kNoData)) {
                vmin = fminf(vmin, o);
                vmax = fmaxf(vmax, o);
            }
        }
    }
    if (lane == 0 && minmax) {
        if (vmin != INFINITY) atomicMin(&minmax[0], orderedKeyL(vmin));
        if (vmax != -INFINITY) atomicMax(&minmax[1], orderedKeyL(vmax));
    }
}

size_t localScratchBytesPerGroup(int cap) { return workBytes(cap); }

// persistent grid: every SM filled to the occupancy the fit kernel's registers allow (at most 2 CTAs per SM)
static int localCtas()
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms * 4;
}
int localGroupsTotal(int groupLanes) { return localCtas() * (kLocalThreads / groupLanes); }

template <int G, int F, int MINB>
static cudaError_t launchT(const LocalModel* dModel, const LocalStations& st, const DevGrid& dem, const int* validIdx,
                           int nTargets, const LocalScratch& scratch, float* out, unsigned* minmax, cudaStream_t s)
{
    constexpr int groupsPerCta = kLocalThreads / G;
    size_t smem = groupsPerCta * workBytes(kCap);
    int xyInSmem = 0;
    if (smem + size_t(st.n) * sizeof(float2) <= 100 * 1024) {
        xyInSmem = 1;
        smem += size_t(st.n) * sizeof(float2);
    }
    auto k = local_detrending_kernel<G, F, MINB>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int perSm = 1, dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k, kLocalThreads, smem);
    if (e != cudaSuccess) return e;
    perSm = perSm < 1 ? 1 : (perSm > 4 ? 4 : perSm);
    const long long need = (nTargets + groupsPerCta - 1) / groupsPerCta;
    const int grid = (int)(need < (long long)sms * perSm ? need : (long long)sms * perSm);
    k<<<grid, kLocalThreads, smem, s>>>(dModel, st, dem, validIdx, nTargets, scratch, xyInSmem, out, minmax);
    return cudaGetLastError();
}

cudaError_t launchLocalDetrendingRaster(const LocalModel* dModel, const LocalStations& st, const DevGrid& dem,
                                        const int* validIdx, int nTargets, const LocalScratch& scratch, int groupLanes,
                                        int elevFunction, float* out, unsigned* minmax, cudaStream_t s)
{
    if (nTargets <= 0) return cudaSuccess;
    static const int minb = getenv("PB_LOCAL_MINB") ? atoi(getenv("PB_LOCAL_MINB")) : 2;
#define PB_LOCAL_CASE(G_, F_)                                                                               \
    return minb == 2 ? launchT<G_, F_, 2>(dModel, st, dem, validIdx, nTargets, scratch, out, minmax, s)     \
         : minb == 3 ? launchT<G_, F_, 3>(dModel, st, dem, validIdx, nTargets, scratch, out, minmax, s)     \
         : minb == 4 ? launchT<G_, F_, 4>(dModel, st, dem, validIdx, nTargets, scratch, out, minmax, s)     \
                     : launchT<G_, F_, 1>(dModel, st, dem, validIdx, nTargets, scratch, out, minmax, s)
    if (groupLanes == 16) {
        switch (elevFunction) {
        case PB_FIT_PIECEWISE_THREE: PB_LOCAL_CASE(16, PB_FIT_PIECEWISE_THREE);
        case PB_FIT_PIECEWISE_THREE_FREE: PB_LOCAL_CASE(16, PB_FIT_PIECEWISE_THREE_FREE);
        default: PB_LOCAL_CASE(16, PB_FIT_PIECEWISE_TWO);
        }
    }
    switch (elevFunction) {
    case PB_FIT_PIECEWISE_THREE: PB_LOCAL_CASE(32, PB_FIT_PIECEWISE_THREE);
    case PB_FIT_PIECEWISE_THREE_FREE: PB_LOCAL_CASE(32, PB_FIT_PIECEWISE_THREE_FREE);
    default: PB_LOCAL_CASE(32, PB_FIT_PIECEWISE_TWO);
    }
#undef PB_LOCAL_CASE
}

}  // namespace pb
