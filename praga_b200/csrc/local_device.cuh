// local_device.cuh — device functions of the local-detrending cell shared by the fused kernel (local_kernels.cu) and the
// three-kernel pipeline (local_pipeline.cu): localSelection and the estimate of one cell. Include only from TUs built with
// -fmad=false.
#pragma once
#include "multiple_detrend.cuh"

namespace pb {

// localSelection (:1087-1170) for the target (xf, yf) by one group: returns n and the working arrays w (the shared-memory
// tile when n <= tileCap, else the group's global scratch) with idx, d, w (regression weight), val, flg filled
template <int G>
__device__ int localSelect(const LocalModel& m, const LocalStations& st, const float2* __restrict__ xy, float xf, float yf,
                           const WorkPtrs& gw, int gcap, unsigned char* tile, int tileCap, int* errFlag, WorkPtrs& w,
                           float& localRadius, const int* __restrict__ sList = nullptr, int nList = 0, int ccap = -1,
                           float listRadius = 0.f)
{
    // listRadius: every station within listRadius (>= the candidate radius) of the target is in sList: rings that end inside it
    // but beyond the candidate radius (raster borders, sparse neighbourhoods) walk the list instead of all S stations
    // ccap: capacity of the candidate arrays gw.oi / gw.w when it differs from gcap (the pipeline keeps them in shared
    // memory). errFlag == nullptr: a selection larger than gcap is the caller's business — the true n is returned and the
    // working arrays are left unfilled
    // sList (optional): ascending station indices that contain EVERY station within the candidate radius of this target (a
    // prefilter shared by the targets of a CTA, local_pipeline.cu); the candidate pass then walks the list instead of all S
    const int lane = grpLane<G>();
    const int S = st.n;
    const bool useCode = m.useLapseRateCode != 0;
    const unsigned ltMask = (1u << lane) - 1u;

    // ------------------------------------------------------------------ localSelection (:1087-1170)
    int n = 0;
    float maxDistance = 0.f;
    const bool takeAll = unsigned(S) <= m.minPoints;
    if (ccap < 0) ccap = gcap;
    if (takeAll) {
        for (int i = lane; i < S && i < gcap; i += G) {
            const float2 p = xy[i];
            const float dx = p.x - xf, dy = p.y - yf;
            gw.idx[i] = i;
            gw.d[i] = sqrtf(dx * dx + dy * dy);
        }
        n = S < gcap ? S : gcap;
        if (S > gcap && lane == 0 && errFlag) *errFlag = 1;
    } else {
        // Candidate pass: ONE scan of all stations keeps, in station order, those within the candidate radius (so that the
        // rings below walk a few dozen entries instead of all S) and the largest distance (the `beyond last point` test).
        // d2 <= rcap2 is conservative (slack >> the rounding of sqrtf), so every station of a ring with r1 <= rcap is a
        // candidate; rings beyond rcap (raster borders, sparse corners) scan the full list as before.
        float rcap = m.candRadius;
        int nc = 0;
        float dmax = 0.f;
        if (rcap > 0.f) {
            // every lane scans ONE contiguous block of the station list: counts first, then (offsets from a prefix over the
            // group's lanes) writes its candidates — station order is kept without a vote per chunk of G stations, which
            // with 4 lanes per cell was a third of the kernel's stall samples
            const float rcap2 = rcap * rcap * 1.0001f;
            float dmax2 = 0.f;
            const int L0 = sList ? nList : S;
            const int per = (L0 + G - 1) / G;
            const int i0 = lane * per, i1 = (i0 + per < L0) ? i0 + per : L0;
            int mine = 0;
            for (int k = i0; k < i1; k++) {
                const int i = sList ? sList[k] : k;
                const float2 p = xy[i];
                const float dx = p.x - xf, dy = p.y - yf;
                const float d2 = dx * dx + dy * dy;
                mine += d2 <= rcap2;
                dmax2 = fmaxf(dmax2, d2);
            }
            int offset = 0;
#pragma unroll
            for (int l = 0; l < G; l++) {
                const int cl = __shfl_sync(grpMask<G>(), mine, l, G);
                offset += (l < lane) ? cl : 0;
                nc += cl;
            }
            if (nc <= ccap) {
                int pos = offset;
                for (int k = i0; k < i1; k++) {
                    const int i = sList ? sList[k] : k;
                    const float2 p = xy[i];
                    const float dx = p.x - xf, dy = p.y - yf;
                    const float d2 = dx * dx + dy * dy;
                    if (d2 <= rcap2) { gw.oi[pos] = i; gw.w[pos] = sqrtf(d2); pos++; }
                }
            }
#pragma unroll
            for (int o = G / 2; o; o >>= 1) dmax2 = fmaxf(dmax2, __shfl_xor_sync(grpMask<G>(), dmax2, o, G));
            dmax = sqrtf(dmax2);                  // sqrtf is monotonic: max of the distances
            if (sList && nList < S) dmax = INFINITY;      // a station outside the shared list lies beyond the candidate radius
            if (nc > ccap) rcap = 0.f;            // too dense for the candidate list: full scans
            grpSync<G>();
        }
        unsigned nrValid = 0, nrPrimaries = 0;
        float stepRadius = 7500.f, r0 = 0.f, r1 = 7500.f;
        bool beyondLastPoint = false;
        while (((!useCode && nrValid < m.minPoints) || (useCode && nrPrimaries < m.minPoints)) && !beyondLastPoint) {
            const bool compact = rcap > 0.f && r1 <= rcap;
            const bool useList = !compact && sList != nullptr && r1 <= listRadius;
            const int L = compact ? nc : (useList ? nList : S);
            bool farther = false;
            // one contiguous block of the list per lane: count the ring's members, prefix over the group, write in place —
            // ring-major / index order without a vote per chunk
            const int per = (L + G - 1) / G;
            const int k0 = lane * per, k1 = (k0 + per < L) ? k0 + per : L;
            int mine = 0, minePrim = 0;
            for (int k = k0; k < k1; k++) {
                float d;
                int i = k;
                if (compact) { i = gw.oi[k]; d = gw.w[k]; }
                else {
                    if (useList) i = sList[k];
                    const float2 p = xy[i];
                    const float dx = p.x - xf, dy = p.y - yf;       // gis::computeDistance(x, y, px, py), gis.cpp:685-691
                    d = sqrtf(dx * dx + dy * dy);
                    farther |= d > r1;
                }
                if (d > r0 && d <= r1) {
                    mine++;
                    minePrim += checkLapseRateCodeDev(st.code[i], useCode, true);
                }
            }
            int offset = 0, total = 0, totalPrim = 0;
#pragma unroll
            for (int l = 0; l < G; l++) {
                const int cl = __shfl_sync(grpMask<G>(), mine, l, G);
                offset += (l < lane) ? cl : 0;
                total += cl;
                totalPrim += __shfl_sync(grpMask<G>(), minePrim, l, G);
            }
            if (total) {
                int pos = n + offset;
                for (int k = k0; k < k1 && mine; k++) {
                    float d;
                    int i = k;
                    if (compact) { i = gw.oi[k]; d = gw.w[k]; }
                    else {
                        if (useList) i = sList[k];
                        const float2 p = xy[i];
                        const float dx = p.x - xf, dy = p.y - yf;
                        d = sqrtf(dx * dx + dy * dy);
                    }
                    if (d > r0 && d <= r1) {
                        if (pos < gcap) { gw.idx[pos] = i; gw.d[pos] = d; }
                        maxDistance = fmaxf(maxDistance, d);
                        pos++;
                    }
                }
                n += total;
                nrPrimaries += totalPrim;
            }
            nrValid = unsigned(n);
            // a station outside the shared list lies beyond listRadius >= r1: it is "farther"
            if (compact) beyondLastPoint = !(dmax > r1);
            else if (useList && nList < S) beyondLastPoint = false;
            else beyondLastPoint = grpBallot<G>(farther) == 0u;
            if (nrValid > m.thr80) stepRadius = 1000.f;
            r0 = r1;
            r1 += stepRadius;
        }
        // group max of maxDistance
#pragma unroll
        for (int o = G / 2; o; o >>= 1) maxDistance = fmaxf(maxDistance, __shfl_xor_sync(grpMask<G>(), maxDistance, o, G));
        if (n > gcap) {
            if (!errFlag) { grpSync<G>(); return n; }
            if (lane == 0) *errFlag = 1;
            n = gcap;
        }
    }
    grpSync<G>();
    localRadius = takeAll ? m.shepardRadius : maxDistance;        // setLocalRadius (:1098, :1167)

    w = gw;
    if (n <= tileCap) {
        w = carve(tile, tileCap);
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
    return n;
}

// interpolate (:2502-2560: idw over the detrended subset, retrend with the fit, clamp) for one target, group-uniform
template <int G, int F>
__device__ float localFinish(const LocalModel& m, const LocalStations& st, const WorkPtrs& w, int n, const double* pv,
                             const MdFit<Fit<F>::N>& fit, bool excludeSupplemental, float xf, float yf, float localRadius)
{
    constexpr int N = Fit<F>::N;
    const bool useCode = m.useLapseRateCode != 0;
    const int ePos = m.elevationPos;
    const bool elevSig = fit.elevSig;
    const int nPred = fit.nPred;
    const double* elevPar = fit.elevPar;
    const int* sigPos = fit.sigPos;
    const double (*othPar)[2] = fit.othPar;

    // ------------------------------------------------------------------ interpolate (:2502-2560), idw over the subset
    float result;
    if (m.useRetrendOnly) result = 0.f;
    else if (m.method == PB_METHOD_SHEPARD_MODIFIED) {
        // modifiedShepardIdw (:948-1028) over the whole subset with radius = getLocalRadius() (:2525-2527). The points still in
        // myPoints (flag bit 0) are the vector; a point computeDistances excludes has distance 0 and weighs nothing. The lanes
        // share the O(n^2) direction term (one row each); every sum over points is taken in point order. s -> w.X, t -> w.Y.
        const int lane = grpLane<G>();
        const float radius = localRadius;
        for (int k = lane; k < n; k += G) {
            float d = w.d[k];
            if (!(w.flg[k] & 1) || (excludeSupplemental && !checkLapseRateCodeDev(st.code[w.idx[k]], useCode, false))) d = 0.f;
            w.d[k] = d;
            w.X[k] = (d > 0.f && d <= radius) ? double((radius - d) / (radius * d)) : 0.;
            w.Y[k] = 0.;
        }
        grpSync<G>();
        double weightSum = 0.0;
        for (int k = 0; k < n; k++) {
            const float d = w.d[k];
            if ((w.flg[k] & 1) && d > 0.f && d <= radius) weightSum += w.X[k];
        }
        result = kNoData;
        if (weightSum != 0.0) {
            const double invWeightSum = 1.0 / weightSum;
            // parameter order of the definition (:948-949): inside, "x" is the caller's y and "y" the caller's x
            const double X = double(yf), Y = double(xf);
            for (int a = lane; a < n; a += G) {
                if (!(w.flg[a] & 1) || w.X[a] == 0.0 || w.d[a] <= 0.f) continue;
                const double xa = st.xd[w.idx[a]], ya = st.yd[w.idx[a]];
                double tt = 0.;
                for (int b = 0; b < n; b++) {
                    if (a == b || !(w.flg[b] & 1) || w.X[b] == 0.0 || w.d[b] <= 0.f) continue;
                    const double cosine = ((X - xa) * (X - st.xd[w.idx[b]]) + (Y - ya) * (Y - st.yd[w.idx[b]])) /
                                          double(w.d[a] * w.d[b]);
                    tt += w.X[b] * (1.0 - cosine);
                }
                w.Y[a] = tt * invWeightSum;
            }
            grpSync<G>();
            double wsum = 0.0;
            for (int k = 0; k < n; k++)
                if (w.flg[k] & 1) wsum += w.X[k] * w.X[k] * (1.0 + w.Y[k]);
            const double inv = 1.0 / wsum;
            double r = 0.0;
            for (int k = 0; k < n; k++)
                if (w.flg[k] & 1) r += ((w.X[k] * w.X[k] * (1.0 + w.Y[k])) * inv) * double(w.val[k]);
            result = float(r);
        }
    } else if (m.method == PB_METHOD_SHEPARD) {
        // shepardIdw (:871-945) behind shepardSearchNeighbour (:806-868) over the subset: the vector is the points still in
        // myPoints (flag bit 0), a point computeDistances excludes has distance 0. At most 10 points survive, so every lane
        // of the group carries the whole estimate (group-uniform, nothing shared is written). Equal distances: subset order.
        constexpr int kMinPts = 5, kMaxPts = 10;      // SHEPARD_MIN_NRPOINTS / SHEPARD_MAX_NRPOINTS
        unsigned nrPoints = 0;
        for (int k = 0; k < n; k++) nrPoints += (w.flg[k] & 1) ? 1u : 0u;
        result = kNoData;
        if (nrPoints > 0) {
            // computeShepardInitialRadius(getPointsBoundingBoxArea(), inputPoints.size(), SHEPARD_AVG_NRPOINTS = 8) (:800-803)
            const float R0 = sqrtf((8u * m.bboxArea) / (float(3.1415926535898) * nrPoints));
            float d10[kMaxPts], d5[kMinPts];
            int i10[kMaxPts], i5[kMinPts];
            int n10 = 0, n5 = 0, cnt = 0;
            auto insert = [](float* dd, int* ii, int& cntIn, int K, float d, int i) {
                if (cntIn == K && !(d < dd[K - 1])) return;
                int pos = cntIn < K ? cntIn : K - 1;
                while (pos > 0 && d < dd[pos - 1]) { dd[pos] = dd[pos - 1]; ii[pos] = ii[pos - 1]; pos--; }
                dd[pos] = d; ii[pos] = i;
                if (cntIn < K) cntIn++;
            };
            for (int k = 0; k < n; k++) {
                if (!(w.flg[k] & 1)) continue;
                float d = w.d[k];
                if (excludeSupplemental && !checkLapseRateCodeDev(st.code[w.idx[k]], useCode, false)) d = 0.f;
                const bool sortable = !(fabs(double(d)) < kEpsilon);      // sortPointsByDistance drops |d| < EPSILON (:134)
                if (sortable) insert(d5, i5, n5, kMinPts, d, k);
                if (d <= R0 && d > 0.f) {
                    cnt++;
                    if (sortable) insert(d10, i10, n10, kMaxPts, d, k);
                }
            }
            float sd[kMaxPts];
            int si[kMaxPts];
            int kk = 0;
            float radius = kNoData;
            if (cnt < kMinPts) {
                kk = n5;
                for (int q = 0; q < kk; q++) { sd[q] = d5[q]; si[q] = i5[q]; }
                if (kk > 0) radius = d5[kk - 1] + float(kEpsilon);
            } else if (cnt > kMaxPts) {
                kk = n10;
                for (int q = 0; q < kk; q++) { sd[q] = d10[q]; si[q] = i10[q]; }
                radius = d10[kk - 1] + float(kEpsilon);
            } else {
                kk = n10;                           // the neighbourhood itself, in subset order
                for (int q = 0; q < kk; q++) { sd[q] = d10[q]; si[q] = i10[q]; }
                for (int a = 1; a < kk; a++) {
                    const float vd = sd[a];
                    const int vi = si[a];
                    int b = a;
                    while (b > 0 && si[b - 1] > vi) { sd[b] = sd[b - 1]; si[b] = si[b - 1]; b--; }
                    sd[b] = vd; si[b] = vi;
                }
                radius = R0;
            }
            if (kk > 0) {
                double S[kMaxPts], T[kMaxPts], px[kMaxPts], py[kMaxPts];
                for (int q = 0; q < kk; q++) { px[q] = st.xd[w.idx[si[q]]]; py[q] = st.yd[w.idx[si[q]]]; S[q] = 0.; T[q] = 0.; }
                double weightSum = 0.;
                const double radius_3 = double(radius) / 3.;
                const double radius_27_4 = 6.75 / double(radius);
                for (int q = 0; q < kk; q++) {
                    const float d = sd[q];
                    if (d > 0.f) {
                        if (double(d) <= radius_3) S[q] = double(1.f / d);
                        else if (d <= radius) {
                            const double tmp = double((d / radius) - 1.f);
                            S[q] = radius_27_4 * tmp * tmp;
                        } else S[q] = 0.;
                        weightSum += S[q];
                    }
                }
                if (weightSum != 0.) {
                    const double X = double(xf), Y = double(yf);
                    for (int a = 0; a < kk; a++) {
                        double tt = 0.;
                        for (int b = 0; b < kk; b++)
                            if (a != b) {
                                const double cosine = ((X - px[a]) * (X - px[b]) + (Y - py[a]) * (Y - py[b])) / double(sd[a] * sd[b]);
                                tt += S[b] * (1. - cosine);
                            }
                        tt /= weightSum;
                        T[a] = tt;
                    }
                    double wsum = 0.;
                    for (int a = 0; a < kk; a++) {
                        T[a] = S[a] * S[a] * (1. + T[a]);
                        wsum += T[a];
                    }
                    double r = 0.;
                    for (int a = 0; a < kk; a++) r += (T[a] / wsum) * double(w.val[si[a]]);
                    result = float(r);
                }
            }
        }
    } else {
        double sum = 0, sumWeights = 0;
        for (int k = 0; k < n; k++) {
            if (!(w.flg[k] & 1)) continue;
            if (excludeSupplemental && !checkLapseRateCodeDev(st.code[w.idx[k]], useCode, false)) continue;
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

}  // namespace pb
