// classic_kernels.cu — K7: the station-space side of preInterpolation that is not the Levenberg-Marquardt branch
// (agrolib/interpolation/interpolation.cpp, paths relative to the reference tree):
//   detrending()                 :1380-1415  per-proxy sequential regression + detrendPoints :1236-1285
//     regressionOrography        :1354-1377  -> regressionOrographyT :433-798 (thermal-inversion heuristic),
//                                               regressionSimpleT :368-406, regressionGeneric :346-365, regressionSimple :304-344
//     statistics::linearRegression  mathFunctions/statistics.cpp:1030-1092
//   optimalDetrending()          :2395-2441  every proxy combination: detrend -> leave-one-out -> MAE -> argmin
//   topographicDistanceOptimize  :2356-2392  golden section on kh, each probe = leave-one-out MAE at kh = int(khFloat)
//   computeResiduals             interpolation/spatialControl.cpp:97-155, computeErrorCrossValidation :305-328
//
// Mapping. These are chains of data-dependent scalar decisions over O(S) sums whose outcome (which combination, which kh)
// changes the whole raster, so they are reproduced operation by operation (this TU is built with -fmad=false): the
// parallelism is ACROSS problems. A problem = one (time step, proxy combination): one lane runs detrending() for it, a
// warp = 32 problems, the working values are stored station-major (n x NP) so that the lanes of a warp read one station
// of 32 problems with one coalesced access, station constants are warp-uniform broadcasts. The leave-one-out is exact
// (f32 distance, f64 weights in station order like inverseDistanceWeighted :1031-1051): thread = (target station, kh);
// the weight of a (target, station) pair does not depend on the combination, so it is computed once and applied to the
// detrended values of 8 problems per pass. kh only takes the integer values 0..maxKh, so the golden section runs on the
// host over a table of MAE(problem, kh) built in one launch instead of ~35 sequential leave-one-out passes.
#include <stdlib.h>
#include "classic.cuh"
#include "epilogue.cuh"
#include "td_device.cuh"

namespace pb {

namespace {

constexpr int kMaxIntervals = 192;
constexpr int kMinRegressionPoints = 5;      // MIN_REGRESSION_POINTS, interpolationConstants.h:4

// meteo/meteo.cpp:1127-1133
__device__ __forceinline__ bool checkLapseRateCodeDev(int code, bool useLapseRateCode, bool useSupplemental)
{
    if (useSupplemental) return !useLapseRateCode || code == PB_LAPSE_PRIMARY || code == PB_LAPSE_SUPPLEMENTAL;
    return !useLapseRateCode || code == PB_LAPSE_PRIMARY || code == PB_LAPSE_SECONDARY;
}

struct Ctx {
    const ClassicSetup& cs;
    const ClassicStations& st;
    float* val;        // this problem's column of the working values: val[i * NP]
    int NP;
    bool warp;         // a whole warp runs this problem (few problems: latency form), else one lane
    __device__ __forceinline__ float value(int i) const { return val[(size_t)i * NP]; }
    __device__ __forceinline__ bool active(int i) const { return st.active ? st.active[i] != 0 : true; }
    __device__ __forceinline__ float proxy(int i, int pos) const { return st.proxy[(size_t)i * cs.nProxy + pos]; }
};

// The items of an ordered sum. One lane: a plain loop. A whole warp (few problems, the latency of ONE problem is the cost):
// the lanes evaluate 32 items at a time (loads and predicates in parallel), then every lane receives the accepted ones in
// index order through shuffles and feeds its own copy of the accumulators: the float / double sums see exactly the
// reference's sequence.
template <class Eval, class Acc>
__device__ __forceinline__ void orderedItems(bool warp, int n, Eval eval, Acc acc)
{
    if (!warp) {
        for (int i = 0; i < n; i++) {
            float a, b;
            if (eval(i, a, b)) acc(a, b);
        }
        return;
    }
    const int lane = threadIdx.x & 31;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        float a = 0.f, b = 0.f;
        const bool ok = i < n && eval(i, a, b);
        unsigned mm = __ballot_sync(0xffffffffu, ok);
        while (mm) {
            const int l = __ffs(mm) - 1;
            mm &= mm - 1;
            acc(__shfl_sync(0xffffffffu, a, l), __shfl_sync(0xffffffffu, b, l));
        }
    }
}

// statistics::linearRegression over the items `get` yields (x, y as float; returns false to skip the item)
template <class G>
__device__ void linearRegressionDev(int n, G get, float& q, float& m, float& r2, bool warp = false)
{
    double SUMx = 0, SUMy = 0, SUMxy = 0, SUMxx = 0;
    long nrValid = 0;
    auto valid = [&](int i, float& x, float& y) -> bool { return get(i, x, y) && x != kNoData && y != kNoData; };
    orderedItems(warp, n, valid, [&](float x, float y) {
        nrValid++;
        SUMx += x;
        SUMy += y;
        SUMxy += x * y;
        SUMxx += x * x;
    });
    const double AVGy = SUMy / double(nrValid);
    const double AVGx = SUMx / double(nrValid);
    m = float((SUMxy - SUMx * SUMy / double(nrValid)) / (SUMxx - SUMx * SUMx / double(nrValid)));
    q = float(AVGy - double(m) * AVGx);
    double SUM_dy = 0, SUMres = 0;
    orderedItems(warp, n, valid, [&](float x, float y) {
        double dy = double(y - (q + m * x));
        dy *= dy;
        SUM_dy += dy;
        double res = double(y) - AVGy;
        res *= res;
        SUMres += res;
    });
    r2 = float((SUMres - SUM_dy) / SUMres);
}

// getMaxHeight / getMinHeight, interpolation.cpp:49-71
__device__ float maxHeightDev(const Ctx& c)
{
    float zMax = kNoData;
    // the f64 elevation travels as its two 32-bit halves
    orderedItems(c.warp, c.cs.n, [&](int i, float& hi, float& lo) -> bool {
        if (!(c.value(i) != kNoData && c.active(i) && checkLapseRateCodeDev(c.st.code[i], c.cs.useLapseRateCode, true))) return false;
        const double z = c.st.z[i];
        hi = __int_as_float(__double2hiint(z));
        lo = __int_as_float(__double2loint(z));
        return true;
    }, [&](float hi, float lo) {
        const double z = __hiloint2double(__float_as_int(hi), __float_as_int(lo));
        if (zMax == kNoData || z > double(zMax)) zMax = float(z);
    });
    return zMax;
}
__device__ float minHeightDev(const Ctx& c)
{
    float zMin = kNoData;
    orderedItems(c.warp, c.cs.n, [&](int i, float& hi, float& lo) -> bool {
        const double z = c.st.z[i];
        if (!(z != double(kNoData) && c.active(i) && checkLapseRateCodeDev(c.st.code[i], c.cs.useLapseRateCode, true))) return false;
        hi = __int_as_float(__double2hiint(z));
        lo = __int_as_float(__double2loint(z));
        return true;
    }, [&](float hi, float lo) {
        const double z = __hiloint2double(__float_as_int(hi), __float_as_int(lo));
        if (zMin == kNoData || z < double(zMin)) zMin = float(z);
    });
    return zMin;
}

// regressionSimple, interpolation.cpp:304-344
__device__ bool regressionSimpleDev(const Ctx& c, int pos, float& coeff, float& intercept, float& r2)
{
    coeff = intercept = r2 = kNoData;
    const bool heightPos = pos == c.cs.indexHeight;
    auto get = [&](int i, float& x, float& y) -> bool {
        if (!c.active(i)) return false;
        if (heightPos && !checkLapseRateCodeDev(c.st.code[i], c.cs.useLapseRateCode, true)) return false;
        const float pv = c.proxy(i, pos);
        if (isEqualF(pv, kNoData)) return false;
        x = pv;
        y = c.value(i);
        return true;
    };
    int cnt = 0;
    orderedItems(c.warp, c.cs.n, get, [&](float, float) { cnt++; });
    if (cnt < kMinRegressionPoints) return false;
    linearRegressionDev(c.cs.n, get, intercept, coeff, r2, c.warp);
    return true;
}

#define PB_SET(P, FIELD, BIT, V) do { (P).FIELD = (V); (P).written |= (BIT); } while (0)

// regressionGeneric, interpolation.cpp:346-365
__device__ bool regressionGenericDev(const Ctx& c, int pos, ClassicProxyState& p)
{
    float q, m, r2;
    if (!regressionSimpleDev(c, pos, m, q, r2)) return false;
    PB_SET(p, slope, CF_SLOPE, m);
    PB_SET(p, intercept, CF_INTERCEPT, q);
    PB_SET(p, r2, CF_R2, r2);
    PB_SET(p, T0, CF_T0, q);
    PB_SET(p, invSig, CF_INVSIG, 0);
    PB_SET(p, invLapse, CF_INVLAPSE, kNoData);
    return r2 >= c.cs.minRegressionR2;
}

// Crit3DProxy::initializeOrography, interpolationSettings.cpp:779-792
__device__ void initializeOrographyDev(ClassicProxyState& p)
{
    PB_SET(p, H0, CF_H0, 0.f);
    PB_SET(p, H1, CF_H1, kNoData);
    PB_SET(p, T0, CF_T0, kNoData);
    PB_SET(p, T1, CF_T1, kNoData);
    PB_SET(p, invLapse, CF_INVLAPSE, kNoData);
    PB_SET(p, slope, CF_SLOPE, kNoData);
    PB_SET(p, intercept, CF_INTERCEPT, kNoData);
    PB_SET(p, r2, CF_R2, kNoData);
    PB_SET(p, invSig, CF_INVSIG, 0);
}

// regressionSimpleT, interpolation.cpp:368-406
__device__ bool regressionSimpleTDev(const Ctx& c, int pos, ClassicProxyState& p)
{
    float slope, t0, r2;
    initializeOrographyDev(p);
    if (!regressionSimpleDev(c, pos, slope, t0, r2)) return false;
    if (r2 < c.cs.minRegressionR2) return false;
    PB_SET(p, slope, CF_SLOPE, slope);
    PB_SET(p, T0, CF_T0, t0);
    PB_SET(p, r2, CF_R2, r2);
    if (slope > 0 && !c.cs.isElaborationVar) {
        PB_SET(p, invLapse, CF_INVLAPSE, slope);
        const float mz = maxHeightDev(c);
        const float maxZ = mz < c.cs.maxHeightInversion ? mz : c.cs.maxHeightInversion;     // MINVALUE
        PB_SET(p, T1, CF_T1, t0 + slope * maxZ);
        PB_SET(p, H1, CF_H1, maxZ);
        PB_SET(p, slope, CF_SLOPE, c.cs.climateLapseRate);
        PB_SET(p, invSig, CF_INVSIG, 1);
    } else {
        PB_SET(p, invSig, CF_INVSIG, 0);
        PB_SET(p, invLapse, CF_INVLAPSE, kNoData);
    }
    return true;
}

// findHeightIntervalAvgValue, interpolation.cpp:409-431
__device__ float heightIntervalAvgDev(const Ctx& c, float heightInf, float heightSup, float maxPointsZ)
{
    float sum = 0.f, nValues = 0.f;
    orderedItems(c.warp, c.cs.n, [&](int i, float& v, float&) -> bool {
        const double z = c.st.z[i];
        if (!(z != double(kNoData) && c.active(i) && checkLapseRateCodeDev(c.st.code[i], c.cs.useLapseRateCode, true))) return false;
        if (!(z >= double(heightInf) && z <= double(heightSup))) return false;
        v = c.value(i);
        return v != kNoData;
    }, [&](float v, float) { sum += v; nValues += 1.f; });
    if (nValues > 1 || (nValues > 0 && heightSup >= maxPointsZ)) return sum / nValues;
    return kNoData;
}

// findLinesIntersection / findLinesIntersectionAboveThreshold, mathFunctions/basicMath.cpp:138-171
__device__ bool linesIntersectionDev(float q1, float m1, float q2, float m2, float& x, float& y)
{
    if (!isEqualF(m1, m2)) {
        x = (q2 - q1) / (m1 - m2);
        y = m1 * (q2 - q1) / (m1 - m2) + q1;
        return true;
    }
    x = kNoData;
    y = kNoData;
    return false;
}
__device__ bool linesIntersectionAboveDev(float q1, float m1, float q2, float m2, float thr, float& x, float& y)
{
    if (!isEqualF(m1, m2)) {
        x = (q2 - q1) / (m1 - m2);
        y = m1 * (q2 - q1) / (m1 - m2) + q1;
        return x > thr;
    }
    x = kNoData;
    y = kNoData;
    return false;
}

// regressionOrographyT, interpolation.cpp:433-798 (climateExists = true, regressionOrography :1365)
__device__ bool regressionOrographyTDev(const Ctx& c, int pos, ClassicProxyState& p, int& error)
{
    const ClassicSetup& cs = c.cs;
    float ih[kMaxIntervals], iv[kMaxIntervals];
    int nI = 0;
    float m, q, r2, r2_values, r2_intervals, q1, m1, r21, q2, m2, r22, x, y;
    const float DELTAZ_INI = 80.f;
    const bool useCode = cs.useLapseRateCode != 0;
    const float maxHeightInv = cs.maxHeightInversion;
    const float sigR2 = cs.minRegressionR2 > 0.2f ? cs.minRegressionR2 : 0.2f;
    const float sigR2Inv = cs.minRegressionR2 > 0.1f ? cs.minRegressionR2 : 0.1f;
    initializeOrographyDev(p);
    const float climateLapseRate = cs.climateLapseRate;
    PB_SET(p, slope, CF_SLOPE, climateLapseRate);
    const float maxPointsZ = maxHeightDev(c);
    float heightInf = minHeightDev(c);
    if (maxPointsZ == heightInf || cs.n < kMinRegressionPoints) return true;

    float heightSup = heightInf, deltaZ = DELTAZ_INI;
    while (heightSup <= maxPointsZ) {
        float avg = kNoData;
        while (avg == kNoData) {
            heightSup = heightSup + deltaZ;
            avg = heightIntervalAvgDev(c, heightInf, heightSup, maxPointsZ);
        }
        if (nI >= kMaxIntervals) { error = 1; return false; }
        ih[nI] = (heightSup + heightInf) / 2.f;
        iv[nI] = avg;
        nI++;
        // expf through the f64 exponential: correctly rounded to f32, like glibc's expf
        deltaZ = DELTAZ_INI * float(exp(double(heightInf / maxHeightInv)));
        heightInf = heightSup;
    }

    PB_SET(p, T1, CF_T1, iv[0]);
    PB_SET(p, H1, CF_H1, ih[0]);
    for (int i = 1; i < nI; i++)
        if (ih[i] <= maxHeightInv && iv[i] >= p.T1 && (double(iv[i]) > (double(iv[0]) + 0.001 * double(ih[i] - ih[0])))) {
            PB_SET(p, H1, CF_H1, ih[i]);
            PB_SET(p, T1, CF_T1, iv[i]);
            PB_SET(p, invSig, CF_INVSIG, 1);
        }
    if (!p.invSig) return regressionGenericDev(c, pos, p);

    // data below / above the inversion height (vectors myData1/myHeights1, myData2/myHeights2): filtered views
    const float H1split = p.H1;
    auto below = [&](int i, float& xx, float& yy) -> bool {
        const double z = c.st.z[i];
        if (!(z != double(kNoData) && checkLapseRateCodeDev(c.st.code[i], useCode, true))) return false;
        if (!(z <= double(H1split))) return false;
        xx = float(z);
        yy = c.value(i);
        return true;
    };
    auto above = [&](int i, float& xx, float& yy) -> bool {
        const double z = c.st.z[i];
        if (!(z != double(kNoData) && checkLapseRateCodeDev(c.st.code[i], useCode, true))) return false;
        if (z <= double(H1split)) return false;
        xx = float(z);
        yy = c.value(i);
        return true;
    };
    int n2 = 0;
    orderedItems(c.warp, cs.n, above, [&](float, float) { n2++; });
    int nI1 = 0;
    for (int i = 0; i < nI; i++)
        if (ih[i] <= H1split) nI1++;
    const int nI2 = nI - nI1;
    auto int1 = [&](int i, float& xx, float& yy) -> bool {
        if (!(ih[i] <= H1split)) return false;
        xx = ih[i];
        yy = iv[i];
        return true;
    };
    auto int2 = [&](int i, float& xx, float& yy) -> bool {
        if (ih[i] <= H1split) return false;
        xx = ih[i];
        yy = iv[i];
        return true;
    };

    // only positive lapse rate
    if (p.invSig && nI1 == nI) {
        if (!regressionSimpleDev(c, pos, m, q, r2)) return false;
        if (r2 >= sigR2) {
            PB_SET(p, invLapse, CF_INVLAPSE, m);
            PB_SET(p, r2, CF_R2, r2);
            PB_SET(p, T0, CF_T0, q);
            PB_SET(p, T1, CF_T1, q + m * p.H1);
        } else {
            linearRegressionDev(nI, int1, q, m, r2, c.warp);
            PB_SET(p, r2, CF_R2, kNoData);
            if (r2 >= sigR2) {
                PB_SET(p, T0, CF_T0, q);
                PB_SET(p, invLapse, CF_INVLAPSE, m);
                PB_SET(p, T1, CF_T1, q + m * p.H1);
            } else {
                PB_SET(p, invLapse, CF_INVLAPSE, 0.f);
                PB_SET(p, T0, CF_T0, iv[0]);
            }
        }
        return true;
    }

    // check inversion significance
    linearRegressionDev(cs.n, below, q1, m1, r2_values, c.warp);
    if (nI1 > 2) linearRegressionDev(nI, int1, q, m, r2_intervals, c.warp);
    else r2_intervals = 0.f;

    if (r2_values < sigR2Inv && r2_intervals < sigR2Inv) {
        if (!regressionSimpleDev(c, pos, m, q, r2)) return false;
        if (double(r2) >= 0.5) {        // case 0: regression with all data much significant
            PB_SET(p, invSig, CF_INVSIG, 0);
            PB_SET(p, H1, CF_H1, kNoData);
            PB_SET(p, invLapse, CF_INVLAPSE, kNoData);
            PB_SET(p, slope, CF_SLOPE, m < 0.f ? m : 0.f);
            PB_SET(p, r2, CF_R2, r2);
            PB_SET(p, T0, CF_T0, q);
            PB_SET(p, T1, CF_T1, kNoData);
            return true;
        }
        // case 1: analysis only above inversion, flat lapse rate below
        PB_SET(p, invLapse, CF_INVLAPSE, 0.f);
        linearRegressionDev(cs.n, above, q2, m2, r2, c.warp);
        if (n2 >= kMinRegressionPoints) {
            if (r2 >= sigR2) {
                PB_SET(p, slope, CF_SLOPE, m2 < 0.f ? m2 : 0.f);
                PB_SET(p, T0, CF_T0, q2 + p.H1 * p.slope);
                PB_SET(p, T1, CF_T1, p.T0);
                PB_SET(p, r2, CF_R2, r2);
                return true;
            } else {
                linearRegressionDev(nI, int2, q2, m2, r2, c.warp);
                if (r2 >= sigR2) {
                    PB_SET(p, slope, CF_SLOPE, m2 < 0.f ? m2 : 0.f);
                    PB_SET(p, T0, CF_T0, q2 + p.H1 * p.slope);
                    PB_SET(p, T1, CF_T1, p.T0);
                    PB_SET(p, r2, CF_R2, r2);
                    return true;
                }
            }
        }
        PB_SET(p, invSig, CF_INVSIG, 0);
        PB_SET(p, H1, CF_H1, kNoData);
        PB_SET(p, invLapse, CF_INVLAPSE, kNoData);
        PB_SET(p, T0, CF_T0, q);
        PB_SET(p, T1, CF_T1, kNoData);
        // case 2: regression with data
        if (!regressionSimpleDev(c, pos, m, q, r2)) return false;
        if (r2 >= sigR2) {
            PB_SET(p, slope, CF_SLOPE, m < 0.f ? m : 0.f);
            PB_SET(p, T0, CF_T0, q);
            PB_SET(p, r2, CF_R2, r2);
            return true;
        } else {
            PB_SET(p, T0, CF_T0, iv[0]);
            if (double(m) > 0.) PB_SET(p, slope, CF_SLOPE, 0.f);
            else PB_SET(p, slope, CF_SLOPE, climateLapseRate);
            return true;
        }
    }

    // significance analysis
    linearRegressionDev(cs.n, below, q1, m1, r21, c.warp);
    linearRegressionDev(cs.n, above, q2, m2, r22, c.warp);
    if (m1 <= 0) r21 = 0;
    PB_SET(p, r2, CF_R2, r22);

    if (r21 >= sigR2Inv && r22 >= sigR2) {
        if (n2 < kMinRegressionPoints && double(m2) > 0.) { m2 = 0.f; q2 = p.T1; }
        linesIntersectionDev(q1, m1, q2, m2, x, y);
        PB_SET(p, T0, CF_T0, q1);
        PB_SET(p, T1, CF_T1, y);
        PB_SET(p, invLapse, CF_INVLAPSE, m1);
        PB_SET(p, slope, CF_SLOPE, m2);
        PB_SET(p, H1, CF_H1, x);
        if (p.H1 > maxHeightInv) {
            PB_SET(p, T1, CF_T1, p.T1 - (p.H1 - maxHeightInv) * p.slope);
            PB_SET(p, H1, CF_H1, maxHeightInv);
            PB_SET(p, invLapse, CF_INVLAPSE, (p.T1 - p.T0) / (p.H1 - p.H0));
        }
        return true;
    } else if (r21 < sigR2Inv && r22 >= sigR2) {
        if (n2 < kMinRegressionPoints && double(m2) > 0.) { m2 = 0.f; q2 = p.T1; }
        linearRegressionDev(nI, int1, q, m, r2, c.warp);
        PB_SET(p, slope, CF_SLOPE, m2);
        if (r2 >= sigR2Inv) {
            if (linesIntersectionAboveDev(q, m, q2, m2, 40.f, x, y)) {
                PB_SET(p, H1, CF_H1, x);
                PB_SET(p, T0, CF_T0, q);
                PB_SET(p, T1, CF_T1, y);
                PB_SET(p, invLapse, CF_INVLAPSE, m);
                if (p.H1 > maxHeightInv) {
                    PB_SET(p, T1, CF_T1, p.T1 - (p.H1 - maxHeightInv) * p.slope);
                    PB_SET(p, H1, CF_H1, maxHeightInv);
                    PB_SET(p, invLapse, CF_INVLAPSE, (p.T1 - p.T0) / (p.H1 - p.H0));
                }
                return true;
            }
        } else {
            PB_SET(p, invLapse, CF_INVLAPSE, 0.f);
            PB_SET(p, T1, CF_T1, q2 + m2 * p.H1);
            PB_SET(p, T0, CF_T0, p.T1);
            return true;
        }
    } else if (r21 >= sigR2Inv && r22 < sigR2) {
        PB_SET(p, T0, CF_T0, q1);
        PB_SET(p, invLapse, CF_INVLAPSE, m1);
        linearRegressionDev(nI, int2, q, m, r2, c.warp);
        if (r2 >= sigR2) {
            PB_SET(p, slope, CF_SLOPE, m < 0.f ? m : 0.f);
            linesIntersectionDev(p.T0, p.invLapse, q, p.slope, x, y);
            PB_SET(p, H1, CF_H1, x);
            PB_SET(p, T1, CF_T1, y);
        } else {
            PB_SET(p, slope, CF_SLOPE, climateLapseRate);
            linesIntersectionDev(p.T0, p.invLapse, p.T1 - p.slope * p.H1, p.slope, x, y);
            PB_SET(p, H1, CF_H1, x);
            PB_SET(p, T1, CF_T1, y);
        }
        return true;
    } else if (r21 < sigR2Inv && r22 < sigR2) {
        linearRegressionDev(nI, int1, q, m, r2, c.warp);
        if (r2 >= sigR2Inv) {
            PB_SET(p, T0, CF_T0, q);
            PB_SET(p, invLapse, CF_INVLAPSE, m);
            PB_SET(p, T1, CF_T1, q + m * p.H1);
        } else {
            PB_SET(p, invLapse, CF_INVLAPSE, 0.f);
            PB_SET(p, T0, CF_T0, iv[0]);
            PB_SET(p, T1, CF_T1, p.T0);
        }
        linearRegressionDev(nI, int2, q, m, r2, c.warp);
        if (r2 >= sigR2) {
            PB_SET(p, slope, CF_SLOPE, m < 0.f ? m : 0.f);
            if (linesIntersectionAboveDev(p.T0, p.invLapse, q, p.slope, 40.f, x, y)) {
                PB_SET(p, H1, CF_H1, x);
                PB_SET(p, T1, CF_T1, y);
                return true;
            }
        } else {
            PB_SET(p, slope, CF_SLOPE, climateLapseRate);
            return true;
        }
    }

    // check max lapse rate (20 C / 1000 m)
    if (double(p.slope) < -0.02) PB_SET(p, slope, CF_SLOPE, -0.02f);
    initializeOrographyDev(p);
    return regressionGenericDev(c, pos, p);
}

// detrendPoints, interpolation.cpp:1236-1285
__device__ void detrendPointsDev(const Ctx& c, int pos, const ClassicProxyState& p)
{
    const bool isHeight = c.cs.kind[pos] == PB_PROXY_HEIGHT;
    for (int i = c.warp ? int(threadIdx.x & 31) : 0; i < c.cs.n; i += c.warp ? 32 : 1) {      // independent per point
        float detrendValue = 0.f;
        const float proxyValue = c.proxy(i, pos);
        if (isHeight) {
            if (proxyValue != kNoData) {
                const float above = p.slope;
                if (p.invSig) {
                    const float H1 = p.H1, H0 = p.H0, below = p.invLapse;
                    if (proxyValue <= H1) {
                        const float t = proxyValue - H0;
                        detrendValue = (t > 0.f ? t : 0.f) * below;
                    } else detrendValue = ((H1 - H0) * below) + (proxyValue - H1) * above;
                } else detrendValue = (proxyValue > 0.f ? proxyValue : 0.f) * above;
            }
        } else {
            if (proxyValue != kNoData)
                if (p.r2 >= c.cs.minRegressionR2) detrendValue = proxyValue * p.slope;
        }
        c.val[(size_t)i * c.NP] -= detrendValue;
    }
    if (c.warp) __syncwarp();
}

__global__ void __launch_bounds__(64) classic_detrend_kernel(ClassicSetup cs, ClassicStations st, ClassicProblem* __restrict__ problems,
                                                             float* __restrict__ work)
{
    const int pIdx = blockIdx.x * blockDim.x + threadIdx.x;
    if (pIdx >= cs.nProblems) return;
    ClassicProblem& pr = problems[pIdx];
    Ctx c{cs, st, work + pIdx, cs.nProblems, false};
    int error = 0;
    for (int pos = 0; pos < cs.nProxy; pos++) {
        pr.significant[pos] = 0;
        if (!pr.active[pos]) continue;
        ClassicProxyState p = pr.proxy[pos];
        p.written = 0;
        bool ok;
        if (cs.kind[pos] == PB_PROXY_HEIGHT) {
            if (cs.isThermalVar) ok = pr.useThermalInversion ? regressionOrographyTDev(c, pos, p, error) : regressionSimpleTDev(c, pos, p);
            else ok = regressionGenericDev(c, pos, p);
        } else ok = regressionGenericDev(c, pos, p);
        pr.proxy[pos] = p;
        if (ok) {
            pr.significant[pos] = 1;
            detrendPointsDev(c, pos, p);
        }
    }
    pr.error = error;
}

// the same with one WARP per problem (a single time step and its handful of combinations: the GPU is otherwise idle): every
// station loop runs through orderedItems, so the lanes share the loads while each sum keeps the reference's order. Every lane
// holds the same state; lane 0 writes it back.
__global__ void __launch_bounds__(128) classic_detrend_warp_kernel(ClassicSetup cs, ClassicStations st, ClassicProblem* __restrict__ problems,
                                                                   float* __restrict__ work)
{
    const int pIdx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (pIdx >= cs.nProblems) return;
    const bool writer = (threadIdx.x & 31) == 0;
    ClassicProblem& pr = problems[pIdx];
    Ctx c{cs, st, work + pIdx, cs.nProblems, true};
    int error = 0;
    for (int pos = 0; pos < cs.nProxy; pos++) {
        if (writer) pr.significant[pos] = 0;
        if (!pr.active[pos]) continue;
        ClassicProxyState p = pr.proxy[pos];
        p.written = 0;
        bool ok;
        if (cs.kind[pos] == PB_PROXY_HEIGHT) {
            if (cs.isThermalVar) ok = pr.useThermalInversion ? regressionOrographyTDev(c, pos, p, error) : regressionSimpleTDev(c, pos, p);
            else ok = regressionGenericDev(c, pos, p);
        } else ok = regressionGenericDev(c, pos, p);
        __syncwarp();
        if (writer) pr.proxy[pos] = p;
        if (ok) {
            if (writer) pr.significant[pos] = 1;
            detrendPointsDev(c, pos, p);
        }
        __syncwarp();
    }
    if (writer) pr.error = error;
}

// work[i * NP + (t * nCombos + c)] = values[t * n + i]
__global__ void classic_init_kernel(const float* __restrict__ values, int T, int nCombos, int n, float* __restrict__ work)
{
    const long long NP = (long long)T * nCombos;
    const long long total = NP * n;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const long long i = k / NP, p = k - i * NP;
        const int t = int(p / nCombos);
        work[k] = values[(size_t)t * n + i];
    }
}

__global__ void classic_gather_kernel(const float* __restrict__ work, int n, int NP, int problem, float* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = work[(size_t)i * NP + problem];
}

// topo[j * nS + i] = topographicDistance(target j -> station i) as computeDistances :244-247 calls it
__global__ void __launch_bounds__(128) topo_matrix_kernel(const float* __restrict__ tx, const float* __restrict__ ty,
                                                          const float* __restrict__ tz, int nT, const float* __restrict__ sx,
                                                          const float* __restrict__ sy, const float* __restrict__ sz, int nS,
                                                          DevGrid dem, float* __restrict__ topo,
                                                          const DevGrid* __restrict__ tdMaps)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (long long)nT * nS) return;
    const int j = int(k / nS), i = int(k - (long long)j * nS);
    const float dx = sx[i] - tx[j], dy = sy[i] - ty[j];
    const float d = sqrtf(dx * dx + dy * dy);
    topo[k] = topographicDistanceCachedDev(tdMaps, i, dem, tx[j], ty[j], tz[j], sx[i], sy[i], sz[i], d);
}

// retrend, classic branch (interpolation.cpp:1316-1347) with the problem's fitted proxies
__device__ float retrendClassicDev(const LooSetup& ls, const ClassicProblem& pr, const float* tproxy)
{
    if (!ls.useDetrendingVar) return 0.f;
    double retrend = 0.;
    for (int pos = 0; pos < ls.nProxy; pos++) {
        if (!(pr.active[pos] && pr.significant[pos])) continue;
        const double v = double(tproxy[pos]);      // getProxyValues(): vector<double> of the float proxy values
        if (v == double(kNoData)) continue;
        const ClassicProxyState& p = pr.proxy[pos];
        const float slope = p.slope;
        if (ls.kind[pos] == PB_PROXY_HEIGHT) {
            if (ls.useThermalInversion && p.invSig) {
                const float H0 = p.H0, H1 = p.H1, below = p.invLapse;
                if (v <= double(H1)) {
                    double t = v - double(H0);
                    t = t > 0. ? t : 0.;
                    retrend += t * double(below);
                } else {
                    const float a = (H1 - H0) * below;
                    retrend += double(a) + (v - double(H1)) * double(slope);
                }
            } else {
                const double t = v > 0. ? v : 0.;
                retrend += t * double(slope);
            }
        } else retrend += v * double(slope);
    }
    return float(retrend);
}

// exact leave-one-out. WARP = (kh index, target j); problems in groups of kLooProblemsPerThread (blockIdx.y).
// The weights of the stations are independent of each other, only the sums are ordered: the 32 lanes compute the weights
// (f32 distance, the two IEEE f64 divisions) of 32 consecutive stations at once, then every sum takes them in station
// order through shuffles (a station at distance 0 and the lanes past the last station contribute the weight +0, which
// leaves a sum as it is). One thread per target walked the stations alone: ~300 dependent cycles per station, 0.15 ms for
// 500 x 500 pairs however few targets there were. Same terms, same order, same bits.
// residual layout: [(problem * nKh + k) * nT + j]
__global__ void __launch_bounds__(128)
loo_exact_kernel(LooSetup ls, LooTargets tg, const float* __restrict__ sx, const float* __restrict__ sy,
                 const float* __restrict__ work, const ClassicProblem* __restrict__ problems, const float* __restrict__ topo,
                 const int* __restrict__ khList, float* __restrict__ residual, float* __restrict__ estimate)
{
    const int nS = ls.nS, nT = ls.nT, NP = ls.nProblems;
    const int lane = threadIdx.x & 31;
    const long long tIdx = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (tIdx >= (long long)nT * ls.nKh) return;
    const int k = int(tIdx / nT), j = int(tIdx - (long long)k * nT);
    const int p0 = blockIdx.y * kLooProblemsPerThread;
    const int np = min(kLooProblemsPerThread, NP - p0);
    const int kh = khList ? khList[k] : 0;

    bool valid = true;
    if (ls.excludeSupplemental) valid = checkLapseRateCodeDev(tg.code[j], ls.useLapseRateCode, false);
    if (valid && ls.excludeOutsideDem && tg.insideDem) valid = tg.insideDem[j] != 0;
    if (!valid) {
        if (lane == 0)
            for (int q = 0; q < np; q++) {
                residual[((size_t)(p0 + q) * ls.nKh + k) * nT + j] = kNoData;
                if (estimate) estimate[((size_t)(p0 + q) * ls.nKh + k) * nT + j] = kNoData;
            }
        return;
    }

    const float x = tg.x[j], y = tg.y[j];
    double sumW = 0.;
    double sum[kLooProblemsPerThread];
#pragma unroll
    for (int q = 0; q < kLooProblemsPerThread; q++) sum[q] = 0.;
    if (!ls.useRetrendOnly) {
        const bool useTopo = topo && kh != 0;
        const float fkh = float(kh);
        for (int base = 0; base < nS; base += 32) {
            const int i = base + lane;
            double w = 0.;
            double pq[kLooProblemsPerThread];
#pragma unroll
            for (int q = 0; q < kLooProblemsPerThread; q++) pq[q] = 0.;
            if (i < nS) {
                // computeDistances(..., excludeSupplemental = false) :208-256, inverseDistanceWeighted :1031-1051
                const float dx = sx[i] - x, dy = sy[i] - y;
                float d = sqrtf(dx * dx + dy * dy);
                if (useTopo) d += (fkh * topo[(size_t)j * nS + i]);
                if (d > 0.f) {
                    const double dk = double(d) / 10000.;
                    w = 1.0 / (dk * dk * dk);
                    const float* v = work + (size_t)i * NP + p0;
#pragma unroll
                    for (int q = 0; q < kLooProblemsPerThread; q++)
                        if (q < np) pq[q] = double(v[q]) * w;
                }
            }
            // station order: lane t of this chunk is station base + t
#pragma unroll 8
            for (int t = 0; t < 32; t++) {
                sumW += __shfl_sync(0xffffffffu, w, t);
#pragma unroll
                for (int q = 0; q < kLooProblemsPerThread; q++)
                    if (q < np) sum[q] += __shfl_sync(0xffffffffu, pq[q], t);
            }
        }
    }
    if (lane != 0) return;
    const float* tproxy = tg.proxy + (size_t)j * (ls.nProxy > 0 ? ls.nProxy : 1);
    float myValue = tg.observed[j];
    if (ls.isPrecVar && myValue != kNoData && myValue < ls.rainfallThreshold) myValue = 0.f;
    for (int q = 0; q < np; q++) {
        float result;
        if (ls.useRetrendOnly) result = 0.f;
        else result = sumW > 0. ? float(sum[q] / sumW) : kNoData;
        float est = kNoData;
        if (!isEqualF(result, kNoData)) {
            if (!ls.useDoNotRetrend) {
                if (ls.useMultipleModel) {
                    double pv[PB_MAX_PROXY];
                    for (int i = 0; i < ls.nProxy; i++) pv[i] = double(tproxy[i]);
                    result += retrendValue(ls.multiple, pv);
                } else result += retrendClassicDev(ls, problems[p0 + q], tproxy);
            }
            switch (ls.clampMode) {
            case CLAMP_PREC: est = (result < ls.rainfallThreshold) ? 0.f : result; break;
            case CLAMP_RH: est = fmaxf(0.f, fminf(result, 100.f)); break;
            case CLAMP_POS: est = fmaxf(result, 0.f); break;
            default: est = result;
            }
        }
        if (ls.isPrecVar && est != kNoData && est < ls.rainfallThreshold) est = 0.f;
        float res = kNoData;
        if (est != kNoData && myValue != kNoData) res = myValue - est;
        residual[((size_t)(p0 + q) * ls.nKh + k) * nT + j] = res;
        if (estimate) estimate[((size_t)(p0 + q) * ls.nKh + k) * nT + j] = est;
    }
}

// neighbourhoodVariability, interpolation.cpp:259-301: computeDistances(..., excludeSupplemental = true), the maxNrPoints
// nearest points with |d| >= EPSILON (sortPointsByDistance :121-160; equal distances in station order), then
// statistics::standardDeviation (float accumulators, statistics.cpp:1189-1215, 1383-1391) and statistics::mean of |dz|
constexpr int kMaxNeighbours = 16;
// WARP = target. Lane l keeps the nearest stations among l, l + 32, ... as a sorted list in REGISTERS (every index a
// compile-time constant; a list indexed by a run-time position lives in local memory), then the 32 lists are merged: in every
// round the lanes offer their heads and the warp takes the smallest (distance, station index) pair — the order of
// sortPointsByDistance (ascending, ties in station order), and exactly the maxNrPoints entries the sequential insertion keeps.
__global__ void __launch_bounds__(128)
neighbourhood_kernel(const float* __restrict__ tx, const float* __restrict__ ty, const float* __restrict__ tz, int nT,
                     const float* __restrict__ sx, const float* __restrict__ sy, const double* __restrict__ sz,
                     const int* __restrict__ scode, const float* __restrict__ svalue, int nS, int useLapseRateCode,
                     const float* __restrict__ topo, int kh, int maxNrPoints, NeighbourOut* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= nT) return;
    const float x = tx[j], y = ty[j], z = tz[j];
    float ld[kMaxNeighbours];
    int li[kMaxNeighbours];
#pragma unroll
    for (int q = 0; q < kMaxNeighbours; q++) { ld[q] = INFINITY; li[q] = 0x7fffffff; }
    const int cap = maxNrPoints < kMaxNeighbours ? maxNrPoints : kMaxNeighbours;
    const bool useTopo = topo && kh != 0;
    float last = INFINITY;                 // ld[cap - 1]
    for (int i = lane; i < nS; i += 32) {
        float d;
        if (!checkLapseRateCodeDev(scode[i], useLapseRateCode, false)) d = 0.f;
        else {
            const float dx = sx[i] - x, dy = sy[i] - y;
            d = sqrtf(dx * dx + dy * dy);
            if (useTopo) d += (float(kh) * topo[(size_t)j * nS + i]);
        }
        if (isEqualF(d, 0.f) || isEqualF(d, kNoData)) continue;
        if (!(d < last)) continue;
        int pos = 0;
#pragma unroll
        for (int q = 0; q < kMaxNeighbours; q++) pos += (q < cap && ld[q] <= d) ? 1 : 0;
#pragma unroll
        for (int q = kMaxNeighbours - 1; q >= 1; q--)
            if (q < cap && q > pos) { ld[q] = ld[q - 1]; li[q] = li[q - 1]; }
#pragma unroll
        for (int q = 0; q < kMaxNeighbours; q++)
            if (q == pos) { ld[q] = d; li[q] = i; }
#pragma unroll
        for (int q = 0; q < kMaxNeighbours; q++)
            if (q == cap - 1) last = ld[q];
    }
    // merge: nd / ni = the warp's list (the same in every lane)
    float nd[kMaxNeighbours];
    int ni[kMaxNeighbours];
    int n = 0;
#pragma unroll
    for (int r = 0; r < kMaxNeighbours; r++) {
        nd[r] = INFINITY;
        ni[r] = 0;
        if (r < cap) {
            float bd = ld[0];
            int bi = li[0];
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, bd, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
            }
            if (bd < INFINITY) {
                nd[r] = bd;
                ni[r] = bi;
                n = r + 1;
                if (li[0] == bi) {          // this lane's head was taken: pop it
#pragma unroll
                    for (int q = 0; q < kMaxNeighbours - 1; q++) { ld[q] = ld[q + 1]; li[q] = li[q + 1]; }
                    ld[kMaxNeighbours - 1] = INFINITY;
                    li[kMaxNeighbours - 1] = 0x7fffffff;
                }
            }
        }
    }
    if (lane != 0) return;
    NeighbourOut o;
    o.ok = n > 1;
    o.devSt = o.avgDeltaZ = o.minDistance = kNoData;
    if (n > 1) {
        o.minDistance = nd[0];
        // mean (double sum) then variance (float accumulators)
        double sum = 0.;
        int nv = 0;
#pragma unroll
        for (int q = 0; q < kMaxNeighbours; q++) {
            if (q < n) {
                const float v = svalue[ni[q]];
                if (!isEqualF(v, kNoData)) { sum += double(v); nv++; }
            }
        }
        const float myMean = nv ? float(sum / double(nv)) : kNoData;
        float squareDiff = 0.f;
        int nrValid = 0;
#pragma unroll
        for (int q = 0; q < kMaxNeighbours; q++) {
            if (q < n) {
                const float v = svalue[ni[q]];
                if (v != kNoData) {
                    const float diff = v - myMean;
                    squareDiff += diff * diff;
                    nrValid++;
                }
            }
        }
        const float variance = nrValid > 1 ? squareDiff / float(nrValid - 1) : kNoData;
        o.devSt = isEqualF(variance, kNoData) ? kNoData : sqrtf(variance);
        if (z != kNoData) {
            double s2 = 0.;
            int nz = 0;
#pragma unroll
            for (int q = 0; q < kMaxNeighbours; q++) {
                if (q < n) {
                    const double zi = sz[ni[q]];
                    if (!(fabs(zi - double(kNoData)) < kEpsilon)) {
                        const float dz = float(fabs(zi - double(z)));
                        if (!isEqualF(dz, kNoData)) { s2 += double(dz); nz++; }
                    }
                }
            }
            o.avgDeltaZ = nz ? float(s2 / double(nz)) : kNoData;
        }
    }
    out[j] = o;
}

// computeErrorCrossValidation (spatialControl.cpp:305-328) + statistics::meanAbsoluteError (statistics.cpp:201-220):
// one thread per job walks its residual row in station order
__global__ void cv_mae_kernel(const float* __restrict__ residual, const float* __restrict__ observed, int n, int nJobs,
                              float* __restrict__ mae)
{
    const int job = blockIdx.x * blockDim.x + threadIdx.x;
    if (job >= nJobs) return;
    const float* r = residual + (size_t)job * n;
    double sum = 0.;
    long cnt = 0;
    for (int j = 0; j < n; j++) {
        const float value = observed[j], res = r[j];
        if (value != kNoData && res != kNoData) {
            const float est = value - res;
            if (est != kNoData) {          // meanAbsoluteError skips NODATA pairs
                sum += fabs(double(value - est));
                cnt++;
            }
        }
    }
    mae[job] = cnt == 0 ? kNoData : float(sum / double(cnt));
}

}  // namespace

cudaError_t launchClassicInit(const float* values, int T, int nCombos, int n, float* work, cudaStream_t s)
{
    if ((long long)T * nCombos * n <= 0) return cudaSuccess;
    classic_init_kernel<<<148 * 4, 256, 0, s>>>(values, T, nCombos, n, work);
    return cudaGetLastError();
}

cudaError_t launchClassicDetrend(const ClassicSetup& cs, const ClassicStations& st, ClassicProblem* problems, float* work,
                                 cudaStream_t s)
{
    if (cs.nProblems <= 0) return cudaSuccess;
    static const bool noWarp = getenv("PB_CLASSIC_NO_WARP") != nullptr;
    if (cs.nProblems <= 64 && !noWarp)
        classic_detrend_warp_kernel<<<(cs.nProblems * 32 + 127) / 128, 128, 0, s>>>(cs, st, problems, work);
    else
        classic_detrend_kernel<<<(cs.nProblems + 63) / 64, 64, 0, s>>>(cs, st, problems, work);
    return cudaGetLastError();
}

cudaError_t launchClassicGather(const float* work, int n, int NP, int problem, float* out, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    classic_gather_kernel<<<(n + 255) / 256, 256, 0, s>>>(work, n, NP, problem, out);
    return cudaGetLastError();
}

cudaError_t launchTopoMatrix(const float* tx, const float* ty, const float* tz, int nT, const float* sx, const float* sy,
                             const float* sz, int nS, const DevGrid& dem, float* topo, cudaStream_t s, const DevGrid* tdMaps)
{
    if (nT <= 0 || nS <= 0) return cudaSuccess;
    const long long total = (long long)nT * nS;
    topo_matrix_kernel<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(tx, ty, tz, nT, sx, sy, sz, nS, dem, topo, tdMaps);
    return cudaGetLastError();
}

cudaError_t launchNeighbourhood(const float* tx, const float* ty, const float* tz, int nT, const float* sx, const float* sy,
                                const double* sz, const int* scode, const float* svalue, int nS, int useLapseRateCode,
                                const float* topo, int kh, int maxNrPoints, NeighbourOut* out, cudaStream_t s)
{
    if (nT <= 0) return cudaSuccess;
    if (maxNrPoints > kMaxNeighbours) maxNrPoints = kMaxNeighbours;
    neighbourhood_kernel<<<(nT * 32 + 127) / 128, 128, 0, s>>>(tx, ty, tz, nT, sx, sy, sz, scode, svalue, nS, useLapseRateCode, topo, kh,
                                                        maxNrPoints, out);
    return cudaGetLastError();
}

cudaError_t launchLooExact(const LooSetup& ls, const LooTargets& t, const float* sx, const float* sy, const float* work,
                           const ClassicProblem* problems, const float* topo, const int* khList, float* residual, cudaStream_t s,
                           float* estimate)
{
    if (ls.nT <= 0 || ls.nProblems <= 0 || ls.nKh <= 0) return cudaSuccess;
    const long long threads = (long long)ls.nT * ls.nKh * 32;          // a warp per (kh, target)
    dim3 grid((unsigned)((threads + 127) / 128), (unsigned)((ls.nProblems + kLooProblemsPerThread - 1) / kLooProblemsPerThread));
    loo_exact_kernel<<<grid, 128, 0, s>>>(ls, t, sx, sy, work, problems, topo, khList, residual, estimate);
    return cudaGetLastError();
}

cudaError_t launchCvMae(const float* residual, const float* observed, int n, int nJobs, float* mae, cudaStream_t s)
{
    if (nJobs <= 0) return cudaSuccess;
    cv_mae_kernel<<<(nJobs + 63) / 64, 64, 0, s>>>(residual, observed, n, nJobs, mae);
    return cudaGetLastError();
}

}  // namespace pb
