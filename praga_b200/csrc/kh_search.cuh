// kh_search.cuh — host pieces of topographicDistanceOptimize (interpolation.cpp:2356-2368) shared by the station-space chain
// (station_space.cu) and the per-area search of glocal detrending (glocal.cu, macroAreaDetrending :1811-1826)
#pragma once
#include <math.h>
#include <algorithm>

#include "common.cuh"

namespace pb {

// computeErrorCrossValidation (spatialControl.cpp:305-328) + statistics::meanAbsoluteError (statistics.cpp:201-220) on the host
inline float cvMaeHost(const float* residual, const float* observed, int n)
{
    double sum = 0.;
    long cnt = 0;
    for (int j = 0; j < n; j++) {
        const float value = observed[j], res = residual[j];
        if (value != kNoData && res != kNoData) {
            const float e = value - res;
            if (e != kNoData) { sum += fabs(double(value - e)); cnt++; }
        }
    }
    return cnt == 0 ? kNoData : float(sum / double(cnt));
}

// goldenSectionSearch (interpolation.cpp:2297-2335) over f(kh) = MAE at kh = int(khFloat), read from the table
template <class MaeOfKh>
inline double goldenSection(MaeOfKh maeOfKh, int maxKh, double a, double b, pb_kh_series* series)
{
    const double GOLDEN = 1.6180339887499;      // GOLDEN_SECTION, commonConstants.h:258
    auto f = [&](double khFloat) -> double {
        int kh = int(khFloat);
        kh = std::max(0, std::min(kh, maxKh));
        return double(maeOfKh(kh));
    };
    auto add = [&](float kh, float err) {
        if (series && series->n < PB_MAX_KH_SERIES) { series->kh[series->n] = kh; series->error[series->n] = err; series->n++; }
    };
    const double tol = 1;
    double x1 = b - (b - a) / GOLDEN;
    double x2 = a + (b - a) / GOLDEN;
    int counter = 0;
    while (fabs(b - a) > tol && counter < 100) {
        counter++;
        if (f(x1) < f(x2)) {
            b = x2;
            x2 = x1;
            x1 = b - (b - a) / GOLDEN;
            add(float(x1), float(f(x1)));
        } else {
            a = x1;
            x1 = x2;
            x2 = a + (b - a) / GOLDEN;
            add(float(x2), float(f(x2)));
        }
    }
    add(float((a + b) / 2), float(f((a + b) / 2)));
    return (a + b) / 2;
}

}  // namespace pb
