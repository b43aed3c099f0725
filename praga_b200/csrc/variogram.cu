// variogram.cu — the host-side variogram estimation that precedes ordinary kriging (BASELINE config 4: "ordinary kriging
// with fitted variogram"; north_star: "the variogram fit stays on the host").
//
// The reference DECLARES krigingEstimateVariogram(myDist, mySemiVar, sizeMyVar, nrMyPoints, myMaxDistance, mySill, myNugget,
// myRange, mySlope, myMode, nrPointData) (agrolib/interpolation/interpolation.h:65-68) and krigLinearPrep (:69) but defines
// neither: there is no reference arithmetic to be identical to, so this is the engine's own definition
// and its parity is UNPINNED by construction; tests/test_variogram.py checks it against known models instead.
// What IS fixed by the reference is the model family the fitted parameters feed (krigingVariogram, kriging.cpp:97-202):
//   spherical   g(h) = c0 + (c - c0) (1.5 h/a - 0.5 (h/a)^3)  for h < a, else c
//   exponential g(h) = c0 + (c - c0) (1 - exp(-3 h / a))
//   gaussian    g(h) = c0 + (c - c0) (1 - exp(-4 (h/a)^2))        
//   linear      g(h) = c0 + slope h
// Step 1 (pb_kriging_empirical_variogram): classical Matheron estimator over all station pairs closer than maxDistance,
//   binned by distance: semiVar_b = sum (v_i - v_j)^2 / (2 N_b), dist_b = mean pair distance of the bin. O(n^2) pairs on the
//   host cores (n = 2000: 2e6 pairs, ~ms).
// Step 2 (pb_kriging_fit_variogram): weighted least squares on the bins (weights N_b / g_model^2 would need iteration; the
//   classical N_b weights are used): for a given range the bounded models are LINEAR in (c0, c - c0), so the range is
//   scanned on a logarithmic grid and refined by golden section, each candidate solved in closed form under c0 >= 0,
//   c - c0 >= 0. mode = 0 tries the three bounded models and keeps the smallest weighted SSE. The nugget is floored at
//   1e-4 of the sill: kriging.cpp divides by the system's diagonal and needs nugget > 0.
#include <math.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#include "context.cuh"
#include "nvtx.cuh"

namespace {

double shapeOf(int mode, double h, double a)
{
    if (!(a > 0)) return 1.0;
    const double r = h / a;
    switch (mode) {
    case PB_KRIGING_SPHERICAL: return h < a ? 1.5 * r - 0.5 * r * r * r : 1.0;
    case PB_KRIGING_EXPONENTIAL: return 1.0 - exp(-3.0 * r);
    case PB_KRIGING_GAUSSIAN: return 1.0 - exp(-4.0 * r * r);
    default: return h;      // linear: "shape" is the distance itself, the second coefficient is the slope
    }
}

// weighted LSQ of g = c0 + b f over the bins, constrained to c0 >= 0, b >= 0; returns the weighted SSE
double solveLinear(const std::vector<double>& f, const std::vector<double>& g, const std::vector<double>& w, double& c0, double& b)
{
    double sw = 0, sf = 0, sg = 0, sff = 0, sfg = 0;
    for (size_t i = 0; i < f.size(); i++) {
        sw += w[i]; sf += w[i] * f[i]; sg += w[i] * g[i]; sff += w[i] * f[i] * f[i]; sfg += w[i] * f[i] * g[i];
    }
    const double det = sw * sff - sf * sf;
    if (fabs(det) > 1e-300 * std::max(1.0, sw * sff)) {
        b = (sw * sfg - sf * sg) / det;
        c0 = (sg - b * sf) / sw;
    } else { b = 0; c0 = sw > 0 ? sg / sw : 0; }
    if (c0 < 0) { c0 = 0; b = sff > 0 ? sfg / sff : 0; }          // on the boundary: one-parameter fits
    if (b < 0) { b = 0; c0 = sw > 0 ? std::max(0.0, sg / sw) : 0; }
    double sse = 0;
    for (size_t i = 0; i < f.size(); i++) {
        const double e = g[i] - (c0 + b * f[i]);
        sse += w[i] * e * e;
    }
    return sse;
}

struct Fit { int mode; double range, nugget, sill, slope, sse; };

Fit fitBounded(int mode, const std::vector<double>& h, const std::vector<double>& g, const std::vector<double>& w, double hMin,
               double hMax)
{
    std::vector<double> f(h.size());
    auto sseAt = [&](double a, double& c0, double& b) {
        for (size_t i = 0; i < h.size(); i++) f[i] = shapeOf(mode, h[i], a);
        return solveLinear(f, g, w, c0, b);
    };
    const int nScan = 96;
    const double lo = std::max(hMin * 0.5, hMax * 1e-4), hi = hMax * 3.0;
    double bestA = lo, bestS = INFINITY, c0, b;
    int bestK = 0;
    for (int k = 0; k < nScan; k++) {
        const double a = lo * pow(hi / lo, double(k) / (nScan - 1));
        const double s = sseAt(a, c0, b);
        if (s < bestS) { bestS = s; bestA = a; bestK = k; }
    }
    // golden section on log(a) between the neighbours of the best grid point
    double x0 = log(lo * pow(hi / lo, double(std::max(bestK - 1, 0)) / (nScan - 1)));
    double x3 = log(lo * pow(hi / lo, double(std::min(bestK + 1, nScan - 1)) / (nScan - 1)));
    const double gr = 0.6180339887498949;
    double x1 = x3 - gr * (x3 - x0), x2 = x0 + gr * (x3 - x0);
    double s1 = sseAt(exp(x1), c0, b), s2 = sseAt(exp(x2), c0, b);
    for (int it = 0; it < 60 && (x3 - x0) > 1e-10; it++) {
        if (s1 < s2) { x3 = x2; x2 = x1; s2 = s1; x1 = x3 - gr * (x3 - x0); s1 = sseAt(exp(x1), c0, b); }
        else { x0 = x1; x1 = x2; s1 = s2; x2 = x0 + gr * (x3 - x0); s2 = sseAt(exp(x2), c0, b); }
    }
    const double aRef = exp(0.5 * (x0 + x3));
    const double sRef = sseAt(aRef, c0, b);
    if (sRef <= bestS) { bestS = sRef; bestA = aRef; } else sseAt(bestA, c0, b);
    Fit r;
    r.mode = mode; r.range = bestA; r.nugget = c0; r.sill = c0 + b; r.slope = 0; r.sse = bestS;
    return r;
}

}  // namespace

extern "C" {

int pb_kriging_empirical_variogram(const double* pos, const double* val, int n, int nBins, double maxDistance, float* binDist,
                                   float* binSemiVar, int32_t* binCount, double* maxDistanceUsed)
{
    PB_NVTX_RANGE();
    if (!pos || !val || n < 2 || nBins < 1 || !binDist || !binSemiVar || !binCount) return PB_ERR_INVALID;
    if (!(maxDistance > 0)) {       // default: half the diagonal of the stations' bounding box
        double x0 = pos[0], x1 = x0, y0 = pos[1], y1 = y0;
        for (int i = 1; i < n; i++) {
            x0 = std::min(x0, pos[2 * i]); x1 = std::max(x1, pos[2 * i]);
            y0 = std::min(y0, pos[2 * i + 1]); y1 = std::max(y1, pos[2 * i + 1]);
        }
        maxDistance = 0.5 * sqrt((x1 - x0) * (x1 - x0) + (y1 - y0) * (y1 - y0));
        if (!(maxDistance > 0)) return PB_ERR_INVALID;
    }
    if (maxDistanceUsed) *maxDistanceUsed = maxDistance;
    const double inv = nBins / maxDistance;
    std::vector<double> sumD(nBins, 0.0), sumG(nBins, 0.0);
    std::vector<long long> cnt(nBins, 0);
#pragma omp parallel
    {
        std::vector<double> d(nBins, 0.0), g(nBins, 0.0);
        std::vector<long long> c(nBins, 0);
#pragma omp for schedule(dynamic, 16)
        for (int i = 0; i < n - 1; i++) {
            const double xi = pos[2 * i], yi = pos[2 * i + 1], vi = val[i];
            for (int j = i + 1; j < n; j++) {
                const double dx = pos[2 * j] - xi, dy = pos[2 * j + 1] - yi;
                const double h = sqrt(dx * dx + dy * dy);
                if (!(h < maxDistance)) continue;
                const int b = std::min(nBins - 1, int(h * inv));
                const double dv = val[j] - vi;
                d[b] += h; g[b] += dv * dv; c[b]++;
            }
        }
#pragma omp critical
        for (int b = 0; b < nBins; b++) { sumD[b] += d[b]; sumG[b] += g[b]; cnt[b] += c[b]; }
    }
    for (int b = 0; b < nBins; b++) {
        binCount[b] = int32_t(std::min<long long>(cnt[b], 0x7fffffff));
        binDist[b] = cnt[b] ? float(sumD[b] / cnt[b]) : float((b + 0.5) / inv);
        binSemiVar[b] = cnt[b] ? float(sumG[b] / (2.0 * cnt[b])) : float(PB_NODATA);
    }
    return PB_OK;
}

int pb_kriging_fit_variogram(const float* binDist, const float* binSemiVar, const int32_t* binCount, int nBins, int mode,
                             double* range, double* nugget, double* sill, double* slope, int* modeOut, double* weightedSse)
{
    PB_NVTX_RANGE();
    if (!binDist || !binSemiVar || nBins < 2 || !range || !nugget || !sill || !slope) return PB_ERR_INVALID;
    if (mode != 0 && mode != PB_KRIGING_SPHERICAL && mode != PB_KRIGING_EXPONENTIAL && mode != PB_KRIGING_GAUSSIAN &&
        mode != PB_KRIGING_LINEAR)
        return PB_ERR_INVALID;
    std::vector<double> h, g, w;
    for (int b = 0; b < nBins; b++) {
        const long long c = binCount ? binCount[b] : 1;
        if (c <= 0 || binSemiVar[b] == float(PB_NODATA) || !(binDist[b] > 0)) continue;
        h.push_back(binDist[b]); g.push_back(binSemiVar[b]); w.push_back(double(c));
    }
    if (h.size() < 3) return PB_ERR_FIT;
    const double hMin = *std::min_element(h.begin(), h.end()), hMax = *std::max_element(h.begin(), h.end());
    Fit best{};
    best.sse = INFINITY;
    if (mode == PB_KRIGING_LINEAR) {
        double c0, b;
        best.sse = solveLinear(h, g, w, c0, b);
        best.mode = PB_KRIGING_LINEAR; best.nugget = c0; best.slope = b; best.range = hMax; best.sill = c0 + b * hMax;
    } else {
        const int modes[3] = {PB_KRIGING_SPHERICAL, PB_KRIGING_EXPONENTIAL, PB_KRIGING_GAUSSIAN};
        for (int m : modes) {
            if (mode != 0 && mode != m) continue;
            const Fit f = fitBounded(m, h, g, w, hMin, hMax);
            if (f.sse < best.sse) best = f;
        }
    }
    if (!(best.sse < INFINITY) || !(best.sill > 0)) return PB_ERR_FIT;
    best.nugget = std::max(best.nugget, 1e-4 * best.sill);       // kriging.cpp needs nugget > 0
    *range = best.range; *nugget = best.nugget; *sill = best.sill; *slope = best.slope;
    if (modeOut) *modeOut = best.mode;
    if (weightedSse) *weightedSse = best.sse;
    return PB_OK;
}

}  // extern "C"
