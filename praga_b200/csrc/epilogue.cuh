// epilogue.cuh — per-target tail of interpolate() shared by the raster/point estimators (K1, TD variant):
// getValueFromXY (gis/gis.cpp:533-546), retrend (interpolation/interpolation.cpp:1288-1351) and the per-variable clamp
// (:2532-2558). Every floating-point operation is an explicit round-to-nearest intrinsic, so the results do not depend on
// the FMA-contraction setting of the including translation unit.
#pragma once
#include "common.cuh"

namespace pb {

// ---------------------------------------------------------------------------------------------------------
// per-target epilogue pieces
// ---------------------------------------------------------------------------------------------------------

// Crit3DRasterGrid::getValueFromXY (gis.cpp:533-546): nearest cell by truncation, flag when outside
__device__ __forceinline__ float gridValueFromXY(const DevGrid& g, double x, double y)
{
    const int r = int(__dmul_rn(__dsub_rn(y, g.lly), g.invCellSize));
    const int row = (g.nrRows - 1) - r;
    const int col = int(__dmul_rn(__dsub_rn(x, g.llx), g.invCellSize));
    if (unsigned(row) >= unsigned(g.nrRows) || unsigned(col) >= unsigned(g.nrCols)) return g.flag;
    return __ldg(g.data + (size_t)row * g.nrCols + col);
}

// lapseRatePiecewise_two / _three / _three_free / functionLinear_intercept
// (mathFunctions/furtherMathFunctions.cpp:115-201); evaluated without FMA contraction like the x86-64 reference build
__device__ __forceinline__ double evalFitFunction(int f, double x, const double* par)
{
    switch (f) {
    case PB_FIT_PIECEWISE_TWO: {
        const double slope = (x < par[0]) ? par[2] : par[3];
        return __dadd_rn(__dmul_rn(slope, __dsub_rn(x, par[0])), par[1]);
    }
    case PB_FIT_PIECEWISE_THREE: {
        const double p2 = par[2] > 10. ? par[2] : 10.;   // MAXVALUE(10, par[2])
        const double xb = __dadd_rn(p2, par[0]);
        if (x < par[0]) return __dadd_rn(__dsub_rn(__dmul_rn(par[4], x), __dmul_rn(par[0], par[4])), par[1]);
        if (x > xb)
            return __dadd_rn(__dadd_rn(__dsub_rn(__dsub_rn(__dmul_rn(par[4], x), __dmul_rn(par[4], par[0])),
                                                 __dmul_rn(par[4], p2)),
                                       __dmul_rn(par[3], p2)),
                             par[1]);
        return __dadd_rn(__dsub_rn(__dmul_rn(par[3], x), __dmul_rn(par[3], par[0])), par[1]);
    }
    case PB_FIT_PIECEWISE_THREE_FREE: {
        const double p2 = par[2] > 10. ? par[2] : 10.;
        const double xb = __dadd_rn(par[0], p2);
        if (x < par[0]) return __dadd_rn(__dsub_rn(__dmul_rn(par[4], x), __dmul_rn(par[0], par[4])), par[1]);
        if (x > xb)
            return __dadd_rn(__dadd_rn(__dsub_rn(__dsub_rn(__dmul_rn(par[5], x), __dmul_rn(par[5], par[0])),
                                                 __dmul_rn(par[5], p2)),
                                       __dmul_rn(par[3], p2)),
                             par[1]);
        return __dadd_rn(__dsub_rn(__dmul_rn(par[3], x), __dmul_rn(par[3], par[0])), par[1]);
    }
    case PB_FIT_LINEAR:
        return __dadd_rn(__dmul_rn(par[0], x), par[1]);
    default:
        return 0.;
    }
}

// retrend (interpolation.cpp:1288-1351); proxyValues[i] as filled by getProxyValuesXY (NODATA when missing)
__device__ __forceinline__ float retrendValue(const DevModel& m, const double* proxyValues)
{
    if (!m.useDetrendingVar) return 0.f;
    double retrend = 0.;
    if (m.useMultipleDetrending) {
        // getMultipleDetrendingValues: elevation first, then the others; ANY active&&significant NODATA -> 0
        if (m.nActiveSig == 0 || m.nFit == 0) return 0.f;
        for (int k = 0; k < m.nActiveSig; k++)
            if (proxyValues[m.activeSigPos[k]] == double(kNoData)) return 0.f;
        double sum = 0.;
        for (int k = 0; k < m.nFit; k++)
            sum = __dadd_rn(sum, evalFitFunction(m.fitFunction[k], proxyValues[m.activeSigPos[k]], m.fitParams[k]));
        retrend = double(float(sum));
    } else {
        for (int pos = 0; pos < m.nProxy; pos++) {
            const DevProxy& p = m.proxy[pos];
            if (!(p.active && p.significant)) continue;
            const double v = proxyValues[pos];
            if (v == double(kNoData)) continue;
            const float slope = p.slope;
            if (p.kind == PB_PROXY_HEIGHT) {
                if (m.useThermalInversion && p.inversionIsSignificative) {
                    const float H0 = p.H0, H1 = p.H1, below = p.invLapseRate;
                    if (v <= double(H1)) {
                        double t = __dsub_rn(v, double(H0));
                        t = t > 0. ? t : 0.;
                        retrend = __dadd_rn(retrend, __dmul_rn(t, double(below)));
                    } else {
                        // ((LR_H1 - LR_H0) * LR_Below) is float arithmetic in the reference, the rest double
                        const float a = __fmul_rn(__fsub_rn(H1, H0), below);
                        const double b = __dmul_rn(__dsub_rn(v, double(H1)), double(slope));
                        retrend = __dadd_rn(retrend, __dadd_rn(double(a), b));
                    }
                } else {
                    const double t = v > 0. ? v : 0.;
                    retrend = __dadd_rn(retrend, __dmul_rn(t, double(slope)));
                }
            } else {
                retrend = __dadd_rn(retrend, __dmul_rn(v, double(slope)));
            }
        }
    }
    return float(retrend);
}

// tail of interpolate(): NODATA passthrough, retrend, clamp (interpolation.cpp:2532-2558)
__device__ __forceinline__ float finishEstimate(const DevModel& m, float result, const double* proxyValues)
{
    if (isEqualF(result, kNoData)) return kNoData;
    if (!m.useDoNotRetrend) result = __fadd_rn(result, retrendValue(m, proxyValues));
    switch (m.clampMode) {
    case CLAMP_PREC: return (result < m.rainfallThreshold) ? 0.f : result;
    case CLAMP_RH: return fmaxf(0.f, fminf(result, 100.f));
    case CLAMP_POS: return fmaxf(result, 0.f);
    default: return result;
    }
}


}  // namespace pb
