/*
 * oracle/oracle_port.cpp — TEST INFRASTRUCTURE, never shipped, never on the product path.
 *
 * Plain C++ CPU restatement of the gridding hot path of ARPA-SIMC/PRAGA (agrolib), written from the
 * reference's behaviour (each function cites the reference file:line it follows; paths are relative to
 * /root/reference/agrolib). It exists so that a checker is available where the reference itself cannot be
 * built (the GPU box has no /root/reference). Parity status: PINNED — tests/test_oracle_port.py compares every
 * entry point below with the reference's own code (oracle/_ref/libpraga_ref.so) on seeded inputs, and with the
 * golden fixtures under tests/golden/ generated from the reference by tests/golden/make_golden.py.
 *
 * Arithmetic follows the reference operation by operation (float vs double, summation order); it is built
 * with -ffp-contract=off so that no FMA is formed, like the reference's x86-64 release build.
 *
 * Same C ABI as oracle/ref_driver.cpp with the prefix port_.
 */
#include <math.h>
#include <omp.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <vector>

#include "../include/praga_b200.h"

namespace {

const float NODATA_F = -9999.f;
const double EPSILON = 0.00001;   // mathFunctions/commonConstants.h:252
const int MIN_REGRESSION_POINTS = 5;   // interpolation/interpolationConstants.h:4

inline bool isEqualF(float a, float b) { return fabs(double(a) - double(b)) < EPSILON; }   // mathFunctions/basicMath.h:25
inline bool isEqualD(double a, double b) { return fabs(a - b) < EPSILON; }                 // basicMath.h:28

// meteo/meteo.cpp:1127-1133
inline bool checkLapseRateCode(int code, bool useLapseRateCode, bool useSupplemental)
{
    if (useSupplemental) return !useLapseRateCode || code == PB_LAPSE_PRIMARY || code == PB_LAPSE_SUPPLEMENTAL;
    return !useLapseRateCode || code == PB_LAPSE_PRIMARY || code == PB_LAPSE_SECONDARY;
}
// interpolation/interpolation.cpp:1190-1234
inline bool isThermal(int v)
{
    return v == PB_VAR_AIR_TEMPERATURE || v == PB_VAR_AIR_DEW_TEMPERATURE || v == PB_VAR_DAILY_TAVG ||
           v == PB_VAR_DAILY_TMAX || v == PB_VAR_DAILY_TMIN || v == PB_VAR_DAILY_ET0_HS || v == PB_VAR_ELABORATION;
}
inline bool useDetrendingVar(int v) { return isThermal(v) || v == PB_VAR_ATM_PRESSURE; }
inline bool isPrec(int v) { return v == PB_VAR_PRECIPITATION || v == PB_VAR_DAILY_PRECIPITATION; }

struct Point {            // Crit3DInterpolationDataPoint, interpolation/interpolationPoint.h:14-31
    double x, y, z;
    float value, regressionWeight;
    int lapseRateCode, index;
    int meteoIndex;           // Crit3DInterpolationDataPoint::index as the caller gave it (macro areas refer to it)
    bool isActive;
    float proxy[PB_MAX_PROXY];
    int nProxy;
    float getProxyValue(int pos) const { return pos < nProxy ? proxy[pos] : NODATA_F; }   // interpolationPoint.cpp:51
};

std::vector<Point> loadPoints(const pb_stations* st)
{
    std::vector<Point> pts(st->n);
    for (int i = 0; i < st->n; i++) {
        Point& p = pts[i];
        p.x = st->x[i]; p.y = st->y[i]; p.z = st->z[i];
        p.value = st->value[i];
        p.regressionWeight = st->regressionWeight ? st->regressionWeight[i] : NODATA_F;
        p.lapseRateCode = st->lapseRateCode ? st->lapseRateCode[i] : PB_LAPSE_PRIMARY;
        p.isActive = st->isActive ? st->isActive[i] != 0 : true;
        p.index = i;
        p.meteoIndex = st->index ? st->index[i] : i;
        p.nProxy = st->nProxy;
        for (int k = 0; k < st->nProxy && k < PB_MAX_PROXY; k++) p.proxy[k] = st->proxyValues[size_t(i) * st->nProxy + k];
    }
    return pts;
}

struct Grid {
    pb_raster_header h;
    const float* data;
    float* const* rows;
    float at(int row, int col) const { return data ? data[size_t(row) * h.nrCols + col] : rows[row][col]; }
    // Crit3DRasterGrid::getValueFromXY, gis/gis.cpp:533-546
    float valueFromXY(double x, double y) const
    {
        const int r = int((y - h.lly) * h.invCellSize);
        const int row = (h.nrRows - 1) - r;
        const int col = int((x - h.llx) * h.invCellSize);
        if (unsigned(row) >= unsigned(h.nrRows) || unsigned(col) >= unsigned(h.nrCols)) return h.flag;
        return at(row, col);
    }
};

// ---- model functions, mathFunctions/furtherMathFunctions.cpp:115-201 (par is mutated by the three-piece forms) ----
double evalFit(int f, double x, double* par)
{
    switch (f) {
    case PB_FIT_PIECEWISE_TWO:                               // lapseRatePiecewise_two :115-132
        if (x < par[0]) return par[2] * (x - par[0]) + par[1];
        return par[3] * (x - par[0]) + par[1];
    case PB_FIT_PIECEWISE_THREE: {                           // lapseRatePiecewise_three :134-147
        par[2] = std::max(10., par[2]);
        double xb = par[2] + par[0];
        if (x < par[0]) return par[4] * x - par[0] * par[4] + par[1];
        else if (x > xb) return par[4] * x - par[4] * par[0] - par[4] * par[2] + par[3] * par[2] + par[1];
        else return par[3] * x - par[3] * par[0] + par[1];
    }
    case PB_FIT_PIECEWISE_THREE_FREE: {                      // lapseRatePiecewise_three_free :149-179
        par[2] = std::max(10., par[2]);
        double xb = par[0] + par[2];
        if (x < par[0]) return par[4] * x - par[0] * par[4] + par[1];
        else if (x > xb) return par[5] * x - par[5] * par[0] - par[5] * par[2] + par[3] * par[2] + par[1];
        else return par[3] * x - par[3] * par[0] + par[1];
    }
    case PB_FIT_LINEAR:                                      // functionLinear_intercept :198-201
        return par[0] * x + par[1];
    default: return 0.;
    }
}

// ---- interpolate() and what it calls ---------------------------------------------------------------------

// retrend, interpolation/interpolation.cpp:1288-1351 (+ getMultipleDetrendingValues :2563-2617, functionSum :181)
float retrend(const pb_settings& s, int var, const double* proxyValues)
{
    if (!useDetrendingVar(var)) return 0.f;
    double retrendValue = 0.;
    if (s.useMultipleDetrending) {
        double act[PB_MAX_PROXY];
        int nAct = 0, elevationPos = -1;
        for (int i = 0; i < s.nProxy; i++)
            if (s.proxy[i].kind == PB_PROXY_HEIGHT && s.currentActive[i] && s.currentSignificant[i]) {
                elevationPos = i;
                if (proxyValues[i] == NODATA_F) return 0.f;
                act[nAct++] = proxyValues[i];
            }
        for (int i = 0; i < s.nProxy; i++)
            if (i != elevationPos && s.currentActive[i] && s.currentSignificant[i]) {
                if (proxyValues[i] == NODATA_F) return 0.f;
                act[nAct++] = proxyValues[i];
            }
        if (nAct == 0) return 0.f;
        if (s.nFitFunctions > 0 && s.nFit > 0) {
            double result = 0.;
            for (int k = 0; k < s.nFitFunctions && k < nAct && k < s.nFit; k++) {
                double par[PB_MAX_FIT_PARAMS];
                memcpy(par, s.fitParams[k], sizeof(par));
                result += evalFit(s.fitFunction[k], act[k], par);
            }
            retrendValue = float(result);
        }
    } else {
        for (int pos = 0; pos < s.nProxy; pos++) {
            if (!(s.currentActive[pos] && s.currentSignificant[pos])) continue;
            const double v = proxyValues[pos];
            if (v == NODATA_F) continue;
            const pb_proxy& p = s.proxy[pos];
            const float proxySlope = p.regressionSlope;
            if (p.kind == PB_PROXY_HEIGHT) {
                if (s.useThermalInversion && p.inversionIsSignificative) {
                    const float H0 = p.lapseRateH0, H1 = p.lapseRateH1, below = p.inversionLapseRate;
                    if (v <= H1) retrendValue += (std::max(v - H0, 0.) * below);
                    else retrendValue += ((H1 - H0) * below) + (v - H1) * proxySlope;
                } else retrendValue += std::max(v, 0.) * proxySlope;
            } else retrendValue += v * proxySlope;
        }
    }
    return float(retrendValue);
}

// getUseTdVar, interpolation.cpp:1220-1234
inline bool useTdVar(int v)
{
    return !(v == PB_VAR_PRECIPITATION || v == PB_VAR_DAILY_PRECIPITATION || v == PB_VAR_GLOBAL_IRRADIANCE ||
             v == PB_VAR_ATM_TRANSMISSIVITY || v == PB_VAR_DAILY_GLOBAL_RADIATION || v == PB_VAR_ATM_PRESSURE ||
             v == PB_VAR_DAILY_WATER_TABLE_DEPTH);
}

// gis::topographicDistance, gis/gis.cpp:1595-1647
float topographicDistance(float x1, float y1, float z1, float x2, float y2, float z2, float distance, const Grid& dem)
{
    const float stepMeter = float(dem.h.cellSize);
    if (distance < stepMeter) return 0;
    const int nrStep = int(distance / stepMeter);
    float Xi, Yi, Zi, Xf, Yf;
    if (z1 < z2) { Xi = x1; Yi = y1; Zi = z1; Xf = x2; Yf = y2; }
    else { Xi = x2; Yi = y2; Zi = z2; Xf = x1; Yf = y1; }
    const float dx = (Xf - Xi) / float(nrStep), dy = (Yf - Yi) / float(nrStep);
    float x = Xi, y = Yi, maxDeltaZ = 0;
    for (int i = 1; i <= nrStep; i++) {
        x = x + dx;
        y = y + dy;
        const float demValue = dem.valueFromXY(x, y);
        if (!isEqualF(demValue, dem.h.flag))
            if (demValue > Zi) maxDeltaZ = std::max(maxDeltaZ, demValue - Zi);
    }
    return maxDeltaZ;
}

// sortPointsByDistance, interpolation.cpp:121-160 (indices into the input list)
std::vector<int> sortByDistance(unsigned maxNr, const std::vector<int>& list, const std::vector<float>& dist)
{
    std::vector<int> idx;
    for (int i : list)
        if (!isEqualF(dist[i], 0) && !isEqualF(dist[i], NODATA_F)) idx.push_back(i);
    std::sort(idx.begin(), idx.end(), [&dist](int a, int b) { return dist[a] < dist[b]; });
    if (idx.size() > maxNr) idx.resize(maxNr);
    return idx;
}

// shepardSearchNeighbour, interpolation.cpp:806-868 (computeShepardInitialRadius :800-803)
float shepardSearchNeighbour(const std::vector<Point>& pts, const std::vector<float>& dist, const pb_settings& s, std::vector<int>& sel)
{
    const unsigned nrPoints = unsigned(pts.size());
    const unsigned avg = 8;                                        // SHEPARD_AVG_NRPOINTS
    const float R0 = float(sqrt((avg * s.pointsBoundingBoxArea) / (float(3.1415926535898) * nrPoints)));
    std::vector<int> first, all;
    for (unsigned i = 0; i < nrPoints; i++) {
        all.push_back(int(i));
        if (dist[i] <= R0 && dist[i] > 0) first.push_back(int(i));
    }
    if (first.size() < 5) {
        sel = sortByDistance(5, all, dist);
        if (sel.empty()) return NODATA_F;
        return dist[sel.back()] + float(EPSILON);
    }
    if (first.size() > 10) {
        sel = sortByDistance(10, first, dist);
        return dist[sel.back()] + float(EPSILON);
    }
    sel = first;
    return R0;
}

// shepardIdw, interpolation.cpp:871-945
float shepardIdw(const std::vector<Point>& pts, const std::vector<float>& dist, const pb_settings& s, float x, float y)
{
    std::vector<int> sel;
    const float radius = shepardSearchNeighbour(pts, dist, s, sel);
    const unsigned n = unsigned(sel.size());
    std::vector<double> weight(n), t(n), S(n);
    double weightSum = 0;
    const double radius_3 = radius / 3.;
    const double radius_27_4 = 6.75 / radius;
    for (unsigned i = 0; i < n; i++) {
        const float d = dist[sel[i]];
        if (d > 0) {
            if (d <= radius_3) S[i] = 1 / d;
            else if (d <= radius) {
                const double tmp = (d / radius) - 1;
                S[i] = radius_27_4 * tmp * tmp;
            } else S[i] = 0;
            weightSum += S[i];
        }
    }
    if (weightSum == 0) return NODATA_F;
    for (unsigned i = 0; i < n; i++) {
        t[i] = 0;
        for (unsigned j = 0; j < n; j++)
            if (i != j) {
                const double cosine = ((x - pts[sel[i]].x) * (x - pts[sel[j]].x) + (y - pts[sel[i]].y) * (y - pts[sel[j]].y)) /
                                      (dist[sel[i]] * dist[sel[j]]);
                t[i] += S[j] * (1 - cosine);
            }
        if (weightSum != 0) t[i] /= weightSum;
    }
    weightSum = 0;
    for (unsigned i = 0; i < n; i++) {
        weight[i] = S[i] * S[i] * (1 + t[i]);
        weightSum += weight[i];
    }
    for (unsigned i = 0; i < n; i++) weight[i] /= weightSum;
    double result = 0;
    for (unsigned i = 0; i < n; i++) result += weight[i] * pts[sel[i]].value;
    return float(result);
}

// modifiedShepardIdw, interpolation.cpp:948-1028. The definition's parameters are (radius, y, x) while the declaration and
// the call site use (radius, x, y): the caller's x arrives in `y` and vice versa. Kept.
float modifiedShepardIdw(const std::vector<Point>& pts, const std::vector<float>& dist, const pb_settings& s, float radius,
                         float y, float x)
{
    std::vector<int> sel;
    if (isEqualF(radius, NODATA_F)) radius = shepardSearchNeighbour(pts, dist, s, sel);
    else for (size_t i = 0; i < pts.size(); i++) sel.push_back(int(i));
    if (sel.empty()) return NODATA_F;
    const size_t n = sel.size();
    std::vector<double> weight(n), t(n, 0.0), sv(n, 0.0);
    double weightSum = 0.0;
    for (size_t i = 0; i < n; i++) {
        const float d = dist[sel[i]];
        if (d > 0.0 && d <= radius) {
            sv[i] = (radius - d) / (radius * d);
            weightSum += sv[i];
        }
    }
    if (weightSum == 0.0) return NODATA_F;
    const double invWeightSum = 1.0 / weightSum;
    for (size_t i = 0; i < n; i++) {
        if (sv[i] == 0.0 || dist[sel[i]] <= 0.0) continue;
        const double xi = pts[sel[i]].x, yi = pts[sel[i]].y;
        for (size_t j = 0; j < n; j++) {
            if (i == j || sv[j] == 0.0 || dist[sel[j]] <= 0.0) continue;
            const double xj = pts[sel[j]].x, yj = pts[sel[j]].y;
            const double cosine = ((x - xi) * (x - xj) + (y - yi) * (y - yj)) / (dist[sel[i]] * dist[sel[j]]);
            t[i] += sv[j] * (1.0 - cosine);
        }
        t[i] *= invWeightSum;
    }
    weightSum = 0.0;
    for (size_t i = 0; i < n; i++) {
        weight[i] = sv[i] * sv[i] * (1.0 + t[i]);
        weightSum += weight[i];
    }
    const double inv = 1.0 / weightSum;
    for (size_t i = 0; i < n; i++) weight[i] *= inv;
    double result = 0.0;
    for (size_t i = 0; i < n; i++) result += weight[i] * pts[sel[i]].value;
    return float(result);
}

// interpolate, interpolation.cpp:2502-2560 with computeDistances :208-256 (topographic distance when a DEM is given
// and getUseTD() holds; no per-station TD rasters), inverseDistanceWeighted :1031-1051 and the Shepard estimators
float interpolate(const std::vector<Point>& pts, const pb_settings& s, int var, float x, float y, const double* proxyValues,
                  bool excludeSupplemental, float z = 0.f, const Grid* dem = nullptr)
{
    const bool td = dem && s.useTD && !s.useLocalDetrending && useTdVar(var) && s.topoDist_Kh != 0;
    if (isPrec(var) && s.precipitationAllZero) return 0.f;
    float result = NODATA_F;
    if (!s.useRetrendOnly) {
        std::vector<float> dist(pts.size());
        for (size_t i = 0; i < pts.size(); i++) {
            float d;
            if (excludeSupplemental && !checkLapseRateCode(pts[i].lapseRateCode, s.useLapseRateCode != 0, false)) d = 0;
            else {
                const float dx = float(pts[i].x) - x, dy = float(pts[i].y) - y;     // gis::computeDistance, gis.cpp:685-691
                d = sqrtf(dx * dx + dy * dy);
                if (td) {
                    const float topo = topographicDistance(x, y, z, float(pts[i].x), float(pts[i].y), float(pts[i].z), d, *dem);
                    d += (s.topoDist_Kh * topo);
                }
            }
            dist[i] = d;
        }
        if (s.method == PB_METHOD_SHEPARD) result = shepardIdw(pts, dist, s, x, y);
        else if (s.method == PB_METHOD_SHEPARD_MODIFIED) {
            float radius = NODATA_F;
            if (s.useLocalDetrending) radius = s.localRadius;
            result = modifiedShepardIdw(pts, dist, s, radius, x, y);
        } else {
            double sum = 0, sumWeights = 0;
            for (size_t i = 0; i < pts.size(); i++) {
                const float d = dist[i];
                if (d > 0.f) {
                    const double dist_km = double(d) / 10000.;
                    const double weight = 1.0 / (dist_km * dist_km * dist_km);
                    sumWeights += weight;
                    sum += double(pts[i].value) * weight;
                }
            }
            result = (sumWeights > 0.0) ? float(sum / sumWeights) : NODATA_F;
        }
    } else result = 0;
    if (isEqualF(result, NODATA_F)) return NODATA_F;
    if (!s.useDoNotRetrend) result += retrend(s, var, proxyValues);
    switch (var) {
    case PB_VAR_PRECIPITATION: case PB_VAR_DAILY_PRECIPITATION:
        return (result < s.rainfallThreshold) ? 0.f : result;
    case PB_VAR_AIR_REL_HUMIDITY: case PB_VAR_DAILY_RH_AVG: case PB_VAR_DAILY_RH_MIN: case PB_VAR_DAILY_RH_MAX:
        return std::max(0.f, std::min(result, 100.f));
    case PB_VAR_DAILY_TRANGE: case PB_VAR_LEAF_WETNESS: case PB_VAR_DAILY_LEAF_WETNESS: case PB_VAR_GLOBAL_IRRADIANCE:
    case PB_VAR_DAILY_GLOBAL_RADIATION: case PB_VAR_ATM_TRANSMISSIVITY: case PB_VAR_WIND_SCALAR_INTENSITY:
    case PB_VAR_WIND_VECTOR_INTENSITY: case PB_VAR_DAILY_WIND_SCALAR_INT_AVG: case PB_VAR_DAILY_WIND_SCALAR_INT_MAX:
    case PB_VAR_DAILY_WIND_VECTOR_INT_AVG: case PB_VAR_DAILY_WIND_VECTOR_INT_MAX: case PB_VAR_ATM_PRESSURE:
        return std::max(result, 0.f);
    default: return result;
    }
}

// ---- statistics::linearRegression, mathFunctions/statistics.cpp:1030-1092 -------------------------------
void linearRegression(const float* x, const float* y, long n, bool zeroIntercept, float* y_intercept, float* mySlope, float* r2)
{
    double SUMx = 0, SUMy = 0, SUMxy = 0, SUMxx = 0, AVGy = 0, AVGx = 0, dy = 0, SUM_dy = 0, SUMres = 0, res = 0;
    long nrValidItems = 0;
    *mySlope = 0; *y_intercept = 0; *r2 = 0;
    for (long i = 0; i < n; i++)
        if (x[i] != NODATA_F && y[i] != NODATA_F) {
            nrValidItems++;
            SUMx += x[i]; SUMy += y[i];
            SUMxy += x[i] * y[i];          // float product, as in the reference
            SUMxx += x[i] * x[i];
        }
    AVGy = SUMy / nrValidItems;
    AVGx = SUMx / nrValidItems;
    if (!zeroIntercept) {
        *mySlope = float((SUMxy - SUMx * SUMy / nrValidItems) / (SUMxx - SUMx * SUMx / nrValidItems));
        *y_intercept = float(AVGy - *mySlope * AVGx);
    } else {
        *mySlope = float(SUMxy / SUMxx);
        *y_intercept = 0.;
    }
    for (long i = 0; i < n; i++)
        if (x[i] != NODATA_F && y[i] != NODATA_F) {
            dy = y[i] - (*y_intercept + *mySlope * x[i]);
            dy *= dy;
            SUM_dy += dy;
            res = y[i] - AVGy;
            res *= res;
            SUMres += res;
        }
    *r2 = float((SUMres - SUM_dy) / SUMres);
}

// ---- classic detrending, interpolation.cpp:304-798, 1236-1285, 1354-1415 --------------------------------
struct Work {
    std::vector<Point> pts;
    pb_settings* s;
    int var;
};

float getMaxHeight(const std::vector<Point>& p, bool useCode)   // interpolation.cpp:60-71
{
    float zMax = NODATA_F;
    for (size_t i = 0; i < p.size(); i++)
        if (p[i].value != NODATA_F && p[i].isActive && checkLapseRateCode(p[i].lapseRateCode, useCode, true))
            if (zMax == NODATA_F || p[i].z > zMax) zMax = float(p[i].z);
    return zMax;
}
float getMinHeight(const std::vector<Point>& p, bool useCode)   // interpolation.cpp:49-58
{
    float zMin = NODATA_F;
    for (size_t i = 0; i < p.size(); i++)
        if (p[i].z != NODATA_F && p[i].isActive && checkLapseRateCode(p[i].lapseRateCode, useCode, true))
            if (zMin == NODATA_F || p[i].z < zMin) zMin = float(p[i].z);
    return zMin;
}

int indexHeight(const pb_settings& s)   // Crit3DInterpolationSettings::addProxy sets indexHeight to the LAST height proxy (:798)
{
    int idx = -1;
    for (int i = 0; i < s.nProxy; i++) if (s.proxy[i].kind == PB_PROXY_HEIGHT) idx = i;
    return idx;
}

// regressionSimple, interpolation.cpp:304-344
bool regressionSimple(Work& w, int proxyPos, bool zeroIntercept, float* coeff, float* intercept, float* r2)
{
    std::vector<float> vals, zs;
    *coeff = *intercept = *r2 = NODATA_F;
    const int ih = indexHeight(*w.s);
    for (size_t i = 0; i < w.pts.size(); i++) {
        const Point& p = w.pts[i];
        if (!p.isActive) continue;
        if (proxyPos != ih || checkLapseRateCode(p.lapseRateCode, w.s->useLapseRateCode != 0, true)) {
            const float pv = p.getProxyValue(proxyPos);
            if (!isEqualF(pv, NODATA_F)) { vals.push_back(p.value); zs.push_back(pv); }
        }
    }
    if (vals.size() >= size_t(MIN_REGRESSION_POINTS)) {
        linearRegression(zs.data(), vals.data(), long(zs.size()), zeroIntercept, intercept, coeff, r2);
        return true;
    }
    return false;
}

// regressionGeneric, interpolation.cpp:346-365
bool regressionGeneric(Work& w, int proxyPos, bool zeroIntercept)
{
    float q, m, r2;
    if (!regressionSimple(w, proxyPos, zeroIntercept, &m, &q, &r2)) return false;
    pb_proxy& p = w.s->proxy[proxyPos];
    p.regressionSlope = m;
    p.regressionIntercept = q;
    p.regressionR2 = r2;
    p.lapseRateT0 = q;
    p.inversionIsSignificative = 0;
    p.inversionLapseRate = NODATA_F;
    return r2 >= w.s->minRegressionR2;
}

void initializeOrography(pb_proxy& p)   // Crit3DProxy::initializeOrography, interpolationSettings.cpp:779-792
{
    p.lapseRateH0 = 0.f;
    p.lapseRateH1 = NODATA_F;
    p.lapseRateT0 = NODATA_F;
    p.lapseRateT1 = NODATA_F;
    p.inversionLapseRate = NODATA_F;
    p.regressionSlope = NODATA_F;
    p.regressionIntercept = NODATA_F;
    p.regressionR2 = NODATA_F;
    p.inversionIsSignificative = 0;
}

// regressionSimpleT, interpolation.cpp:368-406
bool regressionSimpleT(Work& w, int pos)
{
    float slope, t0, r2;
    pb_proxy& p = w.s->proxy[pos];
    initializeOrography(p);
    if (!regressionSimple(w, pos, false, &slope, &t0, &r2)) return false;
    if (r2 < w.s->minRegressionR2) return false;
    p.regressionSlope = slope;
    p.lapseRateT0 = t0;
    p.regressionR2 = r2;
    if (slope > 0 && w.var != PB_VAR_ELABORATION) {
        p.inversionLapseRate = slope;
        float maxZ = std::min(getMaxHeight(w.pts, w.s->useLapseRateCode != 0), w.s->maxHeightInversion);
        p.lapseRateT1 = t0 + slope * maxZ;
        p.lapseRateH1 = maxZ;
        p.regressionSlope = w.s->climateLapseRate;
        p.inversionIsSignificative = 1;
    } else {
        p.inversionIsSignificative = 0;
        p.inversionLapseRate = NODATA_F;
    }
    return true;
}

// findHeightIntervalAvgValue, interpolation.cpp:409-431
float findHeightIntervalAvgValue(bool useCode, const std::vector<Point>& pts, float heightInf, float heightSup, float maxPointsZ)
{
    float mySum = 0.f, nValues = 0.f;
    for (size_t i = 0; i < pts.size(); i++)
        if (pts[i].z != NODATA_F && pts[i].isActive && checkLapseRateCode(pts[i].lapseRateCode, useCode, true))
            if (pts[i].z >= heightInf && pts[i].z <= heightSup) {
                const float v = pts[i].value;
                if (v != NODATA_F) { mySum += v; nValues++; }
            }
    if (nValues > 1 || (nValues > 0 && heightSup >= maxPointsZ)) return mySum / nValues;
    return NODATA_F;
}

bool linesIntersection(float q1, float m1, float q2, float m2, float* x, float* y)   // basicMath.cpp:138-152
{
    if (!isEqualF(m1, m2)) {
        *x = (q2 - q1) / (m1 - m2);
        *y = m1 * (q2 - q1) / (m1 - m2) + q1;
        return true;
    }
    *x = NODATA_F; *y = NODATA_F;
    return false;
}
bool linesIntersectionAbove(float q1, float m1, float q2, float m2, float thr, float* x, float* y)   // basicMath.cpp:154-171
{
    if (!isEqualF(m1, m2)) {
        *x = (q2 - q1) / (m1 - m2);
        *y = m1 * (q2 - q1) / (m1 - m2) + q1;
        return *x > thr;
    }
    *x = NODATA_F; *y = NODATA_F;
    return false;
}

// regressionOrographyT, interpolation.cpp:433-798 (climateExists = true from regressionOrography :1365)
bool regressionOrographyT(Work& w, int pos)
{
    std::vector<float> d1, d2, h1, h2, ih, ih1, ih2, iv, iv1, iv2;
    float m, q, r2, r2_values, r2_intervals, q1, m1, r21, q2, m2, r22, x, y;
    const float DELTAZ_INI = 80.f;
    pb_settings& s = *w.s;
    const bool useCode = s.useLapseRateCode != 0;
    const float maxHeightInv = s.maxHeightInversion;
    pb_proxy& p = s.proxy[pos];
    const float sigR2 = std::max(s.minRegressionR2, 0.2f);
    const float sigR2Inv = std::max(s.minRegressionR2, 0.1f);
    initializeOrography(p);
    const float climateLapseRate = s.climateLapseRate;
    p.regressionSlope = climateLapseRate;
    const float maxPointsZ = getMaxHeight(w.pts, useCode);
    float heightInf = getMinHeight(w.pts, useCode);
    if (maxPointsZ == heightInf || w.pts.size() < size_t(MIN_REGRESSION_POINTS)) return true;

    float heightSup = heightInf, deltaZ = DELTAZ_INI;
    while (heightSup <= maxPointsZ) {
        float avg = NODATA_F;
        while (avg == NODATA_F) {
            heightSup = heightSup + deltaZ;
            avg = findHeightIntervalAvgValue(useCode, w.pts, heightInf, heightSup, maxPointsZ);
        }
        ih.push_back((heightSup + heightInf) / 2.f);
        iv.push_back(avg);
        deltaZ = DELTAZ_INI * expf(heightInf / maxHeightInv);
        heightInf = heightSup;
    }

    p.lapseRateT1 = iv[0];
    p.lapseRateH1 = ih[0];
    for (size_t i = 1; i < iv.size(); i++)
        if (ih[i] <= maxHeightInv && iv[i] >= p.lapseRateT1 && (iv[i] > (iv[0] + 0.001 * (ih[i] - ih[0])))) {
            p.lapseRateH1 = ih[i];
            p.lapseRateT1 = iv[i];
            p.inversionIsSignificative = 1;
        }
    if (!p.inversionIsSignificative) return regressionGeneric(w, pos, false);

    for (size_t i = 0; i < w.pts.size(); i++)
        if (w.pts[i].z != NODATA_F && checkLapseRateCode(w.pts[i].lapseRateCode, useCode, true)) {
            if (w.pts[i].z <= p.lapseRateH1) { d1.push_back(w.pts[i].value); h1.push_back(float(w.pts[i].z)); }
            else { d2.push_back(w.pts[i].value); h2.push_back(float(w.pts[i].z)); }
        }
    for (size_t i = 0; i < iv.size(); i++)
        if (ih[i] <= p.lapseRateH1) { iv1.push_back(iv[i]); ih1.push_back(ih[i]); }
        else { iv2.push_back(iv[i]); ih2.push_back(ih[i]); }

    if (p.inversionIsSignificative && iv1.size() == iv.size()) {
        if (!regressionSimple(w, pos, false, &m, &q, &r2)) return false;
        if (r2 >= sigR2) {
            p.inversionLapseRate = m;
            p.regressionR2 = r2;
            p.lapseRateT0 = q;
            p.lapseRateT1 = q + m * p.lapseRateH1;
        } else {
            linearRegression(ih1.data(), iv1.data(), long(ih1.size()), false, &q, &m, &r2);
            p.regressionR2 = NODATA_F;
            if (r2 >= sigR2) {
                p.lapseRateT0 = q;
                p.inversionLapseRate = m;
                p.lapseRateT1 = q + m * p.lapseRateH1;
            } else {
                p.inversionLapseRate = 0.f;
                p.lapseRateT0 = iv[0];
            }
        }
        return true;
    }

    linearRegression(h1.data(), d1.data(), long(h1.size()), false, &q1, &m1, &r2_values);
    if (iv1.size() > 2) linearRegression(ih1.data(), iv1.data(), long(ih1.size()), false, &q, &m, &r2_intervals);
    else r2_intervals = 0.f;

    if (r2_values < sigR2Inv && r2_intervals < sigR2Inv) {
        if (!regressionSimple(w, pos, false, &m, &q, &r2)) return false;
        if (r2 >= 0.5) {
            p.inversionIsSignificative = 0;
            p.lapseRateH1 = NODATA_F;
            p.inversionLapseRate = NODATA_F;
            p.regressionSlope = std::min(m, 0.0f);
            p.regressionR2 = r2;
            p.lapseRateT0 = q;
            p.lapseRateT1 = NODATA_F;
            return true;
        }
        p.inversionLapseRate = 0.f;
        linearRegression(h2.data(), d2.data(), long(h2.size()), false, &q2, &m2, &r2);
        if (d2.size() >= size_t(MIN_REGRESSION_POINTS)) {
            if (r2 >= sigR2) {
                p.regressionSlope = std::min(m2, 0.0f);
                p.lapseRateT0 = q2 + p.lapseRateH1 * p.regressionSlope;
                p.lapseRateT1 = p.lapseRateT0;
                p.regressionR2 = r2;
                return true;
            } else {
                linearRegression(ih2.data(), iv2.data(), long(ih2.size()), false, &q2, &m2, &r2);
                if (r2 >= sigR2) {
                    p.regressionSlope = std::min(m2, 0.0f);
                    p.lapseRateT0 = q2 + p.lapseRateH1 * p.regressionSlope;
                    p.lapseRateT1 = p.lapseRateT0;
                    p.regressionR2 = r2;
                    return true;
                }
            }
        }
        p.inversionIsSignificative = 0;
        p.lapseRateH1 = NODATA_F;
        p.inversionLapseRate = NODATA_F;
        p.lapseRateT0 = q;
        p.lapseRateT1 = NODATA_F;
        if (!regressionSimple(w, pos, false, &m, &q, &r2)) return false;
        if (r2 >= sigR2) {
            p.regressionSlope = std::min(m, 0.f);
            p.lapseRateT0 = q;
            p.regressionR2 = r2;
            return true;
        } else {
            p.lapseRateT0 = iv[0];
            if (m > 0.) p.regressionSlope = 0.f;
            else p.regressionSlope = climateLapseRate;
            return true;
        }
    }

    linearRegression(h1.data(), d1.data(), long(h1.size()), false, &q1, &m1, &r21);
    linearRegression(h2.data(), d2.data(), long(h2.size()), false, &q2, &m2, &r22);
    if (m1 <= 0) r21 = 0;
    p.regressionR2 = r22;

    if (r21 >= sigR2Inv && r22 >= sigR2) {
        if (h2.size() < size_t(MIN_REGRESSION_POINTS) && m2 > 0.) { m2 = 0.f; q2 = p.lapseRateT1; }
        linesIntersection(q1, m1, q2, m2, &x, &y);
        p.lapseRateT0 = q1;
        p.lapseRateT1 = y;
        p.inversionLapseRate = m1;
        p.regressionSlope = m2;
        p.lapseRateH1 = x;
        if (p.lapseRateH1 > maxHeightInv) {
            p.lapseRateT1 = p.lapseRateT1 - (p.lapseRateH1 - maxHeightInv) * p.regressionSlope;
            p.lapseRateH1 = maxHeightInv;
            p.inversionLapseRate = (p.lapseRateT1 - p.lapseRateT0) / (p.lapseRateH1 - p.lapseRateH0);
        }
        return true;
    } else if (r21 < sigR2Inv && r22 >= sigR2) {
        if (h2.size() < size_t(MIN_REGRESSION_POINTS) && m2 > 0.) { m2 = 0.f; q2 = p.lapseRateT1; }
        linearRegression(ih1.data(), iv1.data(), long(ih1.size()), false, &q, &m, &r2);
        p.regressionSlope = m2;
        if (r2 >= sigR2Inv) {
            if (linesIntersectionAbove(q, m, q2, m2, 40, &x, &y)) {
                p.lapseRateH1 = x;
                p.lapseRateT0 = q;
                p.lapseRateT1 = y;
                p.inversionLapseRate = m;
                if (p.lapseRateH1 > maxHeightInv) {
                    p.lapseRateT1 = p.lapseRateT1 - (p.lapseRateH1 - maxHeightInv) * p.regressionSlope;
                    p.lapseRateH1 = maxHeightInv;
                    p.inversionLapseRate = (p.lapseRateT1 - p.lapseRateT0) / (p.lapseRateH1 - p.lapseRateH0);
                }
                return true;
            }
        } else {
            p.inversionLapseRate = 0.f;
            p.lapseRateT1 = q2 + m2 * p.lapseRateH1;
            p.lapseRateT0 = p.lapseRateT1;
            return true;
        }
    } else if (r21 >= sigR2Inv && r22 < sigR2) {
        p.lapseRateT0 = q1;
        p.inversionLapseRate = m1;
        linearRegression(ih2.data(), iv2.data(), long(ih2.size()), false, &q, &m, &r2);
        if (r2 >= sigR2) {
            p.regressionSlope = std::min(m, 0.f);
            linesIntersection(p.lapseRateT0, p.inversionLapseRate, q, p.regressionSlope, &x, &y);
            p.lapseRateH1 = x;
            p.lapseRateT1 = y;
        } else {
            p.regressionSlope = climateLapseRate;
            linesIntersection(p.lapseRateT0, p.inversionLapseRate, p.lapseRateT1 - p.regressionSlope * p.lapseRateH1,
                              p.regressionSlope, &x, &y);
            p.lapseRateH1 = x;
            p.lapseRateT1 = y;
        }
        return true;
    } else if (r21 < sigR2Inv && r22 < sigR2) {
        linearRegression(ih1.data(), iv1.data(), long(ih1.size()), false, &q, &m, &r2);
        if (r2 >= sigR2Inv) {
            p.lapseRateT0 = q;
            p.inversionLapseRate = m;
            p.lapseRateT1 = q + m * p.lapseRateH1;
        } else {
            p.inversionLapseRate = 0.f;
            p.lapseRateT0 = iv[0];
            p.lapseRateT1 = p.lapseRateT0;
        }
        linearRegression(ih2.data(), iv2.data(), long(ih2.size()), false, &q, &m, &r2);
        if (r2 >= sigR2) {
            p.regressionSlope = std::min(m, 0.f);
            if (linesIntersectionAbove(p.lapseRateT0, p.inversionLapseRate, q, p.regressionSlope, 40, &x, &y)) {
                p.lapseRateH1 = x;
                p.lapseRateT1 = y;
                return true;
            }
        } else {
            p.regressionSlope = climateLapseRate;
            return true;
        }
    }
    if (p.regressionSlope < -0.02) p.regressionSlope = -0.02f;
    initializeOrography(p);
    return regressionGeneric(w, pos, false);
}

// detrendPoints, interpolation.cpp:1236-1285
void detrendPoints(Work& w, int pos)
{
    if (!useDetrendingVar(w.var)) return;
    const pb_proxy& p = w.s->proxy[pos];
    for (size_t i = 0; i < w.pts.size(); i++) {
        float detrendValue = 0;
        const float proxyValue = w.pts[i].getProxyValue(pos);
        if (p.kind == PB_PROXY_HEIGHT) {
            if (proxyValue != NODATA_F) {
                const float LR_above = p.regressionSlope;
                if (p.inversionIsSignificative) {
                    const float H1 = p.lapseRateH1, H0 = p.lapseRateH0, below = p.inversionLapseRate;
                    if (proxyValue <= H1) detrendValue = std::max(proxyValue - H0, 0.f) * below;
                    else detrendValue = ((H1 - H0) * below) + (proxyValue - H1) * LR_above;
                } else detrendValue = std::max(proxyValue, 0.f) * LR_above;
            }
        } else {
            if (proxyValue != NODATA_F)
                if (p.regressionR2 >= w.s->minRegressionR2) detrendValue = proxyValue * p.regressionSlope;
        }
        w.pts[i].value -= detrendValue;
    }
}

// detrending, interpolation.cpp:1380-1415 with regressionOrography :1354-1377
void detrendingClassic(Work& w)
{
    if (!useDetrendingVar(w.var)) return;
    pb_settings& s = *w.s;
    for (int i = 0; i < s.nProxy; i++) { s.currentActive[i] = s.selectedActive[i]; s.currentSignificant[i] = 0; }
    s.currentUseThermalInversion = s.selectedUseThermalInversion;
    for (int pos = 0; pos < s.nProxy; pos++) {
        if (!s.selectedActive[pos]) continue;
        s.currentSignificant[pos] = 0;
        bool ok;
        if (s.proxy[pos].kind == PB_PROXY_HEIGHT) {
            if (isThermal(w.var)) ok = s.selectedUseThermalInversion ? regressionOrographyT(w, pos) : regressionSimpleT(w, pos);
            else ok = regressionGeneric(w, pos, false);
        } else ok = regressionGeneric(w, pos, false);
        if (ok) {
            s.currentSignificant[pos] = 1;
            detrendPoints(w, pos);
        }
    }
}

// ---- multiple detrending, interpolation.cpp:1418-1457, 1506-1728, 1832-2236; LM: furtherMathFunctions.cpp --------

// proxyValidity, interpolation.cpp:1418-1457
bool proxyValidity(const std::vector<Point>& pts, int pos, float stdDevThreshold, double& avg, double& stdDev)
{
    std::vector<float> v;
    avg = stdDev = -9999;
    double sum = 0;
    for (size_t i = 0; i < pts.size(); i++)
        if (pts[i].isActive) {
            const float pv = pts[i].getProxyValue(pos);
            if (pv != NODATA_F) { v.push_back(pv); sum += pv; }
        }
    if (v.size() < 5) return false;
    avg = sum / v.size();
    sum = 0;
    for (size_t i = 0; i < v.size(); i++) sum += (v[i] - avg) * (v[i] - avg);
    stdDev = sqrt(sum / (v.size() - 1));
    if (stdDevThreshold != NODATA_F) return stdDev > stdDevThreshold;
    return true;
}

int elevationFunction(const pb_settings& s)   // getChosenElevationFunction, interpolationSettings.cpp:216-227
{
    int el = -1;
    for (int i = 0; i < s.nProxy; i++) if (s.proxy[i].kind == PB_PROXY_HEIGHT) el = i;
    return el >= 0 ? s.proxy[el].fittingFunction : PB_FIT_NONE;
}

void addFittingFunction(pb_settings& s, int f)
{
    if (s.nFitFunctions < PB_MAX_PROXY) s.fitFunction[s.nFitFunctions] = f;
    s.nFitFunctions++;
}
void addFittingParameters(pb_settings& s, const double* par, int n)
{
    if (s.nFit < PB_MAX_PROXY) {
        s.fitNrParams[s.nFit] = n;
        for (int j = 0; j < n && j < PB_MAX_FIT_PARAMS; j++) s.fitParams[s.nFit][j] = par[j];
    }
    s.nFit++;
}

// calculateFirstGuessCombinations, interpolation.cpp:1556-1622
void calculateFirstGuessCombinations(pb_proxy& p)
{
    const int numSteps = 15;
    const int n = p.nFitParams;
    double stepSize[PB_MAX_FIT_PARAMS], first[PB_MAX_FIT_PARAMS], tmp[PB_MAX_FIT_PARAMS];
    for (int j = 0; j < n; j++) {
        const double mn = p.fitMin[j], mx = p.fitMax[j];
        stepSize[j] = (mx - mn) / numSteps;
        if (p.fitFirstGuess[j] == 0) first[j] = mn;
        else if (p.fitFirstGuess[j] == 1) first[j] = (mn + mx) / 2;
        else first[j] = mx;
    }
    int count = 0;
    auto push = [&](const double* v) {
        if (count < PB_MAX_FIRST_GUESS) memcpy(p.firstGuessCombinations[count], v, sizeof(double) * n);
        count++;
    };
    push(first);
    int paramIndex[2] = {0, 2};
    const int nIdx = (p.fittingFunction == PB_FIT_PIECEWISE_TWO) ? 1 : 2;
    for (int k = 0; k < nIdx; k++) {
        const int pi = paramIndex[k];
        for (int m = 1; m < numSteps + 1; m++) {
            memcpy(tmp, first, sizeof(double) * n);
            // NB the reference indexes firstGuessPosition with k, not with paramIndex[k] (interpolation.cpp:1594, 1600)
            if (p.fitFirstGuess[k] == 0) {
                tmp[pi] = first[pi] + m * stepSize[pi];
                push(tmp);
            } else if (p.fitFirstGuess[k] == 2) {
                tmp[pi] = first[pi] - m * stepSize[pi];
                push(tmp);
            } else if (m < int(numSteps / 2) + 1) {
                tmp[pi] = std::min(first[pi] + m * stepSize[pi], p.fitMax[pi]);
                push(tmp);
                memcpy(tmp, first, sizeof(double) * n);
                tmp[pi] = std::max(first[pi] - m * stepSize[pi], p.fitMin[pi]);
                push(tmp);
            }
        }
    }
    p.nFirstGuessCombinations = std::min(count, PB_MAX_FIRST_GUESS);
}

// setMultipleDetrendingHeightTemperatureRange, interpolation.cpp:1506-1553
bool setMultipleDetrendingHeightTemperatureRange(pb_settings& s)
{
    if (!s.hasPointsRange || !s.useMultipleDetrending) return false;
    for (int i = 0; i < s.nProxy; i++) {
        if (!s.selectedActive[i] || s.proxy[i].kind != PB_PROXY_HEIGHT) continue;
        const double MIN_T = s.pointsRangeMin, MAX_T = s.pointsRangeMax;
        pb_proxy& p = s.proxy[i];
        const int n = p.nFitParams;
        if (n > 0) {
            // tempParam[1] = MIN_T - 2; tempParam[n + 1 (two) | n + 1 (three_free: index 7 of 12) | n + 1 (three: 6 of 10)] = MAX_T + 6
            const int f = elevationFunction(s);
            if (f == PB_FIT_PIECEWISE_TWO || f == PB_FIT_PIECEWISE_THREE_FREE || f == PB_FIT_PIECEWISE_THREE) {
                p.fitMin[1] = MIN_T - 2;
                p.fitMax[1] = MAX_T + 6;
                addFittingFunction(s, f);
            }
        }
        calculateFirstGuessCombinations(p);
    }
    return true;
}

double weightedR2(const std::vector<double>& obs, const std::vector<double>& pred, const std::vector<double>& w)   // furtherMathFunctions.cpp:1245-1275
{
    double ssr = 0, sst = 0, mean = 0, sw = 0;
    for (size_t i = 0; i < obs.size(); i++) { mean += obs[i] * w[i]; sw += w[i]; }
    mean /= sw;
    for (size_t i = 0; i < obs.size(); i++) {
        const double r = w[i] * (obs[i] - pred[i]);
        ssr += r * r;
        const double t = w[i] * (obs[i] - mean);
        sst += t * t;
    }
    return 1.0 - (ssr / sst);
}
double weightedRMSE(const std::vector<double>& obs, const std::vector<double>& pred, const std::vector<double>& w)   // :1301-1322
{
    double s = 0, sw = 0;
    for (size_t i = 0; i < obs.size(); i++) sw += w[i];
    for (size_t i = 0; i < obs.size(); i++) s += w[i] * (obs[i] - pred[i]) * (obs[i] - pred[i]);
    return sqrt(s / (sw * obs.size()));
}

// fittingMarquardt_nDimension, 1-D weighted: furtherMathFunctions.cpp:1954-2118
bool fittingMarquardt1D(int f, const double* pmin, const double* pmax, double* par, const double* delta, int nPar, int maxIter,
                        double eps, const std::vector<double>& x, const std::vector<double>& y, const std::vector<double>& wgt)
{
    const double VFACTOR = 10;
    const int nData = int(y.size());
    double paramChange[PB_MAX_FIT_PARAMS] = {0}, newPar[PB_MAX_FIT_PARAMS] = {0}, lambda[PB_MAX_FIT_PARAMS];
    for (int k = 0; k < nPar; k++) lambda[k] = 0.01;
    double mySSE = 0, diffSSE, newSSE, error;
    for (int i = 0; i < nData; i++) {
        error = y[i] - evalFit(f, x[i], par);
        mySSE += error * error * wgt[i] * wgt[i];
    }
    int iter = 0;
    std::vector<double> firstEst(nData);
    std::vector<std::vector<double>> P(nPar, std::vector<double>(nData)), WP(nPar, std::vector<double>(nData));
    do {
        double g[PB_MAX_FIT_PARAMS] = {0}, a[PB_MAX_FIT_PARAMS][PB_MAX_FIT_PARAMS] = {{0}};
        for (int i = 0; i < nData; i++) firstEst[i] = evalFit(f, x[i], par);
        for (int k = 0; k < nPar; k++) {
            par[k] += delta[k];
            for (int j = 0; j < nData; j++) P[k][j] = (evalFit(f, x[j], par) - firstEst[j]) / delta[k];
            par[k] -= delta[k];
        }
        for (int i = 0; i < nPar; i++) for (int k = 0; k < nData; k++) WP[i][k] = wgt[k] * P[i][k];
        for (int i = 0; i < nPar; i++) for (int j = i; j < nPar; j++) for (int k = 0; k < nData; k++) a[i][j] += (P[i][k] * WP[j][k]);
        for (int i = 0; i < nPar; i++) for (int k = 0; k < nData; k++) g[i] += P[i][k] * wgt[k] * (y[k] - firstEst[k]);
        for (int k = 0; k < nPar; k++) {
            a[k][k] += lambda[k] * a[k][k];
            for (int j = k + 1; j < nPar; j++) a[j][k] = a[k][j];
        }
        for (int j = 0; j < nPar - 1; j++) {
            const double pivot = std::max(a[j][j], EPSILON);
            for (int i = j + 1; i < nPar; i++) {
                const double mult = a[i][j] / pivot;
                for (int k = j + 1; k < nPar; k++) a[i][k] -= mult * a[j][k];
                g[i] -= mult * g[j];
            }
        }
        paramChange[nPar - 1] = g[nPar - 1] / std::max(a[nPar - 1][nPar - 1], EPSILON);
        for (int i = nPar - 2; i >= 0; i--) {
            double top = g[i];
            for (int k = i + 1; k < nPar; k++) top -= a[i][k] * paramChange[k];
            paramChange[i] = top / std::max(a[i][i], EPSILON);
        }
        for (int j = 0; j < nPar; j++) {
            newPar[j] = par[j] + paramChange[j];
            if (newPar[j] > pmax[j] && lambda[j] < 1000) {
                newPar[j] = pmax[j];
                if (lambda[j] < 1000) lambda[j] *= VFACTOR;
            }
            if (newPar[j] < pmin[j]) {
                newPar[j] = pmin[j];
                if (lambda[j] < 1000) lambda[j] *= VFACTOR;
            }
        }
        newSSE = 0;
        for (int i = 0; i < nData; i++) {
            error = y[i] - evalFit(f, x[i], newPar);
            newSSE += error * error * wgt[i] * wgt[i];
        }
        if (newSSE == -9999) return false;
        diffSSE = mySSE - newSSE;
        if (diffSSE > 0) {
            mySSE = newSSE;
            for (int j = 0; j < nPar; j++) { par[j] = newPar[j]; lambda[j] /= VFACTOR; }
        } else {
            for (int j = 0; j < nPar; j++) lambda[j] *= VFACTOR;
        }
        iter++;
    } while (fabs(diffSSE) > eps && iter <= maxIter);
    return fabs(diffSSE) <= eps;
}

// bestFittingMarquardt_nDimension, elevation, weighted: furtherMathFunctions.cpp:1360-1427
double bestFittingMarquardt1D(int f, const double* pmin, const double* pmax, double* par, const double* delta, int nPar, int maxIter,
                              double eps, double deltaR2, const std::vector<double>& x, const std::vector<double>& y,
                              const std::vector<double>& wgt, const pb_proxy& proxy)
{
    const int nData = int(y.size());
    double best[PB_MAX_FIT_PARAMS] = {0};
    double bestR2 = -9999, bestRMSE = -9999;
    int rmseIndex = -9999;
    std::vector<double> ySim(nData);
    for (int k = 0; k < proxy.nFirstGuessCombinations; k++) {
        memcpy(par, proxy.firstGuessCombinations[k], sizeof(double) * nPar);
        fittingMarquardt1D(f, pmin, pmax, par, delta, nPar, maxIter, eps, x, y, wgt);
        for (int i = 0; i < nData; i++) ySim[i] = evalFit(f, x[i], par);
        const double R2 = weightedR2(y, ySim, wgt);
        const double RMSE = weightedRMSE(y, ySim, wgt);
        if (isEqualD(bestRMSE, -9999) || RMSE < bestRMSE) { bestRMSE = RMSE; rmseIndex = k; }
        if (isEqualD(bestR2, -9999) || R2 > (bestR2 + deltaR2)) {
            memcpy(best, par, sizeof(double) * nPar);
            bestR2 = R2;
        }
    }
    memcpy(par, best, sizeof(double) * nPar);
    if (bestR2 < 0 && rmseIndex >= 0) {
        memcpy(par, proxy.firstGuessCombinations[rmseIndex], sizeof(double) * nPar);
        fittingMarquardt1D(f, pmin, pmax, par, delta, nPar, maxIter, eps, x, y, wgt);
    }
    return bestR2;
}

// computeR2adjusted :1175-1194, computeRMSE :1324-1340
double r2Adjusted(const std::vector<double>& obs, const std::vector<double>& sim)
{
    double meanObs = 0, RSS = 0, TSS = 0;
    for (size_t i = 0; i < obs.size(); i++) meanObs += obs[i];
    meanObs /= obs.size();
    for (size_t i = 0; i < obs.size(); i++) {
        RSS += (obs[i] - sim[i]) * (obs[i] - sim[i]);
        TSS += (obs[i] - meanObs) * (obs[i] - meanObs);
    }
    return 1 - RSS / TSS;
}
double rmsePlain(const std::vector<double>& obs, const std::vector<double>& pred)
{
    double sum = 0.0;
    for (size_t i = 0; i < obs.size(); i++) {
        const double residual = (obs[i] - pred[i]) * (obs[i] - pred[i]);
        sum += residual;
    }
    return sqrt(sum / obs.size());
}

// bestFittingMarquardt_nDimension, elevation, NO weights (glocal detrending): furtherMathFunctions.cpp:1433-1529. The descent
// fittingMarquardt_nDimension :2125-2283 is the weighted one :1954-2118 with every weight dropped; with weights of exactly
// 1.0 each product it drops is exact, so the weighted restatement above run on unit weights takes the same steps bit for bit.
double bestFittingMarquardt1DUnweighted(int f, const double* pmin, const double* pmax, double* par, const double* delta, int nPar,
                                        int maxIter, double eps, double deltaR2, const std::vector<double>& x,
                                        const std::vector<double>& y, const pb_proxy& proxy)
{
    const int nData = int(y.size());
    const std::vector<double> ones(nData, 1.0);
    double best[PB_MAX_FIT_PARAMS] = {0};
    double bestR2 = -9999, bestRMSE = -9999;
    int rmseIndex = -9999;
    std::vector<double> ySim(nData);
    double maxZ = x.front();
    for (size_t k = 0; k < x.size(); k++) if (x[k] > maxZ) maxZ = x[k];
    bool isValid = true;
    for (int k = 0; k < proxy.nFirstGuessCombinations; k++) {
        memcpy(par, proxy.firstGuessCombinations[k], sizeof(double) * nPar);
        fittingMarquardt1D(f, pmin, pmax, par, delta, nPar, maxIter, eps, x, y, ones);
        bool rangeFlag = true;
        for (int i = 0; i < nPar; i++)
            if (par[i] < pmin[i] || par[i] > pmax[i]) { rangeFlag = false; break; }
        if (!rangeFlag) continue;
        for (int i = 0; i < nData; i++) ySim[i] = evalFit(f, x[i], par);
        const double R2 = r2Adjusted(y, ySim);
        const double RMSE = rmsePlain(y, ySim);
        if (isEqualD(bestRMSE, -9999) || RMSE < bestRMSE) { bestRMSE = RMSE; rmseIndex = k; }
        if (nPar > 4) isValid = (par[0] + par[2]) < maxZ;
        if (isEqualD(bestR2, -9999) || (R2 > (bestR2 + deltaR2) && isValid)) {
            memcpy(best, par, sizeof(double) * nPar);
            bestR2 = R2;
        }
    }
    memcpy(par, best, sizeof(double) * nPar);
    if (bestR2 < 0 && rmseIndex != -9999) {
        memcpy(par, proxy.firstGuessCombinations[rmseIndex], sizeof(double) * nPar);
        fittingMarquardt1D(f, pmin, pmax, par, delta, nPar, maxIter, eps, x, y, ones);
    }
    return bestR2;
}

bool isVectorNodataOrZero(const double* v, int n)   // interpolation.cpp:2696-2705
{
    bool flag = false;
    for (int i = 0; i < n; i++) if (isEqualD(v[i], -9999) || isEqualD(v[i], 0)) flag = true;
    return flag;
}

// multipleDetrendingElevationFitting, interpolation.cpp:1862-1953 (isWeighted = true from multipleDetrendingMain :1848)
bool multipleDetrendingElevationFitting(Work& w, int elevationPos, bool isWeighted = true)
{
    pb_settings& s = *w.s;
    s.proxy[elevationPos].regressionR2 = NODATA_F;
    if (!useDetrendingVar(w.var)) return true;
    std::vector<Point> ep;
    for (auto& p : w.pts)
        if (checkLapseRateCode(p.lapseRateCode, s.useLapseRateCode != 0, true) && p.getProxyValue(elevationPos) != NODATA_F)
            ep.push_back(p);
    double avg, stdDev;
    bool isValid = proxyValidity(ep, elevationPos, s.proxy[elevationPos].stdDevThreshold, avg, stdDev);
    if (s.useLocalDetrending) isValid = isValid && ep.size() >= unsigned(s.minPointsLocalDetrending);
    s.currentSignificant[elevationPos] = isValid ? 1 : 0;
    if (!isValid) return true;
    std::vector<double> x, y, wgt;
    for (auto& p : ep) {
        x.push_back(p.getProxyValue(elevationPos));
        y.push_back(p.value);
        wgt.push_back(p.regressionWeight != NODATA_F ? p.regressionWeight : 1);
    }
    // setFittingParameters_elevation, interpolation.cpp:1625-1677
    const int f = elevationFunction(s);
    if (!(f == PB_FIT_PIECEWISE_TWO || f == PB_FIT_PIECEWISE_THREE_FREE || f == PB_FIT_PIECEWISE_THREE)) return false;
    addFittingFunction(s, f);
    const pb_proxy& p = s.proxy[elevationPos];
    const int nPar = p.nFitParams;
    if (nPar == 0) return false;
    double pmin[PB_MAX_FIT_PARAMS], pmax[PB_MAX_FIT_PARAMS], delta[PB_MAX_FIT_PARAMS], par[PB_MAX_FIT_PARAMS];
    for (int j = 0; j < nPar; j++) {
        pmin[j] = p.fitMin[j]; pmax[j] = p.fitMax[j];
        delta[j] = (pmax[j] - pmin[j]) / 800;
        par[j] = (pmax[j] + pmin[j]) / 2;
    }
    const double R2 = isWeighted ? bestFittingMarquardt1D(f, pmin, pmax, par, delta, nPar, 1000, 0.002, 0.005, x, y, wgt, p)
                                 : bestFittingMarquardt1DUnweighted(f, pmin, pmax, par, delta, nPar, 1000, 0.002, 0.005, x, y, p);
    s.proxy[elevationPos].regressionR2 = float(R2);
    if (!isEqualD(R2, -9999) || !isVectorNodataOrZero(par, nPar)) addFittingParameters(s, par, nPar);
    else s.currentSignificant[elevationPos] = 0;
    return true;
}

// detrendingElevation, interpolation.cpp:1956-1972
void detrendingElevation(Work& w, int elevationPos)
{
    pb_settings& s = *w.s;
    if (s.nFitFunctions == 0 || s.nFit == 0) return;
    double par[PB_MAX_FIT_PARAMS];
    memcpy(par, s.fitParams[0], sizeof(par));
    for (auto& p : w.pts) {
        const float proxyValue = p.getProxyValue(elevationPos);
        const float detrendValue = float(evalFit(s.fitFunction[0], proxyValue, par));
        p.value -= detrendValue;
    }
}

// fittingMarquardt_nDimension for the other proxies (functionSum of functionLinear_intercept, weighted):
// furtherMathFunctions.cpp:1740-1947; bestFitting wrapper :1535-1580 (delta clamped to >= EPSILON)
void fittingMarquardtOthers(int nPred, double pmin[][2], double pmax[][2], double par[][2], double delta[][2], int maxIter,
                            double eps, const std::vector<std::vector<double>>& x, const std::vector<double>& y,
                            const std::vector<double>& wgt)
{
    const double VFACTOR = 10;
    const int nData = int(y.size());
    const int nTot = 2 * nPred;
    auto fsum = [&](const std::vector<double>& xi, double p[][2]) {
        double r = 0.0;
        for (int c = 0; c < nPred; c++) r += p[c][0] * xi[c] + p[c][1];
        return r;
    };
    auto norm = [&](double p[][2]) {
        double n = 0;
        for (int i = 0; i < nData; i++) {
            const double e = y[i] - fsum(x[i], p);
            n += e * e * wgt[i] * wgt[i];
        }
        return n;
    };
    double lambda[PB_MAX_PROXY][2], change[PB_MAX_PROXY][2], newPar[PB_MAX_PROXY][2];
    for (int i = 0; i < nPred; i++) lambda[i][0] = lambda[i][1] = 0.01;
    double mySSE = norm(par), diffSSE, newSSE;
    int iter = 0;
    std::vector<double> firstEst(nData);
    std::vector<std::vector<double>> P(nTot, std::vector<double>(nData)), WP(nTot, std::vector<double>(nData));
    do {
        std::vector<double> g(nTot, 0.);
        std::vector<std::vector<double>> a(nTot, std::vector<double>(nTot, 0.));
        for (int i = 0; i < nData; i++) firstEst[i] = fsum(x[i], par);
        int cd = 0;
        for (int i = 0; i < nPred; i++)
            for (int k = 0; k < 2; k++) {
                par[i][k] += delta[i][k];
                for (int j = 0; j < nData; j++) P[cd][j] = (fsum(x[j], par) - firstEst[j]) / delta[i][k];
                par[i][k] -= delta[i][k];
                cd++;
            }
        for (int i = 0; i < nTot; i++) for (int k = 0; k < nData; k++) WP[i][k] = wgt[k] * P[i][k];
        for (int i = 0; i < nTot; i++) for (int j = i; j < nTot; j++) for (int k = 0; k < nData; k++) a[i][j] += (P[i][k] * WP[j][k]);
        for (int i = 0; i < nTot; i++) for (int k = 0; k < nData; k++) g[i] += P[i][k] * wgt[k] * (y[k] - firstEst[k]);
        cd = 0;
        for (int i = 0; i < nPred; i++)
            for (int k = 0; k < 2; k++) {
                a[cd][cd] += lambda[i][k] * a[cd][cd];
                for (int j = cd + 1; j < nTot; j++) a[j][cd] = a[cd][j];
                cd++;
            }
        for (int j = 0; j < nTot - 1; j++) {
            const double pivot = a[j][j];          // no EPSILON clamp in this variant (:1884)
            for (int i = j + 1; i < nTot; i++) {
                const double mult = a[i][j] / pivot;
                for (int k = j + 1; k < nTot; k++) a[i][k] -= mult * a[j][k];
                g[i] -= mult * g[j];
            }
        }
        double* ch = &change[0][0];                // correspondenceTag: flat index t <-> (t / 2, t % 2)
        ch[nTot - 1] = g[nTot - 1] / a[nTot - 1][nTot - 1];
        for (int i = nTot - 2; i >= 0; i--) {
            double top = g[i];
            for (int k = i + 1; k < nTot; k++) top -= a[i][k] * ch[k];
            ch[i] = top / a[i][i];
        }
        for (int i = 0; i < nPred; i++)
            for (int j = 0; j < 2; j++) {
                newPar[i][j] = par[i][j] + change[i][j];
                if (newPar[i][j] > pmax[i][j] && lambda[i][j] < 1000) {
                    newPar[i][j] = pmax[i][j];
                    if (lambda[i][j] < 1000) lambda[i][j] *= VFACTOR;
                }
                if (newPar[i][j] < pmin[i][j]) {
                    newPar[i][j] = pmin[i][j];
                    if (lambda[i][j] < 1000) lambda[i][j] *= VFACTOR;
                }
            }
        newSSE = norm(newPar);
        if (newSSE == -9999) return;
        diffSSE = mySSE - newSSE;
        if (diffSSE > 0) {
            mySSE = newSSE;
            for (int i = 0; i < nPred; i++) for (int j = 0; j < 2; j++) { par[i][j] = newPar[i][j]; lambda[i][j] /= VFACTOR; }
        } else {
            for (int i = 0; i < nPred; i++) for (int j = 0; j < 2; j++) lambda[i][j] *= VFACTOR;
        }
        iter++;
    } while (fabs(diffSSE) > eps && iter <= maxIter);
}

// multipleDetrendingOtherProxiesFitting, interpolation.cpp:1975-2171
bool multipleDetrendingOtherProxiesFitting(Work& w, int elevationPos)
{
    if (!useDetrendingVar(w.var)) return true;
    pb_settings& s = *w.s;
    const int proxyNr = s.nProxy;
    uint8_t active[PB_MAX_PROXY];     // myCombination is a COPY taken at entry (:1980)
    memcpy(active, s.currentActive, sizeof(active));
    unsigned nrPredictors = 0;
    for (int pos = 0; pos < proxyNr; pos++)
        if (pos != elevationPos && active[pos]) { s.currentSignificant[pos] = 0; nrPredictors++; }
    if (nrPredictors == 0) return true;
    std::vector<Point> others;
    for (auto& p : w.pts) if (checkLapseRateCode(p.lapseRateCode, s.useLapseRateCode != 0, false)) others.push_back(p);
    unsigned validNr = 0;
    double avg, stdDev;
    for (int pos = 0; pos < proxyNr; pos++)
        if (pos != elevationPos && active[pos]) {
            const bool ok = proxyValidity(others, pos, s.proxy[pos].stdDevThreshold, avg, stdDev);
            s.currentSignificant[pos] = ok ? 1 : 0;
            if (ok) validNr++;
        }
    if (validNr == 0) return true;
    // exclude points with incomplete proxies (from myPoints too: they no longer take part in the IDW) :2040-2069
    {
        std::vector<Point> kept;
        for (auto& p : w.pts) {
            bool isValid = true, atLeastOne = false;
            for (int pos = 0; pos < proxyNr; pos++)
                if (pos != elevationPos && active[pos] && s.currentSignificant[pos]) {
                    const float pv = p.getProxyValue(pos);
                    if (isEqualF(pv, NODATA_F)) { isValid = false; break; }
                    if (!isEqualF(pv, 0)) atLeastOne = true;
                }
            if (isValid && atLeastOne) kept.push_back(p);
        }
        w.pts.swap(kept);
    }
    {
        std::vector<Point> kept;
        for (auto& p : others) {
            bool isValid = true;
            for (int pos = 0; pos < proxyNr; pos++)
                if (pos != elevationPos && active[pos] && s.currentSignificant[pos])
                    if (p.getProxyValue(pos) == NODATA_F) { isValid = false; break; }
            if (isValid) kept.push_back(p);
        }
        others.swap(kept);
    }
    validNr = 0;
    for (int pos = 0; pos < proxyNr; pos++)
        if (pos != elevationPos && active[pos]) {
            const bool ok = proxyValidity(others, pos, s.proxy[pos].stdDevThreshold, avg, stdDev);
            s.currentSignificant[pos] = ok ? 1 : 0;
            if (ok) validNr++;
        }
    if (validNr == 0) return true;
    std::vector<std::vector<double>> predictors;
    std::vector<double> predictands, weights;
    for (auto& p : others) {
        std::vector<double> row;
        for (int pos = 0; pos < proxyNr; pos++)
            if (pos != elevationPos && active[pos] && s.currentSignificant[pos]) row.push_back(p.getProxyValue(pos));
        predictors.push_back(row);
        predictands.push_back(p.value);
        weights.push_back(!isEqualF(p.regressionWeight, NODATA_F) ? p.regressionWeight : 1);
    }
    if (s.useLocalDetrending && int(others.size()) < s.minPointsLocalDetrending) {
        for (int pos = 0; pos < proxyNr; pos++) if (pos != elevationPos) s.currentSignificant[pos] = 0;
        return true;
    }
    // setFittingParameters_otherProxies, interpolation.cpp:1680-1728 (uses the CURRENT combination)
    double pmin[PB_MAX_PROXY][2], pmax[PB_MAX_PROXY][2], delta[PB_MAX_PROXY][2], par[PB_MAX_PROXY][2];
    int nPred = 0;
    for (int i = 0; i < proxyNr; i++)
        if (i != elevationPos && s.currentActive[i] && s.currentSignificant[i]) {
            addFittingFunction(s, PB_FIT_LINEAR);
            const pb_proxy& p = s.proxy[i];
            if (p.nFitParams == 0) return false;
            for (int j = 0; j < 2; j++) {
                pmin[nPred][j] = p.fitMin[j]; pmax[nPred][j] = p.fitMax[j];
                delta[nPred][j] = std::max((pmax[nPred][j] - pmin[nPred][j]) / 1000, EPSILON);
                par[nPred][j] = (pmax[nPred][j] + pmin[nPred][j]) / 2;
            }
            nPred++;
        }
    if (nPred == 0) return false;
    fittingMarquardtOthers(nPred, pmin, pmax, par, delta, 100, 0.005, predictors, predictands, weights);
    for (int i = 0; i < nPred; i++) addFittingParameters(s, par[i], 2);
    return true;
}

// detrendingOtherProxies, interpolation.cpp:2173-2236
void detrendingOtherProxies(Work& w, int elevationPos)
{
    pb_settings& s = *w.s;
    if (s.nFit == 0) return;
    int first = 0;
    if (s.fitNrParams[0] > 2) first = 1;     // the height function/parameters are skipped
    if (s.nFit - first <= 0) return;
    const int nF = s.nFitFunctions - first;
    std::vector<Point> kept;
    for (auto& p : w.pts) {
        double pv[PB_MAX_PROXY];
        int n = 0;
        bool complete = true;
        for (int pos = 0; pos < s.nProxy; pos++)
            if (pos != elevationPos && s.currentActive[pos] && s.currentSignificant[pos] && complete) {
                const double v = p.getProxyValue(pos);
                if (!isEqualD(v, -9999)) pv[n++] = v;
                else complete = false;
            }
        if (complete) {
            double r = 0.0;
            for (int c = 0; c < nF && c < n; c++) {
                double par[PB_MAX_FIT_PARAMS];
                memcpy(par, s.fitParams[first + c], sizeof(par));
                r += evalFit(s.fitFunction[first + c], pv[c], par);
            }
            p.value -= float(r);
            kept.push_back(p);
        }
    }
    w.pts.swap(kept);
}

// multipleDetrendingMain, interpolation.cpp:1832-1859
bool multipleDetrendingMain(Work& w)
{
    pb_settings& s = *w.s;
    for (int i = 0; i < s.nProxy; i++) { s.currentActive[i] = s.selectedActive[i]; s.currentSignificant[i] = 0; }
    s.currentUseThermalInversion = s.selectedUseThermalInversion;
    s.nFit = s.nFitFunctions = 0;    // clearFitting
    int elevationPos = -9999;
    for (int pos = 0; pos < s.nProxy; pos++) if (s.proxy[pos].kind == PB_PROXY_HEIGHT) elevationPos = pos;
    if (elevationPos != -9999 && s.currentActive[elevationPos]) {
        if (!multipleDetrendingElevationFitting(w, elevationPos)) return false;
        if (s.currentSignificant[elevationPos]) detrendingElevation(w, elevationPos);
    }
    if (!multipleDetrendingOtherProxiesFitting(w, elevationPos)) return false;
    detrendingOtherProxies(w, elevationPos);
    return true;
}

// preInterpolation, interpolation.cpp:2444-2499 (optimalDetrending / glocal / TD optimisation are out of scope here)
bool preInterpolation(Work& w)
{
    pb_settings& s = *w.s;
    if (isPrec(w.var)) {
        int nrValid = 0;                        // checkPrecipitationZero, interpolation.cpp:1173-1186
        for (auto& p : w.pts)
            if (p.isActive && !isEqualF(p.value, NODATA_F) && p.value >= s.rainfallThreshold) nrValid++;
        if (nrValid == 0) { s.precipitationAllZero = 1; return true; }
        s.precipitationAllZero = 0;
    }
    if (useDetrendingVar(w.var)) {
        if (s.useMultipleDetrending) {
            if (!s.useLocalDetrending) setMultipleDetrendingHeightTemperatureRange(s);
            if (!multipleDetrendingMain(w)) return false;
        } else detrendingClassic(w);
    }
    return true;
}

// localSelection, interpolation.cpp:1087-1170 (computeShepardInitialRadius :800-803)
bool localSelection(const std::vector<Point>& in, std::vector<Point>& out, float x, float y, pb_settings& s, bool excludeSupplemental)
{
    const float ratioMinPoints = 1.2f;
    const unsigned minPoints = unsigned(s.minPointsLocalDetrending * ratioMinPoints);
    if (in.size() <= minPoints) {
        out = in;
        s.localRadius = float(sqrt((minPoints * s.pointsBoundingBoxArea) / (float(3.141592653589793238462643383) * unsigned(in.size()))));
        return true;
    }
    std::vector<float> distances(in.size());
    for (size_t i = 0; i < in.size(); i++) {
        const float dx = float(in[i].x) - x, dy = float(in[i].y) - y;     // gis::computeDistance, gis.cpp:685-691
        distances[i] = sqrtf(dx * dx + dy * dy);
    }
    unsigned nrValid = 0, nrPrimaries = 0;
    float maxDistance = 0, stepRadius = 7500, r0 = 0, r1 = stepRadius;
    bool beyondLastPoint = false;
    const bool useCode = s.useLapseRateCode != 0;
    std::vector<float> selD;
    out.clear();
    while (((!useCode && nrValid < minPoints) || (useCode && nrPrimaries < minPoints)) && !beyondLastPoint) {
        beyondLastPoint = true;
        for (unsigned i = 0; i < in.size(); i++) {
            const bool excl = useCode && excludeSupplemental && in[i].lapseRateCode == PB_LAPSE_SUPPLEMENTAL;
            if ((!isEqualF(distances[i], NODATA_F) && distances[i] > r0 && distances[i] <= r1) && !excl) {
                out.push_back(in[i]);
                selD.push_back(distances[i]);
                nrValid++;
                if (distances[i] > maxDistance) maxDistance = distances[i];
                if (checkLapseRateCode(in[i].lapseRateCode, useCode, true)) nrPrimaries++;
            }
            if (beyondLastPoint && distances[i] > r1 && !excl) beyondLastPoint = false;
        }
        if (nrValid > unsigned(minPoints * 0.8)) stepRadius = 1000;
        r0 = r1;
        r1 += stepRadius;
    }
    if (!isEqualF(maxDistance, 0))
        for (size_t i = 0; i < out.size(); i++) out[i].regressionWeight = std::max(1.f - selD[i] / maxDistance, float(EPSILON));
    s.localRadius = maxDistance;
    return !out.empty();
}

// one cell of the loop of Project::interpolationDemLocalDetrending, project/project.cpp:3225-3249
float localCell(const std::vector<Point>& pts, pb_settings& s, int var, double x, double y, const double* pv, std::vector<Point>* subsetOut)
{
    Work w{{}, &s, var};
    localSelection(pts, w.pts, float(x), float(y), s, false);
    if (subsetOut) *subsetOut = w.pts;
    preInterpolation(w);
    const float r = interpolate(w.pts, s, var, float(x), float(y), pv, true);
    if (subsetOut) {
        for (auto& q : *subsetOut) {
            float v = NODATA_F;
            for (auto& p : w.pts) if (p.index == q.index) v = p.value;
            q.value = v;
        }
    }
    return r;
}

// ---- glocal detrending (macro areas) ---------------------------------------------------------------------------
// glocalDetrendingFitting, interpolation.cpp:2239-2294
bool glocalDetrendingFitting(const std::vector<Point>& myPoints, pb_settings& s, int var, pb_macro_area* areas, int nAreas)
{
    int elevationPos = -9999;
    for (int pos = 0; pos < s.nProxy; pos++) if (s.proxy[pos].kind == PB_PROXY_HEIGHT) elevationPos = pos;
    for (int i = 0; i < nAreas; i++) {
        pb_macro_area& A = areas[i];
        if (A.nMeteoPoints <= 0) continue;
        Work w{{}, &s, var};
        for (int l = 0; l < A.nMeteoPoints; l++)
            for (size_t k = 0; k < myPoints.size(); k++)
                if (myPoints[k].meteoIndex == A.meteoPoints[l]) { w.pts.push_back(myPoints[k]); break; }
        for (int p = 0; p < s.nProxy; p++) { s.currentActive[p] = s.selectedActive[p]; s.currentSignificant[p] = 0; }
        s.currentUseThermalInversion = s.selectedUseThermalInversion;
        s.nFit = s.nFitFunctions = 0;    // clearFitting
        if (elevationPos != -9999 && s.selectedActive[elevationPos]) {
            if (!multipleDetrendingElevationFitting(w, elevationPos, false)) return false;
            if (s.currentSignificant[elevationPos]) detrendingElevation(w, elevationPos);
        }
        if (!multipleDetrendingOtherProxiesFitting(w, elevationPos)) return false;
        memset(A.active, 0, sizeof(A.active));
        memset(A.significant, 0, sizeof(A.significant));
        for (int p = 0; p < s.nProxy; p++) { A.active[p] = s.currentActive[p]; A.significant[p] = s.currentSignificant[p]; }
        A.useThermalInversion = s.currentUseThermalInversion;
        A.nParameters = s.nFit;
        for (int p = 0; p < PB_MAX_PROXY; p++) {
            A.nrParams[p] = p < s.nFit ? s.fitNrParams[p] : 0;
            for (int j = 0; j < PB_MAX_FIT_PARAMS; j++) A.parameters[p][j] = (p < s.nFit && j < s.fitNrParams[p]) ? s.fitParams[p][j] : 0.;
        }
    }
    return true;
}

// macroAreaDetrending, interpolation.cpp:1759-1829 (topographic-distance optimisation inside an area: not restated)
void macroAreaDetrending(const pb_macro_area& A, pb_settings& s, int var, const std::vector<Point>& pts, std::vector<Point>& subset,
                         int elevationPos)
{
    s.nFit = A.nParameters;
    for (int p = 0; p < A.nParameters; p++) {
        s.fitNrParams[p] = A.nrParams[p];
        memcpy(s.fitParams[p], A.parameters[p], sizeof(s.fitParams[p]));
    }
    for (int p = 0; p < s.nProxy; p++) { s.currentActive[p] = A.active[p]; s.currentSignificant[p] = A.significant[p]; }
    s.currentUseThermalInversion = A.useThermalInversion;
    s.nFitFunctions = 0;
    for (int l = 0; l < A.nParameters; l++) {
        const int n = A.nrParams[l];
        if (n == 2) s.fitFunction[s.nFitFunctions++] = PB_FIT_LINEAR;
        else if (n == 4) s.fitFunction[s.nFitFunctions++] = PB_FIT_PIECEWISE_TWO;
        else if (n == 5) s.fitFunction[s.nFitFunctions++] = PB_FIT_PIECEWISE_THREE;
        else if (n == 6) s.fitFunction[s.nFitFunctions++] = PB_FIT_PIECEWISE_THREE_FREE;
    }
    Work w{{}, &s, var};
    for (int l = 0; l < A.nMeteoPoints; l++)
        for (size_t k = 0; k < pts.size(); k++)
            if (pts[k].meteoIndex == A.meteoPoints[l]) w.pts.push_back(pts[k]);
    if (elevationPos != -9999 && A.active[elevationPos] && A.significant[elevationPos]) detrendingElevation(w, elevationPos);
    detrendingOtherProxies(w, elevationPos);
    subset.swap(w.pts);
}

// getSignificantProxyValuesXY, interpolation.cpp:2650-2677
bool getSignificantProxyValuesXY(float x, float y, const pb_settings& s, const std::vector<Grid>& grids, double* pv)
{
    bool complete = true;
    for (int i = 0; i < s.nProxy; i++) {
        pv[i] = -9999;
        if (s.currentActive[i] && s.currentSignificant[i]) {
            const int gi = s.proxy[i].gridIndex;
            if (gi >= 0 && gi < int(grids.size())) {
                const float v = grids[gi].valueFromXY(x, y);
                if (v != grids[gi].h.flag) pv[i] = v;
                else complete = false;
            }
        }
    }
    return complete;
}

Grid gridOf(const pb_raster& r) { return Grid{r.h, r.data, r.rows}; }

}  // namespace

extern "C" {

int port_max_threads(void) { return omp_get_max_threads(); }

// interpolationRaster, project/interpolationCmd.cpp:353-394 (+ gis::updateMinMaxRasterGrid, gis/gis.cpp:569-614)
int port_interpolation_raster(const pb_stations* st, const pb_settings* s, int variable, const pb_raster* dem,
                              const pb_raster* proxyGrids, int nProxyGrids, int nThreads, pb_raster_out* out, double* elapsed_s)
{
    std::vector<Point> pts = loadPoints(st);
    Grid d = gridOf(*dem);
    std::vector<Grid> grids;
    for (int g = 0; g < nProxyGrids; g++) grids.push_back(gridOf(proxyGrids[g]));
    const int rows = d.h.nrRows, cols = d.h.nrCols;
    if (nThreads > 0) omp_set_num_threads(nThreads);
    auto t0 = std::chrono::steady_clock::now();
    int64_t nValid = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : nValid)
    for (int row = 0; row < rows; row++) {
        double pv[PB_MAX_PROXY];
        float* orow = out->data ? out->data + size_t(row) * cols : out->rows[row];
        for (int col = 0; col < cols; col++) {
            orow[col] = d.h.flag;                                   // initializeGrid(raster), gis.cpp:357-363
            const float z = d.at(row, col);
            if (isEqualF(z, d.h.flag)) continue;
            nValid++;
            // getUtmXYFromRowColSinglePrecision(header, ...), gis.cpp:799-804
            const float x = float(d.h.llx + d.h.cellSize * (col) + 0.5);
            const float y = float(d.h.lly + d.h.cellSize * (rows - row) - 0.5);
            for (int i = 0; i < s->nProxy; i++) pv[i] = -9999;
            if (useDetrendingVar(variable)) {                       // getProxyValuesXY, interpolation.cpp:2620-2647
                for (int i = 0; i < s->nProxy; i++) {
                    if (!s->currentActive[i]) continue;
                    const int gi = s->proxy[i].gridIndex;
                    if (gi < 0 || gi >= nProxyGrids) continue;
                    const float v = grids[gi].valueFromXY(x, y);
                    if (v != grids[gi].h.flag) pv[i] = v;
                }
            }
            orow[col] = interpolate(pts, *s, variable, x, y, pv, true, z, &d);
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (elapsed_s) *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
    bool first = true;
    float mn = NODATA_F, mx = NODATA_F;
    for (int row = 0; row < rows; row++) {
        const float* orow = out->data ? out->data + size_t(row) * cols : out->rows[row];
        for (int col = 0; col < cols; col++) {
            const float v = orow[col];
            if (!isEqualF(v, d.h.flag) && !isEqualF(v, NODATA_F)) {
                if (first) { mn = mx = v; first = false; }
                else if (v < mn) mn = v;
                else if (v > mx) mx = v;
            }
        }
    }
    out->minimum = mn; out->maximum = mx; out->nValid = nValid;
    return first ? PB_ERR_NO_VALID_CELL : PB_OK;
}

int port_pre_interpolation(pb_stations* st, pb_settings* s, int variable, uint8_t* keep, const pb_raster*, int)
{
    Work w{loadPoints(st), s, variable};
    const bool ok = preInterpolation(w);
    if (keep) memset(keep, 0, st->n);
    for (auto& p : w.pts) {
        st->value[p.index] = p.value;
        if (keep) keep[p.index] = 1;
    }
    return ok ? PB_OK : PB_ERR_FIT;
}

// ---- station-space cross-validation drivers: optimalDetrending, topographicDistanceOptimize ------------------------
struct MeteoPt {        // the Crit3DMeteoPoint fields computeResiduals reads (meteo/meteoPoint.h)
    double x, y, z;
    float currentValue, residual;
    int lapseRateCode;
    bool isInsideDem;
    bool active = true;
    std::vector<float> proxy;
};

// computeResiduals, interpolation/spatialControl.cpp:97-155 (every meteo point active and accepted)
void computeResidualsMeteo(int var, std::vector<MeteoPt>& mp, const std::vector<Point>& pts, const pb_settings& s,
                           bool excludeOutsideDem, bool excludeSupplemental, const Grid* dem)
{
    for (auto& m : mp) {
        m.residual = NODATA_F;
        bool valid = !excludeSupplemental || checkLapseRateCode(m.lapseRateCode, s.useLapseRateCode != 0, false);
        valid = valid && (!excludeOutsideDem || m.isInsideDem);
        if (!valid) continue;
        double pv[PB_MAX_PROXY];
        for (int k = 0; k < PB_MAX_PROXY; k++) pv[k] = k < int(m.proxy.size()) ? double(m.proxy[k]) : -9999.;
        float myValue = m.currentValue;
        float est = interpolate(pts, s, var, float(m.x), float(m.y), pv, false, float(m.z), dem);
        if (isPrec(var)) {
            if (myValue != NODATA_F && myValue < s.rainfallThreshold) myValue = 0.f;
            if (est != NODATA_F && est < s.rainfallThreshold) est = 0.f;
        }
        if (est != NODATA_F && myValue != NODATA_F) m.residual = myValue - est;
    }
}

// computeErrorCrossValidation, spatialControl.cpp:305-328 + statistics::meanAbsoluteError, statistics.cpp:201-220
float computeErrorCrossValidation(const std::vector<MeteoPt>& mp)
{
    double sum = 0;
    long n = 0;
    bool any = false;
    for (auto& m : mp) {
        const float value = m.currentValue, residual = m.residual;
        if (value != NODATA_F && residual != NODATA_F) {
            any = true;
            const float est = value - residual;
            if (value != NODATA_F && est != NODATA_F) { sum += fabs(value - est); n++; }
        }
    }
    if (!any || n == 0) return NODATA_F;
    return float(sum / n);
}

// topographicDistanceInternalFunction :2338-2353, goldenSectionSearch :2297-2335, topographicDistanceOptimize :2356-2392
double tdInternal(int var, std::vector<MeteoPt>& mp, const std::vector<Point>& pts, pb_settings& s, const Grid* dem, double khFloat)
{
    s.topoDist_Kh = int(khFloat);
    computeResidualsMeteo(var, mp, pts, s, true, true, dem);
    return double(computeErrorCrossValidation(mp));
}
void topographicDistanceOptimize(int var, std::vector<MeteoPt>& mp, const std::vector<Point>& pts, pb_settings& s, const Grid* dem,
                                 pb_kh_series* series)
{
    if (series) series->n = 0;
    auto add = [&](float kh, float e) {
        if (series && series->n < PB_MAX_KH_SERIES) { series->kh[series->n] = kh; series->error[series->n] = e; series->n++; }
    };
    const double GOLDEN = 1.6180339887499;
    double a = 0, b = double(s.topoDist_maxKh);
    double x1 = b - (b - a) / GOLDEN, x2 = a + (b - a) / GOLDEN;
    int counter = 0;
    while (std::abs(b - a) > 1 && counter < 100) {
        counter++;
        if (tdInternal(var, mp, pts, s, dem, x1) < tdInternal(var, mp, pts, s, dem, x2)) {
            b = x2; x2 = x1; x1 = b - (b - a) / GOLDEN;
            add(float(x1), float(tdInternal(var, mp, pts, s, dem, x1)));
        } else {
            a = x1; x1 = x2; x2 = a + (b - a) / GOLDEN;
            add(float(x2), float(tdInternal(var, mp, pts, s, dem, x2)));
        }
    }
    add(float((a + b) / 2), float(tdInternal(var, mp, pts, s, dem, (a + b) / 2)));
    s.topoDist_Kh = int((a + b) / 2);
}

// passDataToInterpolation, spatialControl.cpp:500-566 (every meteo point active, accepted, not a selection)
bool passDataToInterpolation(const std::vector<MeteoPt>& mp, std::vector<Point>& pts, pb_settings& s)
{
    float xMin = 0, xMax = 0, yMin = 0, yMax = 0, vMin = 0, vMax = 0;
    bool first = true;
    int nrValid = 0;
    pts.clear();
    for (size_t i = 0; i < mp.size(); i++) {
        Point p{};
        p.index = int(i);
        p.value = mp[i].currentValue;
        p.x = mp[i].x; p.y = mp[i].y; p.z = mp[i].z;
        p.lapseRateCode = mp[i].lapseRateCode;
        p.isActive = true;
        p.regressionWeight = NODATA_F;
        p.nProxy = int(mp[i].proxy.size());
        for (int k = 0; k < p.nProxy; k++) p.proxy[k] = mp[i].proxy[k];
        if (first) {
            xMin = xMax = float(p.x); yMin = yMax = float(p.y); vMin = vMax = p.value;
            first = false;
        } else {
            xMin = std::min(xMin, float(p.x)); xMax = std::max(xMax, float(p.x));
            yMin = std::min(yMin, float(p.y)); yMax = std::max(yMax, float(p.y));
            vMin = std::min(vMin, p.value); vMax = std::max(vMax, p.value);
        }
        pts.push_back(p);
        if (checkLapseRateCode(p.lapseRateCode, s.useLapseRateCode != 0, false)) nrValid++;
    }
    if (nrValid == 0) return false;
    s.pointsBoundingBoxArea = (xMax - xMin) * (yMax - yMin);
    s.hasPointsRange = 1;
    s.pointsRangeMin = vMin;
    s.pointsRangeMax = vMax;
    return true;
}

// Crit3DInterpolationSettings::getCombination, interpolationSettings.cpp:869-892
bool getCombination(const pb_settings& s, int k, uint8_t* active, int& inversion)
{
    const int digits = s.nProxy + 1, ih = indexHeight(s);
    auto bit = [&](int pos) { return (k >> (digits - 1 - pos)) & 1; };
    if (k % 2 == 1 && (ih < 0 || !bit(ih))) return false;
    for (int i = 0; i < s.nProxy; i++) active[i] = (bit(i) && s.selectedActive[i]) ? 1 : 0;
    inversion = (bit(digits - 1) && s.selectedUseThermalInversion) ? 1 : 0;
    return true;
}

// detrending() with an explicit combination (interpolation.cpp:1380-1415)
void detrendingCombination(Work& w, const uint8_t* active, int inversion)
{
    pb_settings& s = *w.s;
    uint8_t selSave[PB_MAX_PROXY];
    memcpy(selSave, s.selectedActive, sizeof(selSave));
    const int invSave = s.selectedUseThermalInversion;
    memcpy(s.selectedActive, active, PB_MAX_PROXY);
    s.selectedUseThermalInversion = inversion;
    detrendingClassic(w);
    memcpy(s.selectedActive, selSave, sizeof(selSave));
    s.selectedUseThermalInversion = invSave;
}

// optimalDetrending, interpolation.cpp:2395-2441
void optimalDetrending(int var, std::vector<MeteoPt>& mp, std::vector<Point>& outPts, pb_settings& s, bool excludeOutsideDem,
                       const Grid* dem)
{
    const int nrCombination = 1 << (s.nProxy + 1);
    float minError = NODATA_F;
    int best = 0;
    const bool td = dem && s.useTD && !s.useLocalDetrending && useTdVar(var);
    for (int i = 0; i < nrCombination; i++) {
        uint8_t active[PB_MAX_PROXY] = {0};
        int inversion = 0;
        if (!getCombination(s, i, active, inversion)) continue;
        Work w{{}, &s, var};
        passDataToInterpolation(mp, w.pts, s);
        detrendingCombination(w, active, inversion);
        if (td) topographicDistanceOptimize(var, mp, w.pts, s, dem, nullptr);
        computeResidualsMeteo(var, mp, w.pts, s, excludeOutsideDem, true, dem);
        const float avgError = computeErrorCrossValidation(mp);
        if (!isEqualF(avgError, NODATA_F) && (isEqualF(minError, NODATA_F) || avgError < minError)) {
            minError = avgError;
            best = i;
        }
    }
    uint8_t active[PB_MAX_PROXY] = {0};
    int inversion = 0;
    if (getCombination(s, best, active, inversion)) {
        Work w{{}, &s, var};
        passDataToInterpolation(mp, w.pts, s);
        detrendingCombination(w, active, inversion);
        outPts = w.pts;
    }
}

// preInterpolation, interpolation.cpp:2444-2499, with its meteo points and current DEM
bool preInterpolationEx(Work& w, std::vector<MeteoPt>& mp, bool excludeOutsideDem, const Grid* gp, pb_kh_series* khSeries)
{
    pb_settings* s = w.s;
    const int variable = w.var;
    bool ok = true;
    if (isPrec(variable)) {
        int nrValid = 0;
        for (auto& p : w.pts)
            if (p.isActive && !isEqualF(p.value, NODATA_F) && p.value >= s->rainfallThreshold) nrValid++;
        if (nrValid == 0) { s->precipitationAllZero = 1; return true; }
        s->precipitationAllZero = 0;
    }
    if (useDetrendingVar(variable)) {
        if (s->useMultipleDetrending) {
            if (!s->useLocalDetrending) setMultipleDetrendingHeightTemperatureRange(*s);
            if (!multipleDetrendingMain(w)) ok = false;
        } else if (s->useBestDetrending) optimalDetrending(variable, mp, w.pts, *s, excludeOutsideDem, gp);
        else detrendingClassic(w);
    }
    if (ok && gp && s->useTD && !s->useLocalDetrending && useTdVar(variable) && !s->useGlocalDetrending)
        topographicDistanceOptimize(variable, mp, w.pts, *s, gp, khSeries);
    return ok;
}

std::vector<MeteoPt> loadMeteoPoints(const pb_stations* st, const pb_cv_points* cv)
{
    std::vector<MeteoPt> mp(st->n);
    for (int i = 0; i < st->n; i++) {
        mp[i].x = st->x[i]; mp[i].y = st->y[i]; mp[i].z = st->z[i];
        mp[i].currentValue = (cv && cv->currentValue) ? cv->currentValue[i] : st->value[i];
        mp[i].residual = NODATA_F;
        mp[i].lapseRateCode = st->lapseRateCode ? st->lapseRateCode[i] : PB_LAPSE_PRIMARY;
        mp[i].isInsideDem = (cv && cv->isInsideDem) ? cv->isInsideDem[i] != 0 : true;
        mp[i].proxy.assign(st->proxyValues ? st->proxyValues + size_t(i) * st->nProxy : nullptr,
                           st->proxyValues ? st->proxyValues + size_t(i + 1) * st->nProxy : nullptr);
    }
    return mp;
}

int port_pre_interpolation_ex(pb_stations* st, pb_settings* s, int variable, const pb_cv_points* cv, const pb_raster* dem,
                              uint8_t* keep, pb_kh_series* khSeries)
{
    if (khSeries) khSeries->n = 0;
    Work w{loadPoints(st), s, variable};
    std::vector<MeteoPt> mp = loadMeteoPoints(st, cv);
    Grid g{};
    const Grid* gp = nullptr;
    if (dem) { g = gridOf(*dem); gp = &g; }
    const bool ok = preInterpolationEx(w, mp, cv && cv->excludeOutsideDem, gp, khSeries);
    if (keep) memset(keep, 0, st->n);
    for (auto& p : w.pts) {
        st->value[p.index] = p.value;
        if (keep) keep[p.index] = 1;
    }
    return ok ? PB_OK : PB_ERR_FIT;
}

// sortPointsByDistance :121-160 + neighbourhoodVariability :259-301 (computeDistances with excludeSupplemental = true)
bool neighbourhoodVariability(int var, const std::vector<Point>& pts, const pb_settings& s, float x, float y, float z, int maxNrPoints,
                              const Grid* dem, float& devSt, float& avgDeltaZ, float& minDistance)
{
    const bool td = dem && s.useTD && !s.useLocalDetrending && useTdVar(var) && s.topoDist_Kh != 0;
    std::vector<float> dist(pts.size());
    for (size_t i = 0; i < pts.size(); i++) {
        float d;
        if (!checkLapseRateCode(pts[i].lapseRateCode, s.useLapseRateCode != 0, false)) d = 0;
        else {
            const float dx = float(pts[i].x) - x, dy = float(pts[i].y) - y;
            d = sqrtf(dx * dx + dy * dy);
            if (td) d += (s.topoDist_Kh * topographicDistance(x, y, z, float(pts[i].x), float(pts[i].y), float(pts[i].z), d, *dem));
        }
        dist[i] = d;
    }
    std::vector<int> idx;
    for (int i = 0; i < int(pts.size()); i++)
        if (!isEqualF(dist[i], 0) && !isEqualF(dist[i], NODATA_F)) idx.push_back(i);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return dist[a] < dist[b]; });
    const int n = std::min(maxNrPoints, int(idx.size()));
    if (n <= 1) return false;
    minDistance = dist[idx[0]];
    // statistics::standardDeviation -> variance(vector<float>) :1189-1215 with mean(vector<float>) :1300-1321
    double sum = 0;
    int nv = 0;
    for (int q = 0; q < n; q++) {
        const float v = pts[idx[q]].value;
        if (!isEqualF(v, NODATA_F)) { sum += double(v); nv++; }
    }
    const float myMean = nv ? float(sum / double(nv)) : NODATA_F;
    float squareDiff = 0;
    int nrValid = 0;
    for (int q = 0; q < n; q++) {
        const float v = pts[idx[q]].value;
        if (v != NODATA_F) {
            const float diff = v - myMean;
            squareDiff += diff * diff;
            nrValid++;
        }
    }
    const float variance = nrValid > 1 ? squareDiff / (nrValid - 1) : NODATA_F;
    devSt = isEqualF(variance, NODATA_F) ? NODATA_F : sqrtf(variance);
    if (z != NODATA_F) {
        double s2 = 0;
        int nz = 0;
        for (int q = 0; q < n; q++) {
            const double zi = pts[idx[q]].z;
            if (!isEqualD(zi, double(NODATA_F))) {
                const float dz = float(fabs(zi - z));
                if (!isEqualF(dz, NODATA_F)) { s2 += double(dz); nz++; }
            }
        }
        avgDeltaZ = nz ? float(s2 / double(nz)) : NODATA_F;
    } else avgDeltaZ = NODATA_F;
    return true;
}

// getSpatialThresholdVar, spatialControl.cpp:14-94
float getSpatialThresholdVar(int v, float rainfallThreshold, float value, float stdDev, int nrStdDev, float avgDeltaZ, float minDistance)
{
    const float PREC_THRESHOLD = 1000.f;
    float zWeight, distWeight, threshold;
    switch (v) {
    case PB_VAR_PRECIPITATION: case PB_VAR_DAILY_PRECIPITATION:
        distWeight = std::max(1.f, minDistance / 2000.f);
        if (value <= rainfallThreshold) threshold = std::max(5.f, distWeight + stdDev * (nrStdDev + 1));
        else return PREC_THRESHOLD;
        break;
    case PB_VAR_AIR_TEMPERATURE: case PB_VAR_AIR_DEW_TEMPERATURE: case PB_VAR_DAILY_TMAX: case PB_VAR_DAILY_TMIN: case PB_VAR_DAILY_TAVG:
        threshold = 1.f;
        zWeight = avgDeltaZ / 100.f;
        distWeight = minDistance / 1000.f;
        threshold = std::min(std::min(distWeight + threshold + zWeight, 12.f) + stdDev * nrStdDev, 15.f);
        break;
    case PB_VAR_AIR_REL_HUMIDITY: case PB_VAR_DAILY_RH_MAX: case PB_VAR_DAILY_RH_MIN: case PB_VAR_DAILY_RH_AVG:
        threshold = 20.f;
        zWeight = avgDeltaZ / 10.f;
        distWeight = minDistance / 1000.f;
        threshold += zWeight + distWeight + stdDev * nrStdDev;
        break;
    case PB_VAR_WIND_SCALAR_INTENSITY: case PB_VAR_WIND_VECTOR_INTENSITY: case PB_VAR_DAILY_WIND_SCALAR_INT_AVG:
    case PB_VAR_DAILY_WIND_SCALAR_INT_MAX: case PB_VAR_DAILY_WIND_VECTOR_INT_AVG: case PB_VAR_DAILY_WIND_VECTOR_INT_MAX:
        threshold = 1.f;
        zWeight = avgDeltaZ / 50.f;
        distWeight = minDistance / 2000.f;
        threshold += zWeight + distWeight + stdDev * nrStdDev;
        break;
    case PB_VAR_GLOBAL_IRRADIANCE:
        threshold = 500;
        distWeight = minDistance / 5000.f;
        threshold += distWeight + stdDev * (nrStdDev + 1);
        break;
    case PB_VAR_DAILY_GLOBAL_RADIATION:
        threshold = 10;
        distWeight = minDistance / 5000.f;
        threshold += distWeight + stdDev * (nrStdDev + 1);
        break;
    case PB_VAR_ATM_TRANSMISSIVITY:
        distWeight = minDistance / 20000.f;
        threshold = std::min(std::max(distWeight + stdDev * nrStdDev, 0.3f), 0.8f);
        break;
    case PB_VAR_ATM_PRESSURE:
        zWeight = avgDeltaZ / 10.f;
        threshold = zWeight + stdDev * nrStdDev;
        break;
    case PB_VAR_DAILY_BIC: threshold = PREC_THRESHOLD; break;
    default: threshold = stdDev * nrStdDev;
    }
    return threshold;
}

// spatialQualityControl, spatialControl.cpp:331-427 (every meteo point of st active and accepted at entry)
int port_spatial_quality_control(const pb_stations* st, pb_settings* s, int variable, const pb_cv_points* cv, const pb_raster* dem,
                                 uint8_t* wrongSpatial)
{
    const int n = st->n;
    memset(wrongSpatial, 0, n);
    std::vector<MeteoPt> all = loadMeteoPoints(st, cv);
    Grid g{};
    const Grid* gp = nullptr;
    if (dem) { g = gridOf(*dem); gp = &g; }
    const bool exclOut = cv && cv->excludeOutsideDem;
    // meteo points still accepted -> the vector the reference's loops walk (indices into `all`)
    auto accepted = [&]() {
        std::vector<int> idx;
        for (int i = 0; i < n; i++) if (!wrongSpatial[i]) idx.push_back(i);
        return idx;
    };
    auto subset = [&](const std::vector<int>& idx) {
        std::vector<MeteoPt> mp;
        for (int i : idx) mp.push_back(all[i]);
        return mp;
    };
    std::vector<int> idxA = accepted();
    std::vector<MeteoPt> mpA = subset(idxA);
    Work w{{}, s, variable};
    if (!passDataToInterpolation(mpA, w.pts, *s)) return PB_OK;
    if (!preInterpolationEx(w, mpA, exclOut, gp, nullptr)) return PB_ERR_FIT;
    computeResidualsMeteo(variable, mpA, w.pts, *s, false, false, gp);
    std::vector<int> listIndex;
    for (int i = 0; i < n; i++) {
        float stdDev, avgDeltaZ, minDist;
        if (neighbourhoodVariability(variable, w.pts, *s, float(all[i].x), float(all[i].y), float(all[i].z), 10, gp, stdDev, avgDeltaZ, minDist)) {
            const float myValue = all[i].currentValue, myResidual = mpA[i].residual;
            stdDev = std::max(stdDev, myValue / 100.f);
            if (fabs(myResidual) > getSpatialThresholdVar(variable, s->rainfallThreshold, myValue, stdDev, 2, avgDeltaZ, minDist)) {
                listIndex.push_back(i);
                wrongSpatial[i] = 1;
            }
        }
    }
    if (listIndex.empty()) return PB_OK;
    std::vector<int> idxB = accepted();
    std::vector<MeteoPt> mpB = subset(idxB);
    Work w2{{}, s, variable};
    if (!passDataToInterpolation(mpB, w2.pts, *s)) return PB_OK;
    if (!preInterpolationEx(w2, mpB, exclOut, gp, nullptr)) return PB_ERR_FIT;
    std::vector<float> listResiduals;
    for (size_t k = 0; k < listIndex.size(); k++) {
        const MeteoPt& m = all[listIndex[k]];
        double pv[PB_MAX_PROXY];
        // meteoPoints[i].getProxyValues() with i = position in the list (spatialControl.cpp:393)
        for (int q = 0; q < PB_MAX_PROXY; q++) pv[q] = q < int(all[k].proxy.size()) ? double(all[k].proxy[q]) : -9999.;
        const float est = interpolate(w2.pts, *s, variable, float(m.x), float(m.y), pv, false, float(m.z), gp);
        listResiduals.push_back(est - m.currentValue);
    }
    for (size_t k = 0; k < listIndex.size(); k++) {
        const int i = listIndex[k];
        float stdDev, avgDeltaZ, minDist;
        if (neighbourhoodVariability(variable, w2.pts, *s, float(all[i].x), float(all[i].y), float(all[i].z), 10, gp, stdDev, avgDeltaZ, minDist))
            wrongSpatial[i] = fabs(listResiduals[k]) > getSpatialThresholdVar(variable, s->rainfallThreshold, all[i].currentValue, stdDev, 3,
                                                                             avgDeltaZ, minDist) ? 1 : 0;
        else wrongSpatial[i] = 0;
    }
    return PB_OK;
}

int port_interpolate_points(const pb_stations* st, const pb_settings* s, int variable, const float* x, const float* y,
                            const float*, const double* proxyValues, int m, int excludeSupplemental, float* result)
{
    std::vector<Point> pts = loadPoints(st);
    double none[PB_MAX_PROXY];
    for (auto& v : none) v = -9999;
    for (int i = 0; i < m; i++)
        result[i] = interpolate(pts, *s, variable, x[i], y[i], proxyValues ? proxyValues + size_t(i) * s->nProxy : none,
                                excludeSupplemental != 0);
    return PB_OK;
}

int port_interpolate_points_td(const pb_stations* st, const pb_settings* s, int variable, const float* x, const float* y,
                               const float* z, const double* proxyValues, int m, int excludeSupplemental, const pb_raster* dem,
                               float* result)
{
    std::vector<Point> pts = loadPoints(st);
    Grid d = gridOf(*dem);
    double none[PB_MAX_PROXY];
    for (auto& v : none) v = -9999;
    for (int i = 0; i < m; i++)
        result[i] = interpolate(pts, *s, variable, x[i], y[i], proxyValues ? proxyValues + size_t(i) * s->nProxy : none,
                                excludeSupplemental != 0, z[i], &d);
    return PB_OK;
}

// Project::computeStatisticsCrossValidation, project/project.cpp:2935-2969 (statistics.cpp:180-268, 1030-1092)
void cvStatistics(const std::vector<float>& obs, const std::vector<float>& pre, pb_cv_stats* stats)
{
    if (!stats) return;
    memset(stats, 0, sizeof(*stats));
    stats->n = int(obs.size());
    stats->meanAbsoluteError = stats->meanBiasError = stats->rootMeanSquareError = stats->nashSutcliffe = stats->R2 = NODATA_F;
    if (!obs.empty()) {
        double sumAbs = 0, sigma = 0;
        float sumErr = 0.f;
        long n = 0;
        for (size_t i = 0; i < obs.size(); i++)
            if (obs[i] != NODATA_F && pre[i] != NODATA_F) {
                sumAbs += fabs(obs[i] - pre[i]);
                sumErr += obs[i] - pre[i];
                const double diff = obs[i] - pre[i];
                sigma += diff * diff;
                n++;
            }
        if (n) {
            stats->meanAbsoluteError = float(sumAbs / n);
            stats->meanBiasError = sumErr / n;
            stats->rootMeanSquareError = float(sqrt(sigma / n));
        }
        // NashSutcliffeEfficiency (statistics.cpp:243-268): float accumulators
        double s0 = 0; int nv = 0;
        for (float v : obs) if (!isEqualF(v, NODATA_F)) { s0 += double(v); nv++; }
        if (nv) {
            const float obsAvg = float(s0 / double(nv));
            float sumDev = 0, sumError = 0;
            for (size_t i = 0; i < obs.size(); i++) if (!isEqualF(obs[i], NODATA_F)) sumDev += (obs[i] - obsAvg) * (obs[i] - obsAvg);
            if (sumDev != 0) {
                for (size_t i = 0; i < obs.size(); i++)
                    if (!isEqualF(obs[i], NODATA_F) && !isEqualF(pre[i], NODATA_F)) sumError += (pre[i] - obs[i]) * (pre[i] - obs[i]);
                stats->nashSutcliffe = float(1 - (sumError / sumDev));
            }
        }
        float q, m, r2;
        linearRegression(obs.data(), pre.data(), long(obs.size()), false, &q, &m, &r2);
        stats->R2 = r2;
    }
}

// computeResiduals, interpolation/spatialControl.cpp:97-155 + Project::computeStatisticsCrossValidation, project/project.cpp:2935-2969
// with statistics::meanAbsoluteError / meanError / rootMeanSquareError / NashSutcliffeEfficiency, mathFunctions/statistics.cpp:180-268
int port_compute_residuals(const pb_stations* st, const pb_settings* s, int variable, const float* observed,
                           const uint8_t* useForCv, float* residual, pb_cv_stats* stats)
{
    std::vector<Point> pts = loadPoints(st);
    std::vector<float> obs, pre;
    for (int i = 0; i < st->n; i++) {
        residual[i] = NODATA_F;
        if (useForCv && !useForCv[i]) continue;
        double pv[PB_MAX_PROXY];
        for (int k = 0; k < PB_MAX_PROXY; k++) pv[k] = k < st->nProxy ? double(pts[i].proxy[k]) : -9999.;
        float myValue = observed[i];
        float est = interpolate(pts, *s, variable, float(st->x[i]), float(st->y[i]), pv, false);
        if (isPrec(variable)) {
            if (myValue != NODATA_F && myValue < s->rainfallThreshold) myValue = 0.f;
            if (est != NODATA_F && est < s->rainfallThreshold) est = 0.f;
        }
        if (est != NODATA_F && myValue != NODATA_F) residual[i] = myValue - est;
        const float value = observed[i];
        if (!isEqualF(value, NODATA_F) && !isEqualF(residual[i], NODATA_F)) {
            obs.push_back(value);
            pre.push_back(value - residual[i]);
        }
    }
    cvStatistics(obs, pre, stats);
    return PB_OK;
}

// loop of Project::interpolationDemLocalDetrending, project/project.cpp:3208-3253 (+ updateMinMaxRasterGrid)
int port_local_detrending_raster(const pb_stations* st, const pb_settings* s0, int variable, const pb_raster* dem,
                                 const pb_raster* proxyGrids, int nProxyGrids, int nThreads, int rowBegin, int rowEnd,
                                 pb_raster_out* out, double* elapsed_s)
{
    if (!useDetrendingVar(variable) || !s0->useLocalDetrending) return PB_ERR_INVALID;
    std::vector<Point> pts = loadPoints(st);
    Grid d = gridOf(*dem);
    std::vector<Grid> grids;
    for (int g = 0; g < nProxyGrids; g++) grids.push_back(gridOf(proxyGrids[g]));
    const int rows = d.h.nrRows, cols = d.h.nrCols;
    if (rowEnd < 0) rowEnd = rows;
    if (nThreads > 0) omp_set_num_threads(nThreads);
    pb_settings base = *s0;
    for (int i = 0; i < base.nProxy; i++) { base.currentActive[i] = base.selectedActive[i]; base.currentSignificant[i] = 0; }
    base.currentUseThermalInversion = base.selectedUseThermalInversion;
    if (!setMultipleDetrendingHeightTemperatureRange(base)) return PB_ERR_FIT;
    for (int row = 0; row < rows; row++) {
        float* orow = out->data ? out->data + size_t(row) * cols : out->rows[row];
        for (int col = 0; col < cols; col++) orow[col] = d.h.flag;
    }
    auto t0 = std::chrono::steady_clock::now();
    int64_t nValid = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : nValid)
    for (int row = rowBegin; row < rowEnd; row++) {
        pb_settings s = base;
        double pv[PB_MAX_PROXY];
        float* orow = out->data ? out->data + size_t(row) * cols : out->rows[row];
        for (int col = 0; col < cols; col++) {
            const float z = d.at(row, col);
            if (isEqualF(z, d.h.flag)) continue;
            nValid++;
            const double x = d.h.llx + d.h.cellSize * (col + 0.5);            // getUtmXYFromRowCol, gis.cpp:806-810
            const double y = d.h.lly + d.h.cellSize * (rows - row - 0.5);
            for (int i = 0; i < s.nProxy; i++) {                              // getProxyValuesXY(float x, float y, ...)
                pv[i] = -9999;
                if (!s.currentActive[i]) continue;
                const int gi = s.proxy[i].gridIndex;
                if (gi < 0 || gi >= nProxyGrids) continue;
                const float v = grids[gi].valueFromXY(float(x), float(y));
                if (v != grids[gi].h.flag) pv[i] = v;
            }
            orow[col] = localCell(pts, s, variable, x, y, pv, nullptr);
            s.nFit = s.nFitFunctions = 0;                                     // clearFitting + setCurrentCombination(selected)
            for (int i = 0; i < s.nProxy; i++) { s.currentActive[i] = s.selectedActive[i]; s.currentSignificant[i] = 0; }
            s.currentUseThermalInversion = s.selectedUseThermalInversion;
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (elapsed_s) *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
    bool first = true;
    float mn = NODATA_F, mx = NODATA_F;
    for (int row = 0; row < rows; row++) {
        const float* orow = out->data ? out->data + size_t(row) * cols : out->rows[row];
        for (int col = 0; col < cols; col++) {
            const float v = orow[col];
            if (!isEqualF(v, d.h.flag) && !isEqualF(v, NODATA_F)) {
                if (first) { mn = mx = v; first = false; }
                else if (v < mn) mn = v;
                else if (v > mx) mx = v;
            }
        }
    }
    out->minimum = mn; out->maximum = mx; out->nValid = nValid;
    return first ? PB_ERR_NO_VALID_CELL : PB_OK;
}

// computeResidualsLocalDetrending, spatialControl.cpp:158-228, called as in Project::interpolationCv (project.cpp:3086-3097)
int port_compute_residuals_local(const pb_stations* st, const pb_settings* s0, int variable, const float* observed,
                                 const uint8_t* useForCv, float* residual, pb_cv_stats* stats)
{
    std::vector<Point> pts = loadPoints(st);
    pb_settings s = *s0;
    for (int i = 0; i < s.nProxy; i++) { s.currentActive[i] = s.selectedActive[i]; s.currentSignificant[i] = 0; }
    s.currentUseThermalInversion = s.selectedUseThermalInversion;
    if (!setMultipleDetrendingHeightTemperatureRange(s)) return PB_ERR_FIT;
    std::vector<float> obs, pre;
    for (int i = 0; i < st->n; i++) {
        residual[i] = NODATA_F;
        if (useForCv && !useForCv[i]) continue;
        if (!checkLapseRateCode(pts[i].lapseRateCode, s.useLapseRateCode != 0, false)) continue;   // excludeSupplemental = true
        double pv[PB_MAX_PROXY];
        for (int k = 0; k < PB_MAX_PROXY; k++) pv[k] = k < st->nProxy ? double(pts[i].proxy[k]) : -9999.;
        float myValue = observed[i];
        Work w{{}, &s, variable};
        if (!localSelection(pts, w.pts, float(st->x[i]), float(st->y[i]), s, false)) return PB_ERR_INVALID;
        if (!preInterpolation(w)) return PB_ERR_FIT;
        float est = interpolate(w.pts, s, variable, float(st->x[i]), float(st->y[i]), pv, false);
        if (isPrec(variable)) {
            if (myValue != NODATA_F && myValue < s.rainfallThreshold) myValue = 0.f;
            if (est != NODATA_F && est < s.rainfallThreshold) est = 0.f;
        }
        if (est != NODATA_F && myValue != NODATA_F) residual[i] = myValue - est;
    }
    for (int i = 0; i < st->n; i++) {
        if (useForCv && !useForCv[i]) continue;
        const float value = observed[i];
        if (!isEqualF(value, NODATA_F) && !isEqualF(residual[i], NODATA_F)) {
            obs.push_back(value);
            pre.push_back(value - residual[i]);
        }
    }
    cvStatistics(obs, pre, stats);
    return PB_OK;
}

int port_local_detrending_point(const pb_stations* st, pb_settings* s, int variable, double x, double y, float,
                                const double* proxyValuesIn, int* selIndex, float* selWeight, float* selValue, int* nSel,
                                float* result)
{
    std::vector<Point> pts = loadPoints(st);
    for (int i = 0; i < s->nProxy; i++) { s->currentActive[i] = s->selectedActive[i]; s->currentSignificant[i] = 0; }
    s->currentUseThermalInversion = s->selectedUseThermalInversion;
    if (!setMultipleDetrendingHeightTemperatureRange(*s)) return PB_ERR_FIT;
    std::vector<Point> subset;
    *result = localCell(pts, *s, variable, x, y, proxyValuesIn, &subset);
    *nSel = int(subset.size());
    for (size_t k = 0; k < subset.size(); k++) {
        selIndex[k] = subset[k].index; selWeight[k] = subset[k].regressionWeight; selValue[k] = subset[k].value;
    }
    return PB_OK;
}

// kriging, interpolation/kriging.cpp: krigingVariogram :97-202, matrixInversion :28-81, krigingSetWeight :211-264,
// krigingResult :271-279, krigingRMSE :287-295
// ---- raster map operators around the gridding path ------------------------------------------------
// updateMinMaxRasterGrid, gis/gis.cpp:569-614
static bool minMaxOf(const float* v, size_t n, float flag, float* mn, float* mx)
{
    bool first = true;
    *mn = *mx = NODATA_F;
    for (size_t i = 0; i < n; i++)
        if (!isEqualF(v[i], flag) && !isEqualF(v[i], NODATA_F)) {
            if (first) { *mn = *mx = v[i]; first = false; }
            else if (v[i] < *mn) *mn = v[i];
            else if (v[i] > *mx) *mx = v[i];
        }
    return !first;
}
static float* outRow(pb_raster_out* out, int row, int cols) { return out->data ? out->data + size_t(row) * cols : out->rows[row]; }

// relHumFromTdew(float, float), meteo/meteo.cpp:301-315
static float relHumFromTdew(float Td, float T)
{
    if (isEqualF(Td, NODATA_F) || isEqualF(T, NODATA_F)) return NODATA_F;
    double d = 237.3;
    double c = 17.2693882;
    double esp = 1 / (double(T) + d);
    double rh = pow(exp((c * double(Td)) - ((c * double(T) / (double(T) + d))) * (double(Td) + d)), esp);
    rh *= 100;
    if (rh > 100) return 100;
    return std::max(1.f, float(rh));
}

// tDewFromRelHum(float, float), meteo/meteo.cpp:275-285
void port_tdew_from_relhum(const float* rh, const float* t, int n, float* tdew)
{
    for (int i = 0; i < n; i++) {
        float RH = rh[i];
        const float T = t[i];
        if (isEqualF(RH, NODATA_F) || isEqualF(T, NODATA_F) || RH == 0) { tdew[i] = NODATA_F; continue; }
        RH = (100 < RH) ? 100 : RH;
        const double mySaturatedVaporPres = exp((16.78 * double(T) - 116.9) / (double(T) + 237.3));
        const double actualVaporPres = double(RH) / 100. * mySaturatedVaporPres;
        tdew[i] = float((log(actualVaporPres) * 237.3 + 116.9) / (16.78 - log(actualVaporPres)));
    }
}

// Crit3DHourlyMeteoMaps::computeRelativeHumidityMap, project/meteoMaps.cpp:301-321
int port_relative_humidity_map(const pb_raster* airT, const pb_raster* dewT, pb_raster_out* out)
{
    Grid t = gridOf(*airT), td = gridOf(*dewT);
    const int rows = t.h.nrRows, cols = t.h.nrCols;
    std::vector<float> flat(size_t(rows) * cols);
    for (int row = 0; row < rows; row++)
        for (int col = 0; col < cols; col++) {
            float v = t.h.flag;                                      // emptyGrid()
            const float a = t.at(row, col), b = td.at(row, col);
            if (!isEqualF(a, t.h.flag) && !isEqualF(b, td.h.flag)) v = relHumFromTdew(b, a);
            flat[size_t(row) * cols + col] = v;
            outRow(out, row, cols)[col] = v;
        }
    out->nValid = int64_t(rows) * cols;
    return minMaxOf(flat.data(), flat.size(), t.h.flag, &out->minimum, &out->maximum) ? PB_OK : PB_ERR_NO_VALID_CELL;
}

// Crit3DDailyMeteoMaps::fixDailyThermalConsistency, project/meteoMaps.cpp:103-126
int port_fix_daily_thermal_consistency(const pb_raster* tmax, const pb_raster* tmin, float* tminOut, int64_t* nFixed)
{
    Grid a = gridOf(*tmax), b = gridOf(*tmin);
    int64_t n = 0;
    for (int row = 0; row < a.h.nrRows; row++)
        for (int col = 0; col < a.h.nrCols; col++) {
            float v = b.at(row, col);
            if (!isEqualF(a.at(row, col), a.h.flag) && !isEqualF(v, b.h.flag) && v > a.at(row, col)) {
                v = a.at(row, col) - 0.1f;
                n++;
            }
            tminOut[size_t(row) * a.h.nrCols + col] = v;
        }
    if (nFixed) *nFixed = n;
    return PB_OK;
}

// gis::topographicDistanceMap, gis/gis.cpp:1648-1672
int port_topographic_distance_map(const pb_raster* dem, double x, double y, double z, pb_raster_out* out)
{
    Grid d = gridOf(*dem);
    const int rows = d.h.nrRows, cols = d.h.nrCols;
#pragma omp parallel for schedule(dynamic, 4)
    for (int row = 0; row < rows; row++)
        for (int col = 0; col < cols; col++) {
            const float demValue = d.at(row, col);
            float v = d.h.flag;
            if (!isEqualF(demValue, d.h.flag)) {
                const double gx = d.h.llx + d.h.cellSize * (double(col) + 0.5);          // getXY, gis.cpp:473-477
                const double gy = d.h.lly + d.h.cellSize * (double(rows - row) - 0.5);
                const float dx = float(gx) - float(x), dy = float(gy) - float(y);          // computeDistance, gis.cpp:685-691
                const float distance = sqrtf(dx * dx + dy * dy);
                v = topographicDistance(float(gx), float(gy), demValue, float(x), float(y), float(z), distance, d);
            }
            outRow(out, row, cols)[col] = v;
        }
    out->nValid = int64_t(rows) * cols;
    out->minimum = out->maximum = NODATA_F;
    return PB_OK;
}

// gis::resampleGrid, gis/gis.cpp:1722-1811 (aggrAverage, aggrMedian, aggrCenter)
int port_resample_grid(const pb_raster* src, const pb_raster_header* nh, int method, float nodataRatioThreshold, pb_raster_out* out)
{
    if (method != PB_AGGR_AVERAGE && method != PB_AGGR_MEDIAN && method != PB_AGGR_CENTER && method != PB_AGGR_PREVAILING)
        return PB_ERR_UNSUPPORTED;
    Grid o = gridOf(*src);
    const int rows = nh->nrRows, cols = nh->nrCols;
    const double resampleFactor = nh->cellSize / o.h.cellSize;
    std::vector<float> flat(size_t(rows) * cols);
    for (int row = 0; row < rows; row++)
        for (int col = 0; col < cols; col++) {
            float value = NODATA_F;
            const double x0 = nh->llx + nh->cellSize * (double(col) + 0.5);
            const double y0 = nh->lly + nh->cellSize * (double(rows - row) - 0.5);
            if (resampleFactor <= 1. || method == PB_AGGR_CENTER) {
                const float tmp = o.valueFromXY(x0, y0);                // getRowCol + isOutOfGridRowCol + value
                if (!isEqualF(tmp, o.h.flag)) value = tmp;
            } else {
                const int nrStep = int(floor(resampleFactor)) + 1;
                const double step = nh->cellSize / nrStep, halfStep = step * 0.5;
                const double llx = x0 - (nh->cellSize / 2) + halfStep, lly = y0 - (nh->cellSize / 2) + halfStep;
                std::vector<float> values;
                int maxNrValues = 0;
                for (int i = 0; i < nrStep; i++) {
                    const double x = llx + step * i;
                    for (int j = 0; j < nrStep; j++) {
                        const double y = lly + step * j;
                        const float tmp = o.valueFromXY(x, y);
                        if (!isEqualF(tmp, o.h.flag)) values.push_back(tmp);
                        maxNrValues++;
                    }
                }
                const int nrValues = int(values.size());
                if (maxNrValues > 0 && (float(nrValues) / float(maxNrValues)) > nodataRatioThreshold) {
                    if (method == PB_AGGR_AVERAGE) {                    // statistics::mean, statistics.cpp:1300-1321
                        double sum = 0;
                        int nv = 0;
                        for (float v : values) if (!isEqualF(v, NODATA_F)) { sum += double(v); nv++; }
                        value = (values.empty() || nv == 0) ? NODATA_F : float(sum / double(nv));
                    } else if (method == PB_AGGR_PREVAILING) {          // gis::prevailingValue, gis.cpp:1513-1554 (:1794-1799)
                        if (maxNrValues - nrValues < nrValues) {
                            std::vector<float> distinct;
                            std::vector<unsigned> counters;
                            distinct.push_back(values[0]);
                            counters.push_back(1);
                            for (size_t i = 1; i < values.size(); i++) {
                                bool found = false;
                                for (size_t j = 0; j < distinct.size() && !found; j++)
                                    if (isEqualF(values[i], distinct[j])) { found = true; counters[j]++; }
                                if (!found) { distinct.push_back(values[i]); counters.push_back(1); }
                            }
                            unsigned maxCount = counters[0];
                            value = distinct[0];
                            for (size_t i = 1; i < distinct.size(); i++)
                                if (counters[i] > maxCount) { maxCount = counters[i]; value = distinct[i]; }
                        }
                    } else {                                           // sorting::percentile(values, n, 50, true), basicMath.cpp:327-366
                        if (values.size() >= 3) {
                            std::vector<float> list;
                            for (float v : values) if (!isEqualF(v, NODATA_F)) list.push_back(v);
                            std::sort(list.begin(), list.end());
                            const int n = int(list.size());
                            if (n >= 3) {
                                const double rank = n * double(50.0 / 100.f) - 1.0;
                                if (rank < 0.0) value = list[0];
                                else {
                                    const int low = int(rank);
                                    if (low >= n - 1) value = list[n - 1];
                                    else value = float(list[low] + (rank - low) * (list[low + 1] - list[low]));
                                }
                            }
                        }
                    }
                }
            }
            const float v = isEqualF(value, NODATA_F) ? nh->flag : value;
            flat[size_t(row) * cols + col] = v;
            outRow(out, row, cols)[col] = v;
        }
    out->nValid = int64_t(rows) * cols;
    minMaxOf(flat.data(), flat.size(), nh->flag, &out->minimum, &out->maximum);
    return PB_OK;
}

int port_kriging_points(const double* pos, const double* val, int n, int mode, double range, double nugget, double sill,
                        double slope, const double* xy, int m, double* result, double* rmse, double* setup_s, double* eval_s)
{
    auto t0 = std::chrono::steady_clock::now();
    const int dim = n + 1;
    const double sill_nugget = sill - nugget;
    std::vector<double> V(size_t(dim) * dim, 0.), inv(size_t(dim) * dim, 0.), D(dim), weight(dim);
    auto gamma = [&](double h) {
        const double tmp = h / range;
        switch (mode) {
        case 1: return h < range ? nugget + sill_nugget * (1.5 * tmp - 0.5 * tmp * tmp * tmp) : nugget + sill_nugget;
        case 2: return nugget + sill_nugget * (1. - exp(-3. * tmp));
        case 3: return nugget + sill_nugget * (1. - exp(-4. * tmp * tmp));
        default: return nugget + slope * h;
        }
    };
    for (int i = 0; i < n; i++) { V[size_t(i) * dim + dim - 1] = 1; V[size_t(dim - 1) * dim + i] = 1; }
    V[size_t(dim - 1) * dim + dim - 1] = 0;
    for (int i = 0; i < n; i++)
        for (int j = i; j < n; j++) {
            const double dx = pos[i * 2] - pos[j * 2], dy = pos[i * 2 + 1] - pos[j * 2 + 1];
            V[size_t(i) * dim + j] = V[size_t(j) * dim + i] = gamma(sqrt(dx * dx + dy * dy));
        }
    // Gauss-Jordan without pivoting, the reference's exact elimination order
    double* A = V.data();
    double* I = inv.data();
    for (int i = 0; i < dim; i++) I[size_t(i) * dim + i] = 1.;
    for (int iter = 0; iter < dim - 1; iter++) {
        const int first_row = iter + 1;
        for (int i = first_row; i < dim; i++) {
            const double r = A[size_t(i) * dim + iter] / A[size_t(iter) * dim + iter];
            A[size_t(i) * dim + iter] = 0.;
            for (int j = first_row; j < dim; j++) A[size_t(i) * dim + j] -= r * A[size_t(iter) * dim + j];
            for (int j = 0; j <= iter; j++) I[size_t(i) * dim + j] -= r * I[size_t(iter) * dim + j];
        }
    }
    for (int iter = dim - 1; iter > 0; iter--)
        for (int i = iter - 1; i >= 0; i--) {
            const double r = A[size_t(i) * dim + iter] / A[size_t(iter) * dim + iter];
            A[size_t(i) * dim + iter] = 0.;
            for (int j = 0; j < dim; j++) I[size_t(i) * dim + j] -= r * I[size_t(iter) * dim + j];
        }
    for (int i = 0; i < dim; i++) for (int j = 0; j < dim; j++) I[size_t(i) * dim + j] /= A[size_t(i) * dim + i];
    auto t1 = std::chrono::steady_clock::now();
    for (int p = 0; p < m; p++) {
        for (int i = 0; i < n; i++) {
            const double dx = pos[i * 2] - xy[2 * p], dy = pos[i * 2 + 1] - xy[2 * p + 1];
            D[i] = gamma(sqrt(dx * dx + dy * dy));
        }
        D[n] = 1;
        double res = 0.0, err = 0.0;
        for (int i = 0; i < n; i++) {
            double wi = 0;
            for (int j = 0; j < dim; j++) wi += I[size_t(i) * dim + j] * D[j];
            weight[i] = wi;
        }
        for (int i = 0; i < n; i++) res += weight[i] * val[i];
        for (int i = 0; i < n; i++) err += weight[i] * D[i];
        result[p] = res;
        if (rmse) rmse[p] = sqrt(fabs(err));
    }
    auto t2 = std::chrono::steady_clock::now();
    if (setup_s) *setup_s = std::chrono::duration<double>(t1 - t0).count();
    if (eval_s) *eval_s = std::chrono::duration<double>(t2 - t1).count();
    return PB_OK;
}

}  // extern "C"

// ---- glocal detrending entry points ------------------------------------------------------------------------------
extern "C" {

// preInterpolation :2444-2499 with useGlocalDetrending: height ranges, then glocalDetrendingFitting
int port_glocal_detrending_fitting(const pb_stations* st, pb_settings* s, int variable, pb_macro_area* areas, int nAreas)
{
    std::vector<Point> pts = loadPoints(st);
    if (!useDetrendingVar(variable) || !s->useMultipleDetrending) return PB_OK;
    if (!s->useLocalDetrending) setMultipleDetrendingHeightTemperatureRange(*s);
    glocalDetrendingFitting(pts, *s, variable, areas, nAreas);        // the result is ignored by preInterpolation (:2473)
    return PB_OK;
}

// the gridding part of Project::interpolationDemGlocalDetrending, project/project.cpp:3283-3375 (areas in index order)
int port_glocal_detrending_raster(const pb_stations* st, pb_settings* s, int variable, pb_macro_area* areas, int nAreas,
                                  const pb_raster* dem, const pb_raster* proxyGrids, int nProxyGrids, int, pb_raster_out* out,
                                  double* elapsed_s)
{
    if (!useDetrendingVar(variable) || !s->useGlocalDetrending) return PB_ERR_INVALID;
    std::vector<Point> pts = loadPoints(st);
    Grid d = gridOf(*dem);
    std::vector<Grid> grids;
    for (int g = 0; g < nProxyGrids; g++) grids.push_back(gridOf(proxyGrids[g]));
    const int rows = d.h.nrRows, cols = d.h.nrCols;
    auto t0 = std::chrono::steady_clock::now();
    for (int p = 0; p < s->nProxy; p++) { s->currentActive[p] = s->selectedActive[p]; s->currentSignificant[p] = 0; }
    int rc = port_glocal_detrending_fitting(st, s, variable, areas, nAreas);
    if (rc) return rc;
    bool isOk = true;
    for (int row = 0; row < rows; row++) {
        float* orow = out->data ? out->data + size_t(row) * cols : out->rows[row];
        for (int col = 0; col < cols; col++) orow[col] = d.h.flag;           // initializeGrid(header)
    }
    if (!setMultipleDetrendingHeightTemperatureRange(*s)) return PB_ERR_FIT;
    int elevationPos = -9999;
    for (int pos = 0; pos < s->nProxy; pos++) if (s->proxy[pos].kind == PB_PROXY_HEIGHT) elevationPos = pos;
    pb_settings my = *s;
    for (int a = 0; a < nAreas && isOk; a++) {
        const pb_macro_area& A = areas[a];
        if (A.nAreaCells <= 0) continue;
        std::vector<Point> subset;
        macroAreaDetrending(A, my, variable, pts, subset, elevationPos);
        for (int64_t c = 0; c + 1 < A.nAreaCells; c += 2) {
            if (!isOk) break;
            const int row = int(A.areaCellsDEM[c] / cols);
            const int col = int(A.areaCellsDEM[c]) % cols;
            const float z = d.at(row, col);
            if (isEqualF(z, d.h.flag)) continue;
            const double x = d.h.llx + d.h.cellSize * (col + 0.5);              // getUtmXYFromRowCol, gis.cpp:806-810
            const double y = d.h.lly + d.h.cellSize * (rows - row - 0.5);
            float* cell = (out->data ? out->data + size_t(row) * cols : out->rows[row]) + col;
            double pv[PB_MAX_PROXY];
            if (!getSignificantProxyValuesXY(float(x), float(y), my, grids, pv)) { *cell = NODATA_F; continue; }
            const double v = interpolate(subset, my, variable, float(x), float(y), pv, true, z, nullptr);
            if (isEqualD(v, -9999)) isOk = false;
            if (isEqualF(*cell, NODATA_F)) *cell = float(v * A.areaCellsDEM[c + 1]);
            else *cell += v * A.areaCellsDEM[c + 1];
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (elapsed_s) *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
    int64_t nValid = 0;
    bool first = true;
    float mn = NODATA_F, mx = NODATA_F;
    for (int row = 0; row < rows; row++) {
        const float* orow = out->data ? out->data + size_t(row) * cols : out->rows[row];
        for (int col = 0; col < cols; col++) {
            if (!isEqualF(d.at(row, col), d.h.flag)) nValid++;
            const float v = orow[col];
            if (!isEqualF(v, d.h.flag) && !isEqualF(v, NODATA_F)) {
                if (first) { mn = mx = v; first = false; }
                else if (v < mn) mn = v;
                else if (v > mx) mx = v;
            }
        }
    }
    out->minimum = mn; out->maximum = mx; out->nValid = nValid;
    return isOk ? PB_OK : PB_ERR_NO_VALID_CELL;
}

// Project::computeResidualsAndStatisticsGlocalDetrending (project.cpp:6525-6565) over computeResidualsGlocalDetrending
// (spatialControl.cpp:231-302)
int port_compute_residuals_glocal(const pb_stations* meteo, const float* observed, const pb_cv_points* cv, const pb_stations* points,
                                  pb_settings* s, int variable, const pb_macro_area* areas, int nAreas, const pb_raster*,
                                  int excludeOutsideDem, int excludeSupplemental, float* residual)
{
    std::vector<Point> pts = loadPoints(points);
    std::vector<MeteoPt> mp = loadMeteoPoints(meteo, cv);
    for (int i = 0; i < meteo->n; i++) {
        if (observed) mp[i].currentValue = observed[i];
        mp[i].active = meteo->isActive ? meteo->isActive[i] != 0 : true;
        mp[i].residual = NODATA_F;
    }
    if (!setMultipleDetrendingHeightTemperatureRange(*s)) return PB_ERR_FIT;
    int elevationPos = -9999;
    for (int pos = 0; pos < s->nProxy; pos++) if (s->proxy[pos].kind == PB_PROXY_HEIGHT) elevationPos = pos;
    for (int k = 0; k < nAreas; k++) {
        const pb_macro_area& A = areas[k];
        if (A.nAreaCells <= 0 || A.nMeteoPoints <= 0) continue;
        std::vector<Point> subset;
        macroAreaDetrending(A, *s, variable, pts, subset, elevationPos);
        for (int i = 0; i < A.nMeteoPoints; i++) {
            MeteoPt& m = mp[A.meteoPoints[i]];
            if (!m.active) continue;
            const float weight = 1;
            bool isValid = !excludeSupplemental || checkLapseRateCode(m.lapseRateCode, s->useLapseRateCode != 0, false);
            isValid = isValid && (!excludeOutsideDem || m.isInsideDem);
            if (!isValid) continue;
            double pv[PB_MAX_PROXY];
            for (int q = 0; q < PB_MAX_PROXY; q++) pv[q] = q < int(m.proxy.size()) ? double(m.proxy[q]) : -9999.;
            const float myValue = m.currentValue;
            const float est = interpolate(subset, *s, variable, float(m.x), float(m.y), pv, false, float(m.z), nullptr);
            if (!isEqualF(est, NODATA_F) && !isEqualF(myValue, NODATA_F)) {
                if (isEqualF(m.residual, NODATA_F)) m.residual = (myValue - est) * weight;
                else m.residual += (myValue - est) * weight;
            }
        }
    }
    for (int i = 0; i < meteo->n; i++) residual[i] = mp[i].residual;
    return PB_OK;
}

}  // extern "C"

// Crit3DMeteoGrid::spatialAggregateMeteoGrid (meteo/meteoGrid.cpp:1135-1190) + spatialAggregateMeteoGridPoint (:1193-1240) with
// statistics::mean / variance / standardDeviation (mathFunctions/statistics.cpp:1274-1297, 1160-1186, 1373-1381) and
// sorting::percentile (basicMath.cpp:327-366); fresh aggregation points (z = NODATA)
extern "C" int port_spatial_aggregate_meteo_grid(const pb_raster* raster, const double* x, const double* y, const int64_t* first,
                                                 const int32_t* maxNr, int nCells, int method, float* value)
{
    const Grid g = gridOf(*raster);
    for (int c = 0; c < nCells; c++) {
        value[c] = NODATA_F;
        const int64_t n = first[c + 1] - first[c];
        std::vector<double> z(size_t(n), -9999.);
        double validValues = 0;
        for (int64_t i = 0; i < n; i++) {
            const float v = g.valueFromXY(x[first[c] + i], y[first[c] + i]);
            if (!isEqualF(v, g.h.flag)) { z[size_t(i)] = double(v); validValues = validValues + 1; }
        }
        if (n == 0) continue;
        if (!((validValues / maxNr[c]) > (0 / 100))) continue;
        std::vector<float> vals;
        for (double q : z) if (q != -9999) vals.push_back(float(q));
        double myValue = -9999;
        if (!vals.empty() && !((double(vals.size()) / maxNr[c]) < (0 / 100.0))) {
            if (method == PB_AGGR_AVERAGE) {
                float sum = 0.;
                int nv = 0;
                for (float v : vals) if (v != NODATA_F) { sum += v; nv++; }
                myValue = nv > 0 ? sum / float(nv) : NODATA_F;
            } else if (method == PB_AGGR_STDDEV) {
                float var = NODATA_F;
                if (vals.size() > 1) {
                    float sum = 0.;
                    int nv = 0;
                    for (float v : vals) if (v != NODATA_F) { sum += v; nv++; }
                    const float mean = nv > 0 ? sum / float(nv) : NODATA_F;
                    float squareDiff = 0;
                    int m = 0;
                    for (float v : vals) if (v != NODATA_F) { const float d = v - mean; squareDiff += d * d; m++; }
                    if (m > 1) var = squareDiff / (m - 1);
                }
                myValue = isEqualF(var, NODATA_F) ? NODATA_F : sqrtf(var);
            } else if (method == PB_AGGR_MEDIAN) {
                float r = NODATA_F;
                if (vals.size() >= 3) {
                    std::vector<float> l;
                    for (float v : vals) if (!isEqualF(v, NODATA_F)) l.push_back(v);
                    std::sort(l.begin(), l.end());
                    const int nl = int(l.size());
                    if (nl >= 3) {
                        const double rank = nl * double(50.0 / 100.f) - 1.0;
                        if (rank < 0.0) r = l[0];
                        else {
                            const int low = int(rank);
                            if (low >= nl - 1) r = l[nl - 1];
                            else { const double frac = rank - low; r = float(l[low] + frac * (l[low + 1] - l[low])); }
                        }
                    }
                }
                myValue = r;
            }
        }
        value[c] = float(myValue);
    }
    return PB_OK;
}

// topographicIndex, project/interpolationCmd.cpp:397-482: per window width, the weighted share of lower minus higher (minus half
// the equal) neighbours inside a disc of cellNr cells, the centre's own row and column excluded (:438); widths add up
extern "C" int port_topographic_index(const pb_raster* dem, const float* windowWidths, int nWidths, pb_raster_out* out)
{
    const Grid d = gridOf(*dem);
    const int rows = d.h.nrRows, cols = d.h.nrCols;
    if (nWidths == 0) return PB_ERR_INVALID;
    std::vector<float> flat(size_t(rows) * cols, d.h.flag);
    const float threshold = float(EPSILON);
    for (int w = 0; w < nWidths; w++) {
        const float cellNr = float(round(windowWidths[w] / d.h.cellSize));
#pragma omp parallel for schedule(dynamic, 4)
        for (int row = 0; row < rows; row++)
            for (int col = 0; col < cols; col++) {
                const float z = d.at(row, col);
                if (isEqualF(z, d.h.flag)) continue;
                const int r1 = int(row - cellNr), r2 = int(row + cellNr), c1 = int(col - cellNr), c2 = int(col + cellNr);
                float higherSum = 0, lowerSum = 0, equalSum = 0, weightSum = 0;
                for (int wr = r1; wr <= r2; wr++)
                    for (int wc = c1; wc <= c2; wc++) {
                        if (unsigned(wr) >= unsigned(rows) || unsigned(wc) >= unsigned(cols)) continue;
                        const float value = d.at(wr, wc);
                        if (isEqualF(value, d.h.flag)) continue;
                        if (wr != row && wc != col) {
                            const float dx = float(row - wr), dy = float(col - wc);
                            const float cellDelta = sqrtf(dx * dx + dy * dy);
                            if (cellDelta <= cellNr) {
                                const float weight = 1 - (cellDelta / cellNr);
                                if (value - z > threshold) higherSum += weight;
                                else if (value - z < -threshold) lowerSum += weight;
                                else equalSum += weight;
                                weightSum += weight;
                            }
                        }
                    }
                if (weightSum > 0) {
                    float& o = flat[size_t(row) * cols + col];
                    if (isEqualF(o, d.h.flag)) o = (lowerSum - higherSum - equalSum * 0.5) / weightSum;
                    else o += (lowerSum - higherSum - equalSum * 0.5) / weightSum;
                }
            }
    }
    for (int row = 0; row < rows; row++) memcpy(outRow(out, row, cols), &flat[size_t(row) * cols], sizeof(float) * cols);
    out->nValid = int64_t(rows) * cols;
    minMaxOf(flat.data(), flat.size(), d.h.flag, &out->minimum, &out->maximum);
    return PB_OK;
}
