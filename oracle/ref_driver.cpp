/*
 * oracle/ref_driver.cpp — TEST INFRASTRUCTURE, never shipped, never on the product path.
 *
 * A thin C-ABI wrapper around the UNMODIFIED reference (ARPA-SIMC/PRAGA agrolib) so that the tests and
 * bench.py's cpu_baseline / --impl reference leg can run the reference's own CPU code on the same POD
 * descriptors the product's C ABI (include/praga_b200.h) takes.  It is compiled by oracle/Makefile
 * against the reference headers and linked with the reference's five Qt-free libraries, which are
 * compiled from the sources where they lie under /root/reference (nothing is copied into this repo);
 * the output goes to oracle/_ref/libpraga_ref.so.
 *
 * What runs reference code verbatim:
 *   ref_interpolation_raster  -> interpolationRaster()  agrolib/project/interpolationCmd.cpp:353-394
 *                                (that file needs Qt; the Makefile extracts exactly that function body
 *                                 into oracle/_ref/ at build time and compiles it with Qt-free includes)
 *   ref_pre_interpolation     -> preInterpolation()     agrolib/interpolation/interpolation.cpp:2444
 *   ref_compute_residuals     -> computeResiduals()     agrolib/interpolation/spatialControl.cpp:97
 *   ref_kriging_*             -> kriging.cpp:97-295
 *   ref_interpolate_points    -> interpolate()          interpolation.cpp:2502
 * What is restated here because its only copy lives in Qt-bound code:
 *   ref_local_detrending_raster  <- loop of Project::interpolationDemLocalDetrending, project.cpp:3208-3253
 *   cv statistics                <- Project::computeStatisticsCrossValidation, project.cpp:2935-2969
 */
#include <omp.h>
#include <math.h>
#include <string.h>
#include <chrono>
#include "solarRadiation.h"
#include "radiationSettings.h"
#include <string>
#include <map>
#include <vector>

#include "commonConstants.h"
#include "basicMath.h"
#include "furtherMathFunctions.h"
#include "statistics.h"
#include "gis.h"
#include "meteo.h"
#include "meteoPoint.h"
#include "interpolation.h"
#include "meteoGrid.h"
#include "interpolationSettings.h"
#include "interpolationPoint.h"
#include "spatialControl.h"
#include "kriging.h"

#include "../include/praga_b200.h"

/* extracted at build time from agrolib/project/interpolationCmd.cpp (see Makefile) */
bool interpolationRaster(std::vector<Crit3DInterpolationDataPoint>& dataPoints,
                         Crit3DInterpolationSettings& interpolationSettings, Crit3DMeteoSettings* meteoSettings,
                         gis::Crit3DRasterGrid* outputGrid, gis::Crit3DRasterGrid& raster, meteoVariable variable,
                         bool isParallelComputing);

/* ---- the enum values the C ABI copies by value must match the reference ---- */
static_assert(int(idw) == PB_METHOD_IDW && int(shepard) == PB_METHOD_SHEPARD &&
                  int(shepard_modified) == PB_METHOD_SHEPARD_MODIFIED, "TInterpolationMethod");
static_assert(int(proxyHeight) == PB_PROXY_HEIGHT && int(proxyUrbanFraction) == PB_PROXY_URBAN_FRACTION &&
                  int(proxyOrogIndex) == PB_PROXY_OROG_INDEX && int(proxySeaDistance) == PB_PROXY_SEA_DISTANCE &&
                  int(proxyAspect) == PB_PROXY_ASPECT && int(proxySlope) == PB_PROXY_SLOPE &&
                  int(proxyWaterIndex) == PB_PROXY_WATER_INDEX && int(noProxy) == PB_PROXY_NONE, "TProxyVar");
static_assert(int(piecewiseTwo) == PB_FIT_PIECEWISE_TWO && int(piecewiseThreeFree) == PB_FIT_PIECEWISE_THREE_FREE &&
                  int(piecewiseThree) == PB_FIT_PIECEWISE_THREE && int(linear) == PB_FIT_LINEAR &&
                  int(noFunction) == PB_FIT_NONE, "TFittingFunction");
static_assert(int(primary) == PB_LAPSE_PRIMARY && int(secondary) == PB_LAPSE_SECONDARY &&
                  int(supplemental) == PB_LAPSE_SUPPLEMENTAL, "lapseRateCodeType");
static_assert(int(KRIGING_SPHERICAL) == PB_KRIGING_SPHERICAL && int(KRIGING_LINEAR) == PB_KRIGING_LINEAR, "TkrigingMode");
static_assert(int(airTemperature) == PB_VAR_AIR_TEMPERATURE && int(dailyAirTemperatureMin) == PB_VAR_DAILY_TMIN &&
                  int(dailyAirTemperatureMax) == PB_VAR_DAILY_TMAX && int(dailyAirTemperatureAvg) == PB_VAR_DAILY_TAVG &&
                  int(dailyAirTemperatureRange) == PB_VAR_DAILY_TRANGE && int(precipitation) == PB_VAR_PRECIPITATION &&
                  int(dailyPrecipitation) == PB_VAR_DAILY_PRECIPITATION && int(airRelHumidity) == PB_VAR_AIR_REL_HUMIDITY &&
                  int(airDewTemperature) == PB_VAR_AIR_DEW_TEMPERATURE && int(dailyAirRelHumidityMin) == PB_VAR_DAILY_RH_MIN &&
                  int(dailyAirRelHumidityMax) == PB_VAR_DAILY_RH_MAX && int(dailyAirRelHumidityAvg) == PB_VAR_DAILY_RH_AVG &&
                  int(globalIrradiance) == PB_VAR_GLOBAL_IRRADIANCE && int(atmTransmissivity) == PB_VAR_ATM_TRANSMISSIVITY &&
                  int(dailyGlobalRadiation) == PB_VAR_DAILY_GLOBAL_RADIATION &&
                  int(windScalarIntensity) == PB_VAR_WIND_SCALAR_INTENSITY &&
                  int(windVectorIntensity) == PB_VAR_WIND_VECTOR_INTENSITY && int(windVectorX) == PB_VAR_WIND_VECTOR_X &&
                  int(windVectorY) == PB_VAR_WIND_VECTOR_Y &&
                  int(dailyWindVectorIntensityAvg) == PB_VAR_DAILY_WIND_VECTOR_INT_AVG &&
                  int(dailyWindVectorIntensityMax) == PB_VAR_DAILY_WIND_VECTOR_INT_MAX &&
                  int(dailyWindScalarIntensityAvg) == PB_VAR_DAILY_WIND_SCALAR_INT_AVG &&
                  int(dailyWindScalarIntensityMax) == PB_VAR_DAILY_WIND_SCALAR_INT_MAX &&
                  int(leafWetness) == PB_VAR_LEAF_WETNESS && int(dailyLeafWetness) == PB_VAR_DAILY_LEAF_WETNESS &&
                  int(atmPressure) == PB_VAR_ATM_PRESSURE && int(dailyReferenceEvapotranspirationHS) == PB_VAR_DAILY_ET0_HS &&
                  int(dailyWaterTableDepth) == PB_VAR_DAILY_WATER_TABLE_DEPTH && int(elaborationVar) == PB_VAR_ELABORATION &&
                  int(noMeteoVar) == PB_VAR_NONE && int(dailyBIC) == PB_VAR_DAILY_BIC, "meteoVariable");
static_assert(NODATA == PB_NODATA, "NODATA");
static_assert(int(aggrAverage) == PB_AGGR_AVERAGE && int(aggrMedian) == PB_AGGR_MEDIAN && int(aggrPrevailing) == PB_AGGR_PREVAILING &&
                  int(aggrCenter) == PB_AGGR_CENTER, "aggregationMethod");

namespace {

const char* proxyNameOf(int kind)
{
    switch (kind) {
    case PB_PROXY_HEIGHT: return "elevation";
    case PB_PROXY_URBAN_FRACTION: return "urbanFraction";
    case PB_PROXY_OROG_INDEX: return "orogIndex";
    case PB_PROXY_SEA_DISTANCE: return "seaDistance";
    case PB_PROXY_ASPECT: return "aspect";
    case PB_PROXY_SLOPE: return "slope";
    case PB_PROXY_WATER_INDEX: return "water_index";
    default: return "unknown";
    }
}

typedef double (*fit1d_t)(double, std::vector<double>&);

fit1d_t fitFunctionOf(int f)
{
    switch (f) {
    case PB_FIT_PIECEWISE_TWO: return lapseRatePiecewise_two;
    case PB_FIT_PIECEWISE_THREE_FREE: return lapseRatePiecewise_three_free;
    case PB_FIT_PIECEWISE_THREE: return lapseRatePiecewise_three;
    case PB_FIT_LINEAR: return functionLinear_intercept;
    default: return nullptr;
    }
}

int fitFunctionId(const std::function<double(double, std::vector<double>&)>& f)
{
    auto t = f.target<fit1d_t>();
    if (t == nullptr) return PB_FIT_NONE;
    if (*t == lapseRatePiecewise_two) return PB_FIT_PIECEWISE_TWO;
    if (*t == lapseRatePiecewise_three_free) return PB_FIT_PIECEWISE_THREE_FREE;
    if (*t == lapseRatePiecewise_three) return PB_FIT_PIECEWISE_THREE;
    if (*t == functionLinear_intercept) return PB_FIT_LINEAR;
    return PB_FIT_NONE;
}

/* gis::Crit3DRasterGrid owning its rows, filled from a pb_raster */
void fillGrid(gis::Crit3DRasterGrid& g, const pb_raster& r)
{
    gis::Crit3DRasterHeader h;
    h.nrRows = r.h.nrRows; h.nrCols = r.h.nrCols; h.cellSize = r.h.cellSize; h.invCellSize = r.h.invCellSize;
    h.flag = r.h.flag; h.llCorner.x = r.h.llx; h.llCorner.y = r.h.lly;
    g.initializeGrid(h);
    for (int row = 0; row < h.nrRows; row++) {
        const float* src = r.data ? r.data + size_t(row) * h.nrCols : r.rows[row];
        memcpy(g.value[row], src, sizeof(float) * h.nrCols);
    }
    g.isLoaded = true;
}

struct Scenario {
    Crit3DInterpolationSettings settings;
    Crit3DMeteoSettings meteoSettings;
    Crit3DClimateParameters climate;
    Crit3DTime time;
    std::vector<Crit3DInterpolationDataPoint> points;
    std::vector<gis::Crit3DRasterGrid*> grids;   // proxy grids (owned)
    ~Scenario()
    {
        for (auto& p : points) delete p.point;          // the reference never frees these (interpolationPoint.cpp:49)
        for (auto g : grids) delete g;
    }
};

void buildSettings(Scenario& sc, const pb_settings& s, const pb_raster* proxyGrids, int nProxyGrids)
{
    Crit3DInterpolationSettings& st = sc.settings;
    st.initialize();
    st.setInterpolationMethod(TInterpolationMethod(s.method));
    st.setUseThermalInversion(s.useThermalInversion != 0);
    st.setUseTD(s.useTD != 0);
    st.setUseLocalDetrending(s.useLocalDetrending != 0);
    st.setUseGlocalDetrending(s.useGlocalDetrending != 0);
    st.setUseLapseRateCode(s.useLapseRateCode != 0);
    st.setUseBestDetrending(s.useBestDetrending != 0);
    st.setUseMultipleDetrending(s.useMultipleDetrending != 0);
    st.setUseDoNotRetrend(s.useDoNotRetrend != 0);
    st.setUseRetrendOnly(s.useRetrendOnly != 0);
    st.setPrecipitationAllZero(s.precipitationAllZero != 0);
    st.setMinPointsLocalDetrending(s.minPointsLocalDetrending);
    st.setTopoDist_Kh(s.topoDist_Kh);
    st.setTopoDist_maxKh(s.topoDist_maxKh);
    st.setMinRegressionR2(s.minRegressionR2);
    st.setPointsBoundingBoxArea(s.pointsBoundingBoxArea);
    st.setLocalRadius(s.localRadius);
    if (s.hasPointsRange) st.setPointsRange(s.pointsRangeMin, s.pointsRangeMax);
    sc.meteoSettings.setRainfallThreshold(s.rainfallThreshold);

    for (int g = 0; g < nProxyGrids; g++) {
        gis::Crit3DRasterGrid* grid = new gis::Crit3DRasterGrid();
        fillGrid(*grid, proxyGrids[g]);
        sc.grids.push_back(grid);
    }

    for (int i = 0; i < s.nProxy; i++) {
        const pb_proxy& p = s.proxy[i];
        Crit3DProxy proxy;
        proxy.setName(proxyNameOf(p.kind));
        if (p.gridIndex >= 0 && p.gridIndex < nProxyGrids) proxy.setGrid(sc.grids[p.gridIndex]);
        proxy.setFittingFunctionName(TFittingFunction(p.fittingFunction));
        proxy.setInversionIsSignificative(p.inversionIsSignificative != 0);
        proxy.setRegressionR2(p.regressionR2);
        proxy.setRegressionSlope(p.regressionSlope);
        proxy.setRegressionIntercept(p.regressionIntercept);
        proxy.setLapseRateH0(p.lapseRateH0);
        proxy.setLapseRateH1(p.lapseRateH1);
        proxy.setLapseRateT0(p.lapseRateT0);
        proxy.setLapseRateT1(p.lapseRateT1);
        proxy.setInversionLapseRate(p.inversionLapseRate);
        proxy.setStdDevThreshold(p.stdDevThreshold);
        std::vector<double> range;
        for (int j = 0; j < p.nFitParams; j++) range.push_back(p.fitMin[j]);
        for (int j = 0; j < p.nFitParams; j++) range.push_back(p.fitMax[j]);
        proxy.setFittingParametersRange(range);
        std::vector<int> fg;
        for (int j = 0; j < p.nFitParams; j++) fg.push_back(p.fitFirstGuess[j]);
        proxy.setFittingFirstGuess(fg);
        std::vector<std::vector<double>> combos;
        for (int k = 0; k < p.nFirstGuessCombinations; k++)
            combos.push_back(std::vector<double>(p.firstGuessCombinations[k], p.firstGuessCombinations[k] + p.nFitParams));
        proxy.setFirstGuessCombinations(combos);
        st.addProxy(proxy, s.selectedActive[i] != 0);
    }

    Crit3DProxyCombination sel, cur;
    for (int i = 0; i < s.nProxy; i++) {
        sel.addProxyActive(s.selectedActive[i] != 0);
        sel.addProxySignificant(false);
        cur.addProxyActive(s.currentActive[i] != 0);
        cur.addProxySignificant(s.currentSignificant[i] != 0);
    }
    sel.setUseThermalInversion(s.selectedUseThermalInversion != 0);
    cur.setUseThermalInversion(s.currentUseThermalInversion != 0);
    st.setSelectedCombination(sel);
    st.setCurrentCombination(cur);

    st.clearFitting();
    std::vector<std::vector<double>> params;
    for (int i = 0; i < s.nFit; i++) {
        fit1d_t f = fitFunctionOf(s.fitFunction[i]);
        if (f) st.addFittingFunction(f);
        params.push_back(std::vector<double>(s.fitParams[i], s.fitParams[i] + s.fitNrParams[i]));
    }
    st.setFittingParameters(params);

    /* climate: one lapse rate for every month; 06:00 makes the hourly blend return it exactly (meteo.cpp:173-222) */
    sc.climate.tminLapseRate.assign(12, s.climateLapseRate);
    sc.climate.tmaxLapseRate.assign(12, s.climateLapseRate);
    sc.climate.tdMinLapseRate.assign(12, s.climateLapseRate);
    sc.climate.tdMaxLapseRate.assign(12, s.climateLapseRate);
    sc.time = Crit3DTime(Crit3DDate(15, 1, 2023), 6 * 3600);
}

void exportSettings(Scenario& sc, pb_settings& s)
{
    Crit3DInterpolationSettings& st = sc.settings;
    s.precipitationAllZero = st.getPrecipitationAllZero() ? 1 : 0;
    s.topoDist_Kh = st.getTopoDist_Kh();
    s.localRadius = st.getLocalRadius();
    Crit3DProxyCombination cur = st.getCurrentCombination();
    s.currentUseThermalInversion = cur.getUseThermalInversion() ? 1 : 0;
    for (int i = 0; i < s.nProxy; i++) {
        s.currentActive[i] = cur.isProxyActive(i) ? 1 : 0;
        s.currentSignificant[i] = cur.isProxySignificant(i) ? 1 : 0;
        Crit3DProxy* p = st.getProxy(i);
        pb_proxy& q = s.proxy[i];
        q.inversionIsSignificative = p->getInversionIsSignificative() ? 1 : 0;
        q.regressionR2 = p->getRegressionR2();
        q.regressionSlope = p->getRegressionSlope();
        q.regressionIntercept = p->getRegressionIntercept();
        q.lapseRateH0 = p->getLapseRateH0();
        q.lapseRateH1 = p->getLapseRateH1();
        q.lapseRateT0 = p->getLapseRateT0();
        q.lapseRateT1 = p->getLapseRateT1();
        q.inversionLapseRate = p->getInversionLapseRate();
        std::vector<double> range = p->getFittingParametersRange();
        int np = int(range.size() / 2);
        q.nFitParams = np;
        for (int j = 0; j < np && j < PB_MAX_FIT_PARAMS; j++) { q.fitMin[j] = range[j]; q.fitMax[j] = range[np + j]; }
        std::vector<std::vector<double>> combos = p->getFirstGuessCombinations();
        q.nFirstGuessCombinations = int(combos.size());
        for (size_t k = 0; k < combos.size() && k < PB_MAX_FIRST_GUESS; k++)
            for (size_t j = 0; j < combos[k].size() && j < PB_MAX_FIT_PARAMS; j++) q.firstGuessCombinations[k][j] = combos[k][j];
    }
    auto funcs = st.getFittingFunction();
    auto params = st.getFittingParameters();
    s.nFit = int(params.size());
    for (int i = s.nFit; i < int(funcs.size()) && i < PB_MAX_PROXY; i++) s.fitFunction[i] = fitFunctionId(funcs[i]);
    for (int i = 0; i < s.nFit && i < PB_MAX_PROXY; i++) {
        s.fitFunction[i] = i < int(funcs.size()) ? fitFunctionId(funcs[i]) : PB_FIT_NONE;
        s.fitNrParams[i] = int(params[i].size());
        for (size_t j = 0; j < params[i].size() && j < PB_MAX_FIT_PARAMS; j++) s.fitParams[i][j] = params[i][j];
    }
    s.nFitFunctions = int(funcs.size());
}

// Crit3DMeteoPoint::topographicDistance of the test scenario (Project::loadTopographicDistanceMaps, project.cpp:2368-2430),
// keyed by point index; set through ref_set_topographic_distance_maps
std::map<int, gis::Crit3DRasterGrid*> g_tdMaps;
gis::Crit3DRasterGrid* tdMapOf(int index)
{
    auto it = g_tdMaps.find(index);
    return it == g_tdMaps.end() ? nullptr : it->second;
}

void buildPoints(Scenario& sc, const pb_stations& st)
{
    sc.points.resize(st.n);
    for (int i = 0; i < st.n; i++) {
        Crit3DInterpolationDataPoint& p = sc.points[i];
        p.point->utm.x = st.x[i];
        p.point->utm.y = st.y[i];
        p.point->z = st.z[i];
        p.index = st.index ? st.index[i] : i;
        p.topographicDistance = tdMapOf(p.index);        // passDataToInterpolation, spatialControl.cpp:529
        p.isActive = st.isActive ? st.isActive[i] != 0 : true;
        p.value = st.value[i];
        p.regressionWeight = st.regressionWeight ? st.regressionWeight[i] : float(NODATA);
        p.lapseRateCode = st.lapseRateCode ? lapseRateCodeType(st.lapseRateCode[i]) : primary;
        p.proxyValues.assign(st.proxyValues + size_t(i) * st.nProxy, st.proxyValues + size_t(i + 1) * st.nProxy);
    }
}

void copyOut(const gis::Crit3DRasterGrid& g, pb_raster_out* out)
{
    for (int row = 0; row < g.header->nrRows; row++) {
        float* dst = out->data ? out->data + size_t(row) * g.header->nrCols : out->rows[row];
        memcpy(dst, g.value[row], sizeof(float) * g.header->nrCols);
    }
    out->minimum = g.minimum;
    out->maximum = g.maximum;
}

}  // namespace

extern "C" {

int ref_max_threads(void) { return omp_get_max_threads(); }

/* the points' precomputed topographic-distance maps (what Project::loadTopographicDistanceMaps leaves in
 * Crit3DMeteoPoint::topographicDistance): every driver below hands them to its points; n = 0 drops all */
int ref_set_topographic_distance_maps(const int32_t* pointIndex, const pb_raster* maps, int n)
{
    if (n <= 0) {
        for (auto& kv : g_tdMaps) delete kv.second;
        g_tdMaps.clear();
        return PB_OK;
    }
    for (int k = 0; k < n; k++) {
        auto it = g_tdMaps.find(pointIndex[k]);
        if (it != g_tdMaps.end()) { delete it->second; g_tdMaps.erase(it); }
        if (!maps[k].data && !maps[k].rows) continue;
        gis::Crit3DRasterGrid* g = new gis::Crit3DRasterGrid();
        fillGrid(*g, maps[k]);
        g_tdMaps[pointIndex[k]] = g;
    }
    return PB_OK;
}

/* interpolationRaster, verbatim reference code. nThreads <= 0: all cores (the reference forces that anyway,
 * interpolationCmd.cpp:362-367; OMP_NUM_THREADS / omp_set_num_threads before the call is how we time 1 thread).
 * elapsed_s (optional) = wall time of the interpolationRaster call alone. */
int ref_interpolation_raster(const pb_stations* st, const pb_settings* s, int variable, const pb_raster* dem,
                             const pb_raster* proxyGrids, int nProxyGrids, int nThreads, pb_raster_out* out,
                             double* elapsed_s)
{
    Scenario sc;
    buildSettings(sc, *s, proxyGrids, nProxyGrids);
    buildPoints(sc, *st);
    gis::Crit3DRasterGrid demGrid, outGrid;
    fillGrid(demGrid, *dem);
    sc.settings.setCurrentDEM(&demGrid);
    if (nThreads > 0) omp_set_num_threads(nThreads);
    auto t0 = std::chrono::steady_clock::now();
    bool ok = interpolationRaster(sc.points, sc.settings, &sc.meteoSettings, &outGrid, demGrid, meteoVariable(variable),
                                  nThreads != 1);
    auto t1 = std::chrono::steady_clock::now();
    if (elapsed_s) *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
    int64_t nValid = 0;
    for (int row = 0; row < demGrid.header->nrRows; row++)
        for (int col = 0; col < demGrid.header->nrCols; col++)
            if (!isEqual(demGrid.value[row][col], demGrid.header->flag)) nValid++;
    out->nValid = nValid;
    if (outGrid.isLoaded) copyOut(outGrid, out);
    return ok ? PB_OK : PB_ERR_NO_VALID_CELL;
}

/* preInterpolation, verbatim reference code; station values and settings are written back.
 * keep[i] = 1 if point i is still in myPoints afterwards (multiple detrending erases points, interpolation.cpp:2230). */
int ref_pre_interpolation(pb_stations* st, pb_settings* s, int variable, uint8_t* keep, const pb_raster* proxyGrids,
                          int nProxyGrids)
{
    Scenario sc;
    buildSettings(sc, *s, proxyGrids, nProxyGrids);
    buildPoints(sc, *st);
    for (int i = 0; i < st->n; i++) sc.points[i].index = i;   // used to map survivors back
    std::vector<Crit3DMeteoPoint> meteoPoints;                // only read by optimalDetrending / TD optimisation
    std::string err;
    bool ok = preInterpolation(sc.points, sc.settings, &sc.meteoSettings, &sc.climate, meteoPoints,
                               meteoVariable(variable), sc.time, err);
    if (keep) memset(keep, 0, st->n);
    for (auto& p : sc.points) {
        st->value[p.index] = p.value;
        if (keep) keep[p.index] = 1;
    }
    exportSettings(sc, *s);
    return ok ? PB_OK : PB_ERR_FIT;
}

/* preInterpolation with the meteo points and the current DEM it needs for optimalDetrending (:2395-2441) and
 * topographicDistanceOptimize (:2356-2392), verbatim reference code. One Crit3DMeteoPoint per station (active, accepted,
 * currentValue = cv->currentValue or the station value at entry). optimalDetrending rebuilds the interpolation points from
 * the meteo points (passDataToInterpolation), so their count can differ from st->n only if a point were inactive. */
int ref_pre_interpolation_ex(pb_stations* st, pb_settings* s, int variable, const pb_cv_points* cv, const pb_raster* dem,
                             uint8_t* keep, pb_kh_series* khSeries)
{
    Scenario sc;
    buildSettings(sc, *s, nullptr, 0);
    buildPoints(sc, *st);
    gis::Crit3DRasterGrid demGrid;
    if (dem) {
        fillGrid(demGrid, *dem);
        sc.settings.setCurrentDEM(&demGrid);
    }
    sc.settings.setUseExcludeStationsOutsideDEM(cv && cv->excludeOutsideDem);
    std::vector<Crit3DMeteoPoint> mp(st->n);
    for (int i = 0; i < st->n; i++) {
        sc.points[i].index = i;
        mp[i].active = true;
        mp[i].topographicDistance = tdMapOf(i);
        mp[i].quality = quality::accepted;
        mp[i].currentValue = (cv && cv->currentValue) ? cv->currentValue[i] : st->value[i];
        mp[i].point.utm.x = st->x[i];
        mp[i].point.utm.y = st->y[i];
        mp[i].point.z = st->z[i];
        mp[i].isInsideDem = (cv && cv->isInsideDem) ? cv->isInsideDem[i] != 0 : true;
        mp[i].lapseRateCode = st->lapseRateCode ? lapseRateCodeType(st->lapseRateCode[i]) : primary;
        for (int k = 0; k < st->nProxy; k++) mp[i].proxyValues.push_back(st->proxyValues[size_t(i) * st->nProxy + k]);
    }
    std::string err;
    bool ok = preInterpolation(sc.points, sc.settings, &sc.meteoSettings, &sc.climate, mp, meteoVariable(variable), sc.time, err);
    if (keep) memset(keep, 0, st->n);
    for (auto& p : sc.points) {
        st->value[p.index] = p.value;
        if (keep) keep[p.index] = 1;
    }
    exportSettings(sc, *s);
    s->pointsBoundingBoxArea = sc.settings.getPointsBoundingBoxArea();
    std::vector<double> pr = sc.settings.getPointsRange();
    if (pr.size() == 2) { s->hasPointsRange = 1; s->pointsRangeMin = pr[0]; s->pointsRangeMax = pr[1]; }
    if (khSeries) {
        std::vector<float> kh = sc.settings.getKh_series(), e = sc.settings.getKh_error_series();
        khSeries->n = int(std::min(kh.size(), size_t(PB_MAX_KH_SERIES)));
        for (int i = 0; i < khSeries->n; i++) { khSeries->kh[i] = kh[i]; khSeries->error[i] = e[i]; }
    }
    return ok ? PB_OK : PB_ERR_FIT;
}

/* spatialQualityControl (spatialControl.cpp:331-427) on one accepted, active Crit3DMeteoPoint per station */
int ref_spatial_quality_control(const pb_stations* st, pb_settings* s, int variable, const pb_cv_points* cv, const pb_raster* dem,
                                uint8_t* wrongSpatial)
{
    Scenario sc;
    buildSettings(sc, *s, nullptr, 0);
    gis::Crit3DRasterGrid demGrid;
    if (dem) {
        fillGrid(demGrid, *dem);
        sc.settings.setCurrentDEM(&demGrid);
    }
    sc.settings.setUseExcludeStationsOutsideDEM(cv && cv->excludeOutsideDem);
    std::vector<Crit3DMeteoPoint> mp(st->n);
    for (int i = 0; i < st->n; i++) {
        mp[i].active = true;
        mp[i].topographicDistance = tdMapOf(i);
        mp[i].quality = quality::accepted;
        mp[i].currentValue = (cv && cv->currentValue) ? cv->currentValue[i] : st->value[i];
        mp[i].point.utm.x = st->x[i];
        mp[i].point.utm.y = st->y[i];
        mp[i].point.z = st->z[i];
        mp[i].isInsideDem = (cv && cv->isInsideDem) ? cv->isInsideDem[i] != 0 : true;
        mp[i].lapseRateCode = st->lapseRateCode ? lapseRateCodeType(st->lapseRateCode[i]) : primary;
        for (int k = 0; k < st->nProxy; k++) mp[i].proxyValues.push_back(st->proxyValues[size_t(i) * st->nProxy + k]);
    }
    std::string err;
    bool ok = spatialQualityControl(meteoVariable(variable), mp, sc.settings, &sc.meteoSettings, &sc.climate, sc.time, err);
    for (int i = 0; i < st->n; i++) wrongSpatial[i] = mp[i].quality == quality::wrong_spatial ? 1 : 0;
    exportSettings(sc, *s);
    s->pointsBoundingBoxArea = sc.settings.getPointsBoundingBoxArea();
    std::vector<double> pr = sc.settings.getPointsRange();
    if (pr.size() == 2) { s->hasPointsRange = 1; s->pointsRangeMin = pr[0]; s->pointsRangeMax = pr[1]; }
    return ok ? PB_OK : PB_ERR_FIT;
}

/* One time step as Project::interpolationDem runs it (project/project.cpp:3106-3150; the method lives in the Qt-bound
 * Project class, every call inside is reference code): spatialQualityControl (the part of checkData :466-476 that is
 * not a time-series lookup), passDataToInterpolation, preInterpolation, interpolationRaster; with step->residual also the
 * cross-validation of Project::interpolationCv (:3063-3067 computeResiduals + computeStatisticsCrossValidation :2935). */
int ref_interpolation_dem(const pb_stations* meteo, pb_settings* sq, pb_settings* s, int variable, const pb_raster* dem,
                          const pb_raster* proxyGrids, int nProxyGrids, int nThreads, pb_raster_out* out, pb_dem_step* step)
{
    const meteoVariable myVar = meteoVariable(variable);
    Scenario sc, scq;
    buildSettings(sc, *s, proxyGrids, nProxyGrids);
    gis::Crit3DRasterGrid demGrid, outGrid;
    fillGrid(demGrid, *dem);
    sc.settings.setCurrentDEM(&demGrid);
    std::vector<Crit3DMeteoPoint> mp(meteo->n);
    for (int i = 0; i < meteo->n; i++) {
        mp[i].active = true;
        mp[i].topographicDistance = tdMapOf(i);
        mp[i].quality = quality::accepted;
        mp[i].currentValue = meteo->value[i];
        mp[i].point.utm.x = meteo->x[i];
        mp[i].point.utm.y = meteo->y[i];
        mp[i].point.z = meteo->z[i];
        mp[i].isInsideDem = true;
        mp[i].lapseRateCode = meteo->lapseRateCode ? lapseRateCodeType(meteo->lapseRateCode[i]) : primary;
        for (int k = 0; k < meteo->nProxy; k++) mp[i].proxyValues.push_back(meteo->proxyValues[size_t(i) * meteo->nProxy + k]);
    }
    std::string err;
    if (sq && myVar != precipitation && myVar != dailyPrecipitation && myVar != windVectorX && myVar != windVectorY &&
        myVar != windVectorDirection && myVar != dailyWindVectorDirectionPrevailing) {
        buildSettings(scq, *sq, nullptr, 0);
        scq.settings.setCurrentDEM(&demGrid);
        if (!spatialQualityControl(myVar, mp, scq.settings, &scq.meteoSettings, &scq.climate, scq.time, err)) return PB_ERR_FIT;
        exportSettings(scq, *sq);
        sq->pointsBoundingBoxArea = scq.settings.getPointsBoundingBoxArea();
        std::vector<double> pr = scq.settings.getPointsRange();
        if (pr.size() == 2) { sq->hasPointsRange = 1; sq->pointsRangeMin = pr[0]; sq->pointsRangeMax = pr[1]; }
    }
    if (step && step->wrongSpatial)
        for (int i = 0; i < meteo->n; i++) step->wrongSpatial[i] = mp[i].quality == quality::wrong_spatial ? 1 : 0;
    std::vector<Crit3DInterpolationDataPoint> points;
    if (!passDataToInterpolation(mp, points, sc.settings)) return PB_ERR_INVALID;
    if (sc.settings.getUseMultipleDetrending()) sc.settings.clearFitting();
    if (!preInterpolation(points, sc.settings, &sc.meteoSettings, &sc.climate, mp, myVar, sc.time, err)) return PB_ERR_FIT;
    if (nThreads > 0) omp_set_num_threads(nThreads);
    bool ok = interpolationRaster(points, sc.settings, &sc.meteoSettings, &outGrid, demGrid, myVar, nThreads != 1);
    int64_t nValid = 0;
    for (int row = 0; row < demGrid.header->nrRows; row++)
        for (int col = 0; col < demGrid.header->nrCols; col++)
            if (!isEqual(demGrid.value[row][col], demGrid.header->flag)) nValid++;
    out->nValid = nValid;
    if (outGrid.isLoaded) copyOut(outGrid, out);
    if (step && step->residual) {
        computeResiduals(myVar, mp, points, sc.settings, &sc.meteoSettings, sc.settings.getUseExcludeStationsOutsideDEM(), true);
        std::vector<float> obs, pre;
        for (int i = 0; i < meteo->n; i++) {
            step->residual[i] = mp[i].residual;
            float value = mp[i].currentValue;
            if (!isEqual(value, NODATA) && !isEqual(mp[i].residual, NODATA)) {
                obs.push_back(value);
                pre.push_back(value - mp[i].residual);
            }
        }
        if (step->stats) {
            pb_cv_stats* stats = step->stats;
            memset(stats, 0, sizeof(*stats));
            stats->n = int(obs.size());
            stats->meanAbsoluteError = stats->meanBiasError = stats->rootMeanSquareError = stats->nashSutcliffe = stats->R2 = NODATA;
            if (obs.size() > 0) {
                stats->meanAbsoluteError = statistics::meanAbsoluteError(obs, pre);
                stats->meanBiasError = statistics::meanError(obs, pre);
                stats->rootMeanSquareError = statistics::rootMeanSquareError(obs, pre);
                stats->nashSutcliffe = statistics::NashSutcliffeEfficiency(obs, pre);
                float intercept, slope, r2;
                statistics::linearRegression(obs, pre, int(obs.size()), false, &intercept, &slope, &r2);
                stats->R2 = r2;
            }
        }
    }
    if (step && step->khSeries) {
        std::vector<float> kh = sc.settings.getKh_series(), e = sc.settings.getKh_error_series();
        step->khSeries->n = int(std::min(kh.size(), size_t(PB_MAX_KH_SERIES)));
        for (int i = 0; i < step->khSeries->n; i++) { step->khSeries->kh[i] = kh[i]; step->khSeries->error[i] = e[i]; }
    }
    exportSettings(sc, *s);
    s->pointsBoundingBoxArea = sc.settings.getPointsBoundingBoxArea();
    std::vector<double> pr = sc.settings.getPointsRange();
    if (pr.size() == 2) { s->hasPointsRange = 1; s->pointsRangeMin = pr[0]; s->pointsRangeMax = pr[1]; }
    return ok ? PB_OK : PB_ERR_NO_VALID_CELL;
}

/* tDewFromRelHum(float, float) (meteo/meteo.cpp:275-285): what Crit3DMeteoPoint::getMeteoPointValueH returns as a station's
 * airDewTemperature when none is stored (meteoPoint.cpp:967-972); reference code */
void ref_tdew_from_relhum(const float* rh, const float* t, int n, float* tdew)
{
    for (int i = 0; i < n; i++) tdew[i] = tDewFromRelHum(rh[i], t[i]);
}

/* Crit3DHourlyMeteoMaps::computeRelativeHumidityMap (project/meteoMaps.cpp:301-321; the class is Qt-bound, the loop is
 * restated, relHumFromTdew and updateMinMaxRasterGrid are reference code) */
int ref_relative_humidity_map(const pb_raster* airT, const pb_raster* dewT, pb_raster_out* out)
{
    gis::Crit3DRasterGrid tGrid, tdGrid, rh;
    fillGrid(tGrid, *airT);
    fillGrid(tdGrid, *dewT);
    rh.initializeGrid(tGrid);
    rh.emptyGrid();
    for (long row = 0; row < rh.header->nrRows; row++)
        for (long col = 0; col < rh.header->nrCols; col++) {
            float myT = tGrid.value[row][col], dewPoint = tdGrid.value[row][col];
            if (!isEqual(myT, tGrid.header->flag) && !isEqual(dewPoint, tdGrid.header->flag))
                rh.value[row][col] = relHumFromTdew(dewPoint, myT);
        }
    bool ok = gis::updateMinMaxRasterGrid(&rh);
    out->nValid = int64_t(rh.header->nrRows) * rh.header->nrCols;
    copyOut(rh, out);
    return ok ? PB_OK : PB_ERR_NO_VALID_CELL;
}

/* Crit3DDailyMeteoMaps::fixDailyThermalConsistency (project/meteoMaps.cpp:103-126), loop restated; tmin is rewritten */
int ref_fix_daily_thermal_consistency(const pb_raster* tmax, const pb_raster* tmin, float* tminOut, int64_t* nFixed)
{
    gis::Crit3DRasterGrid a, b;
    fillGrid(a, *tmax);
    fillGrid(b, *tmin);
    int64_t n = 0;
    for (unsigned row = 0; row < unsigned(a.header->nrRows); row++)
        for (unsigned col = 0; col < unsigned(a.header->nrCols); col++) {
            if (!isEqual(a.value[row][col], a.header->flag) && !isEqual(b.value[row][col], b.header->flag))
                if (b.value[row][col] > a.value[row][col]) {
                    b.value[row][col] = a.value[row][col] - 0.1f;
                    n++;
                }
            tminOut[size_t(row) * a.header->nrCols + col] = b.value[row][col];
        }
    if (nFixed) *nFixed = n;
    return PB_OK;
}

/* gis::topographicDistanceMap, verbatim reference code */
int ref_topographic_distance_map(const pb_raster* dem, double x, double y, double z, pb_raster_out* out)
{
    gis::Crit3DRasterGrid demGrid, map;
    fillGrid(demGrid, *dem);
    gis::Crit3DPoint p;
    p.utm.x = x; p.utm.y = y; p.z = z;
    bool ok = gis::topographicDistanceMap(p, demGrid, &map);
    out->nValid = int64_t(demGrid.header->nrRows) * demGrid.header->nrCols;
    copyOut(map, out);
    return ok ? PB_OK : PB_ERR_INVALID;
}

/* gis::resampleGrid, verbatim reference code */
int ref_resample_grid(const pb_raster* src, const pb_raster_header* newHeader, int method, float nodataRatioThreshold,
                      pb_raster_out* out)
{
    gis::Crit3DRasterGrid oldGrid, newGrid;
    fillGrid(oldGrid, *src);
    gis::Crit3DRasterHeader h;
    h.nrRows = newHeader->nrRows; h.nrCols = newHeader->nrCols;
    h.cellSize = newHeader->cellSize; h.invCellSize = newHeader->invCellSize;
    h.llCorner.x = newHeader->llx; h.llCorner.y = newHeader->lly;
    h.flag = newHeader->flag;
    gis::resampleGrid(oldGrid, &newGrid, &h, aggregationMethod(method), nodataRatioThreshold);
    out->nValid = int64_t(h.nrRows) * h.nrCols;
    copyOut(newGrid, out);
    return PB_OK;
}

/* gis::writeEsriGrid / gis::readEsriGrid, verbatim reference code (gis/gisIO.cpp) */
}  // extern "C"
namespace gis { bool readEsriGridHeader(const std::string& fileName, gis::Crit3DRasterHeader* header, std::string& errorStr); }   // gisIO.cpp:118, not in gis.h
extern "C" {
int ref_write_esri(const pb_raster* r, const char* fileName)
{
    gis::Crit3DRasterGrid g;
    fillGrid(g, *r);
    std::string err;
    return gis::writeEsriGrid(fileName, &g, err) ? PB_OK : PB_ERR_INVALID;
}
int ref_read_esri_header(const char* fileName, pb_raster_header* h)
{
    gis::Crit3DRasterHeader header;
    std::string err;
    if (!gis::readEsriGridHeader(fileName, &header, err)) return PB_ERR_INVALID;
    memset(h, 0, sizeof(*h));
    h->nrRows = header.nrRows; h->nrCols = header.nrCols; h->cellSize = header.cellSize; h->invCellSize = header.invCellSize;
    h->llx = header.llCorner.x; h->lly = header.llCorner.y; h->flag = header.flag;
    return PB_OK;
}
int ref_read_esri(const char* fileName, pb_raster_out* out)
{
    gis::Crit3DRasterGrid g;
    std::string err;
    if (!gis::readEsriGrid(fileName, &g, err)) return PB_ERR_INVALID;
    out->nValid = int64_t(g.header->nrRows) * g.header->nrCols;
    copyOut(g, out);
    return PB_OK;
}

/* interpolate() at arbitrary targets (x, y, z, proxyValues[nProxy]) — interpolation.cpp:2502 */
int ref_interpolate_points(const pb_stations* st, const pb_settings* s, int variable, const float* x, const float* y,
                           const float* z, const double* proxyValues, int m, int excludeSupplemental, float* result)
{
    Scenario sc;
    buildSettings(sc, *s, nullptr, 0);
    buildPoints(sc, *st);
    std::vector<double> pv(s->nProxy);
    for (int i = 0; i < m; i++) {
        for (int k = 0; k < s->nProxy; k++) pv[k] = proxyValues[size_t(i) * s->nProxy + k];
        result[i] = interpolate(sc.points, sc.settings, &sc.meteoSettings, meteoVariable(variable), x[i], y[i], z[i], pv,
                                excludeSupplemental != 0);
    }
    return PB_OK;
}

/* interpolate() with a current DEM set, so that topographic distance (computeDistances :226-251) is available */
int ref_interpolate_points_td(const pb_stations* st, const pb_settings* s, int variable, const float* x, const float* y,
                              const float* z, const double* proxyValues, int m, int excludeSupplemental, const pb_raster* dem,
                              float* result)
{
    Scenario sc;
    buildSettings(sc, *s, nullptr, 0);
    buildPoints(sc, *st);
    gis::Crit3DRasterGrid demGrid;
    fillGrid(demGrid, *dem);
    sc.settings.setCurrentDEM(&demGrid);
    std::vector<double> pv(s->nProxy);
    for (int i = 0; i < m; i++) {
        for (int k = 0; k < s->nProxy; k++) pv[k] = proxyValues[size_t(i) * s->nProxy + k];
        result[i] = interpolate(sc.points, sc.settings, &sc.meteoSettings, meteoVariable(variable), x[i], y[i], z[i], pv,
                                excludeSupplemental != 0);
    }
    return PB_OK;
}

/* computeResiduals (spatialControl.cpp:97-155) + Project::computeStatisticsCrossValidation (project.cpp:2935-2969,
 * restated: that copy needs Qt). One Crit3DMeteoPoint per station: active, quality accepted, currentValue =
 * observed[i], position/proxy values of the station. */
int ref_compute_residuals(const pb_stations* st, const pb_settings* s, int variable, const float* observed,
                          const uint8_t* useForCv, float* residual, pb_cv_stats* stats)
{
    Scenario sc;
    buildSettings(sc, *s, nullptr, 0);
    buildPoints(sc, *st);
    std::vector<Crit3DMeteoPoint> mp(st->n);
    for (int i = 0; i < st->n; i++) {
        mp[i].active = useForCv ? useForCv[i] != 0 : true;
        mp[i].quality = quality::accepted;
        mp[i].currentValue = observed[i];
        mp[i].point.utm.x = st->x[i];
        mp[i].point.utm.y = st->y[i];
        mp[i].point.z = st->z[i];
        mp[i].isInsideDem = true;
        mp[i].lapseRateCode = st->lapseRateCode ? lapseRateCodeType(st->lapseRateCode[i]) : primary;
        for (int k = 0; k < st->nProxy; k++) mp[i].proxyValues.push_back(st->proxyValues[size_t(i) * st->nProxy + k]);
        sc.points[i].index = i;
    }
    bool ok = computeResiduals(meteoVariable(variable), mp, sc.points, sc.settings, &sc.meteoSettings, false, false);
    std::vector<float> obs, pre;
    for (int i = 0; i < st->n; i++) {
        residual[i] = mp[i].residual;
        if (mp[i].active) {
            float value = mp[i].currentValue;
            if (!isEqual(value, NODATA) && !isEqual(mp[i].residual, NODATA)) {
                obs.push_back(value);
                pre.push_back(value - mp[i].residual);
            }
        }
    }
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->n = int(obs.size());
        stats->meanAbsoluteError = stats->meanBiasError = stats->rootMeanSquareError = stats->nashSutcliffe = stats->R2 = NODATA;
        if (obs.size() > 0) {
            stats->meanAbsoluteError = statistics::meanAbsoluteError(obs, pre);
            stats->meanBiasError = statistics::meanError(obs, pre);
            stats->rootMeanSquareError = statistics::rootMeanSquareError(obs, pre);
            stats->nashSutcliffe = statistics::NashSutcliffeEfficiency(obs, pre);
            float intercept, slope, r2;
            statistics::linearRegression(obs, pre, int(obs.size()), false, &intercept, &slope, &r2);
            stats->R2 = r2;
        }
    }
    return ok ? PB_OK : PB_ERR_INVALID;
}

/* computeResidualsLocalDetrending (spatialControl.cpp:158-228) called like Project::interpolationCv does
 * (project/project.cpp:3086-3097) + the statistics of computeStatisticsCrossValidation (:2935-2969) */
int ref_compute_residuals_local(const pb_stations* st, const pb_settings* s, int variable, const float* observed,
                                const uint8_t* useForCv, float* residual, pb_cv_stats* stats)
{
    Scenario sc;
    buildSettings(sc, *s, nullptr, 0);
    buildPoints(sc, *st);
    std::vector<Crit3DMeteoPoint> mp(st->n);
    for (int i = 0; i < st->n; i++) {
        mp[i].active = useForCv ? useForCv[i] != 0 : true;
        mp[i].quality = quality::accepted;
        mp[i].currentValue = observed[i];
        mp[i].point.utm.x = st->x[i];
        mp[i].point.utm.y = st->y[i];
        mp[i].point.z = st->z[i];
        mp[i].isInsideDem = true;
        mp[i].lapseRateCode = st->lapseRateCode ? lapseRateCodeType(st->lapseRateCode[i]) : primary;
        for (int k = 0; k < st->nProxy; k++) mp[i].proxyValues.push_back(st->proxyValues[size_t(i) * st->nProxy + k]);
        sc.points[i].index = i;
    }
    if (!setMultipleDetrendingHeightTemperatureRange(sc.settings)) return PB_ERR_FIT;
    bool ok = computeResidualsLocalDetrending(meteoVariable(variable), sc.time, mp, sc.points, sc.settings, &sc.meteoSettings,
                                              &sc.climate, true, true);
    std::vector<float> obs, pre;
    for (int i = 0; i < st->n; i++) {
        residual[i] = mp[i].residual;
        if (mp[i].active) {
            float value = mp[i].currentValue;
            if (!isEqual(value, NODATA) && !isEqual(mp[i].residual, NODATA)) {
                obs.push_back(value);
                pre.push_back(value - mp[i].residual);
            }
        }
    }
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->n = int(obs.size());
        stats->meanAbsoluteError = stats->meanBiasError = stats->rootMeanSquareError = stats->nashSutcliffe = stats->R2 = NODATA;
        if (obs.size() > 0) {
            stats->meanAbsoluteError = statistics::meanAbsoluteError(obs, pre);
            stats->meanBiasError = statistics::meanError(obs, pre);
            stats->rootMeanSquareError = statistics::rootMeanSquareError(obs, pre);
            stats->nashSutcliffe = statistics::NashSutcliffeEfficiency(obs, pre);
            float intercept, slope, r2;
            statistics::linearRegression(obs, pre, int(obs.size()), false, &intercept, &slope, &r2);
            stats->R2 = r2;
        }
    }
    return ok ? PB_OK : PB_ERR_INVALID;
}

/* Local-detrending raster: the loop of Project::interpolationDemLocalDetrending (project/project.cpp:3208-3253) restated
 * (its only copy is a method of the Qt-bound Project class); every call inside it is the reference's own code:
 * setMultipleDetrendingHeightTemperatureRange, getUtmXYFromRowCol, getProxyValuesXY, localSelection, preInterpolation,
 * interpolate, clearFitting/setCurrentCombination, updateMinMaxRasterGrid. Rows outside [rowBegin,rowEnd) keep the flag
 * (bounded samples for timing). */
int ref_local_detrending_raster(const pb_stations* st, const pb_settings* s, int variable, const pb_raster* dem,
                                const pb_raster* proxyGrids, int nProxyGrids, int nThreads, int rowBegin, int rowEnd,
                                pb_raster_out* out, double* elapsed_s)
{
    Scenario sc;
    buildSettings(sc, *s, proxyGrids, nProxyGrids);
    buildPoints(sc, *st);
    gis::Crit3DRasterGrid demGrid, outGrid;
    fillGrid(demGrid, *dem);
    sc.settings.setCurrentDEM(&demGrid);
    const meteoVariable myVar = meteoVariable(variable);
    if (!getUseDetrendingVar(myVar) || !sc.settings.getUseLocalDetrending()) return PB_ERR_INVALID;   // project.cpp:3155
    if (rowEnd < 0) rowEnd = demGrid.header->nrRows;
    if (nThreads > 0) omp_set_num_threads(nThreads);

    Crit3DProxyCombination myCombination = sc.settings.getSelectedCombination();
    sc.settings.setCurrentCombination(myCombination);

    gis::Crit3DRasterHeader myHeader = *(demGrid.header);
    outGrid.initializeGrid(myHeader);
    if (!setMultipleDetrendingHeightTemperatureRange(sc.settings)) return PB_ERR_FIT;

    Crit3DInterpolationSettings myInterpolationSettings = sc.settings;
    std::vector<double> proxyValues(myInterpolationSettings.getProxyNr());
    std::vector<Crit3DMeteoPoint> meteoPoints;
    std::vector<Crit3DInterpolationDataPoint>& interpolationPoints = sc.points;
    Crit3DMeteoSettings* meteoSettings = &sc.meteoSettings;
    Crit3DClimateParameters* climateParameters = &sc.climate;
    const Crit3DTime myTime = sc.time;
    int64_t nValid = 0;

    auto t0 = std::chrono::steady_clock::now();
    #pragma omp parallel for if(nThreads != 1) firstprivate(myInterpolationSettings, proxyValues) reduction(+ : nValid) schedule(dynamic, 1)
    for (long row = rowBegin; row < rowEnd; row++)
    {
        std::string errorStdStr;
        for (long col = 0; col < myHeader.nrCols; col++)
        {
            float z = demGrid.value[row][col];
            if (isEqual(z, myHeader.flag))
                continue;
            nValid++;

            double x, y;
            gis::getUtmXYFromRowCol(myHeader, row, col, &x, &y);

            if (getUseDetrendingVar(myVar))
            {
                getProxyValuesXY(x, y, myInterpolationSettings, proxyValues);
            }

            std::vector <Crit3DInterpolationDataPoint> subsetInterpolationPoints;
            localSelection(interpolationPoints, subsetInterpolationPoints, x, y, myInterpolationSettings, false);

            preInterpolation(subsetInterpolationPoints, myInterpolationSettings, meteoSettings, climateParameters,
                             meteoPoints, myVar, myTime, errorStdStr);

            outGrid.value[row][col] = interpolate(subsetInterpolationPoints, myInterpolationSettings, meteoSettings,
                                                  myVar, x, y, z, proxyValues, true);
            myInterpolationSettings.clearFitting();
            myInterpolationSettings.setCurrentCombination(myInterpolationSettings.getSelectedCombination());

            subsetInterpolationPoints.clear();
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (elapsed_s) *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
    bool ok = gis::updateMinMaxRasterGrid(&outGrid);
    out->nValid = nValid;
    copyOut(outGrid, out);
    return ok ? PB_OK : PB_ERR_NO_VALID_CELL;
}

/* ---- glocal detrending (macro areas) ------------------------------------------------------------------------------ */
namespace {
std::vector<Crit3DMacroArea> buildAreas(const pb_macro_area* areas, int nAreas, int nProxy)
{
    std::vector<Crit3DMacroArea> v(nAreas);
    for (int a = 0; a < nAreas; a++) {
        const pb_macro_area& A = areas[a];
        v[a].setMeteoPoints(std::vector<int>(A.meteoPoints, A.meteoPoints + A.nMeteoPoints));
        v[a].setAreaCellsDEM(std::vector<float>(A.areaCellsDEM, A.areaCellsDEM + A.nAreaCells));
        Crit3DProxyCombination c;
        for (int i = 0; i < nProxy; i++) { c.addProxyActive(A.active[i] != 0); c.addProxySignificant(A.significant[i] != 0); }
        c.setUseThermalInversion(A.useThermalInversion != 0);
        v[a].setCombination(c);
        std::vector<std::vector<double>> par;
        for (int i = 0; i < A.nParameters; i++) par.push_back(std::vector<double>(A.parameters[i], A.parameters[i] + A.nrParams[i]));
        v[a].setParameters(par);
    }
    return v;
}

void exportAreas(const std::vector<Crit3DMacroArea>& v, pb_macro_area* areas, int nProxy)
{
    for (size_t a = 0; a < v.size(); a++) {
        pb_macro_area& A = areas[a];
        Crit3DProxyCombination c = v[a].getCombination();
        for (int i = 0; i < PB_MAX_PROXY; i++) {
            A.active[i] = (i < nProxy && i < int(c.getProxySize()) && c.isProxyActive(i)) ? 1 : 0;
            A.significant[i] = (i < nProxy && i < int(c.getProxySize()) && c.isProxySignificant(i)) ? 1 : 0;
        }
        A.useThermalInversion = c.getUseThermalInversion() ? 1 : 0;
        std::vector<std::vector<double>> par = v[a].getParameters();
        A.nParameters = int(par.size());
        for (int i = 0; i < PB_MAX_PROXY; i++) {
            A.nrParams[i] = i < int(par.size()) ? int(par[i].size()) : 0;
            for (int j = 0; j < PB_MAX_FIT_PARAMS; j++) A.parameters[i][j] = (i < int(par.size()) && j < int(par[i].size())) ? par[i][j] : 0.;
        }
    }
}

void buildMeteoPoints(std::vector<Crit3DMeteoPoint>& mp, const pb_stations* meteo, const float* observed, const pb_cv_points* cv)
{
    mp.resize(meteo->n);
    for (int i = 0; i < meteo->n; i++) {
        mp[i].active = meteo->isActive ? meteo->isActive[i] != 0 : true;
        mp[i].topographicDistance = tdMapOf(i);
        mp[i].quality = quality::accepted;
        mp[i].currentValue = observed ? observed[i] : meteo->value[i];
        mp[i].point.utm.x = meteo->x[i];
        mp[i].point.utm.y = meteo->y[i];
        mp[i].point.z = meteo->z[i];
        mp[i].isInsideDem = (cv && cv->isInsideDem) ? cv->isInsideDem[i] != 0 : true;
        mp[i].lapseRateCode = meteo->lapseRateCode ? lapseRateCodeType(meteo->lapseRateCode[i]) : primary;
        for (int k = 0; k < meteo->nProxy; k++) mp[i].proxyValues.push_back(meteo->proxyValues[size_t(i) * meteo->nProxy + k]);
    }
}
}  // namespace

/* preInterpolation with useGlocalDetrending: setMultipleDetrendingHeightTemperatureRange + glocalDetrendingFitting
 * (interpolation.cpp:2466-2474, 2239-2294), verbatim reference code; the areas receive their fit. */
int ref_glocal_detrending_fitting(const pb_stations* st, pb_settings* s, int variable, pb_macro_area* areas, int nAreas)
{
    Scenario sc;
    buildSettings(sc, *s, nullptr, 0);
    buildPoints(sc, *st);
    sc.settings.setMacroAreas(buildAreas(areas, nAreas, s->nProxy));
    std::vector<Crit3DMeteoPoint> meteoPoints;
    std::string err;
    bool ok = preInterpolation(sc.points, sc.settings, &sc.meteoSettings, &sc.climate, meteoPoints, meteoVariable(variable),
                               sc.time, err);
    exportSettings(sc, *s);
    exportAreas(sc.settings.getMacroAreas(), areas, s->nProxy);
    return ok ? PB_OK : PB_ERR_FIT;
}

/* The gridding part of Project::interpolationDemGlocalDetrending (project/project.cpp:3262-3375) restated (a method of the
 * Qt-bound Project class), from `setCurrentCombination(selected)` on: preInterpolation, the second
 * setMultipleDetrendingHeightTemperatureRange, the loop over macro areas. Every call inside is the reference's own code
 * (macroAreaDetrending, getUtmXYFromRowCol, getSignificantProxyValuesXY, interpolate). updateMinMaxRasterGrid is what the
 * callers of interpolationDemMain run on the result. */
static int glocalRasterDriver(const pb_stations* st, pb_settings* s, int variable, pb_macro_area* areas, int nAreas,
                              const pb_raster* dem, const pb_raster* proxyGrids, int nProxyGrids, int nThreads,
                              pb_raster_out* out, double* elapsed_s, const pb_glocal_meteo* gm);

int ref_glocal_detrending_raster(const pb_stations* st, pb_settings* s, int variable, pb_macro_area* areas, int nAreas,
                                 const pb_raster* dem, const pb_raster* proxyGrids, int nProxyGrids, int nThreads,
                                 pb_raster_out* out, double* elapsed_s)
{
    return glocalRasterDriver(st, s, variable, areas, nAreas, dem, proxyGrids, nProxyGrids, nThreads, out, elapsed_s, nullptr);
}

/* the same with Project::meteoPoints filled (what macroAreaDetrending's kh search walks when settings.useTD is on) */
int ref_glocal_detrending_raster_td(const pb_stations* st, pb_settings* s, int variable, pb_macro_area* areas, int nAreas,
                                    const pb_glocal_meteo* gm, const pb_raster* dem, const pb_raster* proxyGrids, int nProxyGrids,
                                    int nThreads, pb_raster_out* out, double* elapsed_s)
{
    return glocalRasterDriver(st, s, variable, areas, nAreas, dem, proxyGrids, nProxyGrids, nThreads, out, elapsed_s, gm);
}

static int glocalRasterDriver(const pb_stations* st, pb_settings* s, int variable, pb_macro_area* areas, int nAreas,
                              const pb_raster* dem, const pb_raster* proxyGrids, int nProxyGrids, int nThreads,
                              pb_raster_out* out, double* elapsed_s, const pb_glocal_meteo* gm)
{
    Scenario sc;
    buildSettings(sc, *s, proxyGrids, nProxyGrids);
    buildPoints(sc, *st);
    gis::Crit3DRasterGrid DEM, outGrid;
    fillGrid(DEM, *dem);
    sc.settings.setCurrentDEM(&DEM);
    sc.settings.setMacroAreas(buildAreas(areas, nAreas, s->nProxy));
    const meteoVariable myVar = meteoVariable(variable);
    Crit3DInterpolationSettings& interpolationSettings = sc.settings;
    if (!getUseDetrendingVar(myVar) || !interpolationSettings.getUseGlocalDetrending()) return PB_ERR_INVALID;
    std::vector<Crit3DMeteoPoint> meteoPoints;
    if (gm && gm->meteo) buildMeteoPoints(meteoPoints, gm->meteo, gm->observed, gm->cv);
    std::vector<Crit3DInterpolationDataPoint>& interpolationPoints = sc.points;
    Crit3DMeteoSettings* meteoSettings = &sc.meteoSettings;
    gis::Crit3DRasterGrid* myRaster = &outGrid;
    const bool _isParallelComputing = nThreads != 1;
    if (nThreads > 0) omp_set_num_threads(nThreads);
    auto t0 = std::chrono::steady_clock::now();

    Crit3DProxyCombination myCombination = interpolationSettings.getSelectedCombination();
    interpolationSettings.setCurrentCombination(myCombination);
    std::string errorStdStr;
    if (!preInterpolation(interpolationPoints, interpolationSettings, meteoSettings, &sc.climate, meteoPoints, myVar, sc.time,
                          errorStdStr))
        return PB_ERR_FIT;
    exportSettings(sc, *s);
    exportAreas(interpolationSettings.getMacroAreas(), areas, s->nProxy);

    bool isOk = true;
    gis::Crit3DRasterHeader myHeader = *(DEM.header);
    myRaster->initializeGrid(myHeader);
    if (!setMultipleDetrendingHeightTemperatureRange(interpolationSettings)) return PB_ERR_FIT;
    exportSettings(sc, *s);     // the state Project::interpolationSettings is left in (the loop works on a copy)

    int elevationPos = NODATA;
    for (unsigned int pos = 0; pos < interpolationSettings.getCurrentCombination().getProxySize(); pos++)
        if (getProxyPragaName(interpolationSettings.getProxy(pos)->getName()) == proxyHeight) elevationPos = pos;

    Crit3DInterpolationSettings myInterpolationSettings = interpolationSettings;
    std::vector<double> proxyValues(myInterpolationSettings.getProxyNr());

    #pragma omp parallel for if (_isParallelComputing) firstprivate(myInterpolationSettings, proxyValues) shared(isOk)
    for (int areaIndex = 0; areaIndex < myInterpolationSettings.getMacroAreasSize(); areaIndex++)
    {
        if (!isOk) continue;
        Crit3DMacroArea macroArea = myInterpolationSettings.getMacroArea(areaIndex);
        std::vector<float> areaCells = macroArea.getAreaCellsDEM();
        std::vector<Crit3DInterpolationDataPoint> subsetInterpolationPoints;
        double interpolatedValue = NODATA;
        if (areaCells.empty()) continue;

        macroAreaDetrending(macroArea, myVar, myInterpolationSettings, meteoSettings, meteoPoints, interpolationPoints,
                            subsetInterpolationPoints, elevationPos);

        for (std::size_t cellIndex = 0; cellIndex < areaCells.size(); cellIndex = cellIndex + 2)
        {
            if (!isOk) break;
            int row = int(areaCells[cellIndex] / DEM.header->nrCols);
            int col = int(areaCells[cellIndex]) % DEM.header->nrCols;
            float z = DEM.value[row][col];
            if (isEqual(z, myHeader.flag)) continue;
            double x, y;
            gis::getUtmXYFromRowCol(myHeader, row, col, &x, &y);
            if (!getSignificantProxyValuesXY(x, y, myInterpolationSettings, proxyValues))
            {
                myRaster->value[row][col] = NODATA;
                continue;
            }
            interpolatedValue = interpolate(subsetInterpolationPoints, myInterpolationSettings, meteoSettings, myVar, x, y, z,
                                            proxyValues, true);
            if (isEqual(interpolatedValue, NODATA)) isOk = false;
            #pragma omp critical
            {
                if (isEqual(myRaster->value[row][col], NODATA))
                    myRaster->value[row][col] = interpolatedValue * areaCells[cellIndex + 1];
                else
                    myRaster->value[row][col] += interpolatedValue * areaCells[cellIndex + 1];
            }
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (elapsed_s) *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
    gis::updateMinMaxRasterGrid(&outGrid);
    int64_t nValid = 0;
    for (int row = 0; row < DEM.header->nrRows; row++)
        for (int col = 0; col < DEM.header->nrCols; col++)
            if (!isEqual(DEM.value[row][col], DEM.header->flag)) nValid++;
    out->nValid = nValid;
    copyOut(outGrid, out);
    return isOk ? PB_OK : PB_ERR_NO_VALID_CELL;
}

/* Project::computeResidualsAndStatisticsGlocalDetrending (project/project.cpp:6525-6565) restated around the reference's
 * computeResidualsGlocalDetrending (spatialControl.cpp:231-302); the areas carry their fit. The per-area statistics
 * (computeStatisticsGlocalCrossValidation) are not part of the residuals and are left out. */
int ref_compute_residuals_glocal(const pb_stations* meteo, const float* observed, const pb_cv_points* cv, const pb_stations* points,
                                 pb_settings* s, int variable, const pb_macro_area* areas, int nAreas, const pb_raster* dem,
                                 int excludeOutsideDem, int excludeSupplemental, float* residual)
{
    Scenario sc;
    buildSettings(sc, *s, nullptr, 0);
    buildPoints(sc, *points);
    gis::Crit3DRasterGrid DEM;
    fillGrid(DEM, *dem);
    sc.settings.setCurrentDEM(&DEM);
    sc.settings.setMacroAreas(buildAreas(areas, nAreas, s->nProxy));
    std::vector<Crit3DMeteoPoint> meteoPoints;
    buildMeteoPoints(meteoPoints, meteo, observed, cv);
    Crit3DInterpolationSettings& interpolationSettings = sc.settings;
    const meteoVariable myVar = meteoVariable(variable);
    if (!setMultipleDetrendingHeightTemperatureRange(interpolationSettings)) return PB_ERR_FIT;      // project.cpp:3069

    int elevationPos = NODATA;
    for (unsigned int pos = 0; pos < interpolationSettings.getCurrentCombination().getProxySize(); pos++)
        if (getProxyPragaName(interpolationSettings.getProxy(pos)->getName()) == proxyHeight) elevationPos = pos;
    for (size_t j = 0; j < meteoPoints.size(); j++) meteoPoints[j].residual = NODATA;
    for (int k = 0; k < interpolationSettings.getMacroAreasSize(); k++)
    {
        Crit3DMacroArea myArea = interpolationSettings.getMacroArea(k);
        std::vector<int> meteoPointsList = myArea.getMeteoPoints();
        if (myArea.getAreaCellsDEM().empty() || meteoPointsList.empty()) continue;
        interpolationSettings.pushMacroAreaNumber(k);
        if (!computeResidualsGlocalDetrending(myVar, myArea, elevationPos, meteoPoints, sc.points, interpolationSettings,
                                              &sc.meteoSettings, excludeOutsideDem != 0, excludeSupplemental != 0))
            return PB_ERR_INVALID;
    }
    for (int i = 0; i < meteo->n; i++) residual[i] = meteoPoints[i].residual;
    return PB_OK;
}

/* One cell of the loop above with everything the fit produced exported (debugging aid for the parity tests):
 * selected station indices in selection order, their regression weights, the detrended values, fitted settings. */
int ref_local_detrending_point(const pb_stations* st, pb_settings* s, int variable, double x, double y, float z,
                               const double* proxyValuesIn, int* selIndex, float* selWeight, float* selValue, int* nSel,
                               float* result)
{
    Scenario sc;
    buildSettings(sc, *s, nullptr, 0);
    buildPoints(sc, *st);
    for (int i = 0; i < st->n; i++) sc.points[i].index = i;
    const meteoVariable myVar = meteoVariable(variable);
    sc.settings.setCurrentCombination(sc.settings.getSelectedCombination());
    if (!setMultipleDetrendingHeightTemperatureRange(sc.settings)) return PB_ERR_FIT;
    std::vector<double> proxyValues(proxyValuesIn, proxyValuesIn + s->nProxy);
    std::vector<Crit3DMeteoPoint> meteoPoints;
    std::string err;
    std::vector<Crit3DInterpolationDataPoint> subset;
    localSelection(sc.points, subset, x, y, sc.settings, false);
    const size_t n0 = subset.size();
    std::vector<int> order;
    for (auto& p : subset) order.push_back(p.index);
    std::vector<float> w0;
    for (auto& p : subset) w0.push_back(p.regressionWeight);
    preInterpolation(subset, sc.settings, &sc.meteoSettings, &sc.climate, meteoPoints, myVar, sc.time, err);
    *result = interpolate(subset, sc.settings, &sc.meteoSettings, myVar, x, y, z, proxyValues, true);
    *nSel = int(n0);
    for (size_t k = 0; k < n0; k++) { selIndex[k] = order[k]; selWeight[k] = w0[k]; selValue[k] = NODATA; }
    for (auto& p : subset)
        for (size_t k = 0; k < n0; k++) if (order[k] == p.index) selValue[k] = p.value;
    exportSettings(sc, *s);
    return PB_OK;
}

/* kriging.cpp:97-295. File-static state => single threaded, one system at a time. */
int ref_kriging_points(const double* pos, const double* val, int n, int mode, double range, double nugget, double sill,
                       double slope, const double* xy, int m, double* result, double* rmse, double* setup_s,
                       double* eval_s)
{
    std::vector<double> p(pos, pos + 2 * size_t(n)), v(val, val + n);
    auto t0 = std::chrono::steady_clock::now();
    if (!krigingVariogram(p.data(), v.data(), n, TkrigingMode(mode), range, nugget, sill, slope)) return PB_ERR_INVALID;
    auto t1 = std::chrono::steady_clock::now();
    for (int i = 0; i < m; i++) {
        krigingSetWeight(xy[2 * i], xy[2 * i + 1]);
        result[i] = krigingResult();
        if (rmse) rmse[i] = krigingRMSE();
    }
    auto t2 = std::chrono::steady_clock::now();
    krigingFreeMemory();
    if (setup_s) *setup_s = std::chrono::duration<double>(t1 - t0).count();
    if (eval_s) *eval_s = std::chrono::duration<double>(t2 - t1).count();
    return PB_OK;
}

}  // extern "C"

/* Crit3DMeteoGrid::spatialAggregateMeteoGrid (meteo/meteoGrid.cpp:1135-1190), verbatim reference code, on a 1 x nCells meteo
 * grid whose cells carry the given aggregation points (what findGridAggregationPoints would have built); value[c] = the
 * cell's currentValue afterwards (NODATA where the reference leaves the cell untouched). */
extern "C" int ref_spatial_aggregate_meteo_grid(const pb_raster* raster, const double* x, const double* y, const int64_t* first,
                                                const int32_t* maxNr, int nCells, int method, float* value)
{
    gis::Crit3DRasterGrid grid;
    fillGrid(grid, *raster);
    Crit3DMeteoGrid mg;
    Crit3DMeteoGridStructure gs;
    gis::Crit3DLatLonHeader h;
    h.nrRows = 1; h.nrCols = nCells; h.dx = 1; h.dy = 1; h.llCorner.latitude = 0; h.llCorner.longitude = 0;
    gs.setHeader(h);
    mg.setGridStructure(gs);
    mg.initMeteoPoints(1, nCells);
    for (int c = 0; c < nCells; c++) {
        Crit3DMeteoPoint* mp = mg.meteoPointPointer(0, c);
        mp->active = true;
        mp->currentValue = NODATA;
        mp->aggregationPointsMaxNr = maxNr[c];
        for (int64_t i = first[c]; i < first[c + 1]; i++) {
            gis::Crit3DPoint p;
            p.utm.x = x[i]; p.utm.y = y[i]; p.z = NODATA;
            mp->aggregationPoints.push_back(p);
        }
    }
    mg.setIsAggregationDefined(true);
    mg.spatialAggregateMeteoGrid(airTemperature, hourly, Crit3DDate(15, 1, 2023), 6, 0, &grid, &grid, aggregationMethod(method));
    for (int c = 0; c < nCells; c++) value[c] = mg.meteoPointPointer(0, c)->currentValue;
    return PB_OK;
}

/* the raster lookup of Project::passInterpolatedTemperatureToHumidityPoints (project/project.cpp:2840-2845): reference calls only */
extern "C" int ref_raster_sample_points(const pb_raster* raster, const double* x, const double* y, int n, float* value, uint8_t* inGrid)
{
    gis::Crit3DRasterGrid grid;
    fillGrid(grid, *raster);
    for (int i = 0; i < n; i++) {
        int row, col;
        grid.getRowCol(x[i], y[i], row, col);
        const bool in = !gis::isOutOfGridRowCol(row, col, grid);
        value[i] = in ? grid.value[row][col] : float(NODATA);
        if (inGrid) inGrid[i] = in;
    }
    return PB_OK;
}

/* topographicIndex (project/interpolationCmd.cpp:397-482), verbatim reference code (extracted at build time like
 * interpolationRaster): the orography-index proxy raster (the demo's TI_DEM_ER_bacini_500m is one) */
bool topographicIndex(const gis::Crit3DRasterGrid& DEM, std::vector <float> windowWidths, gis::Crit3DRasterGrid& outGrid);
extern "C" int ref_topographic_index(const pb_raster* dem, const float* windowWidths, int nWidths, pb_raster_out* out)
{
    gis::Crit3DRasterGrid DEM, outGrid;
    fillGrid(DEM, *dem);
    bool ok = topographicIndex(DEM, std::vector<float>(windowWidths, windowWidths + nWidths), outGrid);
    if (outGrid.isLoaded) copyOut(outGrid, out);
    out->nValid = int64_t(DEM.header->nrRows) * DEM.header->nrCols;
    return ok ? PB_OK : PB_ERR_INVALID;
}

/* radiation::computeRadiationDEM (solarRadiation.cpp:1045-1069) on the reference's own objects: Crit3DRadiationSettings,
 * Crit3DRadiationMaps(dem, gisSettings) (the constructor derives the static lat / lon / slope / aspect maps from the DEM with
 * gis::computeLatLonMaps / computeSlopeAspectMaps; they are returned so that the GPU side is fed the same inputs), the
 * transmissivity map given. s carries the LOCAL time (gisSettings.isUTC = false). Every output array holds nrRows * nrCols
 * floats; minmax = {min, max} of sunElevation, beam, diffuse, reflected, global as updateRadiationMaps leaves them. */
typedef bool (*RadiationDemFn)(Crit3DRadiationSettings*, const gis::Crit3DRasterGrid&, Crit3DRadiationMaps*, const Crit3DTime&, bool);
static int radiationDemDriver(RadiationDemFn compute, const pb_radiation_settings* s, int utmZone, double startLatitude, const pb_raster* dem,
                              const pb_raster* transmissivity, float* lat, float* lon, float* slope, float* aspect,
                              float* sunElevation, float* beam, float* diffuse, float* reflected, float* global,
                              float* minmax, float* demMaximum, int nThreads, double* elapsed_s)
{
    gis::Crit3DRasterGrid demGrid;
    fillGrid(demGrid, *dem);
    gis::updateMinMaxRasterGrid(&demGrid);
    if (demMaximum) *demMaximum = demGrid.maximum;
    gis::Crit3DGisSettings gisSettings;
    gisSettings.utmZone = utmZone;
    gisSettings.isUTC = false;
    gisSettings.timeZone = s->timeZone;
    gisSettings.startLocation.latitude = startLatitude;
    Crit3DRadiationSettings rad;
    rad.setGisSettings(&gisSettings);
    rad.setRealSky(s->realSky != 0);
    rad.setRealSkyAlgorithm(TradiationRealSkyAlgorithm(s->realSkyAlgorithm));
    rad.setShadowing(s->shadowing != 0);
    rad.setTiltMode(TtiltMode(s->tiltMode));
    rad.setLinkeMode(PARAM_MODE_FIXED);
    rad.setAlbedoMode(PARAM_MODE_FIXED);
    rad.setLinkeDefault(s->linke);
    rad.setAlbedo(s->albedo);
    rad.setTilt(s->tilt);
    rad.setAspect(s->aspect);
    rad.setClearSky(s->clearSky);
    Crit3DRadiationMaps maps(demGrid, gisSettings);
    const int rows = demGrid.header->nrRows, cols = demGrid.header->nrCols;
    auto copyGrid = [&](const gis::Crit3DRasterGrid* g, float* dst) {
        if (!dst) return;
        for (int r = 0; r < rows; r++) memcpy(dst + size_t(r) * cols, g->value[r], sizeof(float) * cols);
    };
    copyGrid(maps.latMap, lat);
    copyGrid(maps.lonMap, lon);
    copyGrid(maps.slopeMap, slope);
    copyGrid(maps.aspectMap, aspect);
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < cols; c++)
            maps.transmissivityMap->value[r][c] = transmissivity->data ? transmissivity->data[size_t(r) * cols + c] : transmissivity->rows[r][c];
    Crit3DTime myTime(Crit3DDate(s->day, s->month, s->year), s->secondsOfDay);
    if (nThreads > 0) omp_set_num_threads(nThreads);
    auto t0 = std::chrono::steady_clock::now();
    const bool ok = compute(&rad, demGrid, &maps, myTime, nThreads != 1);
    auto t1 = std::chrono::steady_clock::now();
    if (elapsed_s) *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
    copyGrid(maps.sunElevationMap, sunElevation);
    copyGrid(maps.beamRadiationMap, beam);
    copyGrid(maps.diffuseRadiationMap, diffuse);
    copyGrid(maps.reflectedRadiationMap, reflected);
    copyGrid(maps.globalRadiationMap, global);
    if (minmax) {
        const gis::Crit3DRasterGrid* gs[5] = {maps.sunElevationMap, maps.beamRadiationMap, maps.diffuseRadiationMap, maps.reflectedRadiationMap,
                                              maps.globalRadiationMap};
        for (int k = 0; k < 5; k++) { minmax[2 * k] = gs[k]->minimum; minmax[2 * k + 1] = gs[k]->maximum; }
    }
    return ok ? PB_OK : PB_ERR_INVALID;
}

extern "C" int ref_compute_radiation_dem(const pb_radiation_settings* s, int utmZone, double startLatitude, const pb_raster* dem,
                                         const pb_raster* transmissivity, float* lat, float* lon, float* slope, float* aspect,
                                         float* sunElevation, float* beam, float* diffuse, float* reflected, float* global,
                                         float* minmax, float* demMaximum, int nThreads, double* elapsed_s)
{
    return radiationDemDriver(radiation::computeRadiationDEM, s, utmZone, startLatitude, dem, transmissivity, lat, lon, slope, aspect,
                              sunElevation, beam, diffuse, reflected, global, minmax, demMaximum, nThreads, elapsed_s);
}

