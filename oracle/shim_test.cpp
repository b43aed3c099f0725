/*
 * oracle/shim_test.cpp — TEST INFRASTRUCTURE. Drives the C++ host shim (praga_b200/host/praga_b200_host.cpp: the
 * reference's own signatures on top of the C ABI) with REAL reference objects built by ref_driver.cpp's helpers, so the
 * drop-in is exercised exactly as a PRAGA maintainer would call it. Linked into oracle/_ref/libpraga_shim_test.so
 * together with the reference libraries; needs libpraga_b200.so (and a B200) at run time.
 */
#include "ref_driver.cpp"

#include "../praga_b200/host/praga_b200_host.h"

extern "C" {

/* pragab200::interpolationRaster / interpolationRasterLocalDetrending on reference objects -> out */
int shim_raster(const pb_stations* st, const pb_settings* s, int variable, const pb_raster* dem, const pb_raster* proxyGrids,
                int nProxyGrids, int local, pb_raster_out* out, char* err, int errLen)
{
    Scenario sc;
    buildSettings(sc, *s, proxyGrids, nProxyGrids);
    buildPoints(sc, *st);
    gis::Crit3DRasterGrid demGrid, outGrid;
    fillGrid(demGrid, *dem);
    sc.settings.setCurrentDEM(&demGrid);
    /* a proxy whose grid IS the DEM (setProxyDEM, project.cpp:184-211): point it at the DEM object itself */
    for (int i = 0; i < s->nProxy; i++)
        if (s->proxy[i].kind == PB_PROXY_HEIGHT && s->proxy[i].gridIndex >= 0) sc.settings.getProxy(i)->setGrid(&demGrid);
    bool ok = local ? pragab200::interpolationRasterLocalDetrending(sc.points, sc.settings, &sc.meteoSettings, &outGrid, demGrid,
                                                                    meteoVariable(variable))
                    : pragab200::interpolationRaster(sc.points, sc.settings, &sc.meteoSettings, &outGrid, demGrid,
                                                     meteoVariable(variable), true);
    if (err && errLen > 0) { strncpy(err, pragab200::lastError().c_str(), errLen - 1); err[errLen - 1] = 0; }
    if (outGrid.isLoaded) copyOut(outGrid, out);
    return ok ? PB_OK : PB_ERR_NO_VALID_CELL;
}

/* pragab200::computeResiduals on reference Crit3DMeteoPoint objects */
int shim_compute_residuals(const pb_stations* st, const pb_settings* s, int variable, const float* observed,
                           const uint8_t* useForCv, float* residual)
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
    }
    bool ok = pragab200::computeResiduals(meteoVariable(variable), mp, sc.points, sc.settings, &sc.meteoSettings, false, false);
    for (int i = 0; i < st->n; i++) residual[i] = mp[i].residual;
    return ok ? PB_OK : PB_ERR_INVALID;
}

/* pragab200::preInterpolation on reference objects (settings, points, meteo points, current DEM); same outputs as
 * ref_pre_interpolation_ex so the two can be compared field by field */
int shim_pre_interpolation(pb_stations* st, pb_settings* s, int variable, const pb_cv_points* cv, const pb_raster* dem,
                           uint8_t* keep, pb_kh_series* khSeries, char* err, int errLen)
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
        mp[i].quality = quality::accepted;
        mp[i].currentValue = (cv && cv->currentValue) ? cv->currentValue[i] : st->value[i];
        mp[i].point.utm.x = st->x[i];
        mp[i].point.utm.y = st->y[i];
        mp[i].point.z = st->z[i];
        mp[i].isInsideDem = (cv && cv->isInsideDem) ? cv->isInsideDem[i] != 0 : true;
        mp[i].lapseRateCode = st->lapseRateCode ? lapseRateCodeType(st->lapseRateCode[i]) : primary;
        for (int k = 0; k < st->nProxy; k++) mp[i].proxyValues.push_back(st->proxyValues[size_t(i) * st->nProxy + k]);
    }
    std::string e;
    bool ok = pragab200::preInterpolation(sc.points, sc.settings, &sc.meteoSettings, &sc.climate, mp, meteoVariable(variable),
                                          sc.time, e);
    if (err && errLen > 0) { strncpy(err, e.c_str(), errLen - 1); err[errLen - 1] = 0; }
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
        std::vector<float> kh = sc.settings.getKh_series(), ke = sc.settings.getKh_error_series();
        khSeries->n = int(std::min(kh.size(), size_t(PB_MAX_KH_SERIES)));
        for (int i = 0; i < khSeries->n; i++) { khSeries->kh[i] = kh[i]; khSeries->error[i] = ke[i]; }
    }
    return ok ? PB_OK : PB_ERR_FIT;
}

/* pragab200::spatialQualityControl on reference Crit3DMeteoPoint objects; same outputs as ref_spatial_quality_control */
int shim_spatial_quality_control(const pb_stations* st, pb_settings* s, int variable, const pb_cv_points* cv, const pb_raster* dem,
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
    bool ok = pragab200::spatialQualityControl(meteoVariable(variable), mp, sc.settings, &sc.meteoSettings, &sc.climate, sc.time, err);
    for (int i = 0; i < st->n; i++) wrongSpatial[i] = mp[i].quality == quality::wrong_spatial ? 1 : 0;
    exportSettings(sc, *s);
    s->pointsBoundingBoxArea = sc.settings.getPointsBoundingBoxArea();
    std::vector<double> pr = sc.settings.getPointsRange();
    if (pr.size() == 2) { s->hasPointsRange = 1; s->pointsRangeMin = pr[0]; s->pointsRangeMax = pr[1]; }
    return ok ? PB_OK : PB_ERR_FIT;
}

/* pragab200::computeResidualsLocalDetrending on reference objects */
int shim_compute_residuals_local(const pb_stations* st, const pb_settings* s, int variable, const float* observed,
                                 const uint8_t* useForCv, float* residual)
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
    bool ok = pragab200::computeResidualsLocalDetrending(meteoVariable(variable), sc.time, mp, sc.points, sc.settings,
                                                         &sc.meteoSettings, &sc.climate, true, true);
    for (int i = 0; i < st->n; i++) residual[i] = mp[i].residual;
    return ok ? PB_OK : PB_ERR_INVALID;
}

/* the stateful kriging API of kriging.h through the shim */
int shim_kriging_points(const double* pos, const double* val, int n, int mode, double range, double nugget, double sill,
                        double slope, const double* xy, int m, double* result, double* rmse)
{
    std::vector<double> p(pos, pos + 2 * size_t(n)), v(val, val + n);
    if (!pragab200::krigingVariogram(p.data(), v.data(), n, short(mode), range, nugget, sill, slope)) return PB_ERR_FIT;
    for (int i = 0; i < m; i++) {
        if (!pragab200::krigingSetWeight(xy[2 * i], xy[2 * i + 1])) return PB_ERR_CUDA;
        result[i] = pragab200::krigingResult();
        rmse[i] = pragab200::krigingRMSE();
    }
    pragab200::krigingFreeMemory();
    return PB_OK;
}

/* pragab200::interpolationRasterGlocalDetrending + computeResidualsGlocalDetrending on reference objects (the macro areas live
 * in the Crit3DInterpolationSettings, like Project keeps them): grid, the areas' fit, the settings and the summed residuals */
int shim_glocal(const pb_stations* st, pb_settings* s, int variable, pb_macro_area* areas, int nAreas, const pb_raster* dem,
                const pb_raster* proxyGrids, int nProxyGrids, pb_raster_out* out, float* residual, char* err, int errLen)
{
    Scenario sc;
    buildSettings(sc, *s, proxyGrids, nProxyGrids);
    buildPoints(sc, *st);
    gis::Crit3DRasterGrid demGrid, outGrid;
    fillGrid(demGrid, *dem);
    sc.settings.setCurrentDEM(&demGrid);
    for (int i = 0; i < s->nProxy; i++)
        if (s->proxy[i].kind == PB_PROXY_HEIGHT && s->proxy[i].gridIndex >= 0) sc.settings.getProxy(i)->setGrid(&demGrid);
    sc.settings.setMacroAreas(buildAreas(areas, nAreas, s->nProxy));
    bool ok = pragab200::interpolationRasterGlocalDetrending(sc.points, sc.settings, &sc.meteoSettings, &outGrid, demGrid,
                                                             meteoVariable(variable));
    if (err && errLen > 0) { strncpy(err, pragab200::lastError().c_str(), errLen - 1); err[errLen - 1] = 0; }
    exportSettings(sc, *s);
    exportAreas(sc.settings.getMacroAreas(), areas, s->nProxy);
    if (outGrid.isLoaded) copyOut(outGrid, out);
    if (residual) {
        std::vector<Crit3DMeteoPoint> mp;
        buildMeteoPoints(mp, st, nullptr, nullptr);
        for (auto& m : mp) m.residual = NODATA;
        int elevationPos = NODATA;
        for (int pos = 0; pos < s->nProxy; pos++) if (s->proxy[pos].kind == PB_PROXY_HEIGHT) elevationPos = pos;
        for (int k = 0; k < sc.settings.getMacroAreasSize(); k++) {          // project.cpp:6545-6562
            Crit3DMacroArea myArea = sc.settings.getMacroArea(k);
            if (myArea.getAreaCellsDEM().empty() || myArea.getMeteoPoints().empty()) continue;
            if (!pragab200::computeResidualsGlocalDetrending(meteoVariable(variable), myArea, elevationPos, mp, sc.points, sc.settings,
                                                             &sc.meteoSettings, true, true))
                return PB_ERR_INVALID;
        }
        for (int i = 0; i < st->n; i++) residual[i] = mp[i].residual;
    }
    return ok ? PB_OK : PB_ERR_NO_VALID_CELL;
}

/* pragab200::krigingEstimateVariogram (the signature the reference declares, interpolation.h:65-68) and
 * pragab200::preInterpolation with useGlocalDetrending on reference objects */
int shim_estimate_variogram(float* dist, float* semiVar, int n, float maxDistance, int* mode, double* sill, double* nugget,
                            double* range, double* slope)
{
    TkrigingMode m = TkrigingMode(*mode);
    bool ok = pragab200::krigingEstimateVariogram(dist, semiVar, n, n, maxDistance, sill, nugget, range, slope, &m, n);
    *mode = int(m);
    return ok ? PB_OK : PB_ERR_FIT;
}

int shim_pre_interpolation_glocal(pb_stations* st, pb_settings* s, int variable, pb_macro_area* areas, int nAreas, char* err, int errLen)
{
    Scenario sc;
    buildSettings(sc, *s, nullptr, 0);
    buildPoints(sc, *st);
    sc.settings.setMacroAreas(buildAreas(areas, nAreas, s->nProxy));
    std::vector<Crit3DMeteoPoint> meteoPoints;
    std::string e;
    bool ok = pragab200::preInterpolation(sc.points, sc.settings, &sc.meteoSettings, &sc.climate, meteoPoints, meteoVariable(variable),
                                          sc.time, e);
    if (err && errLen > 0) { strncpy(err, e.c_str(), errLen - 1); err[errLen - 1] = 0; }
    exportSettings(sc, *s);
    exportAreas(sc.settings.getMacroAreas(), areas, s->nProxy);
    return ok ? PB_OK : PB_ERR_FIT;
}

/* ---- a host SESSION as PRAGA runs one: the DEM, the proxy grids and the output grid are gis::Crit3DRasterGrid objects
 * (float** rows) that live across time steps; every step brings new detrended points and settings and calls
 * pragab200::interpolationRaster. Used by tests (cached vs uncached grids must agree) and by bench.py to time the drop-in
 * through the reference's own data layout: the clock runs around the pragab200 call only; building the reference's point
 * objects from the flat arrays is the harness' work, not the drop-in's. */
struct ShimSession {
    Scenario sc;
    gis::Crit3DRasterGrid dem, out;
    pb_settings flat;
};

void* shim_session_open(const pb_settings* s, const pb_raster* dem, const pb_raster* proxyGrids, int nProxyGrids)
{
    ShimSession* ss = new ShimSession();
    ss->flat = *s;
    buildSettings(ss->sc, *s, proxyGrids, nProxyGrids);
    fillGrid(ss->dem, *dem);
    ss->sc.settings.setCurrentDEM(&ss->dem);
    for (int i = 0; i < s->nProxy; i++)
        if (s->proxy[i].kind == PB_PROXY_HEIGHT && s->proxy[i].gridIndex >= 0) ss->sc.settings.getProxy(i)->setGrid(&ss->dem);
    return ss;
}

/* one time step: returns PB_OK / PB_ERR_NO_VALID_CELL; *elapsed_s = wall time of the pragab200::interpolationRaster call */
int shim_session_step(void* session, const pb_stations* st, int variable, int local, double* elapsed_s, float* minimum, float* maximum)
{
    ShimSession* ss = static_cast<ShimSession*>(session);
    for (auto& p : ss->sc.points) delete p.point;
    ss->sc.points.clear();
    buildPoints(ss->sc, *st);
    auto t0 = std::chrono::steady_clock::now();
    bool ok = local ? pragab200::interpolationRasterLocalDetrending(ss->sc.points, ss->sc.settings, &ss->sc.meteoSettings, &ss->out,
                                                                    ss->dem, meteoVariable(variable))
                    : pragab200::interpolationRaster(ss->sc.points, ss->sc.settings, &ss->sc.meteoSettings, &ss->out, ss->dem,
                                                     meteoVariable(variable), true);
    auto t1 = std::chrono::steady_clock::now();
    if (elapsed_s) *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
    if (minimum) *minimum = ss->out.minimum;
    if (maximum) *maximum = ss->out.maximum;
    return ok ? PB_OK : PB_ERR_NO_VALID_CELL;
}

/* the session's output grid -> a flat array; modifyDem != 0: the DEM cell (row, col) is overwritten IN PLACE first (tests of
 * the cache's invalidation rules) */
int shim_session_output(void* session, pb_raster_out* out)
{
    ShimSession* ss = static_cast<ShimSession*>(session);
    if (!ss->out.isLoaded) return PB_ERR_INVALID;
    copyOut(ss->out, out);
    return PB_OK;
}
void shim_session_poke_dem(void* session, int row, int col, float value, int announce)
{
    ShimSession* ss = static_cast<ShimSession*>(session);
    ss->dem.value[row][col] = value;
    if (announce) pragab200::invalidate(&ss->dem);
}
int shim_session_set_device(int device) { return pragab200::setDevice(device) ? PB_OK : PB_ERR_INVALID; }
int shim_session_timing(pb_timing* t) { return pragab200::lastTiming(t) ? PB_OK : PB_ERR_INVALID; }
void shim_session_cache(int enabled) { pragab200::setRasterCache(enabled != 0); }
void shim_session_close(void* session)
{
    ShimSession* ss = static_cast<ShimSession*>(session);
    pragab200::invalidate(&ss->dem);
    for (auto g : ss->sc.grids) pragab200::invalidate(g);
    delete ss;
}

/* pragab200::computeRadiationDEM on the reference's own Crit3DRadiationSettings / Crit3DRadiationMaps objects (same driver as
 * ref_compute_radiation_dem, the drop-in in place of radiation::computeRadiationDEM) */
int shim_compute_radiation_dem(const pb_radiation_settings* s, int utmZone, double startLatitude, const pb_raster* dem,
                               const pb_raster* transmissivity, float* lat, float* lon, float* slope, float* aspect, float* sunElevation,
                               float* beam, float* diffuse, float* reflected, float* global, float* minmax, float* demMaximum, int nThreads,
                               double* elapsed_s)
{
    const int rc = radiationDemDriver(pragab200::computeRadiationDEM, s, utmZone, startLatitude, dem, transmissivity, lat, lon, slope, aspect,
                                      sunElevation, beam, diffuse, reflected, global, minmax, demMaximum, nThreads, elapsed_s);
    pragab200::invalidateAll();          // the driver's grids die with this call
    return rc;
}

}  // extern "C"
