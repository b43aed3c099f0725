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

}  // extern "C"
