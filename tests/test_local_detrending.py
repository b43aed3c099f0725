"""GPU parity of the local-detrending raster (loop of Project::interpolationDemLocalDetrending, project.cpp:3208-3253:
localSelection + per-cell weighted Levenberg-Marquardt multiple detrending + IDW + retrend) against the reference's
own CPU code. Bar: |GPU - reference| <= 1e-4 on every valid cell, NODATA mask bit-exact."""
import ctypes

import numpy as np
import pytest

from praga_b200 import _abi as A
from praga_b200 import synthetic as syn

TOL = 1e-4
pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["pipeline", "fused"])
def schedule(request, monkeypatch):
    """every test of this module runs under both schedules of K2: the three-kernel pipeline (selection -> flat
    Levenberg-Marquardt descents -> pick + estimate; forced here also for small rasters) and the fused kernel"""
    if request.param == "pipeline":
        monkeypatch.setenv("PB_LOCAL_PIPELINE_MIN", "1")
        monkeypatch.delenv("PB_LOCAL_FUSED", raising=False)
    else:
        monkeypatch.setenv("PB_LOCAL_FUSED", "1")
    return request.param


def local_settings(st, min_points=20, use_urban=True):
    s = syn.settings_multiple_detrending(use_urban)
    s.useLocalDetrending = 1
    s.minPointsLocalDetrending = min_points
    syn.set_points_range(s, st)
    return s


def compare(engine, ref, st, s, var, dem, grids, row_begin=0, row_end=-1, tol=TOL):
    okc, gc, mnc, mxc, nvc, _ = ref.local_detrending_raster(st, s, var, dem, grids, threads=8, row_begin=row_begin,
                                                            row_end=row_end)
    out = np.full(dem.shape, np.float32(dem.flag), dtype=np.float32)
    okg, gg, mng, mxg, nvg = engine.local_detrending_raster(st, s, var, dem, grids, row_begin=row_begin, row_end=row_end,
                                                            out=out)
    assert okg == okc and nvg == nvc
    mg, mc = gg == np.float32(dem.flag), gc == np.float32(dem.flag)
    assert np.array_equal(mg, mc), f"NODATA masks differ on {np.count_nonzero(mg != mc)} cells"
    d = np.abs(gg[~mc].astype(np.float64) - gc[~mc].astype(np.float64))
    nbad = int(np.count_nonzero(d > tol))
    assert nbad == 0, f"{nbad} of {d.size} cells differ by more than {tol}; max {d.max():.3e}"
    if okc:
        assert abs(mng - mnc) <= tol and abs(mxg - mxc) <= tol
    return float(d.max()) if d.size else 0.0, float(np.mean(d == 0)) if d.size else 1.0


def test_local_detrending_c3_density(engine, ref):
    """config 3 station density (5000 stations per 1000 km x 1000 km): ~24 stations per cell, ring search 7.5/1 km."""
    dem = syn.make_dem(120, 140, cell_size=5000.0)
    urban = syn.make_urban(120, 140, cell_size=5000.0)
    st = syn.make_stations(dem, 3360, proxies=(dem, urban))
    s = local_settings(st)
    err, exact = compare(engine, ref, st, s, A.airTemperature, dem, (dem, urban))
    print(f"max |gpu-ref| = {err:.2e}, bit-identical cells: {100 * exact:.1f} %")


def test_local_detrending_lapse_codes_and_dense_rings(engine, ref):
    """dense network (first ring already holds > 64 stations -> global-scratch working set), lapse-rate codes on."""
    dem = syn.make_dem(60, 70, cell_size=500.0, seed=3)
    urban = syn.make_urban(60, 70, cell_size=500.0)
    st = syn.make_stations(dem, 400, proxies=(dem, urban), seed=9, supplemental_frac=0.15)
    s = local_settings(st, min_points=25)
    s.useLapseRateCode = 1
    compare(engine, ref, st, s, A.dailyAirTemperatureMax, dem, (dem, urban))


def test_local_detrending_three_piece_elevation(engine, ref):
    """piecewise_three elevation function: 30 first guesses -> one full warp per cell."""
    dem = syn.make_dem(50, 60, cell_size=6000.0, seed=5)
    urban = syn.make_urban(50, 60, cell_size=6000.0)
    st = syn.make_stations(dem, 900, proxies=(dem, urban), seed=2)
    s = local_settings(st)
    el = s.proxy[0]
    el.fittingFunction = A.piecewiseThree
    el.nFitParams = 5
    for j, (lo, hi, fg) in enumerate(zip((0, -30, 100, -0.01, -0.015), (1500, 50, 1000, 0.01, -0.002), (0, 1, 1, 1, 1))):
        el.fitMin[j], el.fitMax[j], el.fitFirstGuess[j] = lo, hi, fg
    s.proxy[0] = el
    compare(engine, ref, st, s, A.airTemperature, dem, (dem, urban))


def test_local_detrending_few_stations_takes_all(engine, ref):
    """fewer stations than minPoints*1.2: localSelection returns every station with its incoming weights."""
    dem = syn.make_dem(40, 40, cell_size=2000.0, seed=8)
    urban = syn.make_urban(40, 40, cell_size=2000.0)
    st = syn.make_stations(dem, 22, proxies=(dem, urban), seed=4)
    s = local_settings(st)
    compare(engine, ref, st, s, A.airTemperature, dem, (dem, urban))


def test_local_detrending_elevation_only_and_row_band(engine, ref):
    dem = syn.make_dem(80, 90, cell_size=4000.0, seed=12)
    st = syn.make_stations(dem, 1200, proxies=(dem,), seed=6)
    s = local_settings(st, use_urban=False)
    compare(engine, ref, st, s, A.dailyAirTemperatureMin, dem, (dem,), row_begin=20, row_end=55)


def test_local_detrending_rejects_non_detrending_variable(engine):
    import praga_b200 as pb
    dem = syn.make_dem(20, 20)
    st = syn.make_stations(dem, 50, proxies=(dem,))
    s = local_settings(st, use_urban=False)
    with pytest.raises(pb.PragaB200Error):
        engine.local_detrending_raster(st, s, A.precipitation, dem, (dem,))


# ---- leave-one-out in local mode: computeResidualsLocalDetrending (spatialControl.cpp:158-228) ------------------------
def loo_cases():
    dem = syn.make_dem(120, 140, cell_size=5000.0)
    urban = syn.make_urban(120, 140, cell_size=5000.0)
    st = syn.make_stations(dem, 1200, proxies=(dem, urban), seed=11)
    yield "c3_density", st, local_settings(st), A.airTemperature
    dem2 = syn.make_dem(60, 70, cell_size=500.0, seed=3)
    urban2 = syn.make_urban(60, 70, cell_size=500.0)
    st2 = syn.make_stations(dem2, 300, proxies=(dem2, urban2), seed=9, supplemental_frac=0.15)
    s2 = local_settings(st2, min_points=25)
    s2.useLapseRateCode = 1
    yield "codes_dense", st2, s2, A.dailyAirTemperatureMax
    st3 = syn.make_stations(dem2, 22, proxies=(dem2, urban2), seed=4)
    yield "takes_all", st3, local_settings(st3), A.airTemperature


def test_local_residuals_identical_to_reference(engine, ref):
    for name, st, s, var in loo_cases():
        use = np.ones(st.n, dtype=np.uint8)
        use[::17] = 0
        res_r, stats_r = ref.compute_residuals(st, s, var, st.value, use_for_cv=use, local=True)
        res_g, stats_g = engine.compute_residuals(st, s, var, st.value, use_for_cv=use, local=True)
        assert np.array_equal(res_r == np.float32(-9999), res_g == np.float32(-9999)), name
        d = np.abs(res_r.astype(np.float64) - res_g.astype(np.float64))
        assert d.max() <= TOL, f"{name}: max |gpu-ref| residual = {d.max():.3e}"
        assert stats_r["n"] == stats_g["n"], name
        for k in ("meanAbsoluteError", "meanBiasError", "rootMeanSquareError", "nashSutcliffe", "R2"):
            assert abs(stats_r[k] - stats_g[k]) <= TOL, f"{name}: {k} {stats_r[k]} vs {stats_g[k]}"
        print(f"{name}: max |gpu-ref| = {d.max():.2e}, bit-identical: {100 * np.mean(d == 0):.1f} %, MAE = {stats_g['meanAbsoluteError']:.4f}")


# ---- the demo's default estimator in local mode: modifiedShepardIdw over the selected stations with radius = getLocalRadius()
# (interpolation.cpp:2521-2527, 948-1028)
def shepard_local(st, **kw):
    s = local_settings(st, **kw)
    s.method = A.shepard_modified
    s.pointsBoundingBoxArea = float((np.float32(st.x.max()) - np.float32(st.x.min())) * (np.float32(st.y.max()) - np.float32(st.y.min())))
    return s


def test_local_detrending_modified_shepard(engine, ref):
    dem = syn.make_dem(90, 100, cell_size=5000.0)
    urban = syn.make_urban(90, 100, cell_size=5000.0)
    st = syn.make_stations(dem, 1900, proxies=(dem, urban), seed=21)
    err, exact = compare(engine, ref, st, shepard_local(st), A.airTemperature, dem, (dem, urban))
    print(f"max |gpu-ref| = {err:.2e}, bit-identical cells: {100 * exact:.1f} %")
    # lapse-rate codes: supplemental stations take part in the fit but weigh nothing in the estimate
    dem2 = syn.make_dem(60, 70, cell_size=500.0, seed=3)
    urban2 = syn.make_urban(60, 70, cell_size=500.0)
    st2 = syn.make_stations(dem2, 400, proxies=(dem2, urban2), seed=9, supplemental_frac=0.15)
    s2 = shepard_local(st2, min_points=25)
    s2.useLapseRateCode = 1
    compare(engine, ref, st2, s2, A.dailyAirTemperatureMax, dem2, (dem2, urban2))
    # fewer stations than minPoints * 1.2: every station, radius = computeShepardInitialRadius (:1098)
    st3 = syn.make_stations(dem2, 22, proxies=(dem2, urban2), seed=4)
    compare(engine, ref, st3, shepard_local(st3), A.airTemperature, dem2, (dem2, urban2))


def test_local_residuals_modified_shepard(engine, ref):
    dem = syn.make_dem(120, 140, cell_size=5000.0)
    urban = syn.make_urban(120, 140, cell_size=5000.0)
    st = syn.make_stations(dem, 1200, proxies=(dem, urban), seed=11)
    s = shepard_local(st)
    res_r, stats_r = ref.compute_residuals(st, s, A.airTemperature, st.value, local=True)
    res_g, stats_g = engine.compute_residuals(st, s, A.airTemperature, st.value, local=True)
    assert np.array_equal(res_r == np.float32(-9999), res_g == np.float32(-9999))
    assert np.abs(res_r.astype(np.float64) - res_g.astype(np.float64)).max() <= TOL
    assert stats_r["n"] == stats_g["n"] and abs(stats_r["meanAbsoluteError"] - stats_g["meanAbsoluteError"]) <= TOL


@pytest.mark.parametrize("area_scale", [1.0, 0.012, 0.004])
def test_local_detrending_plain_shepard(engine, ref, area_scale):
    """shepardIdw behind shepardSearchNeighbour over each cell's subset (interpolation.cpp:806-945). The initial radius comes
    from getPointsBoundingBoxArea() and the SUBSET size: with the stations' real bounding box every selected station lies
    inside it (the 10-nearest branch); smaller areas reach the 5..10 and the 5-nearest branches too."""
    dem = syn.make_dem(90, 100, cell_size=5000.0)
    urban = syn.make_urban(90, 100, cell_size=5000.0)
    st = syn.make_stations(dem, 1900, proxies=(dem, urban), seed=21, supplemental_frac=0.1)
    s = shepard_local(st)
    s.method = A.shepard
    s.useLapseRateCode = 1
    s.pointsBoundingBoxArea = float(np.float32(s.pointsBoundingBoxArea * area_scale))
    err, exact = compare(engine, ref, st, s, A.airTemperature, dem, (dem, urban))
    print(f"area x{area_scale}: max |gpu-ref| = {err:.2e}, bit-identical cells: {100 * exact:.1f} %")
    assert exact >= 0.999
    if area_scale == 1.0:
        res_r, stats_r = ref.compute_residuals(st, s, A.airTemperature, st.value, local=True)
        res_g, stats_g = engine.compute_residuals(st, s, A.airTemperature, st.value, local=True)
        assert np.array_equal(res_r == np.float32(-9999), res_g == np.float32(-9999))
        assert np.abs(res_r.astype(np.float64) - res_g.astype(np.float64)).max() <= TOL


def test_pipeline_and_fused_schedules_give_the_same_bits(engine, monkeypatch):
    """the same raster through both schedules of K2, the pipeline in chunks smaller than the raster: identical bits,
    identical min / max"""
    dem = syn.make_dem(150, 170, cell_size=5000.0, seed=17)
    urban = syn.make_urban(150, 170, cell_size=5000.0)
    st = syn.make_stations(dem, 5400, proxies=(dem, urban), seed=23)
    s = local_settings(st)
    monkeypatch.setenv("PB_LOCAL_FUSED", "1")
    ra = engine.local_detrending_raster(st, s, A.airTemperature, dem, (dem, urban))
    monkeypatch.delenv("PB_LOCAL_FUSED")
    monkeypatch.setenv("PB_LOCAL_PIPELINE_MIN", "1")
    monkeypatch.setenv("PB_LOCAL_CHUNK", "7000")
    rb = engine.local_detrending_raster(st, s, A.airTemperature, dem, (dem, urban))
    assert ra[0] and rb[0]
    assert np.array_equal(ra[1], rb[1])
    assert ra[2:] == rb[2:]
