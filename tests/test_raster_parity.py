"""GPU parity of interpolationRaster (interpolationCmd.cpp:353-394) against the reference's own CPU code.

Bar (BASELINE.json north_star): |GPU - reference| <= 1e-4 in the variable's units on every valid cell,
NODATA/flag masks bit-exact."""
import copy

import numpy as np
import pytest

from praga_b200 import _abi as A
from praga_b200 import synthetic as syn

TOL = 1e-4
pytestmark = pytest.mark.gpu


def assert_grid_parity(gpu, cpu, flag=-9999.0, tol=TOL):
    assert gpu.shape == cpu.shape
    mg, mc = gpu == np.float32(flag), cpu == np.float32(flag)
    assert np.array_equal(mg, mc), f"NODATA masks differ on {np.count_nonzero(mg != mc)} cells"
    d = np.abs(gpu[~mc].astype(np.float64) - cpu[~mc].astype(np.float64))
    assert d.size == 0 or d.max() <= tol, f"max |gpu-ref| = {d.max():.3e} > {tol}"
    return 0.0 if d.size == 0 else float(d.max())


def run_both(engine, ref, st, s, var, dem, grids):
    okc, gc, mnc, mxc, nvc, _ = ref.interpolation_raster(st, s, var, dem, grids)
    okg, gg, mng, mxg, nvg = engine.interpolation_raster(st, s, var, dem, grids)
    assert okg == okc
    assert nvg == nvc
    err = assert_grid_parity(gg, gc, dem.flag)
    if okc:
        assert abs(mng - mnc) <= TOL and abs(mxg - mxc) <= TOL
    return err, gg, gc


def test_multiple_detrending_idw_temperature(engine, ref):
    """config 2 shape (scaled down): IDW + multipleDetrending, elevation (piecewise two) + urban fraction."""
    dem = syn.make_dem(300, 400)
    urban = syn.make_urban(300, 400)
    st = syn.make_stations(dem, 200, proxies=(dem, urban))
    s = syn.settings_multiple_detrending()
    syn.set_points_range(s, st)
    ok, keep = ref.pre_interpolation(st, s, A.airTemperature, (dem, urban))
    assert ok and s.nFit == 2
    run_both(engine, ref, st.subset(keep), s, A.airTemperature, dem, (dem, urban))


def test_foreign_geometry_proxy_grid(engine, ref):
    """urban proxy on a coarser 250 m grid that does not cover the whole DEM: lookups fall outside -> retrend 0."""
    dem = syn.make_dem(240, 260)
    urban = syn.make_urban(80, 90, cell_size=250.0, llx=syn.LLX + 1300.0, lly=syn.LLY + 700.0, nodata_frac=0.03)
    st = syn.make_stations(dem, 150, proxies=(dem, urban))
    s = syn.settings_multiple_detrending()
    syn.set_points_range(s, st)
    ok, keep = ref.pre_interpolation(st, s, A.airTemperature, (dem, urban))
    assert ok
    run_both(engine, ref, st.subset(keep), s, A.airTemperature, dem, (dem, urban))


@pytest.mark.parametrize("inversion", [False, True])
def test_classic_detrending_idw(engine, ref, inversion):
    """classic detrending(): elevation regression (with/without thermal inversion heuristic) + urban regression."""
    dem = syn.make_dem(200, 300, seed=3)
    urban = syn.make_urban(200, 300)
    st = syn.make_stations(dem, 180, proxies=(dem, urban), seed=5)
    if inversion:   # cold-air pool below 600 m
        st.value[:] = np.where(st.z < 600, st.value - 0.01 * (600 - st.z), st.value).astype(np.float32)
    s = syn.settings_classic_detrending(2)
    s.useThermalInversion = int(inversion)
    s.selectedUseThermalInversion = s.currentUseThermalInversion = int(inversion)
    ok, keep = ref.pre_interpolation(st, s, A.dailyAirTemperatureMax, (dem, urban))
    assert ok
    assert s.currentSignificant[0] == 1
    run_both(engine, ref, st.subset(keep), s, A.dailyAirTemperatureMax, dem, (dem, urban))


def test_no_detrending_humidity_clamped(engine, ref):
    dem = syn.make_dem(180, 220, seed=9)
    st = syn.make_stations(dem, 120, variable="humidity", seed=6)
    st.value[:10] = 100.0
    s = A.default_settings()
    _, gg, _ = run_both(engine, ref, st, s, A.airRelHumidity, dem, ())
    v = gg[gg != np.float32(dem.flag)]
    assert v.min() >= 0.0 and v.max() <= 100.0


def test_precipitation_threshold_and_all_zero(engine, ref):
    dem = syn.make_dem(150, 170, seed=4)
    st = syn.make_stations(dem, 90, variable="precipitation", seed=8)
    s = A.default_settings()
    ok, _ = ref.pre_interpolation(st, s, A.precipitation, ())
    assert ok and s.precipitationAllZero == 0
    _, gg, _ = run_both(engine, ref, st, s, A.precipitation, dem, ())
    v = gg[gg != np.float32(dem.flag)]
    assert np.all((v == 0.0) | (v >= s.rainfallThreshold))
    # all observations below the threshold -> interpolate() short-circuits to exactly 0.f (interpolation.cpp:2506)
    st0 = st.copy()
    st0.value[:] = 0.1
    s0 = A.default_settings()
    ok, _ = ref.pre_interpolation(st0, s0, A.precipitation, ())
    assert ok and s0.precipitationAllZero == 1
    _, gg0, _ = run_both(engine, ref, st0, s0, A.precipitation, dem, ())
    assert np.all(gg0[gg0 != np.float32(dem.flag)] == 0.0)


def test_lapse_rate_code_excludes_supplemental(engine, ref):
    dem = syn.make_dem(120, 140, seed=2)
    st = syn.make_stations(dem, 100, variable="humidity", seed=3, supplemental_frac=0.2)
    s = A.default_settings()
    s.useLapseRateCode = 1
    run_both(engine, ref, st, s, A.airRelHumidity, dem, ())


def test_station_exactly_on_cell_coordinate(engine, ref):
    """a station sitting exactly on a cell's (x, y) has distance 0 and is ignored (interpolation.cpp:1039)."""
    dem = syn.make_dem(64, 64, nodata_frac=0.0, border=False)
    st = syn.make_stations(dem, 40, variable="humidity", seed=1)
    # cell (row 10, col 20): x = llx + cs*col + 0.5, y = lly + cs*(rows-row) - 0.5 (gis.cpp:799-804), both exact in f32
    st.x[0] = dem.llx + dem.cell_size * 20 + 0.5
    st.y[0] = dem.lly + dem.cell_size * (64 - 10) - 0.5
    st.value[0] = 5.0
    s = A.default_settings()
    _, gg, gc = run_both(engine, ref, st, s, A.airRelHumidity, dem, ())
    assert abs(gg[10, 20] - gc[10, 20]) <= TOL


def test_all_flag_dem_returns_false(engine, ref):
    dem = syn.make_dem(32, 32)
    dem.data[:] = dem.flag
    st = syn.make_stations(syn.make_dem(32, 32), 20, variable="humidity")
    s = A.default_settings()
    okc, gc, *_ = ref.interpolation_raster(st, s, A.airRelHumidity, dem, ())
    okg, gg, *_ = engine.interpolation_raster(st, s, A.airRelHumidity, dem, ())
    assert okc is False and okg is False
    assert np.all(gg == np.float32(dem.flag))


def test_empty_station_set_writes_nodata(engine, ref):
    dem = syn.make_dem(40, 40)
    st = syn.make_stations(dem, 5, variable="humidity").subset(np.zeros(5, dtype=bool))
    s = A.default_settings()
    okc, gc, *_ = ref.interpolation_raster(st, s, A.airRelHumidity, dem, ())
    okg, gg, *_ = engine.interpolation_raster(st, s, A.airRelHumidity, dem, ())
    assert okg == okc
    assert np.array_equal(gg, gc)


def test_ragged_shapes_and_many_stations(engine, ref):
    """non-multiple-of-tile station count (> 1 tile) and a raster whose size is not a multiple of anything."""
    dem = syn.make_dem(97, 131, seed=12)
    st = syn.make_stations(dem, 1037, variable="humidity", seed=13)
    s = A.default_settings()
    run_both(engine, ref, st, s, A.dailyAirRelHumidityAvg, dem, ())


def test_row_band_matches_full_raster(engine, ref):
    dem = syn.make_dem(120, 90, seed=21)
    st = syn.make_stations(dem, 77, variable="humidity", seed=22)
    s = A.default_settings()
    _, full, *_ = engine.interpolation_raster(st, s, A.airRelHumidity, dem, ())
    parts = np.full(dem.shape, np.float32(123.0), dtype=np.float32)
    for b0, b1 in ((0, 37), (37, 80), (80, 120)):
        engine.interpolation_raster(st, s, A.airRelHumidity, dem, (), row_begin=b0, row_end=b1, out=parts)
    assert np.array_equal(parts, full)


def test_float_row_pointer_rasters(engine, ref):
    """the reference's own layout: one allocation per row (float**), gis.cpp:301-317."""
    import ctypes as C
    dem = syn.make_dem(50, 60, seed=31)
    st = syn.make_stations(dem, 33, variable="humidity", seed=32)
    s = A.default_settings()
    _, want, *_ = engine.interpolation_raster(st, s, A.airRelHumidity, dem, ())
    rows_in = [np.ascontiguousarray(dem.data[r]).copy() for r in range(50)]
    rows_out = [np.zeros(60, dtype=np.float32) for _ in range(50)]
    RowPtrs = C.POINTER(C.c_float) * 50
    d = A.pb_raster()
    d.h = dem.header()
    d.rows = RowPtrs(*[A.fptr(r) for r in rows_in])
    o = A.pb_raster_out()
    o.rows = RowPtrs(*[A.fptr(r) for r in rows_out])
    stp = st.as_pb()
    rc = engine.lib.pb_interpolation_raster(engine.h, C.byref(stp), C.byref(s), A.airRelHumidity, C.byref(d), None, 0, 0, -1,
                                            C.byref(o))
    assert rc == A.PB_OK
    assert np.array_equal(np.stack(rows_out), want)


def test_points_leave_one_out(engine, ref):
    """interpolate() at the stations themselves: the station's own zero distance is skipped (computeResiduals)."""
    dem = syn.make_dem(200, 200, seed=41)
    urban = syn.make_urban(200, 200)
    st = syn.make_stations(dem, 300, proxies=(dem, urban), seed=42)
    s = syn.settings_multiple_detrending()
    syn.set_points_range(s, st)
    ok, keep = ref.pre_interpolation(st, s, A.airTemperature, (dem, urban))
    st = st.subset(keep)
    x, y, z = st.x.astype(np.float32), st.y.astype(np.float32), st.z.astype(np.float32)
    pv = st.proxy_values.astype(np.float64)
    want = ref.interpolate_points(st, s, A.airTemperature, x, y, z, pv, False)
    got = engine.interpolate_points(st, s, A.airTemperature, x, y, pv, False)
    assert np.abs(got.astype(np.float64) - want).max() <= TOL


@pytest.mark.parametrize("variable", ["temperature", "precipitation"])
def test_compute_residuals_and_cv_statistics(engine, ref, variable):
    """computeResiduals (spatialControl.cpp:97-155) + computeStatisticsCrossValidation (project.cpp:2935-2969)."""
    dem = syn.make_dem(200, 200, seed=51)
    urban = syn.make_urban(200, 200)
    if variable == "temperature":
        st = syn.make_stations(dem, 400, proxies=(dem, urban), seed=52)
        raw = st.value.copy()
        s = syn.settings_multiple_detrending()
        syn.set_points_range(s, st)
        var = A.airTemperature
        ok, keep = ref.pre_interpolation(st, s, var, (dem, urban))
        assert ok and keep.all()
    else:
        st = syn.make_stations(dem, 300, variable="precipitation", seed=53)
        raw = st.value.copy()
        s = A.default_settings()
        var = A.dailyPrecipitation
        ok, _ = ref.pre_interpolation(st, s, var, ())
    use = np.ones(st.n, dtype=np.uint8)
    use[::7] = 0
    want, ws = ref.compute_residuals(st, s, var, raw, use)
    got, gs = engine.compute_residuals(st, s, var, raw, use)
    assert np.array_equal(got == np.float32(-9999), want == np.float32(-9999))
    m = want != np.float32(-9999)
    if variable == "precipitation":
        # the rain threshold is a discontinuity: an estimate within 1e-4 of it may fall on the other side
        close = np.abs(got[m].astype(np.float64) - want[m]) <= TOL
        assert close.mean() > 0.99
    else:
        assert np.abs(got[m].astype(np.float64) - want[m]).max() <= TOL
        assert gs["n"] == ws["n"]
        for k in ("meanAbsoluteError", "meanBiasError", "rootMeanSquareError", "nashSutcliffe", "R2"):
            assert abs(gs[k] - ws[k]) <= 1e-4, (k, gs[k], ws[k])


def test_topographic_distance_raster_and_points(engine, ref):
    """IDW with topographic distance (kh != 0): the ray march through the DEM is reproduced step by step."""
    dem = syn.make_dem(110, 130, seed=6, cell_size=400.0)
    urban = syn.make_urban(110, 130, cell_size=400.0)
    st = syn.make_stations(dem, 90, proxies=(dem, urban), seed=7)
    s = syn.settings_classic_detrending(2)
    ok, keep = ref.pre_interpolation(st, s, A.dailyAirTemperatureMax, (dem, urban))
    st = st.subset(keep)
    s.useTD = 1
    s.topoDist_Kh = 12
    err, gg, gc = run_both(engine, ref, st, s, A.dailyAirTemperatureMax, dem, (dem, urban))
    s0 = copy.copy(s)
    okc, g0, *_ = ref.interpolation_raster(st, A.pb_settings.from_buffer_copy(s), A.dailyAirTemperatureMax, dem, (dem, urban))
    # points: leave-one-out at the stations with their own elevation
    x, y, z = st.x.astype(np.float32), st.y.astype(np.float32), st.z.astype(np.float32)
    pv = st.proxy_values.astype(np.float64)
    want = ref.interpolate_points_td(st, s, A.dailyAirTemperatureMax, x, y, z, pv, dem)
    did = engine.upload(dem)
    got = engine.interpolate_points(st, s, A.dailyAirTemperatureMax, x, y, pv, False, z=z, dem_id=did)
    engine.free(did)
    assert np.abs(got.astype(np.float64) - want).max() <= TOL
    # precipitation does not use topographic distance (getUseTdVar, interpolation.cpp:1220-1234)
    stp = syn.make_stations(dem, 60, variable="precipitation", seed=9)
    sp = A.default_settings()
    sp.useTD = 1
    sp.topoDist_Kh = 12
    run_both(engine, ref, stp, sp, A.precipitation, dem, ())


@pytest.mark.gpu
def test_registered_host_buffers_give_the_same_grid(engine, ref):
    """pb_host_register: page-locked DEM / proxy / output buffers are copied straight from / into the caller's memory by the
    band pipeline of the drop-in call; the grid must be bit-identical to the one that went through the staging ring."""
    dem = syn.make_dem(900, 1000)
    urban = syn.make_urban(900, 1000)
    st = syn.make_stations(dem, 300, proxies=(dem, urban))
    s = syn.settings_multiple_detrending()
    syn.set_points_range(s, st)
    ok, keep = ref.pre_interpolation(st, s, A.airTemperature, (dem, urban))
    st = st.subset(keep)
    out_a = np.empty(dem.shape, dtype=np.float32)
    out_b = np.full(dem.shape, 123.0, dtype=np.float32)
    ra = engine.interpolation_raster(st, s, A.airTemperature, dem, (dem, urban), out=out_a)
    for a in (dem.data, urban.data, out_b):
        engine.host_register(a)
    try:
        rb = engine.interpolation_raster(st, s, A.airTemperature, dem, (dem, urban), out=out_b)
        rc = engine.interpolation_raster(st, s, A.airTemperature, dem, (dem, urban), out=out_b, row_begin=100, row_end=650)
    finally:
        for a in (dem.data, urban.data, out_b):
            engine.host_unregister(a)
    assert ra[0] and rb[0] and ra[2:] == rb[2:]
    assert np.array_equal(out_a, out_b)
    assert rc[0]


@pytest.mark.gpu
def test_eight_distinct_proxy_grids_through_the_band_pipeline(engine, ref):
    """PB_MAX_PROXY = 8 proxies, every one with its own same-geometry grid that is NOT the DEM buffer, pageable buffers: the
    band pipeline stages 9 rasters (DEM + 8 grids) per chunk through a ring of 8 pinned slots — they must never share a slot
    (ADVICE r1). Fitted model given directly: elevation (piecewise two) + 7 linear terms."""
    rows, cols = 640, 520
    dem = syn.make_dem(rows, cols, seed=21)
    elev = A.Raster(dem.data.copy(), dem.llx, dem.lly, dem.cell_size, dem.flag)      # the elevation proxy's own grid buffer
    others = [syn.make_urban(rows, cols, seed=40 + k) for k in range(7)]
    grids = (elev,) + tuple(others)
    st = syn.make_stations(dem, 220, proxies=(dem,) + tuple(others), seed=9)
    kinds = (A.proxyHeight, A.proxyUrbanFraction, A.proxyOrogIndex, A.proxySeaDistance, A.proxyAspect, A.proxySlope,
             A.proxyWaterIndex, A.proxyUrbanFraction)
    s = A.default_settings()
    s.useMultipleDetrending = 1
    s.nProxy = 8
    for i, kind in enumerate(kinds):
        p = A.default_proxy(kind, i)
        p.fittingFunction = A.piecewiseTwo if i == 0 else A.linear
        s.proxy[i] = p
        s.selectedActive[i] = s.currentActive[i] = s.currentSignificant[i] = 1
    s.nFit = s.nFitFunctions = 8
    s.fitFunction[0], s.fitNrParams[0] = A.piecewiseTwo, 4
    for j, v in enumerate((420.0, 12.5, 0.004, -0.0062)):
        s.fitParams[0][j] = v
    for k in range(1, 8):
        s.fitFunction[k], s.fitNrParams[k] = A.linear, 2
        s.fitParams[k][0], s.fitParams[k][1] = 0.3 * k, -0.1 * k
    st.value[:] = (st.value - 12.0).astype(np.float32)
    err, gg, gc = run_both(engine, ref, st, s, A.airTemperature, dem, grids)
    # the retrend term of every proxy is visible in the result: a swapped / mixed staging buffer cannot hide below 1e-4
    assert np.ptp(gc[gc != np.float32(dem.flag)]) > 1.0
