"""The C++ host shim (praga_b200/host/praga_b200_host.cpp) keeps the reference's own signatures — interpolationRaster(),
computeResiduals(), kriging.h — on top of the C ABI. oracle/_ref/libpraga_shim_test.so drives it with REAL reference
objects (Crit3DInterpolationSettings, Crit3DInterpolationDataPoint, gis::Crit3DRasterGrid, Crit3DMeteoPoint) and the
results are compared with the reference's CPU code: this is the drop-in as a PRAGA maintainer would use it."""
import ctypes as C

import numpy as np
import pytest

from praga_b200 import _abi as A
from praga_b200 import synthetic as syn

TOL = 1e-4
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def shim(built_lib):
    from oracle import binding
    import os
    if os.path.isdir("/root/reference/agrolib"):
        binding.build("shim")
    if not binding.have_shim():
        pytest.skip("oracle/_ref/libpraga_shim_test.so not available")
    return binding.load_shim()


def shim_raster(shim, st, s, var, dem, grids, local=False):
    stp, d = st.as_pb(), dem.as_pb()
    g = (A.pb_raster * max(1, len(grids)))()
    for i, r in enumerate(grids):
        g[i] = r.as_pb()
    out = np.empty(dem.shape, dtype=np.float32)
    o = A.pb_raster_out()
    o.data = A.fptr(out)
    err = C.create_string_buffer(512)
    rc = shim.shim_raster(C.byref(stp), C.byref(s), var, C.byref(d), g, len(grids), int(local), C.byref(o), err, 512)
    return rc == A.PB_OK, out, o.minimum, o.maximum, err.value.decode()


def test_shim_interpolation_raster_multiple_detrending(shim, ref):
    dem = syn.make_dem(160, 190)
    urban = syn.make_urban(160, 190)
    st = syn.make_stations(dem, 150, proxies=(dem, urban))
    s = syn.settings_multiple_detrending()
    syn.set_points_range(s, st)
    ok, keep = ref.pre_interpolation(st, s, A.airTemperature, (dem, urban))
    st = st.subset(keep)
    okc, gc, mnc, mxc, _, _ = ref.interpolation_raster(st, s, A.airTemperature, dem, (dem, urban))
    okg, gg, mng, mxg, err = shim_raster(shim, st, s, A.airTemperature, dem, (dem, urban))
    assert okg == okc, err
    assert np.array_equal(gg == np.float32(dem.flag), gc == np.float32(dem.flag))
    m = gc != np.float32(dem.flag)
    assert np.abs(gg[m].astype(np.float64) - gc[m]).max() <= TOL
    assert abs(mng - mnc) <= TOL and abs(mxg - mxc) <= TOL


def test_shim_local_detrending_raster(shim, ref):
    dem = syn.make_dem(50, 60, cell_size=8000.0)
    urban = syn.make_urban(50, 60, cell_size=8000.0)
    st = syn.make_stations(dem, 1100, proxies=(dem, urban))
    s = syn.settings_multiple_detrending()
    s.useLocalDetrending = 1
    s.minPointsLocalDetrending = 20
    syn.set_points_range(s, st)
    okc, gc, *_ = ref.local_detrending_raster(st, s, A.airTemperature, dem, (dem, urban), threads=8)
    okg, gg, _, _, err = shim_raster(shim, st, s, A.airTemperature, dem, (dem, urban), local=True)
    assert okg == okc, err
    assert np.array_equal(gg == np.float32(dem.flag), gc == np.float32(dem.flag))
    m = gc != np.float32(dem.flag)
    assert np.abs(gg[m].astype(np.float64) - gc[m]).max() <= TOL


def test_shim_all_flag_dem_returns_false(shim, ref):
    dem = syn.make_dem(24, 24)
    dem.data[:] = dem.flag
    st = syn.make_stations(syn.make_dem(24, 24), 20, variable="humidity")
    ok, gg, *_ = shim_raster(shim, st, A.default_settings(), A.airRelHumidity, dem, ())
    assert ok is False and np.all(gg == np.float32(dem.flag))


def test_shim_compute_residuals(shim, ref):
    dem = syn.make_dem(150, 150, seed=61)
    urban = syn.make_urban(150, 150)
    st = syn.make_stations(dem, 250, proxies=(dem, urban), seed=62)
    raw = st.value.copy()
    s = syn.settings_multiple_detrending()
    syn.set_points_range(s, st)
    ok, keep = ref.pre_interpolation(st, s, A.airTemperature, (dem, urban))
    assert keep.all()
    want, _ = ref.compute_residuals(st, s, A.airTemperature, raw)
    got = np.empty(st.n, dtype=np.float32)
    stp = st.as_pb()
    rc = shim.shim_compute_residuals(C.byref(stp), C.byref(s), A.airTemperature, A.fptr(raw), None, A.fptr(got))
    assert rc == A.PB_OK
    assert np.abs(got.astype(np.float64) - want).max() <= TOL


def test_shim_kriging_stateful_api(shim, ref):
    rng = np.random.default_rng(8)
    n, m = 120, 40
    pos = rng.random((n, 2)) * 2.0e5 + np.array([5.0e5, 4.8e6])
    val = rng.normal(10, 3, n)
    xy = rng.random((m, 2)) * 2.0e5 + np.array([5.0e5, 4.8e6])
    ok, want, want_rm, _, _ = ref.kriging_points(pos, val, A.KRIGING_SPHERICAL, 7.0e4, 0.1, 4.0, 0.0, xy)
    res, rm = np.empty(m), np.empty(m)
    p, v, q = pos.reshape(-1).copy(), val.copy(), xy.reshape(-1).copy()
    rc = shim.shim_kriging_points(A.dptr(p), A.dptr(v), n, A.KRIGING_SPHERICAL, 7.0e4, 0.1, 4.0, 0.0, A.dptr(q), m, A.dptr(res),
                                  A.dptr(rm))
    assert rc == A.PB_OK
    assert np.abs(res - want).max() <= TOL and np.abs(rm - want_rm).max() <= TOL


@pytest.mark.parametrize("case", ["optimal_inversion", "optimal_codes_warm_valley", "classic_td_kh64", "multiple_td", "td_humidity"])
def test_shim_pre_interpolation(shim, ref, case):
    """pragab200::preInterpolation with the reference's signature on reference objects: points, settings and Kh series
    must come back exactly as the reference's own preInterpolation leaves them."""
    from test_oracle_port import clone, describe_diff, settings_equal
    from test_station_space import scenario
    dem, urban, st, s, var = scenario(case)
    raw = st.value.copy()
    st_r, st_g, s_r, s_g = st.copy(), st.copy(), clone(s), clone(s)
    ok, keep_r, series_r = ref.pre_interpolation_ex(st_r, s_r, var, current_value=raw, dem=dem if "td" in case else None)
    assert ok
    stp = st_g.as_pb()
    cv = A.pb_cv_points()
    cv.currentValue = A.fptr(raw)
    keep_g = np.zeros(st.n, dtype=np.uint8)
    series = A.pb_kh_series()
    d = dem.as_pb()
    err = C.create_string_buffer(512)
    rc = shim.shim_pre_interpolation(C.byref(stp), C.byref(s_g), var, C.byref(cv), C.byref(d) if "td" in case else None,
                                     A.u8ptr(keep_g), C.byref(series), err, 512)
    assert rc == A.PB_OK, err.value.decode()
    assert np.array_equal(keep_r, keep_g.astype(bool))
    assert np.array_equal(st_r.value[keep_r], st_g.value[keep_r])
    assert series_r == [(series.kh[i], series.error[i]) for i in range(series.n)]
    assert settings_equal(s_r, s_g), describe_diff(s_r, s_g)


def test_shim_spatial_quality_control(shim, ref):
    """pragab200::spatialQualityControl on reference meteo points: same flags and settings as the reference."""
    from test_oracle_port import clone, describe_diff, settings_equal
    from test_spatial_quality import sq_scenario
    for case in ("classic_codes", "optimal", "classic_td_kh64", "humidity_plain"):
        dem, urban, st, s, var, bad = sq_scenario(case)
        s_r, s_g = clone(s), clone(s)
        okr, wrong_r = ref.spatial_quality_control(st.copy(), s_r, var, dem=dem if "td" in case else None)
        stp = st.as_pb()
        cv = A.pb_cv_points()
        wrong_g = np.zeros(st.n, dtype=np.uint8)
        d = dem.as_pb()
        rc = shim.shim_spatial_quality_control(C.byref(stp), C.byref(s_g), var, C.byref(cv), C.byref(d) if "td" in case else None,
                                               A.u8ptr(wrong_g))
        assert rc == A.PB_OK and okr, case
        assert np.array_equal(wrong_r, wrong_g.astype(bool)), (case, np.flatnonzero(wrong_r), np.flatnonzero(wrong_g))
        assert settings_equal(s_r, s_g), case + ": " + describe_diff(s_r, s_g)


def test_shim_compute_residuals_local(shim, ref):
    dem = syn.make_dem(60, 70, cell_size=3000.0, seed=3)
    urban = syn.make_urban(60, 70, cell_size=3000.0)
    st = syn.make_stations(dem, 260, proxies=(dem, urban), seed=9, supplemental_frac=0.15)
    s = syn.settings_multiple_detrending()
    s.useLocalDetrending = 1
    s.minPointsLocalDetrending = 20
    s.useLapseRateCode = 1
    syn.set_points_range(s, st)
    use = np.ones(st.n, dtype=np.uint8)
    use[::11] = 0
    res_r, _ = ref.compute_residuals(st, s, A.airTemperature, st.value, use_for_cv=use, local=True)
    stp = st.as_pb()
    res_g = np.empty(st.n, dtype=np.float32)
    obs = st.value.copy()
    rc = shim.shim_compute_residuals_local(C.byref(stp), C.byref(s), A.airTemperature, A.fptr(obs), A.u8ptr(use), A.fptr(res_g))
    assert rc == A.PB_OK
    assert np.array_equal(res_r == np.float32(-9999), res_g == np.float32(-9999))
    assert np.abs(res_r.astype(np.float64) - res_g).max() <= TOL


def test_shim_glocal_detrending_raster_and_residuals(shim, ref):
    """pragab200::interpolationRasterGlocalDetrending / computeResidualsGlocalDetrending / preInterpolation(glocal) on
    reference objects (the macro areas live in Crit3DInterpolationSettings) against the reference's own loop."""
    from test_oracle_port import clone, describe_diff, settings_equal
    dem = syn.make_dem(100, 120, seed=5)
    urban = syn.make_urban(100, 120)
    st = syn.make_stations(dem, 400, proxies=(dem, urban), seed=17)
    s = syn.settings_multiple_detrending()
    s.useGlocalDetrending = 1
    syn.set_points_range(s, st)
    a_r, s_r = syn.make_macro_areas(dem, st, 2, 8), clone(s)
    ok_r, grid_r, mn_r, mx_r, _, _ = ref.glocal_detrending_raster(st.copy(), s_r, A.airTemperature, a_r, dem, (dem, urban), threads=1)
    ok2, res_r = ref.compute_residuals_glocal(st, st.value, st, clone(s_r), A.airTemperature, a_r, dem)
    a_g, s_g = syn.make_macro_areas(dem, st, 2, 8), clone(s)
    stp, d = st.as_pb(), dem.as_pb()
    g = (A.pb_raster * 2)(dem.as_pb(), urban.as_pb())
    grid_g = np.empty(dem.shape, dtype=np.float32)
    o = A.pb_raster_out()
    o.data = A.fptr(grid_g)
    res_g = np.empty(st.n, dtype=np.float32)
    err = C.create_string_buffer(512)
    rc = shim.shim_glocal(C.byref(stp), C.byref(s_g), A.airTemperature, a_g.arr, a_g.n, C.byref(d), g, 2, C.byref(o), A.fptr(res_g),
                          err, 512)
    assert (rc == A.PB_OK) == ok_r, err.value.decode()
    assert a_g.fits() == a_r.fits()
    assert settings_equal(s_r, s_g), describe_diff(s_r, s_g)
    nod = grid_r == np.float32(dem.flag)
    assert np.array_equal(nod, grid_g == np.float32(dem.flag))
    assert np.abs(grid_g[~nod].astype(np.float64) - grid_r[~nod]).max() <= TOL
    assert abs(o.minimum - mn_r) <= TOL and abs(o.maximum - mx_r) <= TOL
    assert ok2 and np.array_equal(res_r, res_g)


def test_shim_variogram_and_glocal_pre_interpolation(shim, ref):
    """pragab200::krigingEstimateVariogram (the signature the reference declares and never defines) and the glocal branch of
    pragab200::preInterpolation on reference objects."""
    import praga_b200 as pb
    from test_oracle_port import clone, describe_diff, settings_equal
    h = ((np.arange(40) + 0.5) * 3.0e3).astype(np.float32)
    r = h.astype(np.float64) / 8.0e4
    g = np.where(h < 8.0e4, 0.1 + 3.9 * (1.5 * r - 0.5 * r ** 3), 4.0).astype(np.float32)
    mode = C.c_int(A.KRIGING_SPHERICAL)
    sill, nug, rng, slope = C.c_double(0), C.c_double(0), C.c_double(0), C.c_double(0)
    rc = shim.shim_estimate_variogram(A.fptr(h), A.fptr(g), h.size, 0.0, C.byref(mode), C.byref(sill), C.byref(nug), C.byref(rng),
                                      C.byref(slope))
    want = pb.fit_variogram(h, g, None, A.KRIGING_SPHERICAL)
    assert rc == A.PB_OK and mode.value == A.KRIGING_SPHERICAL
    assert (sill.value, nug.value, rng.value) == (want["sill"], want["nugget"], want["range"])
    assert abs(rng.value - 8.0e4) / 8.0e4 < 5e-3

    dem = syn.make_dem(100, 120, seed=5)
    urban = syn.make_urban(100, 120)
    st = syn.make_stations(dem, 400, proxies=(dem, urban), seed=17)
    s = syn.settings_multiple_detrending()
    s.useGlocalDetrending = 1
    syn.set_points_range(s, st)
    a_r, s_r = syn.make_macro_areas(dem, st, 2, 8), clone(s)
    assert ref.glocal_detrending_fitting(st.copy(), s_r, A.airTemperature, a_r)
    a_g, s_g = syn.make_macro_areas(dem, st, 2, 8), clone(s)
    st_g = st.copy()                      # keep the buffers behind the descriptor alive
    stp = st_g.as_pb()
    err = C.create_string_buffer(256)
    rc = shim.shim_pre_interpolation_glocal(C.byref(stp), C.byref(s_g), A.airTemperature, a_g.arr, a_g.n, err, 256)
    assert rc == A.PB_OK, err.value.decode()
    assert a_g.fits() == a_r.fits()
    assert settings_equal(s_r, s_g), describe_diff(s_r, s_g)


class ShimSession:
    """a host session as PRAGA runs one: gis::Crit3DRasterGrid DEM / proxy / output objects that live across time steps"""

    def __init__(self, shim, s, dem, grids):
        self.shim, self.shape = shim, dem.shape
        d = dem.as_pb()
        g = (A.pb_raster * max(1, len(grids)))()
        for i, r in enumerate(grids):
            g[i] = r.as_pb()
        self.h = shim.shim_session_open(C.byref(s), C.byref(d), g, len(grids))

    def step(self, st, var, local=False):
        stp = st.as_pb()
        sec, mn, mx = C.c_double(0), C.c_float(0), C.c_float(0)
        rc = self.shim.shim_session_step(self.h, C.byref(stp), var, int(local), C.byref(sec), C.byref(mn), C.byref(mx))
        out = np.empty(self.shape, dtype=np.float32)
        o = A.pb_raster_out()
        o.data = A.fptr(out)
        self.shim.shim_session_output(self.h, C.byref(o))
        return rc == A.PB_OK, out, mn.value, mx.value, sec.value

    def close(self):
        self.shim.shim_session_close(self.h)


def test_shim_session_keeps_rasters_resident_and_notices_changes(shim, ref):
    """Time steps of one session: the DEM / proxy grids go up once (resident cache of the shim), every step must still equal
    the reference's grid; an announced in-place change of the DEM (pragab200::invalidate) is seen by the next step; with the
    cache switched off the grids are the same."""
    dem = syn.make_dem(420, 380, seed=31)
    urban = syn.make_urban(420, 380)
    s = syn.settings_multiple_detrending()
    s.currentSignificant[0] = s.currentSignificant[1] = 1
    s.nFit = s.nFitFunctions = 2
    s.fitFunction[0], s.fitNrParams[0] = A.piecewiseTwo, 4
    s.fitFunction[1], s.fitNrParams[1] = A.linear, 2
    for j, v in enumerate((400.0, 13.3, 0.0045, -0.0065)):
        s.fitParams[0][j] = v
    s.fitParams[1][0], s.fitParams[1][1] = 2.0, -0.6
    base = syn.make_stations(dem, 160, proxies=(dem, urban))
    sess = ShimSession(shim, s, dem, (dem, urban))
    try:
        for k in range(3):
            st = base.copy()
            st.value[:] = (base.value - 12.0 + np.random.default_rng(k).normal(0, 0.3, st.n)).astype(np.float32)
            okc, gc, mnc, mxc, _, _ = ref.interpolation_raster(st, s, A.airTemperature, dem, (dem, urban))
            okg, gg, mng, mxg, sec = sess.step(st, A.airTemperature)
            assert okg == okc
            assert np.array_equal(gg == np.float32(dem.flag), gc == np.float32(dem.flag))
            m = gc != np.float32(dem.flag)
            assert np.abs(gg[m].astype(np.float64) - gc[m]).max() <= TOL
            assert abs(mng - mnc) <= TOL and abs(mxg - mxc) <= TOL
        # the DEM changes in place: a valid cell becomes NODATA and an elevation changes by 800 m; the host announces it
        rows, cols = np.nonzero(dem.data != np.float32(dem.flag))
        r1, c1, r2, c2 = int(rows[1000]), int(cols[1000]), int(rows[5000]), int(cols[5000])
        dem2 = A.Raster(dem.data.copy(), dem.llx, dem.lly, dem.cell_size, dem.flag)
        dem2.data[r1, c1] = dem.flag
        dem2.data[r2, c2] += 800.0
        shim.shim_session_poke_dem(sess.h, r1, c1, float(dem.flag), 0)
        shim.shim_session_poke_dem(sess.h, r2, c2, float(dem2.data[r2, c2]), 1)
        okc, gc, *_ = ref.interpolation_raster(st, s, A.airTemperature, dem2, (dem2, urban))
        okg, gg, *_ = sess.step(st, A.airTemperature)
        assert gg[r1, c1] == np.float32(dem.flag)
        assert np.array_equal(gg == np.float32(dem.flag), gc == np.float32(dem.flag))
        m = gc != np.float32(dem.flag)
        assert np.abs(gg[m].astype(np.float64) - gc[m]).max() <= TOL
        # cache off: same grid
        shim.shim_session_cache(0)
        okg2, gg2, *_ = sess.step(st, A.airTemperature)
        shim.shim_session_cache(1)
        assert np.array_equal(gg, gg2)
    finally:
        sess.close()


def test_shim_compute_radiation_dem(shim, ref):
    """pragab200::computeRadiationDEM in place of radiation::computeRadiationDEM on the reference's own Crit3DRadiationSettings /
    Crit3DRadiationMaps objects: the five maps, their min / max."""
    from test_radiation import transmissivity_map
    dem = syn.make_dem(110, 120, seed=13, cell_size=100.0)
    dem.data[dem.data != np.float32(dem.flag)] *= np.float32(2.0)
    tr = transmissivity_map(dem)
    s = A.default_radiation_settings(2023, 2, 3, 9 * 3600 + 900, time_zone=1)
    want = ref.compute_radiation_dem(s, dem, tr)
    got = ref.compute_radiation_dem(s, dem, tr, shim=shim)
    assert want["ok"] and got["ok"]
    flag = np.float32(dem.flag)
    for nme in ("sunElevation", "beam", "diffuse", "reflected", "global"):
        assert np.array_equal(got[nme] == flag, want[nme] == flag)
        m = want[nme] != flag
        assert np.abs(got[nme][m].astype(np.float64) - want[nme][m]).max() <= TOL, nme
    assert np.abs(got["minmax"] - want["minmax"]).max() <= TOL


def test_shim_registers_the_points_topographic_distance_maps(shim, ref):
    """Crit3DInterpolationDataPoint::topographicDistance (what passDataToInterpolation copies from the meteo points) reaches
    the engine through the shim: the maps are uploaded through the raster cache and registered for the call."""
    from test_td_maps import humidity_case, make_maps
    from test_oracle_port import clone
    dem, st, s, var = humidity_case()
    maps = make_maps(ref, dem, st)
    idx = sorted(maps)
    plain = ref.interpolation_raster(st, clone(s), var, dem, ())[1]
    ref.set_topographic_distance_maps(idx, [maps[i] for i in idx])
    arr = (A.pb_raster * len(idx))()
    for k, i in enumerate(idx):
        arr[k] = maps[i].as_pb()
    shim.ref_set_topographic_distance_maps.restype = C.c_int
    shim.ref_set_topographic_distance_maps((C.c_int32 * len(idx))(*idx), arr, C.c_int(len(idx)))      # the harness' own point builder
    try:
        okc, gc, mnc, mxc, _, _ = ref.interpolation_raster(st, clone(s), var, dem, ())
        okg, gg, mng, mxg, err = shim_raster(shim, st, clone(s), var, dem, ())
    finally:
        ref.set_topographic_distance_maps()
        shim.ref_set_topographic_distance_maps(None, None, C.c_int(0))
    assert okg == okc, err
    assert np.array_equal(gg == np.float32(dem.flag), gc == np.float32(dem.flag))
    assert np.abs(gg.astype(np.float64) - gc).max() <= TOL
    assert np.abs(gc - plain).max() > 1e-2          # the maps mattered
    # and without maps the same call marches
    okg2, gg2, *_ = shim_raster(shim, st, clone(s), var, dem, ())
    assert okg2 and np.abs(gg2.astype(np.float64) - plain).max() <= TOL
