"""Glocal detrending (glocalDetrendingFitting interpolation.cpp:2239-2294, macroAreaDetrending :1759-1829,
the loop of Project::interpolationDemGlocalDetrending project.cpp:3299-3375, computeResidualsGlocalDetrending
spatialControl.cpp:231-302).

CPU part: the restatement (oracle_port.cpp) must agree BIT-EXACTLY with the reference's own code (oracle/_ref).
GPU part: area fits bit-identical, blended raster within 1e-4 (NODATA mask exact), leave-one-out residuals bit-identical."""
import numpy as np
import pytest

from praga_b200 import _abi as A
from praga_b200 import synthetic as syn
from test_oracle_port import clone, describe_diff, port, settings_equal  # noqa: F401  (port is a fixture)


def scenario(case):
    rows, cols = (100, 120) if case != "big" else (300, 340)
    dem = syn.make_dem(rows, cols, seed=5)
    urban = syn.make_urban(rows, cols, nodata_frac=0.02 if case == "urban_nodata" else 0.0)
    n = 400 if case != "big" else 1500
    codes = 0.1 if "lapse_codes" in case else 0.0
    st = syn.make_stations(dem, n, proxies=(dem, urban), seed=17, supplemental_frac=codes)
    s = syn.settings_multiple_detrending()
    s.useGlocalDetrending = 1
    if codes:
        s.useLapseRateCode = 1
    if case == "three_piece":
        el = s.proxy[0]
        el.fittingFunction = A.piecewiseThree
        el.nFitParams = 5
        for j, (lo, hi, fg) in enumerate(zip((0, -30, 100, -0.01, -0.015), (1500, 50, 1000, 0.01, -0.002), (0, 1, 1, 1, 1))):
            el.fitMin[j], el.fitMax[j], el.fitFirstGuess[j] = lo, hi, fg
        s.proxy[0] = el
    if "shepard" in case:       # the demo's default estimator inside every macro area
        s.method = A.shepard_modified if "modified" in case else A.shepard
        s.pointsBoundingBoxArea = float((np.float32(st.x.max()) - np.float32(st.x.min())) * (np.float32(st.y.max()) - np.float32(st.y.min())))
    if case == "missing_proxies":
        st.proxy_values[::9, 1] = A.NODATA
        st.proxy_values[5::13, 0] = A.NODATA
    syn.set_points_range(s, st)
    side = 3 if case in ("three_by_three", "big") else 2

    def areas():
        a = syn.make_macro_areas(dem, st, side, 8)
        if case == "empty_area":        # an area with cells but no stations is never fitted; one without cells never gridded
            a = A.MacroAreas(a.meteo_points[:3] + [np.zeros(0, np.int32)], a.area_cells[:3] + [np.zeros(0, np.float32)])
        return a
    return dem, urban, st, s, areas


CASES = ["two_piece", "three_piece", "lapse_codes", "missing_proxies", "three_by_three", "urban_nodata", "empty_area",
         "modified_shepard", "shepard_lapse_codes"]


@pytest.mark.parametrize("case", CASES)
def test_port_matches_reference_bit_exact(ref, port, case):  # noqa: F811
    dem, urban, st, s, areas = scenario(case)
    got = []
    for o in (ref, port):
        a, so = areas(), clone(s)
        ok, grid, mn, mx, nv, _ = o.glocal_detrending_raster(st.copy(), so, A.airTemperature, a, dem, (dem, urban), threads=1)
        ok2, res = o.compute_residuals_glocal(st, st.value, st, clone(so), A.airTemperature, a, dem)
        assert ok2
        a2, sf = areas(), clone(s)
        assert o.glocal_detrending_fitting(st.copy(), sf, A.airTemperature, a2)
        got.append((ok, grid, mn, mx, nv, a.fits(), so, res, a2.fits(), sf))
    r, p = got
    assert r[0] == p[0] and r[4] == p[4]
    assert r[5] == p[5] and r[8] == p[8], "area fits differ"
    assert r[5] == r[8]
    assert np.array_equal(r[1], p[1]), np.abs(r[1] - p[1]).max()
    assert (r[2], r[3]) == (p[2], p[3])
    assert settings_equal(r[6], p[6]), describe_diff(r[6], p[6])
    assert settings_equal(r[9], p[9]), describe_diff(r[9], p[9])
    assert np.array_equal(r[7], p[7])
    if case not in ("empty_area", "urban_nodata"):
        assert r[0] and (r[1] != np.float32(dem.flag)).sum() == r[4]
    if case == "urban_nodata":      # incomplete significant proxies leave NODATA in the cell (project.cpp:3350-3354)
        assert r[0] and 0 < (r[1] != np.float32(dem.flag)).sum() < r[4]


def test_reference_parallel_order_is_within_rounding(ref):
    """The reference's OpenMP loop adds the areas of a cell in arbitrary order: same grid up to float rounding."""
    dem, urban, st, s, areas = scenario("two_piece")
    g = []
    for threads in (1, 0):
        ok, grid, *_ = ref.glocal_detrending_raster(st.copy(), clone(s), A.airTemperature, areas(), dem, (dem, urban), threads=threads)
        assert ok
        g.append(grid)
    assert np.abs(g[0] - g[1]).max() <= 4e-6


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES + ["big"])
def test_gpu_glocal_matches_reference(engine, ref, case):
    dem, urban, st, s, areas = scenario(case)
    a_r, s_r = areas(), clone(s)
    ok_r, grid_r, mn_r, mx_r, nv_r, _ = ref.glocal_detrending_raster(st.copy(), s_r, A.airTemperature, a_r, dem, (dem, urban), threads=1)
    dem_id, urb_id = engine.upload(dem), engine.upload(urban)
    try:
        a_g, s_g = areas(), clone(s)
        ok_g, grid_g, mn_g, mx_g, nv_g = engine.glocal_detrending_raster(st.copy(), s_g, A.airTemperature, a_g, dem_id, dem.shape,
                                                                         (dem_id, urb_id))
        assert ok_g == ok_r and nv_g == nv_r
        assert a_g.fits() == a_r.fits(), "area fits differ"
        assert settings_equal(s_r, s_g), describe_diff(s_r, s_g)
        if ok_r:
            nod_r = (grid_r == np.float32(dem.flag)) | (grid_r == np.float32(A.NODATA))
            nod_g = (grid_g == np.float32(dem.flag)) | (grid_g == np.float32(A.NODATA))
            assert np.array_equal(nod_r, nod_g), "NODATA masks differ"
            err = np.abs(grid_g.astype(np.float64) - grid_r.astype(np.float64))[~nod_r].max()
            assert err <= 1e-4, err          # north_star tolerance, in the variable's units
            assert abs(mn_g - mn_r) <= 1e-4 and abs(mx_g - mx_r) <= 1e-4
        # resident (cell, weight) pairs: same grid, bit for bit
        cid = engine.macro_areas_upload(a_g)
        try:
            a_c, s_c = areas(), clone(s)
            ok_c, grid_c, *_ = engine.glocal_detrending_raster(st.copy(), s_c, A.airTemperature, a_c, dem_id, dem.shape,
                                                               (dem_id, urb_id), cells_id=cid)
            assert ok_c == ok_g and np.array_equal(grid_c, grid_g) and a_c.fits() == a_g.fits()
        finally:
            engine.macro_areas_free(cid)
        # fitting alone (preInterpolation with useGlocalDetrending)
        a_f, s_f = areas(), clone(s)
        engine.glocal_detrending_fitting(st.copy(), s_f, A.airTemperature, a_f)
        a_fr, s_fr = areas(), clone(s)
        assert ref.glocal_detrending_fitting(st.copy(), s_fr, A.airTemperature, a_fr)
        assert a_f.fits() == a_fr.fits()
        assert settings_equal(s_fr, s_f), describe_diff(s_fr, s_f)
        # leave-one-out over the fitted areas: bit-identical
        ok2, res_r = ref.compute_residuals_glocal(st, st.value, st, clone(s_r), A.airTemperature, a_r, dem)
        res_g = engine.compute_residuals_glocal(st, st.value, st, clone(s_g), A.airTemperature, a_g)
        assert ok2 and np.array_equal(res_r, res_g), np.abs(res_r - res_g).max()
    finally:
        engine.free(dem_id)
        engine.free(urb_id)


def _edge_areas(dem, st, kind):
    a = syn.make_macro_areas(dem, st, 2, 8)
    if kind == "no_stations_anywhere":          # no area is ever fitted: only the height ranges / first guesses are set
        return A.MacroAreas([np.zeros(0, np.int32)] * a.n, [np.zeros(0, np.float32)] * a.n)
    if kind == "cells_without_stations":        # interpolate() over no points gives NODATA: the reference returns false
        return A.MacroAreas(a.meteo_points[:3] + [np.zeros(0, np.int32)], a.area_cells)
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["no_stations_anywhere", "cells_without_stations"])
def test_port_matches_reference_on_degenerate_areas(ref, port, kind):  # noqa: F811
    dem, urban, st, s, _ = scenario("two_piece")
    got = []
    for o in (ref, port):
        a, so = _edge_areas(dem, st, kind), clone(s)
        ok, grid, mn, mx, nv, _ = o.glocal_detrending_raster(st.copy(), so, A.airTemperature, a, dem, (dem, urban), threads=1)
        got.append((ok, a.fits(), so, grid))
    assert got[0][0] == got[1][0] and got[0][1] == got[1][1]
    assert settings_equal(got[0][2], got[1][2]), describe_diff(got[0][2], got[1][2])
    if kind == "no_stations_anywhere":
        assert got[0][0] and np.array_equal(got[0][3], got[1][3]) and (got[0][3] == np.float32(dem.flag)).all()
    else:
        assert not got[0][0]


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["no_stations_anywhere", "cells_without_stations"])
def test_gpu_degenerate_areas(engine, ref, kind):
    dem, urban, st, s, _ = scenario("two_piece")
    a_r, s_r = _edge_areas(dem, st, kind), clone(s)
    ok_r, grid_r, *_ = ref.glocal_detrending_raster(st.copy(), s_r, A.airTemperature, a_r, dem, (dem, urban), threads=1)
    dem_id, urb_id = engine.upload(dem), engine.upload(urban)
    try:
        a_g, s_g = _edge_areas(dem, st, kind), clone(s)
        ok_g, grid_g, *_ = engine.glocal_detrending_raster(st.copy(), s_g, A.airTemperature, a_g, dem_id, dem.shape, (dem_id, urb_id))
        assert ok_g == ok_r
        assert a_g.fits() == a_r.fits()
        assert settings_equal(s_r, s_g), describe_diff(s_r, s_g)
        if ok_r:
            assert np.array_equal(grid_r, grid_g)
    finally:
        engine.free(dem_id)
        engine.free(urb_id)


@pytest.mark.gpu
def test_gpu_glocal_rejects_what_it_does_not_cover(engine):
    import praga_b200 as pb
    dem, urban, st, s, areas = scenario("two_piece")
    dem_id = engine.upload(dem)
    try:
        s_td = clone(s)
        s_td.useTD = 1          # topographic distance inside macro areas needs the meteo points of the kh search: never silently ignored
        with pytest.raises(pb.PragaB200Error) as e:
            engine.glocal_detrending_raster(st.copy(), s_td, A.airTemperature, areas(), dem_id, dem.shape, (dem_id, dem_id))
        assert e.value.status == A.PB_ERR_INVALID
        s_off = clone(s)
        s_off.useGlocalDetrending = 0
        with pytest.raises(pb.PragaB200Error) as e:
            engine.glocal_detrending_raster(st.copy(), s_off, A.airTemperature, areas(), dem_id, dem.shape, (dem_id, dem_id))
        assert e.value.status == A.PB_ERR_INVALID
    finally:
        engine.free(dem_id)


# ---- topographic distance inside macro areas: macroAreaDetrending ends with the area's own kh search over the area's meteo
# points (interpolation.cpp:1811-1826), the area's cells are gridded with distance + kh * topographicDistance
TD_CASES = ["two_piece", "lapse_codes", "modified_shepard", "shepard_lapse_codes"]


def td_scenario(case):
    dem, urban, st, s, areas = scenario(case)
    s.useTD = 1
    s.topoDist_maxKh = 24
    return dem, urban, st, s, areas


def test_reference_glocal_td_differs_from_plain(ref):
    """sanity: with useTD the reference's grid changes (some area ends on kh != 0)."""
    dem, urban, st, s, areas = td_scenario("two_piece")
    ok, grid_td, *_ = ref.glocal_detrending_raster(st.copy(), clone(s), A.airTemperature, areas(), dem, (dem, urban), threads=1,
                                                   meteo=st, observed=st.value)
    s.useTD = 0
    ok0, grid_0, *_ = ref.glocal_detrending_raster(st.copy(), clone(s), A.airTemperature, areas(), dem, (dem, urban), threads=1)
    assert ok and ok0
    assert np.abs(grid_td - grid_0).max() > 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("case", TD_CASES)
def test_gpu_glocal_topographic_distance(engine, ref, case):
    import praga_b200 as pb
    dem, urban, st, s, areas = td_scenario(case)
    a_r, s_r = areas(), clone(s)
    ok_r, grid_r, mn_r, mx_r, nv_r, _ = ref.glocal_detrending_raster(st.copy(), s_r, A.airTemperature, a_r, dem, (dem, urban),
                                                                     threads=1, meteo=st, observed=st.value)
    dem_id, urb_id = engine.upload(dem), engine.upload(urban)
    try:
        a_g, s_g = areas(), clone(s)
        kh = np.full(a_g.n, -1, dtype=np.int32)
        ok_g, grid_g, mn_g, mx_g, nv_g = engine.glocal_detrending_raster(st.copy(), s_g, A.airTemperature, a_g, dem_id, dem.shape,
                                                                         (dem_id, urb_id), meteo=st, observed=st.value, area_kh=kh)
        print("kh per area:", kh.tolist())
        assert ok_g == ok_r and nv_g == nv_r
        assert a_g.fits() == a_r.fits(), "area fits differ"
        assert settings_equal(s_r, s_g), describe_diff(s_r, s_g)
        assert (kh > 0).any(), "no area searched its way to kh != 0: the case does not exercise the ray march"
        nod_r = (grid_r == np.float32(dem.flag)) | (grid_r == np.float32(A.NODATA))
        nod_g = (grid_g == np.float32(dem.flag)) | (grid_g == np.float32(A.NODATA))
        assert np.array_equal(nod_r, nod_g), "NODATA masks differ"
        err = np.abs(grid_g.astype(np.float64) - grid_r.astype(np.float64))[~nod_r].max()
        assert err <= 1e-4, err
        assert abs(mn_g - mn_r) <= 1e-4 and abs(mx_g - mx_r) <= 1e-4
        # resident cells: same bits
        cid = engine.macro_areas_upload(a_g)
        try:
            a_c = areas()
            ok_c, grid_c, *_ = engine.glocal_detrending_raster(st.copy(), clone(s), A.airTemperature, a_c, dem_id, dem.shape,
                                                               (dem_id, urb_id), cells_id=cid, meteo=st, observed=st.value)
            assert ok_c == ok_g and np.array_equal(grid_c, grid_g)
        finally:
            engine.macro_areas_free(cid)
        # leave-one-out: macroAreaDetrending (kh search included) then interpolate() with the area's kh
        ok2, res_r = ref.compute_residuals_glocal(st, st.value, st, clone(s_r), A.airTemperature, a_r, dem)
        res_g = engine.compute_residuals_glocal(st, st.value, st, clone(s_g), A.airTemperature, a_g, dem_id=dem_id)
        assert ok2 and np.array_equal(res_r, res_g), np.abs(res_r - res_g).max()
        # without the meteo points the call refuses instead of skipping the search
        with pytest.raises(pb.PragaB200Error) as e:
            engine.glocal_detrending_raster(st.copy(), clone(s), A.airTemperature, areas(), dem_id, dem.shape, (dem_id, urb_id))
        assert e.value.status == A.PB_ERR_INVALID
    finally:
        engine.free(dem_id)
        engine.free(urb_id)
