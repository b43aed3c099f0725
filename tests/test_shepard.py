"""Shepard estimators of interpolate() (interpolation.cpp:806-1028): shepardIdw and modifiedShepardIdw (with the reference's
swapped x/y inside the direction cosine). CPU: the restatement against the reference, bit-exact. GPU: rasters and
leave-one-out points against the reference; the tolerance is the north star's 1e-4 with bit-exact NODATA masks, and
the test also requires >= 99.9 % of the cells to be bit-identical (same float/double mix, ties aside)."""
import numpy as np
import pytest

from praga_b200 import _abi as A
from praga_b200 import synthetic as syn
from test_oracle_port import clone, port  # noqa: F401


def bbox_area(st):
    x, y = st.x.astype(np.float32), st.y.astype(np.float32)
    return float(np.float32(x.max() - x.min()) * np.float32(y.max() - y.min()))


def scenario(case, rows=70, cols=90, n=160):
    dem = syn.make_dem(rows, cols, seed=5, cell_size=300.0)
    urban = syn.make_urban(rows, cols, cell_size=300.0)
    method = A.shepard_modified if "modified" in case else A.shepard
    if "humidity" in case:
        st = syn.make_stations(dem, n, variable="humidity", seed=13)
        s = A.default_settings()
        var, grids = A.airRelHumidity, ()
    elif "sparse" in case:      # < 5 points inside the initial radius almost everywhere: the 5-nearest branch
        st = syn.make_stations(dem, 9, variable="humidity", seed=2)
        s = A.default_settings()
        var, grids = A.airRelHumidity, ()
    else:
        st = syn.make_stations(dem, n, proxies=(dem, urban), seed=17, supplemental_frac=0.1 if "codes" in case else 0.0)
        s = syn.settings_multiple_detrending()
        if "codes" in case:
            s.useLapseRateCode = 1
        syn.set_points_range(s, st)
        var, grids = A.airTemperature, (dem, urban)
    s.method = method
    s.pointsBoundingBoxArea = bbox_area(st)
    if "_td" in case:           # computeDistances adds kh * topographicDistance in front of the estimator (interpolation.cpp:226-251)
        s.useTD = 1
        s.topoDist_Kh = 9 if "humidity" in case else 4
    if "clustered" in case:     # > 10 points inside the initial radius around the cluster: the 10-nearest branch
        rng = np.random.default_rng(4)
        k = st.n // 2
        st.x[:k] = dem.llx + 9000 + rng.random(k) * 2500
        st.y[:k] = dem.lly + 8000 + rng.random(k) * 2500
        s.pointsBoundingBoxArea = bbox_area(st)
    return dem, urban, st, s, var, grids


CASES = ["shepard_humidity", "modified_humidity", "shepard_sparse", "modified_sparse", "shepard_clustered_humidity",
         "modified_clustered_humidity", "shepard_detrended", "modified_detrended_codes",
         "shepard_humidity_td", "modified_humidity_td", "modified_sparse_humidity_td", "modified_detrended_codes_td"]


def prepared(ref, case):
    dem, urban, st, s, var, grids = scenario(case)
    if grids:
        ok, keep = ref.pre_interpolation(st, s, var, grids)
        assert ok
        st = st.subset(keep)
    return dem, st, s, var, grids


@pytest.mark.parametrize("case", CASES)
def test_port_reproduces_reference_shepard(ref, port, case):  # noqa: F811
    dem, st, s, var, grids = prepared(ref, case)
    a = ref.interpolation_raster(st, clone(s), var, dem, grids)
    b = port.interpolation_raster(st, clone(s), var, dem, grids)
    assert a[0] == b[0] and a[4] == b[4]
    assert np.array_equal(a[1], b[1]), f"max diff {np.abs(a[1] - b[1]).max()}"
    x, y, z = st.x.astype(np.float32), st.y.astype(np.float32), st.z.astype(np.float32)
    pv = st.proxy_values.astype(np.float64)
    if "_td" in case:
        pa = ref.interpolate_points_td(st, clone(s), var, x, y, z, pv, dem)
        pb_ = port.interpolate_points_td(st, clone(s), var, x, y, z, pv, dem)
    else:
        pa = ref.interpolate_points(st, clone(s), var, x, y, z, pv)
        pb_ = port.interpolate_points(st, clone(s), var, x, y, z, pv)
    assert np.array_equal(pa, pb_)


def test_shepard_branches_are_exercised(ref):
    """the three neighbourhood branches must all occur in the scenarios (count inside the initial radius <5, 5..10, >10)."""
    seen = set()
    for case in ("shepard_sparse", "shepard_humidity", "shepard_clustered_humidity"):
        dem, urban, st, s, var, grids = scenario(case)
        R0 = np.float32(np.sqrt((8 * np.float32(s.pointsBoundingBoxArea)) / (np.float32(3.1415926535898) * np.float32(st.n))))
        x, y = st.x.astype(np.float32), st.y.astype(np.float32)
        rows, cols = dem.shape
        for r in range(0, rows, 7):
            for c in range(0, cols, 7):
                cx = np.float32(dem.llx + dem.cell_size * c + 0.5)
                cy = np.float32(dem.lly + dem.cell_size * (rows - r) - 0.5)
                d = np.sqrt((x - cx) ** 2 + (y - cy) ** 2)
                k = int(((d <= R0) & (d > 0)).sum())
                seen.add("few" if k < 5 else ("many" if k > 10 else "mid"))
    assert seen == {"few", "mid", "many"}, seen


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_shepard_raster_and_points(engine, ref, case):
    dem, st, s, var, grids = prepared(ref, case)
    ok_r, grid_r, mn_r, mx_r, nv_r, _ = ref.interpolation_raster(st, clone(s), var, dem, grids)
    ok_g, grid_g, mn_g, mx_g, nv_g = engine.interpolation_raster(st, clone(s), var, dem, grids)
    assert ok_r == ok_g and nv_r == nv_g
    flag = np.float32(dem.flag)
    assert np.array_equal(grid_r == flag, grid_g == flag), "NODATA masks differ"
    assert np.array_equal(grid_r == np.float32(A.NODATA), grid_g == np.float32(A.NODATA))
    diff = np.abs(grid_r.astype(np.float64) - grid_g.astype(np.float64))
    assert diff.max() <= 1e-4, diff.max()
    same = float((grid_r == grid_g).mean())
    assert same >= 0.999, same
    assert abs(mn_r - mn_g) <= 1e-4 and abs(mx_r - mx_g) <= 1e-4
    # leave-one-out at the stations (computeResiduals' inner call)
    x, y, z = st.x.astype(np.float32), st.y.astype(np.float32), st.z.astype(np.float32)
    pv = st.proxy_values.astype(np.float64)
    if "_td" in case:
        pr = ref.interpolate_points_td(st, clone(s), var, x, y, z, pv, dem)
        dem_id = engine.upload(dem)
        try:
            pg = engine.interpolate_points(st, clone(s), var, x, y, pv, z=z, dem_id=dem_id)
        finally:
            engine.free(dem_id)
    else:
        pr = ref.interpolate_points(st, clone(s), var, x, y, z, pv)
        pg = engine.interpolate_points(st, clone(s), var, x, y, pv)
    assert np.array_equal(pr == np.float32(A.NODATA), pg == np.float32(A.NODATA))
    assert np.abs(pr.astype(np.float64) - pg.astype(np.float64)).max() <= 1e-4
    if "_td" in case:           # (computeResiduals with topographic distance: test_spatial_quality / test_station_space)
        return
    res_r, stats_r = ref.compute_residuals(st, clone(s), var, st.value)
    res_g, stats_g = engine.compute_residuals(st, clone(s), var, st.value)
    assert np.abs(res_r.astype(np.float64) - res_g.astype(np.float64)).max() <= 1e-4
    assert stats_r["n"] == stats_g["n"] and abs(stats_r["meanAbsoluteError"] - stats_g["meanAbsoluteError"]) <= 1e-4
