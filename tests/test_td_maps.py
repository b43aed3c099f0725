"""Precomputed per-point topographic-distance maps (Crit3DMeteoPoint::topographicDistance, loaded by
Project::loadTopographicDistanceMaps project.cpp:2368-2430, handed to the interpolation points by passDataToInterpolation
spatialControl.cpp:529): computeDistances (interpolation.cpp:232-247) reads the map at the target's cell where the target lies
inside it and only marches through the DEM where the map holds NODATA (or the point has no map).

The maps of a real project are what gis::topographicDistanceMap wrote for the same DEM, so at DEM cell centres they equal the
march. To prove the maps ARE read, the scenario perturbs them: scaled values, a NODATA block (fallback to the march), one map
on a sub-window of the DEM (targets outside it fall back), half of the points without a map. Whatever the maps hold, engine
and reference must agree: rasters (IDW and both Shepard estimators), explicit targets, the kh search, exact leave-one-out."""
import numpy as np
import pytest

from praga_b200 import _abi as A
from praga_b200 import synthetic as syn
from test_oracle_port import clone, describe_diff, settings_equal
from test_station_space import scenario as station_scenario


def make_maps(ref, dem, st):
    """{point index: Raster}: every second point has a map."""
    maps = {}
    rows, cols = dem.shape
    for i in range(0, st.n, 2):
        m = ref.topographic_distance_map(dem, float(st.x[i]), float(st.y[i]), float(st.z[i]))
        valid = m != np.float32(dem.flag)
        m = np.where(valid, m * np.float32(1.0 + 0.05 * (i % 7)), m).astype(np.float32)
        if i % 6 == 0:          # a block the map does not know: computeDistances falls back to the march there
            m[rows // 3: rows // 2, cols // 4: cols // 2] = np.float32(A.NODATA)
        r = A.Raster(m, dem.llx, dem.lly, dem.cell_size, dem.flag)
        if i % 10 == 0:         # a map on a sub-window of the DEM (another header): targets outside it are out of grid
            r0, r1, c0, c1 = rows // 5, rows - rows // 6, cols // 7, cols - cols // 5
            r = A.Raster(m[r0:r1, c0:c1], dem.llx + c0 * dem.cell_size, dem.lly + (rows - r1) * dem.cell_size, dem.cell_size, dem.flag)
        maps[i] = r
    return maps


class Registered:
    """the same maps registered with the reference driver and (resident) with the engine"""

    def __init__(self, engine, ref, maps):
        self.engine, self.ref = engine, ref
        idx = sorted(maps)
        ref.set_topographic_distance_maps(idx, [maps[i] for i in idx])
        self.ids = [engine.upload(maps[i]) for i in idx] if engine is not None else []
        if engine is not None:
            engine.set_topographic_distance_maps(idx, self.ids)

    def close(self):
        self.ref.set_topographic_distance_maps()
        if self.engine is not None:
            self.engine.set_topographic_distance_maps()
            for i in self.ids:
                self.engine.free(i)


def humidity_case(method=A.idw):
    dem = syn.make_dem(70, 90, seed=5, cell_size=300.0)
    st = syn.make_stations(dem, 60, variable="humidity", seed=13)
    s = A.default_settings()
    s.useTD = 1
    s.topoDist_Kh = 7
    s.method = method
    x, y = st.x.astype(np.float32), st.y.astype(np.float32)
    s.pointsBoundingBoxArea = float(np.float32(x.max() - x.min()) * np.float32(y.max() - y.min()))
    return dem, st, s, A.airRelHumidity


def test_reference_reads_the_maps(ref):
    """sanity of the scenario: with the perturbed maps the reference's raster changes."""
    dem, st, s, var = humidity_case()
    plain = ref.interpolation_raster(st, clone(s), var, dem, ())[1]
    reg = Registered(None, ref, make_maps(ref, dem, st))
    try:
        cached = ref.interpolation_raster(st, clone(s), var, dem, ())[1]
    finally:
        reg.close()
    again = ref.interpolation_raster(st, clone(s), var, dem, ())[1]
    assert np.abs(cached - plain).max() > 1e-2
    assert np.array_equal(again, plain)


@pytest.mark.gpu
@pytest.mark.parametrize("method", [A.idw, A.shepard, A.shepard_modified])
def test_gpu_raster_and_points_with_maps(engine, ref, method):
    dem, st, s, var = humidity_case(method)
    x, y, z = st.x.astype(np.float32), st.y.astype(np.float32), st.z.astype(np.float32)
    pv = st.proxy_values.astype(np.float64)
    dem_id = engine.upload(dem)
    plain = engine.interpolation_raster(st, clone(s), var, dem, ())[1]
    reg = Registered(engine, ref, make_maps(ref, dem, st))
    try:
        ok_r, grid_r, mn_r, mx_r, nv_r, _ = ref.interpolation_raster(st, clone(s), var, dem, ())
        ok_g, grid_g, mn_g, mx_g, nv_g = engine.interpolation_raster(st, clone(s), var, dem, ())
        assert ok_r == ok_g and nv_r == nv_g
        assert np.array_equal(grid_r == np.float32(dem.flag), grid_g == np.float32(dem.flag))
        assert np.abs(grid_r.astype(np.float64) - grid_g.astype(np.float64)).max() <= 1e-4
        # (the IDW kernel sums mean-shifted values: within the tolerance, not bit for bit; the Shepard kernels are bit-exact)
        assert float((grid_r == grid_g).mean()) >= (0.95 if method == A.idw else 0.999)
        assert np.abs(grid_g - plain).max() > 1e-2, "the maps were not read"
        # explicit targets = the stations themselves (not cell centres: the map value of the cell is what counts)
        pr = ref.interpolate_points_td(st, clone(s), var, x, y, z, pv, dem)
        pg = engine.interpolate_points(st, clone(s), var, x, y, pv, z=z, dem_id=dem_id)
        assert np.array_equal(pr == np.float32(A.NODATA), pg == np.float32(A.NODATA))
        assert np.abs(pr.astype(np.float64) - pg.astype(np.float64)).max() <= 1e-4
    finally:
        reg.close()
        engine.free(dem_id)
    # registry dropped: back to the march
    assert np.array_equal(engine.interpolation_raster(st, clone(s), var, dem, ())[1], plain)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["td_humidity", "classic_td_kh64", "optimal_td_kh64", "classic_td_kh64_shepard"])
def test_gpu_kh_search_and_exact_residuals_with_maps(engine, ref, case):
    """preInterpolation's kh search and the exact leave-one-out read the maps like every other computeDistances call."""
    dem, urban, st, s, var = station_scenario(case)
    raw = st.value.copy()
    dem_id = engine.upload(dem)
    reg = Registered(engine, ref, make_maps(ref, dem, st))
    try:
        st_r, st_g, s_r, s_g = st.copy(), st.copy(), clone(s), clone(s)
        ok, keep_r, series_r = ref.pre_interpolation_ex(st_r, s_r, var, current_value=raw, dem=dem)
        assert ok
        keep_g, series_g = engine.pre_interpolation_ex(st_g, s_g, var, current_value=raw, dem_id=dem_id)
        assert np.array_equal(keep_r, keep_g)
        assert np.array_equal(st_r.value[keep_r], st_g.value[keep_g])
        assert series_r == series_g, f"Kh series: {series_r[:3]} vs {series_g[:3]}"
        assert settings_equal(s_r, s_g), describe_diff(s_r, s_g)
        if s_r.topoDist_Kh == 0:
            s_r.topoDist_Kh = s_g.topoDist_Kh = 5      # the leave-one-out below must march / read the maps
        pts = st_r.subset(keep_r)
        pts.index = np.flatnonzero(keep_r).astype(np.int32)
        got = engine.compute_residuals_exact(st.copy(), pts, s_g, var, observed=raw, dem_id=dem_id)
        pv = st.proxy_values.astype(np.float64)
        est = ref.interpolate_points_td(pts, s_r, var, st.x.astype(np.float32), st.y.astype(np.float32), st.z.astype(np.float32),
                                        pv, dem)
        want = np.where(est == np.float32(-9999), np.float32(-9999), raw - est).astype(np.float32)
        assert np.array_equal(got, want), np.abs(got - want).max()
    finally:
        reg.close()
        engine.free(dem_id)
