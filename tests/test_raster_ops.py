"""Raster map operators on either side of the gridding path: computeRelativeHumidityMap,
fixDailyThermalConsistency, gis::topographicDistanceMap, gis::resampleGrid. CPU: the restatement is pinned against the
reference's own code; GPU: bit-exact for the index / selection work (thermal consistency, topographic distance, centre and
median resampling), <= 1e-4 for the floating-point maps (relative humidity through pow/exp, cell averages)."""
import numpy as np
import pytest

from praga_b200 import _abi as A
from praga_b200 import synthetic as syn
from test_oracle_port import port  # noqa: F401  (fixture)

TOL = 1e-4


def temperature_maps(rows=120, cols=150, seed=4):
    dem = syn.make_dem(rows, cols, seed=seed)
    rng = np.random.default_rng(seed)
    flag = np.float32(dem.flag)
    t = (15.0 - 0.0065 * np.where(dem.data == flag, 0, dem.data) + rng.normal(0, 1.0, dem.shape)).astype(np.float32)
    td = (t - np.abs(rng.normal(3.0, 2.5, dem.shape))).astype(np.float32)
    td[rng.random(dem.shape) < 0.03] += np.float32(6.0)                     # dew point above temperature: RH clips at 100
    t[dem.data == flag] = flag
    td[dem.data == flag] = flag
    td[rng.random(dem.shape) < 0.02] = flag                                  # holes that differ between the two maps
    return dem, A.Raster(t, dem.llx, dem.lly, dem.cell_size, dem.flag), A.Raster(td, dem.llx, dem.lly, dem.cell_size, dem.flag)


def new_header(src, factor, rows=None, cols=None, shift=(0.0, 0.0)):
    h = A.pb_raster_header()
    cs = src.cell_size * factor
    h.cellSize, h.invCellSize = cs, 1.0 / cs
    h.nrRows = rows if rows is not None else int(src.shape[0] * src.cell_size / cs)
    h.nrCols = cols if cols is not None else int(src.shape[1] * src.cell_size / cs)
    h.llx, h.lly = src.llx + shift[0], src.lly + shift[1]
    h.flag = src.flag
    return h


RESAMPLE_CASES = [("average", A.PB_AGGR_AVERAGE, 4.0, 0.0), ("average_thr", A.PB_AGGR_AVERAGE, 2.5, 0.5),
                  ("median", A.PB_AGGR_MEDIAN, 3.0, 0.0), ("center", A.PB_AGGR_CENTER, 5.0, 0.0),
                  ("finer", A.PB_AGGR_AVERAGE, 0.5, 0.0), ("prevailing", A.PB_AGGR_PREVAILING, 4.0, 0.0),
                  ("prevailing_sparse", A.PB_AGGR_PREVAILING, 3.0, 0.3)]


def categorical(r):
    """a land-use-like raster (a few integer classes) on the header of r: what aggrPrevailing is for"""
    d = r.data.copy()
    ok = d != np.float32(r.flag)
    d[ok] = np.floor(d[ok] / 150.0) % 7
    return A.Raster(d, r.llx, r.lly, r.cell_size, r.flag)


# ---------------------------------------------------------------- CPU: restatement == reference
def test_port_relative_humidity_and_thermal_consistency(ref, port):  # noqa: F811
    dem, t, td = temperature_maps()
    a, b = ref.relative_humidity_map(t, td), port.relative_humidity_map(t, td)
    assert a[0] and b[0]
    assert np.array_equal(a[1], b[1]) and a[2:] == b[2:]
    assert a[1][dem.data != np.float32(dem.flag)].max() == 100.0 and a[1][a[1] != np.float32(dem.flag)].min() >= 1.0
    fa, na = ref.fix_daily_thermal_consistency(t, td)          # (tmax, tmin) = (t, td): td > t where the dew point was raised
    fb, nb = port.fix_daily_thermal_consistency(t, td)
    assert na == nb and na > 0 and np.array_equal(fa, fb)


def test_port_topographic_distance_map(ref, port):  # noqa: F811
    dem = syn.make_dem(70, 90, seed=6, cell_size=250.0)
    x, y, z = dem.llx + 31 * 250.0 + 40.0, dem.lly + 22 * 250.0 + 3.0, 812.0
    assert np.array_equal(ref.topographic_distance_map(dem, x, y, z), port.topographic_distance_map(dem, x, y, z))


@pytest.mark.parametrize("name,method,factor,thr", RESAMPLE_CASES)
def test_port_resample_grid(ref, port, name, method, factor, thr):  # noqa: F811
    src = syn.make_dem(96, 128, seed=8, nodata_frac=0.3 if name == "prevailing_sparse" else 0.05)
    if method == A.PB_AGGR_PREVAILING:
        src = categorical(src)
    h = new_header(src, factor, shift=(30.0, -20.0))
    a, b = ref.resample_grid(src, h, method, thr), port.resample_grid(src, h, method, thr)
    assert np.array_equal(a[1], b[1]), np.abs(a[1] - b[1]).max()
    assert a[2:] == b[2:]


# ---------------------------------------------------------------- GPU parity
@pytest.mark.gpu
def test_gpu_relative_humidity_map(engine, ref):
    dem, t, td = temperature_maps(400, 520)
    t_id, td_id = engine.upload(t), engine.upload(td)
    try:
        okr, gr, mnr, mxr = ref.relative_humidity_map(t, td)
        okg, gg, mng, mxg, rid = engine.relative_humidity_map(t_id, td_id, t.shape, keep=True)
        assert okg == okr
        flag = np.float32(t.flag)
        assert np.array_equal(gg == flag, gr == flag)
        assert np.abs(gg.astype(np.float64) - gr).max() <= TOL
        assert abs(mng - mnr) <= TOL and abs(mxg - mxr) <= TOL
        assert np.array_equal(engine.download(rid, t.shape), gg)          # the resident copy is the same map
        engine.free(rid)
    finally:
        engine.free(t_id)
        engine.free(td_id)


@pytest.mark.gpu
def test_gpu_thermal_consistency(engine, ref):
    dem, tmax, tmin = temperature_maps(300, 310, seed=9)
    want, n_ref = ref.fix_daily_thermal_consistency(tmax, tmin)
    a, b = engine.upload(tmax), engine.upload(tmin)
    try:
        n = engine.fix_daily_thermal_consistency(a, b)
        got = engine.download(b, tmin.shape)
    finally:
        engine.free(a)
        engine.free(b)
    assert n == n_ref and n > 0
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_gpu_topographic_distance_map_bit_exact(engine, ref):
    dem = syn.make_dem(260, 300, seed=6, cell_size=250.0)
    did = engine.upload(dem)
    try:
        for (cx, cy, z) in ((31.2, 22.1, 812.0), (250.7, 199.4, 90.0), (-3.0, 5.0, 1500.0)):      # last one: outside the grid
            x, y = dem.llx + cx * 250.0, dem.lly + cy * 250.0
            want = ref.topographic_distance_map(dem, x, y, z)
            ok, got, _, _, _ = engine.topographic_distance_map(did, dem.shape, x, y, z)
            assert np.array_equal(got, want), np.abs(got - want).max()
    finally:
        engine.free(did)


@pytest.mark.gpu
@pytest.mark.parametrize("name,method,factor,thr", RESAMPLE_CASES)
def test_gpu_resample_grid(engine, ref, name, method, factor, thr):
    src = syn.make_dem(480, 640, seed=8, nodata_frac=0.3 if name == "prevailing_sparse" else 0.05)
    urban = syn.make_urban(480, 640)
    if method == A.PB_AGGR_PREVAILING:
        src = categorical(src)
    for grid in (src, urban):
        h = new_header(grid, factor, shift=(30.0, -20.0))
        sid = engine.upload(grid)
        try:
            okr, gr, mnr, mxr = ref.resample_grid(grid, h, method, thr)
            okg, gg, mng, mxg, _ = engine.resample_grid(sid, h, method, thr)
        finally:
            engine.free(sid)
        flag = np.float32(h.flag)
        assert np.array_equal(gg == flag, gr == flag), name
        if method == A.PB_AGGR_AVERAGE and factor > 1:
            assert np.abs(gg.astype(np.float64) - gr).max() <= TOL
        else:
            assert np.array_equal(gg, gr), np.abs(gg - gr).max()
        if (gr != flag).any():
            assert abs(mng - mnr) <= TOL and abs(mxg - mxr) <= TOL


@pytest.mark.gpu
def test_gpu_resampled_proxy_feeds_the_raster_call(engine, ref):
    """the use the reference makes of resampleGrid on the path (pragaProject.cpp:2930-2942): a proxy grid with a foreign
    geometry is resampled onto the DEM header and the resampled grid becomes the proxy of interpolationRaster."""
    dem = syn.make_dem(200, 230)
    coarse = syn.make_urban(80, 92, cell_size=250.0)
    h = new_header(dem, 1.0)
    cid = engine.upload(coarse)
    did = engine.upload(dem)
    try:
        ok, res_g, _, _, rid = engine.resample_grid(cid, h, A.PB_AGGR_AVERAGE, 0.0, keep=True)
        okr, res_r, _, _ = ref.resample_grid(coarse, h, A.PB_AGGR_AVERAGE, 0.0)
        assert np.array_equal(res_g, res_r)                      # factor < 1: nearest-cell lookups, exact
        resampled = A.Raster(res_r, dem.llx, dem.lly, dem.cell_size, dem.flag)
        st = syn.make_stations(dem, 120, proxies=(dem, resampled))
        s = syn.settings_multiple_detrending()
        syn.set_points_range(s, st)
        okp, keep = ref.pre_interpolation(st, s, A.airTemperature, (dem, resampled))
        st = st.subset(keep)
        okc, gc, mnc, mxc, nvc, _ = ref.interpolation_raster(st, s, A.airTemperature, dem, (dem, resampled))
        out = np.empty(dem.shape, dtype=np.float32)
        okg, gg, mng, mxg, nvg = engine.interpolation_raster_resident(st, s, A.airTemperature, did, (did, rid), out=out)
        assert okg == okc and nvg == nvc
        m = gc != np.float32(dem.flag)
        assert np.array_equal(gg == np.float32(dem.flag), ~m)
        assert np.abs(gg[m].astype(np.float64) - gc[m]).max() <= TOL
        engine.free(rid)
    finally:
        engine.free(cid)
        engine.free(did)


@pytest.mark.gpu
def test_gpu_raster_sample_points(engine, ref):
    """pb_raster_sample_points = the lookup of Project::passInterpolatedTemperatureToHumidityPoints (project.cpp:2826-2848),
    then the dew point of those stations (tDewFromRelHum) feeds the humidity-from-dew-point chain."""
    import ctypes as C
    from oracle import binding
    lib = binding.load_ref_sample()
    dem, t, td = temperature_maps()
    rng = np.random.default_rng(4)
    n = 5000
    x = dem.llx + (rng.random(n) * 1.2 - 0.1) * dem.shape[1] * dem.cell_size          # some points fall outside the grid
    y = dem.lly + (rng.random(n) * 1.2 - 0.1) * dem.shape[0] * dem.cell_size
    x[:3] = [dem.llx, dem.llx + dem.shape[1] * dem.cell_size, dem.llx - 1e-9]          # the edges
    y[:3] = [dem.lly, dem.lly + dem.shape[0] * dem.cell_size, dem.lly]
    want = np.empty(n, dtype=np.float32)
    win = np.zeros(n, dtype=np.uint8)
    r = t.as_pb()
    lib.ref_raster_sample_points(C.byref(r), A.dptr(x), A.dptr(y), n, A.fptr(want), A.u8ptr(win))
    tid = engine.upload(t)
    try:
        got, gin = engine.raster_sample_points(tid, x, y)
    finally:
        engine.free(tid)
    assert np.array_equal(gin, win.astype(bool)) and np.array_equal(got, want)
    assert 0 < gin.sum() < n and (got[gin] == np.float32(t.flag)).any()


# ---- topographicIndex (interpolationCmd.cpp:397-482): the orography-index proxy raster
TI_CASES = [("one_width", [1500.0]), ("two_widths", [1500.0, 3100.0]), ("narrow", [520.0]), ("one_cell", [260.0]),
            ("zero_cells", [100.0])]


@pytest.mark.parametrize("name,widths", TI_CASES)
def test_port_topographic_index(ref, port, name, widths):  # noqa: F811
    dem = syn.make_dem(70, 90, seed=6, cell_size=250.0, nodata_frac=0.08)
    a, b = ref.topographic_index(dem, widths), port.topographic_index(dem, widths)
    assert a[0] and b[0] and np.array_equal(a[1], b[1]) and a[2:] == b[2:]
    valid = a[1] != np.float32(dem.flag)
    if name in ("zero_cells", "one_cell"):   # 0 cells, or 1 cell (only the diagonal neighbours, sqrt(2) > 1 away): nothing is written
        assert not valid.any()
    else:
        assert valid.any() and np.abs(a[1][valid]).max() <= len(widths) + 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name,widths", TI_CASES + [("wide", [9000.0])])
def test_gpu_topographic_index(engine, ref, name, widths):
    dem = syn.make_dem(140, 180, seed=6, cell_size=250.0, nodata_frac=0.08)
    ok, want, mn, mx = ref.topographic_index(dem, widths)
    did = engine.upload(dem)
    try:
        okg, got, mng, mxg, rid = engine.topographic_index(did, dem.shape, widths, keep=True)
        assert np.array_equal(got, want), np.abs(got - want).max()
        if (want != np.float32(dem.flag)).any():
            assert (mng, mxg) == (mn, mx)
        back = engine.download(rid, dem.shape)          # the resident copy can serve as a proxy grid
        assert np.array_equal(back[1] if isinstance(back, tuple) else back, want)
        engine.free(rid)
    finally:
        engine.free(did)
