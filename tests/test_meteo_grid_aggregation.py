"""DEM -> meteo-grid aggregation: Crit3DMeteoGrid::spatialAggregateMeteoGrid (meteo/meteoGrid.cpp:1135-1190)
+ spatialAggregateMeteoGridPoint (:1193-1240). CPU: the restatement agrees bit-exactly with the reference's own code driven
on a real Crit3DMeteoGrid; GPU: pb_spatial_aggregate_meteo_grid on a resident raster is bit-identical too."""
import numpy as np
import pytest

from praga_b200 import _abi as A
from praga_b200 import synthetic as syn
from test_oracle_port import port  # noqa: F401


def geometry(dem, cell_m=1500.0, step_m=None, seed=1):
    """Aggregation points of a regular meteo grid over the DEM: per cell a step_m lattice of points (like
    assignCellAggregationPoints' DEM-resolution scan, meteoGrid.cpp:667-760); some cells stick out of the DEM, one is empty."""
    rows, cols = dem.shape
    step_m = step_m or dem.cell_size
    x0, y0 = dem.llx - cell_m * 0.5, dem.lly - cell_m * 0.5
    nx, ny = int(cols * dem.cell_size / cell_m) + 2, int(rows * dem.cell_size / cell_m) + 2
    xs, ys, first, max_nr = [], [], [0], []
    rng = np.random.default_rng(seed)
    for j in range(ny):
        for i in range(nx):
            gx = x0 + i * cell_m + step_m / 2 + np.arange(0, cell_m, step_m)
            gy = y0 + j * cell_m + step_m / 2 + np.arange(0, cell_m, step_m)
            px, py = [a.ravel() for a in np.meshgrid(gx, gy, indexing="ij")]
            if (i + j) % 11 == 5:          # a cell whose points were thinned (excludeNoData), max count unchanged
                keep = rng.random(px.size) < 0.5
                px, py = px[keep], py[keep]
            if i == 1 and j == 1:
                px, py = px[:0], py[:0]    # no aggregation points at all
            if i == 2 and j == 1:
                px, py = px[:2], py[:2]    # fewer than MINIMUM_PERCENTILE_DATA values: median -> NODATA
            xs.append(px); ys.append(py)
            first.append(first[-1] + px.size)
            max_nr.append(gx.size * gy.size)
    return np.concatenate(xs), np.concatenate(ys), np.array(first, np.int64), np.array(max_nr, np.int32)


def field(dem):
    """a gridded map on the DEM header: smooth values, the DEM's NODATA mask"""
    r, c = np.indices(dem.shape)
    v = (12.0 + 6.0 * np.sin(r / 17.0) * np.cos(c / 23.0) + 0.01 * (r % 7)).astype(np.float32)
    v[dem.data == np.float32(dem.flag)] = dem.flag
    return A.Raster(v, dem.llx, dem.lly, dem.cell_size, dem.flag)


METHODS = [A.PB_AGGR_AVERAGE, A.PB_AGGR_MEDIAN, A.PB_AGGR_STDDEV, 4]      # 4 = aggrMin: not handled -> NODATA


@pytest.mark.parametrize("method", METHODS)
def test_port_matches_reference(ref, port, method):  # noqa: F811
    dem = syn.make_dem(90, 110, seed=5, nodata_frac=0.15)
    ras = field(dem)
    x, y, first, mx = geometry(dem)
    a = ref.spatial_aggregate_meteo_grid(ras, x, y, first, mx, method)
    b = port.spatial_aggregate_meteo_grid(ras, x, y, first, mx, method)
    assert np.array_equal(a, b)
    if method == 4:
        assert (a == np.float32(A.NODATA)).all()
    else:
        assert (a != np.float32(A.NODATA)).sum() > a.size // 2 and (a == np.float32(A.NODATA)).any()


@pytest.mark.gpu
@pytest.mark.parametrize("method", METHODS)
def test_gpu_matches_reference_bit_exact(engine, ref, method):
    dem = syn.make_dem(300, 340, seed=5, nodata_frac=0.1)
    ras = field(dem)
    x, y, first, mx = geometry(dem, cell_m=5000.0)           # 2500 samples per cell
    want = ref.spatial_aggregate_meteo_grid(ras, x, y, first, mx, method)
    rid = engine.upload(ras)
    aid = engine.aggregation_upload(x, y, first, mx)
    try:
        got = engine.spatial_aggregate_meteo_grid(rid, aid, method)
        assert np.array_equal(want, got), np.abs(want - got).max()
        again = engine.spatial_aggregate_meteo_grid(rid, aid, method)       # fresh points every call
        assert np.array_equal(got, again)
    finally:
        engine.aggregation_free(aid)
        engine.free(rid)
