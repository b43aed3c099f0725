"""The hourly driver kept on the device (pb_meteo_grid_step_batch): PragaProject::interpolationMeteoGrid with
getMeteoGridUpscaleFromDem() (src/pragaProject/pragaProject.cpp:2951-2990) — per hour: air temperature -> DEM map -> meteo-grid
cells; relative humidity through the dew point (passInterpolatedTemperatureToHumidityPoints, project.cpp:2826-2848 ->
tDewFromRelHum at the stations -> dew-point map -> computeRelativeHumidityMap, meteoMaps.cpp:301-321 -> cells); precipitation ->
cells. Checked against the same chain out of the reference's own functions (oracle/_ref): cell values within 1e-4 (NODATA cells
equal), rasters within 1e-4 with equal masks, quality flags and leave-one-out residuals of the dew-point gridding equal."""
import ctypes as C

import numpy as np
import pytest

from praga_b200 import _abi as A
from praga_b200 import synthetic as syn
from test_batch_gridding import settings_for
from test_meteo_grid_aggregation import geometry
from test_oracle_port import clone

TOL = 1e-4
pytestmark = pytest.mark.gpu
HOURS = 3


def ref_sample(ref, raster, x, y):
    from oracle import binding
    lib = binding.load_ref_sample()
    r = raster.as_pb()
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    val = np.empty(x.size, dtype=np.float32)
    ing = np.empty(x.size, dtype=np.uint8)
    lib.ref_raster_sample_points(C.byref(r), A.dptr(x), A.dptr(y), x.size, A.fptr(val), A.u8ptr(ing))
    return val, ing.astype(bool)


def inputs(n=240, seed=41):
    dem = syn.make_dem(160, 180, seed=6, cell_size=500.0)
    urban = syn.make_urban(160, 180, cell_size=500.0)
    grids = (dem, urban)
    base_t = syn.make_stations(dem, n, proxies=grids, seed=seed, supplemental_frac=0.05)
    base_h = syn.make_stations(dem, n, proxies=grids, variable="humidity", seed=seed + 1)
    base_p = syn.make_stations(dem, n, proxies=grids, variable="precipitation", seed=seed + 2)
    rng = np.random.default_rng(seed)
    hours = []
    for h in range(HOURS):
        t = base_t.copy()
        t.value[:] = (base_t.value + rng.normal(0, 0.4, n) + 2.0 * np.sin(h)).astype(np.float32)
        rh = base_h.copy()
        rh.value[:] = np.clip(base_h.value + rng.normal(0, 5, n), 2, 100).astype(np.float32)
        rh.value[rng.choice(n, 2, replace=False)] = np.float32(0.0)           # RH == 0: no dew point
        own_t = (14.0 - 0.006 * rh.z + rng.normal(0, 0.5, n)).astype(np.float32)  # the humidity stations' own thermometers ...
        own_t[rng.random(n) < 0.35] = np.float32(A.NODATA)                      # ... a third of them have none
        if h == 1:
            rh.x[0] += 1.0e6                                                     # one such station lies outside the map
            own_t[0] = np.float32(A.NODATA)
        p = base_p.copy()
        p.value[:] = np.maximum(base_p.value + rng.normal(0, 1.5, n), 0).astype(np.float32)
        hours.append((t, rh, own_t, p))
    return dem, urban, hours


def test_hourly_chain_matches_the_reference(engine, ref):
    dem, urban, hours = inputs()
    grids = (dem, urban)
    x, y, first, mx = geometry(dem, cell_m=5000.0)
    n_cells = mx.size
    did, uid = engine.upload(dem), engine.upload(urban)
    aid = engine.aggregation_upload(x, y, first, mx)
    try:
        batch, want = [], []
        for (t, rh, own_t, p) in hours:
            s_t, q_t = settings_for(A.airTemperature)
            s_h, q_h = settings_for(A.airDewTemperature)
            s_p, q_p = settings_for(A.precipitation)
            batch.append([dict(meteo=t, settings=clone(s_t), quality_settings=clone(q_t), variable=A.airTemperature,
                               aggr_method=A.PB_AGGR_AVERAGE, want_raster=True),
                          dict(meteo=rh, settings=clone(s_h), quality_settings=clone(q_h), variable=A.airRelHumidity, use_dew_point=True,
                               air_temperature=own_t, aggr_method=A.PB_AGGR_MEDIAN, want_raster=True),
                          dict(meteo=p, settings=clone(s_p), quality_settings=clone(q_p), variable=A.precipitation,
                               aggr_method=A.PB_AGGR_AVERAGE)])
            # ---- the reference chain
            r_t = ref.interpolation_dem(t, s_t, A.airTemperature, dem, grids, quality_settings=q_t, cross_validate=True)
            map_t = A.Raster(r_t["grid"], dem.llx, dem.lly, dem.cell_size, dem.flag)
            cells_t = ref.spatial_aggregate_meteo_grid(map_t, x, y, first, mx, A.PB_AGGR_AVERAGE)
            air = own_t.copy()
            need = (rh.value != np.float32(A.NODATA)) & (own_t == np.float32(A.NODATA))
            val, ing = ref_sample(ref, map_t, rh.x[need], rh.y[need])
            idx = np.flatnonzero(need)
            air[idx[ing]] = val[ing]
            tdew = ref.tdew_from_relhum(rh.value, air)
            have = tdew != np.float32(A.NODATA)
            dew = rh.subset(have)
            dew.value[:] = tdew[have]
            r_d = ref.interpolation_dem(dew, s_h, A.airDewTemperature, dem, grids, quality_settings=q_h, cross_validate=True)
            map_d = A.Raster(r_d["grid"], dem.llx, dem.lly, dem.cell_size, dem.flag)
            ok, map_rh, mn, mxv = ref.relative_humidity_map(map_t, map_d)
            cells_h = ref.spatial_aggregate_meteo_grid(A.Raster(map_rh, dem.llx, dem.lly, dem.cell_size, dem.flag), x, y, first, mx,
                                                       A.PB_AGGR_MEDIAN)
            r_p = ref.interpolation_dem(p, s_p, A.precipitation, dem, grids, quality_settings=q_p, cross_validate=True)
            cells_p = ref.spatial_aggregate_meteo_grid(A.Raster(r_p["grid"], dem.llx, dem.lly, dem.cell_size, dem.flag), x, y, first,
                                                       mx, A.PB_AGGR_AVERAGE)
            want.append(dict(t=r_t, cells_t=cells_t, d=r_d, have=have, rh=map_rh, rh_minmax=(mn, mxv), cells_h=cells_h, p=r_p,
                             cells_p=cells_p))
        got = engine.meteo_grid_step_batch(batch, did, dem.shape, (did, uid), aggregation_id=aid, n_cells=n_cells, lanes=2,
                                           cross_validate=True)
        nod, flag = np.float32(A.NODATA), np.float32(dem.flag)

        def cells_close(a, b):
            assert np.array_equal(a == nod, b == nod)
            m = b != nod
            assert np.abs(a[m].astype(np.float64) - b[m]).max() <= TOL

        def grids_close(a, b):
            assert np.array_equal(a == flag, b == flag)
            m = b != flag
            assert np.abs(a[m].astype(np.float64) - b[m]).max() <= TOL

        for h, (g, w) in enumerate(zip(got, want)):
            gt, gh, gp = g
            assert gt["ok"] and gh["ok"] and gp["ok"], (gt["status"], gh["status"], gp["status"])
            grids_close(gt["grid"], w["t"]["grid"])
            assert np.array_equal(gt["wrong_spatial"], w["t"]["wrong_spatial"])
            cells_close(gt["cells"], w["cells_t"])
            # humidity chain: flags / residuals of the dew-point gridding in the order of the humidity stations
            have = w["have"]
            assert np.array_equal(gh["wrong_spatial"][have], w["d"]["wrong_spatial"])
            assert not gh["wrong_spatial"][~have].any()
            res_ref = w["d"]["residual"]
            res_got = gh["residual"][have]
            assert (gh["residual"][~have] == nod).all()
            assert np.array_equal(res_got == nod, res_ref == nod)
            mm = res_ref != nod
            assert np.abs(res_got[mm].astype(np.float64) - res_ref[mm]).max() <= TOL
            grids_close(gh["grid"], w["rh"])
            assert abs(gh["min"] - w["rh_minmax"][0]) <= TOL and abs(gh["max"] - w["rh_minmax"][1]) <= TOL
            cells_close(gh["cells"], w["cells_h"])
            cells_close(gp["cells"], w["cells_p"])
            assert gp["grid"] is None
        # a humidity chain without the hour's temperature map is refused for that variable only
        t, rh, own_t, p = hours[0]
        s_h, q_h = settings_for(A.airDewTemperature)
        import praga_b200 as pb
        with pytest.raises(pb.PragaB200Error):
            engine.meteo_grid_step_batch([[dict(meteo=rh, settings=s_h, quality_settings=q_h, variable=A.airRelHumidity, use_dew_point=True,
                                                air_temperature=own_t)]], did, dem.shape, (did, uid), aggregation_id=aid, n_cells=n_cells)
    finally:
        engine.aggregation_free(aid)
        engine.free(uid)
        engine.free(did)
