#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference (oracle/_ref, built from
/root/reference) in this container. Re-run with:  python tests/golden/make_golden.py

  c1_demo.npz   BASELINE.json configs[0]: the bundled DATA/ demo — DEM_ER_450m (635x337 @450 m) and the stations of
                DATA/METEOPOINT/test.db for DAILY_TMIN 2023-01-05 and DAILY_TMAX 2023-01-10 — gridded with
                algorithm=idw (override of the demo's shepard_modified), classic elevation detrending with thermal
                inversion, optimalDetrending=false, climate lapse rates of January from parameters.ini.
  c1_demo_shipped_*.npz  the same demo day with the settings AS SHIPPED (DATA/PROJECT/test/SETTINGS/parameters.ini):
                algorithm=shepard_modified, optimalDetrending=true, thermalInversion=true, lapseRateCode=true; also stores the
                settings before preInterpolation so that the fit itself (all combinations, leave-one-out, argmin) is checked.
  synth_*.npz   small seeded synthetic cases (praga_b200/synthetic.py) with the reference's outputs.
Each file stores the inputs, the settings AFTER preInterpolation (raw bytes of pb_settings), the detrended station
values, the keep mask and the reference output grid.
"""
import os
import sqlite3
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.binding import Oracle  # noqa: E402
from praga_b200 import _abi as A  # noqa: E402
from praga_b200 import synthetic as syn  # noqa: E402

REF_DATA = "/root/reference/DATA"


def read_flt(base):
    hdr = {}
    for line in open(base + ".hdr"):
        k, v = line.split()
        hdr[k.lower()] = v
    rows, cols = int(hdr["nrows"]), int(hdr["ncols"])
    data = np.fromfile(base + ".flt", dtype="<f4").reshape(rows, cols)
    return A.Raster(data, float(hdr["xllcorner"]), float(hdr["yllcorner"]), float(hdr["cellsize"]), float(hdr["nodata_value"]))


def demo_stations(var_id, date):
    con = sqlite3.connect(f"file:{REF_DATA}/METEOPOINT/test.db?mode=ro", uri=True)
    cur = con.cursor()
    pts = cur.execute("select id_point, utm_x, utm_y, altitude from point_properties where is_active=1 order by id_point").fetchall()
    xs, ys, zs, vs = [], [], [], []
    for pid, x, y, z in pts:
        try:
            row = cur.execute(f"select value from '{pid}_D' where id_variable=? and date(date_time)=?", (var_id, date)).fetchone()
        except sqlite3.OperationalError:
            continue
        if row is None or row[0] is None:
            continue
        xs.append(x); ys.append(y); zs.append(z); vs.append(row[0])
    z = np.array(zs, dtype=np.float64)
    return A.Stations(xs, ys, z, vs, z.astype(np.float32).reshape(-1, 1))


def save(name, dem, grids, st_raw, st_det, keep, s, var, out, ok, mn, mx, nv):
    arrs = dict(dem=dem.data, dem_hdr=np.array([dem.llx, dem.lly, dem.cell_size, dem.flag]),
                x=st_raw.x, y=st_raw.y, z=st_raw.z, value_raw=st_raw.value, value_detrended=st_det.value, keep=keep,
                proxy_values=st_raw.proxy_values, settings=np.frombuffer(bytes(s), dtype=np.uint8), variable=np.int32(var),
                out=out, ok=np.bool_(ok), minimum=np.float32(mn), maximum=np.float32(mx), n_valid=np.int64(nv))
    if st_raw.lapse_rate_code is not None:
        arrs["lapse_rate_code"] = st_raw.lapse_rate_code
    for i, g in enumerate(grids):
        if g is not dem:
            arrs[f"grid{i}"] = g.data
            arrs[f"grid{i}_hdr"] = np.array([g.llx, g.lly, g.cell_size, g.flag])
    np.savez_compressed(os.path.join(HERE, name), **arrs)
    print(name, os.path.getsize(os.path.join(HERE, name)) // 1024, "KiB", "ok" if ok else "FALSE", mn, mx, nv)


def run_case(ref, name, dem, grids, st, s, var):
    raw = st.copy()
    ok_pre, keep = ref.pre_interpolation(st, s, var, grids)
    assert ok_pre
    ok, out, mn, mx, nv, _ = ref.interpolation_raster(st.subset(keep), s, var, dem, grids)
    save(name, dem, grids, raw, st, keep, s, var, out, ok, mn, mx, nv)


def run_case_ex(ref, name, dem, grids, st, s, var):
    """preInterpolation with its meteo points (optimalDetrending re-runs passDataToInterpolation and cross-validates)."""
    raw = st.copy()
    pre = np.frombuffer(bytes(s), dtype=np.uint8).copy()
    ok_pre, keep, _ = ref.pre_interpolation_ex(st, s, var, current_value=raw.value)
    assert ok_pre
    ok, out, mn, mx, nv, _ = ref.interpolation_raster(st.subset(keep), s, var, dem, grids)
    save(name, dem, grids, raw, st, keep, s, var, out, ok, mn, mx, nv)
    path = os.path.join(HERE, name)
    g = dict(np.load(path))
    g["settings_before"] = pre
    np.savez_compressed(path, **g)


def run_local_case(ref, name, dem, grids, st, s, var):
    """loop of Project::interpolationDemLocalDetrending: raw points in, no global preInterpolation."""
    ok, out, mn, mx, nv, _ = ref.local_detrending_raster(st, s, var, dem, grids, threads=8)
    save(name, dem, grids, st, st, np.ones(st.n, dtype=bool), s, var, out, ok, mn, mx, nv)


def run_kriging_case(ref, name, mode, n, m, rng_a, nugget, sill, slope, seed):
    rng = np.random.default_rng(seed)
    pos = rng.random((n, 2)) * 2.5e5 + np.array([5.0e5, 4.8e6])
    val = rng.normal(11.0, 3.0, n)
    xy = rng.random((m, 2)) * 2.5e5 + np.array([5.0e5, 4.8e6])
    ok, res, rm, _, _ = ref.kriging_points(pos, val, mode, rng_a, nugget, sill, slope, xy)
    assert ok
    np.savez_compressed(os.path.join(HERE, name), pos=pos, val=val, xy=xy, params=np.array([mode, rng_a, nugget, sill, slope]),
                        result=res, rmse=rm)
    print(name, os.path.getsize(os.path.join(HERE, name)) // 1024, "KiB")


def run_kriging_c4(ref, name):
    """BASELINE config 4's own system: 2000 stations in the 1000 km box, spherical, range 2e5 m, nugget 0.1, sill 4 (the
    generator state of bench.py KrigingWorkload); 320 targets, 8 of them 30 m from a station."""
    n, m = 2000, 320
    rng = np.random.default_rng(1)
    pos = np.stack([syn.LLX + rng.random(n) * 1.0e6, syn.LLY + rng.random(n) * 1.0e6], axis=1)
    val = rng.normal(12.0, 3.0, n)
    rq = np.random.default_rng(77)
    xy = np.stack([syn.LLX + rq.random(m) * 1.0e6, syn.LLY + rq.random(m) * 1.0e6], axis=1)
    xy[:8] = pos[:8] + 30.0
    ok, res, rm, _, _ = ref.kriging_points(pos, val, A.KRIGING_SPHERICAL, 2.0e5, 0.1, 4.0, 0.0, xy)
    assert ok
    np.savez_compressed(os.path.join(HERE, name), pos=pos, val=val, xy=xy, params=np.array([A.KRIGING_SPHERICAL, 2.0e5, 0.1, 4.0, 0.0]),
                        result=res, rmse=rm)
    print(name, os.path.getsize(os.path.join(HERE, name)) // 1024, "KiB")


def large_station_sets(ref):
    """K1 where it is benchmarked: 5000 stations (10 tiles resident in the persistent kernel, the c3idw shape) and 9500
    stations (more than the 9216 the persistent kernel holds: the tiled idw_kernel<2,false>), multiple detrending, on a
    small DEM so that the fixture stays small."""
    for n, seed in ((5000, 301), (9500, 302)):
        d = syn.make_dem(72, 88, seed=seed, cell_size=1000.0)
        u = syn.make_urban(72, 88, cell_size=1000.0)
        st = syn.make_stations(d, n, proxies=(d, u), seed=seed + 10)
        s = syn.settings_multiple_detrending()
        syn.set_points_range(s, st)
        run_case(ref, f"synth_idw_s{n}.npz", d, (d, u), st, s, A.airTemperature)


def main():
    ref = Oracle("reference")
    if "--round2" in sys.argv:          # only the fixtures added in round 2 (the others are unchanged)
        run_kriging_c4(ref, "kriging_c4_n2000.npz")
        large_station_sets(ref)
        return
    # ---- C1: bundled demo ----
    dem = read_flt(f"{REF_DATA}/DEM/DEM_ER_450m")
    for tag, var_id, var, date, lapse in (("tmin_20230105", 151, A.dailyAirTemperatureMin, "2023-01-05", -0.001),
                                          ("tmax_20230110", 152, A.dailyAirTemperatureMax, "2023-01-10", -0.001)):
        st = demo_stations(var_id, date)
        s = syn.settings_classic_detrending(1)
        s.climateLapseRate = lapse         # [climate] tmin_lapserate / tmax_lapserate, January (parameters.ini)
        run_case(ref, f"c1_demo_{tag}.npz", dem, (dem,), st, s, var)
    # ---- C1 as shipped: algorithm=shepard_modified, optimalDetrending=true, thermalInversion=true, lapseRateCode=true ----
    # two proxies like the shipped project: elevation (table field) + the orography-index raster of DATA/GEO (500 m grid, a
    # geometry foreign to the 450 m DEM; station values sampled from the grid as Project::readProxyValues does, project.cpp:2066-2091)
    ti = read_flt(f"{REF_DATA}/GEO/TI_DEM_ER_bacini_500m")
    st = demo_stations(151, "2023-01-05")
    tiv = syn.sample_grid(ti, st.x, st.y)
    tiv = np.where(tiv == np.float32(ti.flag), np.float32(A.NODATA), tiv)
    st = A.Stations(st.x, st.y, st.z, st.value, np.stack([st.z.astype(np.float32), tiv], axis=1))
    s = syn.settings_classic_detrending(2)
    p1 = A.default_proxy(A.proxyOrogIndex, 1)
    s.proxy[1] = p1
    s.method = A.shepard_modified
    s.useBestDetrending = 1
    s.useLapseRateCode = 1
    s.climateLapseRate = -0.001
    run_case_ex(ref, "c1_demo_shipped_tmin_20230105.npz", dem, (dem, ti), st, s, A.dailyAirTemperatureMin)
    # ---- synthetic ----
    d = syn.make_dem(60, 70, seed=101)
    u = syn.make_urban(60, 70)
    st = syn.make_stations(d, 90, proxies=(d, u), seed=102, supplemental_frac=0.1)
    s = syn.settings_multiple_detrending()
    s.useLapseRateCode = 1
    syn.set_points_range(s, st)
    run_case(ref, "synth_multiple.npz", d, (d, u), st, s, A.airTemperature)
    st = syn.make_stations(d, 70, variable="precipitation", seed=103)
    run_case(ref, "synth_precipitation.npz", d, (), st, A.default_settings(), A.dailyPrecipitation)
    st = syn.make_stations(d, 70, variable="humidity", seed=104)
    run_case(ref, "synth_humidity.npz", d, (), st, A.default_settings(), A.airRelHumidity)
    # topographic distance (kh = 10) on a 400 m DEM
    dt = syn.make_dem(60, 70, seed=105, cell_size=400.0)
    st = syn.make_stations(dt, 50, variable="humidity", seed=106)
    s = A.default_settings()
    s.useTD = 1
    s.topoDist_Kh = 10
    run_case(ref, "synth_topographic_distance.npz", dt, (), st, s, A.airRelHumidity)
    # local detrending at config-3 station density
    dl = syn.make_dem(36, 44, seed=107, cell_size=9000.0)
    ul = syn.make_urban(36, 44, cell_size=9000.0)
    st = syn.make_stations(dl, 640, proxies=(dl, ul), seed=108)
    s = syn.settings_multiple_detrending()
    s.useLocalDetrending = 1
    s.minPointsLocalDetrending = 20
    syn.set_points_range(s, st)
    run_local_case(ref, "local_detrending.npz", dl, (dl, ul), st, s, A.airTemperature)
    # ordinary kriging (kriging.cpp), one fixture per variogram model
    run_kriging_case(ref, "kriging_spherical.npz", A.KRIGING_SPHERICAL, 140, 200, 8.0e4, 0.1, 4.0, 0.0, 201)
    run_kriging_case(ref, "kriging_exponential.npz", A.KRIGING_EXPONENTIAL, 120, 150, 9.0e4, 0.2, 3.0, 0.0, 202)
    run_kriging_case(ref, "kriging_linear.npz", A.KRIGING_LINEAR, 100, 150, 8.0e4, 0.15, 4.0, 2.0e-5, 203)
    run_kriging_c4(ref, "kriging_c4_n2000.npz")
    large_station_sets(ref)


if __name__ == "__main__":
    main()
