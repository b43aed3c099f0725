"""BASELINE config 5 in miniature: an hourly batch (air temperature, relative humidity, precipitation) gridded step by
step the way Project::interpolationDem does (project.cpp:3106-3150) — spatial quality control, preInterpolation,
interpolationRaster — plus the leave-one-out of Project::interpolationCv, with the time steps sharded round-robin over
"ranks" (praga_b200.sharding; no collective). Every step is checked against the same chain of the reference's own
functions (oracle/_ref): identical quality flags and settings, rasters and residuals within 1e-4, NODATA masks equal."""
import numpy as np
import pytest

from praga_b200 import _abi as A
from praga_b200 import sharding
from praga_b200 import synthetic as syn
from test_oracle_port import clone, describe_diff, settings_equal

TOL = 1e-4
pytestmark = pytest.mark.gpu

HOURS = 4
VARS = (A.airTemperature, A.airRelHumidity, A.precipitation)


def batch_inputs(n=260, seed=33):
    dem = syn.make_dem(150, 170, seed=5, cell_size=600.0)
    urban = syn.make_urban(150, 170, cell_size=600.0)
    base = {A.airTemperature: syn.make_stations(dem, n, proxies=(dem, urban), seed=seed, supplemental_frac=0.08),
            A.airRelHumidity: syn.make_stations(dem, n, proxies=(dem, urban), variable="humidity", seed=seed),
            A.precipitation: syn.make_stations(dem, n, proxies=(dem, urban), variable="precipitation", seed=seed)}
    rng = np.random.default_rng(seed)
    ar = np.zeros(n)
    steps = []
    for h in range(HOURS):
        ar = 0.8 * ar + rng.normal(0, 0.4, n)                       # seeded AR(1) anomaly per station
        for var in VARS:
            st = base[var].copy()
            if var == A.airTemperature:
                st.value[:] = (base[var].value + ar + 3.0 * np.sin(h / 24 * 2 * np.pi)).astype(np.float32)
                bad = rng.choice(n, 4, replace=False)                # gross errors for the spatial check
                st.value[bad] += np.float32(9.0) * np.where(rng.random(4) < 0.5, -1, 1).astype(np.float32)
            elif var == A.airRelHumidity:
                st.value[:] = np.clip(base[var].value + 6 * ar, 0, 100).astype(np.float32)
                st.value[rng.choice(n, 3, replace=False)] = np.float32(3.0)
            else:
                wet = h != 2                                          # hour 2: every station below the threshold
                st.value[:] = (np.maximum(base[var].value + 2 * ar, 0) if wet else 0.1 * rng.random(n)).astype(np.float32)
            steps.append((h, var, st))
    return dem, urban, steps


def settings_for(var):
    if var == A.airTemperature:
        s = syn.settings_multiple_detrending()
        s.useLapseRateCode = 1
    else:
        s = A.default_settings()
        s.nProxy = 2
        s.proxy[0] = A.default_proxy(A.proxyHeight, 0)
        s.proxy[1] = A.default_proxy(A.proxyUrbanFraction, 1)
    sq = syn.settings_classic_detrending(2)
    sq.useLapseRateCode = s.useLapseRateCode
    return s, sq


def test_hourly_batch_time_sharded_matches_reference(engine, ref):
    dem, urban, steps = batch_inputs()
    dem_id, urb_id = engine.upload(dem), engine.upload(urban)
    world = 2
    done = {}
    flagged = 0
    try:
        for rank in range(world):                                    # one process per GPU in production; same units here
            for k in sharding.time_steps_for_rank(len(steps), rank, world):
                h, var, st = steps[k]
                s_g, sq_g = settings_for(var)
                s_r, sq_r = clone(s_g), clone(sq_g)
                g = engine.interpolation_dem(st.copy(), s_g, var, dem_id, dem.shape, (dem_id, urb_id), quality_settings=sq_g,
                                             cross_validate=True)
                r = ref.interpolation_dem(st.copy(), s_r, var, dem, (dem, urban), quality_settings=sq_r, cross_validate=True,
                                          threads=8)
                tag = f"hour {h} var {var}"
                assert g["ok"] == r["ok"], tag
                assert np.array_equal(g["wrong_spatial"], r["wrong_spatial"]), tag
                flagged += int(g["wrong_spatial"].sum())
                assert settings_equal(sq_g, sq_r), tag + ": quality settings: " + describe_diff(sq_g, sq_r)
                assert settings_equal(s_g, s_r), tag + ": settings: " + describe_diff(s_g, s_r)
                mg, mr = g["grid"] == np.float32(dem.flag), r["grid"] == np.float32(dem.flag)
                assert np.array_equal(mg, mr), tag
                d = np.abs(g["grid"][~mr].astype(np.float64) - r["grid"][~mr].astype(np.float64))
                assert d.max() <= TOL, f"{tag}: raster max |gpu-ref| = {d.max():.3e}"
                assert g["n_valid"] == r["n_valid"] and abs(g["min"] - r["min"]) <= TOL and abs(g["max"] - r["max"]) <= TOL, tag
                assert np.array_equal(g["residual"] == np.float32(-9999), r["residual"] == np.float32(-9999)), tag
                dr = np.abs(g["residual"].astype(np.float64) - r["residual"].astype(np.float64))
                assert dr.max() <= TOL, f"{tag}: residual max |gpu-ref| = {dr.max():.3e}"
                assert g["stats"]["n"] == r["stats"]["n"], tag
                for key in ("meanAbsoluteError", "meanBiasError", "rootMeanSquareError", "nashSutcliffe", "R2"):
                    assert abs(g["stats"][key] - r["stats"][key]) <= TOL, f"{tag}: {key}"
                done[k] = rank
    finally:
        engine.free(dem_id)
        engine.free(urb_id)
    assert sorted(done) == list(range(len(steps)))                   # every unit gridded exactly once
    assert flagged > 0                                                # the planted errors were caught somewhere
    # the all-dry hour is the precipitationAllZero short-circuit: a raster of zeros
    assert any(var == A.precipitation and h == 2 for h, var, _ in steps)


def test_interpolation_dem_without_quality_control_and_device_only(engine, ref):
    dem, urban, steps = batch_inputs(n=120, seed=5)
    dem_id, urb_id = engine.upload(dem), engine.upload(urban)
    try:
        h, var, st = steps[0]
        s_g, _ = settings_for(var)
        s_r = clone(s_g)
        g = engine.interpolation_dem(st.copy(), s_g, var, dem_id, dem.shape, (dem_id, urb_id), device_only=True)
        r = ref.interpolation_dem(st.copy(), s_r, var, dem, (dem, urban), threads=8)
        assert g["ok"] and r["ok"] and g["n_valid"] == r["n_valid"]
        assert abs(g["min"] - r["min"]) <= TOL and abs(g["max"] - r["max"]) <= TOL
        assert not g["wrong_spatial"].any()
        assert settings_equal(s_g, s_r), describe_diff(s_g, s_r)
        assert engine.timing()["launches"] >= 2
    finally:
        engine.free(dem_id)
        engine.free(urb_id)


def test_batch_call_equals_stand_alone_calls(engine, ref):
    """pb_interpolation_dem_batch (units spread over internal lanes / streams) returns, unit by unit, exactly what the
    stand-alone call returns: rasters, flags, residuals, statistics and the settings written back."""
    dem, urban, steps = batch_inputs(n=200, seed=71)
    dem_id, urb_id = engine.upload(dem), engine.upload(urban)
    try:
        units, alone = [], []
        for h, var, st in steps:
            s_a, sq_a = settings_for(var)
            alone.append((engine.interpolation_dem(st.copy(), s_a, var, dem_id, dem.shape, (dem_id, urb_id), quality_settings=sq_a,
                                                   cross_validate=True), s_a, sq_a))
            s_b, sq_b = settings_for(var)
            units.append((st.copy(), s_b, sq_b, var))
        for lanes in (1, 3, 5):
            fresh = [(u[0].copy(), clone(settings_for(u[3])[0]), clone(settings_for(u[3])[1]), u[3]) for u in units]
            got = engine.interpolation_dem_batch(fresh, dem_id, dem.shape, (dem_id, urb_id), cross_validate=True, lanes=lanes)
            assert len(got) == len(steps)
            for (a, s_a, sq_a), g, u in zip(alone, got, fresh):
                assert g["ok"] == a["ok"]
                assert np.array_equal(g["grid"], a["grid"]), lanes
                assert np.array_equal(g["wrong_spatial"], a["wrong_spatial"])
                assert np.array_equal(g["residual"], a["residual"])
                assert g["stats"] == a["stats"] and g["n_valid"] == a["n_valid"] and g["min"] == a["min"] and g["max"] == a["max"]
                assert settings_equal(u[1], s_a), describe_diff(u[1], s_a)
                assert settings_equal(u[2], sq_a), describe_diff(u[2], sq_a)
        assert engine.timing()["launches"] > len(steps)
    finally:
        engine.free(dem_id)
        engine.free(urb_id)


def test_interpolation_dem_with_the_demo_shipped_settings(engine, ref):
    """Project::interpolationDem with the settings the bundled demo ships (parameters.ini: algorithm=shepard_modified,
    optimalDetrending=true, thermalInversion=true, lapseRateCode=true) for the gridding AND the quality settings: spatial
    quality control, optimalDetrending cross-validated with the Shepard estimator, Shepard raster, Shepard leave-one-out."""
    dem, urban, steps = batch_inputs(n=180, seed=9)
    dem_id, urb_id = engine.upload(dem), engine.upload(urban)
    try:
        for h, var, st in steps[:3]:
            def shipped():
                s = syn.settings_classic_detrending(2)
                s.method = A.shepard_modified
                s.useBestDetrending = 1
                s.useLapseRateCode = 1
                return s
            s_g, sq_g = shipped(), shipped()
            s_r, sq_r = clone(s_g), clone(sq_g)
            g = engine.interpolation_dem(st.copy(), s_g, var, dem_id, dem.shape, (dem_id, urb_id), quality_settings=sq_g,
                                         cross_validate=True)
            r = ref.interpolation_dem(st.copy(), s_r, var, dem, (dem, urban), quality_settings=sq_r, cross_validate=True, threads=8)
            tag = f"hour {h} var {var}"
            assert g["ok"] == r["ok"], tag
            assert np.array_equal(g["wrong_spatial"], r["wrong_spatial"]), tag
            assert settings_equal(sq_g, sq_r), tag + ": quality settings: " + describe_diff(sq_g, sq_r)
            assert settings_equal(s_g, s_r), tag + ": settings: " + describe_diff(s_g, s_r)
            mr = r["grid"] == np.float32(dem.flag)
            assert np.array_equal(g["grid"] == np.float32(dem.flag), mr), tag
            assert np.abs(g["grid"][~mr].astype(np.float64) - r["grid"][~mr]).max() <= TOL, tag
            assert np.array_equal(g["residual"] == np.float32(-9999), r["residual"] == np.float32(-9999)), tag
            assert np.abs(g["residual"].astype(np.float64) - r["residual"]).max() <= TOL, tag
    finally:
        engine.free(dem_id)
        engine.free(urb_id)
