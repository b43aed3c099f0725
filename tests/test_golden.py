"""Golden fixtures generated from the unmodified reference (tests/golden/make_golden.py): the oracle port must
reproduce them bit-exactly on CPU; the CUDA path within 1e-4 with a bit-exact NODATA mask."""
import ctypes
import glob
import os

import numpy as np
import pytest

from praga_b200 import _abi as A

ALL = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))
KRIGING = [p for p in ALL if os.path.basename(p).startswith("kriging_")]
LOCAL = [p for p in ALL if os.path.basename(p).startswith("local_")]
GOLDEN = [p for p in ALL if p not in KRIGING and p not in LOCAL]


def load(path):
    g = np.load(path)
    hdr = g["dem_hdr"]
    dem = A.Raster(g["dem"], hdr[0], hdr[1], hdr[2], hdr[3])
    grids = []
    s = A.pb_settings()
    ctypes.memmove(ctypes.byref(s), g["settings"].tobytes(), ctypes.sizeof(s))
    ngrids = 0
    for i in range(s.nProxy):
        ngrids = max(ngrids, s.proxy[i].gridIndex + 1)
    for i in range(ngrids):
        if f"grid{i}" in g:
            h = g[f"grid{i}_hdr"]
            grids.append(A.Raster(g[f"grid{i}"], h[0], h[1], h[2], h[3]))
        else:
            grids.append(dem)
    codes = g["lapse_rate_code"] if "lapse_rate_code" in g else None
    raw = A.Stations(g["x"], g["y"], g["z"], g["value_raw"], g["proxy_values"], codes)
    det = A.Stations(g["x"], g["y"], g["z"], g["value_detrended"], g["proxy_values"], codes)
    return g, dem, tuple(grids), raw, det, s


def test_fixtures_present():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_port_reproduces_reference_raster(path):
    from oracle import binding
    port = binding.Oracle("port")
    g, dem, grids, raw, det, s = load(path)
    keep = g["keep"].astype(bool)
    ok, out, mn, mx, nv, _ = port.interpolation_raster(det.subset(keep), s, int(g["variable"]), dem, grids)
    assert ok == bool(g["ok"]) and nv == int(g["n_valid"])
    assert np.array_equal(out, g["out"])
    assert mn == g["minimum"] and mx == g["maximum"]


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_gpu_matches_reference_raster(engine, path):
    g, dem, grids, raw, det, s = load(path)
    keep = g["keep"].astype(bool)
    ok, out, mn, mx, nv = engine.interpolation_raster(det.subset(keep), s, int(g["variable"]), dem, grids)
    want = g["out"]
    assert ok == bool(g["ok"]) and nv == int(g["n_valid"])
    flag = np.float32(dem.flag)
    assert np.array_equal(out == flag, want == flag), "NODATA mask differs"
    err = np.abs(out.astype(np.float64) - want.astype(np.float64)).max()
    assert err <= 1e-4, err
    assert abs(mn - g["minimum"]) <= 1e-4 and abs(mx - g["maximum"]) <= 1e-4


@pytest.mark.parametrize("path", LOCAL, ids=[os.path.basename(p) for p in LOCAL])
def test_port_reproduces_reference_local_detrending(path):
    from oracle import binding
    port = binding.Oracle("port")
    g, dem, grids, raw, det, s = load(path)
    ok, out, mn, mx, nv, _ = port.local_detrending_raster(raw, s, int(g["variable"]), dem, grids, threads=4)
    assert ok == bool(g["ok"]) and nv == int(g["n_valid"])
    assert np.array_equal(out, g["out"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", LOCAL, ids=[os.path.basename(p) for p in LOCAL])
def test_gpu_matches_reference_local_detrending(engine, path):
    g, dem, grids, raw, det, s = load(path)
    ok, out, mn, mx, nv = engine.local_detrending_raster(raw, s, int(g["variable"]), dem, grids)
    want = g["out"]
    flag = np.float32(dem.flag)
    assert ok == bool(g["ok"]) and nv == int(g["n_valid"])
    assert np.array_equal(out == flag, want == flag)
    assert np.abs(out.astype(np.float64) - want.astype(np.float64)).max() <= 1e-4


@pytest.mark.parametrize("path", KRIGING, ids=[os.path.basename(p) for p in KRIGING])
def test_port_reproduces_reference_kriging(path):
    from oracle import binding
    port = binding.Oracle("port")
    g = np.load(path)
    mode, rng_a, nugget, sill, slope = g["params"]
    ok, res, rm, _, _ = port.kriging_points(g["pos"], g["val"], int(mode), rng_a, nugget, sill, slope, g["xy"])
    assert ok and np.array_equal(res, g["result"]) and np.array_equal(rm, g["rmse"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", KRIGING, ids=[os.path.basename(p) for p in KRIGING])
def test_gpu_matches_reference_kriging(engine, path):
    g = np.load(path)
    mode, rng_a, nugget, sill, slope = g["params"]
    kid = engine.kriging_setup(g["pos"], g["val"], int(mode), rng_a, nugget, sill, slope)
    res, rm = engine.kriging_points(kid, g["xy"])
    engine.kriging_free(kid)
    assert np.abs(res - g["result"]).max() <= 1e-4 and np.abs(rm - g["rmse"]).max() <= 1e-4


# ---- the demo day with the settings AS SHIPPED (shepard_modified + optimalDetrending): the fit itself is part of the fixture
SHIPPED = [p for p in GOLDEN if "shipped" in os.path.basename(p)]


def _before(g):
    s = A.pb_settings()
    ctypes.memmove(ctypes.byref(s), g["settings_before"].tobytes(), ctypes.sizeof(s))
    return s


def test_shipped_fixture_present():
    assert len(SHIPPED) >= 1


@pytest.mark.parametrize("path", SHIPPED, ids=[os.path.basename(p) for p in SHIPPED])
def test_port_reproduces_reference_optimal_detrending(path):
    from oracle import binding
    port = binding.Oracle("port")
    g, dem, grids, raw, det, s_after = load(path)
    s = _before(g)
    st = raw.copy()
    ok, keep, _ = port.pre_interpolation_ex(st, s, int(g["variable"]), current_value=raw.value)
    assert ok and np.array_equal(keep, g["keep"].astype(bool))
    assert np.array_equal(st.value, g["value_detrended"])
    assert bytes(s) == bytes(s_after)


@pytest.mark.gpu
@pytest.mark.parametrize("path", SHIPPED, ids=[os.path.basename(p) for p in SHIPPED])
def test_gpu_reproduces_reference_optimal_detrending(engine, path):
    g, dem, grids, raw, det, s_after = load(path)
    s = _before(g)
    st = raw.copy()
    keep, _ = engine.pre_interpolation_ex(st, s, int(g["variable"]), current_value=raw.value)
    assert np.array_equal(keep, g["keep"].astype(bool))
    assert np.array_equal(st.value, g["value_detrended"])
    assert bytes(s) == bytes(s_after)
