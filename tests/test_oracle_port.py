"""CPU-only: pins oracle/oracle_port.cpp (the restatement) against the reference's own code (oracle/_ref) on seeded
inputs. The raster estimator must agree BIT-EXACTLY (same arithmetic, no FMA), the station-space fits too."""
import copy
import ctypes

import numpy as np
import pytest

from praga_b200 import _abi as A
from praga_b200 import synthetic as syn


@pytest.fixture(scope="module")
def port():
    from oracle import binding
    binding.build("port")
    return binding.Oracle("port")


def clone(s):
    c = A.pb_settings()
    ctypes.memmove(ctypes.byref(c), ctypes.byref(s), ctypes.sizeof(s))
    return c


def settings_equal(a, b):
    return bytes(a) == bytes(b)


def describe_diff(a, b):
    out = []
    for name, _ in A.pb_settings._fields_:
        va, vb = getattr(a, name), getattr(b, name)
        if name == "proxy":
            for i in range(a.nProxy):
                for pn, _ in A.pb_proxy._fields_:
                    pa, pb_ = getattr(va[i], pn), getattr(vb[i], pn)
                    if hasattr(pa, "__len__"):
                        pa, pb_ = np.array(pa).tolist(), np.array(pb_).tolist()
                    if pa != pb_:
                        out.append(f"proxy[{i}].{pn}: {pa} != {pb_}")
        else:
            if hasattr(va, "__len__"):
                va, vb = np.array(va).tolist(), np.array(vb).tolist()
            if va != vb:
                out.append(f"{name}: {va} != {vb}")
    return "; ".join(out[:6])


@pytest.mark.parametrize("case", ["multiple", "multiple_three", "classic", "classic_inversion", "classic_noinv", "prec", "humidity"])
def test_pre_interpolation_and_raster_bit_exact(ref, port, case):
    dem = syn.make_dem(90, 110, seed=5)
    urban = syn.make_urban(90, 110)
    var = A.airTemperature
    grids = (dem, urban)
    if case in ("multiple", "multiple_three"):
        st = syn.make_stations(dem, 160, proxies=grids, seed=17, supplemental_frac=0.1)
        s = syn.settings_multiple_detrending()
        s.useLapseRateCode = 1
        if case == "multiple_three":
            el = s.proxy[0]
            el.fittingFunction = A.piecewiseThree
            el.nFitParams = 5
            for j, (lo, hi, fg) in enumerate(zip((0, -30, 100, -0.01, -0.015), (1500, 50, 1000, 0.01, -0.002), (0, 1, 1, 1, 1))):
                el.fitMin[j], el.fitMax[j], el.fitFirstGuess[j] = lo, hi, fg
            s.proxy[0] = el
        syn.set_points_range(s, st)
    elif case.startswith("classic"):
        st = syn.make_stations(dem, 140, proxies=grids, seed=19)
        if case == "classic_inversion":
            st.value[:] = np.where(st.z < 700, st.value - 0.012 * (700 - st.z), st.value).astype(np.float32)
        s = syn.settings_classic_detrending(2)
        if case == "classic_noinv":
            s.useThermalInversion = 0
            s.selectedUseThermalInversion = s.currentUseThermalInversion = 0
        var = A.dailyAirTemperatureMin
    elif case == "prec":
        st = syn.make_stations(dem, 80, variable="precipitation", seed=23)
        s = A.default_settings()
        var, grids = A.dailyPrecipitation, ()
    else:
        st = syn.make_stations(dem, 80, variable="humidity", seed=29)
        s = A.default_settings()
        var, grids = A.airRelHumidity, ()

    st_r, st_p, s_r, s_p = st.copy(), st.copy(), clone(s), clone(s)
    ok_r, keep_r = ref.pre_interpolation(st_r, s_r, var, grids)
    ok_p, keep_p = port.pre_interpolation(st_p, s_p, var, grids)
    assert ok_r == ok_p
    assert np.array_equal(keep_r, keep_p)
    assert settings_equal(s_r, s_p), describe_diff(s_r, s_p)
    assert np.array_equal(st_r.value, st_p.value), f"max diff {np.abs(st_r.value - st_p.value).max()}"

    a = ref.interpolation_raster(st_r.subset(keep_r), s_r, var, dem, grids)
    b = port.interpolation_raster(st_p.subset(keep_p), s_p, var, dem, grids)
    assert a[0] == b[0] and a[4] == b[4]
    assert np.array_equal(a[1], b[1]), f"max diff {np.abs(a[1] - b[1]).max()}"
    assert a[2] == b[2] and a[3] == b[3]


def test_points_and_residuals_bit_exact(ref, port):
    dem = syn.make_dem(80, 80, seed=2)
    urban = syn.make_urban(80, 80)
    st = syn.make_stations(dem, 120, proxies=(dem, urban), seed=3)
    raw = st.value.copy()
    s = syn.settings_multiple_detrending()
    syn.set_points_range(s, st)
    ok, keep = ref.pre_interpolation(st, s, A.airTemperature, (dem, urban))
    assert keep.all()
    ra, sa = ref.compute_residuals(st, s, A.airTemperature, raw)
    rb, sb = port.compute_residuals(st, s, A.airTemperature, raw)
    assert np.array_equal(ra, rb)
    assert sa == sb


def test_kriging_bit_exact(ref, port):
    rng = np.random.default_rng(4)
    n = 60
    pos = rng.random((n, 2)) * 1e5
    val = rng.normal(10, 3, n)
    xy = rng.random((25, 2)) * 1e5
    for mode in (A.KRIGING_SPHERICAL, A.KRIGING_EXPONENTIAL, A.KRIGING_GAUSSIAN, A.KRIGING_LINEAR):
        a = ref.kriging_points(pos, val, mode, 4e4, 0.1, 4.0, 1e-4, xy)
        b = port.kriging_points(pos, val, mode, 4e4, 0.1, 4.0, 1e-4, xy)
        assert a[0] and b[0]
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]), mode


def test_topographic_distance_raster_and_points_bit_exact(ref, port):
    """computeDistances with getUseTD() (interpolation.cpp:226-251) + gis::topographicDistance (gis.cpp:1595-1647)."""
    dem = syn.make_dem(70, 80, seed=6, cell_size=400.0)
    st = syn.make_stations(dem, 45, variable="humidity", seed=7)
    s = A.default_settings()
    s.useTD = 1
    s.topoDist_Kh = 8
    a = ref.interpolation_raster(st, s, A.airRelHumidity, dem, ())
    b = port.interpolation_raster(st, s, A.airRelHumidity, dem, ())
    assert a[0] == b[0] and np.array_equal(a[1], b[1])
    s0 = A.default_settings()
    c = ref.interpolation_raster(st, s0, A.airRelHumidity, dem, ())
    assert not np.array_equal(a[1], c[1]), "topographic distance must change the result"
    x, y, z = st.x.astype(np.float32), st.y.astype(np.float32), st.z.astype(np.float32)
    pv = np.zeros((st.n, 0))
    pa = ref.interpolate_points_td(st, s, A.airRelHumidity, x, y, z, pv, dem)
    pb_ = port.interpolate_points_td(st, s, A.airRelHumidity, x, y, z, pv, dem)
    assert np.array_equal(pa, pb_)


def test_local_residuals_port_bit_exact(ref, port):
    """computeResidualsLocalDetrending (spatialControl.cpp:158-228) + CV statistics: restatement == reference."""
    dem = syn.make_dem(60, 70, cell_size=3000.0, seed=3)
    urban = syn.make_urban(60, 70, cell_size=3000.0)
    for n, codes in ((260, 0.15), (22, 0.0)):
        st = syn.make_stations(dem, n, proxies=(dem, urban), seed=9, supplemental_frac=codes)
        s = syn.settings_multiple_detrending()
        s.useLocalDetrending = 1
        s.minPointsLocalDetrending = 20
        s.useLapseRateCode = 1 if codes else 0
        syn.set_points_range(s, st)
        use = np.ones(st.n, dtype=np.uint8)
        use[::13] = 0
        ra, sa = ref.compute_residuals(st, clone(s), A.airTemperature, st.value, use_for_cv=use, local=True)
        rb, sb = port.compute_residuals(st, clone(s), A.airTemperature, st.value, use_for_cv=use, local=True)
        assert np.array_equal(ra, rb), np.abs(ra - rb).max()
        assert sa == sb
        assert sa["n"] > 10
