"""Station-space pipeline of preInterpolation (interpolation.cpp:2444-2499) beyond the Levenberg-Marquardt branch:
classic detrending() with the thermal-inversion heuristic, optimalDetrending() over every proxy combination and the
golden-section search of the topographic-distance coefficient kh. CPU: the restatement is pinned bit-exactly against the
reference. GPU: detrended values, chosen combination, kh, Kh series and every settings field the reference writes back
must be IDENTICAL to the reference's (the decisions change the whole raster, so there is no tolerance here)."""
import numpy as np
import pytest

from praga_b200 import _abi as A
from praga_b200 import synthetic as syn
from test_oracle_port import clone, describe_diff, port, settings_equal  # noqa: F401  (port: fixture)


def scenario(case, n=150, seed=19):
    """(dem, urban, stations, settings, variable) for a named case."""
    dem = syn.make_dem(90, 110, seed=5, cell_size=400.0)
    urban = syn.make_urban(90, 110, cell_size=400.0)
    var = A.dailyAirTemperatureMin
    codes = 0.12 if "codes" in case else 0.0
    st = syn.make_stations(dem, n, proxies=(dem, urban), seed=seed, supplemental_frac=codes)
    if "inversion" in case:
        st.value[:] = np.where(st.z < 700, st.value - 0.012 * (700 - st.z), st.value).astype(np.float32)
    if "warm_valley" in case:      # positive lapse rate below 500 m, strong negative above
        st.value[:] = (12 + np.where(st.z < 500, 0.004 * st.z, 2.0 - 0.009 * (st.z - 500)) +
                       np.random.default_rng(seed).normal(0, 0.3, n)).astype(np.float32)
    s = syn.settings_classic_detrending(2)
    if codes:
        s.useLapseRateCode = 1
    if "noinv" in case:
        s.useThermalInversion = 0
        s.selectedUseThermalInversion = s.currentUseThermalInversion = 0
    if "optimal" in case:
        s.useBestDetrending = 1
    if "td" in case:
        s.useTD = 1
        s.topoDist_maxKh = 64 if "kh64" in case else 128
    if "shepard" in case:       # the demo's default estimator (parameters.ini: algorithm=shepard_modified)
        s.method = A.shepard_modified if "modified" in case else A.shepard
        s.pointsBoundingBoxArea = float((np.float32(st.x.max()) - np.float32(st.x.min())) * (np.float32(st.y.max()) - np.float32(st.y.min())))
    if "humidity" in case:
        st = syn.make_stations(dem, n, variable="humidity", seed=seed)
        s = A.default_settings()
        s.useTD = 1
        var = A.airRelHumidity
    if "pressure" in case:
        var = A.atmPressure
        st.value[:] = (1013.0 - 0.11 * st.z + np.random.default_rng(seed).normal(0, 1.0, n)).astype(np.float32)
    if "multiple" in case:
        s = syn.settings_multiple_detrending()
        s.useTD = 1
        s.topoDist_maxKh = 32
        syn.set_points_range(s, st)
        var = A.airTemperature
    return dem, urban, st, s, var


CPU_CASES = ["classic_inversion", "classic_warm_valley", "classic_noinv", "classic_codes", "classic_pressure",
             "optimal", "optimal_inversion", "optimal_codes_warm_valley", "optimal_noinv",
             "td_humidity", "classic_td_kh64", "optimal_td_kh64", "multiple_td",
             "optimal_shepard", "optimal_modified_shepard_codes_warm_valley", "optimal_modified_shepard_inversion",
             # Shepard estimators behind the kh search (computeDistances adds the topographic distance before any estimator)
             "classic_td_kh64_shepard", "optimal_td_kh64_modified_shepard_inversion"]


def run_oracle(orc, case):
    dem, urban, st, s, var = scenario(case)
    raw = st.value.copy()
    s = clone(s)
    ok, keep, series = orc.pre_interpolation_ex(st, s, var, current_value=raw, dem=dem if "td" in case else None)
    return ok, keep, series, st, s


@pytest.mark.parametrize("case", CPU_CASES)
def test_port_reproduces_reference_station_space(ref, port, case):  # noqa: F811
    a = run_oracle(ref, case)
    b = run_oracle(port, case)
    assert a[0] == b[0]
    assert np.array_equal(a[1], b[1])
    assert a[2] == b[2], f"Kh series differ: {a[2][:4]} vs {b[2][:4]}"
    assert np.array_equal(a[3].value, b[3].value), f"max diff {np.abs(a[3].value - b[3].value).max()}"
    assert settings_equal(a[4], b[4]), describe_diff(a[4], b[4])


def test_optimal_detrending_chooses_among_combinations(ref):
    """sanity of the scenarios: the reference must not always end on the same combination / kh."""
    seen, khs = set(), set()
    for case in ("optimal", "optimal_inversion", "optimal_codes_warm_valley", "optimal_noinv", "optimal_td_kh64"):
        _, _, _, _, s = run_oracle(ref, case)
        seen.add((tuple(s.currentActive[:2]), s.currentUseThermalInversion))
        khs.add(s.topoDist_Kh)
    assert len(seen) >= 2, seen


@pytest.mark.gpu
@pytest.mark.parametrize("case", CPU_CASES)
def test_gpu_station_space_identical_to_reference(engine, ref, case):
    dem, urban, st, s, var = scenario(case)
    raw = st.value.copy()
    st_r, st_g, s_r, s_g = st.copy(), st.copy(), clone(s), clone(s)
    ok, keep_r, series_r = ref.pre_interpolation_ex(st_r, s_r, var, current_value=raw, dem=dem if "td" in case else None)
    assert ok
    dem_id = engine.upload(dem) if "td" in case else -1
    try:
        keep_g, series_g = engine.pre_interpolation_ex(st_g, s_g, var, current_value=raw, dem_id=dem_id)
    finally:
        if dem_id >= 0:
            engine.free(dem_id)
    assert np.array_equal(keep_r, keep_g)
    assert np.array_equal(st_r.value[keep_r], st_g.value[keep_g]), np.abs(st_r.value[keep_r] - st_g.value[keep_g]).max()
    assert series_r == series_g, f"Kh series: {series_r[:3]} vs {series_g[:3]}"
    assert settings_equal(s_r, s_g), describe_diff(s_r, s_g)


@pytest.mark.gpu
def test_gpu_optimal_detrending_many_seeds(engine, ref):
    """the argmin over combinations is a knife edge: 12 seeded station sets, every decision must match."""
    for seed in range(12):
        dem, urban, st, s, var = scenario("optimal_inversion" if seed % 2 else "optimal_codes_warm_valley", n=90 + 7 * seed,
                                          seed=100 + seed)
        raw = st.value.copy()
        st_r, st_g, s_r, s_g = st.copy(), st.copy(), clone(s), clone(s)
        ref.pre_interpolation_ex(st_r, s_r, var, current_value=raw)
        engine.pre_interpolation_ex(st_g, s_g, var, current_value=raw)
        assert np.array_equal(st_r.value, st_g.value), seed
        assert settings_equal(s_r, s_g), f"seed {seed}: " + describe_diff(s_r, s_g)
