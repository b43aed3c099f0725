"""spatialQualityControl (spatialControl.cpp:331-427): the check every checkAndPassDataToInterpolation runs before a
time step is gridded. Two preInterpolation fits, one exact leave-one-out, neighbourhoodVariability at every point and
the per-variable thresholds of getSpatialThresholdVar. The outcome is a set of decisions (which stations are dropped),
so GPU and reference must agree EXACTLY: the wrong_spatial mask and every settings field the reference leaves behind.
CPU: the restatement is pinned against the reference on the same cases."""
import numpy as np
import pytest

from praga_b200 import _abi as A
from praga_b200 import synthetic as syn
from test_oracle_port import clone, describe_diff, port, settings_equal  # noqa: F401  (port: fixture)
from test_station_space import scenario

CASES = ["classic_inversion", "classic_codes", "optimal", "optimal_codes_warm_valley", "classic_pressure",
         "td_humidity", "classic_td_kh64", "multiple_td", "humidity_plain", "wind", "radiation", "transmissivity",
         "optimal_modified_shepard", "classic_shepard_codes", "classic_td_kh64_modified_shepard"]


def sq_scenario(case, n=150, seed=19):
    """a station-space scenario with gross errors planted at a few stations (some obvious, some borderline)."""
    if case in ("humidity_plain", "wind", "radiation", "transmissivity"):
        dem = syn.make_dem(90, 110, seed=5, cell_size=400.0)
        urban = syn.make_urban(90, 110, cell_size=400.0)
        st = syn.make_stations(dem, n, variable="humidity", seed=seed)
        s = A.default_settings()
        var = {"humidity_plain": A.airRelHumidity, "wind": A.windScalarIntensity, "radiation": A.globalIrradiance,
               "transmissivity": A.atmTransmissivity}[case]
        scale = {"humidity_plain": 1.0, "wind": 0.06, "radiation": 8.0, "transmissivity": 0.008}[case]
        st.value[:] = (st.value * scale).astype(np.float32)
    else:
        dem, urban, st, s, var = scenario(case, n=n, seed=seed)
        scale = 1.0
    rng = np.random.default_rng(seed + 1)
    bad = rng.choice(st.n, size=9, replace=False)
    spread = float(np.std(st.value))
    bump = np.array([6, -6, 4, -3, 2.5, 10, -10, 1.5, 3.5]) * max(spread, 1e-3)
    if var in (A.airRelHumidity,):
        bump = np.array([60, -60, 40, -35, 30, 80, -80, 25, 45], dtype=np.float64)
    st.value[bad] = (st.value[bad] + bump).astype(np.float32)
    return dem, urban, st, s, var, bad


def run_oracle(orc, case):
    dem, urban, st, s, var, bad = sq_scenario(case)
    s = clone(s)
    ok, wrong = orc.spatial_quality_control(st, s, var, dem=dem if "td" in case else None)
    return ok, wrong, s, bad


@pytest.mark.parametrize("case", CASES)
def test_port_reproduces_reference_spatial_quality(ref, port, case):  # noqa: F811
    a = run_oracle(ref, case)
    b = run_oracle(port, case)
    assert a[0] and b[0]
    assert np.array_equal(a[1], b[1]), (np.flatnonzero(a[1]), np.flatnonzero(b[1]))
    assert settings_equal(a[2], b[2]), describe_diff(a[2], b[2])


def test_spatial_quality_flags_planted_errors(ref):
    """sanity of the scenarios: the reference flags something, but not everything, in most cases."""
    flagged = {case: int(run_oracle(ref, case)[1].sum()) for case in CASES}
    assert sum(1 for v in flagged.values() if v > 0) >= len(CASES) // 2, flagged
    assert all(v < 40 for v in flagged.values()), flagged


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_spatial_quality_identical_to_reference(engine, ref, case):
    dem, urban, st, s, var, bad = sq_scenario(case)
    s_r, s_g = clone(s), clone(s)
    ok, wrong_r = ref.spatial_quality_control(st.copy(), s_r, var, dem=dem if "td" in case else None)
    assert ok
    dem_id = engine.upload(dem) if "td" in case else -1
    try:
        wrong_g = engine.spatial_quality_control(st.copy(), s_g, var, dem_id=dem_id)
    finally:
        if dem_id >= 0:
            engine.free(dem_id)
    assert np.array_equal(wrong_r, wrong_g), (np.flatnonzero(wrong_r), np.flatnonzero(wrong_g))
    assert settings_equal(s_r, s_g), describe_diff(s_r, s_g)


@pytest.mark.gpu
def test_gpu_spatial_quality_many_seeds(engine, ref):
    for seed in range(10):
        dem, urban, st, s, var, bad = sq_scenario("optimal" if seed % 2 else "classic_codes", n=80 + 11 * seed, seed=300 + seed)
        s_r, s_g = clone(s), clone(s)
        _, wrong_r = ref.spatial_quality_control(st.copy(), s_r, var)
        wrong_g = engine.spatial_quality_control(st.copy(), s_g, var)
        assert np.array_equal(wrong_r, wrong_g), seed
        assert settings_equal(s_r, s_g), f"seed {seed}: " + describe_diff(s_r, s_g)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["classic_td_kh64", "optimal_codes_warm_valley", "td_humidity", "optimal_modified_shepard_codes",
                                  "classic_shepard_inversion", "classic_td_kh64_shepard", "optimal_td_kh64_modified_shepard"])
def test_gpu_exact_residuals_identical(engine, ref, port, case):  # noqa: F811
    """pb_compute_residuals_exact against computeResiduals with the fitted settings (topographic distance included)."""
    dem, urban, st, s, var = scenario(case)
    raw = st.value.copy()
    st_f, s_f = st.copy(), clone(s)
    ok, keep, _ = ref.pre_interpolation_ex(st_f, s_f, var, current_value=raw, dem=dem if "td" in case else None)
    assert ok
    pts = st_f.subset(keep) if hasattr(st_f, "subset") else st_f
    meteo = st.copy()
    dem_id = engine.upload(dem) if "td" in case else -1
    try:
        got = engine.compute_residuals_exact(meteo, pts, s_f, var, observed=raw, dem_id=dem_id)
    finally:
        if dem_id >= 0:
            engine.free(dem_id)
    # the reference's computeResiduals through interpolate() at every station with its own proxies
    pv = st.proxy_values.astype(np.float64)
    if "td" in case:
        est = ref.interpolate_points_td(pts, s_f, var, st.x.astype(np.float32), st.y.astype(np.float32),
                                        st.z.astype(np.float32), pv, dem)
    else:
        est = ref.interpolate_points(pts, s_f, var, st.x.astype(np.float32), st.y.astype(np.float32),
                                     st.z.astype(np.float32), pv)
    want = np.where(est == np.float32(-9999), np.float32(-9999), raw - est).astype(np.float32)
    assert np.array_equal(got, want), np.abs(got - want).max()
