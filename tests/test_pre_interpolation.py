"""GPU parity of preInterpolation (interpolation.cpp:2444-2499), multiple-detrending branch, against the reference:
detrended station values BIT-EXACT, surviving points and every settings field the reference writes back identical."""
import ctypes

import numpy as np
import pytest

from praga_b200 import _abi as A
from praga_b200 import synthetic as syn
from test_oracle_port import clone, describe_diff, settings_equal

pytestmark = pytest.mark.gpu


def scenario(n=220, seed=17, three=False, codes=0.0):
    dem = syn.make_dem(100, 120, seed=5)
    urban = syn.make_urban(100, 120)
    st = syn.make_stations(dem, n, proxies=(dem, urban), seed=seed, supplemental_frac=codes)
    s = syn.settings_multiple_detrending()
    if codes:
        s.useLapseRateCode = 1
    if three:
        el = s.proxy[0]
        el.fittingFunction = A.piecewiseThree
        el.nFitParams = 5
        for j, (lo, hi, fg) in enumerate(zip((0, -30, 100, -0.01, -0.015), (1500, 50, 1000, 0.01, -0.002), (0, 1, 1, 1, 1))):
            el.fitMin[j], el.fitMax[j], el.fitFirstGuess[j] = lo, hi, fg
        s.proxy[0] = el
    syn.set_points_range(s, st)
    return dem, urban, st, s


@pytest.mark.parametrize("case", ["two_piece", "three_piece", "lapse_codes", "missing_proxies"])
def test_pre_interpolation_single_step_bit_exact(engine, ref, case):
    dem, urban, st, s = scenario(three=case == "three_piece", codes=0.12 if case == "lapse_codes" else 0.0)
    if case == "missing_proxies":
        st.proxy_values[::9, 1] = A.NODATA      # incomplete points are erased from myPoints (interpolation.cpp:2040-2069)
        st.proxy_values[5::13, 0] = A.NODATA
    st_r, st_g, s_r, s_g = st.copy(), st.copy(), clone(s), clone(s)
    ok, keep_r = ref.pre_interpolation(st_r, s_r, A.airTemperature, (dem, urban))
    assert ok
    keep_g = engine.pre_interpolation(st_g, s_g, A.airTemperature)
    assert np.array_equal(keep_r, keep_g)
    assert np.array_equal(st_r.value[keep_r], st_g.value[keep_g]), np.abs(st_r.value[keep_r] - st_g.value[keep_g]).max()
    assert settings_equal(s_r, s_g), describe_diff(s_r, s_g)


@pytest.mark.parametrize("T", [24, 70])
def test_pre_interpolation_batch_matches_per_step_reference(engine, ref, T):
    """T 'hours' sharing one station set: every step bit-identical to the reference run on that step alone. Up to 64 steps run
    on the warp-per-descent kernel (one CTA per step), more on the lane-per-descent batch kernel: both are covered."""
    dem, urban, st, s = scenario(n=300, seed=3)
    rng = np.random.default_rng(11)
    values = np.stack([st.value + np.float32(3.0 * np.sin(2 * np.pi * t / 24)) + rng.normal(0, 0.4, st.n).astype(np.float32)
                       for t in range(T)]).astype(np.float32)
    det, keep, fitted = engine.pre_interpolation_batch(st, values, s, A.airTemperature)
    for t in range(T):
        st_t = st.copy()
        st_t.value[:] = values[t]
        s_t = clone(s)
        syn.set_points_range(s_t, st_t)
        ok, keep_r = ref.pre_interpolation(st_t, s_t, A.airTemperature, (dem, urban))
        assert ok and np.array_equal(keep_r, keep[t])
        assert np.array_equal(st_t.value[keep_r], det[t][keep[t]])
        assert settings_equal(s_t, fitted[t]), f"step {t}: " + describe_diff(s_t, fitted[t])


def test_precipitation_all_zero_flag(engine, ref):
    dem = syn.make_dem(60, 60)
    st = syn.make_stations(dem, 70, variable="precipitation", seed=8)
    st.value[:] = 0.1
    s = A.default_settings()
    engine.pre_interpolation(st, s, A.precipitation)
    assert s.precipitationAllZero == 1
