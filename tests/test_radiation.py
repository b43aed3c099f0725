"""Solar-radiation maps: pb_compute_radiation_dem against the reference's own radiation::computeRadiationDEM
(agrolib/solarRadiation/solarRadiation.cpp:1045-1069, compiled into oracle/_ref) on the same DEM, static maps (lat / lon / slope /
aspect as Crit3DRadiationMaps derives them), transmissivity map and settings. Bar: 1e-4 W m-2 (degrees for the sun elevation) on
every cell, flags equal. The kernel restates glibc's acosf bit for bit; the double-precision sin / cos / exp / pow are CUDA's."""
import numpy as np
import pytest

from praga_b200 import _abi as A
from praga_b200 import synthetic as syn

TOL = 1e-4
NAMES = ("sunElevation", "beam", "diffuse", "reflected", "global")


def transmissivity_map(dem, seed=3):
    rng = np.random.default_rng(seed)
    r, c = np.indices(dem.shape)
    t = (0.55 + 0.2 * np.sin(r / 31.0) * np.cos(c / 27.0) + rng.normal(0, 0.02, dem.shape)).astype(np.float32)
    t = np.clip(t, 0.05, 0.8).astype(np.float32)
    t[dem.data == np.float32(dem.flag)] = np.float32(dem.flag)
    return A.Raster(t, dem.llx, dem.lly, dem.cell_size, dem.flag)


CASES = [
    # (label, month, day, seconds of day (local), realSky, algorithm (1 Linke, 0 total transmissivity), shadowing, tilt mode)
    ("summer_noon_linke", 6, 21, 12 * 3600, 1, 1, 1, 2),
    ("winter_morning_linke", 1, 10, 8 * 3600 + 1800, 1, 1, 1, 2),
    ("evening_total_transmissivity", 9, 3, 18 * 3600, 1, 0, 1, 2),
    ("night", 3, 15, 2 * 3600, 1, 1, 1, 2),
    # (shadowing off is not compared: the reference then reads an uninitialised TsunPosition::shadow, solarRadiation.cpp:744-750
    # and :801; the engine takes it as false)
    ("clear_sky_fixed_tilt", 4, 1, 10 * 3600 + 600, 0, 1, 1, 1),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_radiation_maps_match_the_reference(engine, ref, case):
    label, month, day, sec, real_sky, algo, shadowing, tilt_mode = case
    dem = syn.make_dem(150, 170, seed=13, cell_size=100.0)        # steep synthetic relief (slopes up to 28 deg): shadows at low sun
    dem.data[dem.data != np.float32(dem.flag)] *= np.float32(2.0)
    tr = transmissivity_map(dem)
    s = A.default_radiation_settings(2023, month, day, sec, time_zone=1)
    s.realSky, s.realSkyAlgorithm, s.shadowing, s.tiltMode = real_sky, algo, shadowing, tilt_mode
    s.tilt, s.aspect = 25.0, 160.0
    want = ref.compute_radiation_dem(s, dem, tr, utm_zone=32, start_latitude=44.5)
    assert want["ok"]
    s.demMaximum = want["dem_maximum"]
    ids = [engine.upload(dem)]
    for k in ("lat", "lon", "slope", "aspect"):
        ids.append(engine.upload(A.Raster(want[k], dem.llx, dem.lly, dem.cell_size, dem.flag)))
    ids.append(engine.upload(tr))
    try:
        got = engine.compute_radiation_dem(s, *ids, dem.shape)
        flag = np.float32(dem.flag)
        worst = {}
        for k, nme in enumerate(NAMES):
            g, mn, mx = got[nme]
            w = want[nme]
            assert np.array_equal(g == flag, w == flag), f"{label}: {nme} flags differ"
            m = w != flag
            d = np.abs(g[m].astype(np.float64) - w[m].astype(np.float64))
            worst[nme] = (float(d.max()) if d.size else 0.0, float(np.mean(d == 0)) if d.size else 1.0)
            assert d.size == 0 or d.max() <= TOL, f"{label}: {nme} max |gpu - ref| = {d.max():.3e}"
            wmn, wmx = want["minmax"][2 * k], want["minmax"][2 * k + 1]
            assert abs(mn - wmn) <= TOL and abs(mx - wmx) <= TOL, (label, nme, mn, wmn, mx, wmx)
        print(label, {k: (f"{v[0]:.1e}", f"{100 * v[1]:.2f}% identical") for k, v in worst.items()})
        if label == "night":
            assert got["global"][2] == 0.0
        if label == "winter_morning_linke":
            assert (want["beam"][want["beam"] != flag] == 0).mean() > 0.02      # some cells are in the shade of the relief
    finally:
        for i in ids:
            engine.free(i)


@pytest.mark.gpu
def test_radiation_from_a_gridded_transmissivity_map_stays_on_the_device(engine, ref):
    """Project::interpolateDemRadiation (project.cpp:3382-3437): the stations' atmTransmissivity is gridded onto the DEM, the map
    feeds computeRadiationDEM; here the gridded map never leaves the device (pb_raster_from_last_output)."""
    dem = syn.make_dem(120, 130, seed=4, cell_size=500.0)
    st = syn.make_stations(dem, 60, variable="humidity", seed=8)
    st.value[:] = np.clip(st.value / 100.0, 0.1, 0.8).astype(np.float32)
    sett = A.default_settings()
    okc, gc, *_ = ref.interpolation_raster(st, sett, A.atmTransmissivity, dem, ())
    did = engine.upload(dem)
    engine.interpolation_raster_resident(st, sett, A.atmTransmissivity, did, (), device_only=True)
    tid = engine.raster_from_last_output(did)
    s = A.default_radiation_settings(2023, 7, 1, 11 * 3600, time_zone=1)
    want = ref.compute_radiation_dem(s, dem, A.Raster(gc, dem.llx, dem.lly, dem.cell_size, dem.flag))
    s.demMaximum = want["dem_maximum"]
    ids = [engine.upload(A.Raster(want[k], dem.llx, dem.lly, dem.cell_size, dem.flag)) for k in ("lat", "lon", "slope", "aspect")]
    try:
        got = engine.compute_radiation_dem(s, did, *ids, tid, dem.shape, keep=True)
        g, w = got["global"][0], want["global"]
        flag = np.float32(dem.flag)
        assert np.array_equal(g == flag, w == flag)
        m = w != flag
        # the gridded transmissivity itself differs by <= 1e-5 between the two (K1 parity), radiation ~ 1e3 x transmissivity
        assert np.abs(g[m].astype(np.float64) - w[m]).max() <= 2e-2
        assert len(got["ids"]) == 5 and all(i >= 0 for i in got["ids"])
        for i in got["ids"]:
            engine.free(i)
    finally:
        for i in ids + [tid, did]:
            engine.free(i)
