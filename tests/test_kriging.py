"""GPU parity of ordinary kriging (agrolib/interpolation/kriging.cpp:97-295) against the reference's own code:
estimate (krigingResult) and RMSE (krigingRMSE) within 1e-4, through the FP64 direct path and through the
tensor-core (tcgen05) variance path."""
import numpy as np
import pytest

from praga_b200 import _abi as A
from praga_b200 import synthetic as syn

TOL = 1e-4
pytestmark = pytest.mark.gpu


def scenario(n, m, seed=0, box=3.0e5):
    rng = np.random.default_rng(seed)
    pos = rng.random((n, 2)) * box + np.array([5.0e5, 4.8e6])
    val = rng.normal(12.0, 3.0, n)
    xy = rng.random((m, 2)) * box + np.array([5.0e5, 4.8e6])
    return pos, val, xy


@pytest.mark.parametrize("mode", [A.KRIGING_SPHERICAL, A.KRIGING_EXPONENTIAL, A.KRIGING_GAUSSIAN, A.KRIGING_LINEAR])
def test_kriging_points_fp64_direct(engine, ref, mode):
    pos, val, xy = scenario(150, 333, seed=mode)
    ok, want, want_rm, _, _ = ref.kriging_points(pos, val, mode, 8.0e4, 0.1, 4.0, 2.0e-5, xy)
    assert ok
    kid = engine.kriging_setup(pos, val, mode, 8.0e4, 0.1, 4.0, 2.0e-5, fp64_direct=True)
    assert not engine.kriging_uses_tensor_path(kid)
    got, got_rm = engine.kriging_points(kid, xy)
    engine.kriging_free(kid)
    assert np.abs(got - want).max() <= 1e-7
    assert np.abs(got_rm - want_rm).max() <= 1e-7


@pytest.mark.parametrize("mode,n", [(A.KRIGING_SPHERICAL, 150), (A.KRIGING_SPHERICAL, 700), (A.KRIGING_EXPONENTIAL, 300),
                                    (A.KRIGING_GAUSSIAN, 200)])
def test_kriging_points_tensor_path(engine, ref, mode, n):
    """n = 700 spans two accumulator passes (512 TMEM columns each) and ragged tiles."""
    pos, val, xy = scenario(n, 1000, seed=10 + n)
    nug = 0.1 if mode != A.KRIGING_GAUSSIAN else 0.4
    ok, want, want_rm, _, _ = ref.kriging_points(pos, val, mode, 9.0e4, nug, 4.0, 0.0, xy)
    assert ok
    kid = engine.kriging_setup(pos, val, mode, 9.0e4, nug, 4.0, 0.0)
    # the Gaussian model is ill conditioned (pure Gaussian correlation matrix): the engine must notice and take the
    # exact FP64 path, whose elimination order is the reference's
    assert engine.kriging_uses_tensor_path(kid) == (mode != A.KRIGING_GAUSSIAN)
    got, got_rm = engine.kriging_points(kid, xy)
    engine.kriging_free(kid)
    e1, e2 = np.abs(got - want).max(), np.abs(got_rm - want_rm).max()
    print(f"mode {mode} n {n}: max |estimate - ref| = {e1:.2e}, max |rmse - ref| = {e2:.2e}")
    assert e1 <= TOL and e2 <= TOL


def test_kriging_c4_shape_tensor_path(engine, ref):
    """BASELINE config 4 itself: n = 2000 stations in the 1000 km box of the 10k x 10k @100 m DEM, spherical variogram with
    range 2e5 m, nugget 0.1, sill 4 (bench.py KrigingWorkload): 8 accumulator passes of 256 stations, ||L^-1|| at its
    largest. 320 targets (krigingSetWeight is O(n^2) per target on the CPU: ~3 ms each) against the reference's own
    krigingSetWeight / krigingResult / krigingRMSE (kriging.cpp:211-295)."""
    n, m = 2000, 320
    rng = np.random.default_rng(1)                     # the generator state of bench.py KrigingWorkload
    pos = np.stack([syn.LLX + rng.random(n) * 1.0e6, syn.LLY + rng.random(n) * 1.0e6], axis=1)
    val = rng.normal(12.0, 3.0, n)
    rq = np.random.default_rng(77)
    xy = np.stack([syn.LLX + rq.random(m) * 1.0e6, syn.LLY + rq.random(m) * 1.0e6], axis=1)
    xy[:8] = pos[:8] + 30.0                             # targets next to a station (largest weights, strongest cancellation)
    ok, want, want_rm, _, _ = ref.kriging_points(pos, val, A.KRIGING_SPHERICAL, 2.0e5, 0.1, 4.0, 0.0, xy)
    assert ok
    kid = engine.kriging_setup(pos, val, A.KRIGING_SPHERICAL, 2.0e5, 0.1, 4.0, 0.0)
    assert engine.kriging_uses_tensor_path(kid) == 1
    got, got_rm = engine.kriging_points(kid, xy)
    engine.kriging_free(kid)
    e1, e2 = np.abs(got - want).max(), np.abs(got_rm - want_rm).max()
    print(f"C4 shape n {n}: max |estimate - ref| = {e1:.2e}, max |rmse - ref| = {e2:.2e}")
    assert e1 <= TOL and e2 <= TOL


def test_kriging_linear_falls_back_to_fp64(engine, ref):
    pos, val, xy = scenario(120, 200, seed=3)
    kid = engine.kriging_setup(pos, val, A.KRIGING_LINEAR, 8.0e4, 0.2, 4.0, 1.0e-5)
    assert not engine.kriging_uses_tensor_path(kid)
    ok, want, want_rm, _, _ = ref.kriging_points(pos, val, A.KRIGING_LINEAR, 8.0e4, 0.2, 4.0, 1.0e-5, xy)
    got, got_rm = engine.kriging_points(kid, xy)
    engine.kriging_free(kid)
    assert np.abs(got - want).max() <= 1e-6 and np.abs(got_rm - want_rm).max() <= 1e-6


def test_kriging_zero_nugget_is_rejected_like_the_reference_nan(engine):
    import praga_b200 as pb
    pos, val, xy = scenario(40, 10, seed=5)
    with pytest.raises(pb.PragaB200Error):
        engine.kriging_setup(pos, val, A.KRIGING_LINEAR, 8.0e4, 0.0, 4.0, 1.0e-5)


def test_kriging_raster_matches_points(engine, ref):
    """raster form: valid cells at their centres (gis::getUtmXYFromRowCol), flag elsewhere."""
    dem = syn.make_dem(90, 110, cell_size=2000.0)
    rng = np.random.default_rng(7)
    n = 260
    pos = np.stack([dem.llx + rng.random(n) * 110 * 2000.0, dem.lly + rng.random(n) * 90 * 2000.0], axis=1)
    val = rng.normal(10.0, 2.0, n)
    kid = engine.kriging_setup(pos, val, A.KRIGING_SPHERICAL, 6.0e4, 0.1, 3.0, 0.0)
    did = engine.upload(dem)
    est, rm, nv = engine.kriging_raster(kid, did, dem.shape)
    rows, cols = np.nonzero(dem.data != np.float32(dem.flag))
    assert nv == rows.size
    xy = np.stack([dem.llx + dem.cell_size * (cols + 0.5), dem.lly + dem.cell_size * (90 - rows - 0.5)], axis=1)
    ok, want, want_rm, _, _ = ref.kriging_points(pos, val, A.KRIGING_SPHERICAL, 6.0e4, 0.1, 3.0, 0.0, xy)
    assert np.all(est[dem.data == np.float32(dem.flag)] == np.float32(dem.flag))
    assert np.abs(est[rows, cols] - want).max() <= TOL
    assert np.abs(rm[rows, cols] - want_rm).max() <= TOL
    engine.kriging_free(kid)
    engine.free(did)


def test_kriging_c4_raster_tile_with_skipped_stages(engine, ref):
    """The C4 system on raster tiles, where the variance kernel skips K stages: a pair of CTAs covers 256 neighbouring cells
    and multiplies only the stages (32 stations, k-d ordered on the host) that can be inside the variogram range of one of
    them. A 40 x 40 window of 100 m cells in the middle of the 1000 km box and one in its corner, every cell against the
    reference's krigingSetWeight / krigingResult / krigingRMSE; the counter of issued MMAs proves that stages were skipped."""
    n = 2000
    rng = np.random.default_rng(1)
    pos = np.stack([syn.LLX + rng.random(n) * 1.0e6, syn.LLY + rng.random(n) * 1.0e6], axis=1)
    val = rng.normal(12.0, 3.0, n)
    kid = engine.kriging_setup(pos, val, A.KRIGING_SPHERICAL, 2.0e5, 0.1, 4.0, 0.0)
    assert engine.kriging_uses_tensor_path(kid) == 1
    for (ox, oy) in ((4.7e5, 5.1e5), (0.0, 0.0)):
        data = np.full((40, 40), 100.0, dtype=np.float32)
        data[3, 5:9] = -9999.0
        dem = A.Raster(data, syn.LLX + ox, syn.LLY + oy, 100.0, -9999.0)
        did = engine.upload(dem)
        engine.kriging_mma_count(kid)
        est, rm, nv = engine.kriging_raster(kid, did, dem.shape)
        mma = engine.kriging_mma_count(kid)
        rows, cols = np.nonzero(data != np.float32(-9999.0))
        assert nv == rows.size
        xy = np.stack([dem.llx + 100.0 * (cols + 0.5), dem.lly + 100.0 * (40 - rows - 0.5)], axis=1)
        sel = np.random.default_rng(5).choice(rows.size, 160, replace=False)
        ok, want, want_rm, _, _ = ref.kriging_points(pos, val, A.KRIGING_SPHERICAL, 2.0e5, 0.1, 4.0, 0.0, xy[sel])
        assert ok
        e1 = np.abs(est[rows[sel], cols[sel]] - want).max()
        e2 = np.abs(rm[rows[sel], cols[sel]] - want_rm).max()
        dense = 1728 * ((nv + 255) // 256)              # MMAs of the unskipped lower triangle (n padded to 2048) per 256 cells
        print(f"window at ({ox:.0f}, {oy:.0f}): estimate {e1:.2e} rmse {e2:.2e}, MMAs {mma} of {dense}")
        assert e1 <= TOL and e2 <= TOL
        assert 0 < mma < 0.5 * dense
        engine.free(did)
    engine.kriging_free(kid)


def test_kriging_targets_out_of_every_range(engine, ref):
    """targets farther than the range from every station: all weights come from the unbiasedness constraint alone; the
    variance kernel skips every stage (y = 0) and must still return the reference's numbers."""
    pos, val, _ = scenario(300, 1, seed=11)
    xy = np.stack([np.full(40, 5.0e5 + 9.0e5) + np.arange(40) * 50.0, np.full(40, 4.8e6 + 1.0e5)], axis=1)
    ok, want, want_rm, _, _ = ref.kriging_points(pos, val, A.KRIGING_SPHERICAL, 8.0e4, 0.2, 4.0, 0.0, xy)
    assert ok
    kid = engine.kriging_setup(pos, val, A.KRIGING_SPHERICAL, 8.0e4, 0.2, 4.0, 0.0)
    assert engine.kriging_uses_tensor_path(kid) == 1
    engine.kriging_mma_count(kid)
    got, got_rm = engine.kriging_points(kid, xy)
    assert engine.kriging_mma_count(kid) == 0
    engine.kriging_free(kid)
    assert np.abs(got - want).max() <= TOL and np.abs(got_rm - want_rm).max() <= TOL
