"""Python binding of libpraga_b200.so (ctypes over the C ABI of include/praga_b200.h).

This is plumbing for tests and bench.py; the product is the shared library. There is no CPU fallback: if the
library is missing or no B200 is visible, construction raises.
"""
import ctypes as C
import os

import numpy as np

from . import _abi as A

_LIB = None


class PragaB200Error(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{A.STATUS_NAMES[status] if 0 <= status < len(A.STATUS_NAMES) else status}: {message}")
        self.status = status


def lib_path():
    """the in-tree library; PRAGA_B200_LIB selects another BUILD of the same CUDA library (instrumented debug builds)."""
    return os.environ.get("PRAGA_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpraga_b200.so")


def load_library():
    """dlopen libpraga_b200.so (built by praga_b200/build.py). Fails loudly when it is absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise PragaB200Error(A.PB_ERR_CUDA, f"{path} is missing: run `python -m praga_b200.build` "
                                            "(the CUDA extension is mandatory, there is no CPU fallback)")
    lib = C.CDLL(path)
    P = C.POINTER
    lib.pb_version.restype = C.c_char_p
    lib.pb_abi_sizeof.restype = C.c_size_t
    lib.pb_abi_sizeof.argtypes = [C.c_char_p]
    lib.pb_device_count.restype = C.c_int
    lib.pb_create.argtypes = [C.c_int, P(C.c_void_p)]
    lib.pb_destroy.argtypes = [C.c_void_p]
    lib.pb_destroy.restype = None
    lib.pb_last_error.argtypes = [C.c_void_p]
    lib.pb_last_error.restype = C.c_char_p
    lib.pb_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.pb_last_timing.argtypes = [C.c_void_p, P(A.pb_timing)]
    lib.pb_raster_upload.argtypes = [C.c_void_p, P(A.pb_raster), P(C.c_int32)]
    lib.pb_raster_free.argtypes = [C.c_void_p, C.c_int32]
    lib.pb_raster_valid_cells.argtypes = [C.c_void_p, C.c_int32, P(C.c_int64)]
    raster_args = [C.c_void_p, P(A.pb_stations), P(A.pb_settings), C.c_int]
    lib.pb_interpolation_raster.argtypes = raster_args + [P(A.pb_raster), P(A.pb_raster), C.c_int, C.c_int, C.c_int,
                                                          P(A.pb_raster_out)]
    lib.pb_local_detrending_raster.argtypes = lib.pb_interpolation_raster.argtypes
    for name in ("pb_interpolation_raster_resident", "pb_interpolation_raster_device",
                 "pb_local_detrending_raster_resident", "pb_local_detrending_raster_device"):
        getattr(lib, name).argtypes = raster_args + [C.c_int32, P(C.c_int32), C.c_int, C.c_int, C.c_int,
                                                     P(A.pb_raster_out)]
    lib.pb_interpolate_points.argtypes = raster_args + [P(C.c_float), P(C.c_float), P(C.c_double), C.c_int, C.c_int,
                                                        P(C.c_float)]
    lib.pb_pre_interpolation.argtypes = [C.c_void_p, P(A.pb_stations), P(A.pb_settings), C.c_int, P(C.c_uint8)]
    lib.pb_pre_interpolation_ex.argtypes = [C.c_void_p, P(A.pb_stations), P(A.pb_settings), C.c_int, P(A.pb_cv_points), C.c_int32,
                                            P(C.c_uint8), P(A.pb_kh_series)]
    lib.pb_spatial_quality_control.argtypes = [C.c_void_p, P(A.pb_stations), P(A.pb_settings), C.c_int, P(A.pb_cv_points),
                                               C.c_int32, P(C.c_uint8)]
    lib.pb_compute_residuals_exact.argtypes = [C.c_void_p, P(A.pb_stations), P(C.c_float), P(A.pb_cv_points), P(A.pb_stations),
                                               P(A.pb_settings), C.c_int, C.c_int32, C.c_int, C.c_int, P(C.c_float)]
    lib.pb_pre_interpolation_batch.argtypes = [C.c_void_p, P(A.pb_stations), P(C.c_float), C.c_int, P(A.pb_settings), C.c_int,
                                               P(C.c_double), P(C.c_float), P(C.c_uint8), P(A.pb_settings)]
    lib.pb_compute_residuals.argtypes = raster_args + [P(C.c_float), P(C.c_uint8), P(C.c_float), P(A.pb_cv_stats)]
    lib.pb_compute_residuals_local.argtypes = lib.pb_compute_residuals.argtypes
    lib.pb_local_detrending_points.argtypes = lib.pb_interpolate_points.argtypes
    lib.pb_interpolation_dem.argtypes = [C.c_void_p, P(A.pb_stations), P(A.pb_settings), P(A.pb_settings), C.c_int, C.c_int32,
                                         P(C.c_int32), C.c_int, C.c_int, C.c_int, C.c_int, P(A.pb_raster_out), P(A.pb_dem_step)]
    lib.pb_interpolation_dem_batch.argtypes = [C.c_void_p, P(A.pb_dem_batch_item), C.c_int, C.c_int32, P(C.c_int32), C.c_int, C.c_int,
                                               C.c_int, C.c_int, C.c_int]
    lib.pb_meteo_grid_step_batch.argtypes = [C.c_void_p, P(A.pb_meteo_grid_hour), C.c_int, C.c_int32, P(C.c_int32), C.c_int, C.c_int32,
                                             C.c_int]
    lib.pb_compute_radiation_dem.argtypes = [C.c_void_p, P(A.pb_radiation_settings), P(A.pb_radiation_maps), P(A.pb_radiation_out)]
    lib.pb_raster_from_last_output.argtypes = [C.c_void_p, C.c_int32, P(C.c_int32)]
    lib.pb_kriging_setup.argtypes = [C.c_void_p, P(C.c_double), P(C.c_double), C.c_int, C.c_int, C.c_double, C.c_double,
                                     C.c_double, C.c_double, C.c_int, P(C.c_int32)]
    lib.pb_kriging_free.argtypes = [C.c_void_p, C.c_int32]
    lib.pb_kriging_uses_tensor_path.argtypes = [C.c_void_p, C.c_int32]
    lib.pb_kriging_mma_count.argtypes = [C.c_void_p, C.c_int32, P(C.c_ulonglong)]
    lib.pb_kriging_points.argtypes = [C.c_void_p, C.c_int32, P(C.c_double), C.c_int, P(C.c_double), P(C.c_double)]
    lib.pb_kriging_raster.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int, C.c_int, P(A.pb_raster_out),
                                      P(A.pb_raster_out)]
    lib.pb_interpolate_points_td.argtypes = raster_args + [P(C.c_float), P(C.c_float), P(C.c_float), P(C.c_double), C.c_int,
                                                           C.c_int, C.c_int32, P(C.c_float)]
    lib.pb_raster_read_esri_header.argtypes = [C.c_void_p, C.c_char_p, P(A.pb_raster_header)]
    lib.pb_raster_read_esri.argtypes = [C.c_void_p, C.c_char_p, P(C.c_int32), P(A.pb_raster_header), P(C.c_float), P(C.c_float)]
    lib.pb_raster_write_esri.argtypes = [C.c_void_p, C.c_int32, C.c_char_p]
    lib.pb_raster_download.argtypes = [C.c_void_p, C.c_int32, P(A.pb_raster_out)]
    lib.pb_relative_humidity_map.argtypes = [C.c_void_p, C.c_int32, C.c_int32, P(A.pb_raster_out), P(C.c_int32)]
    lib.pb_fix_daily_thermal_consistency.argtypes = [C.c_void_p, C.c_int32, C.c_int32, P(C.c_int64)]
    lib.pb_topographic_distance_map.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_double, P(A.pb_raster_out),
                                                P(C.c_int32)]
    lib.pb_resample_grid.argtypes = [C.c_void_p, C.c_int32, P(A.pb_raster_header), C.c_int, C.c_float, P(A.pb_raster_out),
                                     P(C.c_int32)]
    lib.pb_glocal_detrending_fitting.argtypes = [C.c_void_p, P(A.pb_stations), P(A.pb_settings), C.c_int, P(A.pb_macro_area), C.c_int]
    lib.pb_glocal_detrending_raster.argtypes = [C.c_void_p, P(A.pb_stations), P(A.pb_settings), C.c_int, P(A.pb_macro_area), C.c_int,
                                                C.c_int32, P(C.c_int32), C.c_int, C.c_int, P(A.pb_raster_out)]
    lib.pb_glocal_detrending_raster_resident.argtypes = [C.c_void_p, P(A.pb_stations), P(A.pb_settings), C.c_int, P(A.pb_macro_area),
                                                         C.c_int, C.c_int32, C.c_int32, P(C.c_int32), C.c_int, C.c_int,
                                                         P(A.pb_raster_out)]
    lib.pb_glocal_detrending_raster_td.argtypes = [C.c_void_p, P(A.pb_stations), P(A.pb_settings), C.c_int, P(A.pb_macro_area),
                                                   C.c_int, C.c_int32, P(A.pb_glocal_meteo), C.c_int32, P(C.c_int32), C.c_int,
                                                   C.c_int, P(A.pb_raster_out), P(C.c_int32)]
    lib.pb_compute_residuals_glocal_td.argtypes = [C.c_void_p, P(A.pb_stations), P(C.c_float), P(A.pb_cv_points), P(A.pb_stations),
                                                   P(A.pb_settings), C.c_int, P(A.pb_macro_area), C.c_int, C.c_int32, C.c_int,
                                                   C.c_int, P(C.c_float)]
    lib.pb_set_topographic_distance_maps.argtypes = [C.c_void_p, P(C.c_int32), P(C.c_int32), C.c_int]
    lib.pb_macro_areas_upload.argtypes = [C.c_void_p, P(A.pb_macro_area), C.c_int, P(C.c_int32)]
    lib.pb_macro_areas_free.argtypes = [C.c_void_p, C.c_int32]
    lib.pb_host_register.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.pb_host_unregister.argtypes = [C.c_void_p, C.c_void_p]
    lib.pb_topographic_index.argtypes = [C.c_void_p, C.c_int32, P(C.c_float), C.c_int, P(A.pb_raster_out), P(C.c_int32)]
    lib.pb_raster_sample_points.argtypes = [C.c_void_p, C.c_int32, P(C.c_double), P(C.c_double), C.c_int, P(C.c_float), P(C.c_uint8)]
    lib.pb_aggregation_upload.argtypes = [C.c_void_p, P(C.c_double), P(C.c_double), P(C.c_int64), P(C.c_int32), C.c_int, P(C.c_int32)]
    lib.pb_aggregation_free.argtypes = [C.c_void_p, C.c_int32]
    lib.pb_spatial_aggregate_meteo_grid.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int, P(C.c_float)]
    lib.pb_compute_residuals_glocal.argtypes = [C.c_void_p, P(A.pb_stations), P(C.c_float), P(A.pb_cv_points), P(A.pb_stations),
                                                P(A.pb_settings), C.c_int, P(A.pb_macro_area), C.c_int, C.c_int, C.c_int,
                                                P(C.c_float)]
    lib.pb_kriging_empirical_variogram.argtypes = [P(C.c_double), P(C.c_double), C.c_int, C.c_int, C.c_double, P(C.c_float),
                                                   P(C.c_float), P(C.c_int32), P(C.c_double)]
    lib.pb_kriging_fit_variogram.argtypes = [P(C.c_float), P(C.c_float), P(C.c_int32), C.c_int, C.c_int, P(C.c_double),
                                             P(C.c_double), P(C.c_double), P(C.c_double), P(C.c_int), P(C.c_double)]
    lib.pb_measure_pipe_peaks.argtypes = [C.c_void_p, P(C.c_float), P(C.c_float)]
    lib.pb_measure_fp64_peak.argtypes = [C.c_void_p, P(C.c_float)]
    for name in ("pb_raster_header", "pb_raster", "pb_stations", "pb_proxy", "pb_settings", "pb_raster_out",
                 "pb_cv_stats", "pb_timing", "pb_cv_points", "pb_kh_series", "pb_dem_step", "pb_dem_batch_item",
                 "pb_macro_area", "pb_meteo_grid_var", "pb_meteo_grid_hour", "pb_radiation_settings", "pb_radiation_maps",
                 "pb_radiation_out"):
        got, want = lib.pb_abi_sizeof(name.encode()), C.sizeof(getattr(A, name))
        if got != want:
            raise PragaB200Error(A.PB_ERR_INVALID, f"ABI mismatch for {name}: library {got} B, binding {want} B")
    _LIB = lib
    return lib


def empirical_variogram(pos, val, n_bins=30, max_distance=0.0):
    """Host-side Matheron estimator (pb_kriging_empirical_variogram). pos: (n, 2). Returns (dist, semivar, count, max_distance)."""
    lib = load_library()
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 2)
    val = np.ascontiguousarray(val, dtype=np.float64)
    d, g, c = np.empty(n_bins, np.float32), np.empty(n_bins, np.float32), np.empty(n_bins, np.int32)
    used = C.c_double(0)
    rc = lib.pb_kriging_empirical_variogram(A.dptr(pos), A.dptr(val), val.size, n_bins, float(max_distance), A.fptr(d), A.fptr(g),
                                            A.iptr(c), C.byref(used))
    if rc != A.PB_OK:
        raise PragaB200Error(rc, "pb_kriging_empirical_variogram failed")
    return d, g, c, used.value


def fit_variogram(dist, semivar, count=None, mode=0):
    """Host-side weighted least-squares variogram fit (pb_kriging_fit_variogram). Returns a dict with mode, range, nugget, sill,
    slope, sse: the inputs of Engine.kriging_setup."""
    lib = load_library()
    d = np.ascontiguousarray(dist, dtype=np.float32)
    g = np.ascontiguousarray(semivar, dtype=np.float32)
    c = None if count is None else np.ascontiguousarray(count, dtype=np.int32)
    rng, nug, sill, slope, sse, m = C.c_double(0), C.c_double(0), C.c_double(0), C.c_double(0), C.c_double(0), C.c_int(0)
    rc = lib.pb_kriging_fit_variogram(A.fptr(d), A.fptr(g), None if c is None else A.iptr(c), d.size, mode, C.byref(rng),
                                      C.byref(nug), C.byref(sill), C.byref(slope), C.byref(m), C.byref(sse))
    if rc != A.PB_OK:
        raise PragaB200Error(rc, "pb_kriging_fit_variogram failed")
    return {"mode": m.value, "range": rng.value, "nugget": nug.value, "sill": sill.value, "slope": slope.value, "sse": sse.value}


class Engine:
    """One engine (CUDA context + stream + device-resident raster cache) on one GPU."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.pb_create(device, C.byref(h))
        if rc != A.PB_OK:
            raise PragaB200Error(rc, self.lib.pb_last_error(None).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.pb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, allow=()):
        if rc != A.PB_OK and rc not in allow:
            raise PragaB200Error(rc, self.lib.pb_last_error(self.h).decode())
        return rc

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.pb_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def set_blocking_sync(self, blocking=True):
        """host threads sleep instead of spinning while they wait for the device (more waiting threads than cores)"""
        self.lib.pb_set_blocking_sync.argtypes = [C.c_void_p, C.c_int]
        return self.lib.pb_set_blocking_sync(self.h, int(blocking)) == A.PB_OK

    def timing(self):
        t = A.pb_timing()
        self.lib.pb_last_timing(self.h, C.byref(t))
        return {k: getattr(t, k) for k, _ in A.pb_timing._fields_}

    def measure_pipe_peaks(self):
        """(FP32 TFLOP/s, G MUFU.RSQ/s) measured live on this GPU."""
        a, b = C.c_float(0), C.c_float(0)
        self._check(self.lib.pb_measure_pipe_peaks(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def measure_fp64_peak(self):
        """FP64 FMA TFLOP/s measured live on this GPU (DFMA chain)."""
        a = C.c_float(0)
        self._check(self.lib.pb_measure_fp64_peak(self.h, C.byref(a)))
        return a.value

    # ---- raster cache ----
    def host_register(self, array):
        """Page-lock a long-lived contiguous numpy buffer (DEM / proxy / output grid of the host-raster calls)."""
        self._check(self.lib.pb_host_register(self.h, array.ctypes.data, array.nbytes))

    def host_unregister(self, array):
        self._check(self.lib.pb_host_unregister(self.h, array.ctypes.data))

    def upload(self, raster):
        rid = C.c_int32(-1)
        r = raster.as_pb()
        self._check(self.lib.pb_raster_upload(self.h, C.byref(r), C.byref(rid)))
        return rid.value

    def free(self, rid):
        self._check(self.lib.pb_raster_free(self.h, rid))

    def valid_cells(self, rid):
        n = C.c_int64(0)
        self._check(self.lib.pb_raster_valid_cells(self.h, rid, C.byref(n)))
        return n.value

    # ---- interpolationRaster (interpolationCmd.cpp:353-394) ----
    def interpolation_raster(self, stations, settings, variable, dem, proxy_grids=(), row_begin=0, row_end=-1,
                             out=None, rows=None, local=False):
        """Drop-in call with HOST rasters. Returns (ok, grid, minimum, maximum, nValid); ok False <=> the reference
        returns false (no valid output cell)."""
        st = stations.as_pb()
        d = dem.as_pb()
        grids = (A.pb_raster * max(1, len(proxy_grids)))()
        for i, g in enumerate(proxy_grids):
            grids[i] = d if g is dem else g.as_pb()
        o = A.pb_raster_out()
        if rows is not None:
            o.rows = rows
        else:
            if out is None:
                out = np.empty(dem.shape, dtype=np.float32)
            o.data = A.fptr(out)
        fn = self.lib.pb_local_detrending_raster if local else self.lib.pb_interpolation_raster
        rc = self._check(fn(self.h, C.byref(st), C.byref(settings), variable, C.byref(d), grids, len(proxy_grids),
                            row_begin, row_end, C.byref(o)), allow=(A.PB_ERR_NO_VALID_CELL,))
        return rc == A.PB_OK, out, o.minimum, o.maximum, o.nValid

    def local_detrending_raster(self, stations, settings, variable, dem, proxy_grids=(), row_begin=0, row_end=-1, out=None):
        """loop of Project::interpolationDemLocalDetrending (project.cpp:3208-3253) with HOST rasters."""
        return self.interpolation_raster(stations, settings, variable, dem, proxy_grids, row_begin, row_end, out, local=True)

    def interpolation_raster_resident(self, stations, settings, variable, dem_id, proxy_ids=(), row_begin=0, row_end=-1,
                                      out=None, device_only=False, local=False):
        st = stations.as_pb()
        ids = (C.c_int32 * max(1, len(proxy_ids)))(*proxy_ids)
        o = A.pb_raster_out()
        if not device_only:
            assert out is not None
            o.data = A.fptr(out)
        name = "pb_local_detrending_raster" if local else "pb_interpolation_raster"
        fn = getattr(self.lib, name + ("_device" if device_only else "_resident"))
        rc = self._check(fn(self.h, C.byref(st), C.byref(settings), variable, dem_id, ids, len(proxy_ids), row_begin,
                            row_end, C.byref(o)), allow=(A.PB_ERR_NO_VALID_CELL,))
        return rc == A.PB_OK, out, o.minimum, o.maximum, o.nValid

    def set_topographic_distance_maps(self, point_index=(), map_ids=()):
        """Crit3DMeteoPoint::topographicDistance (Project::loadTopographicDistanceMaps, project.cpp:2368-2430): the points'
        precomputed maps as resident raster ids, keyed by Crit3DInterpolationDataPoint::index; no arguments = drop all."""
        n = len(point_index)
        if n == 0:
            self._check(self.lib.pb_set_topographic_distance_maps(self.h, None, None, 0))
            return
        idx = (C.c_int32 * n)(*[int(i) for i in point_index])
        ids = (C.c_int32 * n)(*[int(i) for i in map_ids])
        self._check(self.lib.pb_set_topographic_distance_maps(self.h, idx, ids, n))

    # ---- interpolate() at explicit targets (interpolation.cpp:2502) ----
    def interpolate_points(self, stations, settings, variable, x, y, proxy_values, exclude_supplemental=False, z=None,
                           dem_id=-1):
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.ascontiguousarray(y, dtype=np.float32)
        m = x.shape[0]
        pv = np.ascontiguousarray(proxy_values, dtype=np.float64).reshape(m, -1)
        res = np.empty(m, dtype=np.float32)
        st = stations.as_pb()
        if z is not None:
            z = np.ascontiguousarray(z, dtype=np.float32)
            self._check(self.lib.pb_interpolate_points_td(self.h, C.byref(st), C.byref(settings), variable, A.fptr(x),
                                                          A.fptr(y), A.fptr(z), A.dptr(pv) if pv.size else None, m,
                                                          int(exclude_supplemental), dem_id, A.fptr(res)))
            return res
        self._check(self.lib.pb_interpolate_points(self.h, C.byref(st), C.byref(settings), variable, A.fptr(x), A.fptr(y),
                                                   A.dptr(pv) if pv.size else None, m, int(exclude_supplemental),
                                                   A.fptr(res)))
        return res

    # ---- computeResiduals + computeStatisticsCrossValidation (spatialControl.cpp:97, project.cpp:2935) ----
    def compute_residuals(self, stations, settings, variable, observed, use_for_cv=None, local=False):
        """computeResiduals (local=True: computeResidualsLocalDetrending, stations NOT detrended) + CV statistics."""
        obs = np.ascontiguousarray(observed, dtype=np.float32)
        res = np.empty(stations.n, dtype=np.float32)
        stats = A.pb_cv_stats()
        st = stations.as_pb()
        use = None if use_for_cv is None else np.ascontiguousarray(use_for_cv, dtype=np.uint8)
        fn = self.lib.pb_compute_residuals_local if local else self.lib.pb_compute_residuals
        self._check(fn(self.h, C.byref(st), C.byref(settings), variable, A.fptr(obs), None if use is None else A.u8ptr(use),
                       A.fptr(res), C.byref(stats)))
        return res, {k: getattr(stats, k) for k, _ in A.pb_cv_stats._fields_}

    # ---- raster map operators around the path ----
    def _grid_call(self, fn, shape, keep, want_host, *args):
        out = np.empty(shape, dtype=np.float32) if want_host else None
        o = A.pb_raster_out()
        if out is not None:
            o.data = A.fptr(out)
        rid = C.c_int32(-1)
        rc = self._check(fn(self.h, *args, C.byref(o), C.byref(rid) if keep else None), allow=(A.PB_ERR_NO_VALID_CELL,))
        return rc == A.PB_OK, out, o.minimum, o.maximum, (rid.value if keep else None)

    def read_esri(self, path):
        """gis::readEsriGrid (gisIO.cpp:516-557) straight into a resident raster. Returns (id, header, min, max)."""
        rid, h = C.c_int32(-1), A.pb_raster_header()
        mn, mx = C.c_float(0), C.c_float(0)
        self._check(self.lib.pb_raster_read_esri(self.h, path.encode(), C.byref(rid), C.byref(h), C.byref(mn), C.byref(mx)))
        return rid.value, h, mn.value, mx.value

    def write_esri(self, rid, path):
        """gis::writeEsriGrid (gisIO.cpp:504-513) from a resident raster; path without extension."""
        self._check(self.lib.pb_raster_write_esri(self.h, rid, path.encode()))

    def download(self, rid, shape):
        out = np.empty(shape, dtype=np.float32)
        o = A.pb_raster_out()
        o.data = A.fptr(out)
        self._check(self.lib.pb_raster_download(self.h, rid, C.byref(o)))
        return out

    def relative_humidity_map(self, air_t_id, dew_t_id, shape, keep=False, want_host=True):
        """computeRelativeHumidityMap (meteoMaps.cpp:301-321). Returns (ok, grid, min, max, resident id or None)."""
        return self._grid_call(self.lib.pb_relative_humidity_map, shape, keep, want_host, air_t_id, dew_t_id)

    def fix_daily_thermal_consistency(self, tmax_id, tmin_id):
        """fixDailyThermalConsistency (meteoMaps.cpp:103-126), in place on the resident tmin map. Returns cells changed."""
        n = C.c_int64(0)
        self._check(self.lib.pb_fix_daily_thermal_consistency(self.h, tmax_id, tmin_id, C.byref(n)))
        return n.value

    def topographic_distance_map(self, dem_id, shape, x, y, z, keep=False, want_host=True):
        """gis::topographicDistanceMap (gis.cpp:1648-1672) for one point."""
        return self._grid_call(self.lib.pb_topographic_distance_map, shape, keep, want_host, dem_id, x, y, z)

    def resample_grid(self, src_id, new_header, method, threshold=0.0, keep=False, want_host=True):
        """gis::resampleGrid (gis.cpp:1722-1811)."""
        return self._grid_call(self.lib.pb_resample_grid, (new_header.nrRows, new_header.nrCols), keep, want_host, src_id,
                               C.byref(new_header), method, threshold)

    def topographic_index(self, dem_id, shape, widths, keep=False, want_host=True):
        """topographicIndex (interpolationCmd.cpp:397-482): the orography-index proxy of a resident DEM."""
        w = np.ascontiguousarray(widths, dtype=np.float32)
        return self._grid_call(self.lib.pb_topographic_index, shape, keep, want_host, dem_id, A.fptr(w), w.size)

    # ---- ordinary kriging (kriging.h:4-15) ----
    def kriging_setup(self, pos, val, mode, rng, nugget, sill, slope=0.0, fp64_direct=False):
        """krigingVariogram (kriging.cpp:97-202). Returns a system id."""
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1)
        val = np.ascontiguousarray(val, dtype=np.float64)
        kid = C.c_int32(-1)
        self._check(self.lib.pb_kriging_setup(self.h, A.dptr(pos), A.dptr(val), val.shape[0], mode, rng, nugget, sill, slope,
                                              1 if fp64_direct else 0, C.byref(kid)))
        return kid.value

    def kriging_free(self, kid):
        self._check(self.lib.pb_kriging_free(self.h, kid))

    def kriging_uses_tensor_path(self, kid):
        return self.lib.pb_kriging_uses_tensor_path(self.h, kid) == 1

    def kriging_mma_count(self, kid):
        """tensor-core MMAs (2 * 256 * 256 * 16 flop each) issued for this system since the last call (measurement aid)"""
        n = C.c_ulonglong(0)
        self._check(self.lib.pb_kriging_mma_count(self.h, kid, C.byref(n)))
        return int(n.value)

    def kriging_points(self, kid, xy, want_rmse=True):
        """krigingSetWeight + krigingResult + krigingRMSE (kriging.cpp:211-295) at m points."""
        xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1)
        m = xy.shape[0] // 2
        res = np.empty(m, dtype=np.float64)
        rm = np.empty(m, dtype=np.float64)
        self._check(self.lib.pb_kriging_points(self.h, kid, A.dptr(xy), m, A.dptr(res), A.dptr(rm) if want_rmse else None))
        return res, (rm if want_rmse else None)

    def kriging_raster(self, kid, dem_id, shape, row_begin=0, row_end=-1, want_rmse=True, device_only=False, est_out=None,
                       rmse_out=None):
        """estimate (and RMSE) grids over the valid cells of a resident DEM. est_out / rmse_out: reusable host grids."""
        est = A.pb_raster_out()
        rm = A.pb_raster_out()
        e = r = None
        if not device_only:
            e = est_out if est_out is not None else np.empty(shape, dtype=np.float32)
            est.data = A.fptr(e)
            if want_rmse:
                r = rmse_out if rmse_out is not None else np.empty(shape, dtype=np.float32)
                rm.data = A.fptr(r)
        self._check(self.lib.pb_kriging_raster(self.h, kid, dem_id, row_begin, row_end, C.byref(est),
                                               C.byref(rm) if want_rmse else None))
        return e, r, est.nValid

    # ---- preInterpolation (interpolation.cpp:2444) ----
    def pre_interpolation(self, stations, settings, variable):
        """one time step, in place (stations.value detrended, settings get the fit). Returns the keep mask."""
        st = stations.as_pb()
        keep = np.zeros(stations.n, dtype=np.uint8)
        self._check(self.lib.pb_pre_interpolation(self.h, C.byref(st), C.byref(settings), variable, A.u8ptr(keep)))
        return keep.astype(bool)

    def pre_interpolation_ex(self, stations, settings, variable, current_value=None, inside_dem=None, exclude_outside_dem=False,
                             dem_id=-1):
        """preInterpolation with every branch (classic detrending, optimalDetrending, kh golden section), in place.
        Returns (keep mask, [(kh, error), ...])."""
        st = stations.as_pb()
        keep = np.zeros(stations.n, dtype=np.uint8)
        cv = A.pb_cv_points()
        cur = None if current_value is None else np.ascontiguousarray(current_value, dtype=np.float32)
        ins = None if inside_dem is None else np.ascontiguousarray(inside_dem, dtype=np.uint8)
        if cur is not None:
            cv.currentValue = A.fptr(cur)
        if ins is not None:
            cv.isInsideDem = A.u8ptr(ins)
        cv.excludeOutsideDem = int(exclude_outside_dem)
        series = A.pb_kh_series()
        self._check(self.lib.pb_pre_interpolation_ex(self.h, C.byref(st), C.byref(settings), variable, C.byref(cv), dem_id,
                                                     A.u8ptr(keep), C.byref(series)))
        return keep.astype(bool), [(series.kh[i], series.error[i]) for i in range(series.n)]

    def spatial_quality_control(self, stations, settings, variable, inside_dem=None, exclude_outside_dem=False, dem_id=-1):
        """spatialQualityControl (spatialControl.cpp:331-427): stations carry the raw current values of the accepted meteo
        points; settings (the spatial-quality settings) are mutated like the reference does. Returns the wrong_spatial mask."""
        st = stations.as_pb()
        cv = A.pb_cv_points()
        ins = None if inside_dem is None else np.ascontiguousarray(inside_dem, dtype=np.uint8)
        if ins is not None:
            cv.isInsideDem = A.u8ptr(ins)
        cv.excludeOutsideDem = int(exclude_outside_dem)
        wrong = np.zeros(stations.n, dtype=np.uint8)
        self._check(self.lib.pb_spatial_quality_control(self.h, C.byref(st), C.byref(settings), variable, C.byref(cv), dem_id,
                                                        A.u8ptr(wrong)))
        return wrong.astype(bool)

    def compute_residuals_exact(self, meteo, points, settings, variable, observed=None, inside_dem=None, dem_id=-1,
                                exclude_outside_dem=False, exclude_supplemental=False):
        """computeResiduals (spatialControl.cpp:97-155) in the reference's exact arithmetic, topographic distance included."""
        m, p = meteo.as_pb(), points.as_pb()
        cv = A.pb_cv_points()
        ins = None if inside_dem is None else np.ascontiguousarray(inside_dem, dtype=np.uint8)
        if ins is not None:
            cv.isInsideDem = A.u8ptr(ins)
        obs = None if observed is None else np.ascontiguousarray(observed, dtype=np.float32)
        res = np.empty(meteo.n, dtype=np.float32)
        self._check(self.lib.pb_compute_residuals_exact(self.h, C.byref(m), None if obs is None else A.fptr(obs), C.byref(cv),
                                                        C.byref(p), C.byref(settings), variable, dem_id, int(exclude_outside_dem),
                                                        int(exclude_supplemental), A.fptr(res)))
        return res

    def interpolation_dem(self, meteo, settings, variable, dem_id, shape, proxy_ids=(), quality_settings=None, cross_validate=False,
                          row_begin=0, row_end=-1, out=None, device_only=False):
        """Project::interpolationDem (project.cpp:3106-3150) for one time step: spatial quality control, preInterpolation,
        interpolationRaster over resident rasters [+ the leave-one-out of Project::interpolationCv]. meteo: RAW values.
        Returns dict(ok, grid, min, max, n_valid, wrong_spatial, residual, stats, kh_series)."""
        st = meteo.as_pb()
        ids = (C.c_int32 * max(1, len(proxy_ids)))(*proxy_ids)
        o = A.pb_raster_out()
        if not device_only:
            if out is None:
                out = np.empty(shape, dtype=np.float32)
            o.data = A.fptr(out)
        wrong = np.zeros(meteo.n, dtype=np.uint8)
        res = np.empty(meteo.n, dtype=np.float32)
        stats = A.pb_cv_stats()
        series = A.pb_kh_series()
        step = A.pb_dem_step()
        step.wrongSpatial = A.u8ptr(wrong)
        step.khSeries = C.pointer(series)
        if cross_validate:
            step.residual = A.fptr(res)
            step.stats = C.pointer(stats)
        rc = self.lib.pb_interpolation_dem(self.h, C.byref(st), None if quality_settings is None else C.byref(quality_settings),
                                           C.byref(settings), variable, dem_id, ids, len(proxy_ids), row_begin, row_end,
                                           int(device_only), C.byref(o), C.byref(step))
        self._check(rc, allow=(A.PB_ERR_NO_VALID_CELL,))
        return dict(ok=rc == A.PB_OK, grid=out, min=o.minimum, max=o.maximum, n_valid=o.nValid, wrong_spatial=wrong.astype(bool),
                    residual=res if cross_validate else None,
                    stats={k: getattr(stats, k) for k, _ in A.pb_cv_stats._fields_} if cross_validate else None,
                    kh_series=[(series.kh[i], series.error[i]) for i in range(series.n)])

    def interpolation_dem_batch(self, units, dem_id, shape, proxy_ids=(), cross_validate=False, device_only=False, lanes=0,
                                outs=None):
        """pb_interpolation_dem_batch. units: list of (meteo stations, settings, quality settings or None, variable).
        Returns a list of dicts like interpolation_dem (grid None when device_only)."""
        n = len(units)
        items = (A.pb_dem_batch_item * max(1, n))()
        keep = []
        for i, (meteo, s, sq, var) in enumerate(units):
            st = meteo.as_pb()
            o = A.pb_raster_out()
            grid = None
            if not device_only:
                grid = outs[i] if outs is not None else np.empty(shape, dtype=np.float32)
                o.data = A.fptr(grid)
            wrong = np.zeros(meteo.n, dtype=np.uint8)
            res = np.empty(meteo.n, dtype=np.float32)
            stats, series, step = A.pb_cv_stats(), A.pb_kh_series(), A.pb_dem_step()
            step.wrongSpatial = A.u8ptr(wrong)
            step.khSeries = C.pointer(series)
            if cross_validate:
                step.residual = A.fptr(res)
                step.stats = C.pointer(stats)
            items[i].meteo = C.pointer(st)
            items[i].settings = C.pointer(s)
            if sq is not None:
                items[i].qualitySettings = C.pointer(sq)
            items[i].variable = var
            items[i].out = C.pointer(o)
            items[i].step = C.pointer(step)
            keep.append((meteo, st, o, grid, wrong, res, stats, series, step))
        ids = (C.c_int32 * max(1, len(proxy_ids)))(*proxy_ids)
        rc = self.lib.pb_interpolation_dem_batch(self.h, items, n, dem_id, ids, len(proxy_ids), 0, -1, int(device_only), lanes)
        self._check(rc, allow=(A.PB_ERR_NO_VALID_CELL,))
        out = []
        for i, (meteo, st, o, grid, wrong, res, stats, series, step) in enumerate(keep):
            out.append(dict(ok=items[i].status == A.PB_OK, status=items[i].status, grid=grid, min=o.minimum, max=o.maximum,
                            n_valid=o.nValid, wrong_spatial=wrong.astype(bool), residual=res if cross_validate else None,
                            stats={k: getattr(stats, k) for k, _ in A.pb_cv_stats._fields_} if cross_validate else None,
                            kh_series=[(series.kh[j], series.error[j]) for j in range(series.n)]))
        return out

    def meteo_grid_step_batch(self, hours, dem_id, shape, proxy_ids=(), aggregation_id=-1, n_cells=0, lanes=0, cross_validate=False):
        """pb_meteo_grid_step_batch: the hourly driver kept on the device (PragaProject::interpolationMeteoGrid with
        getMeteoGridUpscaleFromDem(), pragaProject.cpp:2951-2990). hours: list of lists of dicts with keys meteo, settings,
        variable and optionally quality_settings, use_dew_point, air_temperature, aggr_method (default average), want_raster.
        Returns the same nesting of dicts(status, ok, cells, grid, min, max, n_valid, wrong_spatial, residual, stats)."""
        nh = len(hours)
        harr = (A.pb_meteo_grid_hour * max(1, nh))()
        keep, res = [], []
        for i, hour in enumerate(hours):
            varr = (A.pb_meteo_grid_var * max(1, len(hour)))()
            hres = []
            for k, u in enumerate(hour):
                meteo = u["meteo"]
                st = meteo.as_pb()
                o = A.pb_raster_out()
                grid = None
                if u.get("want_raster"):
                    grid = np.empty(shape, dtype=np.float32)
                    o.data = A.fptr(grid)
                cells = np.empty(max(1, n_cells), dtype=np.float32) if aggregation_id >= 0 else None
                wrong = np.zeros(meteo.n, dtype=np.uint8)
                resid = np.empty(meteo.n, dtype=np.float32)
                stats, step = A.pb_cv_stats(), A.pb_dem_step()
                step.wrongSpatial = A.u8ptr(wrong)
                if cross_validate:
                    step.residual = A.fptr(resid)
                    step.stats = C.pointer(stats)
                v = varr[k]
                v.meteo = C.pointer(st)
                v.settings = C.pointer(u["settings"])
                if u.get("quality_settings") is not None:
                    v.qualitySettings = C.pointer(u["quality_settings"])
                v.variable = u["variable"]
                v.useDewPoint = int(bool(u.get("use_dew_point")))
                air = None
                if u.get("air_temperature") is not None:
                    air = np.ascontiguousarray(u["air_temperature"], dtype=np.float32)
                    v.airTemperature = A.fptr(air)
                v.aggrMethod = u.get("aggr_method", A.PB_AGGR_AVERAGE)
                if cells is not None:
                    v.cellValue = A.fptr(cells)
                v.out = C.pointer(o)
                v.step = C.pointer(step)
                keep.append((meteo, st, o, air, wrong, resid, stats, step))
                hres.append(dict(cells=cells, grid=grid, out=o, wrong=wrong, resid=resid, stats=stats))
            harr[i].vars = varr
            harr[i].nVars = len(hour)
            keep.append(varr)
            res.append((varr, hres))
        ids = (C.c_int32 * max(1, len(proxy_ids)))(*proxy_ids)
        rc = self.lib.pb_meteo_grid_step_batch(self.h, harr, nh, dem_id, ids, len(proxy_ids), aggregation_id, lanes)
        self._check(rc, allow=(A.PB_ERR_NO_VALID_CELL,))
        out = []
        for varr, hres in res:
            row = []
            for k, r in enumerate(hres):
                o = r["out"]
                row.append(dict(status=varr[k].status, ok=varr[k].status == A.PB_OK, cells=r["cells"], grid=r["grid"], min=o.minimum,
                                max=o.maximum, n_valid=o.nValid, wrong_spatial=r["wrong"].astype(bool),
                                residual=r["resid"] if cross_validate else None,
                                stats={f: getattr(r["stats"], f) for f, _ in A.pb_cv_stats._fields_} if cross_validate else None))
            out.append(row)
        return out

    def compute_radiation_dem(self, settings, dem_id, lat_id, lon_id, slope_id, aspect_id, transmissivity_id, shape, want_host=True,
                              keep=False):
        """radiation::computeRadiationDEM (solarRadiation.cpp:1045-1069). Returns dict name -> (grid or None, min, max) for
        sunElevation, beam, diffuse, reflected, global (+ "ids" when keep)."""
        maps = A.pb_radiation_maps(dem_id, lat_id, lon_id, slope_id, aspect_id, transmissivity_id)
        names = ("sunElevation", "beam", "diffuse", "reflected", "global")
        outs, grids = [], []
        ro = A.pb_radiation_out()
        for k, nme in enumerate(names):
            o = A.pb_raster_out()
            g = np.empty(shape, dtype=np.float32) if want_host else None
            if g is not None:
                o.data = A.fptr(g)
            outs.append(o)
            grids.append(g)
            setattr(ro, "global_" if nme == "global" else nme, C.pointer(o))
        ro.keepResident = int(keep)
        self._check(self.lib.pb_compute_radiation_dem(self.h, C.byref(settings), C.byref(maps), C.byref(ro)))
        res = {nme: (grids[k], outs[k].minimum, outs[k].maximum) for k, nme in enumerate(names)}
        if keep:
            res["ids"] = [ro.ids[k] for k in range(5)]
        return res

    def raster_from_last_output(self, like_id):
        rid = C.c_int32(-1)
        self._check(self.lib.pb_raster_from_last_output(self.h, like_id, C.byref(rid)))
        return rid.value

    def pre_interpolation_batch(self, stations, values, settings, variable, points_range=None):
        """T time steps sharing the station set. values: (T, n). Returns (detrended (T, n), keep (T, n), [settings] * T)."""
        values = np.ascontiguousarray(values, dtype=np.float32)
        T, n = values.shape
        assert n == stations.n
        st = stations.as_pb()
        det = np.empty((T, n), dtype=np.float32)
        keep = np.zeros((T, n), dtype=np.uint8)
        fitted = (A.pb_settings * T)()
        pr = None if points_range is None else np.ascontiguousarray(points_range, dtype=np.float64)
        self._check(self.lib.pb_pre_interpolation_batch(self.h, C.byref(st), A.fptr(values), T, C.byref(settings), variable,
                                                        None if pr is None else A.dptr(pr), A.fptr(det), A.u8ptr(keep), fitted))
        return det, keep.astype(bool), fitted

    # ---- glocal detrending (macro areas) ----
    def glocal_detrending_fitting(self, stations, settings, variable, areas):
        """glocalDetrendingFitting (interpolation.cpp:2239-2294) as preInterpolation reaches it: every area of `areas`
        (_abi.MacroAreas) is fitted in one launch and keeps its parameters / combination; settings is mutated like the reference's."""
        st = stations.as_pb()
        self._check(self.lib.pb_glocal_detrending_fitting(self.h, C.byref(st), C.byref(settings), variable, areas.arr, areas.n))

    def macro_areas_upload(self, areas):
        """The areas' (cell, weight) pairs resident on the device (they only change with the glocal weight maps)."""
        mid = C.c_int32(-1)
        self._check(self.lib.pb_macro_areas_upload(self.h, areas.arr, areas.n, C.byref(mid)))
        return mid.value

    def macro_areas_free(self, mid):
        self._check(self.lib.pb_macro_areas_free(self.h, mid))

    def glocal_detrending_raster(self, stations, settings, variable, areas, dem_id, shape, proxy_ids=(), device_only=False,
                                 cells_id=None, out=None, meteo=None, observed=None, inside_dem=None, area_kh=None):
        """gridding part of Project::interpolationDemGlocalDetrending (project.cpp:3283-3375). cells_id: macro_areas_upload id.
        meteo (+ observed, inside_dem): the meteo points of the per-area kh search when settings.useTD is on; area_kh: an
        int32 array (n areas) receiving each area's kh. Returns (ok, grid or None, min, max, nValid); ok False where the
        reference returns false."""
        st = stations.as_pb()
        ids = (C.c_int32 * max(1, len(proxy_ids)))(*proxy_ids)
        grid = None if device_only else (out if out is not None else np.empty(shape, dtype=np.float32))
        o = A.pb_raster_out()
        if grid is not None:
            o.data = A.fptr(grid)
        if meteo is not None:
            mp = meteo.as_pb()
            gm, cv = A.pb_glocal_meteo(), A.pb_cv_points()
            gm.meteo = C.pointer(mp)
            obs = None if observed is None else np.ascontiguousarray(observed, dtype=np.float32)
            if obs is not None:
                gm.observed = A.fptr(obs)
            ins = None if inside_dem is None else np.ascontiguousarray(inside_dem, dtype=np.uint8)
            if ins is not None:
                cv.isInsideDem = A.u8ptr(ins)
                gm.cv = C.pointer(cv)
            rc = self.lib.pb_glocal_detrending_raster_td(self.h, C.byref(st), C.byref(settings), variable, areas.arr, areas.n,
                                                         -1 if cells_id is None else cells_id, C.byref(gm), dem_id, ids,
                                                         len(proxy_ids), int(device_only), C.byref(o),
                                                         None if area_kh is None else A.iptr(area_kh))
        elif cells_id is None:
            rc = self.lib.pb_glocal_detrending_raster(self.h, C.byref(st), C.byref(settings), variable, areas.arr, areas.n, dem_id, ids,
                                                      len(proxy_ids), int(device_only), C.byref(o))
        else:
            rc = self.lib.pb_glocal_detrending_raster_resident(self.h, C.byref(st), C.byref(settings), variable, areas.arr, areas.n,
                                                               cells_id, dem_id, ids, len(proxy_ids), int(device_only), C.byref(o))
        self._check(rc, allow=(A.PB_ERR_NO_VALID_CELL,))
        return rc == A.PB_OK, grid, o.minimum, o.maximum, o.nValid

    def compute_residuals_glocal(self, meteo, observed, points, settings, variable, areas, exclude_outside_dem=True,
                                 exclude_supplemental=True, inside_dem=None, dem_id=None):
        """computeResidualsAndStatisticsGlocalDetrending (project.cpp:6525-6565) over computeResidualsGlocalDetrending
        (spatialControl.cpp:231-302); the areas carry their fit."""
        mp, pt = meteo.as_pb(), points.as_pb()
        obs = np.ascontiguousarray(observed, dtype=np.float32)
        cv = A.pb_cv_points()
        ins = None if inside_dem is None else np.ascontiguousarray(inside_dem, dtype=np.uint8)
        if ins is not None:
            cv.isInsideDem = A.u8ptr(ins)
        res = np.empty(meteo.n, dtype=np.float32)
        if dem_id is not None:      # topographic distance inside the macro areas: the resident DEM of the ray march
            self._check(self.lib.pb_compute_residuals_glocal_td(self.h, C.byref(mp), A.fptr(obs), C.byref(cv), C.byref(pt),
                                                                C.byref(settings), variable, areas.arr, areas.n, dem_id,
                                                                int(exclude_outside_dem), int(exclude_supplemental), A.fptr(res)))
            return res
        self._check(self.lib.pb_compute_residuals_glocal(self.h, C.byref(mp), A.fptr(obs), C.byref(cv), C.byref(pt), C.byref(settings),
                                                         variable, areas.arr, areas.n, int(exclude_outside_dem),
                                                         int(exclude_supplemental), A.fptr(res)))
        return res

    # ---- DEM -> meteo-grid aggregation ----
    def aggregation_upload(self, x, y, first, max_nr):
        """Aggregation points of every meteo-grid cell (Crit3DMeteoPoint::aggregationPoints): x, y concatenated, first[c]..first[c+1]."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        first = np.ascontiguousarray(first, dtype=np.int64)
        max_nr = np.ascontiguousarray(max_nr, dtype=np.int32)
        aid = C.c_int32(-1)
        self._check(self.lib.pb_aggregation_upload(self.h, A.dptr(x), A.dptr(y), first.ctypes.data_as(C.POINTER(C.c_int64)),
                                                   A.iptr(max_nr), max_nr.size, C.byref(aid)))
        self._agg_cells = getattr(self, "_agg_cells", {})
        self._agg_cells[aid.value] = int(max_nr.size)
        return aid.value

    def aggregation_free(self, aid):
        self._check(self.lib.pb_aggregation_free(self.h, aid))

    def spatial_aggregate_meteo_grid(self, raster_id, aid, method):
        """Crit3DMeteoGrid::spatialAggregateMeteoGrid (meteoGrid.cpp:1135-1190) of a resident raster: one value per cell."""
        val = np.empty(self._agg_cells[aid], dtype=np.float32)
        self._check(self.lib.pb_spatial_aggregate_meteo_grid(self.h, raster_id, aid, method, A.fptr(val)))
        return val

    def raster_sample_points(self, raster_id, x, y):
        """Nearest-cell values of a resident raster at points (passInterpolatedTemperatureToHumidityPoints, project.cpp:2826-2848).
        Returns (values, in_grid mask)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        val = np.empty(x.size, dtype=np.float32)
        ins = np.zeros(x.size, dtype=np.uint8)
        self._check(self.lib.pb_raster_sample_points(self.h, raster_id, A.dptr(x), A.dptr(y), x.size, A.fptr(val), A.u8ptr(ins)))
        return val, ins.astype(bool)
