"""ctypes binding of the CPU checkers (TEST INFRASTRUCTURE: imported by tests/, smoke() and bench.py's CPU legs only)."""
import ctypes as C
import os
import subprocess

import numpy as np

from praga_b200 import _abi as A

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libpraga_ref.so")
PORT_LIB = os.path.join(HERE, "liboracle_port.so")
_libs = {}


def build(target="all"):
    """Run oracle/Makefile (compiles the restatement; the reference only where /root/reference exists)."""
    res = subprocess.run(["make", "-s", "-C", HERE, "-j8", target], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)


def have_ref():
    return os.path.exists(REF_LIB)


SHIM_LIB = os.path.join(HERE, "_ref", "libpraga_shim_test.so")


def have_shim():
    return os.path.exists(SHIM_LIB)


def load_shim():
    """oracle/_ref/libpraga_shim_test.so: the C++ host shim driven with reference objects (needs a B200 to run)."""
    lib = C.CDLL(SHIM_LIB)
    P = C.POINTER
    lib.shim_raster.argtypes = [P(A.pb_stations), P(A.pb_settings), C.c_int, P(A.pb_raster), P(A.pb_raster), C.c_int, C.c_int,
                                P(A.pb_raster_out), C.c_char_p, C.c_int]
    lib.shim_compute_residuals.argtypes = [P(A.pb_stations), P(A.pb_settings), C.c_int, P(C.c_float), P(C.c_uint8), P(C.c_float)]
    lib.shim_pre_interpolation.argtypes = [P(A.pb_stations), P(A.pb_settings), C.c_int, P(A.pb_cv_points), P(A.pb_raster),
                                           P(C.c_uint8), P(A.pb_kh_series), C.c_char_p, C.c_int]
    lib.shim_spatial_quality_control.argtypes = [P(A.pb_stations), P(A.pb_settings), C.c_int, P(A.pb_cv_points), P(A.pb_raster),
                                                 P(C.c_uint8)]
    lib.shim_compute_residuals_local.argtypes = lib.shim_compute_residuals.argtypes
    lib.shim_session_open.argtypes = [P(A.pb_settings), P(A.pb_raster), P(A.pb_raster), C.c_int]
    lib.shim_session_open.restype = C.c_void_p
    lib.shim_session_step.argtypes = [C.c_void_p, P(A.pb_stations), C.c_int, C.c_int, P(C.c_double), P(C.c_float), P(C.c_float)]
    lib.shim_session_output.argtypes = [C.c_void_p, P(A.pb_raster_out)]
    lib.shim_session_poke_dem.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int]
    lib.shim_session_poke_dem.restype = None
    lib.shim_session_set_device.argtypes = [C.c_int]
    lib.shim_session_timing.argtypes = [P(A.pb_timing)]
    lib.shim_session_cache.argtypes = [C.c_int]
    lib.shim_session_cache.restype = None
    lib.shim_session_close.argtypes = [C.c_void_p]
    lib.shim_session_close.restype = None
    lib.shim_kriging_points.argtypes = [P(C.c_double), P(C.c_double), C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                        C.c_double, P(C.c_double), C.c_int, P(C.c_double), P(C.c_double)]
    lib.shim_estimate_variogram.argtypes = [P(C.c_float), P(C.c_float), C.c_int, C.c_float, P(C.c_int), P(C.c_double), P(C.c_double),
                                            P(C.c_double), P(C.c_double)]
    lib.shim_pre_interpolation_glocal.argtypes = [P(A.pb_stations), P(A.pb_settings), C.c_int, P(A.pb_macro_area), C.c_int, C.c_char_p,
                                                  C.c_int]
    lib.shim_glocal.argtypes = [P(A.pb_stations), P(A.pb_settings), C.c_int, P(A.pb_macro_area), C.c_int, P(A.pb_raster),
                                P(A.pb_raster), C.c_int, P(A.pb_raster_out), P(C.c_float), C.c_char_p, C.c_int]
    return lib


def load_ref_extra():
    """entry points only the reference driver has (composites of reference calls)."""
    lib = _load(REF_LIB, "ref")
    P = C.POINTER
    lib.ref_interpolation_dem.argtypes = [P(A.pb_stations), P(A.pb_settings), P(A.pb_settings), C.c_int, P(A.pb_raster),
                                          P(A.pb_raster), C.c_int, C.c_int, P(A.pb_raster_out), P(A.pb_dem_step)]
    return lib


def load_ref_sample():
    """reference driver only: the raster lookup of passInterpolatedTemperatureToHumidityPoints"""
    lib = _load(REF_LIB, "ref")
    lib.ref_raster_sample_points.argtypes = [C.POINTER(A.pb_raster), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int,
                                             C.POINTER(C.c_float), C.POINTER(C.c_uint8)]
    return lib


def load_ref_io():
    """ESRI grid I/O of the reference (gis/gisIO.cpp), reference driver only."""
    lib = _load(REF_LIB, "ref")
    P = C.POINTER
    lib.ref_write_esri.argtypes = [P(A.pb_raster), C.c_char_p]
    lib.ref_read_esri_header.argtypes = [C.c_char_p, P(A.pb_raster_header)]
    lib.ref_read_esri.argtypes = [C.c_char_p, P(A.pb_raster_out)]
    return lib


def have_port():
    return os.path.exists(PORT_LIB)


def _load(path, prefix):
    if path in _libs:
        return _libs[path]
    lib = C.CDLL(path)
    P = C.POINTER
    fn = getattr(lib, prefix + "_interpolation_raster")
    fn.argtypes = [P(A.pb_stations), P(A.pb_settings), C.c_int, P(A.pb_raster), P(A.pb_raster), C.c_int, C.c_int,
                   P(A.pb_raster_out), P(C.c_double)]
    fn = getattr(lib, prefix + "_pre_interpolation")
    fn.argtypes = [P(A.pb_stations), P(A.pb_settings), C.c_int, P(C.c_uint8), P(A.pb_raster), C.c_int]
    fn = getattr(lib, prefix + "_pre_interpolation_ex")
    fn.argtypes = [P(A.pb_stations), P(A.pb_settings), C.c_int, P(A.pb_cv_points), P(A.pb_raster), P(C.c_uint8),
                   P(A.pb_kh_series)]
    fn = getattr(lib, prefix + "_spatial_quality_control")
    fn.argtypes = [P(A.pb_stations), P(A.pb_settings), C.c_int, P(A.pb_cv_points), P(A.pb_raster), P(C.c_uint8)]
    fn = getattr(lib, prefix + "_interpolate_points")
    fn.argtypes = [P(A.pb_stations), P(A.pb_settings), C.c_int, P(C.c_float), P(C.c_float), P(C.c_float),
                   P(C.c_double), C.c_int, C.c_int, P(C.c_float)]
    fn = getattr(lib, prefix + "_interpolate_points_td")
    fn.argtypes = [P(A.pb_stations), P(A.pb_settings), C.c_int, P(C.c_float), P(C.c_float), P(C.c_float),
                   P(C.c_double), C.c_int, C.c_int, P(A.pb_raster), P(C.c_float)]
    fn = getattr(lib, prefix + "_compute_residuals")
    fn.argtypes = [P(A.pb_stations), P(A.pb_settings), C.c_int, P(C.c_float), P(C.c_uint8), P(C.c_float),
                   P(A.pb_cv_stats)]
    fn = getattr(lib, prefix + "_compute_residuals_local")
    fn.argtypes = [P(A.pb_stations), P(A.pb_settings), C.c_int, P(C.c_float), P(C.c_uint8), P(C.c_float),
                   P(A.pb_cv_stats)]
    fn = getattr(lib, prefix + "_relative_humidity_map")
    fn.argtypes = [P(A.pb_raster), P(A.pb_raster), P(A.pb_raster_out)]
    fn = getattr(lib, prefix + "_fix_daily_thermal_consistency")
    fn.argtypes = [P(A.pb_raster), P(A.pb_raster), P(C.c_float), P(C.c_int64)]
    fn = getattr(lib, prefix + "_topographic_distance_map")
    fn.argtypes = [P(A.pb_raster), C.c_double, C.c_double, C.c_double, P(A.pb_raster_out)]
    fn = getattr(lib, prefix + "_resample_grid")
    fn.argtypes = [P(A.pb_raster), P(A.pb_raster_header), C.c_int, C.c_float, P(A.pb_raster_out)]
    fn = getattr(lib, prefix + "_kriging_points")
    fn.argtypes = [P(C.c_double), P(C.c_double), C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                   P(C.c_double), C.c_int, P(C.c_double), P(C.c_double), P(C.c_double), P(C.c_double)]
    getattr(lib, prefix + "_max_threads").restype = C.c_int
    fn = getattr(lib, prefix + "_local_detrending_raster")
    fn.argtypes = [P(A.pb_stations), P(A.pb_settings), C.c_int, P(A.pb_raster), P(A.pb_raster), C.c_int, C.c_int,
                   C.c_int, C.c_int, P(A.pb_raster_out), P(C.c_double)]
    fn = getattr(lib, prefix + "_local_detrending_point")
    fn.argtypes = [P(A.pb_stations), P(A.pb_settings), C.c_int, C.c_double, C.c_double, C.c_float, P(C.c_double),
                   P(C.c_int), P(C.c_float), P(C.c_float), P(C.c_int), P(C.c_float)]
    fn = getattr(lib, prefix + "_glocal_detrending_fitting")
    fn.argtypes = [P(A.pb_stations), P(A.pb_settings), C.c_int, P(A.pb_macro_area), C.c_int]
    fn = getattr(lib, prefix + "_glocal_detrending_raster")
    fn.argtypes = [P(A.pb_stations), P(A.pb_settings), C.c_int, P(A.pb_macro_area), C.c_int, P(A.pb_raster), P(A.pb_raster),
                   C.c_int, C.c_int, P(A.pb_raster_out), P(C.c_double)]
    fn = getattr(lib, prefix + "_compute_residuals_glocal")
    fn.argtypes = [P(A.pb_stations), P(C.c_float), P(A.pb_cv_points), P(A.pb_stations), P(A.pb_settings), C.c_int,
                   P(A.pb_macro_area), C.c_int, P(A.pb_raster), C.c_int, C.c_int, P(C.c_float)]
    fn = getattr(lib, prefix + "_spatial_aggregate_meteo_grid")
    fn.argtypes = [P(A.pb_raster), P(C.c_double), P(C.c_double), P(C.c_int64), P(C.c_int32), C.c_int, C.c_int, P(C.c_float)]
    fn = getattr(lib, prefix + "_topographic_index")
    fn.argtypes = [P(A.pb_raster), P(C.c_float), C.c_int, P(A.pb_raster_out)]
    _libs[path] = lib
    return lib


class Oracle:
    """One of the two CPU checkers behind the same Python calls. kind = 'reference' | 'port'."""

    def __init__(self, kind="reference"):
        self.kind = kind
        if kind == "reference":
            if not have_ref():
                raise RuntimeError("oracle/_ref/libpraga_ref.so missing: run `make -C oracle ref` where /root/reference exists")
            self.lib, self.p = _load(REF_LIB, "ref"), "ref"
        else:
            if not have_port():
                build("port")
            self.lib, self.p = _load(PORT_LIB, "port"), "port"

    def _f(self, name):
        return getattr(self.lib, f"{self.p}_{name}")

    def max_threads(self):
        return self._f("max_threads")()

    def interpolation_raster(self, stations, settings, variable, dem, proxy_grids=(), threads=0):
        """interpolationRaster (interpolationCmd.cpp:353-394). Returns (ok, grid, min, max, nValid, seconds)."""
        st = stations.as_pb()
        d = dem.as_pb()
        grids = (A.pb_raster * max(1, len(proxy_grids)))()
        for i, g in enumerate(proxy_grids):
            grids[i] = g.as_pb()
        out = np.empty(dem.shape, dtype=np.float32)
        o = A.pb_raster_out()
        o.data = A.fptr(out)
        sec = C.c_double(0)
        rc = self._f("interpolation_raster")(C.byref(st), C.byref(settings), variable, C.byref(d), grids,
                                             len(proxy_grids), threads, C.byref(o), C.byref(sec))
        return rc == A.PB_OK, out, o.minimum, o.maximum, o.nValid, sec.value

    def local_detrending_raster(self, stations, settings, variable, dem, proxy_grids=(), threads=0, row_begin=0,
                                row_end=-1):
        """loop of Project::interpolationDemLocalDetrending (project.cpp:3208-3253).
        Returns (ok, grid, min, max, nValid, seconds)."""
        st = stations.as_pb()
        d = dem.as_pb()
        grids = (A.pb_raster * max(1, len(proxy_grids)))()
        for i, g in enumerate(proxy_grids):
            grids[i] = g.as_pb()
        out = np.empty(dem.shape, dtype=np.float32)
        o = A.pb_raster_out()
        o.data = A.fptr(out)
        sec = C.c_double(0)
        rc = self._f("local_detrending_raster")(C.byref(st), C.byref(settings), variable, C.byref(d), grids,
                                                len(proxy_grids), threads, row_begin, row_end, C.byref(o), C.byref(sec))
        return rc == A.PB_OK, out, o.minimum, o.maximum, o.nValid, sec.value

    def local_detrending_point(self, stations, settings, variable, x, y, z, proxy_values):
        """one cell of the local-detrending loop; returns dict(sel, weight, value, result); settings get the fit."""
        st = stations.as_pb()
        pv = np.ascontiguousarray(proxy_values, dtype=np.float64)
        sel = np.zeros(stations.n, dtype=np.int32)
        w = np.zeros(stations.n, dtype=np.float32)
        v = np.zeros(stations.n, dtype=np.float32)
        n = C.c_int(0)
        res = C.c_float(0)
        rc = self._f("local_detrending_point")(C.byref(st), C.byref(settings), variable, x, y, z, A.dptr(pv), A.iptr(sel),
                                               A.fptr(w), A.fptr(v), C.byref(n), C.byref(res))
        return dict(ok=rc == A.PB_OK, sel=sel[:n.value], weight=w[:n.value], value=v[:n.value], result=res.value)

    def pre_interpolation(self, stations, settings, variable, proxy_grids=()):
        """preInterpolation (interpolation.cpp:2444): detrends stations.value IN PLACE, mutates settings.
        Returns (ok, keep mask)."""
        st = stations.as_pb()
        grids = (A.pb_raster * max(1, len(proxy_grids)))()
        for i, g in enumerate(proxy_grids):
            grids[i] = g.as_pb()
        keep = np.zeros(stations.n, dtype=np.uint8)
        rc = self._f("pre_interpolation")(C.byref(st), C.byref(settings), variable, A.u8ptr(keep), grids,
                                          len(proxy_grids))
        return rc == A.PB_OK, keep.astype(bool)

    def pre_interpolation_ex(self, stations, settings, variable, current_value=None, inside_dem=None,
                             exclude_outside_dem=False, dem=None):
        """preInterpolation with its meteo points and current DEM (optimalDetrending, kh golden section).
        Returns (ok, keep mask, [(kh, error), ...])."""
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
        d = None if dem is None else dem.as_pb()
        rc = self._f("pre_interpolation_ex")(C.byref(st), C.byref(settings), variable, C.byref(cv),
                                             None if d is None else C.byref(d), A.u8ptr(keep), C.byref(series))
        return rc == A.PB_OK, keep.astype(bool), [(series.kh[i], series.error[i]) for i in range(series.n)]

    def spatial_quality_control(self, stations, settings, variable, inside_dem=None, exclude_outside_dem=False, dem=None):
        """spatialQualityControl (spatialControl.cpp:331-427). Returns (ok, wrong_spatial mask); settings are mutated."""
        st = stations.as_pb()
        cv = A.pb_cv_points()
        ins = None if inside_dem is None else np.ascontiguousarray(inside_dem, dtype=np.uint8)
        if ins is not None:
            cv.isInsideDem = A.u8ptr(ins)
        cv.excludeOutsideDem = int(exclude_outside_dem)
        wrong = np.zeros(stations.n, dtype=np.uint8)
        d = None if dem is None else dem.as_pb()
        rc = self._f("spatial_quality_control")(C.byref(st), C.byref(settings), variable, C.byref(cv),
                                                None if d is None else C.byref(d), A.u8ptr(wrong))
        return rc == A.PB_OK, wrong.astype(bool)

    def interpolation_dem(self, meteo, settings, variable, dem, proxy_grids=(), quality_settings=None, cross_validate=False,
                          threads=0):
        """Project::interpolationDem for one time step out of reference calls (reference oracle only); mirrors
        Engine.interpolation_dem."""
        assert self.kind == "reference"
        lib = load_ref_extra()
        st = meteo.as_pb()
        d = dem.as_pb()
        grids = (A.pb_raster * max(1, len(proxy_grids)))()
        for i, g in enumerate(proxy_grids):
            grids[i] = g.as_pb()
        out = np.empty(dem.shape, dtype=np.float32)
        o = A.pb_raster_out()
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
        rc = lib.ref_interpolation_dem(C.byref(st), None if quality_settings is None else C.byref(quality_settings),
                                       C.byref(settings), variable, C.byref(d), grids, len(proxy_grids), threads, C.byref(o),
                                       C.byref(step))
        return dict(ok=rc == A.PB_OK, rc=rc, grid=out, min=o.minimum, max=o.maximum, n_valid=o.nValid,
                    wrong_spatial=wrong.astype(bool), residual=res if cross_validate else None,
                    stats={k: getattr(stats, k) for k, _ in A.pb_cv_stats._fields_} if cross_validate else None,
                    kh_series=[(series.kh[i], series.error[i]) for i in range(series.n)])

    def set_topographic_distance_maps(self, point_index=(), maps=()):
        """Crit3DMeteoPoint::topographicDistance of the points (reference only): maps = Raster objects; no arguments = drop all."""
        fn = self.lib.ref_set_topographic_distance_maps
        fn.restype = C.c_int
        n = len(point_index)
        if n == 0:
            fn(None, None, C.c_int(0))
            return
        idx = (C.c_int32 * n)(*[int(i) for i in point_index])
        arr = (A.pb_raster * n)()
        for k, g in enumerate(maps):
            arr[k] = g.as_pb()
        fn(idx, arr, C.c_int(n))

    def interpolate_points(self, stations, settings, variable, x, y, z, proxy_values, exclude_supplemental=False):
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.ascontiguousarray(y, dtype=np.float32)
        z = np.ascontiguousarray(z, dtype=np.float32)
        m = x.shape[0]
        pv = np.ascontiguousarray(proxy_values, dtype=np.float64).reshape(m, -1)
        res = np.empty(m, dtype=np.float32)
        st = stations.as_pb()
        self._f("interpolate_points")(C.byref(st), C.byref(settings), variable, A.fptr(x), A.fptr(y), A.fptr(z),
                                      A.dptr(pv) if pv.size else None, m, int(exclude_supplemental), A.fptr(res))
        return res

    def interpolate_points_td(self, stations, settings, variable, x, y, z, proxy_values, dem, exclude_supplemental=False):
        """interpolate() with the current DEM set (topographic distance available)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.ascontiguousarray(y, dtype=np.float32)
        z = np.ascontiguousarray(z, dtype=np.float32)
        m = x.shape[0]
        pv = np.ascontiguousarray(proxy_values, dtype=np.float64).reshape(m, -1)
        res = np.empty(m, dtype=np.float32)
        st = stations.as_pb()
        d = dem.as_pb()
        self._f("interpolate_points_td")(C.byref(st), C.byref(settings), variable, A.fptr(x), A.fptr(y), A.fptr(z),
                                         A.dptr(pv) if pv.size else None, m, int(exclude_supplemental), C.byref(d), A.fptr(res))
        return res

    def compute_residuals(self, stations, settings, variable, observed, use_for_cv=None, local=False):
        """computeResiduals (local=True: computeResidualsLocalDetrending) + the cross-validation statistics."""
        obs = np.ascontiguousarray(observed, dtype=np.float32)
        res = np.empty(stations.n, dtype=np.float32)
        stats = A.pb_cv_stats()
        st = stations.as_pb()
        use = None if use_for_cv is None else np.ascontiguousarray(use_for_cv, dtype=np.uint8)
        self._f("compute_residuals_local" if local else "compute_residuals")(C.byref(st), C.byref(settings), variable, A.fptr(obs),
                                     None if use is None else A.u8ptr(use), A.fptr(res), C.byref(stats))
        return res, {k: getattr(stats, k) for k, _ in A.pb_cv_stats._fields_}

    def topographic_index(self, dem, widths):
        """topographicIndex (interpolationCmd.cpp:397-482). Returns (ok, grid, min, max)."""
        w = np.ascontiguousarray(widths, dtype=np.float32)
        d = dem.as_pb()
        out = np.empty(dem.shape, dtype=np.float32)
        o = A.pb_raster_out()
        o.data = A.fptr(out)
        rc = self._f("topographic_index")(C.byref(d), A.fptr(w), w.size, C.byref(o))
        return rc == A.PB_OK, out, o.minimum, o.maximum

    def spatial_aggregate_meteo_grid(self, raster, x, y, first, max_nr, method):
        """Crit3DMeteoGrid::spatialAggregateMeteoGrid (meteoGrid.cpp:1135-1190): one value per meteo-grid cell."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        first = np.ascontiguousarray(first, dtype=np.int64)
        max_nr = np.ascontiguousarray(max_nr, dtype=np.int32)
        val = np.empty(max_nr.size, dtype=np.float32)
        r = raster.as_pb()
        self._f("spatial_aggregate_meteo_grid")(C.byref(r), A.dptr(x), A.dptr(y), first.ctypes.data_as(C.POINTER(C.c_int64)),
                                                A.iptr(max_nr), max_nr.size, method, A.fptr(val))
        return val

    # ---- glocal detrending (macro areas) ----
    def glocal_detrending_fitting(self, stations, settings, variable, areas):
        """preInterpolation with useGlocalDetrending (glocalDetrendingFitting, interpolation.cpp:2239-2294): the areas
        (praga_b200._abi.MacroAreas) receive their fit, settings what the reference leaves behind."""
        st = stations.as_pb()
        rc = self._f("glocal_detrending_fitting")(C.byref(st), C.byref(settings), variable, areas.arr, areas.n)
        return rc == A.PB_OK

    def glocal_detrending_raster(self, stations, settings, variable, areas, dem, proxy_grids=(), threads=1, meteo=None,
                                 observed=None):
        """gridding part of Project::interpolationDemGlocalDetrending (project.cpp:3283-3375). meteo (+ observed): Project::meteoPoints,
        which macroAreaDetrending's kh search walks when settings.useTD is on (reference only).
        Returns (ok, grid, min, max, nValid, seconds)."""
        st = stations.as_pb()
        d = dem.as_pb()
        grids = (A.pb_raster * max(1, len(proxy_grids)))()
        for i, g in enumerate(proxy_grids):
            grids[i] = g.as_pb()
        out = np.empty(dem.shape, dtype=np.float32)
        o = A.pb_raster_out()
        o.data = A.fptr(out)
        sec = C.c_double(0)
        if meteo is not None:
            mp = meteo.as_pb()
            gm = A.pb_glocal_meteo()
            gm.meteo = C.pointer(mp)
            obs = None if observed is None else np.ascontiguousarray(observed, dtype=np.float32)
            if obs is not None:
                gm.observed = A.fptr(obs)
            fn = self.lib.ref_glocal_detrending_raster_td
            fn.restype = C.c_int
            rc = fn(C.byref(st), C.byref(settings), C.c_int(variable), areas.arr, C.c_int(areas.n), C.byref(gm), C.byref(d), grids,
                    C.c_int(len(proxy_grids)), C.c_int(threads), C.byref(o), C.byref(sec))
            return rc == A.PB_OK, out, o.minimum, o.maximum, o.nValid, sec.value
        rc = self._f("glocal_detrending_raster")(C.byref(st), C.byref(settings), variable, areas.arr, areas.n, C.byref(d), grids,
                                                 len(proxy_grids), threads, C.byref(o), C.byref(sec))
        return rc == A.PB_OK, out, o.minimum, o.maximum, o.nValid, sec.value

    def compute_residuals_glocal(self, meteo, observed, points, settings, variable, areas, dem, exclude_outside_dem=True,
                                 exclude_supplemental=True):
        """computeResidualsAndStatisticsGlocalDetrending (project.cpp:6525-6565), areas already fitted."""
        obs = np.ascontiguousarray(observed, dtype=np.float32)
        res = np.empty(meteo.n, dtype=np.float32)
        mp, pt, d = meteo.as_pb(), points.as_pb(), dem.as_pb()
        rc = self._f("compute_residuals_glocal")(C.byref(mp), A.fptr(obs), None, C.byref(pt), C.byref(settings), variable,
                                                 areas.arr, areas.n, C.byref(d), int(exclude_outside_dem),
                                                 int(exclude_supplemental), A.fptr(res))
        return rc == A.PB_OK, res

    # ---- raster map operators ----
    def tdew_from_relhum(self, rh, t):
        """tDewFromRelHum(float, float) (meteo.cpp:275-285) element-wise: the stations' dew point of the humidity chain."""
        rh = np.ascontiguousarray(rh, dtype=np.float32)
        t = np.ascontiguousarray(t, dtype=np.float32)
        out = np.empty(rh.shape[0], dtype=np.float32)
        f = self._f("tdew_from_relhum")
        f.restype = None
        f(A.fptr(rh), A.fptr(t), rh.shape[0], A.fptr(out))
        return out

    def compute_radiation_dem(self, settings, dem, transmissivity, utm_zone=32, start_latitude=44.5, threads=0, shim=None):
        """radiation::computeRadiationDEM on the reference's objects (reference oracle only). Returns dict with the static maps the
        reference derives from the DEM (lat, lon, slope, aspect), the five output maps, minmax[10], demMaximum, seconds.
        shim: the loaded libpraga_shim_test.so -> the same driver with pragab200::computeRadiationDEM in place of the reference's."""
        assert self.kind == "reference"
        lib = _load(REF_LIB, "ref")
        P = C.POINTER
        fn = lib.ref_compute_radiation_dem if shim is None else shim.shim_compute_radiation_dem
        fn.argtypes = [P(A.pb_radiation_settings), C.c_int, C.c_double, P(A.pb_raster), P(A.pb_raster)] + \
            [P(C.c_float)] * 11 + [C.c_int, P(C.c_double)]
        d, tr = dem.as_pb(), transmissivity.as_pb()
        names = ("lat", "lon", "slope", "aspect", "sunElevation", "beam", "diffuse", "reflected", "global")
        arrs = {k: np.empty(dem.shape, dtype=np.float32) for k in names}
        mm = np.empty(10, dtype=np.float32)
        dmax = C.c_float(0)
        sec = C.c_double(0)
        rc = fn(C.byref(settings), utm_zone, start_latitude, C.byref(d), C.byref(tr),
                                           *[A.fptr(arrs[k]) for k in names], A.fptr(mm), C.byref(dmax), threads, C.byref(sec))
        arrs.update(ok=rc == A.PB_OK, minmax=mm, dem_maximum=dmax.value, seconds=sec.value)
        return arrs

    def relative_humidity_map(self, air_t, dew_t):
        a, b = air_t.as_pb(), dew_t.as_pb()
        out = np.empty(air_t.shape, dtype=np.float32)
        o = A.pb_raster_out()
        o.data = A.fptr(out)
        rc = self._f("relative_humidity_map")(C.byref(a), C.byref(b), C.byref(o))
        return rc == A.PB_OK, out, o.minimum, o.maximum

    def fix_daily_thermal_consistency(self, tmax, tmin):
        a, b = tmax.as_pb(), tmin.as_pb()
        out = np.empty(tmin.shape, dtype=np.float32)
        n = C.c_int64(0)
        self._f("fix_daily_thermal_consistency")(C.byref(a), C.byref(b), A.fptr(out), C.byref(n))
        return out, n.value

    def topographic_distance_map(self, dem, x, y, z):
        d = dem.as_pb()
        out = np.empty(dem.shape, dtype=np.float32)
        o = A.pb_raster_out()
        o.data = A.fptr(out)
        self._f("topographic_distance_map")(C.byref(d), x, y, z, C.byref(o))
        return out

    def resample_grid(self, src, new_header, method, threshold=0.0):
        s = src.as_pb()
        out = np.empty((new_header.nrRows, new_header.nrCols), dtype=np.float32)
        o = A.pb_raster_out()
        o.data = A.fptr(out)
        rc = self._f("resample_grid")(C.byref(s), C.byref(new_header), method, threshold, C.byref(o))
        return rc == A.PB_OK, out, o.minimum, o.maximum

    # ---- ESRI float grids (reference oracle only) ----
    def write_esri(self, raster, path):
        assert self.kind == "reference"
        r = raster.as_pb()
        return load_ref_io().ref_write_esri(C.byref(r), path.encode()) == A.PB_OK

    def read_esri(self, path):
        """gis::readEsriGrid. Returns (header, grid, min, max) or None."""
        assert self.kind == "reference"
        lib = load_ref_io()
        h = A.pb_raster_header()
        if lib.ref_read_esri_header(path.encode(), C.byref(h)) != A.PB_OK:
            return None
        out = np.empty((h.nrRows, h.nrCols), dtype=np.float32)
        o = A.pb_raster_out()
        o.data = A.fptr(out)
        if lib.ref_read_esri(path.encode(), C.byref(o)) != A.PB_OK:
            return None
        return h, out, o.minimum, o.maximum

    def kriging_points(self, pos, val, mode, rng, nugget, sill, slope, xy, want_rmse=True):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1)
        val = np.ascontiguousarray(val, dtype=np.float64)
        xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1)
        m = xy.shape[0] // 2
        res = np.empty(m, dtype=np.float64)
        rm = np.empty(m, dtype=np.float64)
        t_setup, t_eval = C.c_double(0), C.c_double(0)
        rc = self._f("kriging_points")(A.dptr(pos), A.dptr(val), val.shape[0], mode, rng, nugget, sill, slope,
                                       A.dptr(xy), m, A.dptr(res), A.dptr(rm) if want_rmse else None,
                                       C.byref(t_setup), C.byref(t_eval))
        return rc == A.PB_OK, res, rm, t_setup.value, t_eval.value
