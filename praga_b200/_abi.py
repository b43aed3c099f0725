"""ctypes mirror of include/praga_b200.h (POD descriptors + enum values). Pure declarations, no compute."""
import ctypes as C

import numpy as np

NODATA = -9999
PB_MAX_PROXY = 8
PB_MAX_FIT_PARAMS = 6
PB_MAX_FIRST_GUESS = 64

PB_OK, PB_ERR_INVALID, PB_ERR_CUDA, PB_ERR_NO_VALID_CELL, PB_ERR_NOMEM, PB_ERR_UNSUPPORTED, PB_ERR_FIT = range(7)
STATUS_NAMES = ["PB_OK", "PB_ERR_INVALID", "PB_ERR_CUDA", "PB_ERR_NO_VALID_CELL", "PB_ERR_NOMEM", "PB_ERR_UNSUPPORTED",
                "PB_ERR_FIT"]

# TInterpolationMethod (interpolationConstants.h:18)
idw, shepard, shepard_modified = 0, 1, 2
# TProxyVar (interpolationConstants.h:26)
proxyHeight, proxyUrbanFraction, proxyOrogIndex, proxySeaDistance, proxyAspect, proxySlope, proxyWaterIndex, noProxy = range(8)
# TFittingFunction (interpolationConstants.h:41)
piecewiseTwo, piecewiseThreeFree, piecewiseThree, linear, noFunction = range(5)
# lapseRateCodeType (meteo.h:84)
primary, secondary, supplemental = 0, 1, 2
# TkrigingMode (interpolationConstants.h:51)
KRIGING_SPHERICAL, KRIGING_EXPONENTIAL, KRIGING_GAUSSIAN, KRIGING_LINEAR = 1, 2, 3, 4
# meteoVariable (meteo.h:91-108)
airTemperature = 0
dailyAirTemperatureMin = 1
dailyAirTemperatureMax = 3
dailyAirTemperatureAvg = 5
dailyAirTemperatureRange = 7
precipitation = 8
dailyPrecipitation = 9
airRelHumidity = 11
airDewTemperature = 12
dailyAirRelHumidityMin = 13
dailyAirRelHumidityMax = 14
dailyAirRelHumidityAvg = 15
globalIrradiance = 23
atmTransmissivity = 28
dailyGlobalRadiation = 29
windScalarIntensity = 34
windVectorIntensity = 35
windVectorX = 37
windVectorY = 38
leafWetness = 44
atmPressure = 46
dailyReferenceEvapotranspirationHS = 48
elaborationVar = 72
noMeteoVar = 73


class pb_raster_header(C.Structure):
    _fields_ = [("nrRows", C.c_int32), ("nrCols", C.c_int32), ("cellSize", C.c_double), ("invCellSize", C.c_double),
                ("llx", C.c_double), ("lly", C.c_double), ("flag", C.c_float), ("reserved", C.c_int32)]


class pb_raster(C.Structure):
    _fields_ = [("h", pb_raster_header), ("data", C.POINTER(C.c_float)), ("rows", C.POINTER(C.POINTER(C.c_float)))]


class pb_stations(C.Structure):
    _fields_ = [("n", C.c_int32), ("nProxy", C.c_int32), ("x", C.POINTER(C.c_double)), ("y", C.POINTER(C.c_double)),
                ("z", C.POINTER(C.c_double)), ("value", C.POINTER(C.c_float)),
                ("regressionWeight", C.POINTER(C.c_float)), ("lapseRateCode", C.POINTER(C.c_int32)),
                ("isActive", C.POINTER(C.c_uint8)), ("proxyValues", C.POINTER(C.c_float)),
                ("index", C.POINTER(C.c_int32))]


class pb_proxy(C.Structure):
    _fields_ = [("kind", C.c_int32), ("gridIndex", C.c_int32), ("fittingFunction", C.c_int32),
                ("inversionIsSignificative", C.c_int32),
                ("regressionR2", C.c_float), ("regressionSlope", C.c_float), ("regressionIntercept", C.c_float),
                ("lapseRateH0", C.c_float), ("lapseRateH1", C.c_float), ("lapseRateT0", C.c_float),
                ("lapseRateT1", C.c_float), ("inversionLapseRate", C.c_float), ("stdDevThreshold", C.c_float),
                ("nFitParams", C.c_int32),
                ("fitMin", C.c_double * PB_MAX_FIT_PARAMS), ("fitMax", C.c_double * PB_MAX_FIT_PARAMS),
                ("fitFirstGuess", C.c_int32 * PB_MAX_FIT_PARAMS), ("nFirstGuessCombinations", C.c_int32),
                ("firstGuessCombinations", (C.c_double * PB_MAX_FIT_PARAMS) * PB_MAX_FIRST_GUESS)]


class pb_settings(C.Structure):
    _fields_ = [("method", C.c_int32),
                ("useThermalInversion", C.c_int32), ("useTD", C.c_int32), ("useLocalDetrending", C.c_int32),
                ("useGlocalDetrending", C.c_int32), ("useLapseRateCode", C.c_int32), ("useBestDetrending", C.c_int32),
                ("useMultipleDetrending", C.c_int32), ("useDoNotRetrend", C.c_int32), ("useRetrendOnly", C.c_int32),
                ("precipitationAllZero", C.c_int32),
                ("minPointsLocalDetrending", C.c_int32), ("topoDist_Kh", C.c_int32), ("topoDist_maxKh", C.c_int32),
                ("minRegressionR2", C.c_float), ("maxHeightInversion", C.c_float), ("pointsBoundingBoxArea", C.c_float),
                ("localRadius", C.c_float), ("rainfallThreshold", C.c_float), ("climateLapseRate", C.c_float),
                ("hasPointsRange", C.c_int32), ("pointsRangeMin", C.c_double), ("pointsRangeMax", C.c_double),
                ("nProxy", C.c_int32), ("selectedUseThermalInversion", C.c_int32),
                ("currentUseThermalInversion", C.c_int32),
                ("selectedActive", C.c_uint8 * PB_MAX_PROXY), ("currentActive", C.c_uint8 * PB_MAX_PROXY),
                ("currentSignificant", C.c_uint8 * PB_MAX_PROXY),
                ("proxy", pb_proxy * PB_MAX_PROXY),
                ("nFit", C.c_int32), ("nFitFunctions", C.c_int32),
                ("fitFunction", C.c_int32 * PB_MAX_PROXY), ("fitNrParams", C.c_int32 * PB_MAX_PROXY),
                ("fitParams", (C.c_double * PB_MAX_FIT_PARAMS) * PB_MAX_PROXY)]


class pb_raster_out(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_float)), ("rows", C.POINTER(C.POINTER(C.c_float))),
                ("minimum", C.c_float), ("maximum", C.c_float), ("nValid", C.c_int64)]


class pb_cv_stats(C.Structure):
    _fields_ = [("n", C.c_int32), ("meanAbsoluteError", C.c_float), ("meanBiasError", C.c_float),
                ("rootMeanSquareError", C.c_float), ("nashSutcliffe", C.c_float), ("R2", C.c_float)]


class pb_cv_points(C.Structure):
    _fields_ = [("currentValue", C.POINTER(C.c_float)), ("isInsideDem", C.POINTER(C.c_uint8)),
                ("excludeOutsideDem", C.c_int32), ("reserved", C.c_int32), ("suspectProxyValues", C.POINTER(C.c_float))]


class pb_glocal_meteo(C.Structure):
    _fields_ = [("meteo", C.POINTER(pb_stations)), ("observed", C.POINTER(C.c_float)), ("cv", C.POINTER(pb_cv_points))]


PB_AGGR_AVERAGE, PB_AGGR_MEDIAN, PB_AGGR_STDDEV, PB_AGGR_PREVAILING, PB_AGGR_CENTER = 1, 2, 3, 7, 9
PB_MAX_KH_SERIES = 64


class pb_kh_series(C.Structure):
    _fields_ = [("n", C.c_int32), ("kh", C.c_float * PB_MAX_KH_SERIES), ("error", C.c_float * PB_MAX_KH_SERIES)]


class pb_dem_step(C.Structure):
    _fields_ = [("wrongSpatial", C.POINTER(C.c_uint8)), ("residual", C.POINTER(C.c_float)),
                ("stats", C.POINTER(pb_cv_stats)), ("khSeries", C.POINTER(pb_kh_series))]


class pb_dem_batch_item(C.Structure):
    _fields_ = [("meteo", C.POINTER(pb_stations)), ("qualitySettings", C.POINTER(pb_settings)), ("settings", C.POINTER(pb_settings)),
                ("variable", C.c_int32), ("status", C.c_int32), ("out", C.POINTER(pb_raster_out)), ("step", C.POINTER(pb_dem_step))]


class pb_meteo_grid_var(C.Structure):
    _fields_ = [("meteo", C.POINTER(pb_stations)), ("qualitySettings", C.POINTER(pb_settings)), ("settings", C.POINTER(pb_settings)),
                ("variable", C.c_int32), ("useDewPoint", C.c_int32), ("airTemperature", C.POINTER(C.c_float)),
                ("aggrMethod", C.c_int32), ("status", C.c_int32), ("cellValue", C.POINTER(C.c_float)),
                ("out", C.POINTER(pb_raster_out)), ("step", C.POINTER(pb_dem_step))]


class pb_meteo_grid_hour(C.Structure):
    _fields_ = [("vars", C.POINTER(pb_meteo_grid_var)), ("nVars", C.c_int32), ("reserved", C.c_int32)]


class pb_radiation_settings(C.Structure):
    _fields_ = [("realSky", C.c_int32), ("realSkyAlgorithm", C.c_int32), ("shadowing", C.c_int32), ("tiltMode", C.c_int32),
                ("linkeMode", C.c_int32), ("albedoMode", C.c_int32), ("linke", C.c_float), ("albedo", C.c_float), ("tilt", C.c_float),
                ("aspect", C.c_float), ("clearSky", C.c_float), ("demMaximum", C.c_float), ("timeZone", C.c_int32),
                ("year", C.c_int32), ("month", C.c_int32), ("day", C.c_int32), ("secondsOfDay", C.c_int32)]


class pb_radiation_maps(C.Structure):
    _fields_ = [("dem", C.c_int32), ("lat", C.c_int32), ("lon", C.c_int32), ("slope", C.c_int32), ("aspect", C.c_int32),
                ("transmissivity", C.c_int32)]


class pb_radiation_out(C.Structure):
    _fields_ = [("sunElevation", C.POINTER(pb_raster_out)), ("beam", C.POINTER(pb_raster_out)), ("diffuse", C.POINTER(pb_raster_out)),
                ("reflected", C.POINTER(pb_raster_out)), ("global_", C.POINTER(pb_raster_out)), ("keepResident", C.c_int32),
                ("ids", C.c_int32 * 5)]


def default_radiation_settings(year, month, day, seconds_of_day, time_zone=1):
    """Crit3DRadiationSettings::initialize() defaults (radiationSettings.cpp:41-72) + the local time of the step"""
    s = pb_radiation_settings()
    s.realSky, s.realSkyAlgorithm, s.shadowing, s.tiltMode = 1, 1, 1, 2
    s.linke, s.albedo, s.clearSky = 4.0, 0.2, 0.75
    s.timeZone, s.year, s.month, s.day, s.secondsOfDay = time_zone, year, month, day, seconds_of_day
    return s


class pb_macro_area(C.Structure):
    _fields_ = [("nMeteoPoints", C.c_int32), ("reserved", C.c_int32), ("meteoPoints", C.POINTER(C.c_int32)),
                ("nAreaCells", C.c_int64), ("areaCellsDEM", C.POINTER(C.c_float)),
                ("active", C.c_uint8 * PB_MAX_PROXY), ("significant", C.c_uint8 * PB_MAX_PROXY),
                ("useThermalInversion", C.c_int32), ("nParameters", C.c_int32), ("nrParams", C.c_int32 * PB_MAX_PROXY),
                ("parameters", (C.c_double * PB_MAX_FIT_PARAMS) * PB_MAX_PROXY)]


class MacroAreas:
    """std::vector<Crit3DMacroArea> (interpolationSettings.h:149-182): per area the station indices and the flat
    (cell index, weight) float pairs; keeps the numpy buffers alive behind the pb_macro_area array."""

    def __init__(self, meteo_points, area_cells):
        assert len(meteo_points) == len(area_cells)
        self.meteo_points = [np.ascontiguousarray(m, dtype=np.int32) for m in meteo_points]
        self.area_cells = [np.ascontiguousarray(c, dtype=np.float32).ravel() for c in area_cells]
        self.n = len(self.meteo_points)
        self.arr = (pb_macro_area * max(1, self.n))()
        for i in range(self.n):
            a = self.arr[i]
            a.nMeteoPoints = self.meteo_points[i].size
            a.meteoPoints = iptr(self.meteo_points[i]) if self.meteo_points[i].size else None
            a.nAreaCells = self.area_cells[i].size
            a.areaCellsDEM = fptr(self.area_cells[i]) if self.area_cells[i].size else None

    def fit(self, i):
        """(active, significant, parameters) of area i as the last fitting call left them."""
        a = self.arr[i]
        pars = [[a.parameters[k][j] for j in range(a.nrParams[k])] for k in range(a.nParameters)]
        return list(a.active), list(a.significant), pars

    def fits(self):
        return [self.fit(i) for i in range(self.n)]


class pb_timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("kernel_ms", C.c_float), ("d2h_ms", C.c_float), ("total_ms", C.c_float),
                ("launches", C.c_int32), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64)]


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def u8ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def default_settings():
    """Crit3DInterpolationSettings::initialize() defaults (interpolationSettings.cpp:343-382)."""
    s = pb_settings()
    s.method = idw
    s.useThermalInversion = 1
    s.topoDist_maxKh = 128
    s.minRegressionR2 = 0.1
    s.maxHeightInversion = 1000.0
    s.minPointsLocalDetrending = 20
    s.rainfallThreshold = 0.2
    s.climateLapseRate = -0.006
    s.localRadius = 0.0
    return s


def default_proxy(kind, grid_index=-1):
    """Crit3DProxy() defaults (interpolationSettings.cpp:692-717)."""
    p = pb_proxy()
    p.kind = kind
    p.gridIndex = grid_index
    p.fittingFunction = noFunction
    for f in ("regressionR2", "regressionSlope", "lapseRateH0", "lapseRateH1", "lapseRateT0", "lapseRateT1",
              "inversionLapseRate", "stdDevThreshold"):
        setattr(p, f, NODATA)
    return p


class Raster:
    """Host raster: header + contiguous float32 array (row 0 = north), the flat twin of gis::Crit3DRasterGrid."""

    def __init__(self, data, llx, lly, cell_size, flag=-9999.0):
        self.data = np.ascontiguousarray(data, dtype=np.float32)
        assert self.data.ndim == 2
        self.llx, self.lly, self.cell_size, self.flag = float(llx), float(lly), float(cell_size), float(flag)

    @property
    def shape(self):
        return self.data.shape

    def header(self):
        h = pb_raster_header()
        h.nrRows, h.nrCols = self.data.shape
        h.cellSize = self.cell_size
        h.invCellSize = 1.0 / self.cell_size       # gis.cpp:351, gisIO.cpp:183
        h.llx, h.lly, h.flag = self.llx, self.lly, self.flag
        return h

    def as_pb(self):
        r = pb_raster()
        r.h = self.header()
        r.data = fptr(self.data)
        return r


class Stations:
    """SoA of std::vector<Crit3DInterpolationDataPoint> (interpolationPoint.h:14-31)."""

    def __init__(self, x, y, z, value, proxy_values=None, lapse_rate_code=None, is_active=None, regression_weight=None,
                 index=None):
        self.x = np.ascontiguousarray(x, dtype=np.float64)
        self.y = np.ascontiguousarray(y, dtype=np.float64)
        self.z = np.ascontiguousarray(z, dtype=np.float64)
        self.value = np.ascontiguousarray(value, dtype=np.float32).copy()
        n = self.x.shape[0]
        if proxy_values is None:
            proxy_values = np.zeros((n, 0), dtype=np.float32)
        pv = np.ascontiguousarray(proxy_values, dtype=np.float32)
        self.proxy_values = pv.reshape(n, pv.shape[1] if pv.ndim == 2 else (pv.size // n if n else 0))
        self.lapse_rate_code = None if lapse_rate_code is None else np.ascontiguousarray(lapse_rate_code, dtype=np.int32)
        self.is_active = None if is_active is None else np.ascontiguousarray(is_active, dtype=np.uint8)
        self.regression_weight = (None if regression_weight is None
                                  else np.ascontiguousarray(regression_weight, dtype=np.float32))
        # Crit3DInterpolationDataPoint::index (the meteo point behind the interpolation point); None = the position
        self.index = None if index is None else np.ascontiguousarray(index, dtype=np.int32)

    @property
    def n(self):
        return int(self.x.shape[0])

    def copy(self):
        return Stations(self.x, self.y, self.z, self.value.copy(), self.proxy_values, self.lapse_rate_code,
                        self.is_active, self.regression_weight, self.index)

    def subset(self, keep):
        keep = np.asarray(keep, dtype=bool)
        return Stations(self.x[keep], self.y[keep], self.z[keep], self.value[keep], self.proxy_values[keep],
                        None if self.lapse_rate_code is None else self.lapse_rate_code[keep],
                        None if self.is_active is None else self.is_active[keep],
                        None if self.regression_weight is None else self.regression_weight[keep],
                        None if self.index is None else self.index[keep])

    def as_pb(self):
        st = pb_stations()
        st.n = self.n
        st.nProxy = self.proxy_values.shape[1]
        st.x, st.y, st.z = dptr(self.x), dptr(self.y), dptr(self.z)
        st.value = fptr(self.value)
        if self.regression_weight is not None:
            st.regressionWeight = fptr(self.regression_weight)
        if self.lapse_rate_code is not None:
            st.lapseRateCode = iptr(self.lapse_rate_code)
        if self.is_active is not None:
            st.isActive = u8ptr(self.is_active)
        if self.proxy_values.size:
            st.proxyValues = fptr(self.proxy_values)
        if self.index is not None:
            self.index = np.ascontiguousarray(self.index, dtype=np.int32)
            st.index = iptr(self.index)
        return st
