/*
 * praga_b200.h — C ABI of libpraga_b200.so, the B200 (sm_100a) gridding engine that replaces the
 * hot path of ARPA-SIMC/PRAGA's agrolib/interpolation behind the reference's own C++ API.
 *
 * The reference has no FFI for this path: it is reached through statically linked C++ free
 * functions.  Every entry point below names the reference interface it replaces
 * (paths relative to the reference tree).  The C++ shim in praga_b200/host/ keeps the reference
 * signatures (interpolationRaster(), preInterpolation(), computeResiduals(), kriging*) and only
 * flattens the reference objects into the POD descriptors declared here.
 *
 * Conventions
 *  - plain pointers + sizes, no C++ types, no exceptions across the ABI;
 *  - every function returns a pb_status (0 = ok); pb_last_error() gives the message;
 *  - all pointers are HOST pointers unless the name says "device"; the caller owns all of them;
 *  - rasters are row-major, row 0 = northernmost row, exactly like gis::Crit3DRasterGrid::value
 *    (agrolib/gis/gis.h:157-166); either one contiguous block (`data`) or the reference's
 *    per-row allocation (`rows`, a float**) may be passed;
 *  - there is NO CPU fallback: without a usable CUDA device every compute call fails with
 *    PB_ERR_CUDA.
 */
#ifndef PRAGA_B200_H
#define PRAGA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB_NODATA (-9999)            /* agrolib/mathFunctions/commonConstants.h:31 */
#define PB_MAX_PROXY 8
#define PB_MAX_FIT_PARAMS 6          /* lapseRatePiecewise_three_free has 6 (furtherMathFunctions.cpp:149) */
#define PB_MAX_FIRST_GUESS 64        /* calculateFirstGuessCombinations yields 16 or 31 (interpolation.cpp:1556) */

typedef enum pb_status {
    PB_OK = 0,
    PB_ERR_INVALID = 1,        /* bad argument / inconsistent descriptor */
    PB_ERR_CUDA = 2,           /* no device, launch or copy failure (reference-side: return false) */
    PB_ERR_NO_VALID_CELL = 3,  /* gis::updateMinMaxRasterGrid returned false (gis.cpp:601-603) */
    PB_ERR_NOMEM = 4,
    PB_ERR_UNSUPPORTED = 5,    /* option combination outside the hot-path scope */
    PB_ERR_FIT = 6             /* preInterpolation returned false (interpolation.cpp:2476) */
} pb_status;

/* --- enums copied BY VALUE from the reference (checked by static_assert in oracle/ref_driver.cpp) --- */
/* TInterpolationMethod, agrolib/interpolation/interpolationConstants.h:18 */
enum { PB_METHOD_IDW = 0, PB_METHOD_SHEPARD = 1, PB_METHOD_SHEPARD_MODIFIED = 2 };
/* TProxyVar, interpolationConstants.h:26 */
enum { PB_PROXY_HEIGHT = 0, PB_PROXY_URBAN_FRACTION = 1, PB_PROXY_OROG_INDEX = 2, PB_PROXY_SEA_DISTANCE = 3,
       PB_PROXY_ASPECT = 4, PB_PROXY_SLOPE = 5, PB_PROXY_WATER_INDEX = 6, PB_PROXY_NONE = 7 };
/* TFittingFunction, interpolationConstants.h:41 */
enum { PB_FIT_PIECEWISE_TWO = 0, PB_FIT_PIECEWISE_THREE_FREE = 1, PB_FIT_PIECEWISE_THREE = 2, PB_FIT_LINEAR = 3,
       PB_FIT_NONE = 4 };
/* lapseRateCodeType, agrolib/meteo/meteo.h:84 */
enum { PB_LAPSE_PRIMARY = 0, PB_LAPSE_SECONDARY = 1, PB_LAPSE_SUPPLEMENTAL = 2 };
/* TkrigingMode, interpolationConstants.h:51 */
enum { PB_KRIGING_SPHERICAL = 1, PB_KRIGING_EXPONENTIAL = 2, PB_KRIGING_GAUSSIAN = 3, PB_KRIGING_LINEAR = 4 };
/* meteoVariable, agrolib/meteo/meteo.h:91-108 (only the values the gridding path distinguishes) */
enum {
    PB_VAR_AIR_TEMPERATURE = 0, PB_VAR_DAILY_TMIN = 1, PB_VAR_DAILY_TMAX = 3, PB_VAR_DAILY_TAVG = 5,
    PB_VAR_DAILY_TRANGE = 7, PB_VAR_PRECIPITATION = 8, PB_VAR_DAILY_PRECIPITATION = 9,
    PB_VAR_AIR_REL_HUMIDITY = 11, PB_VAR_AIR_DEW_TEMPERATURE = 12, PB_VAR_DAILY_RH_MIN = 13,
    PB_VAR_DAILY_RH_MAX = 14, PB_VAR_DAILY_RH_AVG = 15, PB_VAR_GLOBAL_IRRADIANCE = 23,
    PB_VAR_ATM_TRANSMISSIVITY = 28, PB_VAR_DAILY_GLOBAL_RADIATION = 29, PB_VAR_WIND_SCALAR_INTENSITY = 34,
    PB_VAR_WIND_VECTOR_INTENSITY = 35, PB_VAR_WIND_VECTOR_X = 37, PB_VAR_WIND_VECTOR_Y = 38,
    PB_VAR_DAILY_WIND_VECTOR_INT_AVG = 39, PB_VAR_DAILY_WIND_VECTOR_INT_MAX = 40,
    PB_VAR_DAILY_WIND_SCALAR_INT_AVG = 42, PB_VAR_DAILY_WIND_SCALAR_INT_MAX = 43, PB_VAR_LEAF_WETNESS = 44,
    PB_VAR_DAILY_LEAF_WETNESS = 45, PB_VAR_ATM_PRESSURE = 46, PB_VAR_DAILY_ET0_HS = 48, PB_VAR_DAILY_BIC = 52,
    PB_VAR_DAILY_WATER_TABLE_DEPTH = 67, PB_VAR_ELABORATION = 72, PB_VAR_NONE = 73
};

/* gis::Crit3DRasterHeader, agrolib/gis/gis.h:112-127 */
typedef struct pb_raster_header {
    int32_t nrRows, nrCols;
    double cellSize, invCellSize;
    double llx, lly;               /* llCorner.x / llCorner.y (lower-left corner, metres UTM) */
    float flag;                    /* no-data marker of THIS raster */
    int32_t reserved;
} pb_raster_header;

/* host view of a gis::Crit3DRasterGrid (gis.h:157-199): give `data` (contiguous) or `rows` (float**) */
typedef struct pb_raster {
    pb_raster_header h;
    const float* data;             /* nrRows*nrCols contiguous, or NULL */
    float* const* rows;            /* Crit3DRasterGrid::value, used when data == NULL */
} pb_raster;

/* std::vector<Crit3DInterpolationDataPoint> (agrolib/interpolation/interpolationPoint.h:14-31), SoA */
typedef struct pb_stations {
    int32_t n;                     /* number of points */
    int32_t nProxy;                /* row stride of proxyValues == settings.nProxy */
    const double* x;               /* point->utm.x */
    const double* y;               /* point->utm.y */
    const double* z;               /* point->z */
    float* value;                  /* value (in/out for pb_pre_interpolation: detrended in place) */
    const float* regressionWeight; /* may be NULL => all NODATA */
    const int32_t* lapseRateCode;  /* PB_LAPSE_*; may be NULL => all primary */
    const uint8_t* isActive;       /* may be NULL => all active */
    const float* proxyValues;      /* n*nProxy, station-major; NODATA where missing */
    const int32_t* index;          /* Crit3DInterpolationDataPoint::index; may be NULL */
} pb_stations;

/* Crit3DProxy (interpolationSettings.h:18-102) — fields the path reads or writes */
typedef struct pb_proxy {
    int32_t kind;                  /* PB_PROXY_* = getProxyPragaName(name) */
    int32_t gridIndex;             /* index into the proxyGrids[] argument; -1: no loaded grid */
    int32_t fittingFunction;       /* PB_FIT_* (fittingFunctionName) */
    int32_t inversionIsSignificative;
    float regressionR2, regressionSlope, regressionIntercept;
    float lapseRateH0, lapseRateH1, lapseRateT0, lapseRateT1, inversionLapseRate;
    float stdDevThreshold;
    int32_t nFitParams;            /* fittingParametersRange.size()/2 */
    double fitMin[PB_MAX_FIT_PARAMS];   /* fittingParametersRange[0..n)   */
    double fitMax[PB_MAX_FIT_PARAMS];   /* fittingParametersRange[n..2n)  */
    int32_t fitFirstGuess[PB_MAX_FIT_PARAMS];
    int32_t nFirstGuessCombinations;    /* set by setMultipleDetrendingHeightTemperatureRange */
    double firstGuessCombinations[PB_MAX_FIRST_GUESS][PB_MAX_FIT_PARAMS];
} pb_proxy;

/* Crit3DInterpolationSettings (interpolationSettings.h:185-391) + the two scalars the path reads from
 * Crit3DMeteoSettings (rainfallThreshold) — in/out where the reference mutates its settings */
typedef struct pb_settings {
    int32_t method;                /* PB_METHOD_* */
    int32_t useThermalInversion, useTD, useLocalDetrending, useGlocalDetrending;
    int32_t useLapseRateCode, useBestDetrending, useMultipleDetrending;
    int32_t useDoNotRetrend, useRetrendOnly, precipitationAllZero;
    int32_t minPointsLocalDetrending, topoDist_Kh, topoDist_maxKh;
    float minRegressionR2, maxHeightInversion, pointsBoundingBoxArea, localRadius;
    float rainfallThreshold;       /* Crit3DMeteoSettings::getRainfallThreshold (meteo.h:48) */
    float climateLapseRate;        /* Crit3DClimateParameters::getClimateLapseRate(var, time) (meteo.cpp:184) */
    int32_t hasPointsRange;
    double pointsRangeMin, pointsRangeMax;
    int32_t nProxy;
    int32_t selectedUseThermalInversion, currentUseThermalInversion;
    uint8_t selectedActive[PB_MAX_PROXY];      /* selectedCombination._isActiveList */
    uint8_t currentActive[PB_MAX_PROXY];       /* currentCombination */
    uint8_t currentSignificant[PB_MAX_PROXY];
    pb_proxy proxy[PB_MAX_PROXY];
    /* fitted multiple-detrending model: fittingFunction / fittingParameters, elevation first
     * (interpolation.cpp:1298-1313, 2563-2617) */
    int32_t nFit;                              /* fittingParameters.size() */
    int32_t nFitFunctions;                     /* fittingFunction.size() (can exceed nFit, interpolation.cpp:1506-1553) */
    int32_t fitFunction[PB_MAX_PROXY];         /* PB_FIT_* */
    int32_t fitNrParams[PB_MAX_PROXY];
    double fitParams[PB_MAX_PROXY][PB_MAX_FIT_PARAMS];
} pb_settings;

/* output of a raster call: gis::Crit3DRasterGrid* outputGrid after interpolationRaster */
typedef struct pb_raster_out {
    float* data;                   /* nrRows*nrCols contiguous, or NULL */
    float* const* rows;            /* row pointers (Crit3DRasterGrid::value), used when data == NULL */
    float minimum, maximum;        /* gis::updateMinMaxRasterGrid (gis.cpp:569-614) */
    int64_t nValid;                /* cells with DEM != flag (the metric's "gridded cells") */
} pb_raster_out;

/* cross-validation statistics: Project::computeStatisticsCrossValidation (project/project.cpp:2935-2969) */
typedef struct pb_cv_stats {
    int32_t n;
    float meanAbsoluteError, meanBiasError, rootMeanSquareError, nashSutcliffe, R2;
} pb_cv_stats;

typedef struct pb_context pb_context;

/* --- context ------------------------------------------------------------------------------------ */
int pb_device_count(void);
int pb_create(int device, pb_context** out);
void pb_destroy(pb_context* ctx);
const char* pb_last_error(const pb_context* ctx);   /* ctx may be NULL: last creation error */
const char* pb_version(void);
size_t pb_abi_sizeof(const char* structName);       /* sizeof of a descriptor in THIS build (binding self-check) */
/* run on a caller-owned CUDA stream (a cudaStream_t passed as void*), e.g. torch's current stream */
int pb_set_stream(pb_context* ctx, void* cudaStream);
/* how the host threads of this process wait for the device (cudaSetDeviceFlags): blocking != 0 = sleep on a synchronisation
 * primitive instead of spinning. For hosts that run more waiting threads (ranks x lanes of the batch drivers) than cores. */
int pb_set_blocking_sync(pb_context* ctx, int blocking);

/* --- device-resident raster cache (DEM / proxy grids are reused by every time step) -------------- */
typedef int32_t pb_raster_id;
int pb_raster_upload(pb_context* ctx, const pb_raster* r, pb_raster_id* id);
int pb_raster_free(pb_context* ctx, pb_raster_id id);
/* number of cells with !isEqual(value, flag) (interpolationCmd.cpp:377): the metric's "gridded cells" */
int pb_raster_valid_cells(pb_context* ctx, pb_raster_id id, int64_t* nValid);

/* Optional: page-lock a long-lived CONTIGUOUS host buffer of the caller (a DEM or proxy grid given as pb_raster.data, an
 * output grid given as pb_raster_out.data). The host-raster calls below then copy straight from / into it (no staging copy by
 * the host cores): the drop-in call is otherwise bound by that memcpy, not by PCIe. The caller must unregister the buffer
 * before freeing it. Buffers of the reference's per-row layout (float**) always go through the staging ring. */
int pb_host_register(pb_context* ctx, void* ptr, size_t bytes);
int pb_host_unregister(pb_context* ctx, void* ptr);

/* --- interpolationRaster (agrolib/project/interpolationCmd.h:73-75, .cpp:353-394) ---------------
 * stations are the already pre-interpolated (detrended) points, settings carry the fitted model.
 * `variable` is the reference meteoVariable value. rowBegin/rowEnd select a row band (multi-GPU
 * sharding; 0,-1 = whole raster); cells outside the band are left untouched in `out`.
 * Host-raster form (uploads, computes, downloads; the drop-in call): */
int pb_interpolation_raster(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable,
                            const pb_raster* dem, const pb_raster* proxyGrids, int nProxyGrids,
                            int rowBegin, int rowEnd, pb_raster_out* out);
/* resident form: rasters were uploaded once with pb_raster_upload */
int pb_interpolation_raster_resident(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable,
                                     pb_raster_id dem, const pb_raster_id* proxyGrids, int nProxyGrids,
                                     int rowBegin, int rowEnd, pb_raster_out* out);
/* device-only step (no D2H of the grid; min/max/nValid still returned): used by bench `value` */
int pb_interpolation_raster_device(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable,
                                   pb_raster_id dem, const pb_raster_id* proxyGrids, int nProxyGrids,
                                   int rowBegin, int rowEnd, pb_raster_out* out);

/* --- preInterpolation (agrolib/interpolation/interpolation.cpp:2444-2499): the per-time-step station-space fit.
 * GPU branch: multipleDetrendingMain (:1832-1859; weighted Levenberg-Marquardt fits of the elevation function from
 * 16-31 first guesses and of the other proxies) and checkPrecipitationZero (:1173). One time step cannot fill a GPU, so
 * the batched form runs T independent time steps (the reference's hourly/daily batch drivers) that share the station set
 * st (positions, proxies, codes): values = T x st->n; pointsRange = T x {min, max} as setPointsRange would receive
 * (NULL: float min/max of each row, spatialControl.cpp:533-564); outputs: detrended T x n, keep T x n (1 = the point
 * is still in myPoints, interpolation.cpp:2230), fitted[T] = the settings as the reference leaves them. */
int pb_pre_interpolation_batch(pb_context* ctx, const pb_stations* st, const float* values, int T, const pb_settings* s,
                               int variable, const double* pointsRange, float* detrended, uint8_t* keep, pb_settings* fitted);
/* one time step, in place: st->value is detrended, *s receives the fit (same contract as the reference call) */
int pb_pre_interpolation(pb_context* ctx, pb_stations* st, pb_settings* s, int variable, uint8_t* keep);

/* preInterpolation with EVERY branch of interpolation.cpp:2444-2499, one time step, in place:
 *   - checkPrecipitationZero :1173; multipleDetrendingMain :1832 (as above);
 *   - classic detrending() :1380-1415 (regressionOrographyT :433-798 thermal-inversion heuristic, regressionSimpleT :368,
 *     regressionGeneric :346, detrendPoints :1236), one device lane per (time step, proxy combination);
 *   - optimalDetrending() :2395-2441 when settings.useBestDetrending: every combination of getCombination
 *     (interpolationSettings.cpp:869-892) is detrended, cross-validated (computeResiduals, spatialControl.cpp:97-155, exact
 *     f32-distance / f64-weight leave-one-out on the device) and the smallest MAE wins;
 *   - topographicDistanceOptimize() :2356-2392 when getUseTD() && getUseTdVar(var): golden section on kh over the
 *     MAE(kh) table of all integer kh in [0, topoDist_maxKh], built in one launch from the station-to-station topographic
 *     distances (gis::topographicDistance, gis/gis.cpp:1595-1647, marched through the resident DEM `dem`).
 * The reference walks std::vector<Crit3DMeteoPoint> for the cross-validation; here there is one meteo point per station of
 * `st` in the same order (what passDataToInterpolation, spatialControl.cpp:500-566, builds the points from), described by
 * pb_cv_points (NULL: currentValue = st->value at entry, every point inside the DEM). settings receives what the reference
 * leaves behind: current combination and significance, the proxies' regression fields, topoDist_Kh, points range /
 * bounding-box area (optimalDetrending re-runs passDataToInterpolation); khSeries (may be NULL) the Kh / error series. */
typedef struct pb_cv_points {
    const float* currentValue;     /* Crit3DMeteoPoint::currentValue; NULL => st->value at entry */
    const uint8_t* isInsideDem;    /* Crit3DMeteoPoint::isInsideDem; NULL => all inside */
    int32_t excludeOutsideDem;     /* getUseExcludeStationsOutsideDEM() */
    int32_t reserved;
    /* pb_spatial_quality_control only: n * nProxy values, row k = what the reference passes for the k-th SUSPECT of the
     * second pass: meteoPoints[k].getProxyValues() with k counted over the caller's FULL meteo-point vector, not the
     * suspect's own entry (spatialControl.cpp:393). NULL => st->proxyValues (every meteo point active and accepted). */
    const float* suspectProxyValues;
} pb_cv_points;
#define PB_MAX_KH_SERIES 64
typedef struct pb_kh_series {      /* Crit3DInterpolationSettings::Kh_series / Kh_error_series (interpolationSettings.cpp:134) */
    int32_t n;
    float kh[PB_MAX_KH_SERIES], error[PB_MAX_KH_SERIES];
} pb_kh_series;
int pb_pre_interpolation_ex(pb_context* ctx, pb_stations* st, pb_settings* s, int variable, const pb_cv_points* cv,
                            pb_raster_id dem /* -1: none */, uint8_t* keep, pb_kh_series* khSeries);

/* --- local-detrending raster: the loop of Project::interpolationDemLocalDetrending (agrolib/project/project.cpp:3208-3253),
 * which the reference inlines in its Qt-bound Project class instead of going through interpolationRaster. Per valid
 * cell: localSelection (interpolation.cpp:1087-1170), preInterpolation on the subset (multipleDetrendingMain :1832,
 * weighted Levenberg-Marquardt fits), interpolate (idw, or modifiedShepardIdw with the local radius when method =
 * shepard_modified) and retrend. `st` holds the NOT yet detrended points of
 * checkAndPassDataToInterpolation in their original order; `s` needs useLocalDetrending, useMultipleDetrending,
 * minPointsLocalDetrending, the selected combination, the proxies' fitting ranges and the points range
 * (setPointsRange, spatialControl.cpp:564). Same three forms as interpolationRaster. */
int pb_local_detrending_raster(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable,
                               const pb_raster* dem, const pb_raster* proxyGrids, int nProxyGrids,
                               int rowBegin, int rowEnd, pb_raster_out* out);
int pb_local_detrending_raster_resident(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable,
                                        pb_raster_id dem, const pb_raster_id* proxyGrids, int nProxyGrids,
                                        int rowBegin, int rowEnd, pb_raster_out* out);
int pb_local_detrending_raster_device(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable,
                                      pb_raster_id dem, const pb_raster_id* proxyGrids, int nProxyGrids,
                                      int rowBegin, int rowEnd, pb_raster_out* out);

/* --- glocal detrending: one multiple-detrending fit per MACRO AREA, cells blended over the areas they belong to.
 * pb_macro_area mirrors Crit3DMacroArea (agrolib/interpolation/interpolationSettings.h:149-182): meteoPoints = the
 * Crit3DInterpolationDataPoint::index values of the area's stations (matched against st->index, or the position when that is
 * NULL), areaCellsDEM = the reference's flat (cell index = row * nrCols + col, weight) pairs as floats (cell indices above
 * 2^24 are not exact in that format; the reference has the same limit). The fit part is in/out.
 *   pb_glocal_detrending_fitting <- glocalDetrendingFitting (agrolib/interpolation/interpolation.cpp:2239-2294) as
 *       preInterpolation :2471-2474 reaches it (after setMultipleDetrendingHeightTemperatureRange): per area with stations,
 *       the UNWEIGHTED elevation fit (bestFittingMarquardt_nDimension, mathFunctions/furtherMathFunctions.cpp:1433-1529, over
 *       fittingMarquardt_nDimension :2125-2283), detrendingElevation, the other proxies' fit; the area keeps parameters and
 *       combination. All areas are fitted in ONE launch (one lane group per area). st = the NOT detrended points; *s receives
 *       what the reference leaves in the settings (the last fitted area's combination / parameters, the height ranges).
 *   pb_glocal_detrending_raster  <- the loop of Project::interpolationDemGlocalDetrending (agrolib/project/project.cpp:3262-3375)
 *       after checkAndPassDataToInterpolation: preInterpolation (= the fitting above), then per area macroAreaDetrending
 *       (interpolation.cpp:1759-1829: the area's stations detrended with the area's fit), and for each of its cells
 *       getUtmXYFromRowCol (cell centre, gis/gis.cpp:806-810), getSignificantProxyValuesXY (:2650-2677; incomplete => the
 *       cell is set to NODATA), interpolate() over the area's stations, out[cell] (+)= value * weight. Areas are applied in
 *       index order (the reference's order when isParallelComputing is off; with OpenMP it is unordered, same sum up to
 *       float rounding). Returns PB_ERR_NO_VALID_CELL where the reference returns false (an interpolate() gave NODATA); the
 *       grid is still written. out->minimum / maximum as gis::updateMinMaxRasterGrid would set them.
 *   pb_glocal_detrending_raster_td <- the same with settings.useTD: macroAreaDetrending then ends with
 *       topographicDistanceOptimize over the AREA's meteo points (interpolation.cpp:1811-1826: golden-section search of kh on
 *       the leave-one-out MAE, computeResiduals(..., excludeOutsideDem = true, excludeSupplemental = true)) and the area's
 *       cells are gridded with distance + kh * topographicDistance (computeDistances :226-251). meteo = what
 *       pb_compute_residuals_glocal takes (one entry per Crit3DMeteoPoint, indexed by area.meteoPoints); cells = a
 *       pb_macro_areas_upload id or -1 for the areas' host cells; areaKh (nAreas, may be NULL) receives each area's kh.
 *       Without the meteo points a TD call fails with PB_ERR_INVALID.
 *   pb_compute_residuals_glocal  <- Project::computeResidualsAndStatisticsGlocalDetrending (project.cpp:6525-6565) over
 *       computeResidualsGlocalDetrending (agrolib/interpolation/spatialControl.cpp:231-302), areas already fitted: meteo =
 *       one entry per Crit3DMeteoPoint (position, lapse code, proxy values; meteo->isActive = active), area.meteoPoints index
 *       this array AND are matched against points->index; residual[meteo->n] = sum over the areas holding the point of
 *       (observed - interpolate()), NODATA where never computed. */
typedef struct pb_macro_area {
    int32_t nMeteoPoints;
    int32_t reserved;
    const int32_t* meteoPoints;    /* Crit3DMacroArea::getMeteoPoints() */
    int64_t nAreaCells;            /* number of FLOATS in areaCellsDEM (2 per cell) */
    const float* areaCellsDEM;     /* Crit3DMacroArea::getAreaCellsDEM() */
    /* fit (in/out): areaCombination, areaParameters (elevation first when significant) */
    uint8_t active[PB_MAX_PROXY], significant[PB_MAX_PROXY];
    int32_t useThermalInversion;
    int32_t nParameters;
    int32_t nrParams[PB_MAX_PROXY];
    double parameters[PB_MAX_PROXY][PB_MAX_FIT_PARAMS];
} pb_macro_area;
int pb_glocal_detrending_fitting(pb_context* ctx, const pb_stations* st, pb_settings* s, int variable, pb_macro_area* areas,
                                 int nAreas);
int pb_glocal_detrending_raster(pb_context* ctx, const pb_stations* st, pb_settings* s, int variable, pb_macro_area* areas,
                                int nAreas, pb_raster_id dem, const pb_raster_id* proxyGrids, int nProxyGrids, int deviceOnly,
                                pb_raster_out* out);
/* the areas' (cell, weight) pairs only change with the glocal weight maps: upload them once (like the DEM) and give the id
 * to the resident form, which then moves only the stations per time step; areas[] must describe the same areas. */
typedef int32_t pb_macro_areas_id;
int pb_macro_areas_upload(pb_context* ctx, const pb_macro_area* areas, int nAreas, pb_macro_areas_id* id);
int pb_macro_areas_free(pb_context* ctx, pb_macro_areas_id id);
int pb_glocal_detrending_raster_resident(pb_context* ctx, const pb_stations* st, pb_settings* s, int variable, pb_macro_area* areas,
                                         int nAreas, pb_macro_areas_id cells, pb_raster_id dem, const pb_raster_id* proxyGrids,
                                         int nProxyGrids, int deviceOnly, pb_raster_out* out);
int pb_compute_residuals_glocal(pb_context* ctx, const pb_stations* meteo, const float* observed, const pb_cv_points* cv,
                                const pb_stations* points, pb_settings* s, int variable, const pb_macro_area* areas, int nAreas,
                                int excludeOutsideDem, int excludeSupplemental, float* residual);
/* topographic distance inside macro areas (settings.useTD): the meteo points of the per-area kh search, the resident DEM of
 * the ray march */
typedef struct pb_glocal_meteo {
    const pb_stations* meteo;      /* one entry per Crit3DMeteoPoint (isActive = active) */
    const float* observed;         /* currentValue; NULL: meteo->value */
    const pb_cv_points* cv;        /* isInsideDem; may be NULL */
} pb_glocal_meteo;
int pb_glocal_detrending_raster_td(pb_context* ctx, const pb_stations* st, pb_settings* s, int variable, pb_macro_area* areas,
                                   int nAreas, pb_macro_areas_id cells, const pb_glocal_meteo* meteo, pb_raster_id dem,
                                   const pb_raster_id* proxyGrids, int nProxyGrids, int deviceOnly, pb_raster_out* out,
                                   int32_t* areaKh);
int pb_compute_residuals_glocal_td(pb_context* ctx, const pb_stations* meteo, const float* observed, const pb_cv_points* cv,
                                   const pb_stations* points, pb_settings* s, int variable, const pb_macro_area* areas, int nAreas,
                                   pb_raster_id dem, int excludeOutsideDem, int excludeSupplemental, float* residual);

/* --- precomputed topographic-distance maps of the points: Crit3DMeteoPoint::topographicDistance (meteo/meteoPoint.h:121),
 * one raster per meteo point written by gis::topographicDistanceMap (gis/gis.cpp:1650-1675) and loaded by
 * Project::loadTopographicDistanceMaps (project/project.cpp:2368-2430); passDataToInterpolation hands the pointer to the
 * interpolation point (interpolation/spatialControl.cpp:529). computeDistances (interpolation.cpp:232-247) then reads the
 * map at the target's cell where the target lies inside it and only marches through the DEM where the map holds NODATA.
 * pointIndex[k] = Crit3DInterpolationDataPoint::index of the point (matched against pb_stations.index, or the position
 * when that is NULL), maps[k] = a resident raster (pb_raster_upload / pb_raster_read_esri / pb_topographic_distance_map
 * with keep), or -1 to drop the point's map; n = 0 drops all. Every call with topographic distance on this context
 * (rasters, points, leave-one-out, kh search, macro areas) consults the registry; freeing a raster drops it. */
int pb_set_topographic_distance_maps(pb_context* ctx, const int32_t* pointIndex, const pb_raster_id* maps, int n);

/* --- interpolate() at explicit targets (agrolib/interpolation/interpolation.h:77-79, .cpp:2502-2560) -------
 * The building block of computeResiduals (spatialControl.cpp:97-155): x, y are the f32 target coordinates,
 * proxyValues holds m * settings.nProxy doubles (the targets' own proxy values, NODATA where missing).
 * Stations whose distance to the target is 0 weigh nothing (interpolation.cpp:1039) => leave-one-out. */
int pb_interpolate_points(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable, const float* x,
                          const float* y, const double* proxyValues, int m, int excludeSupplemental, float* result);

/* interpolate() with topographic distance (computeDistances :226-251 + gis::topographicDistance, gis/gis.cpp:1595-1647):
 * z = the targets' elevation (float(point.z)), dem = the resident DEM the ray march samples (getCurrentDEM()). */
int pb_interpolate_points_td(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable, const float* x,
                             const float* y, const float* z, const double* proxyValues, int m, int excludeSupplemental,
                             pb_raster_id dem, float* result);

/* --- leave-one-out cross-validation: computeResiduals (agrolib/interpolation/spatialControl.cpp:97-155) followed by
 * Project::computeStatisticsCrossValidation (agrolib/project/project.cpp:2935-2969). One meteo point per station:
 * observed[i] = its currentValue (NOT detrended), useForCv[i] = active && accepted && valid (NULL = all);
 * residual[i] = observed - interpolate() at the station (NODATA where not computed). `stats` may be NULL. */
int pb_compute_residuals(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable, const float* observed,
                         const uint8_t* useForCv, float* residual, pb_cv_stats* stats);

/* local-detrending chain at explicit targets (K2 in point form): per target localSelection (interpolation.cpp:1087-1170),
 * preInterpolation on the subset (multipleDetrendingMain :1832), interpolate(…, excludeSupplemental) with the targets' own
 * proxy values (m * settings.nProxy doubles). `st` = the NOT detrended interpolation points. */
int pb_local_detrending_points(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable, const float* x,
                               const float* y, const double* proxyValues, int m, int excludeSupplemental, float* result);
/* computeResidualsLocalDetrending (agrolib/interpolation/spatialControl.cpp:158-228) as Project::interpolationCv calls it
 * (agrolib/project/project.cpp:3086-3097) + computeStatisticsCrossValidation (:2935-2969): same contract as
 * pb_compute_residuals, `s` as for pb_local_detrending_raster. Supplemental stations are not cross-validated. */
int pb_compute_residuals_local(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable, const float* observed,
                               const uint8_t* useForCv, float* residual, pb_cv_stats* stats);

/* computeResiduals with the EXACT arithmetic of the reference (f32 distances, f64 weights, sequential order) for any
 * settings the station-space pipeline produces, topographic distance included (computeDistances :226-251 through the
 * resident DEM). meteo = one entry per Crit3DMeteoPoint (active and accepted): position, lapse code, proxy values;
 * observed[i] = currentValue (NULL: meteo->value); points = the interpolation points (already detrended, as
 * preInterpolation leaves them; may be a subset of meteo); residual[meteo->n]. */
int pb_compute_residuals_exact(pb_context* ctx, const pb_stations* meteo, const float* observed, const pb_cv_points* cv,
                               const pb_stations* points, const pb_settings* s, int variable, pb_raster_id dem /* -1: none */,
                               int excludeOutsideDem, int excludeSupplemental, float* residual);

/* --- spatialQualityControl (agrolib/interpolation/spatialControl.cpp:331-427), the check every
 * checkAndPassDataToInterpolation runs before gridding (checkData :466-476; not for precipitation / wind vectors):
 *   passDataToInterpolation :500-566 -> preInterpolation -> computeResiduals(…, false, false) -> per point
 *   neighbourhoodVariability (interpolation.cpp:259-301: the 10 nearest interpolation points, stdDev of their values,
 *   mean |dz|, smallest distance) -> |residual| > getSpatialThresholdVar(…, 2 stdDev) (:14-94) marks the point suspect;
 *   the suspects are then re-estimated from the field WITHOUT them (second preInterpolation) and tested at 3 stdDev.
 * st = one entry per accepted, active meteo point with its current (raw) value; s = the spatial-quality interpolation
 * settings (mutated exactly like the reference mutates them: both preInterpolation fits, points range, bounding box);
 * wrongSpatial[st->n] receives 1 where the reference leaves quality::wrong_spatial. The second pass keeps the
 * reference's indexing of the proxy values (meteoPoints[i] with i = position in the suspect list, :393). */
int pb_spatial_quality_control(pb_context* ctx, const pb_stations* st, pb_settings* s, int variable, const pb_cv_points* cv,
                               pb_raster_id dem /* -1: none */, uint8_t* wrongSpatial);

/* --- one time step of one variable, raw station values -> DEM raster: Project::interpolationDem
 * (agrolib/project/project.cpp:3106-3150), the unit of work of the reference's hourly/daily batch drivers:
 *   checkAndPassDataToInterpolation (spatialControl.cpp:479-497): spatialQualityControl with qualitySettings (NULL or a
 *   variable checkData :466 exempts: skipped), passDataToInterpolation of the accepted points (points range, bounding box);
 *   preInterpolation (every branch, as pb_pre_interpolation_ex); interpolationRaster over the resident DEM / proxy grids.
 * meteo = one entry per active meteo point holding data for this time step (raw values). step (may be NULL) selects the
 * optional outputs: wrongSpatial[n]; residual[n] + stats = the leave-one-out of Project::interpolationCv (:3012-3104:
 * computeResiduals(…, excludeOutsideDem, excludeSupplemental = true) on the same fitted points, then
 * computeStatisticsCrossValidation :2935-2969); khSeries. deviceOnly != 0: the raster stays in HBM (min/max/nValid only).
 * Both settings are mutated as the reference mutates them. pb_last_timing sums the launches of the whole step. */
typedef struct pb_dem_step {
    uint8_t* wrongSpatial;         /* may be NULL */
    float* residual;               /* may be NULL: no cross-validation */
    pb_cv_stats* stats;            /* may be NULL */
    pb_kh_series* khSeries;        /* may be NULL */
} pb_dem_step;
int pb_interpolation_dem(pb_context* ctx, const pb_stations* meteo, pb_settings* qualitySettings, pb_settings* s, int variable,
                         pb_raster_id dem, const pb_raster_id* proxyGrids, int nProxyGrids, int rowBegin, int rowEnd,
                         int deviceOnly, pb_raster_out* out, pb_dem_step* step);

/* --- a batch of independent (time step, variable) units, the loop of the reference's hourly / daily batch drivers over
 * Project::interpolationDemMain (agrolib/project/project.cpp:3526-3556): every item is one pb_interpolation_dem call (own
 * station values and settings, same resident rasters); the units are spread over nLanes internal streams (0: 4) so that the
 * raster kernel of one unit overlaps the station-space kernels of the others. items[i].status receives the unit's pb_status;
 * results are identical to stand-alone calls. */
typedef struct pb_dem_batch_item {
    const pb_stations* meteo;
    pb_settings* qualitySettings;  /* may be NULL */
    pb_settings* settings;
    int32_t variable;
    int32_t status;                /* out */
    pb_raster_out* out;
    pb_dem_step* step;             /* may be NULL */
} pb_dem_batch_item;
int pb_interpolation_dem_batch(pb_context* ctx, pb_dem_batch_item* items, int nItems, pb_raster_id dem,
                               const pb_raster_id* proxyGrids, int nProxyGrids, int rowBegin, int rowEnd, int deviceOnly,
                               int nLanes);

/* --- the hourly gridding driver kept on the device: PragaProject::interpolationMeteoGrid with getMeteoGridUpscaleFromDem()
 * (src/pragaProject/pragaProject.cpp:2951-2990). Per hour and variable the reference runs interpolationDemMain (= one
 * pb_interpolation_dem unit) and then Crit3DMeteoGrid::spatialAggregateMeteoGrid (agrolib/meteo/meteoGrid.cpp:1135-1240) of the
 * gridded map onto the meteo grid; relative humidity with getUseDewPoint() goes through the dew point (:2967-2974):
 *   passInterpolatedTemperatureToHumidityPoints (agrolib/project/project.cpp:2826-2848): a humidity station without air
 *       temperature takes the value of the hour's air temperature MAP at its position;
 *   the stations' dew point (Crit3DMeteoPoint::getMeteoPointValueH -> tDewFromRelHum, agrolib/meteo/meteo.cpp:275-285; host
 *       arithmetic) is gridded like any other variable (interpolationDemMain(airDewTemperature, ...));
 *   computeRelativeHumidityMap (agrolib/project/meteoMaps.cpp:301-321) turns the temperature and dew-point maps into the
 *       relative humidity map, which is then aggregated.
 * Here one HOUR is a list of variables worked through in order by one lane (stream + buffers of its own, like
 * pb_interpolation_dem_batch): the gridded maps stay in HBM, the air temperature map of the hour is kept for the humidity
 * chain, and only the meteo-grid cell values (nCells floats per variable, 160 kB instead of a 16 MB raster at the C5 shape)
 * and the per-station outputs cross PCIe. Hours are independent and spread over nLanes lanes.
 *   meteo           raw station values of the variable at this hour (humidity chain: relative humidity)
 *   airTemperature  humidity chain only: the same stations' own air temperature (PB_NODATA where they have none)
 *   cellValue       out, nCells of the aggregation geometry (pb_aggregation_upload); NULL: no aggregation
 *   out             optional: the map itself (data / rows NULL: stays on the device; minimum / maximum / nValid always set)
 *   step            optional per-station outputs as in pb_interpolation_dem (humidity chain: of the dew-point gridding, in the
 *                   order of `meteo`, PB_NODATA / 0 for the stations without a dew point)
 *   status          out: pb_status of this variable; a humidity chain without an air temperature map of the same hour before
 *                   it is PB_ERR_INVALID */
typedef struct pb_meteo_grid_var {
    const pb_stations* meteo;
    pb_settings* qualitySettings;  /* may be NULL */
    pb_settings* settings;
    int32_t variable;              /* PB_VAR_*; PB_VAR_AIR_REL_HUMIDITY with useDewPoint != 0: the dew-point chain */
    int32_t useDewPoint;
    const float* airTemperature;
    int32_t aggrMethod;            /* PB_AGGR_AVERAGE / PB_AGGR_MEDIAN / PB_AGGR_STDDEV */
    int32_t status;
    float* cellValue;
    pb_raster_out* out;            /* may be NULL */
    pb_dem_step* step;             /* may be NULL */
} pb_meteo_grid_var;
typedef struct pb_meteo_grid_hour {
    pb_meteo_grid_var* vars;
    int32_t nVars;
    int32_t reserved;
} pb_meteo_grid_hour;
int pb_meteo_grid_step_batch(pb_context* ctx, pb_meteo_grid_hour* hours, int nHours, pb_raster_id dem, const pb_raster_id* proxyGrids,
                             int nProxyGrids, int32_t aggregation /* pb_aggregation_id, -1: none */, int nLanes);

/* --- ordinary kriging (agrolib/interpolation/kriging.h:4-15, kriging.cpp). The reference keeps ONE system in file-static
 * globals; here every krigingVariogram call yields a handle and many systems can live on a device.
 *   pb_kriging_setup   <- krigingVariogram(pos[2n], val[n], n, mode, range, nugget, sill, slope)   kriging.cpp:97-202
 *   pb_kriging_points  <- per target: krigingSetWeight(x, y) :211-264, krigingResult() :271-279, krigingRMSE() :287-295
 *   pb_kriging_raster  <- the same over the valid cells (cell centres) of a resident DEM; estimate + RMSE grids
 *   pb_kriging_free    <- krigingFreeMemory() :299-331
 * The variogram parameters are inputs (their fit stays on the host; the reference declares krigingEstimateVariogram,
 * interpolation.h:65-68, but never defines it). flags: PB_KRIGING_FLAG_FP64_DIRECT forces the exact FP64 formulation
 * (reference elimination order); default = tensor-core variance path for the bounded models, FP64 for the linear one. */
typedef int32_t pb_kriging_id;
enum { PB_KRIGING_FLAG_FP64_DIRECT = 1 };
int pb_kriging_setup(pb_context* ctx, const double* pos, const double* val, int n, int mode, double range, double nugget,
                     double sill, double slope, int flags, pb_kriging_id* id);
int pb_kriging_free(pb_context* ctx, pb_kriging_id id);
int pb_kriging_uses_tensor_path(pb_context* ctx, pb_kriging_id id);      /* 1 tensor, 0 FP64 direct, -1 bad id */
/* measurement aid (no reference counterpart): tcgen05 MMAs of 2 * 256 * 256 * 16 flop issued for this system since the last
   call (the variance kernel skips K stages out of the variogram range); resets the counter */
int pb_kriging_mma_count(pb_context* ctx, pb_kriging_id id, unsigned long long* count);
int pb_kriging_points(pb_context* ctx, pb_kriging_id id, const double* xy, int m, double* result, double* rmse /* may be NULL */);
int pb_kriging_raster(pb_context* ctx, pb_kriging_id id, pb_raster_id dem, int rowBegin, int rowEnd,
                      pb_raster_out* estimate, pb_raster_out* rmse /* may be NULL */);

/* --- variogram estimation for ordinary kriging, HOST side (north_star: "the variogram fit stays on the host"; no context,
 * no GPU needed). The reference declares krigingEstimateVariogram(myDist, mySemiVar, sizeMyVar, nrMyPoints, myMaxDistance,
 * mySill, myNugget, myRange, mySlope, myMode, nrPointData) (agrolib/interpolation/interpolation.h:65-68) and never defines it
 * (SURVEY.md 3.5): these two calls are the engine's own definition — parity UNPINNED by construction — for the model family of
 * krigingVariogram (kriging.cpp:97-202).
 *   pb_kriging_empirical_variogram  Matheron estimator over all pairs closer than maxDistance (<= 0: half the bounding-box
 *                                   diagonal; *maxDistanceUsed may be NULL): nBins distance bins, binDist = mean pair distance,
 *                                   binSemiVar = sum (vi - vj)^2 / (2 N) (PB_NODATA for an empty bin), binCount = N
 *   pb_kriging_fit_variogram        weighted (N) least squares of the bins; mode = PB_KRIGING_* or 0 = best of spherical /
 *                                   exponential / gaussian; outputs feed pb_kriging_setup (nugget floored at 1e-4 sill: > 0) */
int pb_kriging_empirical_variogram(const double* pos, const double* val, int n, int nBins, double maxDistance, float* binDist,
                                   float* binSemiVar, int32_t* binCount, double* maxDistanceUsed);
int pb_kriging_fit_variogram(const float* binDist, const float* binSemiVar, const int32_t* binCount, int nBins, int mode,
                             double* range, double* nugget, double* sill, double* slope, int* modeOut, double* weightedSse);

/* --- raster map operators on either side of the gridding path. Inputs are resident rasters;
 * the result goes to `out` (host grid: data or rows; may be NULL) and / or stays on the device as a new resident raster
 * (*outId, to be released with pb_raster_free; NULL: not kept).
 *   pb_raster_download              a resident raster back into a host grid
 *   pb_relative_humidity_map     <- Crit3DHourlyMeteoMaps::computeRelativeHumidityMap (agrolib/project/meteoMaps.cpp:301-321)
 *                                   with relHumFromTdew (agrolib/meteo/meteo.cpp:301-315); header = the air temperature map's
 *   pb_fix_daily_thermal_consistency <- Crit3DDailyMeteoMaps::fixDailyThermalConsistency (meteoMaps.cpp:103-126): tmin is
 *                                   lowered in place to tmax - 0.1 where it exceeds tmax; *nFixed = cells changed
 *   pb_topographic_distance_map  <- gis::topographicDistanceMap (agrolib/gis/gis.cpp:1648-1672): the DEM ray march of
 *                                   gis::topographicDistance (:1595-1647) from (x, y, z) to every valid cell
 *   pb_resample_grid             <- gis::resampleGrid (gis.cpp:1722-1811), aggrAverage / aggrMedian / aggrPrevailing / aggrCenter
 *                                   (enum aggregationMethod, agrolib/mathFunctions/statistics.h:21) */
enum { PB_AGGR_AVERAGE = 1, PB_AGGR_MEDIAN = 2, PB_AGGR_PREVAILING = 7, PB_AGGR_CENTER = 9 };
int pb_raster_download(pb_context* ctx, pb_raster_id id, pb_raster_out* out);
int pb_relative_humidity_map(pb_context* ctx, pb_raster_id airT, pb_raster_id dewT, pb_raster_out* out, pb_raster_id* outId);
int pb_fix_daily_thermal_consistency(pb_context* ctx, pb_raster_id tmax, pb_raster_id tmin, int64_t* nFixed);
int pb_topographic_distance_map(pb_context* ctx, pb_raster_id dem, double x, double y, double z, pb_raster_out* out,
                                pb_raster_id* outId);
int pb_resample_grid(pb_context* ctx, pb_raster_id src, const pb_raster_header* newHeader, int method, float nodataRatioThreshold,
                     pb_raster_out* out, pb_raster_id* outId);

/* --- DEM -> meteo-grid aggregation, the step right after the gridding path in the batch drivers:
 * Crit3DMeteoGrid::spatialAggregateMeteoGrid (agrolib/meteo/meteoGrid.cpp:1135-1190) + spatialAggregateMeteoGridPoint
 * (:1193-1240). The geometry — every active meteo-grid cell's aggregationPoints (utm x, y) and aggregationPointsMaxNr, built by
 * Crit3DMeteoGrid::findGridAggregationPoints (:655-760) on the host — is uploaded once (first[c] .. first[c + 1] index x / y for
 * cell c); then every gridded map resident on the device is reduced to one value per cell: nearest-cell samples
 * (gis::getValueFromXY), coverage test against GRID_MIN_COVERAGE (= 0: at least one valid sample), aggrAverage / aggrMedian /
 * aggrStdDeviation in the reference's float arithmetic and list order (other methods: NODATA, like the reference).
 * value[c] = what the reference stores in the cell's currentValue (NODATA where it leaves the cell untouched). Each call
 * starts from fresh aggregation points (z = NODATA); the reference keeps a stale z where a later map has no data. Only the
 * nCells values cross PCIe (instead of the whole raster). */
enum { PB_AGGR_STDDEV = 3 };
typedef int32_t pb_aggregation_id;
int pb_aggregation_upload(pb_context* ctx, const double* x, const double* y, const int64_t* first, const int32_t* maxNr, int nCells,
                          pb_aggregation_id* id);
int pb_aggregation_free(pb_context* ctx, pb_aggregation_id id);
int pb_spatial_aggregate_meteo_grid(pb_context* ctx, pb_raster_id raster, pb_aggregation_id aggregation, int method, float* value);

/* topographicIndex (agrolib/project/interpolationCmd.cpp:397-482), the orography-index proxy raster of the detrending (the demo
 * ships one, DATA/GEO/TI_DEM_ER_bacini_500m): for every valid DEM cell and window width, over the disc of round(width / cellSize)
 * cells around it (its own row and column excluded, :438) the distance-weighted share of lower minus higher minus half the equal
 * neighbours; the widths add up. Same float accumulation order as the reference: bit-identical. The result (header = the DEM's)
 * goes to `out` and / or stays resident as *outId, ready to be used as a proxy grid of the raster calls. */
int pb_topographic_index(pb_context* ctx, pb_raster_id dem, const float* windowWidths, int nWidths, pb_raster_out* out,
                         pb_raster_id* outId);

/* values of a resident raster at explicit points (nearest cell): Crit3DRasterGrid::getRowCol (agrolib/gis/gis.cpp:485-492),
 * gis::isOutOfGridRowCol (:771-776), value[row][col] — the lookup of Project::passInterpolatedTemperatureToHumidityPoints
 * (agrolib/project/project.cpp:2826-2848): stations with humidity but no temperature take the air temperature of the map just
 * gridded, so that their dew point (tDewFromRelHum, host arithmetic of the meteo point) can be gridded next and turned into
 * relative humidity by pb_relative_humidity_map. value[i] = the cell's value (the raster's flag included, like the
 * reference), PB_NODATA outside the grid; inGrid[i] (may be NULL) = 1 inside. Only n values cross PCIe. */
int pb_raster_sample_points(pb_context* ctx, pb_raster_id raster, const double* x, const double* y, int n, float* value,
                            uint8_t* inGrid);

/* --- solar radiation maps: radiation::computeRadiationDEM (agrolib/solarRadiation/solarRadiation.cpp:1045-1069)
 * = for every valid DEM cell computeRadiationDemPoint (:946-981) -> computeRadiationRsun (:700-834): sun position (NREL SOLPOS,
 * solPos.cpp / sunPosition.cpp), shadowing by the DEM (computeShadow :543-604), real-sky split of the transmissivity
 * (separateTransmissivity_Erbs_Reindl :634-690) or the Linke clear-sky model (:343-392), irradiance on the cell's slope / aspect
 * (getBeamInclined, getDiffuseInclined_Muneer, getReflectedIrradiance) — then gis::updateMinMaxRasterGrid of the maps
 * (updateRadiationMaps :1072-1089). The transmissivity map is what Project::interpolateDemRadiation (agrolib/project/project.cpp:
 * 3382-3437) grids from the stations' atmTransmissivity with this engine; lat / lon / slope / aspect are the static maps of
 * Crit3DRadiationMaps (gis::computeLatLonMaps, gis::computeSlopeAspectMaps), uploaded once. All maps are resident rasters with
 * the DEM's geometry. The time is the LOCAL time (the host applies isUTC / timeZone like computeRadiationRsun :714-719).
 * Linke / albedo: the scalar the reference ends up with for an in-grid cell — the fixed value, the month's for PARAM_MODE_MONTHLY
 * (Linke), PB_NODATA for the map modes, whose accessors only ever read OUT-of-grid cells (radiationSettings.cpp:131-133, :179-181).
 * With shadowing off the reference reads an uninitialised TsunPosition::shadow (:744-750, :801); here it is false. */
typedef struct pb_radiation_settings {
    int32_t realSky, realSkyAlgorithm /* 0 total transmissivity, 1 Linke */, shadowing;
    int32_t tiltMode /* 1 fixed, 2 dem */, linkeMode, albedoMode /* 0 fixed, 1 map, 2 monthly */;
    float linke, albedo, tilt, aspect, clearSky;
    float demMaximum;              /* dem.maximum (gis::updateMinMaxRasterGrid): bound of the shadow march */
    int32_t timeZone;
    int32_t year, month, day, secondsOfDay;
} pb_radiation_settings;
typedef struct pb_radiation_maps {
    pb_raster_id dem, lat, lon, slope, aspect, transmissivity;
} pb_radiation_maps;
typedef struct pb_radiation_out {
    pb_raster_out* sunElevation;   /* each may be NULL or have data / rows NULL: not downloaded; minimum / maximum always set */
    pb_raster_out* beam;
    pb_raster_out* diffuse;
    pb_raster_out* reflected;
    pb_raster_out* global;
    int32_t keepResident;          /* != 0: the five maps stay resident, ids[] (same order) to be released with pb_raster_free */
    pb_raster_id ids[5];
} pb_radiation_out;
int pb_compute_radiation_dem(pb_context* ctx, const pb_radiation_settings* s, const pb_radiation_maps* maps, pb_radiation_out* out);
/* the map the last raster call of this context left on the device (pb_*_device forms, deviceOnly) as a resident raster with
 * the header of `like` (the DEM): e.g. the gridded transmissivity map handed to pb_compute_radiation_dem without a PCIe trip */
int pb_raster_from_last_output(pb_context* ctx, pb_raster_id like, pb_raster_id* id);

/* --- ESRI float grids (.hdr + .flt), the on-disk format at either end of the path, streamed between
 * the file and HBM through pinned staging (no pageable full-size copy):
 *   pb_raster_read_esri_header <- gis::readEsriGridHeader (agrolib/gis/gisIO.cpp:118-193)
 *   pb_raster_read_esri        <- gis::readEsriGrid (:516-557: header, readRasterFloatData :347-422 for 4 / 2 / 1-byte values,
 *                                 updateMinMaxRasterGrid); fileName with or without ".flt"; the grid becomes a resident raster
 *   pb_raster_write_esri       <- gis::writeEsriGrid (:504-513; header text byte-identical to writeEsriGridHeader :432-467);
 *                                 fileName without extension */
int pb_raster_read_esri_header(pb_context* ctx, const char* fileName, pb_raster_header* header);
int pb_raster_read_esri(pb_context* ctx, const char* fileName, pb_raster_id* id, pb_raster_header* header /* may be NULL */,
                        float* minimum /* may be NULL */, float* maximum /* may be NULL */);
int pb_raster_write_esri(pb_context* ctx, pb_raster_id id, const char* fileName);

/* --- timing of the last call (device time measured with CUDA events on the engine's stream) ------ */
typedef struct pb_timing {
    float h2d_ms, kernel_ms, d2h_ms, total_ms;
    int32_t launches;              /* kernels launched by the last call */
    int64_t h2d_bytes, d2h_bytes;
} pb_timing;
int pb_last_timing(const pb_context* ctx, pb_timing* t);
/* live ceilings of the two pipes that bound the IDW station loop: FP32 FMA (TFLOP/s, FFMA2 chain) and
 * MUFU.RSQ (G op/s). Used by bench.py as roofline denominators (MEASURED_PEAKS.json has no FP32 figure). */
int pb_measure_pipe_peaks(pb_context* ctx, float* fp32Tflops, float* mufuGops);
/* FP64 FMA ceiling (TFLOP/s, DFMA chain): roofline denominator of the local-detrending and kriging-estimate kernels */
int pb_measure_fp64_peak(pb_context* ctx, float* fp64Tflops);

#ifdef __cplusplus
}
#endif
#endif /* PRAGA_B200_H */
