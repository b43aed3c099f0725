/*
 * praga_b200_host.h — the C++ host side of the drop-in: the reference's own signatures for the gridding hot path,
 * implemented by flattening the reference objects into the POD descriptors of include/praga_b200.h and calling
 * libpraga_b200.so. Compiled against the reference's headers (agrolib/{gis,meteo,interpolation,mathFunctions}); a
 * maintainer links this translation unit instead of the reference bodies.
 *
 * Same names, argument meaning and error behaviour as the reference:
 *   pragab200::interpolationRaster            <- agrolib/project/interpolationCmd.h:73-75 (.cpp:353-394)
 *   pragab200::interpolationRasterLocalDetrending <- loop of Project::interpolationDemLocalDetrending, project.cpp:3208-3253
 *   pragab200::preInterpolation               <- agrolib/interpolation/interpolation.h (.cpp:2444-2499): precipitation check,
 *                                                multiple detrending, classic detrending(), optimalDetrending(), kh search
 *                                                with useGlocalDetrending: glocalDetrendingFitting (.cpp:2239-2294), the macro
 *                                                areas of the settings receive their fit
 *   pragab200::interpolationRasterGlocalDetrending <- loop of Project::interpolationDemGlocalDetrending, project.cpp:3283-3375
 *   pragab200::computeResidualsGlocalDetrending <- agrolib/interpolation/spatialControl.h:42 (.cpp:231-302)
 *   pragab200::computeResiduals               <- agrolib/interpolation/spatialControl.h:30 (.cpp:97-155)
 *   pragab200::computeResidualsLocalDetrending <- agrolib/interpolation/spatialControl.h (.cpp:158-228)
 *   pragab200::spatialQualityControl          <- agrolib/interpolation/spatialControl.h (.cpp:331-427)
 *   pragab200::krigingVariogram / krigingSetWeight / krigingResult / krigingRMSE / krigingFreeMemory
 *                                             <- agrolib/interpolation/kriging.h:4-15
 *   pragab200::krigingRaster                  (new: the reference never wired kriging into a raster loop)
 *   pragab200::computeRadiationDEM            <- agrolib/solarRadiation/solarRadiation.h (.cpp:1045-1069)
 * A GPU failure returns false (there is no CPU fallback); pragab200::lastError() gives the message.
 */
#ifndef PRAGA_B200_HOST_H
#define PRAGA_B200_HOST_H

#include <string>
#include <vector>

#include "gis.h"
#include "meteo.h"
#include "meteoPoint.h"
#include "interpolationSettings.h"
#include "interpolationPoint.h"
#include "solarRadiation.h"

#include "../../include/praga_b200.h"

namespace pragab200 {

/* device used by the calls below (default 0); one engine context per device is created lazily and kept */
bool setDevice(int device);
const std::string& lastError();

/* Resident raster cache of the raster calls below. The reference keeps its DEM and proxy grids loaded for the whole session
 * (Project::loadDEM, project.cpp:1050-1120; loadProxyGrids :1960-2060) and grids every time step onto them; the drop-in keeps
 * one device copy per (device, grid) — key: the grid's `value` row-pointer array + header — so a time step moves only the
 * stations up and the output grid down. A grid that is modified IN PLACE or freed must be announced with invalidate(); as a
 * safety net every call re-checks a fingerprint of the grid (256 sampled cells, 4 row pointers) and uploads again on a
 * mismatch. setRasterCache(false) restores upload-per-call; limitBytes (0 = keep) bounds the device memory held (LRU). */
/* device time, launches and bytes moved by the last call on the current device (pb_last_timing of the shim's own context) */
bool lastTiming(pb_timing* t);
void setRasterCache(bool enabled, size_t limitBytes = 0);
void invalidate(const gis::Crit3DRasterGrid* grid);
void invalidateAll();

bool interpolationRaster(std::vector<Crit3DInterpolationDataPoint>& dataPoints,
                         Crit3DInterpolationSettings& interpolationSettings, Crit3DMeteoSettings* meteoSettings,
                         gis::Crit3DRasterGrid* outputGrid, gis::Crit3DRasterGrid& raster, meteoVariable variable,
                         bool isParallelComputing);

/* interpolationPoints are the points of checkAndPassDataToInterpolation (NOT detrended); settings need
 * useLocalDetrending + useMultipleDetrending and the points range (setPointsRange) */
bool interpolationRasterLocalDetrending(std::vector<Crit3DInterpolationDataPoint>& interpolationPoints,
                                        Crit3DInterpolationSettings& interpolationSettings,
                                        Crit3DMeteoSettings* meteoSettings, gis::Crit3DRasterGrid* outputGrid,
                                        gis::Crit3DRasterGrid& raster, meteoVariable variable);

/* glocal detrending: interpolationPoints are the points of checkAndPassDataToInterpolation (NOT detrended); the settings carry
 * the macro areas (setMacroAreas), useGlocalDetrending + useMultipleDetrending and the points range. Does what
 * Project::interpolationDemGlocalDetrending does from preInterpolation on: fits every area (the areas in the settings keep
 * parameters and combination), grids each area's cells over the area's stations and blends them by weight into outputGrid.
 * Returns false where the reference does (an interpolate() inside an area gave NODATA). */
bool interpolationRasterGlocalDetrending(std::vector<Crit3DInterpolationDataPoint>& interpolationPoints,
                                         Crit3DInterpolationSettings& interpolationSettings, Crit3DMeteoSettings* meteoSettings,
                                         gis::Crit3DRasterGrid* outputGrid, gis::Crit3DRasterGrid& raster, meteoVariable variable,
                                         const std::vector<Crit3DMeteoPoint>* meteoPoints = nullptr /* Project::meteoPoints: needed
                                         when the settings use topographic distance (the per-area kh search walks them) */);

/* same signature and side effects as the reference (residuals are ADDED to meteoPoints[].residual, one area per call) */
bool computeResidualsGlocalDetrending(meteoVariable myVar, const Crit3DMacroArea& myArea, int elevationPos,
                                      std::vector<Crit3DMeteoPoint>& meteoPoints,
                                      std::vector<Crit3DInterpolationDataPoint>& interpolationPoints,
                                      Crit3DInterpolationSettings& interpolationSettings, Crit3DMeteoSettings* meteoSettings,
                                      bool excludeOutsideDem, bool excludeSupplemental);

/* same contract as the reference: myPoints are detrended in place (points multiple detrending drops are erased), the
 * settings receive the fit (current combination, proxies' regression fields or fitting functions/parameters, kh and the
 * Kh series, points range / bounding box). meteoPoints: the points passDataToInterpolation built myPoints from (one per
 * interpolation point with the same index order once inactive / not accepted ones are skipped). */
bool preInterpolation(std::vector<Crit3DInterpolationDataPoint>& myPoints, Crit3DInterpolationSettings& interpolationSettings,
                      Crit3DMeteoSettings* meteoSettings, Crit3DClimateParameters* climateParameters,
                      std::vector<Crit3DMeteoPoint>& meteoPoints, meteoVariable myVar, Crit3DTime myTime, std::string& errorStr);

bool computeResiduals(meteoVariable myVar, std::vector<Crit3DMeteoPoint>& meteoPoints,
                      const std::vector<Crit3DInterpolationDataPoint>& interpolationPoints,
                      Crit3DInterpolationSettings& interpolationSettings, Crit3DMeteoSettings* meteoSettings,
                      bool excludeOutsideDem, bool excludeSupplemental);

/* the interpolation points are the NOT detrended points of passDataToInterpolation; the caller has run
 * setMultipleDetrendingHeightTemperatureRange like Project::interpolationCv does (project.cpp:3086-3097) */
bool computeResidualsLocalDetrending(meteoVariable myVar, const Crit3DTime& myTime, std::vector<Crit3DMeteoPoint>& meteoPoints,
                                     std::vector<Crit3DInterpolationDataPoint>& interpolationPoints,
                                     Crit3DInterpolationSettings& interpolationSettings, Crit3DMeteoSettings* meteoSettings,
                                     Crit3DClimateParameters* climateParameters, bool excludeOutsideDem, bool excludeSupplemental);

/* marks meteoPoints[i].quality = quality::wrong_spatial exactly where the reference does; the settings (the project's
 * qualityInterpolationSettings) receive both preInterpolation fits, points range and bounding box */
bool spatialQualityControl(meteoVariable myVar, std::vector<Crit3DMeteoPoint>& meteoPoints,
                           Crit3DInterpolationSettings& interpolationSettings, Crit3DMeteoSettings* meteoSettings,
                           Crit3DClimateParameters* climateParameters, const Crit3DTime& myTime, std::string& errorStr);

bool krigingVariogram(double* myPos, double* myVal, int nrItems, short myMode, double myRange, double myNugget,
                      double mySill, double mySlope);
bool krigingSetWeight(double x_p, double y_p);
double krigingResult();
double krigingRMSE();
void krigingFreeMemory();
/* the variogram fit the reference declares but never defines (agrolib/interpolation/interpolation.h:65-68), same parameter
 * list: myDist / mySemiVar = the sizeMyVar bins of the empirical variogram; *myMode in: the model to fit (an invalid value:
 * the best of spherical / exponential / gaussian), out: the model fitted. Host arithmetic (pb_kriging_fit_variogram). */
bool krigingEstimateVariogram(float* myDist, float* mySemiVar, int sizeMyVar, int nrMyPoints, float myMaxDistance, double* mySill,
                              double* myNugget, double* myRange, double* mySlope, TkrigingMode* myMode, int nrPointData);
/* batched form: estimate and RMSE for every valid cell (cell centres) of `raster` with the current system */
bool krigingRaster(gis::Crit3DRasterGrid& raster, gis::Crit3DRasterGrid* estimateGrid, gis::Crit3DRasterGrid* rmseGrid);

/* radiation::computeRadiationDEM (agrolib/solarRadiation/solarRadiation.h, .cpp:1045-1069): same signature and side effects
 * (the five maps of radiationMaps, their minimum / maximum / mapTime, setComputed). The DEM and the static lat / lon / slope /
 * aspect maps stay resident (raster cache above). dem.maximum must be up to date (gis::updateMinMaxRasterGrid), as the
 * reference's shadow march needs it too. */
bool computeRadiationDEM(Crit3DRadiationSettings* radSettings, const gis::Crit3DRasterGrid& dem, Crit3DRadiationMaps* radiationMaps,
                         const Crit3DTime& myTime, bool isParallelComputing);

}  // namespace pragab200

#endif
