/* praga_b200_host.cpp — see praga_b200_host.h. Only flattening + C-ABI calls; no arithmetic of the path lives here. */
#include "praga_b200_host.h"

#include <string.h>
#include <map>

#include "commonConstants.h"
#include "basicMath.h"
#include "furtherMathFunctions.h"
#include "interpolation.h"
#include "interpolationConstants.h"
#include "solarRadiation.h"

#include "../../include/praga_b200.h"

namespace pragab200 {
namespace {

int g_device = 0;
std::map<int, pb_context*> g_ctx;
std::string g_error;

pb_context* context()
{
    auto it = g_ctx.find(g_device);
    if (it != g_ctx.end()) return it->second;
    pb_context* ctx = nullptr;
    if (pb_create(g_device, &ctx) != PB_OK) {
        g_error = pb_last_error(nullptr);
        return nullptr;
    }
    g_ctx[g_device] = ctx;
    return ctx;
}

bool failed(pb_context* ctx, int rc)
{
    if (rc == PB_OK) return false;
    g_error = ctx ? pb_last_error(ctx) : "no context";
    return true;
}

typedef double (*fit1d_t)(double, std::vector<double>&);
int fitFunctionId(const std::function<double(double, std::vector<double>&)>& f)
{
    auto t = f.target<fit1d_t>();
    if (t == nullptr) return PB_FIT_NONE;
    if (*t == lapseRatePiecewise_two) return PB_FIT_PIECEWISE_TWO;
    if (*t == lapseRatePiecewise_three_free) return PB_FIT_PIECEWISE_THREE_FREE;
    if (*t == lapseRatePiecewise_three) return PB_FIT_PIECEWISE_THREE;
    if (*t == functionLinear_intercept) return PB_FIT_LINEAR;
    return PB_FIT_NONE;
}

pb_raster rasterOf(const gis::Crit3DRasterGrid& g)
{
    pb_raster r;
    memset(&r, 0, sizeof(r));
    r.h.nrRows = g.header->nrRows;
    r.h.nrCols = g.header->nrCols;
    r.h.cellSize = g.header->cellSize;
    r.h.invCellSize = g.header->invCellSize;
    r.h.llx = g.header->llCorner.x;
    r.h.lly = g.header->llCorner.y;
    r.h.flag = g.header->flag;
    r.rows = g.value;
    return r;
}

/* Crit3DInterpolationSettings (+ rainfall threshold) -> pb_settings; proxy grids collected into `grids` */
bool flattenSettings(Crit3DInterpolationSettings& st, Crit3DMeteoSettings* meteoSettings, pb_settings& s,
                     std::vector<pb_raster>& grids, const gis::Crit3DRasterGrid* dem)
{
    memset(&s, 0, sizeof(s));
    const unsigned P = unsigned(st.getProxyNr());
    if (P > PB_MAX_PROXY) { g_error = "more than PB_MAX_PROXY proxies"; return false; }
    s.method = int(st.getInterpolationMethod());
    s.useThermalInversion = st.getUseThermalInversion();
    s.useTD = st.getUseTD();
    s.useLocalDetrending = st.getUseLocalDetrending();
    s.useGlocalDetrending = st.getUseGlocalDetrending();
    s.useLapseRateCode = st.getUseLapseRateCode();
    s.useBestDetrending = st.getUseBestDetrending();
    s.useMultipleDetrending = st.getUseMultipleDetrending();
    s.useDoNotRetrend = st.getUseDoNotRetrend();
    s.useRetrendOnly = st.getUseRetrendOnly();
    s.precipitationAllZero = st.getPrecipitationAllZero();
    s.minPointsLocalDetrending = st.getMinPointsLocalDetrending();
    s.topoDist_Kh = st.getTopoDist_Kh();
    s.topoDist_maxKh = st.getTopoDist_maxKh();
    s.minRegressionR2 = st.getMinRegressionR2();
    s.maxHeightInversion = st.getMaxHeightInversion();
    s.pointsBoundingBoxArea = st.getPointsBoundingBoxArea();
    s.localRadius = st.getLocalRadius();
    s.rainfallThreshold = meteoSettings ? meteoSettings->getRainfallThreshold() : 0.f;
    s.climateLapseRate = float(NODATA);
    std::vector<double> range = st.getPointsRange();
    s.hasPointsRange = range.size() >= 2;
    if (s.hasPointsRange) { s.pointsRangeMin = range[0]; s.pointsRangeMax = range[1]; }
    s.nProxy = int(P);
    Crit3DProxyCombination sel = st.getSelectedCombination(), cur = st.getCurrentCombination();
    s.selectedUseThermalInversion = sel.getUseThermalInversion();
    s.currentUseThermalInversion = cur.getUseThermalInversion();
    for (unsigned i = 0; i < P; i++) {
        s.selectedActive[i] = i < sel.getProxySize() && sel.isProxyActive(i);
        s.currentActive[i] = i < cur.getProxySize() && cur.isProxyActive(i);
        s.currentSignificant[i] = i < cur.getProxySize() && cur.isProxySignificant(i);
        Crit3DProxy* p = st.getProxy(i);
        pb_proxy& q = s.proxy[i];
        q.kind = int(getProxyPragaName(p->getName()));
        q.gridIndex = -1;
        gis::Crit3DRasterGrid* g = p->getGrid();
        if (g != nullptr && g->isLoaded) {
            q.gridIndex = int(grids.size());
            grids.push_back(rasterOf(*g));
            if (dem != nullptr && g == dem) grids.back().rows = dem->value;    // same buffer => uploaded once
        }
        q.fittingFunction = int(p->getFittingFunctionName());
        q.inversionIsSignificative = p->getInversionIsSignificative();
        q.regressionR2 = p->getRegressionR2();
        q.regressionSlope = p->getRegressionSlope();
        q.regressionIntercept = p->getRegressionIntercept();
        q.lapseRateH0 = p->getLapseRateH0();
        q.lapseRateH1 = p->getLapseRateH1();
        q.lapseRateT0 = p->getLapseRateT0();
        q.lapseRateT1 = p->getLapseRateT1();
        q.inversionLapseRate = p->getInversionLapseRate();
        q.stdDevThreshold = p->getStdDevThreshold();
        std::vector<double> fr = p->getFittingParametersRange();
        const int np = int(fr.size() / 2);
        if (np > PB_MAX_FIT_PARAMS) { g_error = "too many fitting parameters"; return false; }
        q.nFitParams = np;
        for (int j = 0; j < np; j++) { q.fitMin[j] = fr[j]; q.fitMax[j] = fr[np + j]; }
        std::vector<int> fg = p->getFittingFirstGuess();
        for (int j = 0; j < np && j < int(fg.size()); j++) q.fitFirstGuess[j] = fg[j];
        std::vector<std::vector<double>> combos = p->getFirstGuessCombinations();
        q.nFirstGuessCombinations = int(std::min<size_t>(combos.size(), PB_MAX_FIRST_GUESS));
        for (int k = 0; k < q.nFirstGuessCombinations; k++)
            for (size_t j = 0; j < combos[k].size() && j < PB_MAX_FIT_PARAMS; j++) q.firstGuessCombinations[k][j] = combos[k][j];
    }
    auto funcs = st.getFittingFunction();
    auto params = st.getFittingParameters();
    s.nFit = int(params.size());
    s.nFitFunctions = int(funcs.size());
    if (s.nFit > PB_MAX_PROXY || s.nFitFunctions > PB_MAX_PROXY) { g_error = "too many fitted functions"; return false; }
    for (int i = 0; i < s.nFitFunctions; i++) s.fitFunction[i] = fitFunctionId(funcs[i]);
    for (int i = 0; i < s.nFit; i++) {
        s.fitNrParams[i] = int(std::min<size_t>(params[i].size(), PB_MAX_FIT_PARAMS));
        for (int j = 0; j < s.fitNrParams[i]; j++) s.fitParams[i][j] = params[i][j];
    }
    return true;
}

struct FlatPoints {
    std::vector<double> x, y, z;
    std::vector<float> value, weight, proxy;
    std::vector<int32_t> code, index;
    std::vector<uint8_t> active;
    pb_stations st;
};

void flattenPoints(const std::vector<Crit3DInterpolationDataPoint>& pts, int nProxy, FlatPoints& f)
{
    const size_t n = pts.size();
    f.x.resize(n); f.y.resize(n); f.z.resize(n); f.value.resize(n); f.weight.resize(n); f.code.resize(n);
    f.index.resize(n); f.active.resize(n);
    f.proxy.assign(n * size_t(std::max(nProxy, 1)), float(NODATA));
    for (size_t i = 0; i < n; i++) {
        const Crit3DInterpolationDataPoint& p = pts[i];
        f.x[i] = p.point->utm.x; f.y[i] = p.point->utm.y; f.z[i] = p.point->z;
        f.value[i] = p.value;
        f.weight[i] = p.regressionWeight;
        f.code[i] = int(p.lapseRateCode);
        f.index[i] = p.index;
        f.active[i] = p.isActive;
        for (int k = 0; k < nProxy && k < int(p.proxyValues.size()); k++) f.proxy[i * nProxy + k] = p.proxyValues[k];
    }
    memset(&f.st, 0, sizeof(f.st));
    f.st.n = int(n);
    f.st.nProxy = nProxy;
    f.st.x = f.x.data(); f.st.y = f.y.data(); f.st.z = f.z.data();
    f.st.value = f.value.data(); f.st.regressionWeight = f.weight.data(); f.st.lapseRateCode = f.code.data();
    f.st.isActive = f.active.data(); f.st.proxyValues = f.proxy.data(); f.st.index = f.index.data();
}

// Project::meteoPoints as the C ABI takes them (pb_compute_residuals_glocal, pb_glocal_meteo): one entry per
// Crit3DMeteoPoint; only active, accepted points are cross-validated
struct FlatMeteo {
    std::vector<double> x, y, z;
    std::vector<float> value, proxy;
    std::vector<int32_t> code;
    std::vector<uint8_t> active, inside;
    pb_stations st;
    pb_cv_points cv;
};

void flattenMeteo(const std::vector<Crit3DMeteoPoint>& meteoPoints, int P, FlatMeteo& f)
{
    const int n = int(meteoPoints.size());
    f.x.resize(n); f.y.resize(n); f.z.resize(n); f.value.resize(n); f.code.resize(n); f.active.resize(n); f.inside.resize(n);
    f.proxy.assign(size_t(n) * std::max(P, 1), float(NODATA));
    for (int i = 0; i < n; i++) {
        const Crit3DMeteoPoint& m = meteoPoints[i];
        f.x[i] = m.point.utm.x; f.y[i] = m.point.utm.y; f.z[i] = m.point.z;
        f.value[i] = m.currentValue;
        f.code[i] = int(m.lapseRateCode);
        f.active[i] = m.active && m.quality == quality::accepted;
        f.inside[i] = m.isInsideDem;
        for (int k = 0; k < P && k < int(m.proxyValues.size()); k++) f.proxy[size_t(i) * P + k] = m.proxyValues[k];
    }
    memset(&f.st, 0, sizeof(f.st));
    f.st.n = n; f.st.nProxy = P;
    f.st.x = f.x.data(); f.st.y = f.y.data(); f.st.z = f.z.data(); f.st.value = f.value.data(); f.st.lapseRateCode = f.code.data();
    f.st.isActive = f.active.data(); f.st.proxyValues = f.proxy.data();
    memset(&f.cv, 0, sizeof(f.cv));
    f.cv.isInsideDem = f.inside.data();
}

/* ---- resident raster cache -----------------------------------------------------------------------------------------
 * The reference keeps its DEM and proxy grids loaded for the whole session (Project::loadDEM / loadProxyGrids) and every
 * time step grids onto them; the drop-in therefore keeps one device copy per (device, grid): key = the grid's row-pointer
 * array + header. A call with a cached grid moves only the stations up and the output grid down. A grid that is modified IN
 * PLACE must be announced with pragab200::invalidate(&grid) (INTEGRATION.md shows the two hooks); as a safety net every hit
 * also re-checks a fingerprint of 256 sampled cells and the row pointers, and a mismatch uploads the grid again. */
struct CachedGrid {
    const gis::Crit3DRasterGrid* owner;
    float* const* rows;
    pb_raster_header h;
    pb_raster_id id;
    std::vector<float> sample;
    std::vector<const float*> rowPtr;
    unsigned long long lastUse;
    size_t bytes;
};
std::map<int, std::vector<CachedGrid>> g_cache;
unsigned long long g_useClock = 0;
bool g_cacheEnabled = true;
size_t g_cacheLimit = size_t(16) << 30;

bool sameHeader(const pb_raster_header& a, const pb_raster_header& b)
{
    return a.nrRows == b.nrRows && a.nrCols == b.nrCols && a.cellSize == b.cellSize && a.llx == b.llx && a.lly == b.lly &&
           a.flag == b.flag;
}

void fingerprint(const pb_raster& r, std::vector<float>& sample, std::vector<const float*>& rowPtr)
{
    const int kSamples = 256;
    sample.resize(kSamples);
    unsigned long long x = 0x9E3779B97F4A7C15ull;
    for (int k = 0; k < kSamples; k++) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        const int row = int((x >> 11) % (unsigned long long)r.h.nrRows), col = int((x >> 37) % (unsigned long long)r.h.nrCols);
        sample[k] = r.rows[row][col];
    }
    const int rowsToCheck[4] = {0, r.h.nrRows / 3, (2 * r.h.nrRows) / 3, r.h.nrRows - 1};
    rowPtr.clear();
    for (int q : rowsToCheck) rowPtr.push_back(r.rows[q]);
}

void dropEntry(pb_context* ctx, std::vector<CachedGrid>& v, size_t i)
{
    pb_raster_free(ctx, v[i].id);
    v.erase(v.begin() + long(i));
}

/* device id of a host grid: cached copy when it is still the same grid, else upload (evicting least-recently-used copies
 * beyond the byte limit). *temporary = the caller frees it after the call (cache disabled). */
bool residentId(pb_context* ctx, const gis::Crit3DRasterGrid* owner, const pb_raster& r, pb_raster_id* id, bool* temporary)
{
    *temporary = false;
    if (!g_cacheEnabled) {
        *temporary = true;
        return !failed(ctx, pb_raster_upload(ctx, &r, id));
    }
    std::vector<CachedGrid>& v = g_cache[g_device];
    std::vector<float> sample;
    std::vector<const float*> rowPtr;
    fingerprint(r, sample, rowPtr);
    for (size_t i = 0; i < v.size(); i++) {
        if (v[i].rows != r.rows) continue;
        if (sameHeader(v[i].h, r.h) && v[i].rowPtr == rowPtr && memcmp(v[i].sample.data(), sample.data(), sample.size() * sizeof(float)) == 0) {
            v[i].lastUse = ++g_useClock;
            v[i].owner = owner;
            *id = v[i].id;
            return true;
        }
        dropEntry(ctx, v, i);       // same address, different grid: stale
        break;
    }
    const size_t bytes = size_t(r.h.nrRows) * r.h.nrCols * sizeof(float);
    size_t total = bytes;
    for (const CachedGrid& c : v) total += c.bytes;
    while (total > g_cacheLimit && !v.empty()) {
        size_t lru = 0;
        for (size_t i = 1; i < v.size(); i++) if (v[i].lastUse < v[lru].lastUse) lru = i;
        total -= v[lru].bytes;
        dropEntry(ctx, v, lru);
    }
    CachedGrid c;
    c.owner = owner; c.rows = r.rows; c.h = r.h; c.sample.swap(sample); c.rowPtr.swap(rowPtr); c.lastUse = ++g_useClock; c.bytes = bytes;
    if (failed(ctx, pb_raster_upload(ctx, &r, &c.id))) return false;
    *id = c.id;
    v.push_back(std::move(c));
    return true;
}

/* The points' precomputed topographic-distance maps (Crit3DInterpolationDataPoint::topographicDistance, copied from
 * Crit3DMeteoPoint::topographicDistance by passDataToInterpolation, spatialControl.cpp:529) registered with the engine for the
 * duration of one call: each map becomes a resident raster through the cache above, keyed by the point's index. */
struct TdMapScope {
    pb_context* ctx;
    std::vector<pb_raster_id> temporaries;
    bool registered = false, ok = true;
    explicit TdMapScope(pb_context* c) : ctx(c) {}
    void add(std::vector<int32_t>& index, std::vector<pb_raster_id>& ids, int pointIndex, const gis::Crit3DRasterGrid* map)
    {
        if (!ok || map == nullptr || !map->isLoaded || map->value == nullptr) return;
        pb_raster_id id = -1;
        bool tmp = false;
        if (!residentId(ctx, map, rasterOf(*map), &id, &tmp)) { ok = false; return; }
        if (tmp) temporaries.push_back(id);
        index.push_back(pointIndex);
        ids.push_back(id);
    }
    bool commit(const std::vector<int32_t>& index, const std::vector<pb_raster_id>& ids)
    {
        if (!ok) return false;
        if (index.empty()) return true;
        registered = true;
        // (a map evicted from the raster cache while the others were uploaded is no longer live: the call fails here instead of
        // silently marching; pragab200::setRasterCache raises the limit)
        if (failed(ctx, pb_set_topographic_distance_maps(ctx, index.data(), ids.data(), int(index.size())))) { ok = false; return false; }
        return true;
    }
    bool fromPoints(const std::vector<Crit3DInterpolationDataPoint>& pts, bool useTD)
    {
        if (!useTD) return true;
        std::vector<int32_t> index;
        std::vector<pb_raster_id> ids;
        for (const auto& p : pts) add(index, ids, p.index, p.topographicDistance);
        return commit(index, ids);
    }
    ~TdMapScope()
    {
        if (registered) pb_set_topographic_distance_maps(ctx, nullptr, nullptr, 0);
        for (pb_raster_id t : temporaries) pb_raster_free(ctx, t);
    }
};

bool rasterCall(bool local, std::vector<Crit3DInterpolationDataPoint>& points, Crit3DInterpolationSettings& settings,
                Crit3DMeteoSettings* meteoSettings, gis::Crit3DRasterGrid* outputGrid, gis::Crit3DRasterGrid& raster,
                meteoVariable variable)
{
    // output = DEM header, every cell = flag (interpolationCmd.cpp:357 -> gis.cpp:357-363). A grid that already has this
    // geometry keeps its rows: every cell is overwritten below (valid cells by the estimate, the others by the flag).
    const bool reuse = outputGrid->isLoaded && outputGrid->value != nullptr && outputGrid->header->nrRows == raster.header->nrRows &&
                       outputGrid->header->nrCols == raster.header->nrCols;
    if (reuse) {
        outputGrid->mapTime.setNullTime();       // what clear() resets besides the rows (gis.cpp:265-288)
        outputGrid->minimum = outputGrid->maximum = NODATA;
        *outputGrid->header = *raster.header;
    } else if (!outputGrid->initializeGrid(raster)) return false;
    pb_context* ctx = context();
    if (!ctx) return false;
    pb_settings s;
    std::vector<pb_raster> grids;
    if (!flattenSettings(settings, meteoSettings, s, grids, &raster)) return false;
    FlatPoints fp;
    flattenPoints(points, s.nProxy, fp);
    pb_raster dem = rasterOf(raster);
    pb_raster_out out;
    memset(&out, 0, sizeof(out));
    out.rows = outputGrid->value;
    // rasters: resident copies (cached across calls)
    pb_raster_id demId = -1, ids[PB_MAX_PROXY];
    std::vector<pb_raster_id> temporaries;
    bool tmp = false;
    bool ok = residentId(ctx, &raster, dem, &demId, &tmp);
    if (ok && tmp) temporaries.push_back(demId);
    for (size_t g = 0; ok && g < grids.size(); g++) {
        if (grids[g].rows == dem.rows && sameHeader(grids[g].h, dem.h)) { ids[g] = demId; continue; }
        ok = residentId(ctx, nullptr, grids[g], &ids[g], &tmp);
        if (ok && tmp) temporaries.push_back(ids[g]);
    }
    TdMapScope tdMaps(ctx);
    if (ok) ok = tdMaps.fromPoints(points, s.useTD != 0);
    int rc = PB_ERR_CUDA;
    if (ok)
        rc = local ? pb_local_detrending_raster_resident(ctx, &fp.st, &s, int(variable), demId, ids, int(grids.size()), 0, -1, &out)
                   : pb_interpolation_raster_resident(ctx, &fp.st, &s, int(variable), demId, ids, int(grids.size()), 0, -1, &out);
    const bool bad = ok ? failed(ctx, rc) : true;
    for (pb_raster_id t : temporaries) pb_raster_free(ctx, t);
    if (bad) {
        if (reuse && rc != PB_ERR_NO_VALID_CELL) outputGrid->emptyGrid();      // the call failed before it wrote the grid
        return false;                        // incl. PB_ERR_NO_VALID_CELL <=> updateMinMaxRasterGrid false (gis.cpp:601-603)
    }
    outputGrid->minimum = out.minimum;
    outputGrid->maximum = out.maximum;
    if (!outputGrid->colorScale->isFixedRange()) outputGrid->colorScale->setRange(out.minimum, out.maximum);   // gis.cpp:606-611
    return true;
}

/* kriging: the reference API is stateful (one system, file-static); keep that shape */
pb_kriging_id g_kriging = -1;
std::vector<double> g_lastXY(2);
double g_lastResult = NODATA, g_lastRmse = NODATA;

}  // namespace

bool setDevice(int device)
{
    if (device < 0 || device >= pb_device_count()) { g_error = "device out of range"; return false; }
    g_device = device;
    return true;
}
const std::string& lastError() { return g_error; }

bool lastTiming(pb_timing* t)
{
    auto it = g_ctx.find(g_device);
    return it != g_ctx.end() && t && pb_last_timing(it->second, t) == PB_OK;
}

void setRasterCache(bool enabled, size_t limitBytes)
{
    g_cacheEnabled = enabled;
    if (limitBytes) g_cacheLimit = limitBytes;
    if (!enabled) invalidateAll();
}

void invalidate(const gis::Crit3DRasterGrid* grid)
{
    if (!grid) return;
    for (auto& dev : g_cache) {
        auto it = g_ctx.find(dev.first);
        if (it == g_ctx.end()) continue;
        std::vector<CachedGrid>& v = dev.second;
        for (size_t i = 0; i < v.size();)
            if (v[i].owner == grid || v[i].rows == grid->value) dropEntry(it->second, v, i);
            else i++;
    }
}

void invalidateAll()
{
    for (auto& dev : g_cache) {
        auto it = g_ctx.find(dev.first);
        if (it == g_ctx.end()) continue;
        while (!dev.second.empty()) dropEntry(it->second, dev.second, dev.second.size() - 1);
    }
}

bool interpolationRaster(std::vector<Crit3DInterpolationDataPoint>& dataPoints,
                         Crit3DInterpolationSettings& interpolationSettings, Crit3DMeteoSettings* meteoSettings,
                         gis::Crit3DRasterGrid* outputGrid, gis::Crit3DRasterGrid& raster, meteoVariable variable,
                         bool /*isParallelComputing*/)
{
    return rasterCall(false, dataPoints, interpolationSettings, meteoSettings, outputGrid, raster, variable);
}

bool interpolationRasterLocalDetrending(std::vector<Crit3DInterpolationDataPoint>& interpolationPoints,
                                        Crit3DInterpolationSettings& interpolationSettings,
                                        Crit3DMeteoSettings* meteoSettings, gis::Crit3DRasterGrid* outputGrid,
                                        gis::Crit3DRasterGrid& raster, meteoVariable variable)
{
    // Project::interpolationDemLocalDetrending: current combination = selected combination before the loop (project.cpp:3172-3173)
    interpolationSettings.setCurrentCombination(interpolationSettings.getSelectedCombination());
    return rasterCall(true, interpolationPoints, interpolationSettings, meteoSettings, outputGrid, raster, variable);
}

namespace {
typedef double (*fit1d_t)(double, std::vector<double>&);
fit1d_t fitFunctionPtr(int f)
{
    switch (f) {
    case PB_FIT_PIECEWISE_TWO: return lapseRatePiecewise_two;
    case PB_FIT_PIECEWISE_THREE_FREE: return lapseRatePiecewise_three_free;
    case PB_FIT_PIECEWISE_THREE: return lapseRatePiecewise_three;
    case PB_FIT_LINEAR: return functionLinear_intercept;
    default: return nullptr;
    }
}

/* pb_settings -> the fields preInterpolation mutates in Crit3DInterpolationSettings */
void writeBackSettings(const pb_settings& s, Crit3DInterpolationSettings& st, const pb_kh_series& kh)
{
    st.setPrecipitationAllZero(s.precipitationAllZero != 0);
    st.setTopoDist_Kh(s.topoDist_Kh);
    st.setPointsBoundingBoxArea(s.pointsBoundingBoxArea);
    if (s.hasPointsRange) st.setPointsRange(s.pointsRangeMin, s.pointsRangeMax);
    Crit3DProxyCombination cur = st.getCurrentCombination();
    while (int(cur.getProxySize()) < s.nProxy) { cur.addProxyActive(false); cur.addProxySignificant(false); }
    for (int i = 0; i < s.nProxy; i++) {
        cur.setProxyActive(i, s.currentActive[i] != 0);
        cur.setProxySignificant(i, s.currentSignificant[i] != 0);
        Crit3DProxy* p = st.getProxy(i);
        const pb_proxy& q = s.proxy[i];
        p->setInversionIsSignificative(q.inversionIsSignificative != 0);
        p->setRegressionR2(q.regressionR2);
        p->setRegressionSlope(q.regressionSlope);
        p->setRegressionIntercept(q.regressionIntercept);
        p->setLapseRateH0(q.lapseRateH0);
        p->setLapseRateH1(q.lapseRateH1);
        p->setLapseRateT0(q.lapseRateT0);
        p->setLapseRateT1(q.lapseRateT1);
        p->setInversionLapseRate(q.inversionLapseRate);
        if (s.useMultipleDetrending) {
            std::vector<double> range;
            for (int j = 0; j < q.nFitParams; j++) range.push_back(q.fitMin[j]);
            for (int j = 0; j < q.nFitParams; j++) range.push_back(q.fitMax[j]);
            p->setFittingParametersRange(range);
            std::vector<std::vector<double>> combos;
            for (int k = 0; k < q.nFirstGuessCombinations; k++)
                combos.push_back(std::vector<double>(q.firstGuessCombinations[k], q.firstGuessCombinations[k] + q.nFitParams));
            p->setFirstGuessCombinations(combos);
        }
    }
    cur.setUseThermalInversion(s.currentUseThermalInversion != 0);
    st.setCurrentCombination(cur);
    if (s.useMultipleDetrending) {
        st.clearFitting();
        for (int i = 0; i < s.nFitFunctions; i++)
            if (fit1d_t f = fitFunctionPtr(s.fitFunction[i])) st.addFittingFunction(f);
        std::vector<std::vector<double>> params;
        for (int i = 0; i < s.nFit; i++) params.push_back(std::vector<double>(s.fitParams[i], s.fitParams[i] + s.fitNrParams[i]));
        st.setFittingParameters(params);
    }
    if (kh.n > 0) {
        st.initializeKhSeries();
        for (int i = 0; i < kh.n; i++) st.addToKhSeries(kh.kh[i], kh.error[i]);
    }
}
}  // namespace

namespace {
/* std::vector<Crit3DMacroArea> <-> pb_macro_area[] (interpolationSettings.h:149-182) */
struct FlatAreas {
    std::vector<pb_macro_area> a;
    std::vector<std::vector<int32_t>> pts;
    std::vector<std::vector<float>> cells;
};

void flattenArea(const Crit3DMacroArea& area, int nProxy, pb_macro_area& A, std::vector<int32_t>& pts, std::vector<float>& cells)
{
    memset(&A, 0, sizeof(A));
    std::vector<int> mp = area.getMeteoPoints();
    pts.assign(mp.begin(), mp.end());
    cells = area.getAreaCellsDEM();
    A.nMeteoPoints = int(pts.size());
    A.meteoPoints = pts.empty() ? nullptr : pts.data();
    A.nAreaCells = int64_t(cells.size());
    A.areaCellsDEM = cells.empty() ? nullptr : cells.data();
    Crit3DProxyCombination c = area.getCombination();
    for (int i = 0; i < nProxy && i < int(c.getProxySize()) && i < PB_MAX_PROXY; i++) {
        A.active[i] = c.isProxyActive(i);
        A.significant[i] = c.isProxySignificant(i);
    }
    A.useThermalInversion = c.getUseThermalInversion();
    std::vector<std::vector<double>> par = area.getParameters();
    A.nParameters = int(std::min(par.size(), size_t(PB_MAX_PROXY)));
    for (int i = 0; i < A.nParameters; i++) {
        A.nrParams[i] = int(std::min(par[i].size(), size_t(PB_MAX_FIT_PARAMS)));
        for (int j = 0; j < A.nrParams[i]; j++) A.parameters[i][j] = par[i][j];
    }
}

void flattenAreas(const Crit3DInterpolationSettings& st, int nProxy, FlatAreas& f)
{
    const std::vector<Crit3DMacroArea>& areas = st.getMacroAreas();
    const size_t n = areas.size();
    f.a.resize(n); f.pts.resize(n); f.cells.resize(n);
    for (size_t i = 0; i < n; i++) flattenArea(areas[i], nProxy, f.a[i], f.pts[i], f.cells[i]);
}

/* what glocalDetrendingFitting stores in the areas it fitted (interpolation.cpp:2283-2285, 2291) */
void writeBackAreas(const FlatAreas& f, int nProxy, Crit3DInterpolationSettings& st)
{
    std::vector<Crit3DMacroArea> areas = st.getMacroAreas();
    for (size_t i = 0; i < areas.size() && i < f.a.size(); i++) {
        const pb_macro_area& A = f.a[i];
        if (A.nMeteoPoints <= 0) continue;
        std::vector<std::vector<double>> par;
        for (int k = 0; k < A.nParameters; k++) par.push_back(std::vector<double>(A.parameters[k], A.parameters[k] + A.nrParams[k]));
        areas[i].setParameters(par);
        Crit3DProxyCombination c;
        for (int k = 0; k < nProxy; k++) { c.addProxyActive(A.active[k] != 0); c.addProxySignificant(A.significant[k] != 0); }
        c.setUseThermalInversion(A.useThermalInversion != 0);
        areas[i].setCombination(c);
    }
    st.setMacroAreas(areas);
}
}  // namespace

bool preInterpolation(std::vector<Crit3DInterpolationDataPoint>& myPoints, Crit3DInterpolationSettings& interpolationSettings,
                      Crit3DMeteoSettings* meteoSettings, Crit3DClimateParameters* climateParameters,
                      std::vector<Crit3DMeteoPoint>& meteoPoints, meteoVariable myVar, Crit3DTime myTime, std::string& errorStr)
{
    pb_context* ctx = context();
    if (!ctx) { errorStr = g_error; return false; }
    pb_settings s;
    std::vector<pb_raster> grids;
    gis::Crit3DRasterGrid* dem = interpolationSettings.getCurrentDEM();
    if (!flattenSettings(interpolationSettings, meteoSettings, s, grids, dem)) { errorStr = g_error; return false; }
    if (climateParameters) s.climateLapseRate = climateParameters->getClimateLapseRate(myVar, myTime);
    else s.climateLapseRate = 0.f;
    FlatPoints fp;
    flattenPoints(myPoints, s.nProxy, fp);
    TdMapScope tdMaps(ctx);
    if (!tdMaps.fromPoints(myPoints, s.useTD != 0)) { errorStr = g_error; return false; }
    const int n = fp.st.n;
    // the meteo point behind each interpolation point: Crit3DInterpolationDataPoint::index (passDataToInterpolation :516)
    std::vector<float> current(n);
    std::vector<uint8_t> inside(n, 1);
    bool haveMeteo = !meteoPoints.empty();
    for (int i = 0; i < n && haveMeteo; i++) {
        const int k = myPoints[i].index;
        if (k < 0 || k >= int(meteoPoints.size())) { haveMeteo = false; break; }
        current[i] = meteoPoints[k].currentValue;
        inside[i] = meteoPoints[k].isInsideDem;
    }
    pb_cv_points cv;
    memset(&cv, 0, sizeof(cv));
    if (haveMeteo) { cv.currentValue = current.data(); cv.isInsideDem = inside.data(); }
    cv.excludeOutsideDem = interpolationSettings.getUseExcludeStationsOutsideDEM();
    pb_raster_id demId = -1;
    const bool needDem = s.useTD && !s.useLocalDetrending && !s.useGlocalDetrending;
    if (needDem && dem != nullptr && dem->isLoaded) {
        pb_raster r = rasterOf(*dem);
        if (failed(ctx, pb_raster_upload(ctx, &r, &demId))) { errorStr = g_error; return false; }
    }
    std::vector<uint8_t> keep(n, 1);
    pb_kh_series kh;
    memset(&kh, 0, sizeof(kh));
    if (s.useGlocalDetrending && s.useMultipleDetrending && getUseDetrendingVar(myVar)) {
        // glocalDetrendingFitting (:2471-2474): the points stay as they are, the macro areas receive their fit
        FlatAreas fa;
        flattenAreas(interpolationSettings, s.nProxy, fa);
        const int rcg = pb_glocal_detrending_fitting(ctx, &fp.st, &s, int(myVar), fa.a.data(), int(fa.a.size()));
        if (demId >= 0) pb_raster_free(ctx, demId);
        if (failed(ctx, rcg)) { errorStr = g_error; return false; }
        writeBackSettings(s, interpolationSettings, kh);
        writeBackAreas(fa, s.nProxy, interpolationSettings);
        return true;
    }
    const int rc = pb_pre_interpolation_ex(ctx, &fp.st, &s, int(myVar), &cv, demId, keep.data(), &kh);
    if (demId >= 0) pb_raster_free(ctx, demId);
    if (rc == PB_ERR_FIT) { errorStr = "Error in function preInterpolation"; return false; }
    if (failed(ctx, rc)) { errorStr = g_error; return false; }
    writeBackSettings(s, interpolationSettings, kh);
    std::vector<Crit3DInterpolationDataPoint> out;
    out.reserve(n);
    for (int i = 0; i < n; i++)
        if (keep[i]) {
            myPoints[i].value = fp.value[i];
            out.push_back(myPoints[i]);
        }
    if (int(out.size()) != n) myPoints = out;
    return true;
}

bool interpolationRasterGlocalDetrending(std::vector<Crit3DInterpolationDataPoint>& interpolationPoints,
                                         Crit3DInterpolationSettings& interpolationSettings, Crit3DMeteoSettings* meteoSettings,
                                         gis::Crit3DRasterGrid* outputGrid, gis::Crit3DRasterGrid& raster, meteoVariable variable,
                                         const std::vector<Crit3DMeteoPoint>* meteoPoints)
{
    if (!getUseDetrendingVar(variable) || !interpolationSettings.getUseGlocalDetrending()) return false;     // project.cpp:3264
    interpolationSettings.setCurrentCombination(interpolationSettings.getSelectedCombination());               // :3279-3280
    gis::Crit3DRasterHeader header = *(raster.header);
    if (!outputGrid->initializeGrid(header)) return false;                                                     // :3296-3297
    pb_context* ctx = context();
    if (!ctx) return false;
    pb_settings s;
    std::vector<pb_raster> grids;
    if (!flattenSettings(interpolationSettings, meteoSettings, s, grids, &raster)) return false;
    FlatPoints fp;
    flattenPoints(interpolationPoints, s.nProxy, fp);
    FlatAreas fa;
    flattenAreas(interpolationSettings, s.nProxy, fa);
    pb_raster dem = rasterOf(raster);
    pb_raster_id demId = -1;
    std::vector<pb_raster_id> ids(grids.size(), -1);
    auto release = [&]() {
        if (demId >= 0) pb_raster_free(ctx, demId);
        for (pb_raster_id id : ids) if (id >= 0 && id != demId) pb_raster_free(ctx, id);
    };
    if (failed(ctx, pb_raster_upload(ctx, &dem, &demId))) return false;
    for (size_t g = 0; g < grids.size(); g++) {
        if (grids[g].rows == dem.rows && grids[g].data == dem.data) { ids[g] = demId; continue; }     // the elevation proxy IS the DEM
        if (failed(ctx, pb_raster_upload(ctx, &grids[g], &ids[g]))) { release(); return false; }
    }
    pb_raster_out out;
    memset(&out, 0, sizeof(out));
    out.rows = outputGrid->value;
    int rc;
    if (s.useTD && meteoPoints != nullptr) {
        // macroAreaDetrending's kh search walks Project::meteoPoints (interpolation.cpp:1811-1826)
        FlatMeteo fm;
        flattenMeteo(*meteoPoints, s.nProxy, fm);
        pb_glocal_meteo gm;
        memset(&gm, 0, sizeof(gm));
        gm.meteo = &fm.st;
        gm.observed = fm.value.data();
        gm.cv = &fm.cv;
        TdMapScope tdMaps(ctx);
        if (!tdMaps.fromPoints(interpolationPoints, true)) { release(); return false; }
        rc = pb_glocal_detrending_raster_td(ctx, &fp.st, &s, int(variable), fa.a.data(), int(fa.a.size()), -1, &gm, demId, ids.data(),
                                            int(ids.size()), 0, &out, nullptr);
    } else
        rc = pb_glocal_detrending_raster(ctx, &fp.st, &s, int(variable), fa.a.data(), int(fa.a.size()), demId, ids.data(),
                                         int(ids.size()), 0, &out);
    release();
    if (rc != PB_OK && rc != PB_ERR_NO_VALID_CELL) { failed(ctx, rc); return false; }
    pb_kh_series kh;
    memset(&kh, 0, sizeof(kh));
    writeBackSettings(s, interpolationSettings, kh);
    writeBackAreas(fa, s.nProxy, interpolationSettings);
    outputGrid->minimum = out.minimum;
    outputGrid->maximum = out.maximum;
    if (rc == PB_ERR_NO_VALID_CELL) { failed(ctx, rc); return false; }       // isOk = false (:3358-3359)
    return true;
}

bool computeResidualsGlocalDetrending(meteoVariable myVar, const Crit3DMacroArea& myArea, int /*elevationPos*/,
                                      std::vector<Crit3DMeteoPoint>& meteoPoints,
                                      std::vector<Crit3DInterpolationDataPoint>& interpolationPoints,
                                      Crit3DInterpolationSettings& interpolationSettings, Crit3DMeteoSettings* meteoSettings,
                                      bool excludeOutsideDem, bool excludeSupplemental)
{
    pb_context* ctx = context();
    if (!ctx) return false;
    pb_settings s;
    std::vector<pb_raster> grids;
    if (!flattenSettings(interpolationSettings, meteoSettings, s, grids, interpolationSettings.getCurrentDEM())) return false;
    FlatPoints fp;
    flattenPoints(interpolationPoints, s.nProxy, fp);
    pb_macro_area A;
    std::vector<int32_t> pts;
    std::vector<float> cells;
    flattenArea(myArea, s.nProxy, A, pts, cells);
    if (A.nAreaCells <= 0) { A.nAreaCells = 2; static const float one[2] = {0.f, 1.f}; A.areaCellsDEM = one; }   // this call never reads the cells
    // one Crit3DMeteoPoint per entry: position, lapse code, proxy values, current value; only accepted points are cross-validated
    const int n = int(meteoPoints.size());
    FlatMeteo fm;
    flattenMeteo(meteoPoints, s.nProxy, fm);
    std::vector<float> res(n);
    gis::Crit3DRasterGrid* dem = interpolationSettings.getCurrentDEM();
    if (s.useTD && dem != nullptr && dem->isLoaded) {
        TdMapScope tdMaps(ctx);
        pb_raster_id demId = -1;
        bool tmp = false;
        if (!residentId(ctx, dem, rasterOf(*dem), &demId, &tmp)) return false;
        if (tmp) tdMaps.temporaries.push_back(demId);
        if (!tdMaps.fromPoints(interpolationPoints, true)) return false;
        if (failed(ctx, pb_compute_residuals_glocal_td(ctx, &fm.st, fm.value.data(), &fm.cv, &fp.st, &s, int(myVar), &A, 1, demId,
                                                       excludeOutsideDem, excludeSupplemental, res.data())))
            return false;
    } else if (failed(ctx, pb_compute_residuals_glocal(ctx, &fm.st, fm.value.data(), &fm.cv, &fp.st, &s, int(myVar), &A, 1,
                                                       excludeOutsideDem, excludeSupplemental, res.data())))
        return false;
    for (int i = 0; i < n; i++) {
        if (isEqual(res[i], NODATA)) continue;
        if (isEqual(meteoPoints[i].residual, NODATA)) meteoPoints[i].residual = res[i];
        else meteoPoints[i].residual += res[i];
    }
    return true;
}

bool computeResiduals(meteoVariable myVar, std::vector<Crit3DMeteoPoint>& meteoPoints,
                      const std::vector<Crit3DInterpolationDataPoint>& interpolationPoints,
                      Crit3DInterpolationSettings& interpolationSettings, Crit3DMeteoSettings* meteoSettings,
                      bool excludeOutsideDem, bool excludeSupplemental)
{
    if (myVar == noMeteoVar) return false;
    pb_context* ctx = context();
    if (!ctx) return false;
    pb_settings s;
    std::vector<pb_raster> grids;
    if (!flattenSettings(interpolationSettings, meteoSettings, s, grids, nullptr)) return false;
    FlatPoints fp;
    flattenPoints(interpolationPoints, s.nProxy, fp);
    // topographic distance (computeDistances :226-251): the current DEM resident, the points' precomputed maps registered
    gis::Crit3DRasterGrid* dem = interpolationSettings.getCurrentDEM();
    const bool td = s.useTD && !s.useLocalDetrending && dem != nullptr && dem->isLoaded;
    TdMapScope tdMaps(ctx);
    pb_raster_id demId = -1;
    bool demTemporary = false;
    if (td) {
        if (!residentId(ctx, dem, rasterOf(*dem), &demId, &demTemporary)) return false;
        if (demTemporary) tdMaps.temporaries.push_back(demId);
        if (!tdMaps.fromPoints(interpolationPoints, true)) return false;
    }
    // targets = the meteo points that computeResiduals would evaluate (spatialControl.cpp:113-121)
    std::vector<float> x, y, z, est;
    std::vector<double> pv;
    std::vector<size_t> who;
    for (size_t i = 0; i < meteoPoints.size(); i++) {
        meteoPoints[i].residual = NODATA;
        if (!meteoPoints[i].active) continue;
        bool isValid = (!excludeSupplemental || checkLapseRateCode(meteoPoints[i].lapseRateCode, interpolationSettings.getUseLapseRateCode(), false));
        isValid = (isValid && (!excludeOutsideDem || meteoPoints[i].isInsideDem));
        if (!(isValid && meteoPoints[i].quality == quality::accepted)) continue;
        who.push_back(i);
        x.push_back(float(meteoPoints[i].point.utm.x));
        y.push_back(float(meteoPoints[i].point.utm.y));
        z.push_back(float(meteoPoints[i].point.z));
        std::vector<double> v = meteoPoints[i].getProxyValues();
        for (int k = 0; k < s.nProxy; k++) pv.push_back(k < int(v.size()) ? v[k] : double(NODATA));
    }
    est.resize(who.size());
    if (!who.empty()) {
        const int rc = td ? pb_interpolate_points_td(ctx, &fp.st, &s, int(myVar), x.data(), y.data(), z.data(), pv.data(),
                                                     int(who.size()), 0, demId, est.data())
                          : pb_interpolate_points(ctx, &fp.st, &s, int(myVar), x.data(), y.data(), pv.data(), int(who.size()), 0,
                                                  est.data());
        if (failed(ctx, rc)) return false;
    }
    for (size_t k = 0; k < who.size(); k++) {
        Crit3DMeteoPoint& mp = meteoPoints[who[k]];
        float myValue = mp.currentValue, interpolatedValue = est[k];
        if (myVar == precipitation || myVar == dailyPrecipitation) {
            if (myValue != NODATA && myValue < meteoSettings->getRainfallThreshold()) myValue = 0.;
            if (interpolatedValue != NODATA && interpolatedValue < meteoSettings->getRainfallThreshold()) interpolatedValue = 0.;
        }
        if ((interpolatedValue != NODATA) && (myValue != NODATA)) mp.residual = myValue - interpolatedValue;
    }
    return true;
}

bool computeResidualsLocalDetrending(meteoVariable myVar, const Crit3DTime& myTime, std::vector<Crit3DMeteoPoint>& meteoPoints,
                                     std::vector<Crit3DInterpolationDataPoint>& interpolationPoints,
                                     Crit3DInterpolationSettings& interpolationSettings, Crit3DMeteoSettings* meteoSettings,
                                     Crit3DClimateParameters* climateParameters, bool excludeOutsideDem, bool excludeSupplemental)
{
    (void)myTime; (void)climateParameters;
    if (myVar == noMeteoVar) return false;
    pb_context* ctx = context();
    if (!ctx) return false;
    pb_settings s;
    std::vector<pb_raster> grids;
    if (!flattenSettings(interpolationSettings, meteoSettings, s, grids, nullptr)) return false;
    FlatPoints fp;
    flattenPoints(interpolationPoints, s.nProxy, fp);
    // targets = the meteo points computeResidualsLocalDetrending evaluates (spatialControl.cpp:176-183)
    std::vector<float> x, y, est;
    std::vector<double> pv;
    std::vector<size_t> who;
    for (size_t i = 0; i < meteoPoints.size(); i++) {
        meteoPoints[i].residual = NODATA;
        if (!meteoPoints[i].active) continue;
        bool isValid = (!excludeSupplemental || checkLapseRateCode(meteoPoints[i].lapseRateCode, interpolationSettings.getUseLapseRateCode(), false));
        isValid = (isValid && (!excludeOutsideDem || meteoPoints[i].isInsideDem));
        if (!(isValid && meteoPoints[i].quality == quality::accepted)) continue;
        who.push_back(i);
        x.push_back(float(meteoPoints[i].point.utm.x));
        y.push_back(float(meteoPoints[i].point.utm.y));
        std::vector<double> v = meteoPoints[i].getProxyValues();
        for (int k = 0; k < s.nProxy; k++) pv.push_back(k < int(v.size()) ? v[k] : double(NODATA));
    }
    est.resize(who.size());
    if (!who.empty()) {
        // the inner localSelection / interpolate run with excludeSupplemental = false (the shadowing local, :186)
        const int rc = pb_local_detrending_points(ctx, &fp.st, &s, int(myVar), x.data(), y.data(), pv.data(), int(who.size()), 0, est.data());
        if (failed(ctx, rc)) return false;
    }
    for (size_t k = 0; k < who.size(); k++) {
        Crit3DMeteoPoint& mp = meteoPoints[who[k]];
        float myValue = mp.currentValue, interpolatedValue = est[k];
        if (myVar == precipitation || myVar == dailyPrecipitation) {
            if (myValue != NODATA && myValue < meteoSettings->getRainfallThreshold()) myValue = 0.;
            if (interpolatedValue != NODATA && interpolatedValue < meteoSettings->getRainfallThreshold()) interpolatedValue = 0.;
        }
        if ((interpolatedValue != NODATA) && (myValue != NODATA)) mp.residual = myValue - interpolatedValue;
    }
    return true;
}

bool spatialQualityControl(meteoVariable myVar, std::vector<Crit3DMeteoPoint>& meteoPoints,
                           Crit3DInterpolationSettings& interpolationSettings, Crit3DMeteoSettings* meteoSettings,
                           Crit3DClimateParameters* climateParameters, const Crit3DTime& myTime, std::string& errorStr)
{
    pb_context* ctx = context();
    if (!ctx) { errorStr = g_error; return false; }
    pb_settings s;
    std::vector<pb_raster> grids;
    gis::Crit3DRasterGrid* dem = interpolationSettings.getCurrentDEM();
    if (!flattenSettings(interpolationSettings, meteoSettings, s, grids, dem)) { errorStr = g_error; return false; }
    s.climateLapseRate = climateParameters ? climateParameters->getClimateLapseRate(myVar, myTime) : 0.f;
    // the points passDataToInterpolation would take (spatialControl.cpp:510-516): active, accepted, selected if a selection exists
    bool isSelection = false;
    for (const auto& mp : meteoPoints) if (mp.selected) { isSelection = true; break; }
    std::vector<size_t> who;
    for (size_t i = 0; i < meteoPoints.size(); i++)
        if (meteoPoints[i].active && meteoPoints[i].quality == quality::accepted && (!isSelection || meteoPoints[i].selected)) who.push_back(i);
    const int n = int(who.size()), P = s.nProxy, Pp = std::max(P, 1);
    if (n == 0) return true;
    std::vector<double> x(n), y(n), z(n);
    std::vector<float> value(n), proxy(size_t(n) * Pp, float(NODATA)), suspect(size_t(n) * Pp, float(NODATA));
    std::vector<int32_t> code(n);
    std::vector<uint8_t> inside(n), wrong(n, 0);
    for (int k = 0; k < n; k++) {
        const Crit3DMeteoPoint& mp = meteoPoints[who[k]];
        x[k] = mp.point.utm.x; y[k] = mp.point.utm.y; z[k] = mp.point.z;
        value[k] = mp.currentValue;
        code[k] = int(mp.lapseRateCode);
        inside[k] = mp.isInsideDem;
        for (int q = 0; q < P && q < int(mp.proxyValues.size()); q++) proxy[size_t(k) * P + q] = mp.proxyValues[q];
        // second pass: the reference reads meteoPoints[k] (k = position in the suspect list), spatialControl.cpp:393
        if (size_t(k) < meteoPoints.size())
            for (int q = 0; q < P && q < int(meteoPoints[k].proxyValues.size()); q++) suspect[size_t(k) * P + q] = meteoPoints[k].proxyValues[q];
    }
    pb_stations st;
    memset(&st, 0, sizeof(st));
    st.n = n; st.nProxy = P;
    st.x = x.data(); st.y = y.data(); st.z = z.data(); st.value = value.data();
    st.lapseRateCode = code.data(); st.proxyValues = proxy.data();
    pb_cv_points cv;
    memset(&cv, 0, sizeof(cv));
    cv.isInsideDem = inside.data();
    cv.excludeOutsideDem = interpolationSettings.getUseExcludeStationsOutsideDEM();
    cv.suspectProxyValues = suspect.data();
    pb_raster_id demId = -1;
    const bool needDem = s.useTD && !s.useLocalDetrending && !s.useGlocalDetrending;
    if (needDem && dem != nullptr && dem->isLoaded) {
        pb_raster r = rasterOf(*dem);
        if (failed(ctx, pb_raster_upload(ctx, &r, &demId))) { errorStr = g_error; return false; }
    }
    const int rc = pb_spatial_quality_control(ctx, &st, &s, int(myVar), &cv, demId, wrong.data());
    if (demId >= 0) pb_raster_free(ctx, demId);
    if (rc == PB_ERR_FIT) { errorStr = "Error in function preInterpolation"; return false; }
    if (failed(ctx, rc)) { errorStr = g_error; return false; }
    pb_kh_series kh;
    memset(&kh, 0, sizeof(kh));
    writeBackSettings(s, interpolationSettings, kh);
    for (int k = 0; k < n; k++)
        if (wrong[k]) meteoPoints[who[k]].quality = quality::wrong_spatial;
    return true;
}

bool krigingVariogram(double* myPos, double* myVal, int nrItems, short myMode, double myRange, double myNugget,
                      double mySill, double mySlope)
{
    pb_context* ctx = context();
    if (!ctx) return false;
    krigingFreeMemory();
    return !failed(ctx, pb_kriging_setup(ctx, myPos, myVal, nrItems, myMode, myRange, myNugget, mySill, mySlope, 0, &g_kriging));
}

bool krigingSetWeight(double x_p, double y_p)
{
    pb_context* ctx = context();
    if (!ctx || g_kriging < 0) return false;
    g_lastXY[0] = x_p;
    g_lastXY[1] = y_p;
    return !failed(ctx, pb_kriging_points(ctx, g_kriging, g_lastXY.data(), 1, &g_lastResult, &g_lastRmse));
}

double krigingResult() { return g_lastResult; }
double krigingRMSE() { return g_lastRmse; }

void krigingFreeMemory()
{
    if (g_kriging >= 0) {
        pb_context* ctx = context();
        if (ctx) pb_kriging_free(ctx, g_kriging);
        g_kriging = -1;
    }
}

bool krigingEstimateVariogram(float* myDist, float* mySemiVar, int sizeMyVar, int /*nrMyPoints*/, float myMaxDistance,
                              double* mySill, double* myNugget, double* myRange, double* mySlope, TkrigingMode* myMode,
                              int /*nrPointData*/)
{
    if (!myDist || !mySemiVar || !mySill || !myNugget || !myRange || !mySlope || !myMode) return false;
    // bins beyond myMaxDistance do not take part (count 0)
    std::vector<int32_t> count(size_t(std::max(sizeMyVar, 0)), 1);
    for (int i = 0; i < sizeMyVar; i++) if (myMaxDistance > 0 && myDist[i] > myMaxDistance) count[i] = 0;
    int mode = int(*myMode);
    if (mode < int(KRIGING_SPHERICAL) || mode > int(KRIGING_LINEAR)) mode = 0;
    int modeOut = 0;
    const int rc = pb_kriging_fit_variogram(myDist, mySemiVar, count.data(), sizeMyVar, mode, myRange, myNugget, mySill, mySlope,
                                            &modeOut, nullptr);
    if (rc != PB_OK) { g_error = "variogram fit failed"; return false; }
    *myMode = TkrigingMode(modeOut);
    return true;
}

bool krigingRaster(gis::Crit3DRasterGrid& raster, gis::Crit3DRasterGrid* estimateGrid, gis::Crit3DRasterGrid* rmseGrid)
{
    pb_context* ctx = context();
    if (!ctx || g_kriging < 0 || estimateGrid == nullptr) return false;
    if (!estimateGrid->initializeGrid(raster)) return false;
    if (rmseGrid && !rmseGrid->initializeGrid(raster)) return false;
    pb_raster dem = rasterOf(raster);
    pb_raster_id id = -1;
    if (failed(ctx, pb_raster_upload(ctx, &dem, &id))) return false;
    pb_raster_out e, r;
    memset(&e, 0, sizeof(e));
    memset(&r, 0, sizeof(r));
    e.rows = estimateGrid->value;
    if (rmseGrid) r.rows = rmseGrid->value;
    const int rc = pb_kriging_raster(ctx, g_kriging, id, 0, -1, &e, rmseGrid ? &r : nullptr);
    pb_raster_free(ctx, id);
    if (failed(ctx, rc)) return false;
    gis::updateMinMaxRasterGrid(estimateGrid);
    if (rmseGrid) gis::updateMinMaxRasterGrid(rmseGrid);
    return true;
}

/* radiation::computeRadiationDEM (agrolib/solarRadiation/solarRadiation.cpp:1045-1069) */
bool computeRadiationDEM(Crit3DRadiationSettings* radSettings, const gis::Crit3DRasterGrid& dem, Crit3DRadiationMaps* radiationMaps,
                         const Crit3DTime& myTime, bool /*isParallelComputing*/)
{
    if (radSettings->getAlgorithm() != RADIATION_ALGORITHM_RSUN) return false;
    pb_context* ctx = context();
    if (!ctx) return false;
    // computeRadiationRsun :714-719
    Crit3DTime localTime = myTime;
    if (radSettings->gisSettings->isUTC) localTime = myTime.addSeconds(radSettings->gisSettings->timeZone * 3600);
    pb_radiation_settings s;
    memset(&s, 0, sizeof(s));
    s.realSky = radSettings->getRealSky();
    s.realSkyAlgorithm = int(radSettings->getRealSkyAlgorithm());
    s.shadowing = radSettings->getShadowing();
    s.tiltMode = int(radSettings->getTiltMode());
    s.linkeMode = int(radSettings->getLinkeMode());
    s.albedoMode = int(radSettings->getAlbedoMode());
    // the scalars the reference ends up with for an in-grid cell (computeRadiationDemPoint :958-964): monthly or default Linke;
    // the map accessors only read OUT-of-grid cells (radiationSettings.cpp:131-133, :179-181), i.e. NODATA for every DEM cell
    if (radSettings->getLinkeMode() == PARAM_MODE_MONTHLY) s.linke = radSettings->getLinke(localTime.date.month - 1);
    else s.linke = radSettings->getLinkeMode() == PARAM_MODE_FIXED ? radSettings->getLinkeDefault() : float(NODATA);
    s.albedo = radSettings->getAlbedoMode() == PARAM_MODE_FIXED ? radSettings->getAlbedo() : float(NODATA);
    s.tilt = radSettings->getTilt();
    s.aspect = radSettings->getAspect();
    s.clearSky = radSettings->getClearSky();
    s.demMaximum = dem.maximum;
    s.timeZone = radSettings->gisSettings->timeZone;
    s.year = localTime.date.year; s.month = localTime.date.month; s.day = localTime.date.day;
    s.secondsOfDay = localTime.time;
    // resident copies: the DEM and the four static maps are cached across time steps, the transmissivity map changes every step
    // (its fingerprint differs: uploaded again)
    pb_radiation_maps maps;
    const gis::Crit3DRasterGrid* in[6] = {&dem, radiationMaps->latMap, radiationMaps->lonMap, radiationMaps->slopeMap, radiationMaps->aspectMap,
                                          radiationMaps->transmissivityMap};
    pb_raster_id ids[6];
    std::vector<pb_raster_id> temporaries;
    for (int k = 0; k < 6; k++) {
        if (!in[k] || !in[k]->isLoaded) { g_error = "a radiation input map is not loaded"; return false; }
        bool tmp = false;
        if (!residentId(ctx, in[k], rasterOf(*in[k]), &ids[k], &tmp)) return false;
        if (tmp) temporaries.push_back(ids[k]);
    }
    maps.dem = ids[0]; maps.lat = ids[1]; maps.lon = ids[2]; maps.slope = ids[3]; maps.aspect = ids[4]; maps.transmissivity = ids[5];
    gis::Crit3DRasterGrid* outGrid[5] = {radiationMaps->sunElevationMap, radiationMaps->beamRadiationMap, radiationMaps->diffuseRadiationMap,
                                         radiationMaps->reflectedRadiationMap, radiationMaps->globalRadiationMap};
    pb_raster_out outs[5];
    memset(outs, 0, sizeof(outs));
    for (int k = 0; k < 5; k++) outs[k].rows = outGrid[k]->value;
    pb_radiation_out ro;
    memset(&ro, 0, sizeof(ro));
    ro.sunElevation = &outs[0]; ro.beam = &outs[1]; ro.diffuse = &outs[2]; ro.reflected = &outs[3]; ro.global = &outs[4];
    const int rc = pb_compute_radiation_dem(ctx, &s, &maps, &ro);
    const bool bad = failed(ctx, rc);
    for (pb_raster_id t : temporaries) pb_raster_free(ctx, t);
    if (bad) return false;
    // updateRadiationMaps (:1072-1089)
    gis::updateMinMaxRasterGrid(radiationMaps->transmissivityMap);
    for (int k = 0; k < 5; k++) {
        outGrid[k]->minimum = outs[k].minimum;
        outGrid[k]->maximum = outs[k].maximum;
        if (!outGrid[k]->colorScale->isFixedRange()) outGrid[k]->colorScale->setRange(outs[k].minimum, outs[k].maximum);
        outGrid[k]->setMapTime(myTime);
    }
    radiationMaps->transmissivityMap->setMapTime(myTime);
    radiationMaps->setComputed(true);
    return true;
}

}  // namespace pragab200
