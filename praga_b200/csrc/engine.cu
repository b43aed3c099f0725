// engine.cu — C ABI of libpraga_b200.so (include/praga_b200.h): context, device-resident raster cache,
// flattening of the POD settings into the kernel model, staging and launches.  No CPU fallback: every compute
// entry point needs a CUDA device and fails with PB_ERR_CUDA otherwise.
#include <math.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#include "context.cuh"
#include "local.cuh"
#include "pre.cuh"
#include "shepard.cuh"
#include "nvtx.cuh"

namespace pb {
// idw_kernels.cu
cudaError_t launchCompactValid(const float* dem, long long n, float flag, int* blockCount, long long* dTotal,
                               int* validIdx, int phase, cudaStream_t s, int indexBase = 0);
int compactBlocks(long long n);
bool idwResidentFits(int nPaddedStations);
cudaError_t launchCompactFill(const float* dem, long long n, float flag, int* validIdx, int indexBase, long long* dCount,
                              float* outBand, cudaStream_t s);
int tileStations();
cudaError_t launchIdwRaster(const DevStations& st, const DevModel& m, const DevGrid& dem, const int* validIdx,
                            int nTargets, float* out, unsigned* minmax, cudaStream_t s, const long long* nTargetsDev = nullptr,
                            unsigned* workCounter = nullptr, int reserveSms = 0);
cudaError_t launchIdwPoints(const DevStations& st, const DevModel& m, const float* tx, const float* ty,
                            const double* tproxy, int nTargets, float* out, cudaStream_t s);
cudaError_t launchFill(float* out, long long n, float v, cudaStream_t s);
// td_kernels.cu
cudaError_t launchIdwTdRaster(const DevStations& st, const float* sz, const DevModel& m, const DevGrid& dem, int kh,
                              const int* validIdx, int nTargets, float* out, unsigned* minmax, cudaStream_t s,
                              const DevGrid* tdMaps = nullptr);
cudaError_t launchIdwTdPoints(const DevStations& st, const float* sz, const DevModel& m, const DevGrid& dem, int kh,
                              const float* tx, const float* ty, const float* tz, const double* tproxy, int nTargets, float* out,
                              cudaStream_t s, const DevGrid* tdMaps = nullptr);
cudaError_t launchRowStart(const int* validIdx, long long total, int nrRows, int nrCols, long long* rowStart, cudaStream_t s);
cudaError_t launchFillValid(float* out, const int* validIdx, int n, float v, cudaStream_t s);
// peaks.cu
cudaError_t measurePipePeaks(cudaStream_t s, int sms, float* fp32Tflops, float* mufuGops);
}  // namespace pb

using namespace pb;

namespace {
std::string g_createError;
}  // namespace

namespace pb {
int failCtx(pb_context* ctx, int code, const std::string& msg)
{
    if (ctx) ctx->err = msg;
    else g_createError = msg;
    return code;
}
}  // namespace pb

namespace {

int fail(pb_context* ctx, int code, const std::string& msg) { return pb::failCtx(ctx, code, msg); }

// ---- meteoVariable classification (interpolation.cpp:1190-1234, 2540-2558) ----
bool useDetrendingVar(int v)
{
    return v == PB_VAR_AIR_TEMPERATURE || v == PB_VAR_AIR_DEW_TEMPERATURE || v == PB_VAR_DAILY_TAVG ||
           v == PB_VAR_DAILY_TMAX || v == PB_VAR_DAILY_TMIN || v == PB_VAR_DAILY_ET0_HS || v == PB_VAR_ATM_PRESSURE ||
           v == PB_VAR_ELABORATION;
}
bool useTdVar(int v)
{
    return !(v == PB_VAR_PRECIPITATION || v == PB_VAR_DAILY_PRECIPITATION || v == PB_VAR_GLOBAL_IRRADIANCE ||
             v == PB_VAR_ATM_TRANSMISSIVITY || v == PB_VAR_DAILY_GLOBAL_RADIATION || v == PB_VAR_ATM_PRESSURE ||
             v == PB_VAR_DAILY_WATER_TABLE_DEPTH);
}
// computeDistances adds kh * topographicDistance when getUseTD() (= useTD && !useLocalDetrending,
// interpolationSettings.h:251) && getUseTdVar(var) && kh != 0 (interpolation.cpp:226-251)
bool usesTopographicDistance(const pb_settings* s, int v)
{
    return s->useTD && !s->useLocalDetrending && useTdVar(v) && s->topoDist_Kh != 0;
}
bool isPrecVar(int v) { return v == PB_VAR_PRECIPITATION || v == PB_VAR_DAILY_PRECIPITATION; }
int clampModeOf(int v)
{
    switch (v) {
    case PB_VAR_PRECIPITATION: case PB_VAR_DAILY_PRECIPITATION: return CLAMP_PREC;
    case PB_VAR_AIR_REL_HUMIDITY: case PB_VAR_DAILY_RH_AVG: case PB_VAR_DAILY_RH_MIN: case PB_VAR_DAILY_RH_MAX:
        return CLAMP_RH;
    case PB_VAR_DAILY_TRANGE: case PB_VAR_LEAF_WETNESS: case PB_VAR_DAILY_LEAF_WETNESS: case PB_VAR_GLOBAL_IRRADIANCE:
    case PB_VAR_DAILY_GLOBAL_RADIATION: case PB_VAR_ATM_TRANSMISSIVITY: case PB_VAR_WIND_SCALAR_INTENSITY:
    case PB_VAR_WIND_VECTOR_INTENSITY: case PB_VAR_DAILY_WIND_SCALAR_INT_AVG: case PB_VAR_DAILY_WIND_SCALAR_INT_MAX:
    case PB_VAR_DAILY_WIND_VECTOR_INT_AVG: case PB_VAR_DAILY_WIND_VECTOR_INT_MAX: case PB_VAR_ATM_PRESSURE:
        return CLAMP_POS;
    default: return CLAMP_NONE;
    }
}
// checkLapseRateCode (meteo/meteo.cpp:1127-1133)
bool checkLapseRateCode(int code, bool useLapseRateCode, bool useSupplemental)
{
    if (useSupplemental) return !useLapseRateCode || code == PB_LAPSE_PRIMARY || code == PB_LAPSE_SUPPLEMENTAL;
    return !useLapseRateCode || code == PB_LAPSE_PRIMARY || code == PB_LAPSE_SECONDARY;
}

DevGrid devGridOf(const RasterEntry& r)
{
    DevGrid g{};
    g.data = r.d;
    g.nrRows = r.h.nrRows; g.nrCols = r.h.nrCols;
    g.llx = r.h.llx; g.lly = r.h.lly; g.cellSize = r.h.cellSize; g.invCellSize = r.h.invCellSize;
    g.flag = r.h.flag;
    return g;
}

// pb_settings -> DevModel. proxyGridIds: raster ids for settings.proxy[i].gridIndex (resident form).
int buildModel(pb_context* ctx, const pb_settings* s, int variable, const pb_raster_id* proxyGridIds, int nProxyGrids,
               DevModel& m)
{
    memset(&m, 0, sizeof(m));
    if (s->nProxy < 0 || s->nProxy > PB_MAX_PROXY) return fail(ctx, PB_ERR_INVALID, "settings.nProxy out of range");
    if (s->method != PB_METHOD_IDW && s->method != PB_METHOD_SHEPARD && s->method != PB_METHOD_SHEPARD_MODIFIED)
        return fail(ctx, PB_ERR_INVALID, "unknown TInterpolationMethod");
    if (s->useGlocalDetrending) return fail(ctx, PB_ERR_UNSUPPORTED, "glocal detrending runs through pb_glocal_detrending_fitting / pb_glocal_detrending_raster (macro areas needed)");
    m.nProxy = s->nProxy;
    m.useDetrendingVar = useDetrendingVar(variable);
    m.useMultipleDetrending = s->useMultipleDetrending;
    m.useThermalInversion = s->useThermalInversion;
    m.useDoNotRetrend = s->useDoNotRetrend;
    m.useRetrendOnly = s->useRetrendOnly;
    m.clampMode = clampModeOf(variable);
    m.rainfallThreshold = s->rainfallThreshold;
    m.minRegressionR2 = s->minRegressionR2;
    int slots = 0;
    for (int i = 0; i < s->nProxy; i++) {
        const pb_proxy& p = s->proxy[i];
        DevProxy& d = m.proxy[i];
        d.kind = p.kind;
        d.active = s->currentActive[i];
        d.significant = s->currentSignificant[i];
        d.slope = p.regressionSlope;
        d.r2 = p.regressionR2;
        d.H0 = p.lapseRateH0; d.H1 = p.lapseRateH1; d.invLapseRate = p.inversionLapseRate;
        d.inversionIsSignificative = p.inversionIsSignificative;
        d.gridSlot = -1;
        if (p.gridIndex >= 0 && proxyGridIds) {
            if (p.gridIndex >= nProxyGrids) return fail(ctx, PB_ERR_INVALID, "proxy.gridIndex >= nProxyGrids");
            const pb_raster_id id = proxyGridIds[p.gridIndex];
            if (id < 0 || id >= (int)ctx->rasters.size() || !ctx->rasters[id].used)
                return fail(ctx, PB_ERR_INVALID, "proxy grid id is not a live raster");
            d.gridSlot = slots;
            m.grid[slots++] = devGridOf(ctx->rasters[id]);
        }
    }
    if (s->useMultipleDetrending && m.useDetrendingVar) {
        // getMultipleDetrendingValues order (interpolation.cpp:2563-2617): elevation, then the others
        int elevationPos = -1;
        for (int i = 0; i < s->nProxy; i++)
            if (s->proxy[i].kind == PB_PROXY_HEIGHT && s->currentActive[i] && s->currentSignificant[i]) {
                elevationPos = i;
                m.activeSigPos[m.nActiveSig++] = i;
            }
        for (int i = 0; i < s->nProxy; i++)
            if (i != elevationPos && s->currentActive[i] && s->currentSignificant[i]) m.activeSigPos[m.nActiveSig++] = i;
        // functionSum walks fittingFunction and indexes values/parameters with the same counter
        // (furtherMathFunctions.cpp:181-191); more functions than values/parameters is out of bounds in the reference
        m.nFit = s->nFitFunctions;
        if (s->nFit == 0) m.nFit = 0;   // retrend: fittingParameters.size() > 0 required (interpolation.cpp:1310)
        if (m.nFit > m.nActiveSig || m.nFit > s->nFit || m.nFit > PB_MAX_PROXY)
            return fail(ctx, PB_ERR_INVALID, "fitting functions/parameters/significant proxies are inconsistent");
        for (int k = 0; k < m.nFit; k++) {
            m.fitFunction[k] = s->fitFunction[k];
            for (int j = 0; j < PB_MAX_FIT_PARAMS; j++) m.fitParams[k][j] = j < s->fitNrParams[k] ? s->fitParams[k][j] : 0.;
        }
    }
    return PB_OK;
}

// pack the stations the estimator uses into the device layout. `excludeSupplemental` as in computeDistances
// (interpolation.cpp:217-220): excluded points get distance 0 => never weigh => dropped here.
// Layout of one packed station set; h / d = host staging and its device twin (same offsets).
size_t idwStationsBytes(int nStations)
{
    const int T = tileStations();
    const size_t nPadded = size_t((nStations + T - 1) / T) * T;
    const size_t nAl = (size_t(nStations) + 3) / 4 * 4;
    return (nPadded * 24 + nAl * 16 + 16 + 255) / 256 * 256;
}

void packIdwStations(const pb_stations* st, const pb_settings* s, bool excludeSupplemental, unsigned char* h, unsigned char* d,
                     DevStations& ds, double& valueOffset)
{
    std::vector<int> use;
    use.reserve(st->n);
    for (int i = 0; i < st->n; i++) {
        const int code = st->lapseRateCode ? st->lapseRateCode[i] : PB_LAPSE_PRIMARY;
        if (excludeSupplemental && !checkLapseRateCode(code, s->useLapseRateCode != 0, false)) continue;
        use.push_back(i);
    }
    const int n = (int)use.size();
    const int T = tileStations();
    const int nPadded = ((n + T - 1) / T) * T;
    double mean = 0.;
    for (int k = 0; k < n; k++) mean += double(st->value[use[k]]);
    mean = n ? mean / n : 0.;
    valueOffset = mean;
    const size_t offXY = 0, offV = offXY + size_t(nPadded) * 16, offX = offV + size_t(nPadded) * 8;
    const size_t nAl = (size_t(n) + 3) / 4 * 4;
    const size_t offY = offX + nAl * 4, offVal = offY + nAl * 4, offZ = offVal + nAl * 4;
    float4* xy = reinterpret_cast<float4*>(h + offXY);
    float2* v = reinterpret_cast<float2*>(h + offV);
    float* px = reinterpret_cast<float*>(h + offX);
    float* py = reinterpret_cast<float*>(h + offY);
    float* pv = reinterpret_cast<float*>(h + offVal);
    float* pz = reinterpret_cast<float*>(h + offZ);
    for (int k = 0; k < n; k++) {
        const int i = use[k];
        const float x = float(st->x[i]), y = float(st->y[i]);           // float(point->utm.x) (interpolation.cpp:222)
        const float c = float(double(st->value[i]) - mean);
        xy[k] = make_float4(x, x, y, y);
        v[k] = make_float2(c, c);
        px[k] = x; py[k] = y; pv[k] = c;
        pz[k] = float(st->z[i]);
    }
    for (int k = n; k < nPadded; k++) {   // zero-weight padding: rsqrt(2e36)^3 underflows to 0
        xy[k] = make_float4(1e18f, 1e18f, 1e18f, 1e18f);
        v[k] = make_float2(0.f, 0.f);
    }
    ds.xy = reinterpret_cast<const float4*>(d + offXY);
    ds.v = reinterpret_cast<const float2*>(d + offV);
    ds.x = reinterpret_cast<const float*>(d + offX);
    ds.y = reinterpret_cast<const float*>(d + offY);
    ds.val = reinterpret_cast<const float*>(d + offVal);
    ds.z = reinterpret_cast<const float*>(d + offZ);
    ds.n = n;
    ds.nPadded = nPadded;
}

int stageStations(pb_context* ctx, const pb_stations* st, const pb_settings* s, bool excludeSupplemental, DevStations& ds,
                  double& valueOffset)
{
    if (st->n < 0) return fail(ctx, PB_ERR_INVALID, "stations.n < 0");
    const size_t total = idwStationsBytes(st->n);
    CUDA_TRY(ctx, ctx->hStations.reserve(total));
    CUDA_TRY(ctx, ctx->dStations.reserve(total));
    packIdwStations(st, s, excludeSupplemental, ctx->hStations.p, ctx->dStations.p, ds, valueOffset);
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dStations.p, ctx->hStations.p, total, cudaMemcpyHostToDevice, ctx->stream));
    ctx->timing.h2d_bytes += (int64_t)total;
    return PB_OK;
}

// stations for the Shepard estimators (K8): original order, raw values, f32 + f64 coordinates; R0 =
// computeShepardInitialRadius(pointsBoundingBoxArea, ALL points, SHEPARD_AVG_NRPOINTS = 8) (interpolation.cpp:800-803, 812)
int stageShepardStations(pb_context* ctx, const pb_stations* st, const pb_settings* s, bool excludeSupplemental,
                         ShepardStations& ss, float& R0)
{
    if (st->n < 0) return fail(ctx, PB_ERR_INVALID, "stations.n < 0");
    std::vector<int> use;
    use.reserve(st->n);
    for (int i = 0; i < st->n; i++) {
        const int code = st->lapseRateCode ? st->lapseRateCode[i] : PB_LAPSE_PRIMARY;
        if (excludeSupplemental && !checkLapseRateCode(code, s->useLapseRateCode != 0, false)) continue;
        use.push_back(i);
    }
    const size_t n = use.size(), nAl = (n + 3) / 4 * 4;
    const unsigned nrPoints = unsigned(st->n), minPointsNr = 8;
    R0 = float(sqrt((minPointsNr * s->pointsBoundingBoxArea) / (float(3.1415926535898) * nrPoints)));
    const size_t offXd = 0, offYd = offXd + nAl * 8, offX = offYd + nAl * 8, offY = offX + nAl * 4, offV = offY + nAl * 4,
                 offZ = offV + nAl * 4, total = offZ + nAl * 4 + 16;
    CUDA_TRY(ctx, ctx->hStations.reserve(total));
    CUDA_TRY(ctx, ctx->dStations.reserve(total));
    unsigned char* h = ctx->hStations.p;
    double* xd = reinterpret_cast<double*>(h + offXd);
    double* yd = reinterpret_cast<double*>(h + offYd);
    float* px = reinterpret_cast<float*>(h + offX);
    float* py = reinterpret_cast<float*>(h + offY);
    float* pv = reinterpret_cast<float*>(h + offV);
    float* pz = reinterpret_cast<float*>(h + offZ);
    for (size_t k = 0; k < n; k++) {
        const int i = use[k];
        xd[k] = st->x[i]; yd[k] = st->y[i];
        px[k] = float(st->x[i]); py[k] = float(st->y[i]);
        pv[k] = st->value[i];
        pz[k] = float(st->z[i]);
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dStations.p, h, total, cudaMemcpyHostToDevice, ctx->stream));
    ctx->timing.h2d_bytes += (int64_t)total;
    unsigned char* d = ctx->dStations.p;
    ss.xd = reinterpret_cast<const double*>(d + offXd);
    ss.yd = reinterpret_cast<const double*>(d + offYd);
    ss.x = reinterpret_cast<const float*>(d + offX);
    ss.y = reinterpret_cast<const float*>(d + offY);
    ss.value = reinterpret_cast<const float*>(d + offV);
    ss.z = reinterpret_cast<const float*>(d + offZ);
    ss.n = int(n);
    return PB_OK;
}

int stageTdMapsForImpl(pb_context* ctx, const int32_t* pointIndex, int n, const DevGrid** out)
{
    *out = nullptr;
    const pb_context* owner = ctx->parent ? ctx->parent : ctx;      // a lane reads its parent's registry (and rasters)
    if (owner->tdMaps.empty() || n <= 0) return PB_OK;
    std::vector<DevGrid> g(n);
    bool any = false;
    for (int k = 0; k < n; k++) {
        g[k] = DevGrid{};
        auto it = owner->tdMaps.find(pointIndex[k]);
        if (it == owner->tdMaps.end()) continue;
        const pb_raster_id id = it->second;
        if (id < 0 || id >= (int)owner->rasters.size() || !owner->rasters[id].used)
            return fail(ctx, PB_ERR_INVALID, "a registered topographic-distance map is no longer a live raster");
        g[k] = devGridOf(owner->rasters[id]);
        any = true;
    }
    if (!any) return PB_OK;
    CUDA_TRY(ctx, ctx->dTdMaps.reserve(size_t(n) * sizeof(DevGrid)));
    // (pageable source: the call returns once the descriptors sit in the driver's staging buffer)
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dTdMaps.p, g.data(), size_t(n) * sizeof(DevGrid), cudaMemcpyHostToDevice, ctx->stream));
    ctx->timing.h2d_bytes += (int64_t)(size_t(n) * sizeof(DevGrid));
    *out = reinterpret_cast<const DevGrid*>(ctx->dTdMaps.p);
    return PB_OK;
}

int stageTdMapsImpl(pb_context* ctx, const pb_stations* st, const pb_settings* s, bool excludeSupplemental, const DevGrid** out)
{
    *out = nullptr;
    const pb_context* owner = ctx->parent ? ctx->parent : ctx;
    if (owner->tdMaps.empty()) return PB_OK;
    std::vector<int32_t> idx;
    idx.reserve(st->n);
    for (int i = 0; i < st->n; i++) {
        const int code = st->lapseRateCode ? st->lapseRateCode[i] : PB_LAPSE_PRIMARY;
        if (excludeSupplemental && !checkLapseRateCode(code, s->useLapseRateCode != 0, false)) continue;
        idx.push_back(st->index ? st->index[i] : i);
    }
    return stageTdMapsForImpl(ctx, idx.data(), int(idx.size()), out);
}

constexpr size_t kStageSlotBytes = size_t(8) << 20;

// host raster (contiguous or float**) -> device: row chunks gathered into pinned ring slots by the host cores while the
// previous chunks cross PCIe
int uploadRaster(pb_context* ctx, const pb_raster* r, RasterEntry& e)
{
    if (r->h.nrRows <= 0 || r->h.nrCols <= 0) return fail(ctx, PB_ERR_INVALID, "raster has no cells");
    if (!r->data && !r->rows) return fail(ctx, PB_ERR_INVALID, "raster has neither data nor rows");
    const size_t n = size_t(r->h.nrRows) * r->h.nrCols;
    if (n > size_t(0x7fffffff)) return fail(ctx, PB_ERR_UNSUPPORTED, "raster larger than 2^31-1 cells");
    const int rows = r->h.nrRows, cols = r->h.nrCols;
    if (size_t(cols) * sizeof(float) > kStageSlotBytes) return fail(ctx, PB_ERR_UNSUPPORTED, "raster row larger than a staging slot");
    CUDA_TRY(ctx, ctx->ring.init(kStageSlotBytes));
    void* dptr = nullptr;
    CUDA_TRY(ctx, ctx->pool.acquire(n * sizeof(float), &dptr));
    e.d = static_cast<float*>(dptr);
    e.h = r->h;
    e.n = n;
    const int chunkRows = std::max(1, (int)(kStageSlotBytes / (sizeof(float) * cols)));
    for (int r0 = 0; r0 < rows; r0 += chunkRows) {
        const int r1 = std::min(rows, r0 + chunkRows);
        int slot;
        float* stage = reinterpret_cast<float*>(ctx->ring.acquire(&slot));
        if (r->data && r1 - r0 < 64) memcpy(stage, r->data + size_t(r0) * cols, sizeof(float) * size_t(r1 - r0) * cols);
        else {
#pragma omp parallel for schedule(static)
            for (int row = r0; row < r1; row++) {
                const float* src = r->data ? r->data + size_t(row) * cols : r->rows[row];
                memcpy(stage + size_t(row - r0) * cols, src, sizeof(float) * cols);
            }
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(e.d + size_t(r0) * cols, stage, sizeof(float) * size_t(r1 - r0) * cols,
                                      cudaMemcpyHostToDevice, ctx->stream));
        ctx->ring.submitted(slot, ctx->stream);
    }
    ctx->timing.h2d_bytes += (int64_t)(n * sizeof(float));
    e.used = true;
    e.nValid = -1;
    e.rowStart.clear();
    return PB_OK;
}

void freeEntry(pb_context* ctx, RasterEntry& e)
{
    if (!e.borrowed) {
        if (e.d) ctx->pool.release(e.d);
        if (e.validIdx) ctx->pool.release(e.validIdx);
    }
    e = RasterEntry();
}

// valid-cell list of a DEM (cells with !isEqual(z, flag), interpolationCmd.cpp:377), built once per upload
int ensureValidListImpl(pb_context* ctx, RasterEntry& e)
{
    if (e.nValid >= 0) return PB_OK;
    const long long n = (long long)e.n;
    const int nb = compactBlocks(n);
    void* blockCount = nullptr;
    void* idx = nullptr;
    CUDA_TRY(ctx, ctx->pool.acquire(sizeof(int) * size_t(nb), &blockCount));
    CUDA_TRY(ctx, ctx->pool.acquire(sizeof(int) * size_t(n), &idx));
    e.validIdx = static_cast<int*>(idx);
    if (!ctx->dTotal) {
        CUDA_TRY(ctx, cudaMalloc(&ctx->dTotal, sizeof(long long)));
        CUDA_TRY(ctx, cudaMallocHost(&ctx->hTotal, sizeof(long long)));
    }
    CUDA_TRY(ctx, launchCompactValid(e.d, n, e.h.flag, static_cast<int*>(blockCount), ctx->dTotal, nullptr, 0, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hTotal, ctx->dTotal, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, launchCompactValid(e.d, n, e.h.flag, static_cast<int*>(blockCount), ctx->dTotal, e.validIdx, 1, ctx->stream));
    ctx->timing.launches += 3;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->pool.release(blockCount);
    e.nValid = *ctx->hTotal;
    return PB_OK;
}

// per-row offsets into the valid list (row bands for multi-GPU sharding), computed on the device on first use
int ensureRowStartImpl(pb_context* ctx, RasterEntry& e)
{
    if (!e.rowStart.empty()) return PB_OK;
    const int rows = e.h.nrRows;
    void* d = nullptr;
    CUDA_TRY(ctx, ctx->pool.acquire(sizeof(long long) * size_t(rows + 1), &d));
    CUDA_TRY(ctx, launchRowStart(e.validIdx, e.nValid, rows, e.h.nrCols, static_cast<long long*>(d), ctx->stream));
    ctx->timing.launches++;
    e.rowStart.assign(size_t(rows) + 1, 0);
    CUDA_TRY(ctx, cudaMemcpyAsync(e.rowStart.data(), d, sizeof(long long) * size_t(rows + 1), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->pool.release(d);
    return PB_OK;
}

float keyToFloat(unsigned k)
{
    unsigned u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// ---- host side of the local-detrending raster (what Project::interpolationDemLocalDetrending does before its loop) ----

// calculateFirstGuessCombinations (interpolation.cpp:1556-1622): 15 steps along parameter 0 (and 2 for the three-piece
// functions) around the first-guess position; NB the reference indexes firstGuessPosition with the loop counter.
int firstGuessCombinations(const pb_proxy& p, double out[PB_MAX_FIRST_GUESS][PB_MAX_FIT_PARAMS])
{
    const int numSteps = 15;
    const int n = p.nFitParams;
    double stepSize[PB_MAX_FIT_PARAMS], first[PB_MAX_FIT_PARAMS], tmp[PB_MAX_FIT_PARAMS];
    for (int j = 0; j < n; j++) {
        const double mn = p.fitMin[j], mx = p.fitMax[j];
        stepSize[j] = (mx - mn) / numSteps;
        first[j] = p.fitFirstGuess[j] == 0 ? mn : (p.fitFirstGuess[j] == 1 ? (mn + mx) / 2 : mx);
    }
    int count = 0;
    auto push = [&](const double* v) {
        if (count < PB_MAX_FIRST_GUESS) memcpy(out[count], v, sizeof(double) * n);
        count++;
    };
    push(first);
    const int paramIndex[2] = {0, 2};
    const int nIdx = (p.fittingFunction == PB_FIT_PIECEWISE_TWO) ? 1 : 2;
    for (int k = 0; k < nIdx; k++) {
        const int pi = paramIndex[k];
        for (int m = 1; m < numSteps + 1; m++) {
            memcpy(tmp, first, sizeof(double) * n);
            if (p.fitFirstGuess[k] == 0) {
                tmp[pi] = first[pi] + m * stepSize[pi];
                push(tmp);
            } else if (p.fitFirstGuess[k] == 2) {
                tmp[pi] = first[pi] - m * stepSize[pi];
                push(tmp);
            } else if (m < int(numSteps / 2) + 1) {
                tmp[pi] = std::min(first[pi] + m * stepSize[pi], p.fitMax[pi]);
                push(tmp);
                memcpy(tmp, first, sizeof(double) * n);
                tmp[pi] = std::max(first[pi] - m * stepSize[pi], p.fitMin[pi]);
                push(tmp);
            }
        }
    }
    return std::min(count, PB_MAX_FIRST_GUESS);
}

// pb_settings -> LocalModel: selected combination, setMultipleDetrendingHeightTemperatureRange (interpolation.cpp:1506-1553),
// setFittingParameters_elevation / _otherProxies (:1625-1728), localSelection constants (:1092-1099, :1152)
int buildLocalModel(pb_context* ctx, const pb_settings* s, int variable, int nStations, const pb_raster_id* proxyGridIds,
                    int nProxyGrids, LocalModel& m, bool isLocal = true, const double* range = nullptr)
{
    memset(&m, 0, sizeof(m));
    if (s->nProxy < 0 || s->nProxy > PB_MAX_PROXY) return fail(ctx, PB_ERR_INVALID, "settings.nProxy out of range");
    if (isLocal) {
        if (!useDetrendingVar(variable) || !s->useLocalDetrending)      // project.cpp:3155
            return fail(ctx, PB_ERR_INVALID, "local detrending needs useLocalDetrending and a detrending variable");
        if (!s->useMultipleDetrending)
            // (the reference's settings dialog switches multiple detrending on with local detrending,
            // project/dialogInterpolation.cpp:251-268: only a hand-edited .ini reaches this combination; there the classic
            // regression state would leak from cell to cell in OpenMP schedule order, so there is no single reference answer)
            return fail(ctx, PB_ERR_UNSUPPORTED, "local detrending needs multiple detrending (as the reference's settings dialog enforces)");
    }
    if (s->useGlocalDetrending) return fail(ctx, PB_ERR_UNSUPPORTED, "glocal detrending runs through pb_glocal_detrending_fitting / pb_glocal_detrending_raster (macro areas needed)");
    if (!s->hasPointsRange && !range)     // setMultipleDetrendingHeightTemperatureRange returns false (:1508)
        return fail(ctx, PB_ERR_FIT, "couldn't set temperature ranges for height proxy (points range not set)");
    const double rangeMin = range ? range[0] : s->pointsRangeMin, rangeMax = range ? range[1] : s->pointsRangeMax;
    m.nProxy = s->nProxy;
    m.elevationPos = -1;
    for (int i = 0; i < s->nProxy; i++) if (s->proxy[i].kind == PB_PROXY_HEIGHT) m.elevationPos = i;
    m.useLapseRateCode = s->useLapseRateCode;
    m.minPointsLocal = s->minPointsLocalDetrending;
    const float ratioMinPoints = 1.2f;
    m.minPoints = unsigned(s->minPointsLocalDetrending * ratioMinPoints);
    m.thr80 = unsigned(m.minPoints * 0.8);
    m.bboxArea = s->pointsBoundingBoxArea;
    m.shepardRadius = nStations > 0 ? float(sqrt((m.minPoints * s->pointsBoundingBoxArea) / (float(3.141592653589793238462643383) * unsigned(nStations)))) : 0.f;
    m.clampMode = clampModeOf(variable);
    m.rainfallThreshold = s->rainfallThreshold;
    m.useDoNotRetrend = s->useDoNotRetrend;
    m.useRetrendOnly = s->useRetrendOnly;
    m.method = s->method;
    int slots = 0;
    m.elevFunction = PB_FIT_PIECEWISE_TWO;
    for (int i = 0; i < s->nProxy; i++) {
        const pb_proxy& p = s->proxy[i];
        m.selectedActive[i] = s->selectedActive[i];
        m.kind[i] = p.kind;
        m.stdDevThreshold[i] = p.stdDevThreshold;
        m.gridSlot[i] = -1;
        if (p.gridIndex >= 0 && proxyGridIds) {
            if (p.gridIndex >= nProxyGrids) return fail(ctx, PB_ERR_INVALID, "proxy.gridIndex >= nProxyGrids");
            const pb_raster_id id = proxyGridIds[p.gridIndex];
            if (id < 0 || id >= (int)ctx->rasters.size() || !ctx->rasters[id].used)
                return fail(ctx, PB_ERR_INVALID, "proxy grid id is not a live raster");
            m.gridSlot[i] = slots;
            m.grid[slots++] = devGridOf(ctx->rasters[id]);
        }
        if (!s->selectedActive[i]) continue;
        if (i == m.elevationPos) {
            pb_proxy e = p;
            const int f = e.fittingFunction;
            if (!(f == PB_FIT_PIECEWISE_TWO || f == PB_FIT_PIECEWISE_THREE || f == PB_FIT_PIECEWISE_THREE_FREE))
                return fail(ctx, PB_ERR_FIT, "elevation proxy needs a piecewise fitting function (interpolation.cpp:1640-1660)");
            const int want = f == PB_FIT_PIECEWISE_TWO ? 4 : (f == PB_FIT_PIECEWISE_THREE ? 5 : 6);
            if (e.nFitParams != want) return fail(ctx, PB_ERR_FIT, "wrong number of fitting parameters for the elevation proxy");
            e.fitMin[1] = rangeMin - 2;     // MIN_T - 2 / MAX_T + 6 (:1520-1545)
            e.fitMax[1] = rangeMax + 6;
            m.elevFunction = f;
            m.elevNPar = want;
            for (int j = 0; j < want; j++) {
                m.elevMin[j] = e.fitMin[j];
                m.elevMax[j] = e.fitMax[j];
                m.elevDelta[j] = (e.fitMax[j] - e.fitMin[j]) / 800;
            }
            m.nFirstGuess = firstGuessCombinations(e, m.firstGuess);
            if (m.nFirstGuess < 1) return fail(ctx, PB_ERR_FIT, "no first-guess combination for the elevation proxy");
        } else {
            if (p.nFitParams < 2) return fail(ctx, PB_ERR_FIT, "wrong number of parameters for one of the proxies (interpolation.cpp:1700)");
            for (int j = 0; j < 2; j++) {
                m.otherMin[i][j] = p.fitMin[j];
                m.otherMax[i][j] = p.fitMax[j];
                m.otherDelta[i][j] = std::max((p.fitMax[j] - p.fitMin[j]) / 1000, kEpsilon);
                m.otherStart[i][j] = (p.fitMax[j] + p.fitMin[j]) / 2;
            }
        }
    }
    return PB_OK;
}

// all stations in index order -> device (LocalStations)
int stageLocalStations(pb_context* ctx, const pb_stations* st, LocalStations& ds)
{
    const int n = st->n, P = st->nProxy;
    if (n < 0) return fail(ctx, PB_ERR_INVALID, "stations.n < 0");
    if (P > 0 && !st->proxyValues) return fail(ctx, PB_ERR_INVALID, "stations.proxyValues required");
    const size_t nAl = (size_t(n) + 3) / 4 * 4;
    const size_t offXY = 0, offVal = offXY + nAl * 8, offW = offVal + nAl * 4, offCode = offW + nAl * 4,
                 offProxy = offCode + nAl * 4, offAct = offProxy + (size_t(n) * P + 3) / 4 * 16, offXd = (offAct + nAl + 15) / 16 * 16,
                 offYd = offXd + nAl * 8, total = offYd + nAl * 8 + 16;
    CUDA_TRY(ctx, ctx->hLocal.reserve(total));
    CUDA_TRY(ctx, ctx->dLocalStations.reserve(total));
    unsigned char* h = ctx->hLocal.p;
    float2* xy = reinterpret_cast<float2*>(h + offXY);
    float* val = reinterpret_cast<float*>(h + offVal);
    float* w = reinterpret_cast<float*>(h + offW);
    int* code = reinterpret_cast<int*>(h + offCode);
    float* proxy = reinterpret_cast<float*>(h + offProxy);
    unsigned char* act = h + offAct;
    double* xd = reinterpret_cast<double*>(h + offXd);
    double* yd = reinterpret_cast<double*>(h + offYd);
    for (int i = 0; i < n; i++) {
        xd[i] = st->x[i];
        yd[i] = st->y[i];
        xy[i] = make_float2(float(st->x[i]), float(st->y[i]));
        val[i] = st->value[i];
        w[i] = st->regressionWeight ? st->regressionWeight[i] : kNoData;
        code[i] = st->lapseRateCode ? st->lapseRateCode[i] : PB_LAPSE_PRIMARY;
        act[i] = st->isActive ? (st->isActive[i] != 0) : 1;
    }
    if (P > 0) memcpy(proxy, st->proxyValues, sizeof(float) * size_t(n) * P);
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dLocalStations.p, h, total, cudaMemcpyHostToDevice, ctx->stream));
    ctx->timing.h2d_bytes += (int64_t)total;
    unsigned char* d = ctx->dLocalStations.p;
    ds.xy = reinterpret_cast<const float2*>(d + offXY);
    ds.value = reinterpret_cast<const float*>(d + offVal);
    ds.weight = reinterpret_cast<const float*>(d + offW);
    ds.code = reinterpret_cast<const int*>(d + offCode);
    ds.proxy = reinterpret_cast<const float*>(d + offProxy);
    ds.active = d + offAct;
    ds.xd = reinterpret_cast<const double*>(d + offXd);
    ds.yd = reinterpret_cast<const double*>(d + offYd);
    ds.n = n;
    return PB_OK;
}

// radius within which the ring search of localSelection is expected to end, with margin: the kernel keeps the stations
// inside it as a per-cell candidate list (a speed knob only: rings beyond it fall back to full scans)
float candidateRadius(const pb_stations* st, const LocalModel& m)
{
    if (getenv("PB_LOCAL_NO_CANDIDATES")) return 0.f;
    const int n = st->n;
    if (n <= 0 || unsigned(n) <= m.minPoints) return 0.f;
    double xMin = st->x[0], xMax = xMin, yMin = st->y[0], yMax = yMin;
    int counted = 0;
    for (int i = 0; i < n; i++) {
        xMin = std::min(xMin, st->x[i]); xMax = std::max(xMax, st->x[i]);
        yMin = std::min(yMin, st->y[i]); yMax = std::max(yMax, st->y[i]);
        const int code = st->lapseRateCode ? st->lapseRateCode[i] : PB_LAPSE_PRIMARY;
        if (!m.useLapseRateCode || code == PB_LAPSE_PRIMARY || code == PB_LAPSE_SUPPLEMENTAL) counted++;
    }
    const double area = (xMax - xMin) * (yMax - yMin);
    if (!(area > 0) || counted == 0) return 0.f;
    const double rStar = sqrt(double(m.minPoints) * area / (3.141592653589793 * counted));
    return float(std::max(1.5 * rStar, rStar + 7500.0) + 1000.0);
}

// stage + launch K2 on the targets [t0, t0 + nTargets) of the DEM's valid list
int launchLocal(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable, const pb_raster_id* proxyIds,
                int nProxyGrids, RasterEntry& dem, long long t0, int nTargets)
{
    LocalModel m;
    int rc = buildLocalModel(ctx, s, variable, st->n, proxyIds, nProxyGrids, m);
    if (rc) return rc;
    m.candRadius = candidateRadius(st, m);
    LocalStations ds{};
    rc = stageLocalStations(ctx, st, ds);
    if (rc) return rc;
    CUDA_TRY(ctx, ctx->dLocalModel.reserve(sizeof(LocalModel)));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dLocalModel.p, &m, sizeof(m), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));      // `m` lives on this stack frame
    ctx->timing.h2d_bytes += (int64_t)sizeof(m);
    const int groupLanes = localGroupLanes(m.nFirstGuess);
    LocalScratch scr{};
    scr.cap = std::max(std::min(st->n, 2048), 1);
    scr.perGroup = localScratchBytesPerGroup(scr.cap);
    CUDA_TRY(ctx, ctx->dLocalScratch.reserve(scr.perGroup * size_t(localGroupsTotal(groupLanes))));
    scr.base = ctx->dLocalScratch.p;
    if (!ctx->dLocalError) {
        CUDA_TRY(ctx, cudaMalloc(&ctx->dLocalError, sizeof(int)));
        CUDA_TRY(ctx, cudaMallocHost(&ctx->hLocalError, sizeof(int)));
    }
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->dLocalError, 0, sizeof(int), ctx->stream));
    scr.error = ctx->dLocalError;
    cudaEventRecord(ctx->ev[1], ctx->stream);
    const LocalModel* dModel = reinterpret_cast<const LocalModel*>(ctx->dLocalModel.p);
    // Raster-sized target lists go through the three-kernel pipeline (local_pipeline.cu) in chunks of cells; each chunk's
    // cells that do not fit a pipeline record are then gridded by the fused kernel (list form, count read on the device).
    // PB_LOCAL_FUSED=1 forces the fused kernel for everything (A/B comparison of the two schedules: same bits).
    // PB_LOCAL_PIPELINE_MIN: smallest target list that takes the pipeline (tests lower it to drive small rasters through it).
    const char* forceFused = getenv("PB_LOCAL_FUSED");
    const char* minEnv = getenv("PB_LOCAL_PIPELINE_MIN");
    const int minCells = minEnv ? atoi(minEnv) : 4096;
    const bool pipeline = !(forceFused && atoi(forceFused) != 0) && nTargets >= minCells;
    if (!pipeline) {
        CUDA_TRY(ctx, launchLocalDetrendingRaster(dModel, ds, devGridOf(dem), dem.validIdx + t0, nTargets, scr, groupLanes,
                                                  m.elevFunction, ctx->dOut.p, ctx->dMinMax, ctx->stream));
    } else {
        const int chunkEnv = getenv("PB_LOCAL_CHUNK") ? atoi(getenv("PB_LOCAL_CHUNK")) : 0;
        const int chunk = std::min(nTargets, chunkEnv > 0 ? chunkEnv : (512 << 10));
        const size_t recStride = localRecordStride(m.nFirstGuess);
        // Two chunks are in flight, each on its own stream with its own records: a few descents of a chunk run for hundreds of
        // Levenberg-Marquardt iterations (maxIter = 1000) while every other lane has run out of work; the next chunk's kernels
        // take the SMs those stragglers leave free instead of waiting for them.
        const bool twoStreams = nTargets > chunk && !getenv("PB_LOCAL_ONE_STREAM") && !getenv("PB_LOCAL_TIMING");
        const int nSlots = twoStreams ? 2 : 1;
        const size_t workInts = (localPipelineWorkInts(chunk) + 63) / 64 * 64;
        CUDA_TRY(ctx, ctx->dLocalRecords.reserve(recStride * size_t(chunk) * nSlots));
        CUDA_TRY(ctx, ctx->dLocalWork.reserve(workInts * nSlots));
        CUDA_TRY(ctx, ctx->dLocalOverflow.reserve(size_t(nTargets) + 4));
        int* overflowCount = ctx->dLocalOverflow.p;
        int* overflowList = ctx->dLocalOverflow.p + 4;
        CUDA_TRY(ctx, cudaMemsetAsync(overflowCount, 0, 4 * sizeof(int), ctx->stream));
        cudaStream_t streams[2] = {ctx->stream, ctx->stream};
        if (twoStreams) {
            for (int k = 0; k < 2; k++)
                if (!ctx->sPipe[k]) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->sPipe[k], cudaStreamNonBlocking));
            for (auto& e : ctx->evPipe)
                if (!e) CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            CUDA_TRY(ctx, cudaEventRecord(ctx->evPipe[2], ctx->stream));
            for (int k = 0; k < 2; k++) {
                CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->sPipe[k], ctx->evPipe[2], 0));
                streams[k] = ctx->sPipe[k];
            }
        }
        int launches = 0, k = 0;
        for (long long c0 = 0; c0 < nTargets; c0 += chunk, k++) {
            const int nc = int(std::min<long long>(chunk, nTargets - c0));
            const int slot = k % nSlots;
            CUDA_TRY(ctx, launchLocalPipelineChunk(dModel, ds, devGridOf(dem), dem.validIdx + t0 + c0, nc, m.elevFunction, m.candRadius,
                                                   ctx->dLocalRecords.p + recStride * size_t(chunk) * slot, recStride,
                                                   ctx->dLocalWork.p + workInts * slot, overflowList, overflowCount, ctx->dOut.p,
                                                   ctx->dMinMax, streams[slot], &launches));
        }
        if (twoStreams)
            for (int q = 0; q < 2; q++) {
                CUDA_TRY(ctx, cudaEventRecord(ctx->evPipe[q], ctx->sPipe[q]));
                CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evPipe[q], 0));
            }
        // the cells no chunk could take, through the fused kernel (list form, count read on the device)
        CUDA_TRY(ctx, launchLocalDetrendingRaster(dModel, ds, devGridOf(dem), overflowList, nTargets, scr, groupLanes, m.elevFunction,
                                                  ctx->dOut.p, ctx->dMinMax, ctx->stream, nullptr, overflowCount));
        ctx->timing.launches += launches;            // the caller counts one more (the fused kernel's launch)
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hLocalError, ctx->dLocalError, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    return PB_OK;
}

void resetTiming(pb_context* ctx) { memset(&ctx->timing, 0, sizeof(ctx->timing)); }

// ---- preInterpolation on the device (K3), multiple-detrending branch ----
struct PreBatchBuffers {
    std::vector<LocalModel> models;
    std::vector<PreFit> fits;
};

// what preInterpolation leaves in Crit3DInterpolationSettings (interpolation.cpp:1506-1553, 1832-2171) after the
// multiple-detrending fit `f` of one unit (time step or macro area) run with model `m`
void applyPreFit(const pb_settings& s, const LocalModel& m, const PreFit& f, const double* pointsRange, bool singleStep,
                 pb_settings& o)
{
    if (&o != &s) o = s;
    o.hasPointsRange = 1;
    o.pointsRangeMin = m.elevationPos >= 0 && s.selectedActive[m.elevationPos] ? m.elevMin[1] + 2 : o.pointsRangeMin;
    o.pointsRangeMax = m.elevationPos >= 0 && s.selectedActive[m.elevationPos] ? m.elevMax[1] - 6 : o.pointsRangeMax;
    if (pointsRange) { o.pointsRangeMin = pointsRange[0]; o.pointsRangeMax = pointsRange[1]; }
    else if (s.hasPointsRange && singleStep) { o.pointsRangeMin = s.pointsRangeMin; o.pointsRangeMax = s.pointsRangeMax; }
    for (int i = 0; i < o.nProxy; i++) {
        o.currentActive[i] = o.selectedActive[i];
        o.currentSignificant[i] = 0;
    }
    o.currentUseThermalInversion = o.selectedUseThermalInversion;
    o.nFit = o.nFitFunctions = 0;
    const int e = m.elevationPos;
    for (int i = 0; i < o.nProxy; i++)
        if (o.selectedActive[i] && o.proxy[i].kind == PB_PROXY_HEIGHT) {
            pb_proxy& p = o.proxy[i];
            if (i == e) {
                p.fitMin[1] = m.elevMin[1];
                p.fitMax[1] = m.elevMax[1];
                p.nFirstGuessCombinations = m.nFirstGuess;
                memcpy(p.firstGuessCombinations, m.firstGuess, sizeof(m.firstGuess));
            }
        }
    if (e >= 0 && o.selectedActive[e]) {
        o.proxy[e].regressionR2 = f.elevR2;
        if (f.elevR2 != kNoData || f.elevSig) {
            o.fitFunction[o.nFitFunctions++] = m.elevFunction;
            if (f.elevSig) {
                o.fitNrParams[o.nFit] = m.elevNPar;
                for (int j = 0; j < m.elevNPar; j++) o.fitParams[o.nFit][j] = f.elevPar[j];
                o.nFit++;
            }
        }
        o.currentSignificant[e] = f.elevSig ? 1 : 0;
    }
    for (int c = 0; c < f.nPred; c++) {
        o.currentSignificant[f.sigPos[c]] = 1;
        o.fitFunction[o.nFitFunctions++] = PB_FIT_LINEAR;
        o.fitNrParams[o.nFit] = 2;
        o.fitParams[o.nFit][0] = f.othPar[c][0];
        o.fitParams[o.nFit][1] = f.othPar[c][1];
        o.nFit++;
    }
}

int preInterpolationBatchImpl(pb_context* ctx, const pb_stations* st, const float* values, int T, const pb_settings* s,
                              int variable, const double* pointsRange, float* detrended, uint8_t* keep, pb_settings* fitted)
{
    if (!st || !values || !s || !detrended || T < 0) return fail(ctx, PB_ERR_INVALID, "null argument");
    const int n = st->n;
    if (T == 0 || n == 0) return PB_OK;
    if (!useDetrendingVar(variable)) return fail(ctx, PB_ERR_INVALID, "not a detrending variable (interpolation.cpp:1205)");
    if (!s->useMultipleDetrending)
        return fail(ctx, PB_ERR_UNSUPPORTED, "batched preInterpolation covers multiple detrending; classic detrending: pb_pre_interpolation");
    if (s->useLocalDetrending) return fail(ctx, PB_ERR_INVALID, "local detrending fits per cell (pb_local_detrending_raster)");
    if (s->useBestDetrending) return fail(ctx, PB_ERR_UNSUPPORTED, "optimalDetrending is not on the GPU path (DESIGN.md)");
    resetTiming(ctx);
    cudaEventRecord(ctx->ev[0], ctx->stream);
    std::vector<LocalModel> models(T);
    for (int t = 0; t < T; t++) {
        double range[2];
        if (pointsRange) { range[0] = pointsRange[2 * t]; range[1] = pointsRange[2 * t + 1]; }
        else {      // passDataToInterpolation (spatialControl.cpp:533-551): float min / max of the values
            float mn = values[size_t(t) * n], mx = mn;
            for (int i = 1; i < n; i++) {
                const float v = values[size_t(t) * n + i];
                mn = v < mn ? v : mn;
                mx = v > mx ? v : mx;
            }
            range[0] = mn; range[1] = mx;
        }
        int rc = buildLocalModel(ctx, s, variable, n, nullptr, 0, models[t], false, range);
        if (rc) return rc;
    }
    LocalStations ds{};
    int rc = stageLocalStations(ctx, st, ds);
    if (rc) return rc;
    const int groupLanes = models[0].nFirstGuess <= 16 ? 16 : 32;
    LocalScratch scr{};
    scr.cap = n;
    scr.perGroup = preScratchBytesPerGroup(scr.cap);
    const size_t groups = size_t(preGroupsTotal(groupLanes, T));
    const size_t offModels = 0, offFits = offModels + sizeof(LocalModel) * T, offVal = (offFits + sizeof(PreFit) * T + 15) / 16 * 16,
                 offDet = offVal + sizeof(float) * size_t(T) * n, offKeep = offDet + sizeof(float) * size_t(T) * n,
                 offScr = (offKeep + size_t(T) * n + 255) / 256 * 256, total = offScr + scr.perGroup * groups;
    CUDA_TRY(ctx, ctx->dLocalScratch.reserve(total));
    unsigned char* d = ctx->dLocalScratch.p;
    CUDA_TRY(ctx, cudaMemcpyAsync(d + offModels, models.data(), sizeof(LocalModel) * T, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d + offVal, values, sizeof(float) * size_t(T) * n, cudaMemcpyHostToDevice, ctx->stream));
    ctx->timing.h2d_bytes += (int64_t)(sizeof(LocalModel) * T + sizeof(float) * size_t(T) * n);
    scr.base = d + offScr;
    scr.error = nullptr;
    cudaEventRecord(ctx->ev[1], ctx->stream);
    CUDA_TRY(ctx, launchPreInterpolationBatch(reinterpret_cast<const LocalModel*>(d + offModels), ds,
                                              reinterpret_cast<const float*>(d + offVal), T, scr, groupLanes, models[0].elevFunction,
                                              reinterpret_cast<float*>(d + offDet), d + offKeep,
                                              reinterpret_cast<PreFit*>(d + offFits), ctx->stream));
    ctx->timing.launches++;
    cudaEventRecord(ctx->ev[2], ctx->stream);
    std::vector<PreFit> fits(T);
    std::vector<uint8_t> keepTmp;
    if (!keep) { keepTmp.resize(size_t(T) * n); keep = keepTmp.data(); }
    CUDA_TRY(ctx, cudaMemcpyAsync(detrended, d + offDet, sizeof(float) * size_t(T) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(keep, d + offKeep, size_t(T) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(fits.data(), d + offFits, sizeof(PreFit) * T, cudaMemcpyDeviceToHost, ctx->stream));
    cudaEventRecord(ctx->ev[3], ctx->stream);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->timing.d2h_bytes += (int64_t)(size_t(T) * n * 5 + sizeof(PreFit) * T);
    cudaEventElapsedTime(&ctx->timing.h2d_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->timing.kernel_ms, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->timing.d2h_ms, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&ctx->timing.total_ms, ctx->ev[0], ctx->ev[3]);
    if (fitted)
        for (int t = 0; t < T; t++)
            applyPreFit(*s, models[t], fits[t], pointsRange ? pointsRange + 2 * t : nullptr, T == 1, fitted[t]);
    return PB_OK;
}

constexpr int kMaxBands = 16;
constexpr int kMinBandRows = 128;

// page-locked host memory (cudaHostAlloc'd, or registered through pb_host_register) can be the source / target of an
// asynchronous copy itself: no gather into the staging ring, no host memcpy at all
// (both ends of the range are checked: a caller may have registered only part of a grid)
static bool isPinnedHost(const void* p, size_t bytes)
{
    if (!p || bytes == 0) return false;
    const void* ends[2] = {p, static_cast<const unsigned char*>(p) + bytes - 1};
    for (const void* q : ends) {
        cudaPointerAttributes a{};
        if (cudaPointerGetAttributes(&a, q) != cudaSuccess) { cudaGetLastError(); return false; }
        if (a.type != cudaMemoryTypeHost) return false;
    }
    return true;
}

// streams, events and counters of the band pipelines (created on first use)
static int ensurePipeline(pb_context* ctx)
{
    if (!ctx->sUp) {
        CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->sUp, cudaStreamNonBlocking));
        CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->sDn, cudaStreamNonBlocking));
        for (auto& e : ctx->evBand) CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CUDA_TRY(ctx, cudaMalloc(&ctx->dBand, kMaxBands * (sizeof(long long) + sizeof(unsigned))));
        CUDA_TRY(ctx, cudaMallocHost(&ctx->hBand, kMaxBands * sizeof(long long)));
    }
    CUDA_TRY(ctx, ctx->ring.init(kStageSlotBytes));
    CUDA_TRY(ctx, ctx->ringDn.init(kStageSlotBytes));
    return PB_OK;
}

// rows of the device output grid -> the caller's grid on the download stream: straight into page-locked contiguous memory,
// else in chunks through the pinned ring, each chunk scattered into the caller's rows (float** or contiguous) by the host
// cores while the following ones are still in flight
struct BandDownloader {
    pb_context* ctx;
    pb_raster_out* out;
    int cols, chunkRows;
    bool outPinned;
    struct Chunk { int r0, r1, slot; float* stage; };
    std::vector<Chunk> inflight;
    size_t head = 0;
    BandDownloader(pb_context* c, pb_raster_out* o, int rows, int cols_) : ctx(c), out(o), cols(cols_)
    {
        chunkRows = std::max(1, (int)(kStageSlotBytes / (sizeof(float) * cols)));
        outPinned = out->data && isPinnedHost(out->data, sizeof(float) * size_t(rows) * cols);
    }
    cudaError_t scatter(const Chunk& c)
    {
        cudaError_t e = ctx->ringDn.wait(c.slot);
        if (e != cudaSuccess) return e;
        const int n = c.r1 - c.r0;
        if (n < 32) {
            for (int row = c.r0; row < c.r1; row++)
                memcpy(out->data ? out->data + size_t(row) * cols : out->rows[row], c.stage + size_t(row - c.r0) * cols, sizeof(float) * cols);
            return cudaSuccess;
        }
#pragma omp parallel for schedule(static)
        for (int row = c.r0; row < c.r1; row++) {
            float* dst = out->data ? out->data + size_t(row) * cols : out->rows[row];
            memcpy(dst, c.stage + size_t(row - c.r0) * cols, sizeof(float) * cols);
        }
        return cudaSuccess;
    }
    // rows [r0, r1) of dOut start their way down on sDn (the caller has made sDn wait for the kernel that writes them)
    cudaError_t issue(const float* dOut, int r0, int r1, bool mayScatter)
    {
        cudaStream_t sDn = ctx->sDn;
        if (outPinned)
            return cudaMemcpyAsync(out->data + size_t(r0) * cols, dOut + size_t(r0) * cols, size_t(r1 - r0) * cols * sizeof(float),
                                   cudaMemcpyDeviceToHost, sDn);
        for (int a = r0; a < r1; a += chunkRows) {
            if (mayScatter && inflight.size() - head >= size_t(PinnedRing::kSlots - 1)) {
                cudaError_t e = scatter(inflight[head++]);
                if (e != cudaSuccess) return e;
            }
            Chunk c;
            c.r0 = a;
            c.r1 = std::min(r1, a + chunkRows);
            c.stage = reinterpret_cast<float*>(ctx->ringDn.acquire(&c.slot));
            cudaError_t e = cudaMemcpyAsync(c.stage, dOut + size_t(a) * cols, size_t(c.r1 - a) * cols * sizeof(float),
                                            cudaMemcpyDeviceToHost, sDn);
            if (e != cudaSuccess) return e;
            ctx->ringDn.submitted(c.slot, sDn);
            inflight.push_back(c);
        }
        return cudaSuccess;
    }
    cudaError_t finish()
    {
        for (; head < inflight.size(); head++) {
            cudaError_t e = scatter(inflight[head]);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
};

// K1 over RESIDENT rasters with the output grid coming down band by band while the following bands are still being gridded:
// the per-step cost of a cached DEM is max(kernel, download) instead of their sum. Same kernel and arithmetic as one launch
// over the whole valid list (a band is a slice of that list; every cell is independent).
static int residentRasterPipelined(pb_context* ctx, const DevStations& ds, const DevModel& m, RasterEntry& dem, int rowBegin,
                                   int rowEnd, pb_raster_out* out)
{
    int rc = ensurePipeline(ctx);
    if (rc) return rc;
    rc = ensureRowStartImpl(ctx, dem);
    if (rc) return rc;
    const int cols = dem.h.nrCols, nRows = rowEnd - rowBegin;
    // bands: every band is a persistent launch + a download; a 16 MB grid (2000 x 2000) measured 4.4e9 cells/s end to end with
    // 3 bands against 3.9e9 with 8 (4: 4.0e9, 6: 3.8-4.1e9; three runs each): few large bands for small grids, 8 for large ones
    static const int envBands = getenv("PB_BANDS") ? atoi(getenv("PB_BANDS")) : 0;
    const int wantBands = envBands > 0 ? envBands : (size_t(nRows) * cols * sizeof(float) < (size_t(24) << 20) ? 3 : 8);
    const int nBands = std::max(2, std::min(std::min(kMaxBands, wantBands), nRows / kMinBandRows));
    const int bandRows = (nRows + nBands - 1) / nBands;
    cudaStream_t sK = ctx->stream, sDn = ctx->sDn;
    unsigned* dCounters = reinterpret_cast<unsigned*>(reinterpret_cast<long long*>(ctx->dBand) + kMaxBands);
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->dBand, 0, kMaxBands * (sizeof(long long) + sizeof(unsigned)), sK));
    const DevGrid dg = devGridOf(dem);
    BandDownloader dl(ctx, out, dem.h.nrRows, cols);
    auto drain = [&](int code) { cudaStreamSynchronize(sK); cudaStreamSynchronize(sDn); return code; };
    for (int b = 0; b < nBands; b++) {
        const int r0 = rowBegin + b * bandRows, r1 = std::min(rowEnd, r0 + bandRows);
        if (r0 >= r1) break;
        const long long t0 = dem.rowStart[r0], t1 = dem.rowStart[r1];
        if (t1 > t0) {
            cudaError_t ce = launchIdwRaster(ds, m, dg, dem.validIdx + t0, int(t1 - t0), ctx->dOut.p, ctx->dMinMax, sK, nullptr,
                                             dCounters + b);
            if (ce != cudaSuccess) return drain(failCtx(ctx, PB_ERR_CUDA, std::string("band launch: ") + cudaGetErrorString(ce)));
            ctx->timing.launches++;
        }
        cudaEventRecord(ctx->evBand[kMaxBands + b], sK);
        cudaStreamWaitEvent(sDn, ctx->evBand[kMaxBands + b], 0);
        cudaError_t ce = dl.issue(ctx->dOut.p, r0, r1, true);
        if (ce != cudaSuccess) return drain(failCtx(ctx, PB_ERR_CUDA, std::string("D2H: ") + cudaGetErrorString(ce)));
    }
    cudaEventRecord(ctx->ev[2], sK);
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hMinMax, ctx->dMinMax, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, sK));
    cudaError_t ce = dl.finish();
    if (ce != cudaSuccess) return drain(failCtx(ctx, PB_ERR_CUDA, std::string("D2H: ") + cudaGetErrorString(ce)));
    CUDA_TRY(ctx, cudaStreamSynchronize(sDn));
    cudaEventRecord(ctx->ev[3], sK);
    CUDA_TRY(ctx, cudaStreamSynchronize(sK));
    ctx->timing.d2h_bytes += (int64_t)(size_t(nRows) * cols * sizeof(float)) + 8;
    return PB_OK;
}

enum OutMode { OUT_HOST = 0, OUT_DEVICE_ONLY = 1 };

int rasterCore(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable, pb_raster_id demId,
               const pb_raster_id* proxyIds, int nProxyGrids, int rowBegin, int rowEnd, pb_raster_out* out, OutMode mode,
               bool local = false)
{
    if (!ctx) return PB_ERR_INVALID;
    if (!st || !s || !out) return fail(ctx, PB_ERR_INVALID, "null argument");
    if (demId < 0 || demId >= (int)ctx->rasters.size() || !ctx->rasters[demId].used)
        return fail(ctx, PB_ERR_INVALID, "dem id is not a live raster");
    RasterEntry& dem = ctx->rasters[demId];
    if (rowEnd < 0) rowEnd = dem.h.nrRows;
    if (rowBegin < 0 || rowBegin > rowEnd || rowEnd > dem.h.nrRows) return fail(ctx, PB_ERR_INVALID, "bad row band");
    int rc = ensureValidListImpl(ctx, dem);
    if (rc) return rc;

    DevModel m;
    if (!local) {
        rc = buildModel(ctx, s, variable, proxyIds, nProxyGrids, m);
        if (rc) return rc;
    }

    cudaEventRecord(ctx->ev[0], ctx->stream);
    const bool precZero = !local && isPrecVar(variable) && s->precipitationAllZero;
    DevStations ds{};
    ShepardStations ss{};
    float shepardR0 = 0.f;
    const bool shepardMethod = !local && s->method != PB_METHOD_IDW;
    if (!precZero && !local) {
        if (shepardMethod) rc = stageShepardStations(ctx, st, s, true, ss, shepardR0);
        else rc = stageStations(ctx, st, s, true, ds, m.valueOffset);
        if (rc) return rc;
    }
    cudaEventRecord(ctx->ev[1], ctx->stream);

    // output = DEM header, every cell = flag (initializeGrid(raster), gis.cpp:357-363); valid cells overwritten below
    CUDA_TRY(ctx, ctx->dOut.reserve(dem.n));
    if (ctx->outInitRaster != demId) {
        CUDA_TRY(ctx, launchFill(ctx->dOut.p, (long long)dem.n, dem.h.flag, ctx->stream));
        ctx->timing.launches++;
        ctx->outInitRaster = demId;
    }
    ctx->hMinMax[0] = 0xffffffffu;
    ctx->hMinMax[1] = 0u;
    ctx->hMinMax[2] = 0u;           // work-item counter of the persistent IDW kernel
    ctx->hMinMax[3] = 0u;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dMinMax, ctx->hMinMax, 4 * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));

    long long t0 = 0, t1 = dem.nValid;
    if (rowBegin != 0 || rowEnd != dem.h.nrRows) {
        rc = ensureRowStartImpl(ctx, dem);
        if (rc) return rc;
        t0 = dem.rowStart[rowBegin];
        t1 = dem.rowStart[rowEnd];
    }
    const int nTargets = (int)(t1 - t0);
    const DevGrid dg = devGridOf(dem);
    // K1 with a host output grid: band by band, downloads overlapped with the following bands' kernels
    static const bool noBands = getenv("PB_NO_BAND_PIPELINE") != nullptr;
    if (mode == OUT_HOST && !local && !precZero && !shepardMethod && !usesTopographicDistance(s, variable) && !noBands &&
        idwResidentFits(ds.nPadded) && rowEnd - rowBegin >= 2 * kMinBandRows && (out->data || out->rows)) {
        rc = residentRasterPipelined(ctx, ds, m, dem, rowBegin, rowEnd, out);
        if (rc) return rc;
        cudaEventElapsedTime(&ctx->timing.h2d_ms, ctx->ev[0], ctx->ev[1]);
        cudaEventElapsedTime(&ctx->timing.kernel_ms, ctx->ev[1], ctx->ev[2]);
        cudaEventElapsedTime(&ctx->timing.d2h_ms, ctx->ev[2], ctx->ev[3]);
        cudaEventElapsedTime(&ctx->timing.total_ms, ctx->ev[0], ctx->ev[3]);
        out->nValid = nTargets;
        if (ctx->hMinMax[0] == 0xffffffffu) {
            out->minimum = kNoData;
            out->maximum = kNoData;
            return fail(ctx, PB_ERR_NO_VALID_CELL, "no valid cell in the output raster");
        }
        out->minimum = keyToFloat(ctx->hMinMax[0]);
        out->maximum = keyToFloat(ctx->hMinMax[1]);
        return PB_OK;
    }
    if (local) {
        rc = launchLocal(ctx, st, s, variable, proxyIds, nProxyGrids, dem, t0, nTargets);
        if (rc) return rc;
    } else if (precZero) {
        // interpolate() returns exactly 0.f for every cell (interpolation.cpp:2506-2507)
        CUDA_TRY(ctx, launchFillValid(ctx->dOut.p, dem.validIdx + t0, nTargets, 0.f, ctx->stream));
        if (nTargets > 0) { ctx->hMinMax[0] = 0x80000000u; ctx->hMinMax[1] = 0x80000000u; }
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dMinMax, ctx->hMinMax, 2 * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
    } else if (shepardMethod) {
        // computeDistances adds the topographic distance before any estimator (interpolation.cpp:226-251): the march needs the
        // whole resident DEM also for a row band
        ShepardTd std_{};
        std_.kh = usesTopographicDistance(s, variable) ? s->topoDist_Kh : 0;
        std_.dem = dg;
        if (std_.kh != 0) {
            rc = stageTdMapsImpl(ctx, st, s, true, &std_.maps);
            if (rc) return rc;
        }
        CUDA_TRY(ctx, launchShepardRaster(ss, m, dg, shepardR0, s->method == PB_METHOD_SHEPARD_MODIFIED, dem.validIdx + t0,
                                          nTargets, ctx->dOut.p, ctx->dMinMax, ctx->stream, &std_));
    } else if (usesTopographicDistance(s, variable)) {
        // the march needs the whole DEM (a ray reaches any station): `dem` is the full resident raster also for a row band
        const DevGrid* tdMaps = nullptr;
        rc = stageTdMapsImpl(ctx, st, s, true, &tdMaps);
        if (rc) return rc;
        CUDA_TRY(ctx, launchIdwTdRaster(ds, ds.z, m, dg, s->topoDist_Kh, dem.validIdx + t0, nTargets, ctx->dOut.p, ctx->dMinMax,
                                        ctx->stream, tdMaps));
    } else {
        pb_context* par = ctx->parent;
        if (par && ctx->reserveSms > 0) {
            // a lane of a batch driver: raster kernels run one after the other across lanes on all SMs but `reserveSms`
            std::lock_guard<std::mutex> lock(par->rasterMutex);
            if (!ctx->evRaster) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->evRaster, cudaEventDisableTiming));
            if (par->lastRaster && par->lastRaster != ctx->evRaster) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, par->lastRaster, 0));
            CUDA_TRY(ctx, launchIdwRaster(ds, m, dg, dem.validIdx + t0, nTargets, ctx->dOut.p, ctx->dMinMax, ctx->stream, nullptr, nullptr,
                                          ctx->reserveSms));
            CUDA_TRY(ctx, cudaEventRecord(ctx->evRaster, ctx->stream));
            par->lastRaster = ctx->evRaster;
        } else {
            CUDA_TRY(ctx, launchIdwRaster(ds, m, dg, dem.validIdx + t0, nTargets, ctx->dOut.p, ctx->dMinMax, ctx->stream));
        }
    }
    ctx->timing.launches++;
    cudaEventRecord(ctx->ev[2], ctx->stream);

    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hMinMax, ctx->dMinMax, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    const size_t bandOff = size_t(rowBegin) * dem.h.nrCols, bandN = size_t(rowEnd - rowBegin) * dem.h.nrCols;
    if (mode == OUT_HOST) {
        if (!out->data && !out->rows) return fail(ctx, PB_ERR_INVALID, "output has neither data nor rows");
        // D2H in row chunks through the pinned ring; each chunk is scattered into the caller's rows by the host cores
        // while the following ones are still in flight
        const int cols = dem.h.nrCols;
        CUDA_TRY(ctx, ctx->ring.init(kStageSlotBytes));
        const int chunkRows = std::max(1, (int)(kStageSlotBytes / (sizeof(float) * cols)));
        struct Chunk { int r0, r1, slot; float* stage; };
        std::vector<Chunk> inflight;
        size_t head = 0;
        auto scatter = [&](const Chunk& c) -> cudaError_t {
            cudaError_t e = ctx->ring.wait(c.slot);
            if (e != cudaSuccess) return e;
#pragma omp parallel for schedule(static)
            for (int row = c.r0; row < c.r1; row++) {
                float* dst = out->data ? out->data + size_t(row) * cols : out->rows[row];
                memcpy(dst, c.stage + size_t(row - c.r0) * cols, sizeof(float) * cols);
            }
            return cudaSuccess;
        };
        for (int r0 = rowBegin; r0 < rowEnd; r0 += chunkRows) {
            if (inflight.size() - head >= size_t(PinnedRing::kSlots - 1)) {     // the next acquire would reuse this slot
                cudaError_t e = scatter(inflight[head++]);
                if (e != cudaSuccess) return fail(ctx, PB_ERR_CUDA, std::string("D2H: ") + cudaGetErrorString(e));
            }
            Chunk c;
            c.r0 = r0;
            c.r1 = std::min(rowEnd, r0 + chunkRows);
            c.stage = reinterpret_cast<float*>(ctx->ring.acquire(&c.slot));
            CUDA_TRY(ctx, cudaMemcpyAsync(c.stage, ctx->dOut.p + size_t(r0) * cols, size_t(c.r1 - r0) * cols * sizeof(float),
                                          cudaMemcpyDeviceToHost, ctx->stream));
            ctx->ring.submitted(c.slot, ctx->stream);
            inflight.push_back(c);
        }
        cudaEventRecord(ctx->ev[3], ctx->stream);
        for (; head < inflight.size(); head++) {
            cudaError_t e = scatter(inflight[head]);
            if (e != cudaSuccess) return fail(ctx, PB_ERR_CUDA, std::string("D2H: ") + cudaGetErrorString(e));
        }
        ctx->timing.d2h_bytes += (int64_t)(bandN * sizeof(float));
        (void)bandOff;
    } else {
        cudaEventRecord(ctx->ev[3], ctx->stream);
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->timing.d2h_bytes += 8;
    if (local && ctx->hLocalError && *ctx->hLocalError)
        return fail(ctx, PB_ERR_UNSUPPORTED, "a local selection exceeded 2048 stations (dense first ring): outside the supported range");

    cudaEventElapsedTime(&ctx->timing.h2d_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->timing.kernel_ms, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->timing.d2h_ms, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&ctx->timing.total_ms, ctx->ev[0], ctx->ev[3]);

    out->nValid = nTargets;
    if (ctx->hMinMax[0] == 0xffffffffu) {   // no valid output value: updateMinMaxRasterGrid returns false
        out->minimum = kNoData;
        out->maximum = kNoData;
        return fail(ctx, PB_ERR_NO_VALID_CELL, "no valid cell in the output raster");
    }
    out->minimum = keyToFloat(ctx->hMinMax[0]);
    out->maximum = keyToFloat(ctx->hMinMax[1]);
    return PB_OK;
}

}  // namespace

namespace pb {
// used by station_space.cu
int buildDevModel(pb_context* ctx, const pb_settings* s, int variable, DevModel& m) { return buildModel(ctx, s, variable, nullptr, 0, m); }
int preInterpolationMultiple(pb_context* ctx, pb_stations* st, pb_settings* s, int variable, uint8_t* keep)
{
    std::vector<float> in(st->value, st->value + st->n);
    const double range[2] = {s->pointsRangeMin, s->pointsRangeMax};
    pb_settings tmpl = *s;
    return preInterpolationBatchImpl(ctx, st, in.data(), 1, &tmpl, variable, s->hasPointsRange ? range : nullptr, st->value, keep, s);
}
DevGrid devGridOfRaster(const RasterEntry& r) { return devGridOf(r); }
bool varUsesDetrending(int v) { return useDetrendingVar(v); }
bool varUsesTd(int v) { return useTdVar(v); }
bool varIsPrec(int v) { return isPrecVar(v); }
bool varIsThermal(int v)       // isThermal, interpolation.cpp:1190-1203
{
    return v == PB_VAR_AIR_TEMPERATURE || v == PB_VAR_AIR_DEW_TEMPERATURE || v == PB_VAR_DAILY_TAVG || v == PB_VAR_DAILY_TMAX ||
           v == PB_VAR_DAILY_TMIN || v == PB_VAR_DAILY_ET0_HS || v == PB_VAR_ELABORATION;
}
int varClampMode(int v) { return clampModeOf(v); }
// used by glocal.cu
int buildDevModelGrids(pb_context* ctx, const pb_settings* s, int variable, const pb_raster_id* proxyIds, int nProxyGrids, DevModel& m)
{
    return buildModel(ctx, s, variable, proxyIds, nProxyGrids, m);
}
size_t idwStationsStagingBytes(int nStations) { return idwStationsBytes(nStations); }
void packIdwStationsAt(const pb_stations* st, const pb_settings* s, bool excludeSupplemental, unsigned char* h, unsigned char* d,
                       DevStations& ds, double& valueOffset)
{
    packIdwStations(st, s, excludeSupplemental, h, d, ds, valueOffset);
}
int stageShepardStationsFor(pb_context* ctx, const pb_stations* st, const pb_settings* s, bool excludeSupplemental,
                            ShepardStations& ss, float& R0)
{
    return stageShepardStations(ctx, st, s, excludeSupplemental, ss, R0);
}
int buildPreModel(pb_context* ctx, const pb_settings* s, int variable, int nStations, LocalModel& m)
{
    return buildLocalModel(ctx, s, variable, nStations, nullptr, 0, m, false, nullptr);
}
int stageAllStations(pb_context* ctx, const pb_stations* st, LocalStations& ds) { return stageLocalStations(ctx, st, ds); }
void applyPreFitToSettings(const pb_settings& s, const LocalModel& m, const PreFit& f, pb_settings& o)
{
    applyPreFit(s, m, f, nullptr, true, o);
}
int stageTdMapsFor(pb_context* ctx, const int32_t* pointIndex, int n, const DevGrid** out)
{
    return stageTdMapsForImpl(ctx, pointIndex, n, out);
}
int stageTdMaps(pb_context* ctx, const pb_stations* st, const pb_settings* s, bool excludeSupplemental, const DevGrid** out)
{
    return stageTdMapsImpl(ctx, st, s, excludeSupplemental, out);
}
int ensureValidList(pb_context* ctx, RasterEntry& e) { return ensureValidListImpl(ctx, e); }
int ensureRowStart(pb_context* ctx, RasterEntry& e) { return ensureRowStartImpl(ctx, e); }
}  // namespace pb

extern "C" {

const char* pb_version(void) { return "praga_b200 0.1 (sm_100a)"; }

int pb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int pb_create(int device, pb_context** out)
{
    PB_NVTX_RANGE();
    if (!out) return PB_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(nullptr, PB_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                              " (praga_b200 has no CPU fallback)");
    }
    if (device < 0 || device >= n) return fail(nullptr, PB_ERR_INVALID, "device index out of range");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, PB_ERR_CUDA, cudaGetErrorString(e));
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major != 10)
        return fail(nullptr, PB_ERR_CUDA, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                              ", this library is built for sm_100a only");
    pb_context* ctx = new pb_context();
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&ctx->dMinMax, 4 * sizeof(unsigned)) != cudaSuccess ||
        cudaMallocHost(&ctx->hMinMax, 4 * sizeof(unsigned)) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, PB_ERR_CUDA, "context allocation failed");
    }
    for (auto& ev : ctx->ev) cudaEventCreate(&ev);
    *out = ctx;
    return PB_OK;
}

void pb_destroy(pb_context* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (pb_context* lane : ctx->lanes) pb_destroy(lane);
    ctx->lanes.clear();
    for (auto& r : ctx->rasters) if (r.used) freeEntry(ctx, r);
    pb::krigingDestroyAll(ctx->kriging);
    for (auto& g : ctx->aggregations)
        if (g.used) { cudaFree(g.x); cudaFree(g.y); cudaFree(g.first); cudaFree(g.maxNr); cudaFree(g.scratch); cudaFree(g.value); }
    for (auto& m : ctx->macroAreas) if (m.used) cudaFree(m.cells);
    ctx->dLocalStations.release(); ctx->dLocalScratch.release(); ctx->dLocalModel.release(); ctx->hLocal.release();
    ctx->dLocalRecords.release(); ctx->dLocalWork.release(); ctx->dLocalOverflow.release();
    ctx->dMapT.release(); ctx->dAgg.release(); ctx->hUpload.release();
    for (auto& sp : ctx->sPipe) if (sp) cudaStreamDestroy(sp);
    for (auto& e : ctx->evPipe) if (e) cudaEventDestroy(e);
    if (ctx->dLocalError) cudaFree(ctx->dLocalError);
    if (ctx->hLocalError) cudaFreeHost(ctx->hLocalError);
    ctx->dStations.release(); ctx->hStations.release(); ctx->dOut.release(); ctx->dScratch.release(); ctx->ring.release(); ctx->pool.destroy();
    if (ctx->dTotal) cudaFree(ctx->dTotal);
    if (ctx->hTotal) cudaFreeHost(ctx->hTotal);
    ctx->ringDn.release();
    if (ctx->sUp) { cudaStreamDestroy(ctx->sUp); cudaStreamDestroy(ctx->sDn); }
    for (auto& e : ctx->evBand) if (e) cudaEventDestroy(e);
    if (ctx->dBand) cudaFree(ctx->dBand);
    if (ctx->hBand) cudaFreeHost(ctx->hBand);
    if (ctx->dMinMax) cudaFree(ctx->dMinMax);
    if (ctx->hMinMax) cudaFreeHost(ctx->hMinMax);
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->evRaster) cudaEventDestroy(ctx->evRaster);
    if (ctx->ownStream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* pb_last_error(const pb_context* ctx) { return ctx ? ctx->err.c_str() : g_createError.c_str(); }

int pb_set_stream(pb_context* ctx, void* cudaStream)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    cudaStreamSynchronize(ctx->stream);
    if (ctx->ownStream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cudaStream;
    ctx->ownStream = false;
    return PB_OK;
}

int pb_set_blocking_sync(pb_context* ctx, int blocking)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    cudaSetDevice(ctx->device);
    CUDA_TRY(ctx, cudaSetDeviceFlags(blocking ? cudaDeviceScheduleBlockingSync : cudaDeviceScheduleAuto));
    return PB_OK;
}

int pb_last_timing(const pb_context* ctx, pb_timing* t)
{
    if (!ctx || !t) return PB_ERR_INVALID;
    *t = ctx->timing;
    return PB_OK;
}

int pb_raster_upload(pb_context* ctx, const pb_raster* r, pb_raster_id* id)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (!r || !id) return fail(ctx, PB_ERR_INVALID, "null argument");
    cudaSetDevice(ctx->device);
    int slot = -1;
    for (size_t i = 0; i < ctx->rasters.size(); i++) if (!ctx->rasters[i].used) { slot = (int)i; break; }
    if (slot < 0) { ctx->rasters.emplace_back(); slot = (int)ctx->rasters.size() - 1; }
    int rc = uploadRaster(ctx, r, ctx->rasters[slot]);
    if (rc) { freeEntry(ctx, ctx->rasters[slot]); return rc; }
    if (ctx->outInitRaster == slot) ctx->outInitRaster = -1;
    *id = slot;
    return PB_OK;
}

int pb_raster_free(pb_context* ctx, pb_raster_id id)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (id < 0 || id >= (int)ctx->rasters.size() || !ctx->rasters[id].used) return fail(ctx, PB_ERR_INVALID, "bad raster id");
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    freeEntry(ctx, ctx->rasters[id]);
    if (ctx->outInitRaster == id) ctx->outInitRaster = -1;
    for (auto it = ctx->tdMaps.begin(); it != ctx->tdMaps.end();)       // a freed map no longer belongs to its point
        if (it->second == id) it = ctx->tdMaps.erase(it);
        else ++it;
    return PB_OK;
}

/* Crit3DMeteoPoint::topographicDistance of the points (Project::loadTopographicDistanceMaps); see include/praga_b200.h */
int pb_set_topographic_distance_maps(pb_context* ctx, const int32_t* pointIndex, const pb_raster_id* maps, int n)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (n > 0 && (!pointIndex || !maps)) return fail(ctx, PB_ERR_INVALID, "null argument");
    if (n <= 0) { ctx->tdMaps.clear(); return PB_OK; }
    for (int k = 0; k < n; k++)
        if (maps[k] >= 0 && (maps[k] >= (int)ctx->rasters.size() || !ctx->rasters[maps[k]].used))
            return fail(ctx, PB_ERR_INVALID, "topographic-distance map is not a live raster");
    for (int k = 0; k < n; k++) {
        if (maps[k] < 0) ctx->tdMaps.erase(pointIndex[k]);
        else ctx->tdMaps[pointIndex[k]] = maps[k];
    }
    return PB_OK;
}

int pb_raster_from_last_output(pb_context* ctx, pb_raster_id like, pb_raster_id* id)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (!id || like < 0 || like >= (int)ctx->rasters.size() || !ctx->rasters[like].used) return fail(ctx, PB_ERR_INVALID, "bad raster id");
    cudaSetDevice(ctx->device);
    const size_t n = ctx->rasters[like].n;
    if (!ctx->dOut.p || ctx->dOut.cap < n) return fail(ctx, PB_ERR_INVALID, "no output map of that size on the device");
    int slot = -1;
    for (size_t i = 0; i < ctx->rasters.size(); i++) if (!ctx->rasters[i].used) { slot = (int)i; break; }
    if (slot < 0) { ctx->rasters.emplace_back(); slot = (int)ctx->rasters.size() - 1; }
    void* d = nullptr;
    CUDA_TRY(ctx, ctx->pool.acquire(n * sizeof(float), &d));
    RasterEntry& e = ctx->rasters[slot];
    e = RasterEntry();
    e.d = static_cast<float*>(d);
    e.h = ctx->rasters[like].h;
    e.n = n;
    e.used = true;
    if (ctx->outInitRaster == slot) ctx->outInitRaster = -1;
    CUDA_TRY(ctx, cudaMemcpyAsync(e.d, ctx->dOut.p, n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *id = slot;
    return PB_OK;
}

int pb_raster_valid_cells(pb_context* ctx, pb_raster_id id, int64_t* nValid)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (id < 0 || id >= (int)ctx->rasters.size() || !ctx->rasters[id].used) return fail(ctx, PB_ERR_INVALID, "bad raster id");
    cudaSetDevice(ctx->device);
    int rc = ensureValidListImpl(ctx, ctx->rasters[id]);
    if (rc) return rc;
    if (nValid) *nValid = ctx->rasters[id].nValid;
    return PB_OK;
}

int pb_interpolation_raster_resident(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable,
                                     pb_raster_id dem, const pb_raster_id* proxyGrids, int nProxyGrids, int rowBegin,
                                     int rowEnd, pb_raster_out* out)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    cudaSetDevice(ctx->device);
    resetTiming(ctx);
    return rasterCore(ctx, st, s, variable, dem, proxyGrids, nProxyGrids, rowBegin, rowEnd, out, OUT_HOST);
}

int pb_interpolation_raster_device(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable,
                                   pb_raster_id dem, const pb_raster_id* proxyGrids, int nProxyGrids, int rowBegin,
                                   int rowEnd, pb_raster_out* out)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    cudaSetDevice(ctx->device);
    resetTiming(ctx);
    return rasterCore(ctx, st, s, variable, dem, proxyGrids, nProxyGrids, rowBegin, rowEnd, out, OUT_DEVICE_ONLY);
}

/* Local-detrending raster: the loop of Project::interpolationDemLocalDetrending (project/project.cpp:3208-3253). */
int pb_local_detrending_raster_resident(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable,
                                        pb_raster_id dem, const pb_raster_id* proxyGrids, int nProxyGrids, int rowBegin,
                                        int rowEnd, pb_raster_out* out)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    cudaSetDevice(ctx->device);
    resetTiming(ctx);
    return rasterCore(ctx, st, s, variable, dem, proxyGrids, nProxyGrids, rowBegin, rowEnd, out, OUT_HOST, true);
}

int pb_local_detrending_raster_device(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable,
                                      pb_raster_id dem, const pb_raster_id* proxyGrids, int nProxyGrids, int rowBegin,
                                      int rowEnd, pb_raster_out* out)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    cudaSetDevice(ctx->device);
    resetTiming(ctx);
    return rasterCore(ctx, st, s, variable, dem, proxyGrids, nProxyGrids, rowBegin, rowEnd, out, OUT_DEVICE_ONLY, true);
}

/* preInterpolation (interpolation/interpolation.cpp:2444-2499) for T independent time steps sharing one station set */
int pb_pre_interpolation_batch(pb_context* ctx, const pb_stations* st, const float* values, int T, const pb_settings* s,
                               int variable, const double* pointsRange, float* detrended, uint8_t* keep, pb_settings* fitted)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    cudaSetDevice(ctx->device);
    return preInterpolationBatchImpl(ctx, st, values, T, s, variable, pointsRange, detrended, keep, fitted);
}

/* preInterpolation for one time step, in place like the reference: st->value is detrended, *s receives the fit */
int pb_pre_interpolation(pb_context* ctx, pb_stations* st, pb_settings* s, int variable, uint8_t* keep)
{
    PB_NVTX_RANGE();
    // every branch that needs neither the meteo points' extra fields nor the DEM; optimalDetrending's leave-one-out uses
    // the stations themselves as meteo points (pb_pre_interpolation_ex, station_space.cu)
    return pb_pre_interpolation_ex(ctx, st, s, variable, nullptr, -1, keep, nullptr);
}

// ---------------------------------------------------------------------------------------------------------------------
// The drop-in call with HOST rasters as a band pipeline (K1 path): the rows are cut into bands; band b's DEM / proxy rows
// go up on one stream while band b-1 is gridded on the engine's stream and band b-2's output comes down on a third one.
// The valid cells of a band are compacted on the device and the kernel reads their number there, so the host never
// waits inside the loop. Same kernel, same arithmetic as the resident form; only the schedule differs.
// Not eligible (-> sequential form below): local detrending, Shepard, topographic distance (needs the whole DEM),
// all-zero precipitation, more stations than the persistent kernel keeps in shared memory, small rasters.
// ---------------------------------------------------------------------------------------------------------------------
static bool sameGeometry(const pb_raster_header& a, const pb_raster_header& b)
{
    return a.nrRows == b.nrRows && a.nrCols == b.nrCols && a.cellSize == b.cellSize && a.llx == b.llx && a.lly == b.lly;
}

static int allocRasterEntry(pb_context* ctx, const pb_raster* r, pb_raster_id* id)
{
    if (r->h.nrRows <= 0 || r->h.nrCols <= 0) return fail(ctx, PB_ERR_INVALID, "raster has no cells");
    if (!r->data && !r->rows) return fail(ctx, PB_ERR_INVALID, "raster has neither data nor rows");
    const size_t n = size_t(r->h.nrRows) * r->h.nrCols;
    if (n > size_t(0x7fffffff)) return fail(ctx, PB_ERR_UNSUPPORTED, "raster larger than 2^31-1 cells");
    int slot = -1;
    for (size_t i = 0; i < ctx->rasters.size(); i++) if (!ctx->rasters[i].used) { slot = (int)i; break; }
    if (slot < 0) { ctx->rasters.emplace_back(); slot = (int)ctx->rasters.size() - 1; }
    RasterEntry& e = ctx->rasters[slot];
    void* dptr = nullptr;
    CUDA_TRY(ctx, ctx->pool.acquire(n * sizeof(float), &dptr));
    e = RasterEntry();
    e.d = static_cast<float*>(dptr);
    e.h = r->h;
    e.n = n;
    e.used = true;
    if (ctx->outInitRaster == slot) ctx->outInitRaster = -1;
    *id = slot;
    return PB_OK;
}

// rows [r0, r1) of several host rasters (same column count) -> the same rows of their device buffers, through the pinned
// ring, on `stream`. One parallel region gathers the rows of all rasters of a chunk: the regions' fork/join latency, not
// the copy bandwidth, is what limits the host side of the band pipeline.
struct RowsJob { const pb_raster* r; float* dst; };

static int uploadRows(pb_context* ctx, const RowsJob* jobsIn, int nJobsIn, int r0, int r1, cudaStream_t stream)
{
    if (nJobsIn <= 0) return PB_OK;
    const int cols = jobsIn[0].r->h.nrCols;
    // contiguous page-locked rasters go up straight from the caller's memory
    RowsJob jobs[PB_MAX_PROXY + 1];
    int nJobs = 0;
    for (int j = 0; j < nJobsIn; j++) {
        const pb_raster* r = jobsIn[j].r;
        if (r->data && r1 > r0 && isPinnedHost(r->data, sizeof(float) * size_t(r->h.nrRows) * cols)) {
            CUDA_TRY(ctx, cudaMemcpyAsync(jobsIn[j].dst + size_t(r0) * cols, r->data + size_t(r0) * cols,
                                          sizeof(float) * size_t(r1 - r0) * cols, cudaMemcpyHostToDevice, stream));
            ctx->timing.h2d_bytes += (int64_t)(sizeof(float) * size_t(r1 - r0) * cols);
        } else jobs[nJobs++] = jobsIn[j];
    }
    if (nJobs == 0) return PB_OK;
    const int chunkRows = std::max(1, (int)(kStageSlotBytes / (sizeof(float) * cols)));
    // a group of jobs holds one ring slot each until its copies are submitted: never more than half the ring at once
    // (PB_MAX_PROXY + 1 = 9 jobs would otherwise wrap around the 8 slots and share a staging buffer)
    constexpr int kGroup = PinnedRing::kSlots / 2;
    for (int a = r0; a < r1; a += chunkRows) {
        const int b = std::min(r1, a + chunkRows);
        const int nr = b - a;
        for (int j0 = 0; j0 < nJobs; j0 += kGroup) {
            const int nG = std::min(kGroup, nJobs - j0);
            int slot[kGroup];
            float* stage[kGroup];
            for (int j = 0; j < nG; j++) stage[j] = reinterpret_cast<float*>(ctx->ring.acquire(&slot[j]));
#pragma omp parallel for schedule(static)
            for (int k = 0; k < nG * nr; k++) {
                const int j = k / nr, row = a + (k - j * nr);
                const pb_raster* r = jobs[j0 + j].r;
                const float* src = r->data ? r->data + size_t(row) * cols : r->rows[row];
                memcpy(stage[j] + size_t(row - a) * cols, src, sizeof(float) * cols);
            }
            for (int j = 0; j < nG; j++) {
                CUDA_TRY(ctx, cudaMemcpyAsync(jobs[j0 + j].dst + size_t(a) * cols, stage[j], sizeof(float) * size_t(nr) * cols,
                                              cudaMemcpyHostToDevice, stream));
                ctx->ring.submitted(slot[j], stream);
                ctx->timing.h2d_bytes += (int64_t)(sizeof(float) * size_t(nr) * cols);
            }
        }
    }
    return PB_OK;
}

static int hostRasterPipelined(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable, const pb_raster* dem,
                               const pb_raster* proxyGrids, int nProxyGrids, int rowBegin, int rowEnd, pb_raster_out* out,
                               bool* handled)
{
    *handled = false;
    const int rows = dem->h.nrRows, cols = dem->h.nrCols;
    if (rowEnd < 0) rowEnd = rows;
    if (rowBegin < 0 || rowBegin > rowEnd || rowEnd > rows) return PB_OK;          // the sequential form reports it
    const int nRows = rowEnd - rowBegin;
    static const bool off = getenv("PB_NO_BAND_PIPELINE") != nullptr;
    if (off || nRows < 2 * kMinBandRows || s->method != PB_METHOD_IDW || usesTopographicDistance(s, variable) ||
        (isPrecVar(variable) && s->precipitationAllZero) || !idwResidentFits(st->n + tileStations()) ||
        size_t(cols) * sizeof(float) > kStageSlotBytes || (!out->data && !out->rows))
        return PB_OK;
    *handled = true;

    {
        const int rcp = ensurePipeline(ctx);
        if (rcp) return rcp;
    }

    // device buffers (no data yet) registered as rasters so that the model can refer to the proxy grids
    pb_raster_id demId = -1, ids[PB_MAX_PROXY];
    std::vector<pb_raster_id> owned;
    auto cleanup = [&]() {
        std::string keep = ctx->err;
        for (pb_raster_id id : owned) pb_raster_free(ctx, id);
        ctx->err = keep;
    };
    // every error exit after the first allocRasterEntry gives the temporary raster entries back (the three streams may
    // still be reading / writing them: drained first)
    auto abortPipeline = [&](int code) {
        cudaStreamSynchronize(ctx->sUp); cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->sDn);
        cleanup();
        return code;
    };
#define PIPE_TRY(call)                                                                                                  \
    do {                                                                                                                \
        cudaError_t e_ = (call);                                                                                        \
        if (e_ != cudaSuccess)                                                                                          \
            return abortPipeline(pb::failCtx(ctx, PB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)));    \
    } while (0)
    int rc = allocRasterEntry(ctx, dem, &demId);
    if (rc) return rc;
    owned.push_back(demId);
    bool banded[PB_MAX_PROXY];
    for (int g = 0; g < nProxyGrids; g++) {
        const bool same = proxyGrids[g].data == dem->data && proxyGrids[g].rows == dem->rows &&
                          memcmp(&proxyGrids[g].h, &dem->h, sizeof(pb_raster_header)) == 0;
        banded[g] = false;
        if (same) { ids[g] = demId; continue; }
        rc = allocRasterEntry(ctx, &proxyGrids[g], &ids[g]);
        if (rc) { cleanup(); return rc; }
        owned.push_back(ids[g]);
        banded[g] = sameGeometry(proxyGrids[g].h, dem->h);       // a cell looks up itself: band-local rows are enough
    }
    DevModel m;
    rc = buildModel(ctx, s, variable, ids, nProxyGrids, m);
    if (rc) { cleanup(); return rc; }
    cudaStream_t sK = ctx->stream, sUp = ctx->sUp, sDn = ctx->sDn;
    cudaEventRecord(ctx->ev[0], sK);
    DevStations ds{};
    rc = stageStations(ctx, st, s, true, ds, m.valueOffset);
    if (rc) { cleanup(); return rc; }
    if (!idwResidentFits(ds.nPadded)) { cleanup(); *handled = false; return PB_OK; }
    cudaEventRecord(ctx->ev[1], sK);

    RasterEntry& eDem = ctx->rasters[demId];
    const size_t n = eDem.n;
    PIPE_TRY(ctx->dOut.reserve(n));
    ctx->outInitRaster = -1;
    void* vIdx = nullptr;
    PIPE_TRY(ctx->pool.acquire(sizeof(int) * n, &vIdx));
    eDem.validIdx = static_cast<int*>(vIdx);          // released with the entry
    static const int wantBands = getenv("PB_BANDS") ? atoi(getenv("PB_BANDS")) : 8;
    const int nBands = std::max(2, std::min(std::min(kMaxBands, wantBands), nRows / kMinBandRows));
    const int bandRows = (nRows + nBands - 1) / nBands;
    long long* dTotals = reinterpret_cast<long long*>(ctx->dBand);
    unsigned* dCounters = reinterpret_cast<unsigned*>(dTotals + kMaxBands);
    PIPE_TRY(cudaMemsetAsync(ctx->dBand, 0, kMaxBands * (sizeof(long long) + sizeof(unsigned)), sK));
    ctx->hMinMax[0] = 0xffffffffu;
    ctx->hMinMax[1] = 0u;
    ctx->hMinMax[2] = 0u;
    ctx->hMinMax[3] = 0u;
    PIPE_TRY(cudaMemcpyAsync(ctx->dMinMax, ctx->hMinMax, 4 * sizeof(unsigned), cudaMemcpyHostToDevice, sK));
    // the uploads must not overtake work of a previous call that still reads the recycled buffers
    cudaEventRecord(ctx->evBand[2 * kMaxBands], sK);
    cudaStreamWaitEvent(sUp, ctx->evBand[2 * kMaxBands], 0);

    const DevGrid dg = devGridOf(eDem);
    BandDownloader dl(ctx, out, rows, cols);      // the band comes down straight into a page-locked grid, else through the ring
    auto fail2 = abortPipeline;

    // ---- up: a band's rows of the DEM and of the proxy grids that share its geometry (others: whole, with band 0)
    auto issueUpload = [&](int b) -> int {
        const int r0 = rowBegin + b * bandRows, r1 = std::min(rowEnd, r0 + bandRows);
        if (r0 >= r1) return PB_OK;
        RowsJob jobs[PB_MAX_PROXY + 1];
        int nJobs = 0;
        jobs[nJobs++] = RowsJob{dem, eDem.d};
        int rcu = PB_OK;
        for (int g = 0; rcu == PB_OK && g < nProxyGrids; g++) {
            if (ids[g] == demId) continue;
            if (banded[g]) jobs[nJobs++] = RowsJob{&proxyGrids[g], ctx->rasters[ids[g]].d};
            else if (b == 0) {
                const RowsJob whole{&proxyGrids[g], ctx->rasters[ids[g]].d};
                rcu = uploadRows(ctx, &whole, 1, 0, proxyGrids[g].h.nrRows, sUp);
            }
        }
        if (rcu == PB_OK) rcu = uploadRows(ctx, jobs, nJobs, r0, r1, sUp);
        if (rcu) return rcu;
        cudaEventRecord(ctx->evBand[b], sUp);
        return PB_OK;
    };
    // ---- grid + down for one band: one pass writes the flag into the output band and lists the band's valid cells, K1 runs
    //      over that list (its length is read on the device), then the band's output rows start their way down
    auto issueBand = [&](int b, bool mayScatter) -> int {
        const int r0 = rowBegin + b * bandRows, r1 = std::min(rowEnd, r0 + bandRows);
        if (r0 >= r1) return PB_OK;
        const long long bandCells = (long long)(r1 - r0) * cols;
        cudaStreamWaitEvent(sK, ctx->evBand[b], 0);
        cudaError_t ce = launchCompactFill(eDem.d + size_t(r0) * cols, bandCells, eDem.h.flag, eDem.validIdx + size_t(r0) * cols,
                                           int(size_t(r0) * cols), dTotals + b, ctx->dOut.p + size_t(r0) * cols, sK);
        if (ce == cudaSuccess)
            ce = launchIdwRaster(ds, m, dg, eDem.validIdx + size_t(r0) * cols, (int)bandCells, ctx->dOut.p, ctx->dMinMax, sK,
                                 dTotals + b, dCounters + b);
        if (ce != cudaSuccess) return failCtx(ctx, PB_ERR_CUDA, std::string("band pipeline: ") + cudaGetErrorString(ce));
        ctx->timing.launches += 2;
        cudaEventRecord(ctx->evBand[kMaxBands + b], sK);
        cudaStreamWaitEvent(sDn, ctx->evBand[kMaxBands + b], 0);
        ce = dl.issue(ctx->dOut.p, r0, r1, mayScatter);
        if (ce != cudaSuccess) return failCtx(ctx, PB_ERR_CUDA, std::string("D2H: ") + cudaGetErrorString(ce));
        return PB_OK;
    };
    // The upload engine is the busiest resource of the pipeline and must never wait for the host: the next band's uploads are
    // queued BEFORE this band's kernels. (A second host thread issuing the kernels / downloads was measured slower: the two
    // threads contend for the driver and the spinning one disturbs the OpenMP team that gathers the rows.)
    rc = issueUpload(0);
    for (int b = 0; b < nBands && rc == PB_OK; b++) {
        if (b + 1 < nBands) rc = issueUpload(b + 1);
        if (rc == PB_OK) rc = issueBand(b, true);
    }
    if (rc) return fail2(rc);
    cudaEventRecord(ctx->ev[2], sK);
    PIPE_TRY(cudaMemcpyAsync(ctx->hMinMax, ctx->dMinMax, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, sK));
    PIPE_TRY(cudaMemcpyAsync(ctx->hBand, dTotals, kMaxBands * sizeof(long long), cudaMemcpyDeviceToHost, sK));
    if (dl.finish() != cudaSuccess) { failCtx(ctx, PB_ERR_CUDA, "D2H failed"); return fail2(PB_ERR_CUDA); }
    cudaEventRecord(ctx->ev[3], sK);
    cudaStreamSynchronize(sUp);
    cudaStreamSynchronize(sDn);
    PIPE_TRY(cudaStreamSynchronize(sK));
    ctx->timing.d2h_bytes += (int64_t)(size_t(nRows) * cols * sizeof(float)) + 8;
    cudaEventElapsedTime(&ctx->timing.h2d_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->timing.kernel_ms, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->timing.total_ms, ctx->ev[0], ctx->ev[3]);
    long long nValid = 0;
    for (int b = 0; b < nBands; b++) nValid += reinterpret_cast<long long*>(ctx->hBand)[b];
    out->nValid = nValid;
    cleanup();
    if (ctx->hMinMax[0] == 0xffffffffu) {   // no valid output value: updateMinMaxRasterGrid returns false
        out->minimum = kNoData;
        out->maximum = kNoData;
        return fail(ctx, PB_ERR_NO_VALID_CELL, "no valid cell in the output raster");
    }
    out->minimum = keyToFloat(ctx->hMinMax[0]);
    out->maximum = keyToFloat(ctx->hMinMax[1]);
    return PB_OK;
#undef PIPE_TRY
}

static int hostRasterCall(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable, const pb_raster* dem,
                          const pb_raster* proxyGrids, int nProxyGrids, int rowBegin, int rowEnd, pb_raster_out* out, bool local)
{
    if (!ctx) return PB_ERR_INVALID;
    if (!dem || (nProxyGrids > 0 && !proxyGrids)) return fail(ctx, PB_ERR_INVALID, "null raster argument");
    if (nProxyGrids > PB_MAX_PROXY) return fail(ctx, PB_ERR_INVALID, "too many proxy grids");
    cudaSetDevice(ctx->device);
    resetTiming(ctx);
    if (!local && st && s && out) {
        bool handled = false;
        const int rcp = hostRasterPipelined(ctx, st, s, variable, dem, proxyGrids, nProxyGrids, rowBegin, rowEnd, out, &handled);
        if (handled || rcp != PB_OK) return rcp;
        resetTiming(ctx);
    }
    pb_raster_id demId = -1, ids[PB_MAX_PROXY];
    int nUp = 0;
    int rc = pb_raster_upload(ctx, dem, &demId);
    for (int g = 0; rc == PB_OK && g < nProxyGrids; g++) {
        // a proxy grid that IS the DEM (same buffer and header: setProxyDEM, project.cpp:184-211) is uploaded once
        const bool same = proxyGrids[g].data == dem->data && proxyGrids[g].rows == dem->rows &&
                          memcmp(&proxyGrids[g].h, &dem->h, sizeof(pb_raster_header)) == 0;
        if (same) ids[g] = demId;
        else rc = pb_raster_upload(ctx, &proxyGrids[g], &ids[g]);
        if (rc == PB_OK) nUp = g + 1;
    }
    if (rc == PB_OK) rc = rasterCore(ctx, st, s, variable, demId, ids, nProxyGrids, rowBegin, rowEnd, out, OUT_HOST, local);
    std::string keep = ctx->err;
    for (int g = 0; g < nUp; g++) if (ids[g] != demId) pb_raster_free(ctx, ids[g]);
    if (demId >= 0) pb_raster_free(ctx, demId);
    ctx->err = keep;
    return rc;
}

int pb_interpolation_raster(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable,
                            const pb_raster* dem, const pb_raster* proxyGrids, int nProxyGrids, int rowBegin, int rowEnd,
                            pb_raster_out* out)
{
    PB_NVTX_RANGE();
    return hostRasterCall(ctx, st, s, variable, dem, proxyGrids, nProxyGrids, rowBegin, rowEnd, out, false);
}

int pb_local_detrending_raster(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable,
                               const pb_raster* dem, const pb_raster* proxyGrids, int nProxyGrids, int rowBegin, int rowEnd,
                               pb_raster_out* out)
{
    PB_NVTX_RANGE();
    return hostRasterCall(ctx, st, s, variable, dem, proxyGrids, nProxyGrids, rowBegin, rowEnd, out, true);
}

/* interpolate() at explicit targets (interpolation.cpp:2502): the building block of computeResiduals.
 * proxyValues: m * settings.nProxy doubles (the target's own proxy values). */
static int interpolatePointsImpl(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable, const float* x,
                                 const float* y, const float* z, const double* proxyValues, int m, int excludeSupplemental,
                                 pb_raster_id demId, float* result)
{
    if (!ctx) return PB_ERR_INVALID;
    if (!st || !s || !x || !y || !result || m < 0) return fail(ctx, PB_ERR_INVALID, "null argument");
    cudaSetDevice(ctx->device);
    resetTiming(ctx);
    DevModel mod;
    int rc = buildModel(ctx, s, variable, nullptr, 0, mod);
    if (rc) return rc;
    if (isPrecVar(variable) && s->precipitationAllZero) {
        for (int i = 0; i < m; i++) result[i] = 0.f;
        return PB_OK;
    }
    if (mod.nProxy > 0 && !proxyValues) return fail(ctx, PB_ERR_INVALID, "proxyValues required");
    const bool td = usesTopographicDistance(s, variable);
    if (td) {
        if (!z) return fail(ctx, PB_ERR_INVALID, "topographic distance needs the targets' elevation (pb_interpolate_points_td)");
        if (demId < 0 || demId >= (int)ctx->rasters.size() || !ctx->rasters[demId].used)
            return fail(ctx, PB_ERR_INVALID, "topographic distance needs the resident DEM (pb_interpolate_points_td)");
    }
    cudaEventRecord(ctx->ev[0], ctx->stream);
    DevStations ds{};
    ShepardStations ss{};
    float shepardR0 = 0.f;
    const bool shepardMethod = s->method != PB_METHOD_IDW;
    if (shepardMethod && s->method == PB_METHOD_SHEPARD_MODIFIED && s->useLocalDetrending)
        return fail(ctx, PB_ERR_UNSUPPORTED, "modified Shepard with the local radius is not on the GPU path");
    if (shepardMethod) rc = stageShepardStations(ctx, st, s, excludeSupplemental != 0, ss, shepardR0);
    else rc = stageStations(ctx, st, s, excludeSupplemental != 0, ds, mod.valueOffset);
    if (rc) return rc;
    const size_t mAl = (size_t(m) + 3) / 4 * 4;
    const size_t offX = 0, offY = mAl * 4, offP = offY + mAl * 4, offO = offP + size_t(m) * mod.nProxy * 8;
    const size_t offZ = offO + mAl * 4, total = offZ + mAl * 4;
    CUDA_TRY(ctx, ctx->dScratch.reserve(total));
    unsigned char* d = ctx->dScratch.p;
    CUDA_TRY(ctx, cudaMemcpyAsync(d + offX, x, size_t(m) * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d + offY, y, size_t(m) * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (mod.nProxy > 0)
        CUDA_TRY(ctx, cudaMemcpyAsync(d + offP, proxyValues, size_t(m) * mod.nProxy * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (td) CUDA_TRY(ctx, cudaMemcpyAsync(d + offZ, z, size_t(m) * 4, cudaMemcpyHostToDevice, ctx->stream));
    cudaEventRecord(ctx->ev[1], ctx->stream);
    if (shepardMethod) {
        ShepardTd std_{};
        if (td) {
            std_.kh = s->topoDist_Kh;
            std_.dem = devGridOf(ctx->rasters[demId]);
            std_.tz = reinterpret_cast<const float*>(d + offZ);
            rc = stageTdMapsImpl(ctx, st, s, excludeSupplemental != 0, &std_.maps);
            if (rc) return rc;
        }
        CUDA_TRY(ctx, launchShepardPoints(ss, mod, shepardR0, s->method == PB_METHOD_SHEPARD_MODIFIED,
                                          reinterpret_cast<const float*>(d + offX), reinterpret_cast<const float*>(d + offY),
                                          reinterpret_cast<const double*>(d + offP), m, reinterpret_cast<float*>(d + offO),
                                          ctx->stream, &std_));
    }
    else if (td) {
        const DevGrid* tdMaps = nullptr;
        rc = stageTdMapsImpl(ctx, st, s, excludeSupplemental != 0, &tdMaps);
        if (rc) return rc;
        CUDA_TRY(ctx, launchIdwTdPoints(ds, ds.z, mod, devGridOf(ctx->rasters[demId]), s->topoDist_Kh,
                                        reinterpret_cast<const float*>(d + offX), reinterpret_cast<const float*>(d + offY),
                                        reinterpret_cast<const float*>(d + offZ), reinterpret_cast<const double*>(d + offP), m,
                                        reinterpret_cast<float*>(d + offO), ctx->stream, tdMaps));
    }
    else
        CUDA_TRY(ctx, launchIdwPoints(ds, mod, reinterpret_cast<const float*>(d + offX), reinterpret_cast<const float*>(d + offY),
                                      reinterpret_cast<const double*>(d + offP), m, reinterpret_cast<float*>(d + offO), ctx->stream));
    ctx->timing.launches++;
    cudaEventRecord(ctx->ev[2], ctx->stream);
    CUDA_TRY(ctx, cudaMemcpyAsync(result, d + offO, size_t(m) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    cudaEventRecord(ctx->ev[3], ctx->stream);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->timing.h2d_bytes += (int64_t)(size_t(m) * (8 + mod.nProxy * 8));
    ctx->timing.d2h_bytes += (int64_t)(size_t(m) * 4);
    cudaEventElapsedTime(&ctx->timing.h2d_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->timing.kernel_ms, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->timing.d2h_ms, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&ctx->timing.total_ms, ctx->ev[0], ctx->ev[3]);
    return PB_OK;
}


int pb_interpolate_points(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable, const float* x,
                          const float* y, const double* proxyValues, int m, int excludeSupplemental, float* result)
{
    PB_NVTX_RANGE();
    return interpolatePointsImpl(ctx, st, s, variable, x, y, nullptr, proxyValues, m, excludeSupplemental, -1, result);
}

/* the same with topographic distance available: z = the targets' elevation, dem = the resident DEM the march samples */
int pb_interpolate_points_td(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable, const float* x,
                             const float* y, const float* z, const double* proxyValues, int m, int excludeSupplemental,
                             pb_raster_id dem, float* result)
{
    PB_NVTX_RANGE();
    return interpolatePointsImpl(ctx, st, s, variable, x, y, z, proxyValues, m, excludeSupplemental, dem, result);
}

}  // extern "C"

// statistics::linearRegression (mathFunctions/statistics.cpp:1030-1092): f64 sums of float products, float results
static void linearRegressionHost(const std::vector<float>& x, const std::vector<float>& y, float* intercept, float* slope, float* r2)
{
    double SUMx = 0, SUMy = 0, SUMxy = 0, SUMxx = 0, SUM_dy = 0, SUMres = 0;
    long nrValid = 0;
    *slope = 0; *intercept = 0; *r2 = 0;
    for (size_t i = 0; i < x.size(); i++)
        if (x[i] != kNoData && y[i] != kNoData) {
            nrValid++;
            SUMx += x[i];
            SUMy += y[i];
            SUMxy += x[i] * y[i];
            SUMxx += x[i] * x[i];
        }
    const double AVGy = SUMy / nrValid, AVGx = SUMx / nrValid;
    *slope = float((SUMxy - SUMx * SUMy / nrValid) / (SUMxx - SUMx * SUMx / nrValid));
    *intercept = float(AVGy - *slope * AVGx);
    for (size_t i = 0; i < x.size(); i++)
        if (x[i] != kNoData && y[i] != kNoData) {
            double dy = y[i] - (*intercept + *slope * x[i]);
            dy *= dy;
            SUM_dy += dy;
            double res = y[i] - AVGy;
            res *= res;
            SUMres += res;
        }
    *r2 = float((SUMres - SUM_dy) / SUMres);
}

// Project::computeStatisticsCrossValidation (project/project.cpp:2935-2969) over the (observed, predicted) pairs
static void cvStatistics(const std::vector<float>& obs, const std::vector<float>& pre, pb_cv_stats* stats);
namespace pb {
void cvStatisticsHost(const std::vector<float>& obs, const std::vector<float>& pre, pb_cv_stats* stats) { cvStatistics(obs, pre, stats); }
}
static void cvStatistics(const std::vector<float>& obs, const std::vector<float>& pre, pb_cv_stats* stats)
{
    if (!stats) return;
    memset(stats, 0, sizeof(*stats));
    stats->n = int(obs.size());
    stats->meanAbsoluteError = stats->meanBiasError = stats->rootMeanSquareError = stats->nashSutcliffe = stats->R2 = kNoData;
    if (!obs.empty()) {
        double sumAbs = 0, sigma = 0;
        float sumErr = 0.f;
        long cnt = 0;
        for (size_t i = 0; i < obs.size(); i++)
            if (obs[i] != kNoData && pre[i] != kNoData) {
                sumAbs += fabs(obs[i] - pre[i]);
                sumErr += obs[i] - pre[i];
                const double diff = obs[i] - pre[i];
                sigma += diff * diff;
                cnt++;
            }
        if (cnt) {
            stats->meanAbsoluteError = float(sumAbs / cnt);
            stats->meanBiasError = sumErr / cnt;
            stats->rootMeanSquareError = float(sqrt(sigma / cnt));
        }
        double s0 = 0;
        int nv = 0;
        for (float v : obs) if (!isEqualF(v, kNoData)) { s0 += double(v); nv++; }
        if (nv) {
            const float obsAvg = float(s0 / double(nv));
            float sumDev = 0, sumError = 0;
            for (size_t i = 0; i < obs.size(); i++)
                if (!isEqualF(obs[i], kNoData)) sumDev += (obs[i] - obsAvg) * (obs[i] - obsAvg);
            if (sumDev != 0) {
                for (size_t i = 0; i < obs.size(); i++)
                    if (!isEqualF(obs[i], kNoData) && !isEqualF(pre[i], kNoData)) sumError += (pre[i] - obs[i]) * (pre[i] - obs[i]);
                stats->nashSutcliffe = float(1 - (sumError / sumDev));
            }
        }
        float q, m, r2;
        linearRegressionHost(obs, pre, &q, &m, &r2);
        stats->R2 = r2;
    }
}

/* computeResiduals (interpolation/spatialControl.cpp:97-155) + Project::computeStatisticsCrossValidation
 * (project/project.cpp:2935-2969; statistics::meanAbsoluteError / meanError / rootMeanSquareError /
 * NashSutcliffeEfficiency, mathFunctions/statistics.cpp:180-268). The S x S leave-one-out estimates run on the GPU
 * (K1 point kernel: a station's own zero distance weighs nothing); the O(S) statistics are host arithmetic. */
extern "C" int pb_compute_residuals(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable,
                                    const float* observed, const uint8_t* useForCv, float* residual, pb_cv_stats* stats)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (!st || !s || !observed || !residual) return fail(ctx, PB_ERR_INVALID, "null argument");
    const int n = st->n, P = s->nProxy;
    std::vector<float> x(n), y(n), est(n);
    std::vector<double> pv(size_t(n) * std::max(P, 1), double(kNoData));
    for (int i = 0; i < n; i++) {
        x[i] = float(st->x[i]);
        y[i] = float(st->y[i]);
        for (int k = 0; k < P && k < st->nProxy; k++) pv[size_t(i) * P + k] = double(st->proxyValues[size_t(i) * st->nProxy + k]);
    }
    int rc = pb_interpolate_points(ctx, st, s, variable, x.data(), y.data(), pv.data(), n, 0, est.data());
    if (rc) return rc;
    const bool prec = isPrecVar(variable);
    std::vector<float> obs, pre;
    for (int i = 0; i < n; i++) {
        residual[i] = kNoData;
        if (useForCv && !useForCv[i]) continue;
        float myValue = observed[i], e = est[i];
        if (prec) {
            if (myValue != kNoData && myValue < s->rainfallThreshold) myValue = 0.f;
            if (e != kNoData && e < s->rainfallThreshold) e = 0.f;
        }
        if (e != kNoData && myValue != kNoData) residual[i] = myValue - e;
        const float value = observed[i];
        if (!isEqualF(value, kNoData) && !isEqualF(residual[i], kNoData)) {
            obs.push_back(value);
            pre.push_back(value - residual[i]);
        }
    }
    cvStatistics(obs, pre, stats);
    return PB_OK;
}

/* interpolate() through the local-detrending chain at explicit targets: localSelection at the target, preInterpolation on
 * the subset (multipleDetrendingMain), interpolate with the targets' own proxy values. K2 in point form. */
extern "C" int pb_local_detrending_points(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable, const float* x,
                                          const float* y, const double* proxyValues, int m, int excludeSupplemental, float* result)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (!st || !s || (m > 0 && (!x || !y || !result))) return fail(ctx, PB_ERR_INVALID, "null argument");
    if (m <= 0) return PB_OK;
    cudaSetDevice(ctx->device);
    resetTiming(ctx);
    cudaEventRecord(ctx->ev[0], ctx->stream);
    LocalModel lm;
    int rc = buildLocalModel(ctx, s, variable, st->n, nullptr, 0, lm);
    if (rc) return rc;
    lm.candRadius = candidateRadius(st, lm);
    LocalStations ds{};
    rc = stageLocalStations(ctx, st, ds);
    if (rc) return rc;
    const int P = s->nProxy, Pp = std::max(P, 1);
    const size_t offModel = 0, offX = (sizeof(LocalModel) + 255) / 256 * 256, offY = offX + (size_t(m) * 4 + 255) / 256 * 256,
                 offPv = offY + (size_t(m) * 4 + 255) / 256 * 256, offOut = offPv + (size_t(m) * Pp * 8 + 255) / 256 * 256,
                 total = offOut + size_t(m) * 4;
    CUDA_TRY(ctx, ctx->dScratch.reserve(total + 256));
    unsigned char* d = ctx->dScratch.p;
    std::vector<double> none;
    if (!proxyValues) { none.assign(size_t(m) * Pp, double(kNoData)); proxyValues = none.data(); }
    CUDA_TRY(ctx, cudaMemcpyAsync(d + offModel, &lm, sizeof(lm), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d + offX, x, size_t(m) * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d + offY, y, size_t(m) * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (P > 0) CUDA_TRY(ctx, cudaMemcpyAsync(d + offPv, proxyValues, size_t(m) * P * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));      // `lm` and `none` live on this stack frame
    ctx->timing.h2d_bytes += (int64_t)(sizeof(lm) + size_t(m) * (8 + 8 * P));
    const int groupLanes = localGroupLanes(lm.nFirstGuess);
    LocalScratch scr{};
    scr.cap = std::max(std::min(st->n, 2048), 1);
    scr.perGroup = localScratchBytesPerGroup(scr.cap);
    CUDA_TRY(ctx, ctx->dLocalScratch.reserve(scr.perGroup * size_t(localGroupsTotal(groupLanes))));
    scr.base = ctx->dLocalScratch.p;
    if (!ctx->dLocalError) {
        CUDA_TRY(ctx, cudaMalloc(&ctx->dLocalError, sizeof(int)));
        CUDA_TRY(ctx, cudaMallocHost(&ctx->hLocalError, sizeof(int)));
    }
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->dLocalError, 0, sizeof(int), ctx->stream));
    scr.error = ctx->dLocalError;
    LocalPointTargets pt{reinterpret_cast<const float*>(d + offX), reinterpret_cast<const float*>(d + offY),
                         reinterpret_cast<const double*>(d + offPv), excludeSupplemental};
    cudaEventRecord(ctx->ev[1], ctx->stream);
    CUDA_TRY(ctx, launchLocalDetrendingRaster(reinterpret_cast<const LocalModel*>(d + offModel), ds, DevGrid{}, nullptr, m, scr,
                                              groupLanes, lm.elevFunction, reinterpret_cast<float*>(d + offOut), nullptr, ctx->stream,
                                              &pt));
    ctx->timing.launches++;
    cudaEventRecord(ctx->ev[2], ctx->stream);
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hLocalError, ctx->dLocalError, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(result, d + offOut, size_t(m) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    cudaEventRecord(ctx->ev[3], ctx->stream);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->timing.d2h_bytes += (int64_t)(size_t(m) * 4);
    if (*ctx->hLocalError)
        return fail(ctx, PB_ERR_UNSUPPORTED, "a local selection exceeded 2048 stations (dense first ring): outside the supported range");
    cudaEventElapsedTime(&ctx->timing.h2d_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->timing.kernel_ms, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->timing.d2h_ms, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&ctx->timing.total_ms, ctx->ev[0], ctx->ev[3]);
    return PB_OK;
}

/* computeResidualsLocalDetrending (interpolation/spatialControl.cpp:158-228) as Project::interpolationCv calls it
 * (project/project.cpp:3086-3097: excludeOutsideDem = excludeSupplemental = true for the choice of targets; the inner
 * localSelection / interpolate run with excludeSupplemental = false, :186-200) + computeStatisticsCrossValidation. */
extern "C" int pb_compute_residuals_local(pb_context* ctx, const pb_stations* st, const pb_settings* s, int variable,
                                          const float* observed, const uint8_t* useForCv, float* residual, pb_cv_stats* stats)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (!st || !s || !observed || !residual) return fail(ctx, PB_ERR_INVALID, "null argument");
    const int n = st->n, P = s->nProxy;
    std::vector<int> target;
    for (int i = 0; i < n; i++) {
        residual[i] = kNoData;
        if (useForCv && !useForCv[i]) continue;
        const int code = st->lapseRateCode ? st->lapseRateCode[i] : PB_LAPSE_PRIMARY;
        if (s->useLapseRateCode && !(code == PB_LAPSE_PRIMARY || code == PB_LAPSE_SECONDARY)) continue;   // checkLapseRateCode(…, false)
        target.push_back(i);
    }
    const int m = int(target.size());
    std::vector<float> x(m), y(m), est(m);
    std::vector<double> pv(size_t(m) * std::max(P, 1), double(kNoData));
    for (int k = 0; k < m; k++) {
        const int i = target[k];
        x[k] = float(st->x[i]);
        y[k] = float(st->y[i]);
        for (int q = 0; q < P && q < st->nProxy; q++) pv[size_t(k) * P + q] = double(st->proxyValues[size_t(i) * st->nProxy + q]);
    }
    int rc = pb_local_detrending_points(ctx, st, s, variable, x.data(), y.data(), pv.data(), m, 0, est.data());
    if (rc) return rc;
    const bool prec = isPrecVar(variable);
    std::vector<float> obs, pre;
    for (int k = 0; k < m; k++) {
        const int i = target[k];
        float myValue = observed[i], e = est[k];
        if (prec) {
            if (myValue != kNoData && myValue < s->rainfallThreshold) myValue = 0.f;
            if (e != kNoData && e < s->rainfallThreshold) e = 0.f;
        }
        if (e != kNoData && myValue != kNoData) residual[i] = myValue - e;
    }
    for (int i = 0; i < n; i++) {
        if (useForCv && !useForCv[i]) continue;
        const float value = observed[i];
        if (!isEqualF(value, kNoData) && !isEqualF(residual[i], kNoData)) {
            obs.push_back(value);
            pre.push_back(value - residual[i]);
        }
    }
    cvStatistics(obs, pre, stats);
    return PB_OK;
}

/* live pipe ceilings for the K1 roofline (FP32 FMA TFLOP/s, G MUFU.RSQ/s) */
extern "C" int pb_host_register(pb_context* ctx, void* ptr, size_t bytes)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (!ptr || bytes == 0) return fail(ctx, PB_ERR_INVALID, "null argument");
    cudaSetDevice(ctx->device);
    CUDA_TRY(ctx, cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return PB_OK;
}

extern "C" int pb_host_unregister(pb_context* ctx, void* ptr)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (!ptr) return fail(ctx, PB_ERR_INVALID, "null argument");
    cudaSetDevice(ctx->device);
    CUDA_TRY(ctx, cudaHostUnregister(ptr));
    return PB_OK;
}

extern "C" int pb_measure_pipe_peaks(pb_context* ctx, float* fp32Tflops, float* mufuGops)
{
    PB_NVTX_RANGE();
    if (!ctx || !fp32Tflops || !mufuGops) return PB_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, ctx->device);
    CUDA_TRY(ctx, measurePipePeaks(ctx->stream, prop.multiProcessorCount, fp32Tflops, mufuGops));
    return PB_OK;
}

/* ABI self-description: lets bindings check that their struct layouts match this build */
namespace pb { cudaError_t measureFp64Peak(cudaStream_t s, int sms, float* fp64Tflops); }
extern "C" int pb_measure_fp64_peak(pb_context* ctx, float* fp64Tflops)
{
    PB_NVTX_RANGE();
    if (!ctx || !fp64Tflops) return PB_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, ctx->device);
    CUDA_TRY(ctx, measureFp64Peak(ctx->stream, prop.multiProcessorCount, fp64Tflops));
    return PB_OK;
}

extern "C" size_t pb_abi_sizeof(const char* name)
{
    if (!name) return 0;
    if (!strcmp(name, "pb_cv_points")) return sizeof(pb_cv_points);
    if (!strcmp(name, "pb_kh_series")) return sizeof(pb_kh_series);
    if (!strcmp(name, "pb_raster_header")) return sizeof(pb_raster_header);
    if (!strcmp(name, "pb_raster")) return sizeof(pb_raster);
    if (!strcmp(name, "pb_stations")) return sizeof(pb_stations);
    if (!strcmp(name, "pb_proxy")) return sizeof(pb_proxy);
    if (!strcmp(name, "pb_settings")) return sizeof(pb_settings);
    if (!strcmp(name, "pb_raster_out")) return sizeof(pb_raster_out);
    if (!strcmp(name, "pb_cv_stats")) return sizeof(pb_cv_stats);
    if (!strcmp(name, "pb_timing")) return sizeof(pb_timing);
    if (!strcmp(name, "pb_dem_step")) return sizeof(pb_dem_step);
    if (!strcmp(name, "pb_dem_batch_item")) return sizeof(pb_dem_batch_item);
    if (!strcmp(name, "pb_macro_area")) return sizeof(pb_macro_area);
    if (!strcmp(name, "pb_radiation_settings")) return sizeof(pb_radiation_settings);
    if (!strcmp(name, "pb_radiation_maps")) return sizeof(pb_radiation_maps);
    if (!strcmp(name, "pb_radiation_out")) return sizeof(pb_radiation_out);
    if (!strcmp(name, "pb_meteo_grid_var")) return sizeof(pb_meteo_grid_var);
    if (!strcmp(name, "pb_meteo_grid_hour")) return sizeof(pb_meteo_grid_hour);
    return 0;
}
