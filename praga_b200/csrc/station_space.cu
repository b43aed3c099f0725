// station_space.cu — C ABI of the per-time-step station-space pipeline: preInterpolation with every branch
// (agrolib/interpolation/interpolation.cpp:2444-2499): checkPrecipitationZero, multipleDetrendingMain (K3, engine.cu),
// classic detrending(), optimalDetrending() and topographicDistanceOptimize() (K7, classic_kernels.cu).
// The host side only flattens settings, orders the launches and replays the reference's bookkeeping (which fields of
// Crit3DProxy each combination wrote, the golden-section walk over the MAE(kh) table, the Kh series); every sum over
// stations runs on the device.
#include <math.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <map>
#include <tuple>
#include <vector>

#include "classic.cuh"
#include "kh_search.cuh"
#include "context.cuh"
#include "nvtx.cuh"

namespace pb {
// engine.cu
int buildDevModel(pb_context* ctx, const pb_settings* s, int variable, DevModel& m);
int preInterpolationMultiple(pb_context* ctx, pb_stations* st, pb_settings* s, int variable, uint8_t* keep);
DevGrid devGridOfRaster(const RasterEntry& r);
bool varUsesDetrending(int v);
bool varUsesTd(int v);
bool varIsPrec(int v);
bool varIsThermal(int v);
int varClampMode(int v);
void cvStatisticsHost(const std::vector<float>& obs, const std::vector<float>& pre, pb_cv_stats* stats);
int ensureValidList(pb_context* ctx, RasterEntry& e);
int ensureRowStart(pb_context* ctx, RasterEntry& e);
}  // namespace pb

using namespace pb;

namespace {

int fail(pb_context* ctx, int code, const std::string& msg) { return pb::failCtx(ctx, code, msg); }

template <typename T>
T* carve(unsigned char*& p, size_t count)
{
    T* r = reinterpret_cast<T*>(p);
    p += (count * sizeof(T) + 255) / 256 * 256;
    return r;
}

// Many small host arrays -> their places in one device scratch buffer with ONE asynchronous copy: the arrays are gathered into a
// page-locked mirror of the device range they span (gaps travel as they are: scratch the kernels write later). A stage of the
// station-space chain stages 15 arrays; as separate pageable copies they were 15 driver calls, each taking the driver's lock
// that the other lanes of a batch contend for.
struct UploadBatch {
    struct Item { void* dst; const void* src; size_t bytes; };
    std::vector<Item> items;
    cudaError_t add(void* dst, const void* src, size_t bytes)
    {
        if (bytes) items.push_back(Item{dst, src, bytes});
        return cudaSuccess;
    }
    cudaError_t flush(pb_context* ctx, cudaStream_t stream)
    {
        if (items.empty()) return cudaSuccess;
        unsigned char *lo = static_cast<unsigned char*>(items[0].dst), *hi = lo;
        for (const Item& it : items) {
            unsigned char* d = static_cast<unsigned char*>(it.dst);
            lo = std::min(lo, d);
            hi = std::max(hi, d + it.bytes);
        }
        cudaError_t e = ctx->hUpload.reserve(size_t(hi - lo));
        if (e != cudaSuccess) return e;
        for (const Item& it : items) memcpy(ctx->hUpload.p + (static_cast<unsigned char*>(it.dst) - lo), it.src, it.bytes);
        items.clear();
        return cudaMemcpyAsync(lo, ctx->hUpload.p, size_t(hi - lo), cudaMemcpyHostToDevice, stream);
    }
};


// interpolate() with a Shepard estimator (K8, bit-identical point form behind pb_interpolate_points) at the meteo points,
// followed by the residual rules of computeResiduals (spatialControl.cpp:97-155): the leave-one-out of optimalDetrending,
// spatialQualityControl and interpolationCv when TInterpolationMethod is shepard / shepard_modified (the demo's default).
// Sources = the interpolation points (f64 coordinates, detrended values); targets = the meteo points. NOTE: the call goes
// through the context's point-target scratch, so nothing staged there by the caller survives it.
int shepardLoo(pb_context* ctx, const pb_settings* sp, int variable, int nS, const double* sx, const double* sy, const double* sz,
               const float* sval, const int* scode, int nT, const float* tx, const float* ty, const float* tproxy /* nT * max(P,1) */,
               const int* tcode, const uint8_t* tinside, const float* tobs, int excludeOutsideDem, int excludeSupplemental,
               float* residual, float* estimate, const uint8_t* sactive = nullptr, const float* tz = nullptr,
               pb_raster_id demId = -1, const int32_t* sindex = nullptr)
{
    const int P = sp->nProxy;
    const size_t Pp = std::max(P, 1);
    std::vector<float> val(sval, sval + nS), proxy(size_t(nS) * Pp, kNoData), est(nT);
    std::vector<int32_t> code(scode, scode + nS);
    pb_stations st{};
    st.n = nS;
    st.nProxy = P;
    st.x = sx; st.y = sy; st.z = sz;
    st.value = val.data();
    st.lapseRateCode = code.data();
    st.proxyValues = proxy.data();              // interpolate() never reads the points' own proxy values
    st.isActive = sactive;
    st.index = sindex;
    std::vector<double> pv(size_t(nT) * Pp);
    for (size_t k = 0; k < pv.size(); k++) pv[k] = double(tproxy[k]);
    // (tz / demId: the topographic distance of computeDistances, read only when the settings switch it on with kh != 0)
    int rc = pb_interpolate_points_td(ctx, &st, sp, variable, tx, ty, tz, pv.data(), nT, 0, demId, est.data());
    if (rc) return rc;
    const bool isPrec = varIsPrec(variable);
    for (int j = 0; j < nT; j++) {
        bool valid = !excludeSupplemental || !sp->useLapseRateCode || tcode[j] == PB_LAPSE_PRIMARY || tcode[j] == PB_LAPSE_SECONDARY;
        valid = valid && (!excludeOutsideDem || (tinside ? tinside[j] != 0 : true));
        float res = kNoData, e = kNoData;
        if (valid) {
            float myValue = tobs[j];
            e = est[j];
            if (isPrec) {
                if (myValue != kNoData && myValue < sp->rainfallThreshold) myValue = 0.f;
                if (e != kNoData && e < sp->rainfallThreshold) e = 0.f;
            }
            if (e != kNoData && myValue != kNoData) res = myValue - e;
        }
        if (residual) residual[j] = res;
        if (estimate) estimate[j] = e;
    }
    return PB_OK;
}

void applyWrites(pb_proxy& dst, const ClassicProxyState& p)
{
    if (p.written & CF_SLOPE) dst.regressionSlope = p.slope;
    if (p.written & CF_INTERCEPT) dst.regressionIntercept = p.intercept;
    if (p.written & CF_R2) dst.regressionR2 = p.r2;
    if (p.written & CF_H0) dst.lapseRateH0 = p.H0;
    if (p.written & CF_H1) dst.lapseRateH1 = p.H1;
    if (p.written & CF_T0) dst.lapseRateT0 = p.T0;
    if (p.written & CF_T1) dst.lapseRateT1 = p.T1;
    if (p.written & CF_INVLAPSE) dst.inversionLapseRate = p.invLapse;
    if (p.written & CF_INVSIG) dst.inversionIsSignificative = p.invSig;
}

ClassicProxyState stateOf(const pb_proxy& p)
{
    ClassicProxyState s{};
    s.slope = p.regressionSlope; s.intercept = p.regressionIntercept; s.r2 = p.regressionR2;
    s.H0 = p.lapseRateH0; s.H1 = p.lapseRateH1; s.T0 = p.lapseRateT0; s.T1 = p.lapseRateT1;
    s.invLapse = p.inversionLapseRate; s.invSig = p.inversionIsSignificative;
    s.written = 0;
    return s;
}

// Crit3DInterpolationSettings::getCombination (interpolationSettings.cpp:869-892): bit i of the (P+1)-digit binary string,
// most significant first, switches proxy i; the last digit switches the thermal inversion
bool getCombination(const pb_settings& s, int indexHeight, int k, uint8_t* active, int& useInversion)
{
    const int digits = s.nProxy + 1;
    auto bit = [&](int pos) { return (k >> (digits - 1 - pos)) & 1; };
    if (k % 2 == 1)
        if (indexHeight < 0 || !bit(indexHeight)) return false;
    for (int i = 0; i < s.nProxy; i++) active[i] = (bit(i) && s.selectedActive[i]) ? 1 : 0;
    useInversion = (bit(digits - 1) && s.selectedUseThermalInversion) ? 1 : 0;
    return true;
}

}  // namespace

/* preInterpolation with every branch; see include/praga_b200.h */
extern "C" int pb_pre_interpolation_ex(pb_context* ctx, pb_stations* st, pb_settings* s, int variable, const pb_cv_points* cv,
                                       pb_raster_id demId, uint8_t* keep, pb_kh_series* khSeries)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (!st || !s) return fail(ctx, PB_ERR_INVALID, "null argument");
    cudaSetDevice(ctx->device);
    const int n = st->n, P = s->nProxy;
    if (khSeries) khSeries->n = 0;
    if (P < 0 || P > PB_MAX_PROXY || (P > 0 && st->nProxy < P)) return fail(ctx, PB_ERR_INVALID, "settings.nProxy / stations.nProxy mismatch");

    // checkPrecipitationZero (interpolation.cpp:1173-1186): a count, host arithmetic
    if (varIsPrec(variable)) {
        int nrValid = 0;
        for (int i = 0; i < n; i++) {
            const bool active = st->isActive ? st->isActive[i] != 0 : true;
            if (active && !isEqualF(st->value[i], kNoData) && st->value[i] >= s->rainfallThreshold) nrValid++;
        }
        s->precipitationAllZero = nrValid == 0;
        if (keep) memset(keep, 1, n);
        if (s->precipitationAllZero) return PB_OK;
    }

    const std::vector<float> raw(st->value, st->value + n);      // the meteo points' currentValue unless cv gives it
    std::vector<uint8_t> keepTmp;
    const bool detrendVar = varUsesDetrending(variable);
    const bool useTD = s->useTD && !s->useLocalDetrending && varUsesTd(variable) && !s->useGlocalDetrending;
    const bool classic = detrendVar && !s->useMultipleDetrending;
    const bool optimal = classic && s->useBestDetrending;
    if (detrendVar && s->useMultipleDetrending) {
        if (s->useGlocalDetrending) return fail(ctx, PB_ERR_UNSUPPORTED, "glocal detrending runs through pb_glocal_detrending_fitting / pb_glocal_detrending_raster (macro areas needed)");
        if (s->useLocalDetrending) {
            // preInterpolation on the whole set with local detrending fits nothing the raster loop uses: the raster entry
            // point fits per cell (pb_local_detrending_raster)
            return fail(ctx, PB_ERR_INVALID, "local detrending fits per cell (pb_local_detrending_raster)");
        }
        if (!keep) { keepTmp.assign(n, 1); keep = keepTmp.data(); }
        int rc = preInterpolationMultiple(ctx, st, s, variable, keep);
        if (rc) return rc;
    } else if (keep) memset(keep, 1, n);
    if (!classic && !useTD) return PB_OK;
    if ((optimal || useTD) && n == 0) return PB_OK;
    if (useTD && (demId < 0 || demId >= (int)ctx->rasters.size() || !ctx->rasters[demId].used))
        return fail(ctx, PB_ERR_INVALID, "topographic-distance optimisation needs the resident DEM (getCurrentDEM())");
    // Shepard estimators: the leave-one-out of optimalDetrending and of the kh search goes through the K8 point form, one
    // call per (combination, kh) the searches actually visit
    const bool shepardCv = (optimal || useTD) && s->method != PB_METHOD_IDW;

    // interpolation points = survivors of multiple detrending (points with NODATA proxies are erased, :2230); the meteo
    // points the leave-one-out walks stay ALL stations
    std::vector<int> ip;
    for (int i = 0; i < n; i++)
        if (classic || !keep || keep[i]) ip.push_back(i);
    const int m = (int)ip.size();

    // ---- combinations ----
    int indexHeight = -1;
    for (int i = 0; i < P; i++) if (s->proxy[i].kind == PB_PROXY_HEIGHT) indexHeight = i;
    std::vector<ClassicProblem> problems;
    std::vector<int> comboId;
    if (classic) {
        if (optimal) {
            // passDataToInterpolation side effects (spatialControl.cpp:533-564)
            if (n > 0) {
                float xMin = float(st->x[0]), xMax = xMin, yMin = float(st->y[0]), yMax = yMin;
                const float* v0 = (cv && cv->currentValue) ? cv->currentValue : st->value;
                float vMin = v0[0], vMax = v0[0];
                int nrValid = 0;
                for (int i = 0; i < n; i++) {
                    const float x = float(st->x[i]), y = float(st->y[i]);
                    xMin = std::min(xMin, x); xMax = std::max(xMax, x);
                    yMin = std::min(yMin, y); yMax = std::max(yMax, y);
                    vMin = std::min(vMin, v0[i]); vMax = std::max(vMax, v0[i]);
                    const int code = st->lapseRateCode ? st->lapseRateCode[i] : PB_LAPSE_PRIMARY;
                    const bool ok = !s->useLapseRateCode || code == PB_LAPSE_PRIMARY || code == PB_LAPSE_SECONDARY;
                    if (ok) nrValid++;
                }
                if (nrValid > 0) {
                    s->pointsBoundingBoxArea = (xMax - xMin) * (yMax - yMin);
                    s->hasPointsRange = 1;
                    s->pointsRangeMin = vMin;
                    s->pointsRangeMax = vMax;
                }
            }
            const int nrCombination = 1 << (P + 1);
            for (int k = 0; k < nrCombination; k++) {
                ClassicProblem pr{};
                if (!getCombination(*s, indexHeight, k, pr.active, pr.useThermalInversion)) continue;
                problems.push_back(pr);
                comboId.push_back(k);
            }
        } else {
            ClassicProblem pr{};
            for (int i = 0; i < P; i++) pr.active[i] = s->selectedActive[i];
            pr.useThermalInversion = s->selectedUseThermalInversion;
            problems.push_back(pr);
            comboId.push_back(-1);
        }
        for (auto& pr : problems)
            for (int i = 0; i < P; i++) pr.proxy[i] = stateOf(s->proxy[i]);
    } else {
        problems.push_back(ClassicProblem{});      // one problem: the (already detrended) points, model in settings
        comboId.push_back(-1);
    }
    const int NP = (int)problems.size();
    const int maxKh = std::max(0, s->topoDist_maxKh);
    const int nKh = (useTD && !shepardCv) ? maxKh + 1 : 1;

    // ---- stage everything once: sources = interpolation points (m), targets = meteo points (n) ----
    const size_t Pp = std::max(P, 1);
    size_t bytes = 0;
    auto add = [&](size_t count, size_t size) { bytes += (count * size + 255) / 256 * 256; };
    add(m, 4); add(m, 4); add(m, 4); add(m * Pp, 4); add(m, 4); add(m, 8); add(m, 4); add(m, 1);
    add(n, 4); add(n, 4); add(n, 4); add(n * Pp, 4); add(n, 4); add(n, 4); add(n, 1);
    add(size_t(m) * NP, 4);                                  // work
    add(NP, sizeof(ClassicProblem));
    add(size_t(NP) * nKh * n, 4);                            // residuals
    add(size_t(NP) * nKh, 4);                                // mae
    add(nKh, 4);                                             // kh list
    const bool topoTable = useTD && !shepardCv;              // the Shepard form marches inside its own kernel
    if (topoTable) add(size_t(n) * m, 4);                    // target-to-station topographic distances
    CUDA_TRY(ctx, ctx->dScratch.reserve(bytes + 4096));
    unsigned char* d = ctx->dScratch.p;
    float* dX = carve<float>(d, m);
    float* dY = carve<float>(d, m);
    float* dZf = carve<float>(d, m);
    float* dProxy = carve<float>(d, m * Pp);
    float* dValues = carve<float>(d, m);
    double* dZ = carve<double>(d, m);
    int* dCode = carve<int>(d, m);
    uint8_t* dActive = carve<uint8_t>(d, m);
    float* tX = carve<float>(d, n);
    float* tY = carve<float>(d, n);
    float* tZf = carve<float>(d, n);
    float* tProxy = carve<float>(d, n * Pp);
    float* tObserved = carve<float>(d, n);
    int* tCode = carve<int>(d, n);
    uint8_t* tInside = carve<uint8_t>(d, n);
    ClassicProblem* dProblems = carve<ClassicProblem>(d, NP);       // (everything that is uploaded lies before the work areas)
    int* dKh = carve<int>(d, nKh);
    float* dWork = carve<float>(d, size_t(m) * NP);
    float* dResidual = carve<float>(d, size_t(NP) * nKh * n);
    float* dMae = carve<float>(d, size_t(NP) * nKh);
    float* dTopo = topoTable ? carve<float>(d, size_t(n) * m) : nullptr;

    std::vector<float> hx(m), hy(m), hzf(m), hproxy(m * Pp, kNoData), hval(m);
    std::vector<double> hz(m);
    std::vector<int> hcode(m);
    std::vector<uint8_t> hact(m);
    std::vector<float> gx(n), gy(n), gzf(n), gproxy(n * Pp, kNoData), gobs(n);
    std::vector<int> gcode(n);
    std::vector<uint8_t> gin(n);
    for (int i = 0; i < n; i++) {
        gx[i] = float(st->x[i]);
        gy[i] = float(st->y[i]);
        gzf[i] = float(st->z[i]);
        for (int q = 0; q < P; q++) gproxy[size_t(i) * Pp + q] = st->proxyValues[size_t(i) * st->nProxy + q];
        gcode[i] = st->lapseRateCode ? st->lapseRateCode[i] : PB_LAPSE_PRIMARY;
        gin[i] = (cv && cv->isInsideDem) ? cv->isInsideDem[i] : 1;
        gobs[i] = (cv && cv->currentValue) ? cv->currentValue[i] : raw[i];
    }
    for (int k = 0; k < m; k++) {
        const int i = ip[k];
        hx[k] = gx[i];
        hy[k] = gy[i];
        hz[k] = st->z[i];
        hzf[k] = gzf[i];
        for (int q = 0; q < P; q++) hproxy[size_t(k) * Pp + q] = gproxy[size_t(i) * Pp + q];
        hcode[k] = gcode[i];
        // optimalDetrending rebuilds the points from the meteo points: every one active, value = currentValue
        hact[k] = optimal ? 1 : (st->isActive ? st->isActive[i] : 1);
        hval[k] = optimal ? gobs[i] : st->value[i];
    }
    cudaStream_t stream = ctx->stream;
    UploadBatch batch;
    auto up = [&](void* dst, const void* src, size_t b) { return batch.add(dst, src, b); };
    CUDA_TRY(ctx, up(dX, hx.data(), m * 4));
    CUDA_TRY(ctx, up(dY, hy.data(), m * 4));
    CUDA_TRY(ctx, up(dZf, hzf.data(), m * 4));
    CUDA_TRY(ctx, up(dProxy, hproxy.data(), m * Pp * 4));
    CUDA_TRY(ctx, up(dValues, hval.data(), m * 4));
    CUDA_TRY(ctx, up(dZ, hz.data(), m * 8));
    CUDA_TRY(ctx, up(dCode, hcode.data(), m * 4));
    CUDA_TRY(ctx, up(dActive, hact.data(), m));
    CUDA_TRY(ctx, up(tX, gx.data(), n * 4));
    CUDA_TRY(ctx, up(tY, gy.data(), n * 4));
    CUDA_TRY(ctx, up(tZf, gzf.data(), n * 4));
    CUDA_TRY(ctx, up(tProxy, gproxy.data(), n * Pp * 4));
    CUDA_TRY(ctx, up(tObserved, gobs.data(), n * 4));
    CUDA_TRY(ctx, up(tCode, gcode.data(), n * 4));
    CUDA_TRY(ctx, up(tInside, gin.data(), n));
    CUDA_TRY(ctx, up(dProblems, problems.data(), sizeof(ClassicProblem) * NP));
    std::vector<int> khList(nKh);
    for (int k = 0; k < nKh; k++) khList[k] = topoTable ? k : 0;
    CUDA_TRY(ctx, up(dKh, khList.data(), nKh * 4));
    CUDA_TRY(ctx, batch.flush(ctx, stream));
    memset(&ctx->timing, 0, sizeof(ctx->timing));

    // ---- detrending() of every problem ----
    CUDA_TRY(ctx, launchClassicInit(dValues, 1, NP, m, dWork, stream));
    ctx->timing.launches++;
    if (classic) {
        ClassicSetup cs{};
        cs.n = m; cs.nProxy = P; cs.nProblems = NP;
        cs.isThermalVar = varIsThermal(variable);
        cs.isElaborationVar = variable == PB_VAR_ELABORATION;
        cs.useLapseRateCode = s->useLapseRateCode;
        cs.indexHeight = indexHeight;
        cs.minRegressionR2 = s->minRegressionR2;
        cs.maxHeightInversion = s->maxHeightInversion;
        cs.climateLapseRate = s->climateLapseRate;
        for (int i = 0; i < P; i++) cs.kind[i] = s->proxy[i].kind;
        ClassicStations cst{dZ, dProxy, dCode, dActive};
        CUDA_TRY(ctx, launchClassicDetrend(cs, cst, dProblems, dWork, stream));
        ctx->timing.launches++;
        CUDA_TRY(ctx, cudaMemcpyAsync(problems.data(), dProblems, sizeof(ClassicProblem) * NP, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(stream));
        for (auto& pr : problems)
            if (pr.error) return fail(ctx, PB_ERR_UNSUPPORTED, "regressionOrographyT: more than 192 height intervals");
    }

    // ---- leave-one-out tables ----
    std::vector<float> mae;
    std::vector<float> hWork;                 // host copy of the detrended values of every problem (Shepard cross-validation)
    std::vector<int32_t> ipIndex(m);          // Crit3DInterpolationDataPoint::index: the points' precomputed topographic-distance maps
    for (int k = 0; k < m; k++) ipIndex[k] = st->index ? st->index[ip[k]] : ip[k];
    // Shepard estimators: MAE of combination p at kh, computed when a search asks for it (one K8 point-form call over that
    // combination's detrended points and regression fields) and kept
    std::map<std::tuple<int, int, int>, float> shepardMae;
    int shepardRc = PB_OK;
    auto shepardMaeAt = [&](int p, int kh, int excludeOutsideDem) -> float {
        if (shepardRc) return kNoData;
        const auto key = std::make_tuple(p, kh, excludeOutsideDem);
        auto it = shepardMae.find(key);
        if (it != shepardMae.end()) return it->second;
        if (hWork.empty()) {
            hWork.resize(size_t(m) * NP);
            shepardRc = cudaMemcpyAsync(hWork.data(), dWork, hWork.size() * 4, cudaMemcpyDeviceToHost, stream) == cudaSuccess &&
                        cudaStreamSynchronize(stream) == cudaSuccess ? PB_OK : fail(ctx, PB_ERR_CUDA, "download of the detrended points");
            if (shepardRc) return kNoData;
        }
        std::vector<double> sxd(m), syd(m);
        for (int k = 0; k < m; k++) { sxd[k] = st->x[ip[k]]; syd[k] = st->y[ip[k]]; }
        std::vector<float> col(m), res(n);
        pb_settings sp = *s;
        if (classic)
            for (int i = 0; i < P; i++) {
                applyWrites(sp.proxy[i], problems[p].proxy[i]);
                sp.currentActive[i] = problems[p].active[i];
                sp.currentSignificant[i] = problems[p].significant[i];
            }
        if (classic) sp.currentUseThermalInversion = problems[p].useThermalInversion;
        sp.topoDist_Kh = kh;
        for (int k = 0; k < m; k++) col[k] = hWork[size_t(k) * NP + p];
        const pb_timing acc = ctx->timing;
        shepardRc = shepardLoo(ctx, &sp, variable, m, sxd.data(), syd.data(), hz.data(), col.data(), hcode.data(), n, gx.data(),
                               gy.data(), gproxy.data(), gcode.data(), gin.data(), gobs.data(), excludeOutsideDem, 1, res.data(),
                               nullptr, hact.data(), gzf.data(), useTD ? demId : -1, ipIndex.data());
        if (shepardRc) return kNoData;
        ctx->timing.launches += acc.launches;
        ctx->timing.kernel_ms += acc.kernel_ms;
        const float v = cvMaeHost(res.data(), gobs.data(), n);
        shepardMae[key] = v;
        return v;
    };
    auto runLoo = [&](int excludeOutsideDem) -> int {
        LooSetup ls{};
        ls.nS = m; ls.nT = n; ls.nProxy = P; ls.nProblems = NP; ls.nKh = nKh;
        ls.useLapseRateCode = s->useLapseRateCode;
        ls.excludeSupplemental = 1;
        ls.excludeOutsideDem = excludeOutsideDem;
        ls.useDetrendingVar = detrendVar;
        ls.useThermalInversion = s->useThermalInversion;
        ls.useDoNotRetrend = s->useDoNotRetrend;
        ls.useRetrendOnly = s->useRetrendOnly;
        ls.clampMode = varClampMode(variable);
        ls.isPrecVar = varIsPrec(variable);
        ls.rainfallThreshold = s->rainfallThreshold;
        ls.useMultipleModel = 0;
        for (int i = 0; i < P; i++) ls.kind[i] = s->proxy[i].kind;
        if (!classic && detrendVar && s->useMultipleDetrending) {
            int rc = buildDevModel(ctx, s, variable, ls.multiple);
            if (rc) return rc;
            ls.useMultipleModel = 1;
        }
        LooTargets lt{tX, tY, tProxy, tCode, tInside, tObserved};
        CUDA_TRY(ctx, launchLooExact(ls, lt, dX, dY, dWork, dProblems, dTopo, dKh, dResidual, stream));
        CUDA_TRY(ctx, launchCvMae(dResidual, tObserved, n, NP * nKh, dMae, stream));
        ctx->timing.launches += 2;
        mae.resize(size_t(NP) * nKh);
        CUDA_TRY(ctx, cudaMemcpyAsync(mae.data(), dMae, mae.size() * 4, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(stream));
        return PB_OK;
    };

    if (topoTable) {
        const DevGrid dem = devGridOfRaster(ctx->rasters[demId]);
        const DevGrid* tdMaps = nullptr;
        int rcm = stageTdMapsFor(ctx, ipIndex.data(), m, &tdMaps);
        if (rcm) return rcm;
        CUDA_TRY(ctx, launchTopoMatrix(tX, tY, tZf, n, dX, dY, dZf, m, dem, dTopo, stream, tdMaps));
        ctx->timing.launches++;
    }

    int best = 0;
    std::vector<int> khOf(NP, s->topoDist_Kh);
    if (shepardCv) {
        // the same searches as below, the table replaced by shepardMaeAt
        const int excl = cv ? cv->excludeOutsideDem : 0;
        if (useTD)
            for (int p = 0; p < NP; p++) {
                pb_kh_series tmp{};
                khOf[p] = int(goldenSection([&](int kh) { return shepardMaeAt(p, kh, 1); }, maxKh, 0., double(maxKh), &tmp));
            }
        if (optimal) {
            float minError = kNoData;
            int bestCombinationIndex = 0;
            best = -1;
            for (int p = 0; p < NP; p++) {
                const int kh = useTD ? std::max(0, std::min(khOf[p], maxKh)) : 0;
                const float avgError = shepardMaeAt(p, kh, excl);
                if (!isEqualF(avgError, kNoData) && (isEqualF(minError, kNoData) || avgError < minError)) {
                    minError = avgError;
                    bestCombinationIndex = comboId[p];
                    best = p;
                }
            }
            if (best < 0) {
                for (int p = 0; p < NP; p++) if (comboId[p] == bestCombinationIndex) best = p;
                if (best < 0) best = 0;
            }
        }
        if (useTD) {
            if (khSeries) khSeries->n = 0;
            s->topoDist_Kh = int(goldenSection([&](int kh) { return shepardMaeAt(best, kh, 1); }, maxKh, 0., double(maxKh), khSeries));
        }
        if (shepardRc) return shepardRc;
    } else if (optimal || useTD) {
        // topographicDistanceOptimize runs computeResiduals(..., excludeOutsideDem = true, excludeSupplemental = true)
        int rc = runLoo(useTD ? 1 : (cv ? cv->excludeOutsideDem : 0));
        if (rc) return rc;
        std::vector<float> maeTd = mae;
        if (useTD)
            for (int p = 0; p < NP; p++) {
                pb_kh_series tmp{};
                const float* row = &maeTd[size_t(p) * nKh];
                const double bestKh = goldenSection([row](int kh) { return row[kh]; }, maxKh, 0., double(maxKh), &tmp);
                khOf[p] = int(bestKh);
            }
        if (optimal) {
            // computeResiduals(..., getUseExcludeStationsOutsideDEM(), true) at each combination's kh
            const int excl = cv ? cv->excludeOutsideDem : 0;
            const bool sameFilter = !useTD || excl == 1 || !(cv && cv->isInsideDem);
            if (!sameFilter) {
                rc = runLoo(excl);
                if (rc) return rc;
            }
            float minError = kNoData;
            int bestCombinationIndex = 0;       // index into the reference's 0..2^(P+1) numbering
            best = -1;
            for (int p = 0; p < NP; p++) {
                const int kh = useTD ? std::max(0, std::min(khOf[p], maxKh)) : 0;
                const float avgError = mae[size_t(p) * nKh + kh];
                if (!isEqualF(avgError, kNoData) && (isEqualF(minError, kNoData) || avgError < minError)) {
                    minError = avgError;
                    bestCombinationIndex = comboId[p];
                    best = p;
                }
            }
            if (best < 0) {
                // no combination produced an error: the reference falls back to combination 0 (all proxies off)
                for (int p = 0; p < NP; p++) if (comboId[p] == bestCombinationIndex) best = p;
                if (best < 0) best = 0;
            }
        }
        if (useTD) {
            // the search preInterpolation runs last, on the final points (:2493-2496): same table row, series kept
            if (khSeries) khSeries->n = 0;
            const float* row = &maeTd[size_t(best) * nKh];
            const double bestKh = goldenSection([row](int kh) { return row[kh]; }, maxKh, 0., double(maxKh), khSeries);
            s->topoDist_Kh = int(bestKh);
        }
    }

    // ---- what the reference leaves in the settings and the points ----
    if (classic) {
        // optimalDetrending walks the combinations in order on one settings object, then detrends with the best one again
        std::vector<int> order;
        if (optimal) { for (int p = 0; p < NP; p++) order.push_back(p); }
        order.push_back(best);
        for (int p : order)
            for (int i = 0; i < P; i++)
                if (problems[p].active[i]) applyWrites(s->proxy[i], problems[p].proxy[i]);
        const ClassicProblem& fin = problems[best];
        for (int i = 0; i < P; i++) {
            s->currentActive[i] = fin.active[i];
            s->currentSignificant[i] = fin.significant[i];
        }
        s->currentUseThermalInversion = fin.useThermalInversion;
        if (!hWork.empty()) {       // the device staging did not survive the Shepard calls: the host copy has every problem
            for (int k = 0; k < m; k++) st->value[ip[k]] = hWork[size_t(k) * NP + best];
        } else {
            CUDA_TRY(ctx, launchClassicGather(dWork, m, NP, best, dValues, stream));
            ctx->timing.launches++;
            std::vector<float> out(m);
            CUDA_TRY(ctx, cudaMemcpyAsync(out.data(), dValues, m * 4, cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(ctx, cudaStreamSynchronize(stream));
            for (int k = 0; k < m; k++) st->value[ip[k]] = out[k];
        }
    }
    return PB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// exact computeResiduals / interpolate at explicit targets with the FITTED settings, and neighbourhoodVariability:
// the building blocks of spatialQualityControl (spatialControl.cpp:331-427)
// ---------------------------------------------------------------------------------------------------------------------
namespace {

struct LooArrays {
    // targets (meteo points)
    std::vector<float> tx, ty, tz, tproxy, tobs;
    std::vector<int> tcode;
    std::vector<uint8_t> tin;
    // sources (interpolation points, detrended values)
    std::vector<float> sx, sy, szf, sval;
    std::vector<double> sz;
    std::vector<int> scode;
    std::vector<double> sxd, syd;           // the Shepard estimators take the direction cosines from the f64 coordinates
    std::vector<int32_t> sindex;            // Crit3DInterpolationDataPoint::index: the points' precomputed topographic-distance maps
};

// problem = what the settings say about classic detrending (current combination + the proxies' regression fields)
ClassicProblem problemOfSettings(const pb_settings& s)
{
    ClassicProblem pr{};
    for (int i = 0; i < s.nProxy; i++) {
        pr.active[i] = s.currentActive[i];
        pr.significant[i] = s.currentSignificant[i];
        pr.proxy[i] = stateOf(s.proxy[i]);
    }
    pr.useThermalInversion = s.currentUseThermalInversion;
    return pr;
}

int exactLoo(pb_context* ctx, const LooArrays& a, const pb_settings* s, int variable, pb_raster_id demId, int excludeOutsideDem,
             int excludeSupplemental, float* residual, float* estimate, NeighbourOut* nb)
{
    const int nT = (int)a.tx.size(), nS = (int)a.sx.size(), P = s->nProxy;
    if (nT == 0) return PB_OK;
    const size_t Pp = std::max(P, 1);
    const bool detrendVar = varUsesDetrending(variable);
    const bool useTD = s->useTD && !s->useLocalDetrending && varUsesTd(variable) && s->topoDist_Kh != 0;
    if (useTD && (demId < 0 || demId >= (int)ctx->rasters.size() || !ctx->rasters[demId].used))
        return fail(ctx, PB_ERR_INVALID, "topographic distance needs the resident DEM (getCurrentDEM())");
    const bool shepardMethod = s->method != PB_METHOD_IDW;
    if (shepardMethod && (residual || estimate)) {
        if (a.sxd.size() != size_t(nS)) return fail(ctx, PB_ERR_INVALID, "Shepard leave-one-out needs the f64 station coordinates");
        if (nb) {       // the neighbourhood statistics do not depend on the estimator: device pass first, then the estimates
            int rc = exactLoo(ctx, a, s, variable, demId, excludeOutsideDem, excludeSupplemental, nullptr, nullptr, nb);
            if (rc) return rc;
        }
        return shepardLoo(ctx, s, variable, nS, a.sxd.data(), a.syd.data(), a.sz.data(), a.sval.data(), a.scode.data(), nT,
                          a.tx.data(), a.ty.data(), a.tproxy.data(), a.tcode.data(), a.tin.data(), a.tobs.data(), excludeOutsideDem,
                          excludeSupplemental, residual, estimate, nullptr, a.tz.data(), useTD ? demId : -1,
                          a.sindex.size() == size_t(nS) ? a.sindex.data() : nullptr);
    }
    size_t bytes = 0;
    auto add = [&](size_t count, size_t size) { bytes += (count * size + 255) / 256 * 256; };
    add(nT, 4); add(nT, 4); add(nT, 4); add(nT * Pp, 4); add(nT, 4); add(nT, 4); add(nT, 1);
    add(nS, 4); add(nS, 4); add(nS, 4); add(nS, 4); add(nS, 8); add(nS, 4);
    add(1, sizeof(ClassicProblem)); add(nT, 4); add(nT, 4); add(1, 4); add(nT, sizeof(NeighbourOut));
    if (useTD) add(size_t(nT) * nS, 4);
    CUDA_TRY(ctx, ctx->dScratch.reserve(bytes + 4096));
    unsigned char* d = ctx->dScratch.p;
    float* tX = carve<float>(d, nT);
    float* tY = carve<float>(d, nT);
    float* tZ = carve<float>(d, nT);
    float* tProxy = carve<float>(d, nT * Pp);
    float* tObs = carve<float>(d, nT);
    int* tCode = carve<int>(d, nT);
    uint8_t* tIn = carve<uint8_t>(d, nT);
    float* sX = carve<float>(d, nS);
    float* sY = carve<float>(d, nS);
    float* sZf = carve<float>(d, nS);
    float* sVal = carve<float>(d, nS);
    double* sZ = carve<double>(d, nS);
    int* sCode = carve<int>(d, nS);
    ClassicProblem* dProblem = carve<ClassicProblem>(d, 1);
    float* dRes = carve<float>(d, nT);
    float* dEst = carve<float>(d, nT);
    int* dKh = carve<int>(d, 1);
    NeighbourOut* dNb = carve<NeighbourOut>(d, nT);
    float* dTopo = useTD ? carve<float>(d, size_t(nT) * nS) : nullptr;
    cudaStream_t stream = ctx->stream;
    UploadBatch batch;
    auto up = [&](void* dst, const void* src, size_t b) { return batch.add(dst, src, b); };
    CUDA_TRY(ctx, up(tX, a.tx.data(), nT * 4));
    CUDA_TRY(ctx, up(tY, a.ty.data(), nT * 4));
    CUDA_TRY(ctx, up(tZ, a.tz.data(), nT * 4));
    CUDA_TRY(ctx, up(tProxy, a.tproxy.data(), a.tproxy.size() * 4));
    CUDA_TRY(ctx, up(tObs, a.tobs.data(), nT * 4));
    CUDA_TRY(ctx, up(tCode, a.tcode.data(), nT * 4));
    CUDA_TRY(ctx, up(tIn, a.tin.data(), nT));
    CUDA_TRY(ctx, up(sX, a.sx.data(), nS * 4));
    CUDA_TRY(ctx, up(sY, a.sy.data(), nS * 4));
    CUDA_TRY(ctx, up(sZf, a.szf.data(), nS * 4));
    CUDA_TRY(ctx, up(sVal, a.sval.data(), nS * 4));
    CUDA_TRY(ctx, up(sZ, a.sz.data(), nS * 8));
    CUDA_TRY(ctx, up(sCode, a.scode.data(), nS * 4));
    const ClassicProblem pr = problemOfSettings(*s);
    CUDA_TRY(ctx, up(dProblem, &pr, sizeof(pr)));
    const int kh = useTD ? s->topoDist_Kh : 0;
    CUDA_TRY(ctx, up(dKh, &kh, 4));
    CUDA_TRY(ctx, batch.flush(ctx, stream));
    if (useTD) {
        const DevGrid* tdMaps = nullptr;
        if (a.sindex.size() == size_t(nS)) {
            int rcm = stageTdMapsFor(ctx, a.sindex.data(), nS, &tdMaps);
            if (rcm) return rcm;
        }
        CUDA_TRY(ctx, launchTopoMatrix(tX, tY, tZ, nT, sX, sY, sZf, nS, devGridOfRaster(ctx->rasters[demId]), dTopo, stream, tdMaps));
        ctx->timing.launches++;
    }
    if (residual || estimate) {
        LooSetup ls{};
        ls.nS = nS; ls.nT = nT; ls.nProxy = P; ls.nProblems = 1; ls.nKh = 1;
        ls.useLapseRateCode = s->useLapseRateCode;
        ls.excludeSupplemental = excludeSupplemental;
        ls.excludeOutsideDem = excludeOutsideDem;
        ls.useDetrendingVar = detrendVar;
        ls.useThermalInversion = s->useThermalInversion;
        ls.useDoNotRetrend = s->useDoNotRetrend;
        ls.useRetrendOnly = s->useRetrendOnly;
        ls.clampMode = varClampMode(variable);
        ls.isPrecVar = varIsPrec(variable);
        ls.rainfallThreshold = s->rainfallThreshold;
        for (int i = 0; i < P; i++) ls.kind[i] = s->proxy[i].kind;
        if (detrendVar && s->useMultipleDetrending) {
            int rc = buildDevModel(ctx, s, variable, ls.multiple);
            if (rc) return rc;
            ls.useMultipleModel = 1;
        }
        if (ls.isPrecVar && s->precipitationAllZero) {
            // interpolate() returns 0 for every target (interpolation.cpp:2506): retrend-only with nothing to retrend
            ls.useRetrendOnly = 1;
            ls.useDoNotRetrend = 1;
        }
        LooTargets lt{tX, tY, tProxy, tCode, tIn, tObs};
        CUDA_TRY(ctx, launchLooExact(ls, lt, sX, sY, sVal /* NP = 1: work[i] = value i */, dProblem, dTopo, dKh, dRes, stream, dEst));
        ctx->timing.launches++;
        if (residual) CUDA_TRY(ctx, cudaMemcpyAsync(residual, dRes, nT * 4, cudaMemcpyDeviceToHost, stream));
        if (estimate) CUDA_TRY(ctx, cudaMemcpyAsync(estimate, dEst, nT * 4, cudaMemcpyDeviceToHost, stream));
    }
    if (nb) {
        CUDA_TRY(ctx, launchNeighbourhood(tX, tY, tZ, nT, sX, sY, sZ, sCode, sVal, nS, s->useLapseRateCode, dTopo, kh, 10, dNb, stream));
        ctx->timing.launches++;
        CUDA_TRY(ctx, cudaMemcpyAsync(nb, dNb, nT * sizeof(NeighbourOut), cudaMemcpyDeviceToHost, stream));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(stream));
    return PB_OK;
}

// getSpatialThresholdVar, spatialControl.cpp:14-94
float spatialThreshold(int v, float rainfallThreshold, float value, float stdDev, int nrStdDev, float avgDeltaZ, float minDistance)
{
    const float PREC_THRESHOLD = 1000.f;
    float zWeight, distWeight, threshold;
    if (v == PB_VAR_PRECIPITATION || v == PB_VAR_DAILY_PRECIPITATION) {
        distWeight = std::max(1.f, minDistance / 2000.f);
        if (value <= rainfallThreshold) threshold = std::max(5.f, distWeight + stdDev * (nrStdDev + 1));
        else return PREC_THRESHOLD;
    } else if (v == PB_VAR_AIR_TEMPERATURE || v == PB_VAR_AIR_DEW_TEMPERATURE || v == PB_VAR_DAILY_TMAX || v == PB_VAR_DAILY_TMIN ||
               v == PB_VAR_DAILY_TAVG) {
        threshold = 1.f;
        zWeight = avgDeltaZ / 100.f;
        distWeight = minDistance / 1000.f;
        threshold = std::min(std::min(distWeight + threshold + zWeight, 12.f) + stdDev * nrStdDev, 15.f);
    } else if (v == PB_VAR_AIR_REL_HUMIDITY || v == PB_VAR_DAILY_RH_MAX || v == PB_VAR_DAILY_RH_MIN || v == PB_VAR_DAILY_RH_AVG) {
        threshold = 20.f;
        zWeight = avgDeltaZ / 10.f;
        distWeight = minDistance / 1000.f;
        threshold += zWeight + distWeight + stdDev * nrStdDev;
    } else if (v == PB_VAR_WIND_SCALAR_INTENSITY || v == PB_VAR_WIND_VECTOR_INTENSITY || v == PB_VAR_DAILY_WIND_SCALAR_INT_AVG ||
               v == PB_VAR_DAILY_WIND_SCALAR_INT_MAX || v == PB_VAR_DAILY_WIND_VECTOR_INT_AVG || v == PB_VAR_DAILY_WIND_VECTOR_INT_MAX) {
        threshold = 1.f;
        zWeight = avgDeltaZ / 50.f;
        distWeight = minDistance / 2000.f;
        threshold += zWeight + distWeight + stdDev * nrStdDev;
    } else if (v == PB_VAR_GLOBAL_IRRADIANCE) {
        threshold = 500;
        distWeight = minDistance / 5000.f;
        threshold += distWeight + stdDev * (nrStdDev + 1);
    } else if (v == PB_VAR_DAILY_GLOBAL_RADIATION) {
        threshold = 10;
        distWeight = minDistance / 5000.f;
        threshold += distWeight + stdDev * (nrStdDev + 1);
    } else if (v == PB_VAR_ATM_TRANSMISSIVITY) {
        distWeight = minDistance / 20000.f;
        threshold = std::min(std::max(distWeight + stdDev * nrStdDev, 0.3f), 0.8f);
    } else if (v == PB_VAR_ATM_PRESSURE) {
        zWeight = avgDeltaZ / 10.f;
        threshold = zWeight + stdDev * nrStdDev;
    } else if (v == PB_VAR_DAILY_BIC) threshold = PREC_THRESHOLD;
    else threshold = stdDev * nrStdDev;
    return threshold;
}

void fillTargets(LooArrays& a, const pb_stations* st, const std::vector<int>& tIdx, const std::vector<int>& proxyFrom,
                 const float* observed, const uint8_t* insideDem, int P)
{
    const size_t nT = tIdx.size(), Pp = std::max(P, 1);
    a.tx.resize(nT); a.ty.resize(nT); a.tz.resize(nT); a.tobs.resize(nT); a.tcode.resize(nT); a.tin.resize(nT);
    a.tproxy.assign(nT * Pp, kNoData);
    for (size_t k = 0; k < nT; k++) {
        const int i = tIdx[k], pi = proxyFrom[k];
        a.tx[k] = float(st->x[i]);
        a.ty[k] = float(st->y[i]);
        a.tz[k] = float(st->z[i]);
        a.tobs[k] = observed[i];
        a.tcode[k] = st->lapseRateCode ? st->lapseRateCode[i] : PB_LAPSE_PRIMARY;
        a.tin[k] = insideDem ? insideDem[i] : 1;
        for (int q = 0; q < P; q++) a.tproxy[k * Pp + q] = st->proxyValues[size_t(pi) * st->nProxy + q];
    }
}

void fillSources(LooArrays& a, const pb_stations* pts, const uint8_t* keep)
{
    a.sx.clear(); a.sy.clear(); a.szf.clear(); a.sval.clear(); a.sz.clear(); a.scode.clear(); a.sxd.clear(); a.syd.clear();
    a.sindex.clear();
    for (int i = 0; i < pts->n; i++) {
        if (keep && !keep[i]) continue;
        a.sindex.push_back(pts->index ? pts->index[i] : i);
        a.sxd.push_back(pts->x[i]);
        a.syd.push_back(pts->y[i]);
        a.sx.push_back(float(pts->x[i]));
        a.sy.push_back(float(pts->y[i]));
        a.szf.push_back(float(pts->z[i]));
        a.sz.push_back(pts->z[i]);
        a.sval.push_back(pts->value[i]);
        a.scode.push_back(pts->lapseRateCode ? pts->lapseRateCode[i] : PB_LAPSE_PRIMARY);
    }
}

// a pb_stations view over the entries `idx` of st (host copies owned by the struct)
struct SubStations {
    std::vector<double> x, y, z;
    std::vector<float> value, weight, proxy;
    std::vector<int32_t> code, index;
    pb_stations st{};
    SubStations(const pb_stations* src, const std::vector<int>& idx, const float* values)
    {
        const size_t n = idx.size(), P = size_t(std::max(src->nProxy, 0));
        x.resize(n); y.resize(n); z.resize(n); value.resize(n); code.resize(n); index.resize(n);
        proxy.resize(n * std::max<size_t>(P, 1));
        for (size_t k = 0; k < n; k++) {
            const int i = idx[k];
            x[k] = src->x[i]; y[k] = src->y[i]; z[k] = src->z[i];
            value[k] = values[i];
            code[k] = src->lapseRateCode ? src->lapseRateCode[i] : PB_LAPSE_PRIMARY;
            index[k] = src->index ? src->index[i] : i;      // the view keeps the points' identity
            for (size_t q = 0; q < P; q++) proxy[k * P + q] = src->proxyValues[size_t(i) * P + q];
        }
        st.n = int(n);
        st.nProxy = src->nProxy;
        st.x = x.data(); st.y = y.data(); st.z = z.data();
        st.value = value.data();
        st.lapseRateCode = code.data();
        st.proxyValues = proxy.data();
        st.index = index.data();
    }
};

// passDataToInterpolation's writes into the settings (spatialControl.cpp:560-565)
void passDataSideEffects(const pb_stations* st, pb_settings* s)
{
    const int n = st->n;
    if (n == 0) return;
    float xMin = float(st->x[0]), xMax = xMin, yMin = float(st->y[0]), yMax = yMin, vMin = st->value[0], vMax = vMin;
    for (int i = 1; i < n; i++) {
        const float x = float(st->x[i]), y = float(st->y[i]);
        xMin = std::min(xMin, x); xMax = std::max(xMax, x);
        yMin = std::min(yMin, y); yMax = std::max(yMax, y);
        vMin = std::min(vMin, st->value[i]); vMax = std::max(vMax, st->value[i]);
    }
    s->pointsBoundingBoxArea = (xMax - xMin) * (yMax - yMin);
    s->hasPointsRange = 1;
    s->pointsRangeMin = vMin;
    s->pointsRangeMax = vMax;
}

}  // namespace

/* computeResiduals (spatialControl.cpp:97-155), exact arithmetic, with topographic distance when the settings ask for it */
extern "C" int pb_compute_residuals_exact(pb_context* ctx, const pb_stations* meteo, const float* observed, const pb_cv_points* cv,
                                          const pb_stations* points, const pb_settings* s, int variable, pb_raster_id dem,
                                          int excludeOutsideDem, int excludeSupplemental, float* residual)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (!meteo || !points || !s || !residual) return fail(ctx, PB_ERR_INVALID, "null argument");
    cudaSetDevice(ctx->device);
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    LooArrays a;
    std::vector<int> all(meteo->n);
    for (int i = 0; i < meteo->n; i++) all[i] = i;
    fillTargets(a, meteo, all, all, observed ? observed : meteo->value, cv ? cv->isInsideDem : nullptr, s->nProxy);
    fillSources(a, points, nullptr);
    return exactLoo(ctx, a, s, variable, dem, excludeOutsideDem, excludeSupplemental, residual, nullptr, nullptr);
}

/* spatialQualityControl (spatialControl.cpp:331-427); see include/praga_b200.h */
extern "C" int pb_spatial_quality_control(pb_context* ctx, const pb_stations* st, pb_settings* s, int variable, const pb_cv_points* cv,
                                          pb_raster_id dem, uint8_t* wrongSpatial)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (!st || !s || !wrongSpatial) return fail(ctx, PB_ERR_INVALID, "null argument");
    cudaSetDevice(ctx->device);
    const int n = st->n, P = s->nProxy;
    memset(wrongSpatial, 0, n);
    if (n == 0) return PB_OK;
    const float* observed = (cv && cv->currentValue) ? cv->currentValue : st->value;
    const uint8_t* inside = cv ? cv->isInsideDem : nullptr;
    std::vector<int> all(n);
    for (int i = 0; i < n; i++) all[i] = i;
    auto nrValidOf = [&](const std::vector<int>& idx) {
        int c = 0;
        for (int i : idx) {
            const int code = st->lapseRateCode ? st->lapseRateCode[i] : PB_LAPSE_PRIMARY;
            if (!s->useLapseRateCode || code == PB_LAPSE_PRIMARY || code == PB_LAPSE_SECONDARY) c++;
        }
        return c;
    };
    if (nrValidOf(all) == 0) return PB_OK;           // passDataToInterpolation returns false

    // ---- first pass: every accepted point ----
    SubStations ptsA(st, all, observed);
    passDataSideEffects(&ptsA.st, s);
    std::vector<uint8_t> keepA(n, 1);
    pb_cv_points cvA{};
    cvA.currentValue = observed;
    cvA.isInsideDem = inside;
    cvA.excludeOutsideDem = cv ? cv->excludeOutsideDem : 0;
    int rc = pb_pre_interpolation_ex(ctx, &ptsA.st, s, variable, &cvA, dem, keepA.data(), nullptr);
    if (rc) return rc;
    LooArrays a;
    fillTargets(a, st, all, all, observed, inside, P);
    fillSources(a, &ptsA.st, keepA.data());
    std::vector<float> residual(n);
    std::vector<NeighbourOut> nb(n);
    rc = exactLoo(ctx, a, s, variable, dem, 0, 0, residual.data(), nullptr, nb.data());
    if (rc) return rc;
    std::vector<int> listIndex;
    for (int i = 0; i < n; i++) {
        if (!nb[i].ok) continue;
        const float myValue = observed[i], myResidual = residual[i];
        const float stdDev = std::max(nb[i].devSt, myValue / 100.f);
        if (fabs(myResidual) > spatialThreshold(variable, s->rainfallThreshold, myValue, stdDev, 2, nb[i].avgDeltaZ, nb[i].minDistance)) {
            listIndex.push_back(i);
            wrongSpatial[i] = 1;
        }
    }
    if (listIndex.empty()) return PB_OK;

    // ---- second pass: the suspects against the field interpolated without them ----
    std::vector<int> rest;
    for (int i = 0; i < n; i++) if (!wrongSpatial[i]) rest.push_back(i);
    if (rest.empty() || nrValidOf(rest) == 0) return PB_OK;
    SubStations ptsB(st, rest, observed);
    passDataSideEffects(&ptsB.st, s);
    std::vector<uint8_t> keepB(rest.size(), 1);
    std::vector<float> obsB(rest.size());
    std::vector<uint8_t> inB(rest.size(), 1);
    for (size_t k = 0; k < rest.size(); k++) { obsB[k] = observed[rest[k]]; if (inside) inB[k] = inside[rest[k]]; }
    pb_cv_points cvB{};
    cvB.currentValue = obsB.data();
    cvB.isInsideDem = inB.data();
    cvB.excludeOutsideDem = cvA.excludeOutsideDem;
    rc = pb_pre_interpolation_ex(ctx, &ptsB.st, s, variable, &cvB, dem, keepB.data(), nullptr);
    if (rc) return rc;
    // interpolate() at the suspects; the reference passes meteoPoints[i].getProxyValues() with i = position in the list,
    // not the suspect's own index (spatialControl.cpp:393): kept
    std::vector<int> proxyFrom(listIndex.size());
    for (size_t k = 0; k < listIndex.size(); k++) proxyFrom[k] = int(k);
    LooArrays b;
    fillTargets(b, st, listIndex, proxyFrom, observed, inside, P);
    if (cv && cv->suspectProxyValues)
        for (size_t k = 0; k < listIndex.size(); k++)
            for (int q = 0; q < P; q++) b.tproxy[k * std::max(P, 1) + q] = cv->suspectProxyValues[k * size_t(st->nProxy) + q];
    fillSources(b, &ptsB.st, keepB.data());
    std::vector<float> est(listIndex.size());
    std::vector<float> dummy(listIndex.size());
    std::vector<NeighbourOut> nbB(listIndex.size());
    rc = exactLoo(ctx, b, s, variable, dem, 0, 0, dummy.data(), est.data(), nbB.data());
    if (rc) return rc;
    for (size_t k = 0; k < listIndex.size(); k++) {
        const int i = listIndex[k];
        if (nbB[k].ok) {
            const float myValue = observed[i];
            const float myResidual = est[k] - myValue;
            wrongSpatial[i] = fabs(myResidual) > spatialThreshold(variable, s->rainfallThreshold, myValue, nbB[k].devSt, 3,
                                                                  nbB[k].avgDeltaZ, nbB[k].minDistance)
                                  ? 1 : 0;
        } else wrongSpatial[i] = 0;
    }
    return PB_OK;
}

// PB_STAGE_TIMING=1: wall time a lane spends in each stage of pb_interpolation_dem, summed over the batch and printed on stderr
// when the batch returns (a diagnostic of where the lanes wait for each other)
static std::atomic<long long> g_stageUs[5];
static std::atomic<long long> g_stageCalls;
static bool stageTiming()
{
    static const bool on = getenv("PB_STAGE_TIMING") != nullptr;
    return on;
}
static void stageReport(const char* what, double wallMs, int nLanes)
{
    if (!stageTiming()) return;
    const long long n = std::max<long long>(1, g_stageCalls.exchange(0));
    fprintf(stderr, "[pb stages] %s: %lld griddings, %d lanes, wall %.2f ms; per gridding: qc %.3f pre %.3f raster %.3f loo %.3f other %.3f ms\n",
            what, n, nLanes, wallMs, g_stageUs[0].exchange(0) / 1e3 / n, g_stageUs[1].exchange(0) / 1e3 / n,
            g_stageUs[2].exchange(0) / 1e3 / n, g_stageUs[3].exchange(0) / 1e3 / n, g_stageUs[4].exchange(0) / 1e3 / n);
}

/* Project::interpolationDem (project/project.cpp:3106-3150) [+ the leave-one-out of Project::interpolationCv :3012-3104]:
 * one time step of one variable, from raw station values to the DEM raster; see include/praga_b200.h */
extern "C" int pb_interpolation_dem(pb_context* ctx, const pb_stations* meteo, pb_settings* qualitySettings, pb_settings* s, int variable,
                                    pb_raster_id dem, const pb_raster_id* proxyGrids, int nProxyGrids, int rowBegin, int rowEnd,
                                    int deviceOnly, pb_raster_out* out, pb_dem_step* step)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (!meteo || !s || !out) return fail(ctx, PB_ERR_INVALID, "null argument");
    cudaSetDevice(ctx->device);
    const int n = meteo->n;
    int launches = 0;
    float kernelMs = 0.f;
    int64_t h2d = 0, d2h = 0;
    // every stage below starts from a zeroed pb_timing (a stage that does not reset it would otherwise be counted with the
    // totals of the stage before) and is added up here
    auto account = [&]() {
        launches += ctx->timing.launches; kernelMs += ctx->timing.kernel_ms; h2d += ctx->timing.h2d_bytes; d2h += ctx->timing.d2h_bytes;
        memset(&ctx->timing, 0, sizeof(ctx->timing));
    };
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    std::vector<uint8_t> wrong(std::max(n, 1), 0);
    auto tStage = std::chrono::steady_clock::now();
    auto lap = [&](int k) {
        if (!stageTiming()) return;
        const auto now = std::chrono::steady_clock::now();
        g_stageUs[k] += std::chrono::duration_cast<std::chrono::microseconds>(now - tStage).count();
        tStage = now;
    };
    // checkData (spatialControl.cpp:466-476): the spatial check is skipped for precipitation and the wind vector components
    const bool spatialVar = variable != PB_VAR_PRECIPITATION && variable != PB_VAR_DAILY_PRECIPITATION && variable != PB_VAR_WIND_VECTOR_X &&
                            variable != PB_VAR_WIND_VECTOR_Y && variable != 36 /* windVectorDirection */ &&
                            variable != 41 /* dailyWindVectorDirectionPrevailing */;
    if (qualitySettings && spatialVar && n > 0) {
        int rc = pb_spatial_quality_control(ctx, meteo, qualitySettings, variable, nullptr, dem, wrong.data());
        if (rc) return rc;
        account();
    }
    if (step && step->wrongSpatial) memcpy(step->wrongSpatial, wrong.data(), n);
    lap(0);
    // passDataToInterpolation (spatialControl.cpp:500-566)
    std::vector<int> accepted;
    int nrValid = 0;
    for (int i = 0; i < n; i++) {
        if (wrong[i]) continue;
        accepted.push_back(i);
        const int code = meteo->lapseRateCode ? meteo->lapseRateCode[i] : PB_LAPSE_PRIMARY;
        if (!s->useLapseRateCode || code == PB_LAPSE_PRIMARY || code == PB_LAPSE_SECONDARY) nrValid++;
    }
    if (nrValid == 0) return fail(ctx, PB_ERR_INVALID, "No data available");
    SubStations pts(meteo, accepted, meteo->value);
    passDataSideEffects(&pts.st, s);
    if (s->useMultipleDetrending) s->nFit = s->nFitFunctions = 0;       // clearFitting (project.cpp:3540)
    std::vector<float> raw(pts.value);
    std::vector<uint8_t> keep(accepted.size(), 1);
    pb_cv_points cv{};
    cv.currentValue = raw.data();
    int rc = pb_pre_interpolation_ex(ctx, &pts.st, s, variable, &cv, dem, keep.data(), step ? step->khSeries : nullptr);
    if (rc) return rc;
    account();
    lap(1);
    std::vector<int> kept;
    for (size_t k = 0; k < accepted.size(); k++) if (keep[k]) kept.push_back(int(k));
    SubStations ipts(&pts.st, kept, pts.st.value);
    rc = deviceOnly ? pb_interpolation_raster_device(ctx, &ipts.st, s, variable, dem, proxyGrids, nProxyGrids, rowBegin, rowEnd, out)
                    : pb_interpolation_raster_resident(ctx, &ipts.st, s, variable, dem, proxyGrids, nProxyGrids, rowBegin, rowEnd, out);
    if (rc) return rc;
    account();
    lap(2);
    if (step && step->residual) {
        // computeResiduals(…, getUseExcludeStationsOutsideDEM(), true) over the accepted meteo points (project.cpp:3063-3067)
        SubStations mp(meteo, accepted, meteo->value);
        std::vector<float> res(accepted.size());
        LooArrays a;
        std::vector<int> all(accepted.size());
        for (size_t k = 0; k < all.size(); k++) all[k] = int(k);
        fillTargets(a, &mp.st, all, all, raw.data(), nullptr, s->nProxy);
        fillSources(a, &ipts.st, nullptr);
        rc = exactLoo(ctx, a, s, variable, dem, 0, 1, res.data(), nullptr, nullptr);
        if (rc) return rc;
        account();
        std::vector<float> obs, pre;
        for (int i = 0; i < n; i++) step->residual[i] = kNoData;
        for (size_t k = 0; k < accepted.size(); k++) {
            step->residual[accepted[k]] = res[k];
            const float value = raw[k];
            if (!(fabs(double(value) - double(kNoData)) < kEpsilon) && !(fabs(double(res[k]) - double(kNoData)) < kEpsilon)) {
                obs.push_back(value);
                pre.push_back(value - res[k]);
            }
        }
        cvStatisticsHost(obs, pre, step->stats);
    }
    lap(3);
    if (stageTiming()) g_stageCalls++;
    ctx->timing.launches = launches;
    ctx->timing.kernel_ms = kernelMs;
    ctx->timing.h2d_bytes = h2d;
    ctx->timing.d2h_bytes = d2h;
    return PB_OK;
}

// SMs a lane's raster kernel leaves free for the other lanes' station-space kernels (PB_LANE_RESERVE_SMS overrides)
static int laneReserveSms()
{
    static const int v = getenv("PB_LANE_RESERVE_SMS") ? atoi(getenv("PB_LANE_RESERVE_SMS")) : 16;
    return v;
}

/* A batch of independent (time step, variable) units — what the reference's hourly / daily drivers loop over — run through
 * pb_interpolation_dem on several LANES: child contexts of ctx with their own stream and buffers, each driven by its own
 * host thread, all reading the parent's resident rasters. One step alone is latency bound (a dozen small station-space
 * kernels and their host round trips around one raster kernel); with lanes the raster kernel of one step overlaps the
 * station-space work of the others. Every unit gives exactly the result of a stand-alone pb_interpolation_dem call. */
extern "C" int pb_interpolation_dem_batch(pb_context* ctx, pb_dem_batch_item* items, int nItems, pb_raster_id dem,
                                          const pb_raster_id* proxyGrids, int nProxyGrids, int rowBegin, int rowEnd, int deviceOnly,
                                          int nLanes)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (nItems < 0 || (nItems > 0 && !items)) return fail(ctx, PB_ERR_INVALID, "null argument");
    if (nItems == 0) return PB_OK;
    if (dem < 0 || dem >= (int)ctx->rasters.size() || !ctx->rasters[dem].used) return fail(ctx, PB_ERR_INVALID, "dem id is not a live raster");
    cudaSetDevice(ctx->device);
    if (nLanes <= 0) nLanes = 4;
    nLanes = std::min(std::min(nLanes, nItems), 16);
    // the lists the raster calls need are built once on the parent, then shared
    int rc = ensureValidList(ctx, ctx->rasters[dem]);
    if (rc) return rc;
    if (rowBegin != 0 || (rowEnd >= 0 && rowEnd != ctx->rasters[dem].h.nrRows)) {
        rc = ensureRowStart(ctx, ctx->rasters[dem]);
        if (rc) return rc;
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    while ((int)ctx->lanes.size() < nLanes) {
        pb_context* lane = nullptr;
        rc = pb_create(ctx->device, &lane);
        if (rc) return fail(ctx, rc, "could not create a lane context");
        ctx->lanes.push_back(lane);
    }
    for (int l = 0; l < nLanes; l++) {
        pb_context* lane = ctx->lanes[l];
        lane->rasters = ctx->rasters;                       // same ids, same device memory
        for (auto& e : lane->rasters) e.borrowed = true;
        lane->outInitRaster = -1;
        lane->parent = ctx;
        lane->reserveSms = nLanes > 1 ? laneReserveSms() : 0;
    }
    ctx->lastRaster = nullptr;
    std::atomic<int> next(0);
    std::vector<int> launches(nLanes, 0);
    std::vector<float> kernelMs(nLanes, 0.f);
    std::vector<int64_t> h2d(nLanes, 0), d2h(nLanes, 0);
    auto work = [&](int l) {
        pb_context* lane = ctx->lanes[l];
        cudaSetDevice(lane->device);
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= nItems) break;
            pb_dem_batch_item& it = items[i];
            if (!it.meteo || !it.settings || !it.out) { it.status = PB_ERR_INVALID; continue; }
            it.status = pb_interpolation_dem(lane, it.meteo, it.qualitySettings, it.settings, it.variable, dem, proxyGrids, nProxyGrids,
                                             rowBegin, rowEnd, deviceOnly, it.out, it.step);
            launches[l] += lane->timing.launches;
            kernelMs[l] += lane->timing.kernel_ms;
            h2d[l] += lane->timing.h2d_bytes;
            d2h[l] += lane->timing.d2h_bytes;
        }
    };
    std::vector<std::thread> threads;
    for (int l = 1; l < nLanes; l++) threads.emplace_back(work, l);
    work(0);
    for (auto& th : threads) th.join();
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    int firstError = PB_OK;
    for (int l = 0; l < nLanes; l++) {
        pb_context* lane = ctx->lanes[l];
        cudaStreamSynchronize(lane->stream);
        for (auto& e : lane->rasters) e = RasterEntry();    // views only: nothing to release
        lane->rasters.clear();
        ctx->timing.launches += launches[l];
        ctx->timing.kernel_ms += kernelMs[l];
        ctx->timing.h2d_bytes += h2d[l];
        ctx->timing.d2h_bytes += d2h[l];
    }
    for (int i = 0; i < nItems; i++)
        if (items[i].status != PB_OK && items[i].status != PB_ERR_NO_VALID_CELL && firstError == PB_OK) firstError = items[i].status;
    if (firstError != PB_OK) return fail(ctx, firstError, "a unit of the batch failed (see the items' status)");
    return PB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// pb_meteo_grid_step_batch: the hourly driver kept on the device (see include/praga_b200.h)
// ---------------------------------------------------------------------------------------------------------------------
namespace pb {
cudaError_t launchSpatialAggregateDev(const DevGrid& raster, const AggregationGeometry& g, int method, float* scratch, float* value,
                                      cudaStream_t s);
cudaError_t launchRelativeHumidityMapDev(const float* airT, float flagT, const float* dewT, float flagTd, long long n, float outFlag,
                                         float* out, unsigned* minmax, cudaStream_t s);
cudaError_t launchSamplePointsDev(const DevGrid& g, const double* x, const double* y, int n, float* value, unsigned char* inGrid,
                                  cudaStream_t s);
}  // namespace pb

namespace {

// tDewFromRelHum(float RH, float T), agrolib/meteo/meteo.cpp:275-285 (host arithmetic of the meteo point, like the reference)
float tDewFromRelHumHost(float RH, float T)
{
    auto isEq = [](float a, float b) { return fabs(double(a) - double(b)) < kEpsilon; };
    if (isEq(RH, kNoData) || isEq(T, kNoData) || RH == 0) return kNoData;
    RH = (100 < RH) ? 100 : RH;
    const double mySaturatedVaporPres = exp((16.78 * double(T) - 116.9) / (double(T) + 237.3));
    const double actualVaporPres = double(RH) / 100. * mySaturatedVaporPres;
    return float((log(actualVaporPres) * 237.3 + 116.9) / (16.78 - log(actualVaporPres)));
}

float keyToFloatS(unsigned k)
{
    unsigned u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// the map a lane has just gridded (lane->dOut, header of the DEM) -> the meteo-grid cells
int aggregateLaneMap(pb_context* lane, const AggregationGeometry& g, const DevGrid& map, int method, float* cellValue)
{
    if (method != PB_AGGR_AVERAGE && method != PB_AGGR_MEDIAN && method != PB_AGGR_STDDEV) {
        for (int c = 0; c < g.nCells; c++) cellValue[c] = kNoData;      // spatialAggregateMeteoGridPoint: NODATA (meteoGrid.cpp:1234-1237)
        return PB_OK;
    }
    CUDA_TRY(lane, lane->dAgg.reserve(size_t(std::max<long long>(g.nPoints, 1)) + size_t(g.nCells)));
    float* scratch = lane->dAgg.p;
    float* value = scratch + std::max<long long>(g.nPoints, 1);
    CUDA_TRY(lane, launchSpatialAggregateDev(map, g, method, scratch, value, lane->stream));
    CUDA_TRY(lane, cudaMemcpyAsync(cellValue, value, size_t(g.nCells) * 4, cudaMemcpyDeviceToHost, lane->stream));
    CUDA_TRY(lane, cudaStreamSynchronize(lane->stream));
    return PB_OK;
}

// one variable of one hour on a lane; tMapValid: lane->dMapT holds the hour's air temperature map
int meteoGridVar(pb_context* lane, pb_meteo_grid_var& v, pb_raster_id dem, const pb_raster_id* proxyGrids, int nProxyGrids,
                 const AggregationGeometry* g, bool& tMapValid, int& launches, int64_t& h2d, int64_t& d2h)
{
    if (!v.meteo || !v.settings) return fail(lane, PB_ERR_INVALID, "null argument");
    const RasterEntry& eDem = lane->rasters[dem];
    const size_t nCellsDem = eDem.n;
    DevGrid mapGrid = devGridOfRaster(eDem);
    pb_raster_out localOut{};
    pb_raster_out* out = v.out ? v.out : &localOut;
    auto account = [&]() { launches += lane->timing.launches; h2d += lane->timing.h2d_bytes; d2h += lane->timing.d2h_bytes; };

    if (v.variable == PB_VAR_AIR_REL_HUMIDITY && v.useDewPoint) {
        if (!tMapValid) return fail(lane, PB_ERR_INVALID, "the dew-point chain needs the air temperature map of the same hour before it");
        if (!v.airTemperature) return fail(lane, PB_ERR_INVALID, "the dew-point chain needs the stations' air temperature (PB_NODATA where missing)");
        const pb_stations* st = v.meteo;
        const int n = st->n;
        DevGrid tGrid = mapGrid;
        tGrid.data = lane->dMapT.p;
        // passInterpolatedTemperatureToHumidityPoints (project.cpp:2826-2848)
        std::vector<float> airT(v.airTemperature, v.airTemperature + n);
        std::vector<int> need;
        for (int i = 0; i < n; i++)
            if (!(fabs(double(st->value[i]) - double(kNoData)) < kEpsilon) && (fabs(double(airT[i]) - double(kNoData)) < kEpsilon)) need.push_back(i);
        if (!need.empty()) {
            const size_t m = need.size(), mAl = (m + 1) / 2 * 2;
            CUDA_TRY(lane, lane->dScratch.reserve(mAl * 16 + mAl * 4 + mAl + 64));
            std::vector<double> xy(2 * mAl);
            for (size_t k = 0; k < m; k++) { xy[k] = st->x[need[k]]; xy[mAl + k] = st->y[need[k]]; }
            double* dX = reinterpret_cast<double*>(lane->dScratch.p);
            double* dY = dX + mAl;
            float* dV = reinterpret_cast<float*>(dY + mAl);
            unsigned char* dIn = reinterpret_cast<unsigned char*>(dV + mAl);
            std::vector<float> val(m);
            std::vector<uint8_t> in(m);
            CUDA_TRY(lane, cudaMemcpyAsync(dX, xy.data(), 2 * mAl * 8, cudaMemcpyHostToDevice, lane->stream));
            CUDA_TRY(lane, launchSamplePointsDev(tGrid, dX, dY, int(m), dV, dIn, lane->stream));
            CUDA_TRY(lane, cudaMemcpyAsync(val.data(), dV, m * 4, cudaMemcpyDeviceToHost, lane->stream));
            CUDA_TRY(lane, cudaMemcpyAsync(in.data(), dIn, m, cudaMemcpyDeviceToHost, lane->stream));
            CUDA_TRY(lane, cudaStreamSynchronize(lane->stream));
            launches++;
            h2d += int64_t(2 * mAl * 8);
            d2h += int64_t(m * 5);
            for (size_t k = 0; k < m; k++) if (in[k]) airT[need[k]] = val[k];      // the map's value, its flag included
        }
        // the stations' dew point: Crit3DMeteoPoint::getMeteoPointValueH(airDewTemperature) (meteoPoint.cpp:967-972)
        std::vector<int> have;
        std::vector<float> tdew(n, kNoData);
        for (int i = 0; i < n; i++) {
            tdew[i] = tDewFromRelHumHost(st->value[i], airT[i]);
            if (!(fabs(double(tdew[i]) - double(kNoData)) < kEpsilon)) have.push_back(i);      // checkAndPassDataToInterpolation: valid values only
        }
        SubStations dew(st, have, tdew.data());
        pb_dem_step stepLocal{};
        std::vector<uint8_t> wrong(std::max<size_t>(have.size(), 1), 0);
        std::vector<float> resid(std::max<size_t>(have.size(), 1), kNoData);
        if (v.step) {
            stepLocal = *v.step;
            stepLocal.wrongSpatial = v.step->wrongSpatial ? wrong.data() : nullptr;
            stepLocal.residual = v.step->residual ? resid.data() : nullptr;
        }
        int rc = pb_interpolation_dem(lane, &dew.st, v.qualitySettings, v.settings, PB_VAR_AIR_DEW_TEMPERATURE, dem, proxyGrids,
                                      nProxyGrids, 0, -1, 1, out, v.step ? &stepLocal : nullptr);
        account();
        if (rc != PB_OK) return rc;
        if (v.step) {
            if (v.step->wrongSpatial) { memset(v.step->wrongSpatial, 0, n); for (size_t k = 0; k < have.size(); k++) v.step->wrongSpatial[have[k]] = wrong[k]; }
            if (v.step->residual) { for (int i = 0; i < n; i++) v.step->residual[i] = kNoData; for (size_t k = 0; k < have.size(); k++) v.step->residual[have[k]] = resid[k]; }
        }
        // computeRelativeHumidityMap (meteoMaps.cpp:301-321): in place over the dew-point map (a cell reads its own two inputs)
        lane->hMinMax[0] = 0xffffffffu; lane->hMinMax[1] = 0u; lane->hMinMax[2] = 0u; lane->hMinMax[3] = 0u;
        CUDA_TRY(lane, cudaMemcpyAsync(lane->dMinMax, lane->hMinMax, 4 * sizeof(unsigned), cudaMemcpyHostToDevice, lane->stream));
        CUDA_TRY(lane, launchRelativeHumidityMapDev(lane->dMapT.p, eDem.h.flag, lane->dOut.p, eDem.h.flag, (long long)nCellsDem,
                                                   eDem.h.flag, lane->dOut.p, lane->dMinMax, lane->stream));
        CUDA_TRY(lane, cudaMemcpyAsync(lane->hMinMax, lane->dMinMax, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, lane->stream));
        launches++;
        if (v.out && (v.out->data || v.out->rows)) {
            const int rows = eDem.h.nrRows, cols = eDem.h.nrCols;
            if (v.out->data)
                CUDA_TRY(lane, cudaMemcpyAsync(v.out->data, lane->dOut.p, nCellsDem * 4, cudaMemcpyDeviceToHost, lane->stream));
            else
                for (int r = 0; r < rows; r++)
                    CUDA_TRY(lane, cudaMemcpyAsync(v.out->rows[r], lane->dOut.p + size_t(r) * cols, size_t(cols) * 4, cudaMemcpyDeviceToHost,
                                                   lane->stream));
            d2h += int64_t(nCellsDem * 4);
        }
        CUDA_TRY(lane, cudaStreamSynchronize(lane->stream));
        if (lane->hMinMax[0] == 0xffffffffu) { out->minimum = kNoData; out->maximum = kNoData; }
        else { out->minimum = keyToFloatS(lane->hMinMax[0]); out->maximum = keyToFloatS(lane->hMinMax[1]); }
    } else {
        const int deviceOnly = !(v.out && (v.out->data || v.out->rows));
        int rc = pb_interpolation_dem(lane, v.meteo, v.qualitySettings, v.settings, v.variable, dem, proxyGrids, nProxyGrids, 0, -1,
                                      deviceOnly, out, v.step);
        account();
        if (rc != PB_OK && rc != PB_ERR_NO_VALID_CELL) return rc;
        if (v.variable == PB_VAR_AIR_TEMPERATURE) {
            // mapHourlyTair of this hour, kept for the humidity chain
            CUDA_TRY(lane, lane->dMapT.reserve(nCellsDem));
            CUDA_TRY(lane, cudaMemcpyAsync(lane->dMapT.p, lane->dOut.p, nCellsDem * 4, cudaMemcpyDeviceToDevice, lane->stream));
            tMapValid = true;
        }
        if (rc == PB_ERR_NO_VALID_CELL) {
            if (g && v.cellValue) for (int c = 0; c < g->nCells; c++) v.cellValue[c] = kNoData;
            return rc;
        }
    }
    if (g && v.cellValue) {
        mapGrid.data = lane->dOut.p;
        int rc = aggregateLaneMap(lane, *g, mapGrid, v.aggrMethod, v.cellValue);
        if (rc) return rc;
        launches++;
        d2h += int64_t(g->nCells) * 4;
    }
    return PB_OK;
}

}  // namespace

extern "C" int pb_meteo_grid_step_batch(pb_context* ctx, pb_meteo_grid_hour* hours, int nHours, pb_raster_id dem,
                                        const pb_raster_id* proxyGrids, int nProxyGrids, int32_t aggregation, int nLanes)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (nHours < 0 || (nHours > 0 && !hours)) return fail(ctx, PB_ERR_INVALID, "null argument");
    if (nHours == 0) return PB_OK;
    if (dem < 0 || dem >= (int)ctx->rasters.size() || !ctx->rasters[dem].used) return fail(ctx, PB_ERR_INVALID, "dem id is not a live raster");
    const AggregationGeometry* g = nullptr;
    if (aggregation >= 0) {
        if (aggregation >= (int)ctx->aggregations.size() || !ctx->aggregations[aggregation].used)
            return fail(ctx, PB_ERR_INVALID, "bad aggregation id");
        g = &ctx->aggregations[aggregation];
    }
    cudaSetDevice(ctx->device);
    if (nLanes <= 0) nLanes = 4;
    nLanes = std::min(std::min(nLanes, nHours), 16);
    int rc = ensureValidList(ctx, ctx->rasters[dem]);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    while ((int)ctx->lanes.size() < nLanes) {
        pb_context* lane = nullptr;
        rc = pb_create(ctx->device, &lane);
        if (rc) return fail(ctx, rc, "could not create a lane context");
        ctx->lanes.push_back(lane);
    }
    for (int l = 0; l < nLanes; l++) {
        pb_context* lane = ctx->lanes[l];
        lane->rasters = ctx->rasters;                       // same ids, same device memory
        for (auto& e : lane->rasters) e.borrowed = true;
        lane->outInitRaster = -1;
        lane->parent = ctx;
        lane->reserveSms = nLanes > 1 ? laneReserveSms() : 0;
    }
    ctx->lastRaster = nullptr;
    std::atomic<int> next(0);
    std::vector<int> launches(nLanes, 0);
    std::vector<int64_t> h2d(nLanes, 0), d2h(nLanes, 0);
    auto work = [&](int l) {
        pb_context* lane = ctx->lanes[l];
        cudaSetDevice(lane->device);
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= nHours) break;
            pb_meteo_grid_hour& hour = hours[i];
            bool tMapValid = false;
            for (int k = 0; k < hour.nVars; k++) {
                pb_meteo_grid_var& v = hour.vars[k];
                v.status = meteoGridVar(lane, v, dem, proxyGrids, nProxyGrids, g, tMapValid, launches[l], h2d[l], d2h[l]);
            }
        }
    };
    const auto tBatch = std::chrono::steady_clock::now();
    std::vector<std::thread> threads;
    for (int l = 1; l < nLanes; l++) threads.emplace_back(work, l);
    work(0);
    for (auto& th : threads) th.join();
    stageReport("pb_meteo_grid_step_batch", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tBatch).count(), nLanes);
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    int firstError = PB_OK;
    std::string firstMsg;
    for (int l = 0; l < nLanes; l++) {
        pb_context* lane = ctx->lanes[l];
        cudaStreamSynchronize(lane->stream);
        for (auto& e : lane->rasters) e = RasterEntry();    // views only: nothing to release
        lane->rasters.clear();
        ctx->timing.launches += launches[l];
        ctx->timing.h2d_bytes += h2d[l];
        ctx->timing.d2h_bytes += d2h[l];
    }
    for (int i = 0; i < nHours; i++)
        for (int k = 0; k < hours[i].nVars; k++) {
            const int st = hours[i].vars[k].status;
            if (st != PB_OK && st != PB_ERR_NO_VALID_CELL && firstError == PB_OK) firstError = st;
        }
    if (firstError != PB_OK) return fail(ctx, firstError, "a variable of the batch failed (see the variables' status)");
    return PB_OK;
}

