// glocal.cu — glocal detrending: one multiple-detrending fit per MACRO AREA, cells blended over the
// areas they belong to. Reference (paths relative to agrolib/):
//   glocalDetrendingFitting                          interpolation/interpolation.cpp:2239-2294 (from preInterpolation :2471-2474)
//   macroAreaDetrending                              :1759-1829
//   getSignificantProxyValuesXY                      :2650-2677
//   loop of Project::interpolationDemGlocalDetrending project/project.cpp:3299-3375
//   computeResidualsGlocalDetrending                 interpolation/spatialControl.cpp:231-302
//   Project::computeResidualsAndStatisticsGlocalDetrending  project/project.cpp:6525-6565
//
// Mapping:
//   * FIT: all areas in ONE launch of the batched preInterpolation kernel (K3, pre_kernels.cu): a unit is an area = a station
//     subset of the same time step, one lane group per area, lane g = first guess g of the UNWEIGHTED elevation fit
//     (bestFittingMarquardt_nDimension, mathFunctions/furtherMathFunctions.cpp:1433-1529). Parameters bit-identical.
//   * RASTER: per area, in index order on one stream: the area's (cell, weight) pairs become explicit targets
//     (glocal_targets_kernel: cell centre of getUtmXYFromRowCol rounded to f32, significant proxies looked up), the IDW
//     station loop of K1 in point form grids them over the area's detrended stations, glocal_blend_kernel accumulates
//     value * weight into the output (first touch replaces NODATA). The few-hundred-station detrend of an area
//     (detrendingElevation :1956-1972, detrendingOtherProxies :2173-2236) is host arithmetic in the reference's float/double
//     mix (this TU is built without FMA contraction on either side).
//   * LOO: per area the exact leave-one-out kernel (pb_compute_residuals_exact) on the area's stations; sums over areas on
//     the host like the reference's `residual +=`.
#include <math.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#include "context.cuh"
#include "epilogue.cuh"
#include "kh_search.cuh"
#include "pre.cuh"
#include "shepard.cuh"
#include "nvtx.cuh"

namespace pb {
// engine.cu
int buildDevModelGrids(pb_context* ctx, const pb_settings* s, int variable, const pb_raster_id* proxyIds, int nProxyGrids, DevModel& m);
size_t idwStationsStagingBytes(int nStations);
void packIdwStationsAt(const pb_stations* st, const pb_settings* s, bool excludeSupplemental, unsigned char* h, unsigned char* d,
                       DevStations& ds, double& valueOffset);
int buildPreModel(pb_context* ctx, const pb_settings* s, int variable, int nStations, LocalModel& m);
int stageShepardStationsFor(pb_context* ctx, const pb_stations* st, const pb_settings* s, bool excludeSupplemental,
                            ShepardStations& ss, float& R0);
int stageAllStations(pb_context* ctx, const pb_stations* st, LocalStations& ds);
void applyPreFitToSettings(const pb_settings& s, const LocalModel& m, const PreFit& f, pb_settings& o);
DevGrid devGridOfRaster(const RasterEntry& r);
bool varUsesDetrending(int v);
bool varUsesTd(int v);
// idw_kernels.cu
cudaError_t launchIdwPoints(const DevStations& st, const DevModel& m, const float* tx, const float* ty, const double* tproxy,
                            int nTargets, float* out, cudaStream_t s);
cudaError_t launchFill(float* out, long long n, float v, cudaStream_t s);
// td_kernels.cu
cudaError_t launchIdwTdPoints(const DevStations& st, const float* sz, const DevModel& m, const DevGrid& dem, int kh,
                              const float* tx, const float* ty, const float* tz, const double* tproxy, int nTargets, float* out,
                              cudaStream_t s, const DevGrid* tdMaps = nullptr);

// which proxies getSignificantProxyValuesXY looks up (active && significant, grid loaded)
struct GlocalLookup {
    int nProxy;
    int slot[PB_MAX_PROXY];          // -1: value stays NODATA
    DevGrid grid[PB_MAX_PROXY];
};

enum : unsigned char { CELL_SKIP = 0, CELL_NODATA = 1, CELL_OK = 2 };

// Crit3DRasterGrid::getValueFromXY (gis.cpp:533-546)
__device__ __forceinline__ float glocalGridValueXY(const DevGrid& g, double x, double y)
{
    const int r = int((y - g.lly) * g.invCellSize);
    const int row = (g.nrRows - 1) - r;
    const int col = int((x - g.llx) * g.invCellSize);
    if (unsigned(row) >= unsigned(g.nrRows) || unsigned(col) >= unsigned(g.nrCols)) return g.flag;
    return __ldg(g.data + (size_t)row * g.nrCols + col);
}

// row / col of an areaCells entry exactly as project.cpp:3337-3338 computes them: a FLOAT division for the row, an integer
// remainder for the column
__device__ __forceinline__ void glocalRowCol(float idxf, int nrCols, int& row, int& col)
{
    row = int(__fdiv_rn(idxf, float(nrCols)));
    col = int(idxf) % nrCols;
}

// the area's cells -> explicit targets: getUtmXYFromRowCol (gis.cpp:806-810) rounded to f32 (the float parameters of
// getSignificantProxyValuesXY and interpolate), the significant proxies' values
__global__ void __launch_bounds__(256)
glocal_targets_kernel(const float* __restrict__ cells, int nCells, DevGrid dem, GlocalLookup lk, float* __restrict__ tx,
                      float* __restrict__ ty, float* __restrict__ tz, double* __restrict__ pv, unsigned char* __restrict__ state)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    int row, col;
    glocalRowCol(cells[2 * c], dem.nrCols, row, col);
    unsigned char st = CELL_SKIP;
    float xf = 0.f, yf = 0.f, zf = 0.f;
    double v[PB_MAX_PROXY];
#pragma unroll
    for (int i = 0; i < PB_MAX_PROXY; i++) v[i] = -9999.;
    if (unsigned(row) < unsigned(dem.nrRows) && unsigned(col) < unsigned(dem.nrCols)) {
        const float z = __ldg(dem.data + (size_t)row * dem.nrCols + col);
        if (!isEqualF(z, dem.flag)) {
            const double x = dem.llx + dem.cellSize * (col + 0.5);
            const double y = dem.lly + dem.cellSize * (dem.nrRows - row - 0.5);
            xf = float(x);
            yf = float(y);
            zf = z;                                  // interpolate(..., x, y, z = DEM.value[row][col], ...) (project.cpp:3340, 3356)
            bool complete = true;
#pragma unroll
            for (int i = 0; i < PB_MAX_PROXY; i++)
                if (i < lk.nProxy && lk.slot[i] >= 0) {
                    const DevGrid& g = lk.grid[lk.slot[i]];
                    const float gv = glocalGridValueXY(g, double(xf), double(yf));
                    if (gv != g.flag) v[i] = double(gv);
                    else complete = false;
                }
            st = complete ? CELL_OK : CELL_NODATA;
        }
    }
    tx[c] = xf;
    ty[c] = yf;
    tz[c] = zf;
    state[c] = st;
    for (int i = 0; i < lk.nProxy; i++) pv[(size_t)c * lk.nProxy + i] = v[i];
}

// out[cell] (+)= interpolate() * weight (project.cpp:3361-3369); an incomplete proxy set leaves NODATA (:3350-3354)
__global__ void __launch_bounds__(256)
glocal_blend_kernel(const float* __restrict__ cells, int nCells, int nrCols, const unsigned char* __restrict__ state,
                    const float* __restrict__ res, float* __restrict__ out, int* __restrict__ notOk)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    const unsigned char st = state[c];
    if (st == CELL_SKIP) return;
    int row, col;
    glocalRowCol(cells[2 * c], nrCols, row, col);
    float* cell = out + (size_t)row * nrCols + col;
    if (st == CELL_NODATA) { *cell = kNoData; return; }
    const double interpolatedValue = double(res[c]);
    if (fabs(interpolatedValue - (-9999.)) < kEpsilon) *notOk = 1;
    const double prod = interpolatedValue * double(cells[2 * c + 1]);
    const float cur = *cell;
    *cell = isEqualF(cur, kNoData) ? float(prod) : float(double(cur) + prod);
}

// One macro area in ONE launch (plain idw, no topographic distance): thread = an entry of the area's (cell, weight) list.
// glocal_targets_kernel + the K1 point form + glocal_blend_kernel cost three launches and 28 B of scratch traffic per entry
// for ~140 station pairs of arithmetic; here the entry is decoded, its significant proxies are looked up, the area's stations
// (shared memory, <= kGlocalTile per pass) are gridded with K1's weight (rsqrt(d2)^3: the 1/10000 scale cancels in the ratio),
// FP32 over chunks of 32 stations and FP64 across chunks, a station ON the cell takes the exact loop of the reference
// (f32 distance, f64 weights, d > 0), then retrend / clamp with the area's model and the blend of project.cpp:3361-3369.
// Areas are launched in index order on one stream, a cell occurs once per area: the blend order is the reference's.
constexpr int kGlocalTile = 512;          // stations per shared-memory pass (K1's packed tile: {x, x, y, y} + {v, v})
struct GlocalEntry {                       // one (cell, weight) entry of an area, decoded
    int row, col;
    bool live, complete;
    float xf, yf;
    double pv[PB_MAX_PROXY];
};
__device__ __forceinline__ void glocalDecode(const float* __restrict__ cells, int c, int nCells, const DevGrid& dem, const GlocalLookup& lk,
                                             GlocalEntry& e)
{
    e.row = e.col = 0;
    e.live = false;
    e.complete = true;
    e.xf = e.yf = 0.f;
#pragma unroll
    for (int i = 0; i < PB_MAX_PROXY; i++) e.pv[i] = -9999.;
    if (c >= nCells) return;
    glocalRowCol(cells[2 * c], dem.nrCols, e.row, e.col);
    if (unsigned(e.row) >= unsigned(dem.nrRows) || unsigned(e.col) >= unsigned(dem.nrCols)) return;
    const float z = __ldg(dem.data + (size_t)e.row * dem.nrCols + e.col);
    if (isEqualF(z, dem.flag)) return;
    e.live = true;
    e.xf = float(dem.llx + dem.cellSize * (e.col + 0.5));
    e.yf = float(dem.lly + dem.cellSize * (dem.nrRows - e.row - 0.5));
#pragma unroll
    for (int i = 0; i < PB_MAX_PROXY; i++)
        if (i < lk.nProxy && lk.slot[i] >= 0) {
            const DevGrid& g = lk.grid[lk.slot[i]];
            const float gv = glocalGridValueXY(g, double(e.xf), double(e.yf));
            if (gv != g.flag) e.pv[i] = double(gv);
            else e.complete = false;
        }
}
__device__ __forceinline__ void glocalFinish(const float* __restrict__ cells, int c, const GlocalEntry& e, const DevGrid& dem,
                                             const DevStations& st, const DevModel& m, double SW, double SWV, float* __restrict__ out,
                                             int* __restrict__ notOk)
{
    if (!e.live) return;
    float* cell = out + (size_t)e.row * dem.nrCols + e.col;
    if (!e.complete) { *cell = kNoData; return; }             // an incomplete proxy set leaves NODATA (:3350-3354)
    float result;
    if (m.useRetrendOnly) result = 0.f;
    else {
        if (!(SW <= 1.0e300) || !(fabs(SWV) <= 1.0e300)) {
            // a station on the cell (inf / NaN sums): computeDistances + inverseDistanceWeighted in the reference's arithmetic
            // (:208-256, :1031-1051); DevStations::val is centred on the host
            SW = 0.;
            SWV = 0.;
            for (int i = 0; i < st.n; i++) {
                const float dx = st.x[i] - e.xf, dy = st.y[i] - e.yf;
                const float d = sqrtf(dx * dx + dy * dy);
                if (d > 0.f) {
                    const double dk = double(d) / 10000.;
                    const double w = 1.0 / (dk * dk * dk);
                    SW += w;
                    SWV += double(st.val[i]) * w;
                }
            }
        }
        result = SW > 0. ? float(SWV / SW + m.valueOffset) : kNoData;
    }
    const double interpolatedValue = double(finishEstimate(m, result, e.pv));
    if (fabs(interpolatedValue - (-9999.)) < kEpsilon) *notOk = 1;
    const double prod = interpolatedValue * double(cells[2 * c + 1]);
    const float cur = *cell;
    *cell = isEqualF(cur, kNoData) ? float(prod) : float(double(cur) + prod);
}

__global__ void __launch_bounds__(256)
glocal_area_kernel(const float* __restrict__ cells, int nCells, DevGrid dem, GlocalLookup lk, DevStations st, DevModel m,
                   float* __restrict__ out, int* __restrict__ notOk)
{
    __shared__ float4 sxy[kGlocalTile];
    __shared__ float2 sv[kGlocalTile];
    // two neighbouring entries per thread: K1's packed FP32x2 station loop
    const int c0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    GlocalEntry ea, eb;
    glocalDecode(cells, c0, nCells, dem, lk, ea);
    glocalDecode(cells, c0 + 1, nCells, dem, lk, eb);
    const bool work = (ea.live && ea.complete) || (eb.live && eb.complete);
    const f32x2 nx = pack2(-ea.xf, -eb.xf), ny = pack2(-ea.yf, -eb.yf);
    double SWa = 0., SWVa = 0., SWb = 0., SWVb = 0.;
    const int used = (st.n + 31) / 32 * 32;                 // the padding of the packed tile weighs +0
    PB_DEV_ASSERT(used <= st.nPadded);
    for (int j0 = 0; j0 < used; j0 += kGlocalTile) {
        const int nj = min(kGlocalTile, used - j0);
        __syncthreads();
        for (int j = threadIdx.x; j < nj; j += blockDim.x) { sxy[j] = st.xy[j0 + j]; sv[j] = st.v[j0 + j]; }
        __syncthreads();
        if (work && !m.useRetrendOnly) {
            for (int b = 0; b < nj; b += 32) {
                f32x2 sw = pack2(0.f, 0.f), swv = pack2(0.f, 0.f);
#pragma unroll 8
                for (int j = b; j < b + 32; j++) {
                    const float4 pxy = sxy[j];
                    const float2 pvv = sv[j];
                    const f32x2 dx = add2(pack2(pxy.x, pxy.y), nx), dy = add2(pack2(pxy.z, pxy.w), ny);
                    const f32x2 d2 = fma2(dx, dx, mul2(dy, dy));
                    float a, bq;
                    unpack2(d2, a, bq);
                    const f32x2 r = pack2(rsqrt_approx(a), rsqrt_approx(bq));
                    const f32x2 w = mul2(mul2(r, r), r);
                    sw = add2(sw, w);
                    swv = fma2(w, pack2(pvv.x, pvv.y), swv);
                }
                float a, bq, va, vb;
                unpack2(sw, a, bq);
                unpack2(swv, va, vb);
                SWa += double(a); SWb += double(bq);
                SWVa += double(va); SWVb += double(vb);
            }
        }
    }
    glocalFinish(cells, c0, ea, dem, st, m, SWa, SWVa, out, notOk);
    glocalFinish(cells, c0 + 1, eb, dem, st, m, SWb, SWVb, out, notOk);
}

__device__ __forceinline__ unsigned orderedKeyG(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// gis::updateMinMaxRasterGrid (gis.cpp:569-614) over the blended grid
__global__ void __launch_bounds__(256) glocal_minmax_kernel(const float* __restrict__ v, long long n, float flag, unsigned* __restrict__ minmax)
{
    float vmin = INFINITY, vmax = -INFINITY;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const float x = v[k];
        if (!isEqualF(x, flag) && !isEqualF(x, kNoData)) { vmin = fminf(vmin, x); vmax = fmaxf(vmax, x); }
    }
    for (int o = 16; o; o >>= 1) {
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (vmin != INFINITY) atomicMin(&minmax[0], orderedKeyG(vmin));
        if (vmax != -INFINITY) atomicMax(&minmax[1], orderedKeyG(vmax));
    }
}

}  // namespace pb

using namespace pb;

namespace {

int fail(pb_context* ctx, int code, const std::string& msg) { return pb::failCtx(ctx, code, msg); }

float keyToFloatG(unsigned k)
{
    unsigned u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

bool isEqualDh(double a, double b) { return fabs(a - b) < kEpsilon; }

// model functions (mathFunctions/furtherMathFunctions.cpp:115-201) on the host, for the per-area detrend of a few hundred
// stations; par is mutated by the three-piece forms like in the reference
double evalFitHost(int f, double x, double* par)
{
    switch (f) {
    case PB_FIT_PIECEWISE_TWO:
        if (x < par[0]) return par[2] * (x - par[0]) + par[1];
        return par[3] * (x - par[0]) + par[1];
    case PB_FIT_PIECEWISE_THREE: {
        par[2] = std::max(10., par[2]);
        const double xb = par[2] + par[0];
        if (x < par[0]) return par[4] * x - par[0] * par[4] + par[1];
        else if (x > xb) return par[4] * x - par[4] * par[0] - par[4] * par[2] + par[3] * par[2] + par[1];
        else return par[3] * x - par[3] * par[0] + par[1];
    }
    case PB_FIT_PIECEWISE_THREE_FREE: {
        par[2] = std::max(10., par[2]);
        const double xb = par[0] + par[2];
        if (x < par[0]) return par[4] * x - par[0] * par[4] + par[1];
        else if (x > xb) return par[5] * x - par[5] * par[0] - par[5] * par[2] + par[3] * par[2] + par[1];
        else return par[3] * x - par[3] * par[0] + par[1];
    }
    case PB_FIT_LINEAR: return par[0] * x + par[1];
    default: return 0.;
    }
}

int elevationPosOf(const pb_settings& s)
{
    int e = -1;
    for (int pos = 0; pos < s.nProxy; pos++) if (s.proxy[pos].kind == PB_PROXY_HEIGHT) e = pos;
    return e;
}

// the stations of one area after macroAreaDetrending (:1759-1829), with the settings interpolate() then sees
struct AreaStations {
    std::vector<double> x, y, z;
    std::vector<float> value, weight, proxy;
    std::vector<int32_t> code, index;
    std::vector<uint8_t> active;
    pb_stations st{};
    pb_settings s{};
};

void macroAreaDetrendingHost(const pb_macro_area& A, const pb_settings& base, const pb_stations* pts, AreaStations& out)
{
    pb_settings& s = out.s;
    s = base;
    s.useGlocalDetrending = 0;     // interpolate() itself never looks at it
    // setFittingParameters(area), setCurrentCombination(area), fitting functions by parameter count (:1766-1785)
    s.nFit = A.nParameters;
    for (int p = 0; p < PB_MAX_PROXY; p++) {
        s.fitNrParams[p] = p < A.nParameters ? A.nrParams[p] : 0;
        for (int j = 0; j < PB_MAX_FIT_PARAMS; j++) s.fitParams[p][j] = p < A.nParameters ? A.parameters[p][j] : 0.;
    }
    for (int p = 0; p < s.nProxy; p++) { s.currentActive[p] = A.active[p]; s.currentSignificant[p] = A.significant[p]; }
    s.currentUseThermalInversion = A.useThermalInversion;
    s.nFitFunctions = 0;
    for (int l = 0; l < A.nParameters; l++) {
        const int n = A.nrParams[l];
        if (n == 2) s.fitFunction[s.nFitFunctions++] = PB_FIT_LINEAR;
        else if (n == 4) s.fitFunction[s.nFitFunctions++] = PB_FIT_PIECEWISE_TWO;
        else if (n == 5) s.fitFunction[s.nFitFunctions++] = PB_FIT_PIECEWISE_THREE;
        else if (n == 6) s.fitFunction[s.nFitFunctions++] = PB_FIT_PIECEWISE_THREE_FREE;
    }
    // the area's interpolation points, in the area's order (:1794-1803: every match, no break)
    const int P = std::max(pts->nProxy, 0);
    std::vector<int> sel;
    for (int l = 0; l < A.nMeteoPoints; l++)
        for (int k = 0; k < pts->n; k++)
            if ((pts->index ? pts->index[k] : k) == A.meteoPoints[l]) sel.push_back(k);
    const int e = elevationPosOf(s);
    auto proxyOf = [&](int k, int pos) { return pos < P ? pts->proxyValues[size_t(k) * P + pos] : kNoData; };
    std::vector<float> value(sel.size());
    for (size_t q = 0; q < sel.size(); q++) value[q] = pts->value[sel[q]];
    // detrendingElevation (:1956-1972)
    if (e >= 0 && A.active[e] && A.significant[e] && s.nFitFunctions > 0 && s.nFit > 0) {
        double par[PB_MAX_FIT_PARAMS];
        memcpy(par, s.fitParams[0], sizeof(par));
        for (size_t q = 0; q < sel.size(); q++) {
            const float proxyValue = proxyOf(sel[q], e);
            const float detrendValue = float(evalFitHost(s.fitFunction[0], proxyValue, par));
            value[q] -= detrendValue;
        }
    }
    // detrendingOtherProxies (:2173-2236)
    std::vector<uint8_t> kept(sel.size(), 1);
    if (s.nFit > 0) {
        const int first = s.fitNrParams[0] > 2 ? 1 : 0;
        if (s.nFit - first > 0) {
            const int nF = s.nFitFunctions - first;
            for (size_t q = 0; q < sel.size(); q++) {
                double pvv[PB_MAX_PROXY];
                int n = 0;
                bool complete = true;
                for (int pos = 0; pos < s.nProxy; pos++)
                    if (pos != e && s.currentActive[pos] && s.currentSignificant[pos] && complete) {
                        const double v = proxyOf(sel[q], pos);
                        if (!isEqualDh(v, -9999)) pvv[n++] = v;
                        else complete = false;
                    }
                if (complete) {
                    double r = 0.0;
                    for (int c = 0; c < nF && c < n; c++) {
                        double par[PB_MAX_FIT_PARAMS];
                        memcpy(par, s.fitParams[first + c], sizeof(par));
                        r += evalFitHost(s.fitFunction[first + c], pvv[c], par);
                    }
                    value[q] -= float(r);
                } else kept[q] = 0;
            }
        }
    }
    for (size_t q = 0; q < sel.size(); q++) {
        if (!kept[q]) continue;
        const int k = sel[q];
        out.x.push_back(pts->x[k]); out.y.push_back(pts->y[k]); out.z.push_back(pts->z[k]);
        out.value.push_back(value[q]);
        out.weight.push_back(pts->regressionWeight ? pts->regressionWeight[k] : kNoData);
        out.code.push_back(pts->lapseRateCode ? pts->lapseRateCode[k] : PB_LAPSE_PRIMARY);
        out.active.push_back(pts->isActive ? pts->isActive[k] : 1);
        out.index.push_back(pts->index ? pts->index[k] : k);
        for (int pos = 0; pos < P; pos++) out.proxy.push_back(pts->proxyValues[size_t(k) * P + pos]);
    }
    if (out.proxy.empty()) out.proxy.push_back(0.f);
    pb_stations& st = out.st;
    st.n = int(out.x.size());
    st.nProxy = pts->nProxy;
    st.x = out.x.data(); st.y = out.y.data(); st.z = out.z.data();
    st.value = out.value.data();
    st.regressionWeight = out.weight.data();
    st.lapseRateCode = out.code.data();
    st.isActive = out.active.data();
    st.proxyValues = out.proxy.data();
    st.index = out.index.data();
}

// setMultipleDetrendingHeightTemperatureRange (:1506-1553) as far as it shows in pb_settings beyond what the fit already
// wrote: one more fitting function per active height proxy with parameter ranges
void secondHeightRangeCall(pb_settings& s)
{
    if (!s.hasPointsRange || !s.useMultipleDetrending) return;
    for (int i = 0; i < s.nProxy; i++)
        if (s.selectedActive[i] && s.proxy[i].kind == PB_PROXY_HEIGHT && s.proxy[i].nFitParams > 0) {
            const int f = s.proxy[i].fittingFunction;
            if ((f == PB_FIT_PIECEWISE_TWO || f == PB_FIT_PIECEWISE_THREE_FREE || f == PB_FIT_PIECEWISE_THREE) &&
                s.nFitFunctions < PB_MAX_PROXY)
                s.fitFunction[s.nFitFunctions++] = f;
        }
}

// the area's active meteo points, in the area's order, as the targets of a leave-one-out (computeResiduals skips the others)
struct AreaMeteo {
    std::vector<int> who;
    std::vector<double> x, y, z;
    std::vector<float> val, proxy;
    std::vector<int32_t> code;
    std::vector<uint8_t> inside;
    pb_stations mp{};
    pb_cv_points cva{};
};

int areaMeteoPoints(pb_context* ctx, const pb_macro_area& A, const pb_stations* meteo, const float* obs, const pb_cv_points* cv,
                    AreaMeteo& o)
{
    const int P = std::max(meteo->nProxy, 0);
    for (int i = 0; i < A.nMeteoPoints; i++) {
        const int j = A.meteoPoints[i];
        if (j < 0 || j >= meteo->n) return fail(ctx, PB_ERR_INVALID, "macro area meteo point out of range");
        if (meteo->isActive && !meteo->isActive[j]) continue;
        o.who.push_back(j);
    }
    const size_t m = o.who.size();
    o.x.resize(m); o.y.resize(m); o.z.resize(m); o.val.resize(m); o.code.resize(m);
    o.proxy.assign(std::max<size_t>(m * P, 1), kNoData);
    o.inside.assign(std::max<size_t>(m, 1), 1);
    for (size_t q = 0; q < m; q++) {
        const int j = o.who[q];
        o.x[q] = meteo->x[j]; o.y[q] = meteo->y[j]; o.z[q] = meteo->z[j];
        o.val[q] = obs[j];
        o.code[q] = meteo->lapseRateCode ? meteo->lapseRateCode[j] : PB_LAPSE_PRIMARY;
        if (cv && cv->isInsideDem) o.inside[q] = cv->isInsideDem[j];
        for (int p = 0; p < P; p++) o.proxy[q * P + p] = meteo->proxyValues[size_t(j) * P + p];
    }
    o.mp = pb_stations{};
    o.mp.n = int(m);
    o.mp.nProxy = meteo->nProxy;
    o.mp.x = o.x.data(); o.mp.y = o.y.data(); o.mp.z = o.z.data();
    o.mp.value = o.val.data();
    o.mp.lapseRateCode = o.code.data();
    o.mp.proxyValues = o.proxy.data();
    o.cva = pb_cv_points{};
    o.cva.isInsideDem = o.inside.data();
    return PB_OK;
}

// topographicDistanceOptimize at the end of macroAreaDetrending (:1811-1826): golden-section search of kh over the
// leave-one-out MAE of the AREA's meteo points against the area's detrended stations (computeResiduals(..., true, true)), each
// visited kh one exact leave-one-out launch. sub.s.topoDist_Kh receives the result.
int areaKhSearch(pb_context* ctx, const AreaMeteo& am, AreaStations& sub, int variable, pb_raster_id demId, pb_timing& sum)
{
    const int maxKh = std::max(0, sub.s.topoDist_maxKh);
    const int m = am.mp.n;
    std::vector<float> res(std::max(m, 1));
    std::vector<float> table(size_t(maxKh) + 1, 0.f);
    std::vector<uint8_t> known(size_t(maxKh) + 1, 0);
    int rc = PB_OK;
    auto f = [&](int kh) -> float {
        if (m == 0 || sub.st.n == 0) return kNoData;        // computeErrorCrossValidation over nothing (:323-324)
        if (rc) return kNoData;
        if (known[kh]) return table[kh];
        sub.s.topoDist_Kh = kh;
        rc = pb_compute_residuals_exact(ctx, &am.mp, am.val.data(), &am.cva, &sub.st, &sub.s, variable, demId, 1, 1, res.data());
        if (rc) return kNoData;
        sum.kernel_ms += ctx->timing.kernel_ms; sum.total_ms += ctx->timing.total_ms; sum.launches += ctx->timing.launches;
        known[kh] = 1;
        table[kh] = cvMaeHost(res.data(), am.val.data(), m);
        return table[kh];
    };
    pb_kh_series series{};
    const double best = goldenSection(f, maxKh, 0., double(maxKh), &series);
    if (rc) return rc;
    sub.s.topoDist_Kh = int(best);
    return PB_OK;
}

// glocalDetrendingFitting on the device: every area with stations is one unit of the batched preInterpolation kernel
int glocalFit(pb_context* ctx, const pb_stations* st, pb_settings* s, int variable, pb_macro_area* areas, int nAreas)
{
    if (!varUsesDetrending(variable) || !s->useMultipleDetrending) return PB_OK;      // preInterpolation :2464-2467
    if (s->useLocalDetrending) return fail(ctx, PB_ERR_INVALID, "glocal and local detrending are exclusive (project.cpp:3546-3552)");
    if (!s->hasPointsRange) return fail(ctx, PB_ERR_FIT, "couldn't set temperature ranges for height proxy (points range not set)");
    pb_settings s0 = *s;
    s0.useGlocalDetrending = 0;
    LocalModel m;
    int rc = buildPreModel(ctx, &s0, variable, st->n, m);
    if (rc) return rc;
    // the subsets, in each area's order (first match only, :2259-2266)
    std::vector<int> idx, off(nAreas + 1, 0), unitArea;
    int cap = 1;
    for (int a = 0; a < nAreas; a++) {
        const pb_macro_area& A = areas[a];
        if (A.nMeteoPoints <= 0) continue;
        if (!A.meteoPoints) return fail(ctx, PB_ERR_INVALID, "macro area without meteoPoints");
        const int begin = int(idx.size());
        for (int l = 0; l < A.nMeteoPoints; l++)
            for (int k = 0; k < st->n; k++)
                if ((st->index ? st->index[k] : k) == A.meteoPoints[l]) { idx.push_back(k); break; }
        unitArea.push_back(a);
        off[unitArea.size()] = int(idx.size());
        cap = std::max(cap, int(idx.size()) - begin);
    }
    const int T = int(unitArea.size());
    // the state glocalDetrendingFitting starts every area from, kept when no area has stations
    for (int i = 0; i < s->nProxy; i++) { s->currentActive[i] = s->selectedActive[i]; }
    if (T == 0) {
        // no area has stations: only setMultipleDetrendingHeightTemperatureRange acted (:2466-2469): height ranges, first
        // guesses and one more fitting function
        pb_settings o;
        PreFit none{};
        none.elevR2 = kNoData;
        applyPreFitToSettings(s0, m, none, o);
        for (int i = 0; i < s->nProxy; i++)
            if (i == m.elevationPos && s->selectedActive[i]) {
                const float r2 = s->proxy[i].regressionR2;
                s->proxy[i] = o.proxy[i];
                s->proxy[i].regressionR2 = r2;
            }
        secondHeightRangeCall(*s);
        return PB_OK;
    }
    LocalStations ds{};
    rc = stageAllStations(ctx, st, ds);
    if (rc) return rc;
    const int groupLanes = m.nFirstGuess <= 16 ? 16 : 32;
    LocalScratch scr{};
    scr.cap = cap;
    scr.perGroup = preScratchBytesPerGroup(scr.cap);
    const size_t groups = size_t(preGroupsTotal(groupLanes, T));
    const size_t offModel = 0, offFits = (sizeof(LocalModel) + 15) / 16 * 16, offIdx = (offFits + sizeof(PreFit) * T + 15) / 16 * 16,
                 offOff = offIdx + (idx.size() * 4 + 15) / 16 * 16, offScr = (offOff + size_t(T + 1) * 4 + 255) / 256 * 256,
                 total = offScr + scr.perGroup * groups;
    CUDA_TRY(ctx, ctx->dLocalScratch.reserve(total));
    unsigned char* d = ctx->dLocalScratch.p;
    CUDA_TRY(ctx, cudaMemcpyAsync(d + offModel, &m, sizeof(LocalModel), cudaMemcpyHostToDevice, ctx->stream));
    if (!idx.empty()) CUDA_TRY(ctx, cudaMemcpyAsync(d + offIdx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d + offOff, off.data(), size_t(T + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    ctx->timing.h2d_bytes += (int64_t)(sizeof(LocalModel) + idx.size() * 4 + size_t(T + 1) * 4);
    scr.base = d + offScr;
    scr.error = nullptr;
    PreSubsets sub{};
    sub.idx = reinterpret_cast<const int*>(d + offIdx);
    sub.off = reinterpret_cast<const int*>(d + offOff);
    sub.unweightedElevation = 1;
    CUDA_TRY(ctx, launchPreInterpolationBatch(reinterpret_cast<const LocalModel*>(d + offModel), ds, ds.value, T, scr, groupLanes,
                                              m.elevFunction, nullptr, nullptr, reinterpret_cast<PreFit*>(d + offFits), ctx->stream,
                                              &sub));
    ctx->timing.launches++;
    std::vector<PreFit> fits(T);
    CUDA_TRY(ctx, cudaMemcpyAsync(fits.data(), d + offFits, sizeof(PreFit) * T, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->timing.d2h_bytes += (int64_t)(sizeof(PreFit) * T);
    pb_settings o = *s;
    for (int u = 0; u < T; u++) {
        applyPreFitToSettings(s0, m, fits[u], o);
        pb_macro_area& A = areas[unitArea[u]];
        memset(A.active, 0, sizeof(A.active));
        memset(A.significant, 0, sizeof(A.significant));
        for (int p = 0; p < o.nProxy; p++) { A.active[p] = o.currentActive[p]; A.significant[p] = o.currentSignificant[p]; }
        A.useThermalInversion = o.currentUseThermalInversion;
        A.nParameters = o.nFit;
        for (int p = 0; p < PB_MAX_PROXY; p++) {
            A.nrParams[p] = p < o.nFit ? o.fitNrParams[p] : 0;
            for (int j = 0; j < PB_MAX_FIT_PARAMS; j++) A.parameters[p][j] = (p < o.nFit && j < o.fitNrParams[p]) ? o.fitParams[p][j] : 0.;
        }
    }
    o.useGlocalDetrending = s->useGlocalDetrending;      // the settings keep the last fitted area's state
    *s = o;
    return PB_OK;
}

}  // namespace

extern "C" {

int pb_glocal_detrending_fitting(pb_context* ctx, const pb_stations* st, pb_settings* s, int variable, pb_macro_area* areas,
                                 int nAreas)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (!st || !s || (!areas && nAreas > 0) || nAreas < 0) return fail(ctx, PB_ERR_INVALID, "null argument");
    cudaSetDevice(ctx->device);
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    return glocalFit(ctx, st, s, variable, areas, nAreas);
}

static int glocalRasterImpl(pb_context* ctx, const pb_stations* st, pb_settings* s, int variable, pb_macro_area* areas, int nAreas,
                            const ResidentMacroAreas* resident, pb_raster_id demId, const pb_raster_id* proxyGrids, int nProxyGrids,
                            int deviceOnly, pb_raster_out* out, const pb_glocal_meteo* gm = nullptr, int32_t* areaKh = nullptr)
{
    if (!st || !s || !out || (!areas && nAreas > 0) || nAreas < 0) return fail(ctx, PB_ERR_INVALID, "null argument");
    cudaSetDevice(ctx->device);
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    if (!varUsesDetrending(variable) || !s->useGlocalDetrending)       // project.cpp:3264
        return fail(ctx, PB_ERR_INVALID, "glocal detrending needs useGlocalDetrending and a detrending variable");
    const bool shepardMethod = s->method != PB_METHOD_IDW;      // K8 point form per area (one staging copy and sync per area)
    if (demId < 0 || demId >= (int)ctx->rasters.size() || !ctx->rasters[demId].used)
        return fail(ctx, PB_ERR_INVALID, "dem id is not a live raster");
    if (!deviceOnly && !out->data && !out->rows) return fail(ctx, PB_ERR_INVALID, "output has neither data nor rows");
    if (resident && int(resident->firstFloat.size()) != nAreas + 1) return fail(ctx, PB_ERR_INVALID, "resident cells describe other areas");
    // topographic distance: macroAreaDetrending searches kh per area over the area's METEO points (:1811-1826)
    const bool useTD = s->useTD && varUsesTd(variable);
    if (useTD && (!gm || !gm->meteo))
        return fail(ctx, PB_ERR_INVALID, "topographic distance inside macro areas needs the meteo points (pb_glocal_detrending_raster_td)");
    cudaEventRecord(ctx->ev[0], ctx->stream);
    for (int i = 0; i < s->nProxy; i++) { s->currentActive[i] = s->selectedActive[i]; s->currentSignificant[i] = 0; }
    s->currentUseThermalInversion = s->selectedUseThermalInversion;
    int rc = glocalFit(ctx, st, s, variable, areas, nAreas);     // preInterpolation (:3283)
    if (rc) return rc;
    secondHeightRangeCall(*s);                                   // :3300
    cudaEventRecord(ctx->ev[1], ctx->stream);

    RasterEntry& dem = ctx->rasters[demId];
    rc = ensureValidList(ctx, dem);
    if (rc) return rc;
    const DevGrid dg = devGridOfRaster(dem);
    long long maxCells = 1;
    for (int a = 0; a < nAreas; a++) {
        if (areas[a].nAreaCells > 0 && !areas[a].areaCellsDEM && !resident) return fail(ctx, PB_ERR_INVALID, "macro area without areaCellsDEM");
        if (areas[a].nAreaCells / 2 > 0x7fffff00LL) return fail(ctx, PB_ERR_INVALID, "macro area with too many cells");
        if (resident && resident->firstFloat[a + 1] - resident->firstFloat[a] != areas[a].nAreaCells)
            return fail(ctx, PB_ERR_INVALID, "resident cells describe other areas");
        maxCells = std::max<long long>(maxCells, areas[a].nAreaCells / 2);
    }
    // every area's stations (macroAreaDetrending on the host), packed behind each other in ONE staging buffer: one copy, no
    // synchronisation between the areas
    std::vector<AreaStations> subs(nAreas);
    std::vector<size_t> stOff(nAreas + 1, 0);
    pb_timing khTiming{};
    for (int a = 0; a < nAreas; a++) {
        if (areas[a].nAreaCells / 2 > 0) {
            macroAreaDetrendingHost(areas[a], *s, st, subs[a]);
            if (useTD) {
                // (the searches run before anything of the raster is staged: they go through the shared scratch buffers)
                AreaMeteo am;
                rc = areaMeteoPoints(ctx, areas[a], gm->meteo, gm->observed ? gm->observed : gm->meteo->value, gm->cv, am);
                if (!rc) rc = areaKhSearch(ctx, am, subs[a], variable, demId, khTiming);
                if (rc) return rc;
            }
        }
        if (areaKh) areaKh[a] = (useTD && areas[a].nAreaCells / 2 > 0) ? subs[a].s.topoDist_Kh : 0;
        stOff[a + 1] = stOff[a] + (areas[a].nAreaCells / 2 > 0 ? idwStationsStagingBytes(subs[a].st.n) : 0);
    }
    if (useTD) { memset(&ctx->timing, 0, sizeof(ctx->timing)); ctx->timing.launches = khTiming.launches; }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));           // nothing in flight may still read the staging buffer
    CUDA_TRY(ctx, ctx->hStations.reserve(stOff[nAreas] + 256));
    CUDA_TRY(ctx, ctx->dStations.reserve(stOff[nAreas] + 256));
    std::vector<DevStations> dss(nAreas);
    std::vector<DevModel> mods(nAreas);
    for (int a = 0; a < nAreas; a++) {
        if (areas[a].nAreaCells / 2 <= 0) continue;
        rc = buildDevModelGrids(ctx, &subs[a].s, variable, proxyGrids, nProxyGrids, mods[a]);
        if (rc) return rc;
        if (!shepardMethod)
            packIdwStationsAt(&subs[a].st, &subs[a].s, true, ctx->hStations.p + stOff[a], ctx->dStations.p + stOff[a], dss[a],
                              mods[a].valueOffset);
    }
    if (stOff[nAreas] > 0 && !shepardMethod)
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dStations.p, ctx->hStations.p, stOff[nAreas], cudaMemcpyHostToDevice, ctx->stream));
    ctx->timing.h2d_bytes += (int64_t)stOff[nAreas];

    const int P = std::max(s->nProxy, 0);
    const size_t mAl = (size_t(maxCells) + 3) / 4 * 4;
    const size_t offCells = 0, offX = offCells + (resident ? 0 : mAl * 8), offY = offX + mAl * 4, offZ = offY + mAl * 4,
                 offRes = offZ + mAl * 4, offPv = offRes + mAl * 4, offState = offPv + mAl * 8 * std::max(P, 1), offFlag = (offState + mAl + 15) / 16 * 16,
                 total = offFlag + 16;
    CUDA_TRY(ctx, ctx->dScratch.reserve(total));
    unsigned char* d = ctx->dScratch.p;
    int* dNotOk = reinterpret_cast<int*>(d + offFlag);
    CUDA_TRY(ctx, cudaMemsetAsync(dNotOk, 0, sizeof(int), ctx->stream));
    CUDA_TRY(ctx, ctx->dOut.reserve(dem.n));
    CUDA_TRY(ctx, launchFill(ctx->dOut.p, (long long)dem.n, dem.h.flag, ctx->stream));      // initializeGrid(header) :3297
    ctx->outInitRaster = -1;
    ctx->timing.launches++;

    for (int a = 0; a < nAreas; a++) {
        const pb_macro_area& A = areas[a];
        const int nCells = int(A.nAreaCells / 2);
        if (nCells <= 0) continue;
        GlocalLookup lk{};
        lk.nProxy = P;
        for (int i = 0; i < PB_MAX_PROXY; i++) {
            lk.slot[i] = (i < P && subs[a].s.currentActive[i] && subs[a].s.currentSignificant[i]) ? mods[a].proxy[i].gridSlot : -1;
            lk.grid[i] = mods[a].grid[i];
        }
        const float* dCells;
        if (resident) dCells = resident->cells + resident->firstFloat[a];
        else {
            CUDA_TRY(ctx, cudaMemcpyAsync(d + offCells, A.areaCellsDEM, size_t(nCells) * 8, cudaMemcpyHostToDevice, ctx->stream));
            ctx->timing.h2d_bytes += (int64_t)(size_t(nCells) * 8);
            dCells = reinterpret_cast<const float*>(d + offCells);
        }
        const int grid = (nCells + 255) / 256;
        static const bool noFuse = getenv("PB_GLOCAL_NO_FUSE") != nullptr;
        if (!shepardMethod && !(useTD && subs[a].s.topoDist_Kh != 0) && !noFuse) {
            glocal_area_kernel<<<(nCells + 511) / 512, 256, 0, ctx->stream>>>(dCells, nCells, dg, lk, dss[a], mods[a], ctx->dOut.p, dNotOk);
            CUDA_TRY(ctx, cudaGetLastError());
            ctx->timing.launches++;
            continue;
        }
        glocal_targets_kernel<<<grid, 256, 0, ctx->stream>>>(dCells, nCells, dg, lk, reinterpret_cast<float*>(d + offX),
                                                             reinterpret_cast<float*>(d + offY), reinterpret_cast<float*>(d + offZ),
                                                             reinterpret_cast<double*>(d + offPv), d + offState);
        CUDA_TRY(ctx, cudaGetLastError());
        if (shepardMethod) {
            // shepardIdw / modifiedShepardIdw over the area's points: neighbourhood radius from the settings' bounding box and
            // the AREA's point count (computeShepardInitialRadius :800-803 inside shepardSearchNeighbour :812)
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));           // the station staging buffer is reused per area
            ShepardStations ss{};
            float R0 = 0.f;
            rc = stageShepardStationsFor(ctx, &subs[a].st, &subs[a].s, true, ss, R0);
            if (rc) return rc;
            ShepardTd std_{};
            std_.kh = useTD ? subs[a].s.topoDist_Kh : 0;
            std_.dem = dg;
            std_.tz = reinterpret_cast<const float*>(d + offZ);
            if (std_.kh != 0) {
                rc = stageTdMaps(ctx, &subs[a].st, &subs[a].s, true, &std_.maps);
                if (rc) return rc;
            }
            CUDA_TRY(ctx, launchShepardPoints(ss, mods[a], R0, s->method == PB_METHOD_SHEPARD_MODIFIED,
                                              reinterpret_cast<const float*>(d + offX), reinterpret_cast<const float*>(d + offY),
                                              reinterpret_cast<const double*>(d + offPv), nCells,
                                              reinterpret_cast<float*>(d + offRes), ctx->stream, &std_));
        } else if (useTD && subs[a].s.topoDist_Kh != 0) {
            const DevGrid* tdMaps = nullptr;        // (restaged per area: the copy is ordered behind the previous area's kernel)
            rc = stageTdMaps(ctx, &subs[a].st, &subs[a].s, true, &tdMaps);
            if (rc) return rc;
            CUDA_TRY(ctx, launchIdwTdPoints(dss[a], dss[a].z, mods[a], dg, subs[a].s.topoDist_Kh,
                                            reinterpret_cast<const float*>(d + offX), reinterpret_cast<const float*>(d + offY),
                                            reinterpret_cast<const float*>(d + offZ), reinterpret_cast<const double*>(d + offPv),
                                            nCells, reinterpret_cast<float*>(d + offRes), ctx->stream, tdMaps));
        } else
        CUDA_TRY(ctx, launchIdwPoints(dss[a], mods[a], reinterpret_cast<const float*>(d + offX), reinterpret_cast<const float*>(d + offY),
                                      reinterpret_cast<const double*>(d + offPv), nCells, reinterpret_cast<float*>(d + offRes),
                                      ctx->stream));
        glocal_blend_kernel<<<grid, 256, 0, ctx->stream>>>(dCells, nCells, dem.h.nrCols, d + offState,
                                                           reinterpret_cast<const float*>(d + offRes), ctx->dOut.p, dNotOk);
        CUDA_TRY(ctx, cudaGetLastError());
        ctx->timing.launches += 3;
    }
    ctx->hMinMax[0] = 0xffffffffu;
    ctx->hMinMax[1] = 0u;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dMinMax, ctx->hMinMax, 2 * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
    {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
        const int grid = int(std::min<long long>((long long)sms * 8, ((long long)dem.n + 255) / 256));
        glocal_minmax_kernel<<<std::max(grid, 1), 256, 0, ctx->stream>>>(ctx->dOut.p, (long long)dem.n, dem.h.flag, ctx->dMinMax);
        CUDA_TRY(ctx, cudaGetLastError());
        ctx->timing.launches++;
    }
    cudaEventRecord(ctx->ev[2], ctx->stream);
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hMinMax, ctx->dMinMax, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hMinMax + 2, dNotOk, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (!deviceOnly) {
        const int cols = dem.h.nrCols, rows = dem.h.nrRows;
        if (out->data) {
            CUDA_TRY(ctx, cudaMemcpyAsync(out->data, ctx->dOut.p, dem.n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        } else {
            for (int row = 0; row < rows; row++)
                CUDA_TRY(ctx, cudaMemcpyAsync(out->rows[row], ctx->dOut.p + size_t(row) * cols, size_t(cols) * sizeof(float),
                                              cudaMemcpyDeviceToHost, ctx->stream));
        }
        ctx->timing.d2h_bytes += (int64_t)(dem.n * sizeof(float));
    }
    cudaEventRecord(ctx->ev[3], ctx->stream);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    const int notOk = int(ctx->hMinMax[2]);
    cudaEventElapsedTime(&ctx->timing.h2d_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->timing.kernel_ms, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->timing.d2h_ms, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&ctx->timing.total_ms, ctx->ev[0], ctx->ev[3]);
    out->nValid = dem.nValid;
    if (ctx->hMinMax[0] == 0xffffffffu) {
        out->minimum = kNoData;
        out->maximum = kNoData;
    } else {
        out->minimum = keyToFloatG(ctx->hMinMax[0]);
        out->maximum = keyToFloatG(ctx->hMinMax[1]);
    }
    if (notOk) return fail(ctx, PB_ERR_NO_VALID_CELL, "interpolate() returned NODATA inside a macro area (project.cpp:3358-3359)");
    return PB_OK;
}

int pb_glocal_detrending_raster(pb_context* ctx, const pb_stations* st, pb_settings* s, int variable, pb_macro_area* areas,
                                int nAreas, pb_raster_id demId, const pb_raster_id* proxyGrids, int nProxyGrids, int deviceOnly,
                                pb_raster_out* out)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    return glocalRasterImpl(ctx, st, s, variable, areas, nAreas, nullptr, demId, proxyGrids, nProxyGrids, deviceOnly, out);
}

int pb_glocal_detrending_raster_resident(pb_context* ctx, const pb_stations* st, pb_settings* s, int variable, pb_macro_area* areas,
                                         int nAreas, pb_macro_areas_id cells, pb_raster_id demId, const pb_raster_id* proxyGrids,
                                         int nProxyGrids, int deviceOnly, pb_raster_out* out)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (cells < 0 || cells >= (int)ctx->macroAreas.size() || !ctx->macroAreas[cells].used)
        return fail(ctx, PB_ERR_INVALID, "macro areas id is not live");
    return glocalRasterImpl(ctx, st, s, variable, areas, nAreas, &ctx->macroAreas[cells], demId, proxyGrids, nProxyGrids, deviceOnly,
                            out);
}

/* the same with topographic distance: the per-area kh search needs the meteo points; cells < 0 = the areas' host cells */
int pb_glocal_detrending_raster_td(pb_context* ctx, const pb_stations* st, pb_settings* s, int variable, pb_macro_area* areas,
                                   int nAreas, pb_macro_areas_id cells, const pb_glocal_meteo* meteo, pb_raster_id demId,
                                   const pb_raster_id* proxyGrids, int nProxyGrids, int deviceOnly, pb_raster_out* out,
                                   int32_t* areaKh)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    const ResidentMacroAreas* resident = nullptr;
    if (cells >= 0) {
        if (cells >= (int)ctx->macroAreas.size() || !ctx->macroAreas[cells].used)
            return fail(ctx, PB_ERR_INVALID, "macro areas id is not live");
        resident = &ctx->macroAreas[cells];
    }
    return glocalRasterImpl(ctx, st, s, variable, areas, nAreas, resident, demId, proxyGrids, nProxyGrids, deviceOnly, out, meteo,
                            areaKh);
}

int pb_macro_areas_upload(pb_context* ctx, const pb_macro_area* areas, int nAreas, pb_macro_areas_id* id)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (!areas || nAreas <= 0 || !id) return fail(ctx, PB_ERR_INVALID, "null argument");
    cudaSetDevice(ctx->device);
    ResidentMacroAreas r;
    r.firstFloat.assign(nAreas + 1, 0);
    for (int a = 0; a < nAreas; a++) {
        if (areas[a].nAreaCells < 0 || (areas[a].nAreaCells > 0 && !areas[a].areaCellsDEM))
            return fail(ctx, PB_ERR_INVALID, "macro area without areaCellsDEM");
        r.firstFloat[a + 1] = r.firstFloat[a] + areas[a].nAreaCells;
    }
    CUDA_TRY(ctx, cudaMalloc(&r.cells, std::max<size_t>(size_t(r.firstFloat[nAreas]), 1) * sizeof(float)));
    for (int a = 0; a < nAreas; a++)
        if (areas[a].nAreaCells > 0)
            CUDA_TRY(ctx, cudaMemcpyAsync(r.cells + r.firstFloat[a], areas[a].areaCellsDEM, size_t(areas[a].nAreaCells) * sizeof(float),
                                          cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    r.used = true;
    int slot = -1;
    for (size_t i = 0; i < ctx->macroAreas.size(); i++) if (!ctx->macroAreas[i].used) { slot = int(i); break; }
    if (slot < 0) { ctx->macroAreas.emplace_back(); slot = int(ctx->macroAreas.size()) - 1; }
    ctx->macroAreas[slot] = r;
    *id = slot;
    return PB_OK;
}

int pb_macro_areas_free(pb_context* ctx, pb_macro_areas_id id)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (id < 0 || id >= (int)ctx->macroAreas.size() || !ctx->macroAreas[id].used) return fail(ctx, PB_ERR_INVALID, "bad macro areas id");
    cudaSetDevice(ctx->device);
    cudaFree(ctx->macroAreas[id].cells);
    ctx->macroAreas[id] = ResidentMacroAreas();
    return PB_OK;
}

static int residualsGlocalImpl(pb_context* ctx, const pb_stations* meteo, const float* observed, const pb_cv_points* cv,
                               const pb_stations* points, pb_settings* s, int variable, const pb_macro_area* areas, int nAreas,
                               pb_raster_id demId, int excludeOutsideDem, int excludeSupplemental, float* residual)
{
    if (!meteo || !points || !s || !residual || (!areas && nAreas > 0)) return fail(ctx, PB_ERR_INVALID, "null argument");
    cudaSetDevice(ctx->device);
    if (!s->hasPointsRange || !s->useMultipleDetrending)        // setMultipleDetrendingHeightTemperatureRange false (project.cpp:3069)
        return fail(ctx, PB_ERR_FIT, "couldn't set temperature ranges for height proxy");
    const bool useTD = s->useTD && varUsesTd(variable);
    if (useTD && (demId < 0 || demId >= (int)ctx->rasters.size() || !ctx->rasters[demId].used))
        return fail(ctx, PB_ERR_INVALID, "topographic distance inside macro areas needs the resident DEM (pb_compute_residuals_glocal_td)");
    for (int i = 0; i < meteo->n; i++) residual[i] = kNoData;
    const float* obs = observed ? observed : meteo->value;
    pb_timing sum{};
    for (int k = 0; k < nAreas; k++) {
        const pb_macro_area& A = areas[k];
        if (A.nAreaCells <= 0 || A.nMeteoPoints <= 0) continue;
        AreaStations sub;
        macroAreaDetrendingHost(A, *s, points, sub);
        AreaMeteo am;
        int rc = areaMeteoPoints(ctx, A, meteo, obs, cv, am);
        if (rc) return rc;
        if (useTD) {            // macroAreaDetrending ends with the area's kh search (:1811-1826)
            rc = areaKhSearch(ctx, am, sub, variable, demId, sum);
            if (rc) return rc;
        }
        if (am.who.empty()) continue;
        const size_t m = am.who.size();
        std::vector<float> res(m);
        rc = pb_compute_residuals_exact(ctx, &am.mp, am.val.data(), &am.cva, &sub.st, &sub.s, variable, useTD ? demId : -1,
                                        excludeOutsideDem, excludeSupplemental, res.data());
        if (rc) return rc;
        sum.kernel_ms += ctx->timing.kernel_ms; sum.total_ms += ctx->timing.total_ms; sum.launches += ctx->timing.launches;
        for (size_t q = 0; q < m; q++) {
            if (isEqualF(res[q], kNoData)) continue;
            float& r = residual[am.who[q]];
            if (isEqualF(r, kNoData)) r = res[q] * 1.f;
            else r += res[q] * 1.f;
        }
    }
    ctx->timing = sum;
    return PB_OK;
}

int pb_compute_residuals_glocal(pb_context* ctx, const pb_stations* meteo, const float* observed, const pb_cv_points* cv,
                                const pb_stations* points, pb_settings* s, int variable, const pb_macro_area* areas, int nAreas,
                                int excludeOutsideDem, int excludeSupplemental, float* residual)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    return residualsGlocalImpl(ctx, meteo, observed, cv, points, s, variable, areas, nAreas, -1, excludeOutsideDem,
                               excludeSupplemental, residual);
}

int pb_compute_residuals_glocal_td(pb_context* ctx, const pb_stations* meteo, const float* observed, const pb_cv_points* cv,
                                   const pb_stations* points, pb_settings* s, int variable, const pb_macro_area* areas, int nAreas,
                                   pb_raster_id dem, int excludeOutsideDem, int excludeSupplemental, float* residual)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    return residualsGlocalImpl(ctx, meteo, observed, cv, points, s, variable, areas, nAreas, dem, excludeOutsideDem,
                               excludeSupplemental, residual);
}

}  // extern "C"
