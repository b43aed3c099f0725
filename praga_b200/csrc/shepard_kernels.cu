// shepard_kernels.cu — K8: the two Shepard estimators interpolate() offers besides plain IDW
// (agrolib/interpolation/interpolation.cpp, paths relative to the reference tree):
//   shepardSearchNeighbour   :806-868   first neighbourhood inside computeShepardInitialRadius (:800-803), then
//                                        < 5 points: the 5 nearest of ALL points; > 10: the 10 nearest of the neighbourhood;
//                                        otherwise the neighbourhood itself (sortPointsByDistance :121-160)
//   shepardIdw               :871-945   Shepard 1968 weights with the direction term
//   modifiedShepardIdw       :948-1028  s = (R - d) / (R d); its DEFINITION takes (..., radius, y, x) while every caller passes
//                                        (..., radius, x, y): inside, the direction cosine mixes the target's y with the
//                                        stations' x. Reproduced as is (it is the demo's default algorithm).
// One thread per target. Pass 1 walks all stations once (warp-uniform addresses -> broadcast loads) keeping the count inside
// the initial radius, the 10 nearest inside it and the 5 nearest overall; exact f32 distances (IEEE sqrt) are only taken for
// pairs a squared-distance test cannot rule out. Pass 2 is the O(k^2), k <= 10, weight computation in the reference's
// float/double mix (this TU is built with -fmad=false). Ties of equal distances are broken by station index (the
// reference's std::sort leaves them unspecified).
#include "epilogue.cuh"
#include "shepard.cuh"
#include "td_device.cuh"

namespace pb {

namespace {

constexpr int kMinPts = 5, kMaxPts = 10;       // SHEPARD_MIN_NRPOINTS / SHEPARD_MAX_NRPOINTS, interpolationConstants.h:6-8

struct Near {
    float d;
    int i;
};

// keep the K smallest (d, index) pairs, ascending
template <int K>
__device__ __forceinline__ void insertNear(Near (&a)[K], int& n, float d, int i)
{
    if (n == K && !(d < a[K - 1].d)) return;      // equal distance, larger index: stays out
    int pos = n < K ? n : K - 1;
    while (pos > 0 && d < a[pos - 1].d) {
        a[pos] = a[pos - 1];
        pos--;
    }
    a[pos].d = d;
    a[pos].i = i;
    if (n < K) n++;
}

__device__ __forceinline__ unsigned orderedKeyS(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

template <bool POINTS, bool MODIFIED, bool TD>
__global__ void __launch_bounds__(128)
shepard_kernel(ShepardStations st, DevModel m, DevGrid dem, float R0, const int* __restrict__ validIdx, int nTargets,
               const float* __restrict__ tx, const float* __restrict__ ty, const double* __restrict__ tproxy,
               float* __restrict__ out, unsigned* __restrict__ minmax, ShepardTd td)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    float vmin = INFINITY, vmax = -INFINITY;
    if (t < nTargets) {
        float x, y, z = 0.f;
        int cell;
        if (POINTS) { cell = t; x = tx[t]; y = ty[t]; if (TD) z = td.tz[t]; }
        else {
            cell = validIdx[t];
            const int row = cell / dem.nrCols, col = cell - row * dem.nrCols;
            // getUtmXYFromRowColSinglePrecision(header, ...), gis.cpp:799-804
            x = float(dem.llx + dem.cellSize * double(col) + 0.5);
            y = float(dem.lly + dem.cellSize * double(dem.nrRows - row) - 0.5);
            if (TD) z = dem.data[cell];          // interpolate(..., myX, myY, myZ = the DEM value, ...) (interpolationCmd.cpp:379-388)
        }

        float result = kNoData;
        if (m.useRetrendOnly) result = 0.f;
        else {
            // ---- pass 1: shepardSearchNeighbour ----
            Near in10[kMaxPts], all5[kMinPts];
            int n10 = 0, n5 = 0, cnt = 0;
            const float R0sq = R0 * R0 * 1.000001f;
            for (int i = 0; i < st.n; i++) {
                const float dx = __ldg(st.x + i) - x, dy = __ldg(st.y + i) - y;
                const float d2 = dx * dx + dy * dy;
                const float w5 = n5 == kMinPts ? all5[kMinPts - 1].d : INFINITY;
                // (with topographic distance the total distance is >= the planar one for kh >= 0: the planar test still rules out)
                if ((!TD || td.kh >= 0) && d2 > R0sq && d2 > w5 * w5 * 1.000001f) continue;
                float d = sqrtf(d2);
                if (TD) {
                    // computeDistances :226-251: distance += kh * topographicDistance(x, y, z, point x, y, z, distance, DEM)
                    const float topo = topographicDistanceCachedDev(td.maps, i, td.dem, x, y, z, __ldg(st.x + i), __ldg(st.y + i), __ldg(st.z + i), d);
                    d += (td.kh * topo);
                }
                if (!(fabs(double(d)) < kEpsilon)) insertNear(all5, n5, d, i);       // sortPointsByDistance drops |d| < EPSILON
                if (d <= R0 && d > 0.f) {
                    cnt++;
                    if (!(fabs(double(d)) < kEpsilon)) insertNear(in10, n10, d, i);
                }
            }
            Near sel[kMaxPts];
            int k = 0;
            float radius = kNoData;
            if (cnt < kMinPts) {
                k = n5;
                for (int q = 0; q < k; q++) sel[q] = all5[q];
                if (k > 0) radius = all5[k - 1].d + float(kEpsilon);
            } else if (cnt > kMaxPts) {
                k = n10;
                for (int q = 0; q < k; q++) sel[q] = in10[q];
                radius = in10[k - 1].d + float(kEpsilon);
            } else {
                // the neighbourhood itself, in station order
                k = n10;
                for (int q = 0; q < k; q++) sel[q] = in10[q];
                for (int a = 1; a < k; a++) {
                    const Near v = sel[a];
                    int b = a;
                    while (b > 0 && sel[b - 1].i > v.i) { sel[b] = sel[b - 1]; b--; }
                    sel[b] = v;
                }
                radius = R0;
            }

            // ---- pass 2: weights ----
            if (k > 0) {
                double S[kMaxPts], T[kMaxPts], px[kMaxPts], py[kMaxPts];
                for (int q = 0; q < k; q++) { px[q] = st.xd[sel[q].i]; py[q] = st.yd[sel[q].i]; S[q] = 0.; T[q] = 0.; }
                double weightSum = 0.;
                if (!MODIFIED) {
                    const double radius_3 = double(radius) / 3.;
                    const double radius_27_4 = 6.75 / double(radius);
                    for (int q = 0; q < k; q++) {
                        const float d = sel[q].d;
                        if (d > 0.f) {
                            if (double(d) <= radius_3) S[q] = double(1.f / d);
                            else if (d <= radius) {
                                const double tmp = double((d / radius) - 1.f);
                                S[q] = radius_27_4 * tmp * tmp;
                            } else S[q] = 0.;
                            weightSum += S[q];
                        }
                    }
                    if (weightSum != 0.) {
                        const double X = double(x), Y = double(y);
                        for (int a = 0; a < k; a++) {
                            double tt = 0.;
                            for (int b = 0; b < k; b++)
                                if (a != b) {
                                    const double cosine = ((X - px[a]) * (X - px[b]) + (Y - py[a]) * (Y - py[b])) /
                                                          double(sel[a].d * sel[b].d);
                                    tt += S[b] * (1. - cosine);
                                }
                            tt /= weightSum;
                            T[a] = tt;
                        }
                        double wsum = 0.;
                        for (int a = 0; a < k; a++) {
                            T[a] = S[a] * S[a] * (1. + T[a]);      // weight[i]
                            wsum += T[a];
                        }
                        double r = 0.;
                        for (int a = 0; a < k; a++) r += (T[a] / wsum) * double(st.value[sel[a].i]);
                        result = float(r);
                    }
                } else {
                    for (int q = 0; q < k; q++) {
                        const float d = sel[q].d;
                        if (d > 0.f && d <= radius) {
                            S[q] = double((radius - d) / (radius * d));
                            weightSum += S[q];
                        }
                    }
                    if (weightSum != 0.) {
                        const double invWeightSum = 1.0 / weightSum;
                        // parameter order of the definition (:948-949): inside, "x" is the caller's y and "y" the caller's x
                        const double X = double(y), Y = double(x);
                        for (int a = 0; a < k; a++) {
                            if (S[a] == 0. || sel[a].d <= 0.f) continue;
                            double tt = 0.;
                            for (int b = 0; b < k; b++) {
                                if (a == b || S[b] == 0. || sel[b].d <= 0.f) continue;
                                const double cosine = ((X - px[a]) * (X - px[b]) + (Y - py[a]) * (Y - py[b])) /
                                                      double(sel[a].d * sel[b].d);
                                tt += S[b] * (1.0 - cosine);
                            }
                            T[a] = tt * invWeightSum;
                        }
                        double wsum = 0.;
                        for (int a = 0; a < k; a++) {
                            T[a] = S[a] * S[a] * (1.0 + T[a]);
                            wsum += T[a];
                        }
                        const double inv = 1.0 / wsum;
                        double r = 0.;
                        for (int a = 0; a < k; a++) r += (T[a] * inv) * double(st.value[sel[a].i]);
                        result = float(r);
                    }
                }
            }
        }

        double pv[PB_MAX_PROXY];
        if (POINTS) {
            for (int i = 0; i < m.nProxy; i++) pv[i] = tproxy[(size_t)cell * m.nProxy + i];
        } else {
            for (int i = 0; i < m.nProxy; i++) {
                pv[i] = double(kNoData);
                const DevProxy& p = m.proxy[i];
                if (m.useDetrendingVar && p.active && p.gridSlot >= 0) {
                    const DevGrid& g = m.grid[p.gridSlot];
                    const float v = gridValueFromXY(g, double(x), double(y));
                    if (v != g.flag) pv[i] = double(v);
                }
            }
        }
        const float o = finishEstimate(m, result, pv);
        out[cell] = o;
        if (!isEqualF(o, dem.flag) && !isEqualF(o, kNoData)) { vmin = o; vmax = o; }
    }
    if (minmax) {
        for (int o = 16; o; o >>= 1) {
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        }
        if ((threadIdx.x & 31) == 0) {
            if (vmin != INFINITY) atomicMin(&minmax[0], orderedKeyS(vmin));
            if (vmax != -INFINITY) atomicMax(&minmax[1], orderedKeyS(vmax));
        }
    }
}

}  // namespace

cudaError_t launchShepardRaster(const ShepardStations& st, const DevModel& m, const DevGrid& dem, float R0, bool modified,
                                const int* validIdx, int nTargets, float* out, unsigned* minmax, cudaStream_t s, const ShepardTd* td)
{
    if (nTargets <= 0) return cudaSuccess;
    const int grid = (nTargets + 127) / 128;
    const bool useTd = td && td->kh != 0;
    const ShepardTd t = useTd ? *td : ShepardTd{};
    if (useTd) {
        if (modified)
            shepard_kernel<false, true, true><<<grid, 128, 0, s>>>(st, m, dem, R0, validIdx, nTargets, nullptr, nullptr, nullptr, out, minmax, t);
        else
            shepard_kernel<false, false, true><<<grid, 128, 0, s>>>(st, m, dem, R0, validIdx, nTargets, nullptr, nullptr, nullptr, out, minmax, t);
    } else if (modified)
        shepard_kernel<false, true, false><<<grid, 128, 0, s>>>(st, m, dem, R0, validIdx, nTargets, nullptr, nullptr, nullptr, out, minmax, t);
    else
        shepard_kernel<false, false, false><<<grid, 128, 0, s>>>(st, m, dem, R0, validIdx, nTargets, nullptr, nullptr, nullptr, out, minmax, t);
    return cudaGetLastError();
}

cudaError_t launchShepardPoints(const ShepardStations& st, const DevModel& m, float R0, bool modified, const float* tx,
                                const float* ty, const double* tproxy, int nTargets, float* out, cudaStream_t s, const ShepardTd* td)
{
    if (nTargets <= 0) return cudaSuccess;
    const int grid = (nTargets + 127) / 128;
    DevGrid none{};
    none.flag = kNoData;
    const bool useTd = td && td->kh != 0;
    const ShepardTd t = useTd ? *td : ShepardTd{};
    if (useTd) {
        if (modified)
            shepard_kernel<true, true, true><<<grid, 128, 0, s>>>(st, m, none, R0, nullptr, nTargets, tx, ty, tproxy, out, nullptr, t);
        else
            shepard_kernel<true, false, true><<<grid, 128, 0, s>>>(st, m, none, R0, nullptr, nTargets, tx, ty, tproxy, out, nullptr, t);
    } else if (modified)
        shepard_kernel<true, true, false><<<grid, 128, 0, s>>>(st, m, none, R0, nullptr, nTargets, tx, ty, tproxy, out, nullptr, t);
    else
        shepard_kernel<true, false, false><<<grid, 128, 0, s>>>(st, m, none, R0, nullptr, nTargets, tx, ty, tproxy, out, nullptr, t);
    return cudaGetLastError();
}

}  // namespace pb
