// detrend_device.cuh — device side of multiple detrending: the reference's Levenberg-Marquardt proxy fits, run by a
// GROUP of G lanes (16 or 32) per problem (a raster cell with its locally selected stations, or a whole time step).
//
// Reference (paths relative to agrolib/):
//   multipleDetrendingMain                interpolation/interpolation.cpp:1832-1859
//   multipleDetrendingElevationFitting    :1862-1953   (setFittingParameters_elevation :1625-1677)
//   detrendingElevation                   :1956-1972
//   multipleDetrendingOtherProxiesFitting :1975-2171   (setFittingParameters_otherProxies :1680-1728)
//   detrendingOtherProxies                :2173-2236
//   proxyValidity                         :1418-1457
//   bestFittingMarquardt_nDimension       mathFunctions/furtherMathFunctions.cpp:1360-1427 (elevation, weighted)
//   fittingMarquardt_nDimension           :1954-2118 (1-D weighted), :1740-1947 (multi-proxy weighted)
//   computeWeighted_R2 / _RMSE            :1245-1275, :1301-1322
//   lapseRatePiecewise_* / functionLinear_intercept / functionSum   :115-201
//
// Mapping: bestFittingMarquardt runs one INDEPENDENT Levenberg-Marquardt descent per first guess (16 for the
// two-piece lapse rate, 30/31 for the three-piece ones): lane g of the group runs first guess g, sequentially over
// the data points in the reference's order, in FP64, with every accumulator in registers. Nothing is reduced across
// lanes inside a fit, so each descent follows the reference's operation order exactly (this TU is compiled with
// -fmad=false: no FMA contraction, like the reference's x86-64 build) and takes the same accept/reject path.
// The winner is then picked with the reference's sequential rule (R2 > best + deltaR2) through warp shuffles.
#pragma once
#include "common.cuh"

namespace pb {

// ---- group helpers (G lanes per problem, a power of two: 32 = one problem per warp, 16 = two, 4 = eight) ----
template <int G> __device__ __forceinline__ unsigned grpShift() { return G == 32 ? 0u : ((threadIdx.x & 31u) & ~(unsigned(G) - 1u)); }
template <int G> __device__ __forceinline__ unsigned grpMask() { return G == 32 ? 0xffffffffu : (((1u << G) - 1u) << grpShift<G>()); }
template <int G> __device__ __forceinline__ int grpLane() { return threadIdx.x & (G - 1); }
template <int G> __device__ __forceinline__ unsigned grpBallot(bool p) { return __ballot_sync(grpMask<G>(), p) >> grpShift<G>(); }
template <int G> __device__ __forceinline__ double grpBcast(double v, int src) { return __shfl_sync(grpMask<G>(), v, src, G); }
template <int G> __device__ __forceinline__ void grpSync() { __syncwarp(grpMask<G>()); }

__device__ __forceinline__ bool isEqualD(double a, double b) { return fabs(a - b) < kEpsilon; }

// checkLapseRateCode, meteo/meteo.cpp:1127-1133
__device__ __forceinline__ bool checkLapseRateCodeDev(int code, bool useLapseRateCode, bool useSupplemental)
{
    if (useSupplemental) return !useLapseRateCode || code == PB_LAPSE_PRIMARY || code == PB_LAPSE_SUPPLEMENTAL;
    return !useLapseRateCode || code == PB_LAPSE_PRIMARY || code == PB_LAPSE_SECONDARY;
}

// n / d for a divisor whose correctly rounded reciprocal r = 1/d is known: two FMA correction steps on q = n*r give the
// correctly rounded IEEE quotient (Markstein), the same bits as the division instruction sequence at a sixth of its cost.
// Quotients outside the safe exponent range (zeros, subnormals, huge values, NaN) take the plain division.
__device__ __forceinline__ double divByKnown(double n, double d, double r)
{
    const double q0 = n * r;
    const double aq = fabs(q0);
    // the corrected quotient is computed unconditionally (straight-line code: the lanes of a group sit on different data
    // points of different descents, a branch here diverges on every zero derivative); the rare cases select afterwards
    const double e0 = __fma_rn(-d, q0, n);
    const double q1 = __fma_rn(e0, r, q0);
    const double e1 = __fma_rn(-d, q1, n);
    double q = __fma_rn(e1, r, q1);
    if (!(aq > 1e-280 && aq < 1e280)) {
        q = q0;                     // n == 0: (+-0) / d keeps the sign of n * sign(d)
        if (n != 0) q = n / d;      // subnormal / huge / NaN quotients: the division instruction sequence
    }
    return q;
}

// straight-line form of divByKnown for the inner loops: the corrected quotient is always computed, a zero numerator
// selects q0 = (+-0) (same value and sign as the division), and the quotients outside the safe exponent range only raise
// `slow`, so that the caller redoes them with divByKnown after the loop body. No branch => the independent columns of a
// Jacobian row interleave instead of being serialised by reconvergence barriers.
__device__ __forceinline__ double divFast(double n, double d, double r, bool& slow)
{
    const double q0 = n * r;
    const double e0 = __fma_rn(-d, q0, n);
    const double q1 = __fma_rn(e0, r, q0);
    const double e1 = __fma_rn(-d, q1, n);
    const double q = __fma_rn(e1, r, q1);
    // the guards are integer tests on the operands' bits (the FP64 pipe is what bounds the descents: no DSETP here): biased
    // exponent of q0 in [94, 1952] = |q0| within about (1e-280, 1e280); a zero numerator of either sign keeps q0
    const unsigned ex = (unsigned(__double2hiint(q0)) >> 20) & 0x7ffu;
    const bool nz = ((__double2hiint(n) & 0x7fffffff) | __double2loint(n)) != 0;
    slow |= (ex - 94u > 1858u) && nz;
    return nz ? q : q0;
}

// ---- model functions with the reference's in-place clamp of par[2] (furtherMathFunctions.cpp:138, 158) ----
template <int F> struct Fit;
template <> struct Fit<PB_FIT_PIECEWISE_TWO> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void clamp(double*) {}
    static __device__ __forceinline__ double eval(double x, const double* p)
    {
        const double slope = (x < p[0]) ? p[2] : p[3];      // same products and sums as the two returns, no branch
        return slope * (x - p[0]) + p[1];
    }
};
template <> struct Fit<PB_FIT_PIECEWISE_THREE> {
    static constexpr int N = 5;
    static __device__ __forceinline__ void clamp(double* p) { p[2] = (10. > p[2]) ? 10. : p[2]; }
    static __device__ __forceinline__ double eval(double x, const double* p)
    {
        const double xb = p[2] + p[0];
        if (x < p[0]) return p[4] * x - p[0] * p[4] + p[1];
        else if (x > xb) return p[4] * x - p[4] * p[0] - p[4] * p[2] + p[3] * p[2] + p[1];
        else return p[3] * x - p[3] * p[0] + p[1];
    }
};
template <> struct Fit<PB_FIT_PIECEWISE_THREE_FREE> {
    static constexpr int N = 6;
    static __device__ __forceinline__ void clamp(double* p) { p[2] = (10. > p[2]) ? 10. : p[2]; }
    static __device__ __forceinline__ double eval(double x, const double* p)
    {
        const double xb = p[0] + p[2];
        if (x < p[0]) return p[4] * x - p[0] * p[4] + p[1];
        else if (x > xb) return p[5] * x - p[5] * p[0] - p[5] * p[2] + p[3] * p[2] + p[1];
        else return p[3] * x - p[3] * p[0] + p[1];
    }
};

// run-time dispatched evaluation on a private copy of the parameters (detrend / retrend)
__device__ __forceinline__ double evalFitRt(int f, double x, const double* par)
{
    double p[PB_MAX_FIT_PARAMS];
#pragma unroll
    for (int j = 0; j < PB_MAX_FIT_PARAMS; j++) p[j] = par[j];
    switch (f) {
    case PB_FIT_PIECEWISE_TWO: return Fit<PB_FIT_PIECEWISE_TWO>::eval(x, p);
    case PB_FIT_PIECEWISE_THREE: Fit<PB_FIT_PIECEWISE_THREE>::clamp(p); return Fit<PB_FIT_PIECEWISE_THREE>::eval(x, p);
    case PB_FIT_PIECEWISE_THREE_FREE:
        Fit<PB_FIT_PIECEWISE_THREE_FREE>::clamp(p);
        return Fit<PB_FIT_PIECEWISE_THREE_FREE>::eval(x, p);
    case PB_FIT_LINEAR: return p[0] * x + p[1];
    default: return 0.;
    }
}

// ---------------------------------------------------------------------------------------------------------
// fittingMarquardt_nDimension, 1-D weighted (furtherMathFunctions.cpp:1954-2118): ONE descent, run by one lane.
// The reference builds P[k][j] = (f(x_j; p + delta_k) - f(x_j; p)) / delta_k for k outer, j inner, then a = P^T W P
// and g = P^T W (y - f). Here j is the outer loop and every accumulator lives in registers; each accumulator still
// receives its terms in ascending j, so the sums are bit-identical. The perturb/restore sequence of the reference
// (par[k] += delta; ...; par[k] -= delta, which does not always restore the same bits) is reproduced: column k is
// evaluated with the already "restored" entries below k, and the restored vector is what the step is added to.
// ---------------------------------------------------------------------------------------------------------
// unroll factor of the data loops of lmIter (each accumulator still receives its terms in ascending point order; the point
// computations of consecutive iterations overlap). A TU may define PB_LM_UNROLL before including this header.
#ifndef PB_LM_UNROLL
#define PB_LM_UNROLL 1
#endif
constexpr int kLmUnroll = PB_LM_UNROLL;

// state of one descent: lmElevation = lmInit + lmIter until it reports the end
template <int F>
struct LmState {
    double p[Fit<F>::N], lambda[Fit<F>::N];
    double mySSE;
    int iter;
    // SCORED form of lmIter only: the curve-dependent sums of computeWeighted_R2 / _RMSE (:1245-1275, :1301-1322) for the
    // CURRENT parameters p, i.e. ssr = sum (W (Y - f))^2 and sc = sum W (Y - f)^2, taken in the pass that evaluates the SSE of
    // a candidate (same f, same terms, same order as weightedScoresWith) and kept when the candidate is accepted
    double ssr, sc;
};

template <int F>
__device__ __forceinline__ void lmInit(LmState<F>& s, const double* par, const double* X, const double* Y, const double* W,
                                       int nData)
{
    constexpr int N = Fit<F>::N;
#pragma unroll
    for (int k = 0; k < N; k++) { s.lambda[k] = 0.01; s.p[k] = par[k]; }
    s.mySSE = 0;
    if (nData > 0) Fit<F>::clamp(s.p);
    for (int i = 0; i < nData; i++) {
        const double error = Y[i] - Fit<F>::eval(X[i], s.p);
        const double w = W[i];
        s.mySSE += error * error * w * w;
    }
    s.iter = 0;
}

// one data point of the forward-difference Jacobian: P[k] = (f(x; p + delta_k) - f(x; p)) / delta_k with the reference's
// perturb / restore sequence (see above); returns f(x; p). Shared by the lane-per-descent and the warp-per-descent forms.
template <int F>
__device__ __forceinline__ double lmPointJacobian(double x, const double* p0, const double* up, const double* rs, double sp2,
                                                  const double* dl, const double* rdl, double* P)
{
    constexpr int N = Fit<F>::N;
    const double f0 = Fit<F>::eval(x, p0);
    double num[N];
    bool slow = false;
#pragma unroll
    for (int k = 0; k < N; k++) {
        double q[N];
#pragma unroll
        for (int c = 0; c < N; c++) q[c] = (c < k) ? ((F != PB_FIT_PIECEWISE_TWO && c == 2) ? sp2 : rs[c]) : p0[c];
        q[k] = up[k];
        num[k] = Fit<F>::eval(x, q) - f0;
        P[k] = divFast(num[k], dl[k], rdl[k], slow);
    }
    if (slow) {
#pragma unroll
        for (int k = 0; k < N; k++) P[k] = divByKnown(num[k], dl[k], rdl[k]);
    }
    return f0;
}

// branch-free form for the lane-per-descent loops: a quotient outside the fast division's exponent range only raises `slow`
// (the caller redoes the whole pass over the data with lmPointJacobian). Without the fix-up branch the body of the data loop
// is one basic block, so the points of an unrolled loop interleave.
template <int F>
__device__ __forceinline__ double lmPointJacobianFast(double x, const double* p0, const double* up, const double* rs, double sp2,
                                                      const double* dl, const double* rdl, double* P, bool& slow)
{
    constexpr int N = Fit<F>::N;
    const double f0 = Fit<F>::eval(x, p0);
#pragma unroll
    for (int k = 0; k < N; k++) {
        double q[N];
#pragma unroll
        for (int c = 0; c < N; c++) q[c] = (c < k) ? ((F != PB_FIT_PIECEWISE_TWO && c == 2) ? sp2 : rs[c]) : p0[c];
        q[k] = up[k];
        P[k] = divFast(Fit<F>::eval(x, q) - f0, dl[k], rdl[k], slow);
    }
    return f0;
}

// one Levenberg-Marquardt iteration; true when the descent is over (the reference's do-while condition turned false)
template <int F, bool SCORED = false>
__device__ __forceinline__ bool lmIter(LmState<F>& s, const double* __restrict__ pmin, const double* __restrict__ pmax,
                                       const double* dl, const double* rdl, int maxIter, double eps, const double* X,
                                       const double* Y, const double* W, int nData)
{
    constexpr int N = Fit<F>::N;
    const double VFACTOR = 10;
    double diffSSE;
    // perturbed parameter vectors: up[k] = value of par[k] while column k is evaluated; rs[k] = after the restore
    double p0[N], up[N], rs[N];
#pragma unroll
    for (int k = 0; k < N; k++) p0[k] = s.p[k];
#pragma unroll
    for (int k = 0; k < N; k++) {
        s.p[k] += dl[k];
        if (nData > 0) Fit<F>::clamp(s.p);       // the model function clamps par[2] in place when it is evaluated
        up[k] = s.p[k];
        s.p[k] -= dl[k];
        rs[k] = s.p[k];
    }
    if (nData > 0) Fit<F>::clamp(s.p);           // evaluations of the following columns / of newPar see it clamped
    // for the three-piece forms s.p[2] may have been re-clamped after its own restore: columns k > 2 see s.p[2]
    double g[N], a[N][N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        g[i] = 0;
#pragma unroll
        for (int j = 0; j < N; j++) a[i][j] = 0;
    }
    bool anySlow = false;
#pragma unroll kLmUnroll
    for (int j = 0; j < nData; j++) {
        const double x = X[j], y = Y[j], w = W[j];
        double P[N], WP[N];
        const double f0 = lmPointJacobianFast<F>(x, p0, up, rs, s.p[2], dl, rdl, P, anySlow);
#pragma unroll
        for (int k = 0; k < N; k++) WP[k] = w * P[k];
#pragma unroll
        for (int i = 0; i < N; i++) {
#pragma unroll
            for (int c = i; c < N; c++) a[i][c] += (P[i] * WP[c]);
            g[i] += P[i] * w * (y - f0);
        }
    }
    if (anySlow) {
        // a quotient left the fast division's exponent range (subnormal / huge): the pass again with the plain division where
        // it is needed, same terms in the same order
#pragma unroll
        for (int i = 0; i < N; i++) {
            g[i] = 0;
#pragma unroll
            for (int j = 0; j < N; j++) a[i][j] = 0;
        }
#pragma unroll 1
        for (int j = 0; j < nData; j++) {
            const double x = X[j], y = Y[j], w = W[j];
            double P[N], WP[N];
            const double f0 = lmPointJacobian<F>(x, p0, up, rs, s.p[2], dl, rdl, P);
#pragma unroll
            for (int k = 0; k < N; k++) WP[k] = w * P[k];
#pragma unroll
            for (int i = 0; i < N; i++) {
#pragma unroll
                for (int c = i; c < N; c++) a[i][c] += (P[i] * WP[c]);
                g[i] += P[i] * w * (y - f0);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < N; k++) {
        a[k][k] += s.lambda[k] * a[k][k];
#pragma unroll
        for (int j = k + 1; j < N; j++) a[j][k] = a[k][j];
    }
#pragma unroll
    for (int j = 0; j < N - 1; j++) {
        const double pivot = (a[j][j] > kEpsilon) ? a[j][j] : kEpsilon;     // MAXVALUE(a[j][j], EPSILON)
#pragma unroll
        for (int i = j + 1; i < N; i++) {
            const double mult = a[i][j] / pivot;
#pragma unroll
            for (int k = j + 1; k < N; k++) a[i][k] -= mult * a[j][k];
            g[i] -= mult * g[j];
        }
    }
    double change[N];
    change[N - 1] = g[N - 1] / ((a[N - 1][N - 1] > kEpsilon) ? a[N - 1][N - 1] : kEpsilon);
#pragma unroll
    for (int i = N - 2; i >= 0; i--) {
        double top = g[i];
#pragma unroll
        for (int k = i + 1; k < N; k++) top -= a[i][k] * change[k];
        change[i] = top / ((a[i][i] > kEpsilon) ? a[i][i] : kEpsilon);
    }
    double np[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
        const double hi = pmax[j], lo = pmin[j];
        np[j] = s.p[j] + change[j];
        if (np[j] > hi && s.lambda[j] < 1000) {
            np[j] = hi;
            if (s.lambda[j] < 1000) s.lambda[j] *= VFACTOR;
        }
        if (np[j] < lo) {
            np[j] = lo;
            if (s.lambda[j] < 1000) s.lambda[j] *= VFACTOR;
        }
    }
    double newSSE = 0, newSsr = 0, newSc = 0;
    if (nData > 0) Fit<F>::clamp(np);
#pragma unroll kLmUnroll
    for (int i = 0; i < nData; i++) {
        const double error = Y[i] - Fit<F>::eval(X[i], np);
        const double w = W[i];
        newSSE += error * error * w * w;
        if (SCORED) {
            const double r = w * error;          // W (Y - sim)
            newSsr += r * r;
            newSc += r * error;                  // W (Y - sim) (Y - sim), left to right
        }
    }
    diffSSE = s.mySSE - newSSE;
    if (diffSSE > 0) {
        s.mySSE = newSSE;
        if (SCORED) { s.ssr = newSsr; s.sc = newSc; }
#pragma unroll
        for (int j = 0; j < N; j++) { s.p[j] = np[j]; s.lambda[j] /= VFACTOR; }
    } else {
#pragma unroll
        for (int j = 0; j < N; j++) s.lambda[j] *= VFACTOR;
    }
    s.iter++;
    return !(fabs(diffSSE) > eps && s.iter <= maxIter);
}

template <int F>
__device__ __noinline__ void lmElevation(const double* __restrict__ pmin, const double* __restrict__ pmax,
                                         const double* __restrict__ delta, double* par, int maxIter, double eps,
                                         const double* X, const double* Y, const double* W, int nData)
{
    constexpr int N = Fit<F>::N;
    double dl[N], rdl[N];
#pragma unroll
    for (int k = 0; k < N; k++) { dl[k] = delta[k]; rdl[k] = 1.0 / dl[k]; }
    LmState<F> s;
    lmInit<F>(s, par, X, Y, W, nData);
    while (!lmIter<F>(s, pmin, pmax, dl, rdl, maxIter, eps, X, Y, W, nData)) {}
#pragma unroll
    for (int k = 0; k < N; k++) par[k] = s.p[k];
}

// ---------------------------------------------------------------------------------------------------------
// The same descent run by a whole WARP (single time step / a handful of macro areas: the GPU is otherwise idle and the
// latency of one descent is what counts). The data points of an iteration are spread over the lanes for everything that is
// independent per point (model evaluations, the four divisions, the products); every SUM is then taken in ascending point
// order by ONE lane per accumulator (lane t owns accumulator t; the terms of a 32-point chunk go through shared memory), so
// each accumulator receives exactly the terms, in exactly the order, of the sequential form: same bits, same accept /
// reject path. The small N x N solve and the step are computed redundantly by every lane (uniform data).
// sm: kLmWarpTerms x kLmWarpStride doubles of shared memory owned by this warp.
// ---------------------------------------------------------------------------------------------------------
constexpr int kLmWarpTerms = 6 * 7 / 2 + 6 + 1;        // upper triangle + gradient (+ 1 spare) for N = 6
constexpr int kLmWarpStride = 33;                       // doubles per term row (32 lanes + 1: the row readers hit different banks)

template <int F>
__device__ __forceinline__ double lmWarpSse(const double* p /* clamped */, const double* X, const double* Y, const double* W, int nData,
                                            double* sm)
{
    const int lane = threadIdx.x & 31;
    double sse = 0;
    for (int base = 0; base < nData; base += 32) {
        const int j = base + lane;
        double term = 0.;                                  // lanes past the last point add +0: the sum keeps its bits
        if (j < nData) {
            const double error = Y[j] - Fit<F>::eval(X[j], p);
            const double w = W[j];
            term = error * error * w * w;
        }
        sm[lane] = term;
        __syncwarp();
        // every lane: the same ordered sum (broadcast reads); a fixed trip count, so the 32 loads are in flight at once and
        // only the chain of additions is serial
#pragma unroll
        for (int l = 0; l < 32; l++) sse += sm[l];
        __syncwarp();
    }
    return sse;
}

template <int F>
__device__ __noinline__ void lmElevationWarp(const double* __restrict__ pmin, const double* __restrict__ pmax,
                                             const double* __restrict__ delta, double* par, int maxIter, double eps,
                                             const double* X, const double* Y, const double* W, int nData, double* sm)
{
    constexpr int N = Fit<F>::N;
    constexpr int NA = N * (N + 1) / 2;
    const double VFACTOR = 10;
    const int lane = threadIdx.x & 31;
    double dl[N], rdl[N], p[N], lambda[N];
#pragma unroll
    for (int k = 0; k < N; k++) { dl[k] = delta[k]; rdl[k] = 1.0 / dl[k]; lambda[k] = 0.01; p[k] = par[k]; }
    if (nData > 0) Fit<F>::clamp(p);
    double mySSE = lmWarpSse<F>(p, X, Y, W, nData, sm);
    int iter = 0;
    double diffSSE;
    do {
        double p0[N], up[N], rs[N];
#pragma unroll
        for (int k = 0; k < N; k++) p0[k] = p[k];
#pragma unroll
        for (int k = 0; k < N; k++) {
            p[k] += dl[k];
            if (nData > 0) Fit<F>::clamp(p);
            up[k] = p[k];
            p[k] -= dl[k];
            rs[k] = p[k];
        }
        if (nData > 0) Fit<F>::clamp(p);
        double acc = 0;                          // lane t < NA: a[i][c] (row-major upper triangle); NA <= t < NA + N: g[t - NA]
        for (int base = 0; base < nData; base += 32) {
            const int j = base + lane;
            if (j < nData) {
                const double x = X[j], y = Y[j], w = W[j];
                double P[N], WP[N];
                const double f0 = lmPointJacobian<F>(x, p0, up, rs, p[2 < N ? 2 : 0], dl, rdl, P);
#pragma unroll
                for (int k = 0; k < N; k++) WP[k] = w * P[k];
                int t = 0;
#pragma unroll
                for (int i = 0; i < N; i++) {
#pragma unroll
                    for (int c = i; c < N; c++) sm[(t++) * kLmWarpStride + lane] = (P[i] * WP[c]);
                }
#pragma unroll
                for (int i = 0; i < N; i++) sm[(NA + i) * kLmWarpStride + lane] = P[i] * w * (y - f0);
            } else {
                // past the last point: the terms are +0 (a sum keeps its bits), so the accumulation below has a fixed trip count
#pragma unroll
                for (int t = 0; t < NA + N; t++) sm[t * kLmWarpStride + lane] = 0.;
            }
            __syncwarp();
            if (lane < NA + N) {
                const double* src = sm + lane * kLmWarpStride;
#pragma unroll
                for (int l = 0; l < 32; l++) acc += src[l];      // 32 loads in flight, only the additions are serial
            }
            __syncwarp();
        }
        double g[N], a[N][N];
        {
            int t = 0;
#pragma unroll
            for (int i = 0; i < N; i++) {
#pragma unroll
                for (int c = i; c < N; c++) a[i][c] = __shfl_sync(0xffffffffu, acc, t++);
            }
#pragma unroll
            for (int i = 0; i < N; i++) g[i] = __shfl_sync(0xffffffffu, acc, NA + i);
        }
#pragma unroll
        for (int k = 0; k < N; k++) {
            a[k][k] += lambda[k] * a[k][k];
#pragma unroll
            for (int j = k + 1; j < N; j++) a[j][k] = a[k][j];
        }
#pragma unroll
        for (int j = 0; j < N - 1; j++) {
            const double pivot = (a[j][j] > kEpsilon) ? a[j][j] : kEpsilon;
#pragma unroll
            for (int i = j + 1; i < N; i++) {
                const double mult = a[i][j] / pivot;
#pragma unroll
                for (int k = j + 1; k < N; k++) a[i][k] -= mult * a[j][k];
                g[i] -= mult * g[j];
            }
        }
        double change[N];
        change[N - 1] = g[N - 1] / ((a[N - 1][N - 1] > kEpsilon) ? a[N - 1][N - 1] : kEpsilon);
#pragma unroll
        for (int i = N - 2; i >= 0; i--) {
            double top = g[i];
#pragma unroll
            for (int k = i + 1; k < N; k++) top -= a[i][k] * change[k];
            change[i] = top / ((a[i][i] > kEpsilon) ? a[i][i] : kEpsilon);
        }
        double np[N];
#pragma unroll
        for (int j = 0; j < N; j++) {
            const double hi = pmax[j], lo = pmin[j];
            np[j] = p[j] + change[j];
            if (np[j] > hi && lambda[j] < 1000) {
                np[j] = hi;
                if (lambda[j] < 1000) lambda[j] *= VFACTOR;
            }
            if (np[j] < lo) {
                np[j] = lo;
                if (lambda[j] < 1000) lambda[j] *= VFACTOR;
            }
        }
        if (nData > 0) Fit<F>::clamp(np);
        const double newSSE = lmWarpSse<F>(np, X, Y, W, nData, sm);
        diffSSE = mySSE - newSSE;
        if (diffSSE > 0) {
            mySSE = newSSE;
#pragma unroll
            for (int j = 0; j < N; j++) { p[j] = np[j]; lambda[j] /= VFACTOR; }
        } else {
#pragma unroll
            for (int j = 0; j < N; j++) lambda[j] *= VFACTOR;
        }
        iter++;
    } while (fabs(diffSSE) > eps && iter <= maxIter);
#pragma unroll
    for (int k = 0; k < N; k++) par[k] = p[k];
}

// computeWeighted_R2 (:1245-1275) and computeWeighted_RMSE (:1301-1322) of one fitted curve
template <int F>
__device__ __forceinline__ void weightedScores(double* par, const double* X, const double* Y, const double* W, int nData,
                                               double& R2, double& RMSE)
{
    double mean = 0, sw = 0;
    for (int i = 0; i < nData; i++) { mean += Y[i] * W[i]; sw += W[i]; }
    mean /= sw;
    if (nData > 0) Fit<F>::clamp(par);
    double ssr = 0, sst = 0, s = 0;
    for (int i = 0; i < nData; i++) {
        const double sim = Fit<F>::eval(X[i], par);
        const double r = W[i] * (Y[i] - sim);
        ssr += r * r;
        const double t = W[i] * (Y[i] - mean);
        sst += t * t;
        s += W[i] * (Y[i] - sim) * (Y[i] - sim);
    }
    R2 = 1.0 - (ssr / sst);
    RMSE = sqrt(s / (sw * nData));
}

// the parts of computeWeighted_R2 / _RMSE that do not depend on the fitted curve (weighted mean, sum of weights, total sum of
// squares): the same for every first guess of a problem, so the queue form computes them once per lane instead of once per
// descent. Same sums in the same order as weightedScores.
struct ScoreBase { double mean, sw, sst; };
__device__ __forceinline__ ScoreBase weightedScoreBase(const double* Y, const double* W, int nData)
{
    ScoreBase b;
    double mean = 0, sw = 0;
    for (int i = 0; i < nData; i++) { mean += Y[i] * W[i]; sw += W[i]; }
    mean /= sw;
    double sst = 0;
    for (int i = 0; i < nData; i++) {
        const double t = W[i] * (Y[i] - mean);
        sst += t * t;
    }
    b.mean = mean; b.sw = sw; b.sst = sst;
    return b;
}
template <int F>
__device__ __forceinline__ void weightedScoresWith(const ScoreBase& b, double* par, const double* X, const double* Y, const double* W,
                                                   int nData, double& R2, double& RMSE)
{
    if (nData > 0) Fit<F>::clamp(par);
    double ssr = 0, s = 0;
    for (int i = 0; i < nData; i++) {
        const double sim = Fit<F>::eval(X[i], par);
        const double r = W[i] * (Y[i] - sim);
        ssr += r * r;
        s += W[i] * (Y[i] - sim) * (Y[i] - sim);
    }
    R2 = 1.0 - (ssr / b.sst);
    RMSE = sqrt(s / (b.sw * nData));
}

// computeR2adjusted (:1175-1194) and computeRMSE (:1324-1340) of one fitted curve (the unweighted best fit of glocal detrending)
template <int F>
__device__ __forceinline__ void plainScores(double* par, const double* X, const double* Y, int nData, double& R2, double& RMSE)
{
    double mean = 0;
    for (int i = 0; i < nData; i++) mean += Y[i];
    mean /= double(nData);
    if (nData > 0) Fit<F>::clamp(par);
    double rss = 0, tss = 0;
    for (int i = 0; i < nData; i++) {
        const double sim = Fit<F>::eval(X[i], par);
        rss += (Y[i] - sim) * (Y[i] - sim);
        tss += (Y[i] - mean) * (Y[i] - mean);
    }
    R2 = 1 - rss / tss;
    RMSE = sqrt(rss / double(nData));
}

constexpr int kFitResultStride = PB_MAX_FIT_PARAMS + 3;       // parameters, R2, RMSE, flags of one first guess
constexpr int kFitResultDoubles = PB_MAX_FIRST_GUESS * kFitResultStride;

// bestFittingMarquardt_nDimension (:1360-1427): all first guesses of the group in parallel, reference pick order.
// firstGuess: nGuess x PB_MAX_FIT_PARAMS doubles. par (out, group-uniform) = chosen parameters. Returns bestR2.
// unweighted: the form without weights (:1433-1529, glocal detrending) — W must hold 1.0 everywhere (the descent :2125-2283
// is the weighted one with the weights dropped: every product by exactly 1.0 is exact); a fit that ends outside its
// parameter ranges is skipped, the scores are computeR2adjusted / computeRMSE, and the three-piece forms are only taken
// while their second inversion point stays below the highest station.
template <int G, int F>
__device__ double bestFitElevation(const double* __restrict__ pmin, const double* __restrict__ pmax,
                                   const double* __restrict__ delta, const double* __restrict__ firstGuess, int nGuess,
                                   double* par, int maxIter, double eps, double deltaR2, const double* X, const double* Y,
                                   const double* W, int nData, bool unweighted = false, double* results = nullptr)
{
    constexpr int N = Fit<F>::N;
    const int lane = grpLane<G>();
    double best[N];
#pragma unroll
    for (int j = 0; j < N; j++) best[j] = 0;
    double bestR2 = -9999, bestRMSE = -9999;
    int rmseIndex = -9999;
    double maxZ = 0;
    if (unweighted && nData > 0) {
        maxZ = X[0];
        for (int i = 0; i < nData; i++) if (X[i] > maxZ) maxZ = X[i];
    }
    if (results && nGuess > G) {
        // More first guesses than lanes (small groups: more problems per warp). Every lane owns the guesses lane, lane + G, ...
        // and walks them as ONE loop of Levenberg-Marquardt iterations: a lane whose descent ends scores it, stores it and
        // starts its next guess while the others keep iterating, so the lanes' totals (sums of 2-60 iterations each) even
        // out instead of every round waiting for its longest descent. Results go to `results` (kFitResultStride doubles
        // per guess); the pick below reads them in guess order: same rule, same winner.
        double dl[N], rdl[N];
#pragma unroll
        for (int j = 0; j < N; j++) { dl[j] = delta[j]; rdl[j] = 1.0 / dl[j]; }
        LmState<F> st;
        int k = lane;
        bool active = k < nGuess;
        ScoreBase base{};
        if (active && !unweighted) base = weightedScoreBase(Y, W, nData);
        if (active) lmInit<F>(st, firstGuess + k * PB_MAX_FIT_PARAMS, X, Y, W, nData);
        while (grpBallot<G>(active)) {
            if (active && lmIter<F>(st, pmin, pmax, dl, rdl, maxIter, eps, X, Y, W, nData)) {
                double q[N], R2 = 0, RMSE = 0;
                bool inRange = true, secondBelow = true;
#pragma unroll
                for (int j = 0; j < N; j++) q[j] = st.p[j];
                if (unweighted) {
#pragma unroll
                    for (int j = 0; j < N; j++) inRange = inRange && !(q[j] < pmin[j] || q[j] > pmax[j]);
                    plainScores<F>(q, X, Y, nData, R2, RMSE);
                    if (N > 4) secondBelow = (q[0] + q[2]) < maxZ;
                } else {
                    weightedScoresWith<F>(base, q, X, Y, W, nData, R2, RMSE);
                }
                double* r = results + (size_t)k * kFitResultStride;
#pragma unroll
                for (int j = 0; j < N; j++) r[j] = q[j];
                r[PB_MAX_FIT_PARAMS] = R2;
                r[PB_MAX_FIT_PARAMS + 1] = RMSE;
                r[PB_MAX_FIT_PARAMS + 2] = double((inRange ? 1 : 0) | (secondBelow ? 2 : 0));
                k += G;
                active = k < nGuess;
                if (active) lmInit<F>(st, firstGuess + k * PB_MAX_FIT_PARAMS, X, Y, W, nData);
            }
        }
        grpSync<G>();
        for (int kk = 0; kk < nGuess; kk++) {
            const double* r = results + (size_t)kk * kFitResultStride;
            const int flags = int(r[PB_MAX_FIT_PARAMS + 2]);
            if (!(flags & 1)) continue;
            const double r2 = r[PB_MAX_FIT_PARAMS], rmse = r[PB_MAX_FIT_PARAMS + 1];
            if (isEqualD(bestRMSE, -9999) || rmse < bestRMSE) { bestRMSE = rmse; rmseIndex = kk; }
            if (isEqualD(bestR2, -9999) || (r2 > (bestR2 + deltaR2) && (flags & 2))) {
#pragma unroll
                for (int j = 0; j < N; j++) best[j] = r[j];
                bestR2 = r2;
            }
        }
        grpSync<G>();
    } else
    for (int c0 = 0; c0 < nGuess; c0 += G) {
        const int k = c0 + lane;
        double q[N], R2 = 0, RMSE = 0;
        bool inRange = true, secondBelow = true;
#pragma unroll
        for (int j = 0; j < N; j++) q[j] = 0;
        if (k < nGuess) {
#pragma unroll
            for (int j = 0; j < N; j++) q[j] = firstGuess[k * PB_MAX_FIT_PARAMS + j];
            lmElevation<F>(pmin, pmax, delta, q, maxIter, eps, X, Y, W, nData);
            if (unweighted) {
#pragma unroll
                for (int j = 0; j < N; j++) inRange = inRange && !(q[j] < pmin[j] || q[j] > pmax[j]);
                plainScores<F>(q, X, Y, nData, R2, RMSE);
                if (N > 4) secondBelow = (q[0] + q[2]) < maxZ;
            } else {
                weightedScores<F>(q, X, Y, W, nData, R2, RMSE);
            }
        }
        grpSync<G>();
        const int nk = (nGuess - c0 < G) ? nGuess - c0 : G;
        const unsigned okMask = grpBallot<G>(inRange), belowMask = grpBallot<G>(secondBelow);
        for (int kk = 0; kk < nk; kk++) {
            const double r2 = grpBcast<G>(R2, kk), rmse = grpBcast<G>(RMSE, kk);
            double cand[N];
#pragma unroll
            for (int j = 0; j < N; j++) cand[j] = grpBcast<G>(q[j], kk);
            if (!((okMask >> kk) & 1u)) continue;                // rangeFlag false: `continue` (:1481-1482)
            if (isEqualD(bestRMSE, -9999) || rmse < bestRMSE) { bestRMSE = rmse; rmseIndex = c0 + kk; }
            if (isEqualD(bestR2, -9999) || (r2 > (bestR2 + deltaR2) && ((belowMask >> kk) & 1u))) {
#pragma unroll
                for (int j = 0; j < N; j++) best[j] = cand[j];
                bestR2 = r2;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < N; j++) par[j] = best[j];
    if (bestR2 < 0 && rmseIndex >= 0) {     // refit from the best-RMSE guess (every lane computes the same descent)
        double q[N];
#pragma unroll
        for (int j = 0; j < N; j++) q[j] = firstGuess[rmseIndex * PB_MAX_FIT_PARAMS + j];
        lmElevation<F>(pmin, pmax, delta, q, maxIter, eps, X, Y, W, nData);
#pragma unroll
        for (int j = 0; j < N; j++) par[j] = q[j];
    }
    return bestR2;
}

// ---------------------------------------------------------------------------------------------------------
// fittingMarquardt_nDimension for the other proxies (furtherMathFunctions.cpp:1740-1947): functionSum of nPred
// functionLinear_intercept terms, 2*nPred parameters, weighted, no pivot clamp. One descent, group-uniform (every lane
// computes it). Predictor c of data point j is pred(j, c).
// ---------------------------------------------------------------------------------------------------------
template <int NP, class PredFn>
__device__ __noinline__ void lmOthers(int nPredRt, double (*pmin)[2], double (*pmax)[2], double (*par)[2],
                                      double (*delta)[2], int maxIter, double eps, PredFn pred, const double* Y,
                                      const double* W, int nData)
{
    constexpr int MT = NP > 0 ? 2 * NP : 2 * PB_MAX_PROXY;     // NP > 0: compile-time size, everything in registers
    const double VFACTOR = 10;
    const int nPred = NP > 0 ? NP : nPredRt;
    const int nTot = 2 * nPred;
    double lambda[MT], change[MT], np[MT], p[MT], dl[MT];
    #pragma unroll
    for (int t = 0; t < nTot; t++) { lambda[t] = 0.01; p[t] = par[t >> 1][t & 1]; dl[t] = delta[t >> 1][t & 1]; }
    auto fsum = [&](int j, const double* q) {
        double r = 0.0;
        #pragma unroll
        for (int c = 0; c < nPred; c++) r += q[2 * c] * pred(j, c) + q[2 * c + 1];
        return r;
    };
    auto norm = [&](const double* q) {
        double n = 0;
        for (int i = 0; i < nData; i++) {
            const double e = Y[i] - fsum(i, q);
            n += e * e * W[i] * W[i];
        }
        return n;
    };
    double mySSE = norm(p), diffSSE;
    int iter = 0;
    do {
        double g[MT], a[MT][MT], p0[MT], up[MT], rs[MT];
        #pragma unroll
        for (int i = 0; i < nTot; i++) {
            g[i] = 0;
            p0[i] = p[i];
            #pragma unroll
            for (int j = 0; j < nTot; j++) a[i][j] = 0;
        }
        #pragma unroll
        for (int t = 0; t < nTot; t++) {
            p[t] += dl[t];
            up[t] = p[t];
            p[t] -= dl[t];
            rs[t] = p[t];
        }
        for (int j = 0; j < nData; j++) {
            const double f0 = fsum(j, p0);
            double P[MT], WP[MT];
            #pragma unroll
            for (int t = 0; t < nTot; t++) {
                double q[MT];
                #pragma unroll
                for (int c = 0; c < nTot; c++) q[c] = (c < t) ? rs[c] : p0[c];
                q[t] = up[t];
                P[t] = (fsum(j, q) - f0) / dl[t];
                WP[t] = W[j] * P[t];
            }
            #pragma unroll
            for (int i = 0; i < nTot; i++) {
                #pragma unroll
                for (int c = i; c < nTot; c++) a[i][c] += (P[i] * WP[c]);
                g[i] += P[i] * W[j] * (Y[j] - f0);
            }
        }
        #pragma unroll
        for (int t = 0; t < nTot; t++) {
            a[t][t] += lambda[t] * a[t][t];
            #pragma unroll
            for (int j = t + 1; j < nTot; j++) a[j][t] = a[t][j];
        }
        #pragma unroll
        for (int j = 0; j < nTot - 1; j++) {
            const double pivot = a[j][j];
            #pragma unroll
            for (int i = j + 1; i < nTot; i++) {
                const double mult = a[i][j] / pivot;
                #pragma unroll
                for (int k = j + 1; k < nTot; k++) a[i][k] -= mult * a[j][k];
                g[i] -= mult * g[j];
            }
        }
        change[nTot - 1] = g[nTot - 1] / a[nTot - 1][nTot - 1];
        #pragma unroll
        for (int i = nTot - 2; i >= 0; i--) {
            double top = g[i];
            #pragma unroll
            for (int k = i + 1; k < nTot; k++) top -= a[i][k] * change[k];
            change[i] = top / a[i][i];
        }
        #pragma unroll
        for (int t = 0; t < nTot; t++) {
            const double hi = pmax[t >> 1][t & 1], lo = pmin[t >> 1][t & 1];
            np[t] = p[t] + change[t];
            if (np[t] > hi && lambda[t] < 1000) {
                np[t] = hi;
                if (lambda[t] < 1000) lambda[t] *= VFACTOR;
            }
            if (np[t] < lo) {
                np[t] = lo;
                if (lambda[t] < 1000) lambda[t] *= VFACTOR;
            }
        }
        const double newSSE = norm(np);
        diffSSE = mySSE - newSSE;
        if (diffSSE > 0) {
            mySSE = newSSE;
            #pragma unroll
            for (int t = 0; t < nTot; t++) { p[t] = np[t]; lambda[t] /= VFACTOR; }
        } else {
            #pragma unroll
            for (int t = 0; t < nTot; t++) lambda[t] *= VFACTOR;
        }
        iter++;
    } while (fabs(diffSSE) > eps && iter <= maxIter);
    #pragma unroll
    for (int t = 0; t < nTot; t++) par[t >> 1][t & 1] = p[t];
}

// lmOthers run by a whole warp (few units, see lmElevationWarp): data points over the lanes, every sum in ascending point
// order by one lane per accumulator through shared memory; NP = 1..3 predictors (at most 27 accumulators). Same bits.
template <int NP, class PredFn>
__device__ __noinline__ void lmOthersWarp(double (*pmin)[2], double (*pmax)[2], double (*par)[2], double (*delta)[2], int maxIter,
                                          double eps, PredFn pred, const double* Y, const double* W, int nData, double* sm)
{
    constexpr int MT = 2 * NP;
    constexpr int NA = MT * (MT + 1) / 2;
    static_assert(NA + MT <= kLmWarpTerms, "too many accumulators for the warp form");
    const double VFACTOR = 10;
    const int lane = threadIdx.x & 31;
    double lambda[MT], change[MT], np[MT], p[MT], dl[MT];
#pragma unroll
    for (int t = 0; t < MT; t++) { lambda[t] = 0.01; p[t] = par[t >> 1][t & 1]; dl[t] = delta[t >> 1][t & 1]; }
    auto fsum = [&](int j, const double* q) {
        double r = 0.0;
#pragma unroll
        for (int c = 0; c < NP; c++) r += q[2 * c] * pred(j, c) + q[2 * c + 1];
        return r;
    };
    auto norm = [&](const double* q) {
        double n = 0;
        for (int base = 0; base < nData; base += 32) {
            const int j = base + lane;
            double term = 0.;                   // +0 past the last point
            if (j < nData) {
                const double e = Y[j] - fsum(j, q);
                term = e * e * W[j] * W[j];
            }
            sm[lane] = term;
            __syncwarp();
#pragma unroll
            for (int l = 0; l < 32; l++) n += sm[l];
            __syncwarp();
        }
        return n;
    };
    double mySSE = norm(p), diffSSE;
    int iter = 0;
    do {
        double g[MT], a[MT][MT], p0[MT], up[MT], rs[MT];
#pragma unroll
        for (int t = 0; t < MT; t++) {
            p0[t] = p[t];
        }
#pragma unroll
        for (int t = 0; t < MT; t++) {
            p[t] += dl[t];
            up[t] = p[t];
            p[t] -= dl[t];
            rs[t] = p[t];
        }
        double acc = 0;
        for (int base = 0; base < nData; base += 32) {
            const int j = base + lane;
            if (j < nData) {
                const double f0 = fsum(j, p0);
                double P[MT], WP[MT];
#pragma unroll
                for (int t = 0; t < MT; t++) {
                    double q[MT];
#pragma unroll
                    for (int c = 0; c < MT; c++) q[c] = (c < t) ? rs[c] : p0[c];
                    q[t] = up[t];
                    P[t] = (fsum(j, q) - f0) / dl[t];
                    WP[t] = W[j] * P[t];
                }
                int k = 0;
#pragma unroll
                for (int i = 0; i < MT; i++) {
#pragma unroll
                    for (int c = i; c < MT; c++) sm[(k++) * kLmWarpStride + lane] = (P[i] * WP[c]);
                }
#pragma unroll
                for (int i = 0; i < MT; i++) sm[(NA + i) * kLmWarpStride + lane] = P[i] * W[j] * (Y[j] - f0);
            } else {
#pragma unroll
                for (int t = 0; t < NA + MT; t++) sm[t * kLmWarpStride + lane] = 0.;      // +0 past the last point
            }
            __syncwarp();
            if (lane < NA + MT) {
                const double* src = sm + lane * kLmWarpStride;
#pragma unroll
                for (int l = 0; l < 32; l++) acc += src[l];
            }
            __syncwarp();
        }
        {
            int k = 0;
#pragma unroll
            for (int i = 0; i < MT; i++) {
#pragma unroll
                for (int c = i; c < MT; c++) a[i][c] = __shfl_sync(0xffffffffu, acc, k++);
            }
#pragma unroll
            for (int i = 0; i < MT; i++) g[i] = __shfl_sync(0xffffffffu, acc, NA + i);
        }
#pragma unroll
        for (int t = 0; t < MT; t++) {
            a[t][t] += lambda[t] * a[t][t];
#pragma unroll
            for (int j = t + 1; j < MT; j++) a[j][t] = a[t][j];
        }
#pragma unroll
        for (int j = 0; j < MT - 1; j++) {
            const double pivot = a[j][j];
#pragma unroll
            for (int i = j + 1; i < MT; i++) {
                const double mult = a[i][j] / pivot;
#pragma unroll
                for (int k = j + 1; k < MT; k++) a[i][k] -= mult * a[j][k];
                g[i] -= mult * g[j];
            }
        }
        change[MT - 1] = g[MT - 1] / a[MT - 1][MT - 1];
#pragma unroll
        for (int i = MT - 2; i >= 0; i--) {
            double top = g[i];
#pragma unroll
            for (int k = i + 1; k < MT; k++) top -= a[i][k] * change[k];
            change[i] = top / a[i][i];
        }
#pragma unroll
        for (int t = 0; t < MT; t++) {
            const double hi = pmax[t >> 1][t & 1], lo = pmin[t >> 1][t & 1];
            np[t] = p[t] + change[t];
            if (np[t] > hi && lambda[t] < 1000) {
                np[t] = hi;
                if (lambda[t] < 1000) lambda[t] *= VFACTOR;
            }
            if (np[t] < lo) {
                np[t] = lo;
                if (lambda[t] < 1000) lambda[t] *= VFACTOR;
            }
        }
        const double newSSE = norm(np);
        diffSSE = mySSE - newSSE;
        if (diffSSE > 0) {
            mySSE = newSSE;
#pragma unroll
            for (int t = 0; t < MT; t++) { p[t] = np[t]; lambda[t] /= VFACTOR; }
        } else {
#pragma unroll
            for (int t = 0; t < MT; t++) lambda[t] *= VFACTOR;
        }
        iter++;
    } while (fabs(diffSSE) > eps && iter <= maxIter);
#pragma unroll
    for (int t = 0; t < MT; t++) par[t >> 1][t & 1] = p[t];
}

}  // namespace pb
