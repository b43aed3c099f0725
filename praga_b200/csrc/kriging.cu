// kriging.cu — ordinary kriging: host set-up of the system, FP64 estimate kernel, FP64 direct variance path and the
// C-ABI entry points. The tensor-core variance path lives in kriging_tc.cu.
//
// Reference (agrolib/interpolation/kriging.cpp; file-static state, one system at a time, not thread safe):
//   krigingVariogram  :97-202   V = [[gamma(|s_i - s_j|), 1], [1^T, 0]], then matrixInversion(V) :28-81 (Gauss-Jordan, no pivoting)
//   krigingSetWeight  :211-264  D_j = gamma(|p - s_j|), D_n = 1;  weight_i = sum_j Vinv[i][j] D_j   (i < n)
//   krigingResult     :271-279  sum_i weight_i val_i
//   krigingRMSE       :287-295  sqrt|sum_i weight_i D_i|   (the Lagrange term is NOT added)
//   krigingFreeMemory :299-331
//
// What runs where:
//   estimate  = u . D with u = Vinv[:, 0:n] val  (V is symmetric)  -> O(n) per cell, FP64 on the CUDA cores: the sum
//               cancels from ~1e4 down to ~1e1, FP32 variogram values would already cost 2e-4.
//   variance  = D^T Vinv[0:n, :] D, the only dense contraction. Two implementations:
//     TENSOR (default, bounded variogram models): with the covariance C = sill - Gamma = L L^T (Cholesky, SPD thanks
//        to the nugget), c = sill - D, M = L^-1, y = M c, z = M 1:
//            sum_i weight_i D_i = sill - ( |y|^2 - (z.y)(z.y - 1) / |z|^2 )
//        Whitening removes the cancellation of the raw form (sum|M c| ~ 40 against 6e4), so FP16x2-split operands with
//        FP32 accumulation on tcgen05 hold ~4e-6 on the RMSE (measured); the raw form loses 2e-3 in FP32.
//     DIRECT (FP64, exact reference formula): Vinv by the reference's own elimination order on the host, W = A D with a
//        tiled FP64 kernel. Used for the linear model (no sill, C is not SPD), on request, and as the anchor in tests.
#include <math.h>
#include <omp.h>
#include <string.h>
#include <algorithm>
#include <functional>
#include <string>
#include <vector>

#include "context.cuh"
#include "kriging.cuh"
#include "nvtx.cuh"

namespace pb {

// ---------------------------------------------------------------------------------------------------------
// device: variogram models (kriging.cpp:160-192, 228-253)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double variogram(const KrigingParams& k, double h)
{
    const double tmp = h / k.range;
    switch (k.mode) {
    case PB_KRIGING_SPHERICAL:
        return (h < k.range) ? k.nugget + k.sillNugget * (1.5 * tmp - 0.5 * tmp * tmp * tmp) : k.nugget + k.sillNugget;
    case PB_KRIGING_EXPONENTIAL: return k.nugget + k.sillNugget * (1. - exp(-3. * tmp));
    case PB_KRIGING_GAUSSIAN: return k.nugget + k.sillNugget * (1. - exp(-4. * tmp * tmp));
    default: return k.nugget + k.slope * h;
    }
}

__device__ __forceinline__ void targetXY(const KrigingTargets& t, long long c, double& x, double& y)
{
    if (t.xy) {
        x = t.xy[2 * c];
        y = t.xy[2 * c + 1];
    } else {      // raster cell centre, gis::getUtmXYFromRowCol (gis/gis.cpp:806-810)
        const int cell = t.validIdx[c];
        const int row = cell / t.grid.nrCols, col = cell - row * t.grid.nrCols;
        x = t.grid.llx + t.grid.cellSize * (col + 0.5);
        y = t.grid.lly + t.grid.cellSize * (t.grid.nrRows - row - 0.5);
    }
}

constexpr int kEstThreads = 128;
constexpr int kEstTile = 384;       // stations per shared-memory tile (x, y, u: 9 KiB: an estimate CTA fits an SM beside the two
                                    // 105 KiB CTAs of the variance kernel, see evaluate())

// estimate_c = sum_j u_j gamma(|p_c - s_j|) + u_n ; optionally stores D[j][c] (direct variance path)
__global__ void __launch_bounds__(kEstThreads)
kriging_estimate_kernel(KrigingParams k, const double* __restrict__ pos, const double* __restrict__ u, KrigingTargets t,
                        long long c0, int m, double* __restrict__ est, double* __restrict__ D, int ldD)
{
    __shared__ double sx[kEstTile], sy[kEstTile], su[kEstTile];
    const int c = blockIdx.x * kEstThreads + threadIdx.x;
    double x = 0, y = 0;
    if (c < m) targetXY(t, c0 + c, x, y);
    double acc = 0;
    for (int j0 = 0; j0 < k.n; j0 += kEstTile) {
        const int nj = min(kEstTile, k.n - j0);
        __syncthreads();
        for (int j = threadIdx.x; j < nj; j += kEstThreads) {
            sx[j] = pos[2 * (j0 + j)];
            sy[j] = pos[2 * (j0 + j) + 1];
            su[j] = u[j0 + j];
        }
        __syncthreads();
        if (c < m) {
            for (int j = 0; j < nj; j++) {
                const double dx = sx[j] - x, dy = sy[j] - y;
                const double g = variogram(k, sqrt(dx * dx + dy * dy));
                acc += su[j] * g;
                if (D) D[(size_t)(j0 + j) * ldD + c] = g;
            }
        }
    }
    if (c < m) {
        est[c] = acc + u[k.n];
        if (D) D[(size_t)k.n * ldD + c] = 1.;
    }
}

// the estimate on the tensor path (no D column is kept): the FP64 pipe bounds this kernel, so the pair loop is stripped to
// what the sum needs. h = d2 * rsqrt(d2) (1-2 ulp instead of the correctly rounded sqrt), h / range as a multiplication by the
// reciprocal, and the spherical model folded to gamma = nugget + h (A - B d2): 17 FP64 instructions per pair instead of
// ~40. The differences are ~1e-16 relative on gamma.
// Bounded model (spherical): gamma = sill beyond the range, so with S = sum_j u_j (0 up to rounding: the unbiasedness row of V)
//   estimate = u_n + sum_j u_j gamma_j = (u_n + sill S) + sum_{j: |p - s_j| < range} u_j (gamma_j - sill)
// and only the stations inside the range contribute. The cells of a CTA are one run of the valid-cell list (neighbours in
// a raster row): stations farther than the range from the CTA's bounding box are dropped for the whole CTA by an ORDERED
// compaction into the shared-memory tile (station order is kept, so the sum is deterministic): at C4 (range 200 km in a
// 1000 km box) ~1/8 of the pairs are left. Unbounded models (exponential, Gaussian) keep every station.
template <int MODE>
__global__ void __launch_bounds__(kEstThreads)
kriging_estimate_fast_kernel(KrigingParams k, const double* __restrict__ pos, const double* __restrict__ u, double sumU,
                             KrigingTargets t, long long c0, int m, double* __restrict__ est)
{
    constexpr bool kBounded = MODE == PB_KRIGING_SPHERICAL;
    constexpr int kWarps = kEstThreads / 32;
    __shared__ double sx[kEstTile], sy[kEstTile], su[kEstTile];
    __shared__ double sBox[4][kWarps];
    __shared__ int sWarpCnt[kWarps];
    const int c = blockIdx.x * kEstThreads + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool valid = c < m;
    double x = 0, y = 0;
    if (valid) targetXY(t, c0 + c, x, y);
    const double invR = 1.0 / k.range, range2 = k.range * k.range, sill = k.nugget + k.sillNugget;
    const double sphA = 1.5 * k.sillNugget * invR, sphB = 0.5 * k.sillNugget * invR * invR * invR;
    double bx0 = 0, bx1 = 0, by0 = 0, by1 = 0;
    if (kBounded) {
        bx0 = valid ? x : 1e300; bx1 = valid ? x : -1e300; by0 = valid ? y : 1e300; by1 = valid ? y : -1e300;
        for (int o = 16; o > 0; o >>= 1) {
            bx0 = fmin(bx0, __shfl_xor_sync(~0u, bx0, o));
            bx1 = fmax(bx1, __shfl_xor_sync(~0u, bx1, o));
            by0 = fmin(by0, __shfl_xor_sync(~0u, by0, o));
            by1 = fmax(by1, __shfl_xor_sync(~0u, by1, o));
        }
        if (lane == 0) { sBox[0][warp] = bx0; sBox[1][warp] = bx1; sBox[2][warp] = by0; sBox[3][warp] = by1; }
        __syncthreads();
        for (int w = 0; w < kWarps; w++) {
            bx0 = fmin(bx0, sBox[0][w]); bx1 = fmax(bx1, sBox[1][w]);
            by0 = fmin(by0, sBox[2][w]); by1 = fmax(by1, sBox[3][w]);
        }
    }
    double acc0 = 0, acc1 = 0;
    int j0 = 0;
    while (j0 < k.n) {
        // fill the tile with the next (up to kEstTile) stations that can be inside the range of a cell of this CTA
        int cnt = 0;
        __syncthreads();                                // the previous tile has been consumed
        while (j0 < k.n && cnt + kEstThreads <= kEstTile) {
            const int j = j0 + threadIdx.x;
            bool keep = false;
            double px = 0, py = 0;
            if (j < k.n) {
                px = pos[2 * j];
                py = pos[2 * j + 1];
                if (kBounded) {
                    const double ex = fmax(fmax(bx0 - px, px - bx1), 0.), ey = fmax(fmax(by0 - py, py - by1), 0.);
                    keep = ex * ex + ey * ey < range2;
                } else keep = true;
            }
            const unsigned b = __ballot_sync(~0u, keep);
            if (lane == 0) sWarpCnt[warp] = __popc(b);
            __syncthreads();
            int base = cnt, tot = 0;
            for (int w = 0; w < kWarps; w++) {
                if (w < warp) base += sWarpCnt[w];
                tot += sWarpCnt[w];
            }
            if (keep) {
                const int idx = base + __popc(b & ((1u << lane) - 1u));
                sx[idx] = px;
                sy[idx] = py;
                su[idx] = u[j];
            }
            cnt += tot;
            j0 += kEstThreads;
            __syncthreads();                            // sWarpCnt is rewritten by the next chunk; the tile is visible
        }
        if (valid) {
#pragma unroll 4
            for (int j = 0; j < cnt; j++) {
                const double dx = sx[j] - x, dy = sy[j] - y;
                const double d2 = dx * dx + dy * dy;
                const double h = d2 > 0. ? d2 * rsqrt(d2) : 0.;
                double g;
                if (MODE == PB_KRIGING_SPHERICAL) g = (d2 < range2) ? h * (sphA - sphB * d2) - k.sillNugget : 0.;   // gamma - sill
                else if (MODE == PB_KRIGING_EXPONENTIAL) g = k.nugget + k.sillNugget * (1. - exp(-3. * (h * invR)));
                else g = k.nugget + k.sillNugget * (1. - exp(-4. * (d2 * invR * invR)));
                if (j & 1) acc1 += su[j] * g;         // two accumulators: the dependent add chain is the latency floor
                else acc0 += su[j] * g;
            }
        }
    }
    if (valid) est[c] = (acc0 + acc1) + (kBounded ? u[k.n] + sill * sumU : u[k.n]);
}

// direct variance: q_c = sum_i D[i][c] * (sum_j A[i][j] D[j][c]), i < n, j <= n. 64 x 64 output tile, K step 16,
// 16 x 16 threads with a 4 x 4 register block each; the i-reduction is finished with one atomicAdd per (thread, cell).
constexpr int kQT = 64, kQK = 16;
__global__ void __launch_bounds__(256)
kriging_quadform_kernel(const double* __restrict__ A, int n, int ldA, const double* __restrict__ D, int ldD, int m,
                        double* __restrict__ q)
{
    __shared__ double sA[kQK][kQT + 1], sD[kQK][kQT + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i0 = blockIdx.y * kQT, c0 = blockIdx.x * kQT;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = 0;
    const int K = n + 1;
    for (int k0 = 0; k0 < K; k0 += kQK) {
        for (int e = threadIdx.x; e < kQK * kQT; e += 256) {
            const int kk = e & (kQK - 1), ii = e >> 4;      // A tile: consecutive threads walk along j (contiguous)
            const int i = i0 + ii, j = k0 + kk;
            sA[kk][ii] = (i < n && j < K) ? A[(size_t)i * ldA + j] : 0.;
            const int cc = e & (kQT - 1), k2 = e >> 6;      // D tile: consecutive threads walk along the cells
            const int c = c0 + cc, j2 = k0 + k2;
            sD[k2][cc] = (c < m && j2 < K) ? D[(size_t)j2 * ldD + c] : 0.;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kQK; kk++) {
            double av[4], dv[4];
#pragma unroll
            for (int a = 0; a < 4; a++) av[a] = sA[kk][ty * 4 + a];
#pragma unroll
            for (int b = 0; b < 4; b++) dv[b] = sD[kk][tx * 4 + b];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[a][b] += av[a] * dv[b];
        }
        __syncthreads();
    }
#pragma unroll
    for (int b = 0; b < 4; b++) {
        const int c = c0 + tx * 4 + b;
        if (c >= m) continue;
        double s = 0;
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const int i = i0 + ty * 4 + a;
            if (i < n) s += acc[a][b] * D[(size_t)i * ldD + c];
        }
        atomicAdd(&q[c], s);
    }
}

// epilogue shared by both variance paths: rmse = sqrt|q|; scatter to raster grids or to the point arrays
__global__ void kriging_finish_kernel(KrigingTargets t, long long c0, int m, const double* __restrict__ est,
                                      const double* __restrict__ q, double* __restrict__ outEst, double* __restrict__ outRmse,
                                      float* __restrict__ gridEst, float* __restrict__ gridRmse)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    const double r = q ? sqrt(fabs(q[c])) : 0.;
    if (t.xy) {
        outEst[c0 + c] = est[c];
        if (outRmse) outRmse[c0 + c] = r;
    } else {
        const int cell = t.validIdx[c0 + c];
        gridEst[cell] = float(est[c]);
        if (gridRmse) gridRmse[cell] = float(r);
    }
}

// ---------------------------------------------------------------------------------------------------------
// host set-up
// ---------------------------------------------------------------------------------------------------------
static double variogramHost(const KrigingParams& k, double h)
{
    const double tmp = h / k.range;
    switch (k.mode) {
    case PB_KRIGING_SPHERICAL:
        return (h < k.range) ? k.nugget + k.sillNugget * (1.5 * tmp - 0.5 * tmp * tmp * tmp) : k.nugget + k.sillNugget;
    case PB_KRIGING_EXPONENTIAL: return k.nugget + k.sillNugget * (1. - exp(-3. * tmp));
    case PB_KRIGING_GAUSSIAN: return k.nugget + k.sillNugget * (1. - exp(-4. * tmp * tmp));
    default: return k.nugget + k.slope * h;
    }
}

// matrixInversion (kriging.cpp:28-81): Gauss-Jordan without pivoting in the reference's elimination order; the row
// loops are independent inside one elimination step, so they are spread over the host cores without changing a bit.
static bool gaussJordanNoPivot(std::vector<double>& V, std::vector<double>& I, int dim)
{
    double* A = V.data();
    double* inv = I.data();
    std::fill(I.begin(), I.end(), 0.);
    for (int i = 0; i < dim; i++) inv[size_t(i) * dim + i] = 1.;
    for (int iter = 0; iter < dim - 1; iter++) {
        const int first = iter + 1;
        const double piv = A[size_t(iter) * dim + iter];
#pragma omp parallel for schedule(static)
        for (int i = first; i < dim; i++) {
            const double r = A[size_t(i) * dim + iter] / piv;
            A[size_t(i) * dim + iter] = 0.;
            for (int j = first; j < dim; j++) A[size_t(i) * dim + j] -= r * A[size_t(iter) * dim + j];
            for (int j = 0; j <= iter; j++) inv[size_t(i) * dim + j] -= r * inv[size_t(iter) * dim + j];
        }
    }
    for (int iter = dim - 1; iter > 0; iter--) {
        const double piv = A[size_t(iter) * dim + iter];
#pragma omp parallel for schedule(static)
        for (int i = 0; i < iter; i++) {
            const double r = A[size_t(i) * dim + iter] / piv;
            A[size_t(i) * dim + iter] = 0.;
            for (int j = 0; j < dim; j++) inv[size_t(i) * dim + j] -= r * inv[size_t(iter) * dim + j];
        }
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < dim; i++) {
        const double d = A[size_t(i) * dim + i];
        for (int j = 0; j < dim; j++) inv[size_t(i) * dim + j] /= d;
    }
    for (size_t e = 0; e < I.size(); e++)
        if (!std::isfinite(inv[e])) return false;      // zero pivot (nugget == 0): the reference returns NaN estimates
    return true;
}

// lower Cholesky factor in place (row-major, lower triangle); false if the matrix is not positive definite
static bool choleskyLower(std::vector<double>& C, int n)
{
    double* a = C.data();
    for (int k = 0; k < n; k++) {
        const double d = a[size_t(k) * n + k];
        if (!(d > 0.)) return false;
        const double l = sqrt(d);
        a[size_t(k) * n + k] = l;
#pragma omp parallel for schedule(static)
        for (int i = k + 1; i < n; i++) a[size_t(i) * n + k] /= l;
#pragma omp parallel for schedule(dynamic, 16)
        for (int i = k + 1; i < n; i++) {
            const double lik = a[size_t(i) * n + k];
            double* ri = a + size_t(i) * n;
            for (int j = k + 1; j <= i; j++) ri[j] -= lik * a[size_t(j) * n + k];
        }
    }
    return true;
}

// M = L^-1 (lower triangular), row-major
static void invertLower(const std::vector<double>& L, std::vector<double>& M, int n)
{
    std::fill(M.begin(), M.end(), 0.);
    // column c of M solves L m = e_c by forward substitution; columns are independent
#pragma omp parallel for schedule(dynamic, 8)
    for (int c = 0; c < n; c++) {
        std::vector<double> m(n, 0.);
        m[c] = 1. / L[size_t(c) * n + c];
        for (int i = c + 1; i < n; i++) {
            double s = 0;
            const double* li = L.data() + size_t(i) * n;
            for (int j = c; j < i; j++) s += li[j] * m[j];
            m[i] = -s / li[i];
        }
        for (int i = c; i < n; i++) M[size_t(i) * n + c] = m[i];
    }
}

// k-d ordering of the stations (tensor path, bounded model): recursive bisection along the longer side of the bounding
// box, split points at multiples of 32, so that every K stage of the variance kernel (32 consecutive stations) is a spatially
// compact leaf. The whitened quantities |y|^2 and z.y are invariant under a permutation of the stations (C' = P C P^T),
// so the order is free; a compact stage that lies out of every cell's range is skipped by the kernel (kriging_tc.cu).
static void kdOrder(const double* pos, int* idx, int lo, int hi)
{
    if (hi - lo <= kTcBK) return;
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
    for (int i = lo; i < hi; i++) {
        x0 = std::min(x0, pos[2 * idx[i]]); x1 = std::max(x1, pos[2 * idx[i]]);
        y0 = std::min(y0, pos[2 * idx[i] + 1]); y1 = std::max(y1, pos[2 * idx[i] + 1]);
    }
    const int axis = (x1 - x0 >= y1 - y0) ? 0 : 1;
    const int leaves = (hi - lo + kTcBK - 1) / kTcBK;
    const int mid = lo + (leaves / 2) * kTcBK;
    std::nth_element(idx + lo, idx + mid, idx + hi, [&](int a, int b) {
        const double va = pos[2 * a + axis], vb = pos[2 * b + axis];
        return va < vb || (va == vb && a < b);
    });
    kdOrder(pos, idx, lo, mid);
    kdOrder(pos, idx, mid, hi);
}

void krigingFree(KrigingSystem* k)
{
    if (!k) return;
    if (k->dPos) cudaFree(k->dPos);
    if (k->dU) cudaFree(k->dU);
    if (k->dA) cudaFree(k->dA);
    if (k->dMhi) cudaFree(k->dMhi);
    if (k->dMlo) cudaFree(k->dMlo);
    if (k->dZ) cudaFree(k->dZ);
    if (k->dKbBox) cudaFree(k->dKbBox);
    if (k->dMmaCount) cudaFree(k->dMmaCount);
    if (k->sEst) {
        cudaStreamDestroy(k->sEst);
        for (auto& e : k->evEst) cudaEventDestroy(e);
        for (auto& e : k->evFin) cudaEventDestroy(e);
    }
    if (k->sDn) {
        cudaStreamDestroy(k->sDn);
        for (auto& e : k->evDn) cudaEventDestroy(e);
    }
    if (k->dWork) cudaFree(k->dWork);
    delete k;
}

void krigingDestroyAll(std::vector<KrigingSystem*>& v)
{
    for (auto k : v) krigingFree(k);
    v.clear();
}

}  // namespace pb

using namespace pb;

namespace {

int fail(pb_context* ctx, int code, const std::string& msg) { return pb::failCtx(ctx, code, msg); }

KrigingSystem* systemOf(pb_context* ctx, pb_kriging_id id)
{
    if (id < 0 || id >= (int)ctx->kriging.size()) return nullptr;
    return ctx->kriging[id];
}

// run estimate + variance over targets [0, m) in chunks; results to host doubles (points) or device grids (raster)
int evaluate(pb_context* ctx, KrigingSystem* k, const KrigingTargets& t, long long m, bool wantRmse, double* dOutEst,
             double* dOutRmse, float* gridEst, float* gridRmse, const std::function<int(long long)>& afterChunk = nullptr)
{
    const int chunk = k->tensor ? kTcChunk : 8192;
    const size_t ldD = chunk;
    const size_t needD = (!k->tensor && wantRmse) ? size_t(k->n + 1) * ldD * sizeof(double) : 0;
    const size_t need = needD + 3 * size_t(chunk) * sizeof(double);          // two estimate buffers + the variance's
    if (k->workBytes < need) {
        if (k->dWork) cudaFree(k->dWork);
        k->dWork = nullptr;
        k->workBytes = 0;
        CUDA_TRY(ctx, cudaMalloc(&k->dWork, need));
        k->workBytes = need;
    }
    // tensor path with the variance wanted: the estimate of chunk c + 1 (FP64 pipe, its own low-priority stream) runs beside
    // the variance kernel of chunk c (tensor + FP32 pipes): two estimate buffers, events both ways
    const bool overlap = k->tensor && wantRmse && m > chunk && !getenv("PB_KRIGING_NO_OVERLAP");
    double* dEstBuf[2] = {reinterpret_cast<double*>(k->dWork), reinterpret_cast<double*>(k->dWork) + chunk};
    double* dQ = dEstBuf[1] + chunk;
    double* dD = needD ? dQ + chunk : nullptr;
    if (overlap && !k->sEst) {
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        CUDA_TRY(ctx, cudaStreamCreateWithPriority(&k->sEst, cudaStreamNonBlocking, least));
        for (auto& e : k->evEst) CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : k->evFin) CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    cudaStream_t sE = overlap ? k->sEst : ctx->stream;
    if (overlap) {                                     // the estimate stream starts behind whatever precedes this call
        CUDA_TRY(ctx, cudaEventRecord(k->evFin[0], ctx->stream));
        CUDA_TRY(ctx, cudaStreamWaitEvent(sE, k->evFin[0], 0));
    }
    auto launchEstimate = [&](long long c0, int mc, double* dEst) -> int {
        const int estGrid = (mc + kEstThreads - 1) / kEstThreads;
        if (dD || k->params.mode == PB_KRIGING_LINEAR)
            kriging_estimate_kernel<<<estGrid, kEstThreads, 0, sE>>>(k->params, k->dPos, k->dU, t, c0, mc, dEst, dD, (int)ldD);
        else if (k->params.mode == PB_KRIGING_SPHERICAL)
            kriging_estimate_fast_kernel<PB_KRIGING_SPHERICAL><<<estGrid, kEstThreads, 0, sE>>>(k->params, k->dPos, k->dU, k->sumU, t, c0, mc, dEst);
        else if (k->params.mode == PB_KRIGING_EXPONENTIAL)
            kriging_estimate_fast_kernel<PB_KRIGING_EXPONENTIAL><<<estGrid, kEstThreads, 0, sE>>>(k->params, k->dPos, k->dU, k->sumU, t, c0, mc, dEst);
        else
            kriging_estimate_fast_kernel<PB_KRIGING_GAUSSIAN><<<estGrid, kEstThreads, 0, sE>>>(k->params, k->dPos, k->dU, k->sumU, t, c0, mc, dEst);
        ctx->timing.launches++;
        CUDA_TRY(ctx, cudaGetLastError());
        return PB_OK;
    };
    long long nChunk = 0;
    if (overlap) {
        const int rc = launchEstimate(0, (int)std::min<long long>(chunk, m), dEstBuf[0]);
        if (rc) return rc;
        CUDA_TRY(ctx, cudaEventRecord(k->evEst[0], sE));
    }
    for (long long c0 = 0; c0 < m; c0 += chunk, nChunk++) {
        const int mc = (int)std::min<long long>(chunk, m - c0);
        double* dEst = dEstBuf[overlap ? (nChunk & 1) : 0];
        if (!overlap) {
            const int rc = launchEstimate(c0, mc, dEst);
            if (rc) return rc;
        } else if (c0 + chunk < m) {
            // the next chunk's estimate: its buffer was read by the finish kernel of chunk nChunk - 1
            const int nb = int((nChunk + 1) & 1);
            if (nChunk >= 1) CUDA_TRY(ctx, cudaStreamWaitEvent(sE, k->evFin[nb], 0));
            const int rc = launchEstimate(c0 + chunk, (int)std::min<long long>(chunk, m - c0 - chunk), dEstBuf[nb]);
            if (rc) return rc;
            CUDA_TRY(ctx, cudaEventRecord(k->evEst[nb], sE));
        }
        if (wantRmse) {
            if (k->tensor) {
                CUDA_TRY(ctx, launchKrigingVarianceTc(*k, t, c0, mc, dQ, ctx->stream));
                ctx->timing.launches++;
            } else {
                CUDA_TRY(ctx, cudaMemsetAsync(dQ, 0, sizeof(double) * mc, ctx->stream));
                dim3 grid((mc + kQT - 1) / kQT, (k->n + kQT - 1) / kQT);
                kriging_quadform_kernel<<<grid, 256, 0, ctx->stream>>>(k->dA, k->n, k->n + 1, dD, (int)ldD, mc, dQ);
                ctx->timing.launches++;
            }
        }
        if (overlap) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, k->evEst[nChunk & 1], 0));
        kriging_finish_kernel<<<(mc + 255) / 256, 256, 0, ctx->stream>>>(t, c0, mc, dEst, wantRmse ? dQ : nullptr, dOutEst,
                                                                          dOutRmse, gridEst, gridRmse);
        ctx->timing.launches++;
        CUDA_TRY(ctx, cudaGetLastError());
        if (overlap) CUDA_TRY(ctx, cudaEventRecord(k->evFin[nChunk & 1], ctx->stream));
        if (afterChunk) {
            const int rc = afterChunk(c0 + mc);      // the kernels of the targets [0, c0 + mc) are in the stream
            if (rc) return rc;
        }
    }
    return PB_OK;
}

}  // namespace

extern "C" {

/* krigingVariogram (kriging.cpp:97-202): builds and factorises the system for one set of stations and values */
int pb_kriging_setup(pb_context* ctx, const double* pos, const double* val, int n, int mode, double range, double nugget,
                     double sill, double slope, int flags, pb_kriging_id* id)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (!pos || !val || !id || n < 1) return fail(ctx, PB_ERR_INVALID, "null argument / no stations");
    cudaSetDevice(ctx->device);
    KrigingSystem* k = new KrigingSystem();
    k->n = n;
    k->params.n = n;
    k->params.mode = (mode >= PB_KRIGING_SPHERICAL && mode <= PB_KRIGING_GAUSSIAN) ? mode : PB_KRIGING_LINEAR;
    k->params.range = range;
    k->params.nugget = nugget;
    k->params.sillNugget = sill - nugget;
    k->params.slope = slope;
    k->params.sill = nugget + (sill - nugget);          // the value the reference's bounded models reach (:171)
    const int dim = n + 1;
    bool tensor = !(flags & PB_KRIGING_FLAG_FP64_DIRECT) && k->params.mode != PB_KRIGING_LINEAR;
    // tensor path, bounded model: stations in k-d order (kdOrder); the exact path keeps the caller's order (it eliminates in
    // the reference's order)
    std::vector<double> posK, valK;
    const double* posIn = pos;
    const double* valIn = val;
    const bool reorder = tensor && k->params.mode == PB_KRIGING_SPHERICAL && n > kTcBK;
    if (reorder) {
        std::vector<int> idx(n);
        for (int i = 0; i < n; i++) idx[i] = i;
        kdOrder(pos, idx.data(), 0, n);
        posK.resize(size_t(2) * n);
        valK.resize(n);
        for (int i = 0; i < n; i++) {
            posK[2 * i] = pos[2 * idx[i]];
            posK[2 * i + 1] = pos[2 * idx[i] + 1];
            valK[i] = val[idx[i]];
        }
        pos = posK.data();
        val = valK.data();
    }
    std::vector<double> G(size_t(n) * n);
    auto buildG = [&]() {
#pragma omp parallel for schedule(dynamic, 16)
        for (int i = 0; i < n; i++)
            for (int j = i; j < n; j++) {
                const double dx = pos[i * 2] - pos[j * 2], dy = pos[i * 2 + 1] - pos[j * 2 + 1];
                const double g = variogramHost(k->params, sqrt(dx * dx + dy * dy));
                G[size_t(i) * n + j] = g;
                G[size_t(j) * n + i] = g;
            }
    };
    buildG();
    std::vector<double> u(dim);
    std::vector<double> M;
    if (tensor) {
        // covariance form: C = sill - Gamma = L L^T
        std::vector<double> C(size_t(n) * n);
        const double c0 = k->params.sill;
        for (size_t e = 0; e < C.size(); e++) C[e] = c0 - G[e];
        double normC = 0;
        for (size_t e = 0; e < C.size(); e++) normC += C[e] * C[e];
        if (!choleskyLower(C, n)) tensor = false;     // not SPD (nugget 0, degenerate model): exact path instead
        else {
            M.resize(size_t(n) * n);
            invertLower(C, M, n);
            // condition guard: cond_F(C) <= |C|_F |M|_F^2. An ill-conditioned system (Gaussian model: the reference puts
            // the nugget on gamma(0), so C is a pure Gaussian correlation matrix) makes the reference's own no-pivot
            // elimination result rounding-dependent; only the same elimination order reproduces it -> exact path.
            double normM = 0;
            for (size_t e = 0; e < M.size(); e++) normM += M[e] * M[e];
            if (!(sqrt(normC) * normM < 1.0e8)) tensor = false;
        }
        if (tensor) {
            // z = M 1, t = M val; u1 = C^-1 (1 u2 - val), u2 = (1^T C^-1 val) / (1^T C^-1 1)   [V u = (val; 0)]
            std::vector<double> z(n), tv(n);
#pragma omp parallel for schedule(static)
            for (int i = 0; i < n; i++) {
                double sz = 0, st = 0;
                const double* mi = M.data() + size_t(i) * n;
                for (int j = 0; j <= i; j++) { sz += mi[j]; st += mi[j] * val[j]; }
                z[i] = sz;
                tv[i] = st;
            }
            double zz = 0, zt = 0;
            for (int i = 0; i < n; i++) { zz += z[i] * z[i]; zt += z[i] * tv[i]; }
            const double u2 = zt / zz;
            // u1 = M^T (z u2 - t)
#pragma omp parallel for schedule(static)
            for (int j = 0; j < n; j++) {
                double s = 0;
                for (int i = j; i < n; i++) s += M[size_t(i) * n + j] * (z[i] * u2 - tv[i]);
                u[j] = s;
            }
            u[n] = u2;
            int rc = krigingUploadWhitened(*k, M, z, zz, pos);
            if (rc != 0) { krigingFree(k); return fail(ctx, PB_ERR_CUDA, "kriging: upload of the whitened system failed"); }
        }
    }
    if (!tensor) {
        if (reorder) {          // the whitened system was rejected: back to the caller's station order
            pos = posIn;
            val = valIn;
            buildG();
        }
        std::vector<double> V(size_t(dim) * dim, 0.), I(size_t(dim) * dim);
        for (int i = 0; i < n; i++) {
            memcpy(&V[size_t(i) * dim], &G[size_t(i) * n], sizeof(double) * n);
            V[size_t(i) * dim + n] = 1.;
            V[size_t(n) * dim + i] = 1.;
        }
        if (!gaussJordanNoPivot(V, I, dim)) {
            krigingFree(k);
            return fail(ctx, PB_ERR_FIT, "kriging: singular system (zero pivot; the reference needs nugget > 0)");
        }
        // u_j = sum_i val_i Vinv[i][j]
        for (int j = 0; j < dim; j++) {
            double s = 0;
            for (int i = 0; i < n; i++) s += val[i] * I[size_t(i) * dim + j];
            u[j] = s;
        }
        if (cudaMalloc(&k->dA, sizeof(double) * size_t(n) * dim) != cudaSuccess ||
            cudaMemcpy(k->dA, I.data(), sizeof(double) * size_t(n) * dim, cudaMemcpyHostToDevice) != cudaSuccess) {
            krigingFree(k);
            return fail(ctx, PB_ERR_CUDA, "kriging: device allocation failed");
        }
    }
    k->tensor = tensor;
    k->sumU = 0;
    for (int j = 0; j < n; j++) k->sumU += u[j];
    if (cudaMalloc(&k->dPos, sizeof(double) * 2 * n) != cudaSuccess || cudaMalloc(&k->dU, sizeof(double) * dim) != cudaSuccess ||
        cudaMemcpy(k->dPos, pos, sizeof(double) * 2 * n, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(k->dU, u.data(), sizeof(double) * dim, cudaMemcpyHostToDevice) != cudaSuccess) {
        krigingFree(k);
        return fail(ctx, PB_ERR_CUDA, "kriging: device allocation failed");
    }
    int slot = -1;
    for (size_t i = 0; i < ctx->kriging.size(); i++) if (!ctx->kriging[i]) { slot = (int)i; break; }
    if (slot < 0) { ctx->kriging.push_back(nullptr); slot = (int)ctx->kriging.size() - 1; }
    ctx->kriging[slot] = k;
    *id = slot;
    return PB_OK;
}

/* krigingFreeMemory (kriging.cpp:299-331) */
int pb_kriging_free(pb_context* ctx, pb_kriging_id id)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    KrigingSystem* k = systemOf(ctx, id);
    if (!k) return fail(ctx, PB_ERR_INVALID, "bad kriging id");
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    krigingFree(k);
    ctx->kriging[id] = nullptr;
    return PB_OK;
}

int pb_kriging_uses_tensor_path(pb_context* ctx, pb_kriging_id id)
{
    PB_NVTX_RANGE();
    KrigingSystem* k = ctx ? systemOf(ctx, id) : nullptr;
    return k ? (k->tensor ? 1 : 0) : -1;
}

/* measurement aid: tensor-core MMAs (M = 256, N = 256, K = 16: 2 * 256 * 256 * 16 flop each) the variance kernel has issued
   for this system since the last call; the counter is reset. 0 on the FP64 direct path. */
int pb_kriging_mma_count(pb_context* ctx, pb_kriging_id id, unsigned long long* count)
{
    if (!ctx || !count) return PB_ERR_INVALID;
    KrigingSystem* k = systemOf(ctx, id);
    if (!k) return fail(ctx, PB_ERR_INVALID, "kriging: unknown system id");
    *count = 0;
    if (!k->dMmaCount) return PB_OK;
    cudaSetDevice(ctx->device);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaMemcpy(count, k->dMmaCount, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CUDA_TRY(ctx, cudaMemset(k->dMmaCount, 0, sizeof(unsigned long long)));
    return PB_OK;
}

/* krigingSetWeight + krigingResult + krigingRMSE (kriging.cpp:211-295) for m target points (x0, y0, x1, y1, ...) */
int pb_kriging_points(pb_context* ctx, pb_kriging_id id, const double* xy, int m, double* result, double* rmse)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    KrigingSystem* k = systemOf(ctx, id);
    if (!k) return fail(ctx, PB_ERR_INVALID, "bad kriging id");
    if (!xy || !result || m < 0) return fail(ctx, PB_ERR_INVALID, "null argument");
    if (m == 0) return PB_OK;
    cudaSetDevice(ctx->device);
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    double *dXY = nullptr, *dRes = nullptr, *dRm = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&dXY, sizeof(double) * 2 * m));
    CUDA_TRY(ctx, cudaMalloc(&dRes, sizeof(double) * 2 * m));
    dRm = dRes + m;
    CUDA_TRY(ctx, cudaMemcpyAsync(dXY, xy, sizeof(double) * 2 * m, cudaMemcpyHostToDevice, ctx->stream));
    KrigingTargets t{};
    t.xy = dXY;
    cudaEventRecord(ctx->ev[1], ctx->stream);
    int rc = evaluate(ctx, k, t, m, rmse != nullptr, dRes, dRm, nullptr, nullptr);
    cudaEventRecord(ctx->ev[2], ctx->stream);
    if (rc == PB_OK) {
        CUDA_TRY(ctx, cudaMemcpyAsync(result, dRes, sizeof(double) * m, cudaMemcpyDeviceToHost, ctx->stream));
        if (rmse) CUDA_TRY(ctx, cudaMemcpyAsync(rmse, dRm, sizeof(double) * m, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        cudaEventElapsedTime(&ctx->timing.kernel_ms, ctx->ev[1], ctx->ev[2]);
    }
    cudaFree(dXY);
    cudaFree(dRes);
    return rc;
}

/* kriging over the valid cells of a resident DEM (cell centres): estimate and (optionally) RMSE grids.
 * estimate->data / rmse->data (or ->rows) receive the host grids; NULL data and rows = keep on the device only. */
int pb_kriging_raster(pb_context* ctx, pb_kriging_id id, pb_raster_id demId, int rowBegin, int rowEnd,
                      pb_raster_out* estimate, pb_raster_out* rmse)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    KrigingSystem* k = systemOf(ctx, id);
    if (!k) return fail(ctx, PB_ERR_INVALID, "bad kriging id");
    if (!estimate) return fail(ctx, PB_ERR_INVALID, "null argument");
    if (demId < 0 || demId >= (int)ctx->rasters.size() || !ctx->rasters[demId].used)
        return fail(ctx, PB_ERR_INVALID, "dem id is not a live raster");
    cudaSetDevice(ctx->device);
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    RasterEntry& dem = ctx->rasters[demId];
    if (rowEnd < 0) rowEnd = dem.h.nrRows;
    if (rowBegin < 0 || rowBegin > rowEnd || rowEnd > dem.h.nrRows) return fail(ctx, PB_ERR_INVALID, "bad row band");
    int rc = ensureValidList(ctx, dem);
    if (rc) return rc;
    const size_t N = dem.n;
    CUDA_TRY(ctx, ctx->dOut.reserve(2 * N));
    ctx->outInitRaster = -1;
    float* gEst = ctx->dOut.p;
    float* gRm = rmse ? ctx->dOut.p + N : nullptr;
    // both grids start as the DEM's flag
    CUDA_TRY(ctx, launchFill(gEst, (long long)(rmse ? 2 * N : N), dem.h.flag, ctx->stream));
    long long t0 = 0, t1 = dem.nValid;
    if (rowBegin != 0 || rowEnd != dem.h.nrRows) {
        rc = ensureRowStart(ctx, dem);
        if (rc) return rc;
        t0 = dem.rowStart[rowBegin];
        t1 = dem.rowStart[rowEnd];
    }
    KrigingTargets t{};
    t.validIdx = dem.validIdx + t0;
    t.grid.data = dem.d;
    t.grid.nrRows = dem.h.nrRows; t.grid.nrCols = dem.h.nrCols;
    t.grid.llx = dem.h.llx; t.grid.lly = dem.h.lly; t.grid.cellSize = dem.h.cellSize; t.grid.invCellSize = dem.h.invCellSize;
    t.grid.flag = dem.h.flag;
    // the rows of both grids come down chunk by chunk on a second stream while the following chunks are evaluated: a chunk's
    // rows are copied once the NEXT chunk's kernels are in the compute stream (a copy into pageable memory blocks the host)
    pb_raster_out* outs[2] = {estimate, rmse};
    float* grids[2] = {gEst, gRm};
    const int cols = dem.h.nrCols;
    const bool toHost = (estimate->data || estimate->rows) || (rmse && (rmse->data || rmse->rows));
    if (toHost) {
        rc = ensureRowStart(ctx, dem);
        if (rc) return rc;
        if (!k->sDn) {
            CUDA_TRY(ctx, cudaStreamCreateWithFlags(&k->sDn, cudaStreamNonBlocking));
            for (auto& e : k->evDn) CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
    }
    auto copyRows = [&](int ra, int rb) -> int {
        if (ra >= rb) return PB_OK;
        for (int g = 0; g < 2; g++) {
            pb_raster_out* o = outs[g];
            if (!o) continue;
            if (o->data) {
                CUDA_TRY(ctx, cudaMemcpyAsync(o->data + size_t(ra) * cols, grids[g] + size_t(ra) * cols,
                                              sizeof(float) * size_t(rb - ra) * cols, cudaMemcpyDeviceToHost, k->sDn));
            } else if (o->rows) {
                for (int r = ra; r < rb; r++)
                    CUDA_TRY(ctx, cudaMemcpyAsync(o->rows[r], grids[g] + size_t(r) * cols, sizeof(float) * cols,
                                                  cudaMemcpyDeviceToHost, k->sDn));
            } else continue;
            ctx->timing.d2h_bytes += (int64_t)(sizeof(float) * size_t(rb - ra) * cols);
        }
        return PB_OK;
    };
    int rowCopied = rowBegin, rowPending = rowBegin, nChunk = 0;       // rows [rowCopied, rowPending) wait for evDn[(nChunk - 1) & 1]
    auto afterChunk = [&](long long done) -> int {
        // rows whose cells are all among the first `done` targets of the band
        const long long* rs = dem.rowStart.data();
        int rTo = int(std::upper_bound(rs + rowPending, rs + rowEnd + 1, t0 + done) - rs) - 1;
        if (t0 + done >= t1) rTo = rowEnd;
        if (rowPending > rowCopied) {           // the previous chunk's rows: its event was recorded before this chunk's kernels
            CUDA_TRY(ctx, cudaStreamWaitEvent(k->sDn, k->evDn[(nChunk - 1) & 1], 0));
            const int rc2 = copyRows(rowCopied, rowPending);
            if (rc2) return rc2;
            rowCopied = rowPending;
        }
        CUDA_TRY(ctx, cudaEventRecord(k->evDn[nChunk & 1], ctx->stream));
        rowPending = std::max(rowPending, rTo);
        nChunk++;
        return PB_OK;
    };
    cudaEventRecord(ctx->ev[1], ctx->stream);
    rc = evaluate(ctx, k, t, t1 - t0, rmse != nullptr, nullptr, nullptr, gEst, gRm,
                  toHost ? std::function<int(long long)>(afterChunk) : std::function<int(long long)>());
    cudaEventRecord(ctx->ev[2], ctx->stream);
    if (rc) { if (k->sDn) cudaStreamSynchronize(k->sDn); return rc; }
    for (int g = 0; g < 2; g++)
        if (outs[g]) outs[g]->nValid = t1 - t0;
    if (toHost) {
        if (nChunk > 0) CUDA_TRY(ctx, cudaStreamWaitEvent(k->sDn, k->evDn[(nChunk - 1) & 1], 0));
        else CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));          // no valid cell: only the fill kernel ran
        rc = copyRows(rowCopied, rowEnd);
        if (rc) return rc;
        CUDA_TRY(ctx, cudaStreamSynchronize(k->sDn));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->timing.kernel_ms, ctx->ev[1], ctx->ev[2]);
    ctx->timing.total_ms = ctx->timing.kernel_ms;
    return PB_OK;
}

}  // extern "C"
