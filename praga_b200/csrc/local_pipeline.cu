// local_pipeline.cu — K2 as a three-kernel pipeline over a chunk of DEM cells (the raster form of local detrending):
//
//   A  local_select_kernel    per cell: localSelection (interpolation.cpp:1087-1170) and the staging of the elevation fit
//                             (multipleDetrendingElevationFitting :1862-1953, first half) into a per-cell RECORD in HBM
//   B  local_descent_kernel   every (cell, first guess) pair is ONE Levenberg-Marquardt descent (bestFittingMarquardt_nDimension
//                             mathFunctions/furtherMathFunctions.cpp:1360-1427 over fittingMarquardt_nDimension :1954-2118):
//                             a persistent CTA per SM keeps a ring of cell SLOTS in shared memory (producer warps copy the
//                             records in and pre-compute each first guess's starting SSE / scores), every consumer LANE pulls
//                             descents from a queue and runs them one Levenberg-Marquardt iteration per loop trip, scoring fused
//                             into the iteration: lanes never wait for the longest descent of a cell, a group or a warp
//   C  local_finish_kernel    per cell: the reference's pick over the first guesses (R2 > best + 0.005 in guess order), the rest
//                             of multipleDetrendingMain (detrendingElevation, other proxies' fit and detrending), interpolate
//                             (idw / modified Shepard) + retrend + clamp (:2502-2560)
//
// The arithmetic is the fused kernel's (same device functions, same order of every sum): the rasters are bit-identical
// (tests/test_local_detrending.py compares both forms). What changes is the schedule: in the fused kernel a group of 4 lanes
// owned a cell from selection to estimate, and 18 of 32 lanes were busy because descents end after 2-60 iterations.
// Cells whose selection does not fit a record (more than kRecCap stations, or more than kFitCap elevation points) are listed
// and gridded by the fused kernel afterwards. Built with -fmad=false.
#include <stdlib.h>
#include <algorithm>
#include "local_device.cuh"

namespace pb {

constexpr int kRecCap = 64;         // selected stations a record holds
constexpr int kFitCap = 48;         // elevation-fit points a descent slot holds
constexpr int kResStride = 8;       // doubles per first guess in a record's result area: N <= 6 parameters, ssr, sc

struct CellHeader {                 // 32 bytes at the start of a record
    int n, nE, flags, cell;
    float localRadius, xf, yf;
    int pad;
};
enum { CELL_FIT = 1, CELL_OVERFLOW = 2 };

// record = header | work arrays (WorkPtrs layout, capacity kRecCap) | per first guess {sse0, ssr0, sc0} (32 guesses) | results
constexpr int kMaxGuess = 32;
__host__ __device__ constexpr size_t recWorkOff() { return 32; }
__host__ __device__ constexpr size_t recInitOff() { return 32 + workBytes(kRecCap); }
__host__ __device__ constexpr size_t recResOff() { return recInitOff() + kMaxGuess * 3 * 8; }
size_t localRecordStride(int nGuess) { return (recResOff() + size_t(nGuess) * kResStride * 8 + 255) / 256 * 256; }

// lmInit (furtherMathFunctions.cpp:1954-1975: the SSE of the first guess) of one first guess + the curve-dependent sums of
// computeWeighted_R2 / _RMSE at that starting point (a descent that never accepts a step ends on it): same sums, same order as
// lmInit / weightedScoresWith
template <int F>
__device__ __forceinline__ void descentInit(const double* __restrict__ guess, const double* X, const double* Y, const double* W,
                                            int nE, double* __restrict__ init)
{
    constexpr int N = Fit<F>::N;
    double p[N];
#pragma unroll
    for (int k = 0; k < N; k++) p[k] = guess[k];
    if (nE > 0) Fit<F>::clamp(p);
    double sse = 0, ssr = 0, sc = 0;
    for (int i = 0; i < nE; i++) {
        const double error = Y[i] - Fit<F>::eval(X[i], p);
        const double w = W[i];
        sse += error * error * w * w;
        const double r = w * error;
        ssr += r * r;
        sc += r * error;
    }
    init[0] = sse;
    init[1] = ssr;
    init[2] = sc;
}

// ---------------------------------------------------------------------------------------------------------------------
// A: selection + staging. One group of 4 lanes per cell, many cells in flight per SM (no LM code: few registers).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kSelG = 4;
constexpr int kSelThreads = 256;

// candidate / selection scratch of one group: only the arrays localSelect touches (candidate list oi / w, selection idx / d,
// and val / flg, which it also fills when a selection is too large for the record)
__host__ __device__ constexpr size_t selScratchBytes(int cap) { return (size_t(cap) * 21 + 15) / 16 * 16; }
__device__ __forceinline__ WorkPtrs carveSel(unsigned char* base, int cap)
{
    WorkPtrs p{};
    p.idx = reinterpret_cast<int*>(base);
    p.d = reinterpret_cast<float*>(p.idx + cap);
    p.oi = reinterpret_cast<int*>(p.d + cap);
    p.w = reinterpret_cast<float*>(p.oi + cap);
    p.val = p.w + cap;
    p.flg = reinterpret_cast<unsigned char*>(p.val + cap);
    return p;
}

__global__ void __launch_bounds__(kSelThreads, 3)
local_select_kernel(const LocalModel* __restrict__ mp, LocalStations st, DevGrid dem, const int* __restrict__ validIdx, int nCells,
                    unsigned char* __restrict__ selScratch, int selCap, int* __restrict__ errFlag, int xyInSmem,
                    unsigned char* __restrict__ records, size_t recStride, int* __restrict__ overflowList,
                    int* __restrict__ overflowCount)
{
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int G = kSelG;
    constexpr int groupsPerCta = kSelThreads / G;
    const LocalModel& m = *mp;
    const int grp = threadIdx.x / G;
    const int lane = grpLane<G>();
    float2* sxy = reinterpret_cast<float2*>(smem);
    if (xyInSmem) {
        for (int i = threadIdx.x; i < st.n; i += kSelThreads) sxy[i] = st.xy[i];
        __syncthreads();
    }
    const float2* xy = xyInSmem ? sxy : st.xy;
    const long long groupsTotal = (long long)gridDim.x * groupsPerCta;
    const long long gid = (long long)blockIdx.x * groupsPerCta + grp;
    const WorkPtrs gw = carveSel(selScratch + gid * selScratchBytes(selCap), selCap);
    const int ePos = m.elevationPos;
    const bool hasElev = ePos >= 0 && m.selectedActive[ePos];

    for (long long t = gid; t < nCells; t += groupsTotal) {
        const int cell = validIdx[t];
        const int row = cell / dem.nrCols, col = cell - row * dem.nrCols;
        // gis::getUtmXYFromRowCol (gis.cpp:806-810), then narrowed to float by the callee signatures (interpolation.h:77-95,162)
        const double x = dem.llx + dem.cellSize * (col + 0.5);
        const double y = dem.lly + dem.cellSize * (dem.nrRows - row - 0.5);
        const float xf = float(x), yf = float(y);
        unsigned char* rec = records + size_t(t) * recStride;
        WorkPtrs w;
        float localRadius = 0.f;
        // the record's work area plays the role of the fused kernel's shared-memory tile (capacity kRecCap)
        const int n = localSelect<G>(m, st, xy, xf, yf, gw, selCap, rec + recWorkOff(), kRecCap, errFlag, w, localRadius);
        int nE = 0, flags = 0;
        if (n > kRecCap) flags = CELL_OVERFLOW;
        else if (hasElev) {
            bool isValid;
            nE = mdStageElevation<G>(m, st, w, n, true, isValid);
            if (isValid) flags = (nE <= kFitCap) ? CELL_FIT : CELL_OVERFLOW;
            if (flags == CELL_FIT) {
                // starting point of every descent of this cell (the lanes share the first guesses)
                double* init = reinterpret_cast<double*>(rec + recInitOff());
                const int nGuess = m.nFirstGuess < kMaxGuess ? m.nFirstGuess : kMaxGuess;
                for (int g = lane; g < nGuess; g += G) {
                    const double* guess = &m.firstGuess[g][0];
                    switch (m.elevFunction) {
                    case PB_FIT_PIECEWISE_THREE: descentInit<PB_FIT_PIECEWISE_THREE>(guess, w.X, w.Y, w.W, nE, init + g * 3); break;
                    case PB_FIT_PIECEWISE_THREE_FREE: descentInit<PB_FIT_PIECEWISE_THREE_FREE>(guess, w.X, w.Y, w.W, nE, init + g * 3); break;
                    default: descentInit<PB_FIT_PIECEWISE_TWO>(guess, w.X, w.Y, w.W, nE, init + g * 3); break;
                    }
                }
            }
        }
        if (lane == 0) {
            CellHeader h;
            h.n = n; h.nE = nE; h.flags = flags; h.cell = cell;
            h.localRadius = localRadius; h.xf = xf; h.yf = yf; h.pad = 0;
            *reinterpret_cast<CellHeader*>(rec) = h;
            if (flags & CELL_OVERFLOW) overflowList[atomicAdd(overflowCount, 1)] = cell;
        }
        grpSync<G>();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// B: the descents.
// ---------------------------------------------------------------------------------------------------------------------
template <int F> struct Pipe {
    static constexpr int N = Fit<F>::N;
    static constexpr int NG = (F == PB_FIT_PIECEWISE_TWO) ? 16 : 32;            // first guesses per cell (16, or 30 / 31)
    static constexpr int NSLOT = (F == PB_FIT_PIECEWISE_TWO) ? 128 : 96;         // cells resident in shared memory per CTA
    static constexpr int RING = (F == PB_FIT_PIECEWISE_TWO) ? 2048 : 4096;       // >= NSLOT * NG, power of two
    static constexpr int FREE = 128;                                             // >= NSLOT, power of two
    // slot: X, Y, W (kFitCap doubles each), per first guess {sse0, ssr0, sc0}, header {cell, nE, remaining, pad}; the stride in
    // 8-byte units is odd so that lanes working on different slots spread over all shared-memory banks
    static constexpr int offInit = 3 * kFitCap * 8;
    static constexpr int offHdr = offInit + NG * 3 * 8;
    static constexpr int slotBytesRaw = offHdr + 16;
    static constexpr int slotBytes = ((slotBytesRaw / 8) % 2 == 1) ? slotBytesRaw : slotBytesRaw + 8;
    static constexpr int offRing = NSLOT * slotBytes;
    static constexpr int offFree = offRing + RING * 4;
    static constexpr int offCtl = offFree + FREE * 4;
    static constexpr int offModel = offCtl + 64;                                  // pmin, pmax, delta (3 x 8 doubles) + first guesses
    static constexpr int smemBytes = offModel + 3 * 8 * 8 + NG * PB_MAX_FIT_PARAMS * 8;
};

struct PipeCtl {                    // shared-memory control block
    unsigned qReserve;              // item positions handed to producers
    unsigned qHead;                 // tickets handed to consumer lanes
    unsigned fTail, fHead;          // free-slot ring
    unsigned prodDone;              // producer warps that ran out of cells
    unsigned pad[3];
};

__device__ __forceinline__ unsigned ldVolatile(const unsigned* p) { return *reinterpret_cast<const volatile unsigned*>(p); }
__device__ __forceinline__ void stVolatile(unsigned* p, unsigned v) { *reinterpret_cast<volatile unsigned*>(p) = v; }

// registers per thread: what ceil(warps / 4) warps leave each other in the 16 K registers of one SM sub-partition (unit: 8)
__host__ __device__ constexpr int descentRegs(int warps)
{
    return (16384 / (((warps + 3) / 4) * 32)) / 8 * 8 > 255 ? 255 : (16384 / (((warps + 3) / 4) * 32)) / 8 * 8;
}

template <int F, int NCONS, int NPROD>
__global__ void __maxnreg__(descentRegs(NCONS + NPROD))
local_descent_kernel(const LocalModel* __restrict__ mp, unsigned char* __restrict__ records, size_t recStride, int nCells,
                     unsigned* __restrict__ cellCounter)
{
    using P = Pipe<F>;
    constexpr int N = P::N;
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned* ring = reinterpret_cast<unsigned*>(smem + P::offRing);
    unsigned* fring = reinterpret_cast<unsigned*>(smem + P::offFree);
    PipeCtl* ctl = reinterpret_cast<PipeCtl*>(smem + P::offCtl);
    double* sModel = reinterpret_cast<double*>(smem + P::offModel);          // pmin[8], pmax[8], delta[8], firstGuess[NG][6]
    const LocalModel& m = *mp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nGuess = m.nFirstGuess < P::NG ? m.nFirstGuess : P::NG;

    for (int i = tid; i < P::RING; i += blockDim.x) ring[i] = 0u;
    for (int i = tid; i < P::FREE; i += blockDim.x) fring[i] = (i < P::NSLOT) ? ((1u << 16) | unsigned(i)) : 0u;
    if (tid == 0) {
        ctl->qReserve = 0; ctl->qHead = 0; ctl->fTail = P::NSLOT; ctl->fHead = 0; ctl->prodDone = 0;
    }
    if (tid < 8) {
        sModel[tid] = tid < N ? m.elevMin[tid] : 0.;
        sModel[8 + tid] = tid < N ? m.elevMax[tid] : 0.;
        sModel[16 + tid] = tid < N ? m.elevDelta[tid] : 1.;
    }
    for (int i = tid; i < nGuess * PB_MAX_FIT_PARAMS; i += blockDim.x) sModel[24 + i] = (&m.firstGuess[0][0])[i];
    __syncthreads();
    const double* pmin = sModel;
    const double* pmax = sModel + 8;
    const double* firstGuess = sModel + 24;

    if (warp >= NCONS) {
        // ------------------------------------------------------------------ producer warp
        for (;;) {
            unsigned t = 0;
            if (lane == 0) t = atomicAdd(cellCounter, 1u);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= unsigned(nCells)) break;
            unsigned char* rec = records + size_t(t) * recStride;
            const CellHeader h = *reinterpret_cast<const CellHeader*>(rec);
            if (!(h.flags & CELL_FIT) || (h.flags & CELL_OVERFLOW)) continue;
            // a free slot
            unsigned fp = 0;
            if (lane == 0) fp = atomicAdd(&ctl->fHead, 1u);
            fp = __shfl_sync(0xffffffffu, fp, 0);
            const unsigned wantTag = (fp / P::FREE + 1u) & 0xffffu;
            unsigned fe;
            while (((fe = ldVolatile(&fring[fp & (P::FREE - 1)])) >> 16) != wantTag) __nanosleep(64);
            const int slot = int(fe & 0xffffu);
            __threadfence_block();
            unsigned char* sb = smem + size_t(slot) * P::slotBytes;
            double* sX = reinterpret_cast<double*>(sb);
            double* sY = sX + kFitCap;
            double* sW = sY + kFitCap;
            double* sInit = reinterpret_cast<double*>(sb + P::offInit);
            int* sHdr = reinterpret_cast<int*>(sb + P::offHdr);
            const WorkPtrs rw = carve(rec + recWorkOff(), kRecCap);
            const int nE = h.nE;
            for (int j = lane; j < nE; j += 32) { sX[j] = rw.X[j]; sY[j] = rw.Y[j]; sW[j] = rw.W[j]; }
            if (lane == 0) { sHdr[0] = int(t); sHdr[1] = nE; sHdr[2] = nGuess; sHdr[3] = 0; }
            {
                const double* rInit = reinterpret_cast<const double*>(rec + recInitOff());
                for (int i = lane; i < nGuess * 3; i += 32) sInit[i] = rInit[i];
            }
            __syncwarp();
            __threadfence_block();
            unsigned pos = 0;
            if (lane == 0) pos = atomicAdd(&ctl->qReserve, unsigned(nGuess));
            pos = __shfl_sync(0xffffffffu, pos, 0);
            for (int g = lane; g < nGuess; g += 32) {
                const unsigned q = pos + unsigned(g);
                unsigned* e = &ring[q & (P::RING - 1)];
                while (ldVolatile(e) != 0u) __nanosleep(64);          // the entry of the previous lap has been read
                stVolatile(e, (((q / P::RING + 1u) & 0xffffu) << 16) | (unsigned(slot) << 5) | unsigned(g));
            }
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            atomicAdd(&ctl->prodDone, 1u);
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer lanes
    double dl[N], rdl[N];
#pragma unroll
    for (int k = 0; k < N; k++) { dl[k] = sModel[16 + k]; rdl[k] = 1.0 / dl[k]; }
    LmState<F> s;
    bool have = false;
    int ticket = -1, slot = 0, guess = 0, nData = 0;
    const double *X = nullptr, *Y = nullptr, *W = nullptr;
    for (;;) {
        if (!have) {
            if (ticket < 0) ticket = int(atomicAdd(&ctl->qHead, 1u));
            unsigned* ep = &ring[unsigned(ticket) & (P::RING - 1)];
            const unsigned e = ldVolatile(ep);
            if ((e >> 16) == ((unsigned(ticket) / P::RING + 1u) & 0xffffu)) {
                __threadfence_block();
                stVolatile(ep, 0u);                                   // consumed: the position may be written again
                slot = int((e >> 5) & 0x7ffu);
                guess = int(e & 31u);
                unsigned char* sb = smem + size_t(slot) * P::slotBytes;
                X = reinterpret_cast<const double*>(sb);
                Y = X + kFitCap;
                W = Y + kFitCap;
                const double* sInit = reinterpret_cast<const double*>(sb + P::offInit);
                nData = reinterpret_cast<const int*>(sb + P::offHdr)[1];
#pragma unroll
                for (int k = 0; k < N; k++) { s.lambda[k] = 0.01; s.p[k] = firstGuess[guess * PB_MAX_FIT_PARAMS + k]; }
                if (nData > 0) Fit<F>::clamp(s.p);
                s.mySSE = sInit[guess * 3 + 0];
                s.ssr = sInit[guess * 3 + 1];
                s.sc = sInit[guess * 3 + 2];
                s.iter = 0;
                have = true;
                ticket = -1;
            }
        }
        if (!__any_sync(0xffffffffu, have)) {
            // nobody in this warp holds a descent: leave once the producers are done and no ticket of ours can still be served
            const bool prodDone = ldVolatile(&ctl->prodDone) == unsigned(NPROD);
            __threadfence_block();
            const unsigned tail = ldVolatile(&ctl->qReserve);
            const bool mine = prodDone && unsigned(ticket) >= tail;
            if (__all_sync(0xffffffffu, mine)) break;
            __nanosleep(128);
            continue;
        }
        if (have) {
            if (lmIter<F, true>(s, pmin, pmax, dl, rdl, 1000, 0.002, X, Y, W, nData)) {
                unsigned char* sb = smem + size_t(slot) * P::slotBytes;
                int* sHdr = reinterpret_cast<int*>(sb + P::offHdr);
                double* res = reinterpret_cast<double*>(records + size_t(sHdr[0]) * recStride + recResOff()) + guess * kResStride;
#pragma unroll
                for (int k = 0; k < N; k++) res[k] = s.p[k];
                res[6] = s.ssr;
                res[7] = s.sc;
                __threadfence_block();
                if (atomicSub(reinterpret_cast<unsigned*>(&sHdr[2]), 1u) == 1u) {      // last descent of the cell: the slot is free
                    const unsigned fp = atomicAdd(&ctl->fTail, 1u);
                    stVolatile(&fring[fp & (P::FREE - 1)], (((fp / P::FREE + 1u) & 0xffffu) << 16) | unsigned(slot));
                }
                have = false;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// C: pick + the rest of multipleDetrendingMain + estimate. One group of 4 lanes per cell working on the cell's record.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kFinThreads = 256;

template <int F>
__global__ void __launch_bounds__(kFinThreads)
local_finish_kernel(const LocalModel* __restrict__ mp, LocalStations st, DevGrid dem, unsigned char* __restrict__ records,
                    size_t recStride, int nCells, float* __restrict__ out, unsigned* __restrict__ minmax)
{
    constexpr int G = kSelG;
    constexpr int N = Fit<F>::N;
    constexpr int groupsPerCta = kFinThreads / G;
    const LocalModel& m = *mp;
    const int grp = threadIdx.x / G;
    const int lane = grpLane<G>();
    const long long groupsTotal = (long long)gridDim.x * groupsPerCta;
    const long long gid = (long long)blockIdx.x * groupsPerCta + grp;
    const int nGuess = m.nFirstGuess;
    const double deltaR2 = 0.005;

    float vmin = INFINITY, vmax = -INFINITY;
    for (long long t = gid; t < nCells; t += groupsTotal) {
        unsigned char* rec = records + size_t(t) * recStride;
        const CellHeader h = *reinterpret_cast<const CellHeader*>(rec);
        if (h.flags & CELL_OVERFLOW) continue;                         // gridded by the fused kernel
        const WorkPtrs w = carve(rec + recWorkOff(), kRecCap);
        const int n = h.n;
        double elevPar[N];
#pragma unroll
        for (int j = 0; j < N; j++) elevPar[j] = 0;
        double R2 = 0;
        const bool fitRan = (h.flags & CELL_FIT) != 0;
        if (fitRan) {
            // bestFittingMarquardt_nDimension's pick (:1395-1424) over the descents' results, in guess order
            const int nE = h.nE;
            const ScoreBase base = weightedScoreBase(w.Y, w.W, nE);
            const double* res = reinterpret_cast<const double*>(rec + recResOff());
            double best[N];
#pragma unroll
            for (int j = 0; j < N; j++) best[j] = 0;
            double bestR2 = -9999, bestRMSE = -9999;
            int rmseIndex = -9999;
            for (int kk = 0; kk < nGuess; kk++) {
                const double* r = res + kk * kResStride;
                const double r2 = 1.0 - (r[6] / base.sst);
                const double rmse = sqrt(r[7] / (base.sw * nE));
                if (isEqualD(bestRMSE, -9999) || rmse < bestRMSE) { bestRMSE = rmse; rmseIndex = kk; }
                if (isEqualD(bestR2, -9999) || r2 > (bestR2 + deltaR2)) {
#pragma unroll
                    for (int j = 0; j < N; j++) best[j] = r[j];
                    bestR2 = r2;
                }
            }
#pragma unroll
            for (int j = 0; j < N; j++) elevPar[j] = best[j];
            if (bestR2 < 0 && rmseIndex >= 0) {
                // the reference refits from the best-RMSE first guess (:1417-1423): that descent is deterministic and its
                // result is already in the record (the parameters a descent ends on are the clamped ones it was scored with)
                const double* r = res + rmseIndex * kResStride;
#pragma unroll
                for (int j = 0; j < N; j++) elevPar[j] = r[j];
            }
            R2 = bestR2;
        }
        grpSync<G>();
        MdFit<N> fit;
        mdFinish<G, F>(m, st, w, n, true, fitRan, R2, elevPar, fit);
        double pv[PB_MAX_PROXY];
        for (int i = 0; i < m.nProxy; i++) {
            pv[i] = double(kNoData);
            if (m.selectedActive[i] && m.gridSlot[i] >= 0) {
                const DevGrid& g = m.grid[m.gridSlot[i]];
                const float v = gridValueXY(g, double(h.xf), double(h.yf));
                if (v != g.flag) pv[i] = double(v);
            }
        }
        const float o = localFinish<G, F>(m, st, w, n, pv, fit, true, h.xf, h.yf, h.localRadius);
        if (lane == 0) {
            out[h.cell] = o;
            if (!isEqualF(o, dem.flag) && !isEqualF(o, kNoData)) {
                vmin = fminf(vmin, o);
                vmax = fmaxf(vmax, o);
            }
        }
    }
    if (lane == 0 && minmax) {
        if (vmin != INFINITY) atomicMin(&minmax[0], orderedKeyL(vmin));
        if (vmax != -INFINITY) atomicMax(&minmax[1], orderedKeyL(vmax));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
static int smCount()
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

int localSelectGroups() { return smCount() * 4 * (kSelThreads / kSelG); }
size_t localSelectScratchBytes(int cap) { return selScratchBytes(cap) * size_t(localSelectGroups()); }

template <int F>
static cudaError_t launchDescent(const LocalModel* dModel, unsigned char* records, size_t recStride, int nCells, unsigned* cellCounter,
                                 cudaStream_t s)
{
    // registers bound the CTA: each of the four SM sub-partitions owns 16 K registers and holds whole warps, so 12 warps (3 per
    // sub-partition) may use 168 registers each; the two-piece descent needs 146, the three-piece ones 182 (a few spills at 168)
    constexpr int NPROD = 1;
    constexpr int NCONS = 11;
    auto k = local_descent_kernel<F, NCONS, NPROD>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Pipe<F>::smemBytes);
    if (e != cudaSuccess) return e;
    const int grid = std::min(smCount(), std::max(1, nCells / 8));
    k<<<grid, (NCONS + NPROD) * 32, Pipe<F>::smemBytes, s>>>(dModel, records, recStride, nCells, cellCounter);
    return cudaGetLastError();
}

template <int F>
static cudaError_t launchFinish(const LocalModel* dModel, const LocalStations& st, const DevGrid& dem, unsigned char* records,
                                size_t recStride, int nCells, float* out, unsigned* minmax, cudaStream_t s)
{
    const int groupsPerCta = kFinThreads / kSelG;
    const long long need = (nCells + groupsPerCta - 1) / groupsPerCta;
    const int grid = int(std::min<long long>(need, (long long)smCount() * 8));
    local_finish_kernel<F><<<grid, kFinThreads, 0, s>>>(dModel, st, dem, records, recStride, nCells, out, minmax);
    return cudaGetLastError();
}

// one chunk of cells [validIdx, validIdx + nCells) through A, B, C. records: nCells * recStride bytes; selScratch:
// localSelectScratchBytes(selCap); counters: {cellCounter, overflowCount} zeroed here; overflowList: nCells ints.
cudaError_t launchLocalPipelineChunk(const LocalModel* dModel, const LocalStations& st, const DevGrid& dem, const int* validIdx,
                                     int nCells, int elevFunction, int nFirstGuess, unsigned char* records, size_t recStride,
                                     unsigned char* selScratch, int selCap, int* errFlag, unsigned* counters, int* overflowList,
                                     float* out, unsigned* minmax, cudaStream_t s, int* launches)
{
    if (nCells <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned), s);
    if (e != cudaSuccess) return e;
    {
        size_t smem = 0;
        int xyInSmem = 0;
        if (size_t(st.n) * sizeof(float2) <= 96 * 1024) { xyInSmem = 1; smem = size_t(st.n) * sizeof(float2); }
        e = cudaFuncSetAttribute(local_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024));
        if (e != cudaSuccess) return e;
        const int groupsPerCta = kSelThreads / kSelG;
        const long long need = (nCells + groupsPerCta - 1) / groupsPerCta;
        const int grid = int(std::min<long long>(need, (long long)smCount() * 4));
        local_select_kernel<<<grid, kSelThreads, smem, s>>>(dModel, st, dem, validIdx, nCells, selScratch, selCap, errFlag, xyInSmem,
                                                            records, recStride, overflowList, reinterpret_cast<int*>(counters + 1));
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    (void)nFirstGuess;
    switch (elevFunction) {
    case PB_FIT_PIECEWISE_THREE:
        e = launchDescent<PB_FIT_PIECEWISE_THREE>(dModel, records, recStride, nCells, counters, s);
        if (e == cudaSuccess) e = launchFinish<PB_FIT_PIECEWISE_THREE>(dModel, st, dem, records, recStride, nCells, out, minmax, s);
        break;
    case PB_FIT_PIECEWISE_THREE_FREE:
        e = launchDescent<PB_FIT_PIECEWISE_THREE_FREE>(dModel, records, recStride, nCells, counters, s);
        if (e == cudaSuccess) e = launchFinish<PB_FIT_PIECEWISE_THREE_FREE>(dModel, st, dem, records, recStride, nCells, out, minmax, s);
        break;
    default:
        e = launchDescent<PB_FIT_PIECEWISE_TWO>(dModel, records, recStride, nCells, counters, s);
        if (e == cudaSuccess) e = launchFinish<PB_FIT_PIECEWISE_TWO>(dModel, st, dem, records, recStride, nCells, out, minmax, s);
        break;
    }
    if (launches) *launches += 3;
    return e;
}

}  // namespace pb
