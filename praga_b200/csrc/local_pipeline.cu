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
#include <stdio.h>
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
constexpr int kSelThreads = 128;     // 32 cells in flight per CTA; 5 CTAs per SM (shared memory)
constexpr int kSelSpanTilesMax = 8;  // tiles (of kSelThreads / kSelG consecutive targets) that share one prefilter list, at most
constexpr int kSelListCap = 1024;    // stations the CTA-wide prefilter list holds (beyond: the groups scan all stations)
constexpr int kCandCap = 80;         // per-cell candidate list (shared memory); denser neighbourhoods scan the prefilter list per ring

// shared-memory scratch of one group: candidate list (oi, w: kCandCap) + selection (idx, d: kRecCap)
constexpr int kSelGroupBytes = kCandCap * 8 + kRecCap * 8;
__device__ __forceinline__ WorkPtrs carveSel(unsigned char* base)
{
    WorkPtrs p{};
    p.oi = reinterpret_cast<int*>(base);
    p.w = reinterpret_cast<float*>(p.oi + kCandCap);
    p.idx = reinterpret_cast<int*>(p.w + kCandCap);
    p.d = reinterpret_cast<float*>(p.idx + kRecCap);
    return p;
}

__device__ __forceinline__ unsigned orderedKeyF(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float keyToFloatF(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void __launch_bounds__(kSelThreads, 5)
local_select_kernel(const LocalModel* __restrict__ mp, LocalStations st, DevGrid dem, const int* __restrict__ validIdx, int nCells,
                    unsigned char* __restrict__ records, size_t recStride, int* __restrict__ overflowList,
                    int* __restrict__ overflowCount, unsigned* __restrict__ fitHist, int spanTiles, unsigned* __restrict__ spanCounter)
{
    __shared__ __align__(16) unsigned char sGroup[(kSelThreads / kSelG) * kSelGroupBytes];
    __shared__ int sList[kSelListCap];
    __shared__ unsigned sBox[4];            // ordered keys of min x, max x, min y, max y of the span's targets
    __shared__ int sWarpCount[kSelThreads / 32];
    __shared__ unsigned sSpan;
    constexpr int G = kSelG;
    constexpr int groupsPerCta = kSelThreads / G;
    const int spanCells = groupsPerCta * spanTiles;
    const LocalModel& m = *mp;
    const int grp = threadIdx.x / G;
    const int lane = grpLane<G>();
    const int warp = threadIdx.x >> 5, wl = threadIdx.x & 31;
    const float2* xy = st.xy;
    const WorkPtrs gw = carveSel(sGroup + grp * kSelGroupBytes);
    const int ePos = m.elevationPos;
    const bool hasElev = ePos >= 0 && m.selectedActive[ePos];
    const int S = st.n;
    // the prefilter pays when the candidate radius covers a small part of the station set
    const bool prefilter = m.candRadius > 0.f && unsigned(S) > m.minPoints && S > 256;

    // a SPAN = spanTiles x groupsPerCta consecutive targets of the valid list (neighbouring cells of a row, a few rows at
    // most): their candidate circles overlap almost entirely, so the CTA first lists, in station order, the stations inside the
    // span's bounding box grown by the candidate radius (a superset of every target's candidates); each group's candidate pass
    // then walks that list instead of all S stations. One barrier per span: the groups run their cells independently.
    // the list reaches twice the candidate radius: the rings of targets with few stations around them (raster borders) go on
    // beyond the candidate radius and would otherwise scan all S stations per ring
    const float listRadius = 2.f * m.candRadius;
    for (;;) {
        __syncthreads();                     // every group is done with the previous span's list
        if (threadIdx.x == 0) sSpan = atomicAdd(spanCounter, 1u);          // spans are handed out dynamically (uneven cost)
        __syncthreads();
        const long long first = (long long)sSpan * spanCells;
        if (first >= nCells) break;
        int nList = 0;
        const int* list = nullptr;
        if (prefilter) {
            if (threadIdx.x == 0) { sBox[0] = 0xffffffffu; sBox[1] = 0u; sBox[2] = 0xffffffffu; sBox[3] = 0u; }
            __syncthreads();
            {
                float bx0 = INFINITY, bx1 = -INFINITY, by0 = INFINITY, by1 = -INFINITY;
                for (long long t = first + threadIdx.x; t < first + spanCells && t < nCells; t += kSelThreads) {
                    const int cell = validIdx[t];
                    const int row = cell / dem.nrCols, col = cell - row * dem.nrCols;
                    const float xf = float(dem.llx + dem.cellSize * (col + 0.5));
                    const float yf = float(dem.lly + dem.cellSize * (dem.nrRows - row - 0.5));
                    bx0 = fminf(bx0, xf); bx1 = fmaxf(bx1, xf); by0 = fminf(by0, yf); by1 = fmaxf(by1, yf);
                }
                for (int o = 16; o; o >>= 1) {
                    bx0 = fminf(bx0, __shfl_xor_sync(0xffffffffu, bx0, o)); bx1 = fmaxf(bx1, __shfl_xor_sync(0xffffffffu, bx1, o));
                    by0 = fminf(by0, __shfl_xor_sync(0xffffffffu, by0, o)); by1 = fmaxf(by1, __shfl_xor_sync(0xffffffffu, by1, o));
                }
                if (wl == 0 && bx0 != INFINITY) {
                    atomicMin(&sBox[0], orderedKeyF(bx0)); atomicMax(&sBox[1], orderedKeyF(bx1));
                    atomicMin(&sBox[2], orderedKeyF(by0)); atomicMax(&sBox[3], orderedKeyF(by1));
                }
            }
            __syncthreads();
            // grown by more than the list radius (the per-target tests are d2 <= rcap^2 * 1.0001 and d <= r1 <= listRadius): a
            // superset for sure
            const float R = listRadius * 1.001f + 1.f;
            const float x0 = keyToFloatF(sBox[0]) - R, x1 = keyToFloatF(sBox[1]) + R;
            const float y0 = keyToFloatF(sBox[2]) - R, y1 = keyToFloatF(sBox[3]) + R;
            // every warp owns one contiguous range of stations: count, prefix over the warps, write — station order is kept
            const int per = ((S + kSelThreads / 32 - 1) / (kSelThreads / 32) + 31) / 32 * 32;
            const int w0 = warp * per, w1 = min(S, w0 + per);
            int cnt = 0;
            for (int i = w0 + wl; i - wl < w1; i += 32) {
                bool in = false;
                if (i < w1) { const float2 p = xy[i]; in = p.x >= x0 && p.x <= x1 && p.y >= y0 && p.y <= y1; }
                cnt += __popc(__ballot_sync(0xffffffffu, in));
            }
            if (wl == 0) sWarpCount[warp] = cnt;
            __syncthreads();
            int base = 0, total = 0;
            for (int q = 0; q < kSelThreads / 32; q++) { const int c = sWarpCount[q]; if (q < warp) base += c; total += c; }
            if (total <= kSelListCap) {
                int pos = base;
                for (int i = w0 + wl; i - wl < w1; i += 32) {
                    bool in = false;
                    if (i < w1) { const float2 p = xy[i]; in = p.x >= x0 && p.x <= x1 && p.y >= y0 && p.y <= y1; }
                    const unsigned b = __ballot_sync(0xffffffffu, in);
                    if (in) sList[pos + __popc(b & ((1u << wl) - 1u))] = i;
                    pos += __popc(b);
                }
                nList = total;
                list = sList;
            }
            __syncthreads();
        }
        for (int j = 0; j < spanTiles; j++) {
            const long long t = first + (long long)j * groupsPerCta + grp;
            if (t >= nCells) break;
            const int cell = validIdx[t];
            const int row = cell / dem.nrCols, col = cell - row * dem.nrCols;
            // gis::getUtmXYFromRowCol (gis.cpp:806-810), then narrowed to float by the callee signatures (interpolation.h:77-95,162)
            const double x = dem.llx + dem.cellSize * (col + 0.5);
            const double y = dem.lly + dem.cellSize * (dem.nrRows - row - 0.5);
            const float xf = float(x), yf = float(y);
            unsigned char* rec = records + size_t(t) * recStride;
            WorkPtrs w;
            float localRadius = 0.f;
            // the record's work area plays the role of the fused kernel's shared-memory tile (capacity kRecCap); a selection that
            // does not fit it is left to the fused kernel
            const int n = localSelect<G>(m, st, xy, xf, yf, gw, kRecCap, rec + recWorkOff(), kRecCap, nullptr, w, localRadius, list,
                                         nList, kCandCap, listRadius);
            int nE = 0, flags = 0;
            PB_DEV_ASSERT(nList <= kSelListCap);
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
                else if (flags & CELL_FIT) atomicAdd(&fitHist[nE], 1u);      // the descents are handed out by fit size (order kernel)
            }
            grpSync<G>();
        }
    }
}

// the cells of a chunk that run descents, ordered by the number of points of their elevation fit (largest first): the lanes
// of a warp pull consecutive queue entries, so they work on fits of the same length and their data loops end together
constexpr int kHistBins = kFitCap + 1;
__global__ void local_order_scan_kernel(unsigned* __restrict__ hist /* in: counts, out: first position of each bin */,
                                        unsigned* __restrict__ total)
{
    if (threadIdx.x == 0) {
        unsigned run = 0;
        for (int b = kHistBins - 1; b >= 0; b--) { const unsigned c = hist[b]; hist[b] = run; run += c; }
        *total = run;
    }
}
__global__ void __launch_bounds__(256)
local_order_scatter_kernel(const unsigned char* __restrict__ records, size_t recStride, int nCells, unsigned* __restrict__ binPos,
                           int* __restrict__ order)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nCells) return;
    const CellHeader* h = reinterpret_cast<const CellHeader*>(records + size_t(t) * recStride);
    if (h->flags == CELL_FIT) order[atomicAdd(&binPos[h->nE], 1u)] = t;
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
local_descent_kernel(const LocalModel* __restrict__ mp, unsigned char* __restrict__ records, size_t recStride,
                     const int* __restrict__ order, const unsigned* __restrict__ nOrder, unsigned* __restrict__ cellCounter)
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
            if (lane == 0) {
                t = atomicAdd(cellCounter, 1u);
                t = t < *nOrder ? unsigned(order[t]) : 0xffffffffu;       // cells with a fit, largest fits first
            }
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t == 0xffffffffu) break;
            unsigned char* rec = records + size_t(t) * recStride;
            const CellHeader h = *reinterpret_cast<const CellHeader*>(rec);
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
            PB_DEV_ASSERT(nE <= kFitCap && slot < P::NSLOT);      // the selection kernel sent larger fits to the overflow list
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
                PB_DEV_ASSERT(slot < P::NSLOT && guess < nGuess);
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
constexpr int kFinThreads = 256;     // 4 CTAs per SM (64 registers): the kernel waits on the records' loads, occupancy beats registers
                                     // (2 CTAs at 128 registers: 169.8 ms on the 3.7M-cell probe, 3: 168.0, 4: 163.8, 6: 170.6)

template <int F>
__global__ void __launch_bounds__(kFinThreads, 4)
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


template <int F, int NCONS>
static cudaError_t launchDescentT(const LocalModel* dModel, unsigned char* records, size_t recStride, int nCells, const int* order,
                                  const unsigned* nOrder, unsigned* cellCounter, cudaStream_t s)
{
    constexpr int NPROD = 1;
    auto k = local_descent_kernel<F, NCONS, NPROD>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Pipe<F>::smemBytes);
    if (e != cudaSuccess) return e;
    const int grid = std::min(smCount(), std::max(1, nCells / 8));
    k<<<grid, (NCONS + NPROD) * 32, Pipe<F>::smemBytes, s>>>(dModel, records, recStride, order, nOrder, cellCounter);
    return cudaGetLastError();
}

template <int F>
static cudaError_t launchDescent(const LocalModel* dModel, unsigned char* records, size_t recStride, int nCells, const int* order,
                                 const unsigned* nOrder, unsigned* cellCounter, cudaStream_t s)
{
    // registers bound the CTA: each of the four SM sub-partitions owns 16 K registers and holds whole warps, so 12 warps (3 per
    // sub-partition) may use 168 registers each, 16 warps 128; the two-piece descent needs 146, the three-piece ones 182
    static const int wide = getenv("PB_LOCAL_WARPS16") ? atoi(getenv("PB_LOCAL_WARPS16")) : 0;
    if (F == PB_FIT_PIECEWISE_TWO && wide) return launchDescentT<F, 15>(dModel, records, recStride, nCells, order, nOrder, cellCounter, s);
    return launchDescentT<F, 11>(dModel, records, recStride, nCells, order, nOrder, cellCounter, s);
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

// one chunk of cells [validIdx, validIdx + nCells) through A, (order), B, C on stream s. records: nCells * recStride bytes;
// work: localPipelineWorkInts(nCells) ints = {cell counter, -, fit count, span counter, histogram[64], order[nCells]}, private
// to the chunk; overflowList / overflowCount: the cells left to the fused kernel, appended to by every chunk of a call (the
// caller zeroes the count once and runs the fused kernel over the list when all chunks are done).
size_t localPipelineWorkInts(int nCells) { return 4 + 64 + size_t(nCells); }

cudaError_t launchLocalPipelineChunk(const LocalModel* dModel, const LocalStations& st, const DevGrid& dem, const int* validIdx,
                                     int nCells, int elevFunction, float candRadius, unsigned char* records, size_t recStride,
                                     int* work, int* overflowList, int* overflowCount, float* out, unsigned* minmax, cudaStream_t s,
                                     int* launches)
{
    if (nCells <= 0) return cudaSuccess;
    static_assert(kHistBins <= 64, "histogram area too small");
    unsigned* counters = reinterpret_cast<unsigned*>(work);
    unsigned* hist = counters + 4;
    int* order = work + 4 + 64;
    cudaError_t e = cudaMemsetAsync(work, 0, (4 + 64) * sizeof(int), s);
    if (e != cudaSuccess) return e;
    cudaEvent_t evStart = nullptr;
    if (getenv("PB_LOCAL_TIMING")) { cudaEventCreate(&evStart); cudaEventRecord(evStart, s); }
    {
        // a span should not be much wider than the candidate radius, or its station list grows with it
        const int groupsPerCta = kSelThreads / kSelG;
        int spanTiles = candRadius > 0.f ? int(double(candRadius) / dem.cellSize / groupsPerCta) : 1;
        spanTiles = std::max(1, std::min(kSelSpanTilesMax, spanTiles));
        const int spanCells = groupsPerCta * spanTiles;
        const long long need = (nCells + spanCells - 1) / spanCells;
        const int grid = int(std::min<long long>(need, (long long)smCount() * 5));
        local_select_kernel<<<grid, kSelThreads, 0, s>>>(dModel, st, dem, validIdx, nCells, records, recStride, overflowList,
                                                         overflowCount, hist, spanTiles, counters + 3);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        local_order_scan_kernel<<<1, 32, 0, s>>>(hist, counters + 2);
        local_order_scatter_kernel<<<(nCells + 255) / 256, 256, 0, s>>>(records, recStride, nCells, hist, order);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    // PB_LOCAL_TIMING=1: per-phase device times of this chunk on stderr (a diagnostic: it synchronises the stream)
    static const bool timing = getenv("PB_LOCAL_TIMING") != nullptr;
    cudaEvent_t ev[4] = {};
    if (timing) {
        for (auto& x : ev) cudaEventCreate(&x);
        cudaEventDestroy(ev[0]);
        ev[0] = evStart;
        cudaEventRecord(ev[1], s);
    }
    switch (elevFunction) {
    case PB_FIT_PIECEWISE_THREE:
        e = launchDescent<PB_FIT_PIECEWISE_THREE>(dModel, records, recStride, nCells, order, counters + 2, counters, s);
        if (timing) cudaEventRecord(ev[2], s);
        if (e == cudaSuccess) e = launchFinish<PB_FIT_PIECEWISE_THREE>(dModel, st, dem, records, recStride, nCells, out, minmax, s);
        break;
    case PB_FIT_PIECEWISE_THREE_FREE:
        e = launchDescent<PB_FIT_PIECEWISE_THREE_FREE>(dModel, records, recStride, nCells, order, counters + 2, counters, s);
        if (timing) cudaEventRecord(ev[2], s);
        if (e == cudaSuccess) e = launchFinish<PB_FIT_PIECEWISE_THREE_FREE>(dModel, st, dem, records, recStride, nCells, out, minmax, s);
        break;
    default:
        e = launchDescent<PB_FIT_PIECEWISE_TWO>(dModel, records, recStride, nCells, order, counters + 2, counters, s);
        if (timing) cudaEventRecord(ev[2], s);
        if (e == cudaSuccess) e = launchFinish<PB_FIT_PIECEWISE_TWO>(dModel, st, dem, records, recStride, nCells, out, minmax, s);
        break;
    }
    if (timing) {
        cudaEventRecord(ev[3], s);
        cudaEventSynchronize(ev[3]);
        float a = 0, b = 0, c = 0;
        cudaEventElapsedTime(&a, ev[0], ev[1]);
        cudaEventElapsedTime(&b, ev[1], ev[2]);
        cudaEventElapsedTime(&c, ev[2], ev[3]);
        fprintf(stderr, "[pb local pipeline] %d cells: select+order %.3f ms, descents %.3f ms, finish %.3f ms\n", nCells, a, b, c);
        for (auto& x : ev) cudaEventDestroy(x);
    }
    if (launches) *launches += 5;
    return e;
}

}  // namespace pb
