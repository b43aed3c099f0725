// context.cuh — the engine context behind the opaque pb_context of include/praga_b200.h, shared by the .cu files that
// implement C-ABI entry points (engine.cu, kriging.cu).
#pragma once
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace pb {

struct RasterEntry {
    bool used = false;
    pb_raster_header h{};
    float* d = nullptr;
    size_t n = 0;
    // DEM role (built lazily): ascending list of valid cells + per-row offsets into it
    int* validIdx = nullptr;           // capacity n
    long long nValid = -1;
    std::vector<long long> rowStart;   // nrRows + 1 prefix counts (host), built on first row-band request
    bool borrowed = false;             // a lane context's view of its parent's raster: never released here
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinnedBuf {
    unsigned char* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost(&p, n);
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// ring of pinned staging slots: host gather/scatter of one slot overlaps the PCIe copy of the others
struct PinnedRing {
    static constexpr int kSlots = 8;
    unsigned char* p = nullptr;
    size_t slotBytes = 0;
    cudaEvent_t ev[kSlots]{};
    bool pending[kSlots]{};
    int next = 0;
    cudaError_t init(size_t bytesPerSlot)
    {
        if (p && slotBytes >= bytesPerSlot) return cudaSuccess;
        release();
        cudaError_t e = cudaMallocHost(&p, bytesPerSlot * kSlots);
        if (e != cudaSuccess) { p = nullptr; return e; }
        slotBytes = bytesPerSlot;
        for (auto& x : ev) cudaEventCreateWithFlags(&x, cudaEventDisableTiming);
        return cudaSuccess;
    }
    // next slot, waiting for the copy that last used it
    unsigned char* acquire(int* slot)
    {
        const int s = next;
        next = (next + 1) % kSlots;
        if (pending[s]) { cudaEventSynchronize(ev[s]); pending[s] = false; }
        *slot = s;
        return p + size_t(s) * slotBytes;
    }
    void submitted(int slot, cudaStream_t stream) { cudaEventRecord(ev[slot], stream); pending[slot] = true; }
    cudaError_t wait(int slot)
    {
        cudaError_t e = cudaSuccess;
        if (pending[slot]) { e = cudaEventSynchronize(ev[slot]); pending[slot] = false; }
        return e;
    }
    void release()
    {
        if (p) {
            for (int s = 0; s < kSlots; s++) if (pending[s]) cudaEventSynchronize(ev[s]);
            cudaFreeHost(p);
            for (auto& x : ev) if (x) cudaEventDestroy(x);
        }
        p = nullptr; slotBytes = 0; next = 0;
        for (auto& x : ev) x = nullptr;
        for (auto& x : pending) x = false;
    }
};

// device buffers recycled by size: the drop-in call uploads and frees rasters every time step, cudaMalloc/cudaFree would
// cost more than the kernel
struct DevicePool {
    struct Buf { void* p; size_t bytes; bool used; };
    std::vector<Buf> bufs;
    cudaError_t acquire(size_t bytes, void** out)
    {
        for (auto& b : bufs)
            if (!b.used && b.bytes >= bytes && b.bytes <= bytes + bytes / 4 + 4096) { b.used = true; *out = b.p; return cudaSuccess; }
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {     // give cached memory back and retry once
            trim();
            e = cudaMalloc(&p, bytes);
            if (e != cudaSuccess) return e;
        }
        bufs.push_back({p, bytes, true});
        *out = p;
        return cudaSuccess;
    }
    void release(void* p)
    {
        for (auto& b : bufs) if (b.p == p) { b.used = false; return; }
    }
    void trim()
    {
        for (size_t i = 0; i < bufs.size();)
            if (!bufs[i].used) { cudaFree(bufs[i].p); bufs.erase(bufs.begin() + i); }
            else i++;
    }
    void destroy() { for (auto& b : bufs) cudaFree(b.p); bufs.clear(); }
};

// aggregation geometry of a meteo grid (Crit3DMeteoPoint::aggregationPoints of every cell, meteo/meteoPoint.h:86-87),
// resident on the device: the batch drivers aggregate every gridded map onto the same cells (raster_ops.cu)
struct AggregationGeometry {
    bool used = false;
    double* x = nullptr;
    double* y = nullptr;
    long long* first = nullptr;        // nCells + 1
    int* maxNr = nullptr;              // aggregationPointsMaxNr
    float* scratch = nullptr;          // one float per point
    float* value = nullptr;            // nCells
    int nCells = 0;
    long long nPoints = 0;
};

// the (cell, weight) pairs of a set of macro areas, resident on the device (glocal.cu): they only change when the
// glocal weight maps do, every time step reuses them
struct ResidentMacroAreas {
    bool used = false;
    float* cells = nullptr;                // concatenated areaCellsDEM
    std::vector<long long> firstFloat;     // nAreas + 1 offsets (in floats) into cells
};

struct KrigingSystem;                       // kriging.cu
void krigingDestroyAll(std::vector<KrigingSystem*>& v);

}  // namespace pb

struct pb_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool ownStream = true;
    std::string err;
    std::vector<pb::RasterEntry> rasters;
    pb::DevBuf<unsigned char> dStations;
    pb::PinnedBuf hStations;
    pb::PinnedBuf hUpload;                  // page-locked mirror of one stage's small uploads (station_space.cu UploadBatch)
    pb::DevBuf<float> dOut;
    int outInitRaster = -1;             // dOut currently holds the flag pattern of this raster id
    pb::DevBuf<unsigned char> dScratch;     // point targets
    pb::DevBuf<unsigned char> dLocalStations, dLocalScratch, dLocalModel;   // local-detrending raster (K2)
    pb::DevBuf<unsigned char> dLocalRecords;                                // K2 pipeline: per-cell records of the two chunks in flight
    pb::DevBuf<int> dLocalWork, dLocalOverflow;                             // K2 pipeline: per-chunk counters / order; cells left to the fused kernel
    cudaStream_t sPipe[2] = {nullptr, nullptr};                             // K2 pipeline: chunks alternate between two streams
    cudaEvent_t evPipe[3] = {nullptr, nullptr, nullptr};
    pb::PinnedBuf hLocal;
    int* dLocalError = nullptr;
    int* hLocalError = nullptr;         // pinned
    pb::PinnedRing ring;                    // raster staging (H2D / D2H), 8 slots
    pb::PinnedRing ringDn;                  // band pipeline of the drop-in call: D2H staging (ring is then H2D only)
    cudaStream_t sUp = nullptr, sDn = nullptr;      // band pipeline: upload / download streams (kernels: stream)
    cudaEvent_t evBand[2 * 16 + 1]{};       // per band: rows uploaded, band gridded; + one fence
    unsigned char* dBand = nullptr;         // per band: valid-cell count (long long) and work-item counter (unsigned)
    unsigned char* hBand = nullptr;         // pinned copy of the counts
    pb::DevicePool pool;                    // raster-sized device buffers
    long long* dTotal = nullptr;            // valid-cell count of the last compaction
    long long* hTotal = nullptr;            // pinned
    unsigned* dMinMax = nullptr;
    unsigned* hMinMax = nullptr;        // pinned
    cudaEvent_t ev[4]{};
    pb_timing timing{};
    std::vector<pb::KrigingSystem*> kriging;    // pb_kriging_id -> system (kriging.cu)
    std::vector<pb::AggregationGeometry> aggregations;   // pb_aggregation_id -> geometry (raster_ops.cu)
    std::vector<pb::ResidentMacroAreas> macroAreas;      // pb_macro_areas_id -> resident cells (glocal.cu)
    std::vector<pb_context*> lanes;             // child contexts (own stream and buffers) of pb_interpolation_dem_batch
    // batch drivers: a lane's raster kernels leave `reserveSms` SMs to the station-space kernels of the other lanes and are
    // chained one after the other across lanes (the persistent K1 kernel fills every SM it gets: two of them in flight, or one
    // on all SMs, would make every small kernel of every other lane wait for a whole raster)
    pb_context* parent = nullptr;
    int reserveSms = 0;
    std::mutex rasterMutex;                     // parent: guards lastRaster
    cudaEvent_t lastRaster = nullptr;           // parent: completion of the raster kernel launched last by any lane
    cudaEvent_t evRaster = nullptr;             // lane: completion of this lane's last raster kernel
    // Crit3DMeteoPoint::topographicDistance: point index -> resident raster (pb_set_topographic_distance_maps), and the
    // descriptors of the stations staged by the current call
    std::unordered_map<int32_t, pb_raster_id> tdMaps;
    pb::DevBuf<unsigned char> dTdMaps;
    pb::DevBuf<float> dMapT;                    // a lane's air temperature map of the hour (pb_meteo_grid_step_batch)
    pb::DevBuf<float> dAgg;                     // a lane's aggregation scratch (one float per point) + cell values
};

namespace pb {

int failCtx(pb_context* ctx, int code, const std::string& msg);     // engine.cu
int ensureValidList(pb_context* ctx, RasterEntry& e);               // engine.cu
int ensureRowStart(pb_context* ctx, RasterEntry& e);                // engine.cu
// engine.cu: descriptors of the precomputed topographic-distance maps of n staged stations (pointIndex[k] = the station's
// Crit3DInterpolationDataPoint::index), uploaded on ctx->stream; *out = nullptr when none of them has a map
int stageTdMapsFor(pb_context* ctx, const int32_t* pointIndex, int n, const DevGrid** out);
// the same for the stations the IDW / Shepard staging keeps (excludeSupplemental filter, same order)
int stageTdMaps(pb_context* ctx, const pb_stations* st, const pb_settings* s, bool excludeSupplemental, const DevGrid** out);

}  // namespace pb

#define CUDA_TRY(ctx, call)                                                                              \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return pb::failCtx(ctx, PB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
    } while (0)
