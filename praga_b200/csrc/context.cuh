// context.cuh — the engine context behind the opaque pb_context of include/praga_b200.h, shared by the .cu files that
// implement C-ABI entry points (engine.cu, kriging.cu).
#pragma once
#include <string>
#include <vector>

#include "common.cuh"

namespace pb {

struct RasterEntry {
    bool used = false;
    pb_raster_header h{};
    float* d = nullptr;
    size_t n = 0;
    // DEM role (built lazily): ascending list of valid cells + per-row offsets into it
    int* validIdx = nullptr;
    long long nValid = -1;
    std::vector<long long> rowStart;   // nrRows + 1 prefix counts (host)
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
    pb::DevBuf<float> dOut;
    int outInitRaster = -1;             // dOut currently holds the flag pattern of this raster id
    pb::DevBuf<unsigned char> dScratch;     // point targets
    pb::DevBuf<unsigned char> dLocalStations, dLocalScratch, dLocalModel;   // local-detrending raster (K2)
    pb::PinnedBuf hLocal;
    int* dLocalError = nullptr;
    int* hLocalError = nullptr;         // pinned
    pb::PinnedBuf hStage;                   // raster staging (H2D / D2H)
    unsigned* dMinMax = nullptr;
    unsigned* hMinMax = nullptr;        // pinned
    cudaEvent_t ev[4]{};
    pb_timing timing{};
    std::vector<pb::KrigingSystem*> kriging;    // pb_kriging_id -> system (kriging.cu)
};

namespace pb {

int failCtx(pb_context* ctx, int code, const std::string& msg);     // engine.cu
int ensureValidList(pb_context* ctx, RasterEntry& e);               // engine.cu

}  // namespace pb

#define CUDA_TRY(ctx, call)                                                                              \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return pb::failCtx(ctx, PB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
    } while (0)
