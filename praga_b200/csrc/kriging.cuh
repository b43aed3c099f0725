// kriging.cuh — shared declarations of the ordinary-kriging path (kriging.cu: set-up, FP64 kernels, C ABI;
// kriging_tc.cu: tcgen05 variance contraction)
#pragma once
#include <cuda_fp16.h>
#include <vector>

#include "common.cuh"

namespace pb {

struct KrigingParams {
    int n, mode;                       // mode = PB_KRIGING_*
    double range, nugget, sillNugget, slope;
    double sill;                       // nugget + sillNugget
};

// targets: explicit points (xy != NULL: x0, y0, x1, y1, ...) or cells of a resident raster (validIdx + grid)
struct KrigingTargets {
    const double* xy;
    const int* validIdx;
    DevGrid grid;
};

// geometry of the tcgen05 variance kernel (kriging_tc.cu)
constexpr int kTcM = 128;              // MMA M = TMEM lanes
constexpr int kTcMTiles = 2;           // M tiles per CTA: 256 cells share every L^-1 tile that is loaded
constexpr int kTcCells = kTcM * kTcMTiles;
constexpr int kTcNT = 128;             // stations i per MMA N tile
constexpr int kTcBK = 32;              // stations j per pipeline stage
constexpr int kTcPass = 256;           // stations i per pass: 2 M tiles x 256 columns = the 512 TMEM columns
constexpr int kTcStages = 3;           // shared-memory pipeline depth
constexpr int kTcTileBytes = kTcNT * kTcBK * 2;     // one FP16 piece of one (i-tile, k-block) tile, canonical UMMA layout
constexpr int kTcChunk = 1 << 19;      // targets per evaluate() chunk on the tensor path

struct KrigingSystem {
    int n = 0;
    KrigingParams params{};
    bool tensor = false;
    double* dPos = nullptr;            // 2n
    double* dU = nullptr;              // n + 1: estimate = sum_j u_j D_j + u_n
    double* dA = nullptr;              // direct path: Vinv rows 0..n-1, n x (n+1)
    // whitened system (tensor path): M = L^-1 of C = sill - Gamma, FP16 hi/lo pieces pre-tiled for bulk copies
    unsigned char* dMhi = nullptr;     // tiles[(it * nKb + kb) * 2 + piece][kTcTileBytes]   (dMlo unused: pieces are interleaved)
    unsigned char* dMlo = nullptr;
    float* dZ = nullptr;               // z = M 1, padded to nPad
    unsigned long long* dMmaCount = nullptr;   // tcgen05 MMAs (256 x 256 x 16) issued by the pair kernel since the last read
    double* dKbBox = nullptr;          // per K stage (32 stations): xmin, xmax, ymin, ymax (+inf / -inf for padding stages)
    float2* dPosF = nullptr;           // not owned separately (inside dZ allocation)
    int nPad = 0, nIt = 0, nKb = 0;
    double zz = 0;                     // |z|^2
    double sumU = 0;                   // sum_j u_j, j < n (0 up to rounding)
    cudaStream_t sEst = nullptr;       // estimate stream + events of evaluate()'s estimate / variance overlap
    cudaEvent_t evEst[2] = {nullptr, nullptr}, evFin[2] = {nullptr, nullptr};
    cudaStream_t sDn = nullptr;        // download stream + events of pb_kriging_raster's chunk pipeline
    cudaEvent_t evDn[2] = {nullptr, nullptr};
    unsigned char* dWork = nullptr;
    size_t workBytes = 0;
};

void krigingFree(KrigingSystem* k);
int krigingUploadWhitened(KrigingSystem& k, const std::vector<double>& M, const std::vector<double>& z, double zz,
                          const double* pos);
cudaError_t launchKrigingVarianceTc(const KrigingSystem& k, const KrigingTargets& t, long long c0, int m, double* q,
                                    cudaStream_t s);
cudaError_t launchFill(float* out, long long n, float v, cudaStream_t s);     // idw_kernels.cu

}  // namespace pb
