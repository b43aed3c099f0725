// raster_io.cu — the on-disk raster format at either end of the path: ESRI float grids (.hdr + .flt)
// streamed between the file and HBM through the pinned staging ring, without a pageable full-size copy in between.
//
//   gis::readEsriGridHeader   agrolib/gis/gisIO.cpp:118-193  (keys case-insensitive; NCOLS NROWS CELLSIZE XLLCORNER YLLCORNER
//                                                            NODATA_value / NODATA; >= 6 of them; DATATYPE = bytes per value)
//   gis::readRasterFloatData  :347-422  (row-wise fread; 4-byte float, 2-byte int16 or 1-byte values)
//   gis::readEsriGrid         :516-557  (then updateMinMaxRasterGrid, gis.cpp:569-614)
//   gis::writeEsriGridHeader  :432-467, gis::writeEsriGridFlt :477-497, gis::writeEsriGrid :504-513
#include <errno.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "context.cuh"
#include "nvtx.cuh"

namespace pb {

__device__ __forceinline__ unsigned orderedKeyIo(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// gis::updateMinMaxRasterGrid (gis.cpp:569-614): min / max over the cells that are neither the flag nor NODATA
__global__ void __launch_bounds__(256) raster_minmax_kernel(const float* __restrict__ v, long long n, float flag, unsigned* __restrict__ minmax)
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
        if (vmin != INFINITY) atomicMin(&minmax[0], orderedKeyIo(vmin));
        if (vmax != -INFINITY) atomicMax(&minmax[1], orderedKeyIo(vmax));
    }
}

}  // namespace pb

using namespace pb;

namespace {

int fail(pb_context* ctx, int code, const std::string& msg) { return pb::failCtx(ctx, code, msg); }

constexpr size_t kIoSlotBytes = size_t(8) << 20;

std::string upperCase(std::string s)
{
    std::transform(s.begin(), s.end(), s.begin(), ::toupper);
    return s;
}

// base name without a trailing ".flt" (readEsriGridHeader :123-128 removes the LAST occurrence of ".flt")
std::string stripFlt(const std::string& fileName)
{
    std::string fn = fileName;
    const std::size_t found = fn.rfind(".flt");
    if (found != std::string::npos) fn.replace(found, 4, "");
    return fn;
}

// readEsriGridHeader; nrBytes: bytes per value (DATATYPE key, default 4)
int readHeader(pb_context* ctx, const std::string& fileName, pb_raster_header& h, int& nrBytes)
{
    const std::string fn = stripFlt(fileName) + ".hdr";
    std::ifstream f(fn.c_str());
    if (f.fail()) return fail(ctx, PB_ERR_INVALID, "Missing file: " + fn);
    memset(&h, 0, sizeof(h));
    h.flag = kNoData;                       // Crit3DRasterHeader() default (gis.cpp:89-98)
    nrBytes = 4;
    int nrKeys = 0;
    std::string line;
    try {
        while (f.good()) {
            std::getline(f, line);
            std::istringstream ss(line);
            std::string key, value;
            ss >> key;
            ss >> value;
            if (key.empty() || value.empty()) continue;
            const std::string up = upperCase(key);
            if (up == "NCOLS" || up == "NROWS" || up == "CELLSIZE" || up == "XLLCORNER" || up == "YLLCORNER" || up == "NODATA_VALUE" ||
                up == "NODATA")
                nrKeys++;
            if (up == "NCOLS") h.nrCols = std::stoi(value);
            else if (up == "NROWS") h.nrRows = std::stoi(value);
            else if (up == "DATATYPE") {
                nrBytes = std::stoi(value);
                if (nrBytes > 4) return fail(ctx, PB_ERR_INVALID, "Wrong data type:" + value + " The maximum allowed is 4 (float).");
            } else if (up == "XLLCORNER") h.llx = std::stod(value);
            else if (up == "YLLCORNER") h.lly = std::stod(value);
            else if (up == "CELLSIZE") {
                h.cellSize = std::stod(value);
                h.invCellSize = 1.0 / h.cellSize;
            } else if (up == "NODATA_VALUE" || up == "NODATA") h.flag = std::stof(value);
        }
    } catch (const std::exception&) {
        return fail(ctx, PB_ERR_INVALID, "Malformed value in header file: " + fn);
    }
    if (nrKeys < 6) return fail(ctx, PB_ERR_INVALID, "Missing keys in header file.");
    if (h.nrRows <= 0 || h.nrCols <= 0) return fail(ctx, PB_ERR_INVALID, "Header without rows / columns.");
    return PB_OK;
}

// file extension rule of readEsriGrid (:525-533): a 4-character suffix starting with '.'
std::string extensionOf(const std::string& fileName)
{
    if (fileName.size() > 4) {
        std::string suffix = fileName.substr(fileName.size() - 4);
        if (suffix[0] == '.') {
            std::transform(suffix.begin(), suffix.end(), suffix.begin(), ::tolower);
            return suffix;
        }
    }
    return "";
}

}  // namespace

extern "C" {

int pb_raster_read_esri_header(pb_context* ctx, const char* fileName, pb_raster_header* header)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (!fileName || !header) return fail(ctx, PB_ERR_INVALID, "null argument");
    int nrBytes;
    return readHeader(ctx, fileName, *header, nrBytes);
}

int pb_raster_read_esri(pb_context* ctx, const char* fileName, pb_raster_id* id, pb_raster_header* header, float* minimum,
                        float* maximum)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (!fileName || !id) return fail(ctx, PB_ERR_INVALID, "null argument");
    cudaSetDevice(ctx->device);
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    const std::string name(fileName), ext = extensionOf(name);
    if (!(ext.empty() || ext == ".flt")) return fail(ctx, PB_ERR_UNSUPPORTED, "only ESRI float grids (.hdr + .flt) are read on this path");
    pb_raster_header h;
    int nrBytes;
    int rc = readHeader(ctx, name, h, nrBytes);
    if (rc) return rc;
    if (nrBytes != 4 && nrBytes != 2 && nrBytes != 1) return fail(ctx, PB_ERR_INVALID, "Unsupported raster format.");
    const std::string flt = ext.empty() ? name + ".flt" : name;
    FILE* fp = fopen(flt.c_str(), "rb");
    if (!fp) return fail(ctx, PB_ERR_INVALID, "Error in opening raster file.");
    const int rows = h.nrRows, cols = h.nrCols;
    const size_t n = size_t(rows) * cols;
    if (n > size_t(0x7fffffff) || size_t(cols) * sizeof(float) > kIoSlotBytes) { fclose(fp); return fail(ctx, PB_ERR_UNSUPPORTED, "raster too large"); }
    // a new resident raster
    int slot = -1;
    for (size_t i = 0; i < ctx->rasters.size(); i++) if (!ctx->rasters[i].used) { slot = (int)i; break; }
    if (slot < 0) { ctx->rasters.emplace_back(); slot = (int)ctx->rasters.size() - 1; }
    void* d = nullptr;
    if (cudaError_t e = ctx->pool.acquire(n * sizeof(float), &d); e != cudaSuccess) { fclose(fp); return fail(ctx, PB_ERR_CUDA, cudaGetErrorString(e)); }
    RasterEntry& entry = ctx->rasters[slot];
    entry = RasterEntry();
    entry.d = static_cast<float*>(d);
    entry.h = h;
    entry.n = n;
    entry.used = true;
    if (ctx->outInitRaster == slot) ctx->outInitRaster = -1;
    if (cudaError_t e = ctx->ring.init(kIoSlotBytes); e != cudaSuccess) { fclose(fp); return fail(ctx, PB_ERR_CUDA, cudaGetErrorString(e)); }
    const int chunkRows = std::max(1, (int)(kIoSlotBytes / (sizeof(float) * cols)));
    std::vector<unsigned char> narrow;
    if (nrBytes != 4) narrow.resize(size_t(cols) * nrBytes);
    bool ok = true;
    for (int r0 = 0; r0 < rows && ok; r0 += chunkRows) {
        const int r1 = std::min(rows, r0 + chunkRows);
        int s;
        float* stage = reinterpret_cast<float*>(ctx->ring.acquire(&s));
        if (nrBytes == 4) {
            const size_t want = size_t(r1 - r0) * cols;
            ok = fread(stage, sizeof(float), want, fp) == want;
        } else {
            for (int row = r0; row < r1 && ok; row++) {
                ok = fread(narrow.data(), nrBytes, cols, fp) == size_t(cols);
                float* dst = stage + size_t(row - r0) * cols;
                if (nrBytes == 2) for (int c = 0; c < cols; c++) dst[c] = static_cast<float>(reinterpret_cast<const int16_t*>(narrow.data())[c]);
                else for (int c = 0; c < cols; c++) dst[c] = static_cast<float>(narrow[c]);
            }
        }
        if (!ok) break;
        if (cudaMemcpyAsync(entry.d + size_t(r0) * cols, stage, sizeof(float) * size_t(r1 - r0) * cols, cudaMemcpyHostToDevice,
                            ctx->stream) != cudaSuccess) { ok = false; break; }
        ctx->ring.submitted(s, ctx->stream);
    }
    fclose(fp);
    if (!ok) {
        cudaStreamSynchronize(ctx->stream);
        ctx->pool.release(entry.d);
        entry = RasterEntry();
        return fail(ctx, PB_ERR_INVALID, "Error reading raster data.");
    }
    ctx->timing.h2d_bytes = (int64_t)(n * sizeof(float));
    // updateMinMaxRasterGrid
    ctx->hMinMax[0] = 0xffffffffu;
    ctx->hMinMax[1] = 0u;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dMinMax, ctx->hMinMax, 2 * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
    raster_minmax_kernel<<<1184, 256, 0, ctx->stream>>>(entry.d, (long long)n, h.flag, ctx->dMinMax);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.launches++;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hMinMax, ctx->dMinMax, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    auto keyToFloat = [](unsigned k) { unsigned u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k; float f; memcpy(&f, &u, 4); return f; };
    const bool any = ctx->hMinMax[0] != 0xffffffffu;
    if (minimum) *minimum = any ? keyToFloat(ctx->hMinMax[0]) : kNoData;
    if (maximum) *maximum = any ? keyToFloat(ctx->hMinMax[1]) : kNoData;
    if (header) *header = h;
    *id = slot;
    return PB_OK;
}

int pb_raster_write_esri(pb_context* ctx, pb_raster_id id, const char* fileName)
{
    PB_NVTX_RANGE();
    if (!ctx) return PB_ERR_INVALID;
    if (!fileName) return fail(ctx, PB_ERR_INVALID, "null argument");
    if (id < 0 || id >= (int)ctx->rasters.size() || !ctx->rasters[id].used) return fail(ctx, PB_ERR_INVALID, "bad raster id");
    cudaSetDevice(ctx->device);
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    const RasterEntry& e = ctx->rasters[id];
    const std::string base(fileName);
    {   // writeEsriGridHeader: same text, byte for byte
        const std::string hdr = base + ".hdr";
        std::ofstream f(hdr.c_str());
        if (f.fail()) return fail(ctx, PB_ERR_INVALID, "Error writing file: " + hdr + '\n' + strerror(errno));
        char xll[64], yll[64];
        snprintf(xll, sizeof(xll), "%.03f", e.h.llx);
        snprintf(yll, sizeof(yll), "%.03f", e.h.lly);
        f << "ncols         " << e.h.nrCols << "\n";
        f << "nrows         " << e.h.nrRows << "\n";
        f << "xllcorner     " << xll << "\n";
        f << "yllcorner     " << yll << "\n";
        f << "cellsize      " << e.h.cellSize << "\n";
        f << "NODATA_value  " << e.h.flag << "\n";
        f << "NODATA        " << e.h.flag << "\n";
        f << "byteorder     LSBFIRST" << "\n";
    }
    const std::string flt = base + ".flt";
    FILE* fp = fopen(flt.c_str(), "wb");
    if (!fp) return fail(ctx, PB_ERR_INVALID, "Error writing file: " + flt + '\n' + strerror(errno));
    const int rows = e.h.nrRows, cols = e.h.nrCols;
    if (size_t(cols) * sizeof(float) > kIoSlotBytes) { fclose(fp); return fail(ctx, PB_ERR_UNSUPPORTED, "raster row larger than a staging slot"); }
    if (cudaError_t ce = ctx->ring.init(kIoSlotBytes); ce != cudaSuccess) { fclose(fp); return fail(ctx, PB_ERR_CUDA, cudaGetErrorString(ce)); }
    const int chunkRows = std::max(1, (int)(kIoSlotBytes / (sizeof(float) * cols)));
    struct Chunk { int r0, r1, slot; float* stage; };
    std::vector<Chunk> inflight;
    size_t head = 0;
    bool ok = true;
    auto flush = [&](const Chunk& c) {
        if (ctx->ring.wait(c.slot) != cudaSuccess) { ok = false; return; }
        const size_t want = size_t(c.r1 - c.r0) * cols;
        if (fwrite(c.stage, sizeof(float), want, fp) != want) ok = false;
    };
    for (int r0 = 0; r0 < rows && ok; r0 += chunkRows) {
        if (inflight.size() - head >= size_t(PinnedRing::kSlots - 1)) flush(inflight[head++]);
        Chunk c;
        c.r0 = r0;
        c.r1 = std::min(rows, r0 + chunkRows);
        c.stage = reinterpret_cast<float*>(ctx->ring.acquire(&c.slot));
        if (cudaMemcpyAsync(c.stage, e.d + size_t(r0) * cols, sizeof(float) * size_t(c.r1 - r0) * cols, cudaMemcpyDeviceToHost,
                            ctx->stream) != cudaSuccess) { ok = false; break; }
        ctx->ring.submitted(c.slot, ctx->stream);
        inflight.push_back(c);
    }
    for (; head < inflight.size() && ok; head++) flush(inflight[head]);
    cudaStreamSynchronize(ctx->stream);
    fclose(fp);
    ctx->timing.d2h_bytes = (int64_t)(size_t(rows) * cols * sizeof(float));
    if (!ok) return fail(ctx, PB_ERR_CUDA, "Error writing raster data.");
    return PB_OK;
}

}  // extern "C"
