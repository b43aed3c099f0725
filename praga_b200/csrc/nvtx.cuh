// nvtx.cuh — one NVTX range per C-ABI entry point (tracing): nsys / ncu --nvtx show the drop-in's calls by the
// names of include/praga_b200.h. nvtx3 is header-only and costs nothing unless a tool is attached.
#pragma once
#include <nvtx3/nvToolsExt.h>

namespace pb {
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace pb
#define PB_NVTX_RANGE() pb::NvtxRange pbNvtxRange_(__func__)
