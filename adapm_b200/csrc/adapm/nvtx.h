// NVTX ranges around the worker API calls and the sync round ("NVTX ranges per API call + sync phase").
// Header-only NVTX 3: without an attached tool (Nsight Systems / Compute) the first call resolves to no-op stubs, so the
// ranges cost a few nanoseconds; builds without the CUDA toolkit headers (sanitizer build of the host core) compile them
// away.
#pragma once
#if defined(__has_include)
#if __has_include(<nvtx3/nvToolsExt.h>) && !defined(ADAPM_NO_NVTX)
#include <nvtx3/nvToolsExt.h>
#define ADAPM_HAVE_NVTX 1
#endif
#endif

namespace adapm {
#if defined(ADAPM_HAVE_NVTX)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
inline bool nvtx_compiled_in() { return true; }
#else
struct NvtxRange {
  explicit NvtxRange(const char*) {}
};
inline bool nvtx_compiled_in() { return false; }
#endif
}  // namespace adapm

#define ADAPM_NVTX_CAT2(a, b) a##b
#define ADAPM_NVTX_CAT(a, b) ADAPM_NVTX_CAT2(a, b)
#define ADAPM_NVTX(name) ::adapm::NvtxRange ADAPM_NVTX_CAT(adapm_nvtx_range_, __LINE__)(name)
