// Device-side pre-pass for Intent (experimental, opt-in from the models: ADAPM_INTENT_PREPASS=1).
//
// Most keys of a training batch are already usable from local memory when their intent is signalled (owned, replica,
// placeholder, relocation in flight). For those, signalling intent only means "extend the end clock of my intent":
// an atomicMax on intent_end[slot][worker], which is safe at any time - the sync round reads the end clocks, it never
// resets them while the slot lives. Only the keys WITHOUT a usable local slot need the sync thread (allocate a
// placeholder, request the key from its owner). This kernel does the first part on the device and compacts the second
// part into a short list that the caller hands to Worker::Intent - taking ~98 % of the per-key host work
// (sync_engine.cc: collect_intents) out of the sync thread in steady state.
//
// Best effort like every intent: an extension that races with the expiry of the previous intent of the same slot
// (phase A already decided to drop the replica / phase B to relocate the key away) is lost; the key is then simply
// not local when it is used and the next batch that contains it goes through the host path.
#include <cuda_runtime.h>

#include "ops.h"
#include "pm_kernels.cuh"

namespace adapm {
namespace cudaops {

namespace {

__global__ void __launch_bounds__(256)
intent_prepass_kernel(const __grid_constant__ Ctx c, const Key* __restrict__ keys, int64_t n, Clock end, int worker,
                      Key* __restrict__ out_keys, unsigned int* __restrict__ out_count) {
  const int lane = threadIdx.x & 31;
  dev::cta_enter(c);
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < n; i0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i0 + threadIdx.x;
    bool need = false;
    Key k = -1;
    if (i < n) {
      k = keys[i];
      if (k >= 0 && k < c.L.num_keys) need = !extend_intent_if_local(c, k, worker, end);   // protocol.h, shared with the host
      else need = true;   // out of range: hand it to the host path, whose Intent() raises for it (like Worker::Intent)
    }
    // warp-aggregated append of the keys that still need the sync thread
    const unsigned mask = __ballot_sync(0xffffffffu, need);
    if (mask) {
      unsigned base = 0;
      if (lane == (__ffs(mask) - 1)) base = atomicAdd(out_count, (unsigned)__popc(mask));
      base = __shfl_sync(0xffffffffu, base, __ffs(mask) - 1);
      if (need) out_keys[base + __popc(mask & ((1u << lane) - 1u))] = k;
    }
  }
  dev::cta_exit(c);
}

}  // namespace

void intent_prepass(CudaBackend& be, cudaStream_t stream, const Key* keys, int64_t n, Clock end, int worker, Key* out_keys,
                    unsigned int* out_count) {
  if (n <= 0) return;
  ADAPM_CHECK(worker >= 0 && worker < (int)be.ctx().L.workers, "intent_prepass: bad worker id");
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)be.num_sms() * 4);
  intent_prepass_kernel<<<blocks, 256, 0, stream>>>(be.ctx(), keys, n, end, worker, out_keys, out_count);
  ADAPM_COUNT_LAUNCH();
  ADAPM_CUDA_CHECK(cudaGetLastError());
}

}  // namespace cudaops
}  // namespace adapm
