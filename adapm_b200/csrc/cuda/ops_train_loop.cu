// Native step driver for the word2vec SGNS training loop (the C++ counterpart of the loop in bench.py /
// apps/word2vec.py): per step Intent for a future batch, bounded run-ahead, (prefetched) host->device copy of the
// step's key batch, negative-sampling kernel, fused SGNS kernel, device->host copy of the loss, clock tick - all
// issued from C++ with the GIL released, so the step rate does not depend on Python launch jitter (8 ranks share one
// host). Reference equivalent: the per-sentence loop of apps/word2vec.cc:548-756, batched.
#include <cuda_runtime.h>

#include <vector>

#include "../adapm/node.h"
#include "ops.h"
#include "pm_kernels.cuh"

namespace adapm {
namespace cudaops {

SgnsLoop::SgnsLoop(Worker* worker, CudaBackend* be, const SgnsLoopConfig& cfg) : w_(worker), be_(be), c_(cfg) {
  ADAPM_CUDA_CHECK(cudaSetDevice(be_->device()));
  const size_t kb = (size_t)2 * c_.batch_pairs * sizeof(Key);
  depth_ = std::max(2, c_.max_inflight + 2);
  key_bufs_.resize(depth_);
  copied_.resize(depth_);
  done_.resize(depth_);
  done_valid_.assign(depth_, false);
  for (int i = 0; i < depth_; ++i) {
    ADAPM_CUDA_CHECK(cudaMalloc((void**)&key_bufs_[i], kb));
    ADAPM_CUDA_CHECK(cudaEventCreateWithFlags(&copied_[i], cudaEventDisableTiming));
    ADAPM_CUDA_CHECK(cudaEventCreateWithFlags(&done_[i], cudaEventDisableTiming));
  }
  ADAPM_CUDA_CHECK(cudaMalloc((void**)&neg_, (size_t)c_.batch_pairs * c_.negative * sizeof(Key)));
  ADAPM_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
}

SgnsLoop::~SgnsLoop() {
  cudaSetDevice(be_->device());
  cudaDeviceSynchronize();
  for (int i = 0; i < depth_; ++i) { cudaFree(key_bufs_[i]); cudaEventDestroy(copied_[i]); cudaEventDestroy(done_[i]); }
  cudaFree(neg_);
  cudaStreamDestroy(copy_stream_);
}

// batches[s]: pinned host [2, B] keys (resident = false) or device [2, B] keys (resident = true) of step s;
// intent_keys[s] / intent_counts[s]: the distinct keys of batch s (host memory) for Intent.
void SgnsLoop::run(cudaStream_t stream, int64_t first, int64_t n, bool resident, const std::vector<uintptr_t>& batches,
                   const std::vector<uintptr_t>& intent_keys, const std::vector<int64_t>& intent_counts, float* loss_dev,
                   float* loss_host, unsigned long long* stats, int64_t step_no0, float alpha) {
  ADAPM_CUDA_CHECK(cudaSetDevice(be_->device()));
  be_->track_stream(stream);
  const int B = c_.batch_pairs;
  const size_t kb = (size_t)2 * B * sizeof(Key);
  const int64_t total = (int64_t)batches.size();
  auto prefetch = [&](int64_t s) {   // host -> device copy of step s's keys on the copy stream
    if (resident || s >= total || s >= first + n) return;
    const int slot = (int)(s % depth_);
    if (done_valid_[slot]) ADAPM_CUDA_CHECK(cudaStreamWaitEvent(copy_stream_, done_[slot], 0));   // the buffer's previous user
    ADAPM_CUDA_CHECK(cudaMemcpyAsync(key_bufs_[slot], (const void*)batches[s], kb, cudaMemcpyHostToDevice, copy_stream_));
    ADAPM_CUDA_CHECK(cudaEventRecord(copied_[slot], copy_stream_));
  };
  prefetch(first);
  for (int64_t s = first; s < first + n; ++s) {
    if (c_.signal_intent && s + c_.read_ahead < (int64_t)intent_keys.size()) {
      const int64_t f = s + c_.read_ahead;
      const Clock at = w_->currentClock() + c_.read_ahead;
      w_->Intent((const Key*)intent_keys[f], (size_t)intent_counts[f], at, at + 1);
    }
    const int slot = (int)(s % depth_);
    // bounded run-ahead: at most max_inflight steps queued behind the one that is running
    const int64_t back = s - (c_.max_inflight + 1);
    if (back >= first) ADAPM_CUDA_CHECK(cudaEventSynchronize(done_[(int)(back % depth_)]));
    const Key* keys;
    if (resident) {
      keys = (const Key*)batches[s];
    } else {
      ADAPM_CUDA_CHECK(cudaStreamWaitEvent(stream, copied_[slot], 0));
      keys = key_bufs_[slot];
    }
    ADAPM_CUDA_CHECK(cudaMemsetAsync(loss_dev, 0, sizeof(float), stream));
    const uint64_t seed = ((uint64_t)c_.model_seed * 1000003ull + (uint64_t)c_.rank * 7919ull + (uint64_t)(step_no0 + (s - first))) & 0xFFFFFFFFFFFFull;
    sample_keys(*be_, stream, c_.sampler_kind, c_.prob, c_.alias, c_.n_table, c_.first_key, c_.key_stride, neg_,
                (int64_t)B * c_.negative, seed, c_.local_only, 64, c_.sampler_stats);
    sgns_step(*be_, stream, keys, keys + B, neg_, B, c_.negative, c_.embed_dim, alpha, loss_dev, stats, 0);
    if (!resident && loss_host) ADAPM_CUDA_CHECK(cudaMemcpyAsync(loss_host + s, loss_dev, sizeof(float), cudaMemcpyDeviceToHost, stream));
    ADAPM_CUDA_CHECK(cudaEventRecord(done_[slot], stream));
    done_valid_[slot] = true;
    prefetch(s + 1);
    w_->advanceClock();
  }
}

}  // namespace cudaops
}  // namespace adapm
