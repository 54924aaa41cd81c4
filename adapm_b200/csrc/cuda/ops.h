// C++ entry points of the fused application kernels (implemented in ops_*.cu, bound to
// Python in ops_bind.cc).
#pragma once
#include <cstdint>
#include "cuda_backend.h"

namespace adapm {
namespace cudaops {

// word2vec SGNS: fused pull + score + AdaGrad + push   (ops_sgns.cu)
void sgns_step(CudaBackend& be, cudaStream_t stream, const Key* centers, const Key* contexts, const Key* negatives,
               int n_pairs, int neg, int d, float alpha, float* loss_out, unsigned long long* stats);

// key sampling (ops_sampler.cu). kind: 0 alias table, 1 uniform, 2 log-uniform
void sample_keys(CudaBackend& be, cudaStream_t stream, int kind, const float* prob, const int32_t* alias,
                 int64_t n_table, Key first, Key stride, Key* out, int64_t n, uint64_t seed, bool local_only,
                 int max_tries, unsigned long long* stats);

}  // namespace cudaops
}  // namespace adapm
