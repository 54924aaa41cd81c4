// C++ entry points of the fused application kernels (implemented in ops_*.cu, bound to
// Python in ops_bind.cc).
#pragma once
#include <cstdint>
#include "cuda_backend.h"

namespace adapm {
namespace cudaops {
}  // namespace cudaops
}  // namespace adapm
