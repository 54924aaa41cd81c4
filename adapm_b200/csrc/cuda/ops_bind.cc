// Python bindings of the fused application kernels.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "ops.h"

namespace py = pybind11;

namespace adapm {
namespace cudaops {

void bind(py::module_& m) {
  (void)m;
}

}  // namespace cudaops
}  // namespace adapm
