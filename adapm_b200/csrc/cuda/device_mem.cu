// Device memory helpers behind fabric.h's cudamem namespace: heap allocation, CUDA IPC
// export/import (the NVLink/NVSwitch peer mapping between one-process-per-GPU ranks) and
// peer access between devices of one process.
#include <cuda_runtime.h>
#include <cstring>
#include <sstream>
#include "../adapm/fabric.h"
#include "group.cuh"

namespace adapm {
namespace cudamem {

bool available() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return n > 0;
}
int device_count() {
  int n = 0;
  ADAPM_CUDA_CHECK(cudaGetDeviceCount(&n));
  return n;
}
void set_device(int dev) { ADAPM_CUDA_CHECK(cudaSetDevice(dev)); }
char* alloc_zeroed(uint64_t bytes) {
  void* p = nullptr;
  ADAPM_CUDA_CHECK(cudaMalloc(&p, bytes));
  ADAPM_CUDA_CHECK(cudaMemset(p, 0, bytes));
  ADAPM_CUDA_CHECK(cudaDeviceSynchronize());
  return (char*)p;
}
void free_dev(char* p) { cudaFree(p); }
void export_handle(char* p, unsigned char* out128) {
  static_assert(sizeof(cudaIpcMemHandle_t) <= 128, "ipc handle size");
  cudaIpcMemHandle_t h;
  ADAPM_CUDA_CHECK(cudaIpcGetMemHandle(&h, p));
  memset(out128, 0, 128);
  memcpy(out128, &h, sizeof(h));
}
char* import_handle(const unsigned char* in128) {
  cudaIpcMemHandle_t h;
  memcpy(&h, in128, sizeof(h));
  void* p = nullptr;
  ADAPM_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  return (char*)p;
}
void close_handle(char* p) { cudaIpcCloseMemHandle(p); }
void enable_peer(int my_dev, int peer_dev) {
  if (my_dev == peer_dev) return;
  int can = 0;
  ADAPM_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, my_dev, peer_dev));
  if (!can) throw Error("device " + std::to_string(my_dev) + " cannot access peer " + std::to_string(peer_dev) + " (no NVLink/P2P path)");
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_dev, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return; }
  ADAPM_CUDA_CHECK(e);
}

}  // namespace cudamem
}  // namespace adapm
