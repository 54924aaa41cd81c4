// Fabric: how ranks find each other's control block and symmetric heaps.
//
//   inproc  ranks are objects/threads of one process (CPU tests; several logical ranks
//           on one GPU); heaps are plain allocations visible to everybody.
//   shm     one process per rank on one host (the torchrun model). The control block is
//           a POSIX shm segment. CPU heaps are shm segments mapped by every rank; CUDA
//           heaps are cudaMalloc'ed and exported with CUDA IPC so that peers read and
//           write them directly over NVLink/NVSwitch.
//
// This is the B200-native stand-in for ZMQVan + Van + Postoffice node management
// (include/zmq_van.h:30-250, src/van.cc:267-357): there are no sockets and no message
// (de)serialisation anywhere.
#pragma once
#include <atomic>
#include <memory>
#include <thread>
#include <string>
#include <vector>
#include "config.h"
#include "control.h"

namespace adapm {

class Fabric {
 public:
  static std::shared_ptr<Fabric> create(const Options& opt);
  virtual ~Fabric() {}

  ControlBlock* control() { return ctl_; }
  int rank() const { return rank_; }
  int world() const { return world_; }
  bool cuda() const { return cuda_; }
  int device() const { return device_; }

  // Collective: allocate `bytes` for this rank (zero-filled) and map all peers' heaps.
  virtual void allocate_heaps(uint64_t bytes) = 0;
  char* heap(int r) const { return heaps_[r]; }

  void node_barrier(const char* what) {
    ctl_->node_barrier.wait(world_, timeout_s_, what);
  }

  // Failure detector (the reference's heartbeat skeleton, van.cc:529-548, made effective): a thread that checks
  // every `period_ms` whether the processes of the other ranks are still alive; a dead peer breaks all barriers so
  // that every blocked call on this rank raises instead of waiting for the watchdog timeout. No-op for in-process
  // ranks. stop_failure_detector() is called before the last shutdown barrier (peers exit at different times).
  void start_failure_detector(int period_ms = 200);
  void stop_failure_detector();
  int dead_peer() const { return dead_peer_.load(); }   // -1: none detected

 protected:
  ControlBlock* ctl_ = nullptr;
  int rank_ = 0, world_ = 1, device_ = -1;
  bool cuda_ = false;
  double timeout_s_ = 300;
  std::vector<char*> heaps_;
  bool peers_are_processes_ = false;
  std::thread fd_thread_;
  std::atomic<bool> fd_stop_{false};
  std::atomic<int> dead_peer_{-1};
};

// device helpers implemented in cuda/device_mem.cu (only linked in the cuda build)
namespace cudamem {
bool available();
int device_count();
void set_device(int dev);
char* alloc_zeroed(uint64_t bytes);               // cudaMalloc + memset
void free_dev(char* p);
void export_handle(char* p, unsigned char* out128);   // cudaIpcGetMemHandle
char* import_handle(const unsigned char* in128);      // cudaIpcOpenMemHandle
void close_handle(char* p);
void enable_peer(int my_dev, int peer_dev);
}  // namespace cudamem

}  // namespace adapm
