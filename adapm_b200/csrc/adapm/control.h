// Host-side control plane shared by all ranks of a job.
//
// Replaces the reference's scheduler process, ADD_NODE rendezvous, BARRIER messages and
// heartbeats (src/van.cc:40-210, src/postoffice.cc:149-200): all ranks of one box map one
// small control block (heap memory for in-process ranks, a POSIX shm segment for
// one-process-per-GPU) and coordinate through atomics in it. Nothing here is on the
// data path; the data path is peer memory (layout.h).
#pragma once
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>
#include "base.h"
#include "log.h"

namespace adapm {

// Sense-reversing barrier usable across processes (lives in shared memory).
struct ShmBarrier {
  std::atomic<int32_t> count;
  std::atomic<int32_t> generation;
  std::atomic<int32_t> broken;

  void init() { count.store(0); generation.store(0); broken.store(0); }

  // Returns the generation that was completed. Throws after timeout_s (watchdog).
  void wait(int participants, double timeout_s, const char* what) {
    if (participants <= 1) return;
    int gen = generation.load(std::memory_order_acquire);
    int arrived = count.fetch_add(1, std::memory_order_acq_rel) + 1;
    if (arrived == participants) {
      count.store(0, std::memory_order_relaxed);
      generation.store(gen + 1, std::memory_order_release);
      return;
    }
    auto t0 = std::chrono::steady_clock::now();
    int spins = 0;
    while (generation.load(std::memory_order_acquire) == gen) {
      if (broken.load(std::memory_order_relaxed)) throw Error(std::string("barrier broken by a failed peer: ") + what);
      if (++spins < 200) { std::this_thread::yield(); continue; }
      std::this_thread::sleep_for(std::chrono::microseconds(spins < 2000 ? 20 : 200));
      if ((spins & 1023) == 0) {
        double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (el > timeout_s) {
          broken.store(1);
          throw Error(std::string("watchdog: barrier '") + what + "' timed out (a peer died or hangs)");
        }
      }
    }
  }
};

struct alignas(64) RankControl {
  std::atomic<int64_t> worker_clock[MAX_LOCAL_WORKERS];
  std::atomic<int32_t> worker_state[MAX_LOCAL_WORKERS];  // 0 = not registered, 1 = running, 2 = finalized
  std::atomic<uint64_t> rounds_done;     // completed sync rounds of this rank
  std::atomic<int32_t> stop_requested;   // server wants to shut the sync loop down
  std::atomic<int32_t> sweep_requested;  // a WaitSync caller wants guaranteed propagation
  std::atomic<int32_t> snap_stop;        // snapshots published before the round-start barrier
  std::atomic<int32_t> snap_sweep;
  std::atomic<int32_t> attached;
  std::atomic<int32_t> alive;            // heartbeat counter (failure detection)
  unsigned char ipc_handle[128];         // cudaIpcMemHandle_t of the heap (cuda backend)
  uint64_t heap_bytes;
  int32_t device;
  int32_t pid;
  int32_t vmm_ok;                        // this rank can allocate its heap with the VMM API and pass it as a POSIX fd
  int32_t mc_ok;                         // ... and its device supports NVLS multicast objects
};

// Host RPC mailboxes (rpc.h): one multi-producer / single-consumer ring of fragments per rank.
constexpr int MAIL_SLOTS = 64;
constexpr int MAIL_BODY = 976;

struct MailSlot {
  std::atomic<uint32_t> state;   // 0 = free, 1 = published
  int32_t sender, app_id, customer_id, head, timestamp;
  int32_t flags;                 // bit 0: request (else response)
  uint32_t msg_id, frag_off, total_len, frag_len;
  char body[MAIL_BODY];
};
static_assert(sizeof(MailSlot) == 1024 - 4, "MailSlot layout");

struct Mailbox {
  std::atomic<uint64_t> tail;    // next ticket (producers)
  std::atomic<uint64_t> head;    // next ticket to consume (owner)
  MailSlot slots[MAIL_SLOTS];
};

struct ControlBlock {
  uint32_t magic;
  int32_t world;
  int32_t workers;
  std::atomic<int32_t> initialized;
  ShmBarrier sync_barrier;    // one participant per rank: the sync threads
  ShmBarrier node_barrier;    // one participant per rank: Server::barrier / setup
  ShmBarrier worker_barrier;  // world * workers participants: Worker::Barrier
  std::atomic<int32_t> round_sweep;  // rank 0's decision for the current round
  std::atomic<int32_t> round_stop;
  std::atomic<int64_t> allreduce_buf[64];
  RankControl ranks[MAX_RANKS];
  Mailbox mail[MAX_RANKS];
};

constexpr uint32_t kControlMagic = 0xADA9B200u;

}  // namespace adapm
