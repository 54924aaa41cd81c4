// Store: layout construction + the backend interface (CPU executor here, CUDA in cuda/).
#pragma once
#include <algorithm>
#include <atomic>
#include <map>
#include <memory>
#include <vector>
#include "config.h"
#include "fabric.h"
#include "layout.h"
#include "protocol.h"

namespace adapm {

// Describes value lengths: uniform, or one length per key (reference: ColoKVServer(len) vs
// ColoKVServer(value_lengths), coloc_kv_server_handle.h:996-999, bindings.cc:88-94).
struct ValueSpec {
  int64_t num_keys = 0;
  uint32_t uniform_len = 0;
  std::vector<uint32_t> lens;  // empty => uniform
  uint32_t len_of(Key k) const { return lens.empty() ? uniform_len : lens[k]; }
};

inline size_t dtype_size(const std::string& d) {
  if (d == "float32") return 4;
  if (d == "float64" || d == "int64") return 8;
  throw Error("unsupported dtype " + d);
}

inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

// Builds the (rank-independent) heap layout. `key_class_out` receives the class of each key
// when there is more than one length class.
inline Layout build_layout(const ValueSpec& spec, const Options& opt, std::vector<uint8_t>* key_class_out) {
  Layout L;
  memset(&L, 0, sizeof(L));
  L.num_keys = spec.num_keys;
  L.world = opt.world;
  L.workers = opt.workers;
  L.val_bytes = (uint32_t)dtype_size(opt.dtype);
  ADAPM_CHECK(spec.num_keys > 0, "num_keys must be positive");

  std::vector<uint32_t> class_len;
  std::vector<std::vector<int64_t>> home_count;  // [class][rank]
  if (spec.lens.empty()) {
    ADAPM_CHECK(spec.uniform_len > 0, "value length must be positive");
    class_len.push_back(spec.uniform_len);
    home_count.assign(1, std::vector<int64_t>(opt.world, 0));
    for (int r = 0; r < opt.world; ++r)
      home_count[0][r] = spec.num_keys / opt.world + ((spec.num_keys % opt.world) > r ? 1 : 0);
  } else {
    ADAPM_CHECK((int64_t)spec.lens.size() == spec.num_keys, "value_lengths must have one entry per key");
    // Any number of distinct value lengths (reference: one allocation per key, coloc_kv_server_handle.h:162-170,
    // 996-999) maps onto at most MAX_CLASSES size classes of the row slabs. Up to MAX_CLASSES distinct lengths: one
    // exact class each. More: the sorted lengths are cut into MAX_CLASSES groups of bounded relative spread (the
    // smallest ratio that fits, found by bisection); a class's slot holds its longest row (rounded up to 16 bytes), the
    // keys keep their own length in a per-key table (Layout::per_key_len).
    std::map<uint32_t, int> idx;
    for (uint32_t l : spec.lens) idx.emplace(l, 0);
    if ((int)idx.size() <= MAX_CLASSES) {
      int i = 0;
      for (auto& kv : idx) { kv.second = i++; class_len.push_back(kv.first); }
    } else {
      std::vector<uint32_t> d;
      for (auto& kv : idx) d.push_back(kv.first);
      auto groups_for = [&](double ratio, std::vector<size_t>* starts) {
        int g = 0; double lo = 0;
        for (size_t i = 0; i < d.size(); ++i) {
          if (i == 0 || (double)d[i] > lo * ratio) { ++g; lo = (double)d[i]; if (starts) starts->push_back(i); }
        }
        return g;
      };
      double a = 1.0, b = (double)d.back() / (double)d.front() + 1.0;
      for (int it = 0; it < 60; ++it) {
        const double m = 0.5 * (a + b);
        if (groups_for(m, nullptr) <= MAX_CLASSES) b = m; else a = m;
      }
      std::vector<size_t> starts;
      groups_for(b, &starts);
      for (size_t g = 0; g < starts.size(); ++g) {
        const size_t end = g + 1 < starts.size() ? starts[g + 1] : d.size();
        const uint32_t longest = d[end - 1];
        class_len.push_back((longest + 3u) / 4u * 4u);
        for (size_t i = starts[g]; i < end; ++i) idx[d[i]] = (int)g;
      }
      L.per_key_len = 1;
    }
    home_count.assign(class_len.size(), std::vector<int64_t>(opt.world, 0));
    if (key_class_out) key_class_out->resize(spec.num_keys);
    for (int64_t k = 0; k < spec.num_keys; ++k) {
      int c = idx[spec.lens[k]];
      ADAPM_CHECK(spec.lens[k] > 0, "value length of key " << k << " is zero");
      if (key_class_out) (*key_class_out)[k] = (uint8_t)c;
      home_count[c][k % opt.world]++;
    }
  }
  L.num_classes = (int)class_len.size();

  uint32_t slot = 0;
  for (int c = 0; c < L.num_classes; ++c) {
    int64_t n_c = 0, home_max = 0;
    for (int r = 0; r < opt.world; ++r) { n_c += home_count[c][r]; home_max = std::max(home_max, home_count[c][r]); }
    // A key can transiently occupy two slots on a rank (relocation source kept readable for one round while
    // the key is already requested back; worst case: every key), so "every key everywhere" needs slack above n_c.
    const int64_t full = n_c <= (1 << 16) ? 2 * n_c + 64 : n_c + n_c / 4;
    int64_t cap;
    if (opt.world == 1) cap = n_c;
    else if (n_c <= (1 << 16) && opt.pool_factor <= 0) cap = full;
    else if (opt.pool_factor > 0) {
      cap = std::min<int64_t>(full, (int64_t)std::ceil(home_max * opt.pool_factor) + 1024);
    } else {
      // auto: twice the home share, or as many slots as the pool memory budget buys (a long intent look-ahead keeps
      // many replicas / relocated rows alive per rank: with 8 ranks, home is only 1/8 of the keys)
      const int64_t slot_bytes = 2 * (int64_t)class_len[c] * L.val_bytes + 64;
      const int64_t by_budget = (int64_t)(opt.pool_bytes / L.num_classes) / slot_bytes;
      cap = std::min<int64_t>(full, std::max<int64_t>(2 * home_max + 1024, by_budget));
    }
    cap = std::max<int64_t>(cap, std::min<int64_t>(full, opt.min_pool));
    cap = std::max<int64_t>(cap, home_max);
    L.cls[c].len = class_len[c];
    L.cls[c].cap = (uint32_t)cap;
    L.cls[c].slot_begin = slot;
    slot += (uint32_t)cap;
  }
  L.total_slots = slot;

  uint64_t off = 0;
  auto take = [&](uint64_t bytes) { uint64_t o = off; off = align_up(off + bytes, 256); return o; };
  L.off_dir = take((uint64_t)L.num_keys);
  L.off_slot_of = take((uint64_t)L.num_keys * 4);
  L.off_key_class = take(L.num_classes > 1 ? (uint64_t)L.num_keys : 1);
  L.off_key_len = take(L.per_key_len ? (uint64_t)L.num_keys * 4 : 1);
  L.off_meta = take((uint64_t)L.total_slots * 4);
  L.off_version = take((uint64_t)L.total_slots * 4);
  L.off_ver_seen = take((uint64_t)L.total_slots * 4);
  L.off_want = take((uint64_t)L.total_slots * 8);
  L.off_slot_key = take((uint64_t)L.total_slots * 8);
  L.off_intent_end = take((uint64_t)L.total_slots * 8 * L.workers);
  L.off_flags = take((uint64_t)L.total_slots);
  L.off_dirty = take((uint64_t)L.total_slots);
  L.off_want_owner = take((uint64_t)L.total_slots);
  L.off_peer_slot = take((uint64_t)L.total_slots * 4);
  L.off_free_top = take(MAX_CLASSES * 4);
  L.off_counters = take(C_NUM_COUNTERS * 8);
  L.locality_stats = opt.locality_stats ? 1u : 0u;
  L.off_access = opt.locality_stats ? take((uint64_t)L.num_keys * 8) : 0;
  L.off_sync = take(sizeof(SyncArea));
  for (int c = 0; c < L.num_classes; ++c) {
    uint64_t row_bytes = (uint64_t)L.cls[c].len * L.val_bytes;
    L.cls[c].rows_off = take((uint64_t)L.cls[c].cap * row_bytes);
    L.cls[c].base_off = take((uint64_t)L.cls[c].cap * row_bytes);
    L.cls[c].free_off = take((uint64_t)L.cls[c].cap * 4);
  }
  L.heap_bytes = align_up(off, 4096);
  return L;
}

// Result of a batched worker op.
struct OpResult {
  uint64_t n_local = 0;
  uint64_t n_remote = 0;
  uint64_t n_failed = 0;
  uint64_t n_retry = 0;   // Set only: keys whose relocation is in flight - repeat them after the next sync round
};

// Where the caller's key/value buffers live and how the op is ordered.
//   cpu backend : host pointers, the op completes before the call returns.
//   cuda backend: on_device=false -> host pointers, staged through pinned memory, synchronous;
//                 on_device=true  -> device pointers, the op is enqueued on `stream`
//                 (has_stream=false: the worker's own stream) and the call returns a ticket.
struct IoDesc {
  bool on_device = false;
  bool has_stream = false;   // false: use the worker's own stream
  void* stream = nullptr;    // cudaStream_t (0 is the legacy default stream, hence has_stream)
  const int64_t* offsets = nullptr;  // device path on mixed-length stores: value offset of every key (device ptr)
};

// A whole sync round as one request (backends whose rounds run without host sequencing, see has_fused_round).
struct RoundRequest {
  RoundParams rp;            // rp.sweep is decided by the backend (any rank's want_sweep)
  const IntentRec* recs = nullptr;
  size_t n_recs = 0;
  uint8_t* status = nullptr; // [n_recs] 0 registered, 1 deferred, 2 dropped
  bool want_stop = false;    // this rank would like to leave the round loop
  bool want_sweep = false;   // this rank needs a guaranteed-propagation round (WaitSync)
};
struct RoundOutcome {
  bool all_stop = false;     // every rank wanted to stop: nothing was done, leave the loop
  bool any_sweep = false;
};

// One backend instance per rank. Thread-safety: worker ops may be called concurrently from
// several worker threads; the round functions are only called from the rank's sync thread.
class Backend {
 public:
  virtual ~Backend() {}
  virtual const Ctx& ctx() const = 0;
  virtual bool is_cuda() const { return false; }

  // Populate directory + initial allocation (every key at its home rank key % world,
  // reference coloc_kv_server.h:86-90). Collective.
  // key_lens: the per-key row lengths (only read when the layout has per_key_len set)
  virtual void init_store(const std::vector<uint8_t>& key_class, const std::vector<uint32_t>& key_lens) = 0;

  // ---- worker data path. `vals` holds the concatenated rows in key order.
  // host pointers for the cpu backend; for cuda see cuda/cuda_backend.h (device or host).
  // Return value: 0 = completed inline, otherwise a ticket for wait_ticket()/ticket_done().
  virtual uint64_t pull(int worker, const Key* keys, size_t n, void* vals, bool local_only, uint8_t* ok,
                        OpResult* res, const IoDesc& io) = 0;
  // `todo` (Set with retry, host array of n bytes or nullptr): only keys with todo[i] != 0 are processed; completed
  // keys are cleared, keys that must be repeated after the next sync round stay set (res->n_retry of them). With
  // `todo` the op is synchronous on every backend.
  virtual uint64_t push(int worker, const Key* keys, size_t n, const void* vals, bool set, OpResult* res,
                        const IoDesc& io, uint8_t* todo = nullptr) = 0;
  virtual void wait_ticket(uint64_t) {}
  virtual bool ticket_done(uint64_t) { return true; }
  virtual void wait_worker(int /*worker*/) {}   // all ops issued by this worker are complete
  virtual bool key_is_local(Key k) = 0;
  // Debug/tracing: state (SlotState, S_FREE if not resident) and directory owner of some keys.
  virtual void peek_states(const Key* keys, size_t n, uint8_t* state_out, uint8_t* owner_out) = 0;

  // ---- sync round (sync thread only)
  // status[i]: 0 registered, 1 deferred, 2 dropped
  virtual void register_intents(const IntentRec* recs, size_t n, const RoundParams& rp, uint8_t* status) = 0;
  virtual void phase_a(const RoundParams& rp) = 0;
  virtual void phase_b(const RoundParams& rp) = 0;
  virtual void phase_c(const RoundParams& rp) = 0;
  virtual void round_fence() = 0;   // wait until the round work issued so far is complete and visible
  // Grace period: returns when every worker op of THIS rank that started before the call is done.
  virtual void grace() = 0;
  // Device-resident round: stop/sweep agreement, the barriers between the phases and the grace period all happen
  // inside the backend (cuda: on the device, one enqueue per round); the sync thread only feeds intents.
  virtual bool has_fused_round() const { return false; }
  virtual RoundOutcome fused_round(const RoundRequest&) { throw Error("this backend has no fused round"); }
  virtual void read_counters(uint64_t* out) = 0;
  virtual void reset_counters() = 0;
  // copy `bytes` at heap offset `off` of THIS rank into host memory (statistics / debugging)
  virtual void read_heap(uint64_t off, void* dst, size_t bytes) = 0;
  // Kernel timeline (ADAPM_SYNC_TRACE=1, cuda backend): mark a point on a stream / write all records as TSV.
  virtual void trace_mark(const char* /*name*/, void* /*stream*/) {}
  virtual void dump_trace(const std::string& /*path*/) {}
};

std::unique_ptr<Backend> make_cpu_backend(const Options& opt, const Layout& L, std::shared_ptr<Fabric> fabric);
std::unique_ptr<Backend> make_cuda_backend(const Options& opt, const Layout& L, std::shared_ptr<Fabric> fabric);

}  // namespace adapm
