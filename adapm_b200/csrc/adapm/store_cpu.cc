// CPU executor of the protocol (protocol.h) over host-mapped heaps.
// Used for the contract tests, for BASELINE config #1 (apps/simple at world_size=2 on CPU)
// and as the reference implementation the CUDA kernels are tested against.
#include "store.h"

namespace adapm {

namespace {

template <class Val>
class CpuBackend : public Backend {
 public:
  CpuBackend(const Options& opt, const Layout& L, std::shared_ptr<Fabric> fabric) : fabric_(fabric) {
    memset(&ctx_, 0, sizeof(ctx_));
    ctx_.L = L;
    ctx_.rank = opt.rank;
    ctx_.technique = (int)opt.techniques;
    fabric_->allocate_heaps(L.heap_bytes);
    for (int r = 0; r < L.world; ++r) ctx_.heap[r] = fabric_->heap(r);
    for (int w = 0; w < MAX_LOCAL_WORKERS; ++w) op_seq_[w].store(0);
  }

  const Ctx& ctx() const override { return ctx_; }

  void init_store(const std::vector<uint8_t>& key_class, const std::vector<uint32_t>& key_lens) override {
    const Ctx& c = ctx_;
    const int me = c.rank;
    const Layout& L = c.L;
    uint8_t* dir = dir_of(c, me);
    int32_t* so = slot_of(c, me);
    if (L.num_classes > 1) memcpy(at<uint8_t>(c, me, L.off_key_class), key_class.data(), (size_t)L.num_keys);
    if (L.per_key_len) memcpy(at<uint32_t>(c, me, L.off_key_len), key_lens.data(), (size_t)L.num_keys * 4);
    std::vector<uint32_t> next(L.num_classes);
    for (int k = 0; k < L.num_classes; ++k) next[k] = L.cls[k].slot_begin;
    for (int64_t key = 0; key < L.num_keys; ++key) {
      int home = (int)(key % L.world);
      dir[key] = (uint8_t)home;
      if (home == me) {
        int cl = class_of_key(c, key);
        uint32_t s = next[cl]++;
        ADAPM_CHECK(s < L.cls[cl].slot_begin + L.cls[cl].cap, "pool too small for home keys");
        so[key] = (int32_t)s;
        meta_of(c, me)[s] = meta_make(S_OWNED, 0, 1);
        slot_key_of(c, me)[s] = key;
      } else {
        so[key] = -1;
      }
    }
    for (int k = 0; k < L.num_classes; ++k) {
      int32_t* stack = at<int32_t>(c, me, L.cls[k].free_off);
      int32_t n = 0;
      uint32_t end = L.cls[k].slot_begin + L.cls[k].cap;
      for (uint32_t s = end; s-- > next[k];) stack[n++] = (int32_t)s;  // pop order = ascending slots
      free_top_of(c, me)[k] = n;
    }
    mem::fence();
    fabric_->node_barrier("init_store");
  }

  struct OpGuard {
    std::atomic<uint64_t>& s;
    explicit OpGuard(std::atomic<uint64_t>& seq) : s(seq) { s.fetch_add(1, std::memory_order_acq_rel); }
    ~OpGuard() { s.fetch_add(1, std::memory_order_acq_rel); }
  };

  uint64_t pull(int worker, const Key* keys, size_t n, void* vals, bool local_only, uint8_t* ok, OpResult* res,
                const IoDesc& io) override {
    ADAPM_CHECK(!io.on_device, "the cpu backend takes host pointers");
    OpGuard guard(op_seq_[worker]);
    HostGroup g;
    Val* out = reinterpret_cast<Val*>(vals);
    uint64_t nl = 0, nr = 0, nf = 0;
    for (size_t i = 0; i < n; ++i) {
      const Key key = keys[i];
      ADAPM_CHECK(key >= 0 && key < ctx_.L.num_keys, "[ERROR] Pull key " << key << ", which is outside the configured key range [0," << ctx_.L.num_keys << ")");
      bool local = false;
      bool good = pull_key<Val>(ctx_, g, key, out, local_only, &local);
      if (ok) ok[i] = good ? 1 : 0;
      if (!good) ++nf; else if (local) ++nl; else ++nr;
      out += key_len(ctx_, key, class_of_key(ctx_, key));
    }
    count(ctx_, C_PULL_LOCAL, nl);
    count(ctx_, C_PULL_REMOTE, nr);
    if (res) { res->n_local = nl; res->n_remote = nr; res->n_failed = nf; }
    return 0;
  }

  uint64_t push(int worker, const Key* keys, size_t n, const void* vals, bool set, OpResult* res,
                const IoDesc& io, uint8_t* todo) override {
    ADAPM_CHECK(!io.on_device, "the cpu backend takes host pointers");
    OpGuard guard(op_seq_[worker]);
    HostGroup g;
    const Val* in = reinterpret_cast<const Val*>(vals);
    uint64_t nl = 0, nr = 0, nf = 0, nretry = 0;
    for (size_t i = 0; i < n; ++i) {
      const Key key = keys[i];
      ADAPM_CHECK(key >= 0 && key < ctx_.L.num_keys, "[ERROR] Push key " << key << ", which is outside the configured key range [0," << ctx_.L.num_keys << ")");
      bool local = false;
      if (!todo || todo[i]) {
        int good = set ? set_key<Val>(ctx_, g, key, in, &local) : (push_key<Val>(ctx_, g, key, in, &local) ? SET_OK : SET_FAIL);
        if (good == SET_RETRY && todo) ++nretry;
        else {
          if (todo) todo[i] = 0;
          if (good != SET_OK) ++nf; else if (local) ++nl; else ++nr;
        }
      }
      in += key_len(ctx_, key, class_of_key(ctx_, key));
    }
    mem::fence();
    count(ctx_, C_PUSH_LOCAL, nl);
    count(ctx_, C_PUSH_REMOTE, nr);
    if (res) { res->n_local = nl; res->n_remote = nr; res->n_failed = nf; res->n_retry = nretry; }
    return 0;
  }

  bool key_is_local(Key k) override { return is_local(ctx_, k); }
  void peek_states(const Key* keys, size_t n, uint8_t* state_out, uint8_t* owner_out) override {
    for (size_t i = 0; i < n; ++i) {
      int32_t s = mem::ld_relaxed(slot_of(ctx_, ctx_.rank) + keys[i]);
      state_out[i] = s >= 0 ? (uint8_t)meta_state(mem::ld_acquire(meta_of(ctx_, ctx_.rank) + s)) : (uint8_t)S_FREE;
      owner_out[i] = mem::ld_relaxed(dir_of(ctx_, ctx_.rank) + keys[i]);
    }
  }

  void register_intents(const IntentRec* recs, size_t n, const RoundParams& rp, uint8_t* status) override {
    uint64_t reg = 0, def = 0;
    for (size_t i = 0; i < n; ++i) {
      int st = register_intent<Val>(ctx_, recs[i], rp.clocks);
      status[i] = (uint8_t)st;
      if (st == 0) ++reg; else if (st == 1) ++def;
    }
    count(ctx_, C_INTENTS_REGISTERED, reg);
    count(ctx_, C_INTENTS_DEFERRED, def);
  }
  void phase_a(const RoundParams& rp) override {
    HostGroup g;
    const uint32_t S = ctx_.L.total_slots;
    for (uint32_t s = 0; s < S; ++s)
      if (phase_a_wants(ctx_, s, rp)) {
        SlotWork w;
        w.slot = s;
        phase_a_resolve<Val>(ctx_, w, rp);
        if (w.op != OP_NONE) { mem::fence(); row_op_execute<Val>(ctx_, g, w); mem::fence(); }
        phase_a_commit<Val>(ctx_, w);
      }
    mem::fence();
  }
  void phase_b(const RoundParams& rp) override {
    const uint32_t S = ctx_.L.total_slots;
    for (uint32_t s = 0; s < S; ++s)
      if (phase_b_candidate(ctx_, s)) phase_b_slot(ctx_, s, rp);
    mem::fence();
  }
  void phase_c(const RoundParams& rp) override {
    HostGroup g;
    const uint32_t S = ctx_.L.total_slots;
    for (uint32_t s = 0; s < S; ++s)
      if (phase_c_wants(ctx_, s, rp)) {
        SlotWork w;
        w.slot = s;
        phase_c_resolve<Val>(ctx_, w, rp);
        if (w.op != OP_NONE) { mem::fence(); row_op_execute<Val>(ctx_, g, w); mem::fence(); }
        phase_c_commit<Val>(ctx_, w);
      }
    mem::fence();
  }
  void round_fence() override { mem::fence(); }
  void grace() override {
    for (int w = 0; w < ctx_.L.workers; ++w) {
      uint64_t s = op_seq_[w].load(std::memory_order_acquire);
      if (s & 1) {
        int spins = 0;
        while (op_seq_[w].load(std::memory_order_acquire) == s) {
          if (++spins < 100) std::this_thread::yield();
          else std::this_thread::sleep_for(std::chrono::microseconds(20));
        }
      }
    }
  }
  void read_counters(uint64_t* out) override {
    const uint64_t* c = counters_of(ctx_, ctx_.rank);
    for (int i = 0; i < C_NUM_COUNTERS; ++i) out[i] = mem::ld_relaxed(c + i);
  }
  void read_heap(uint64_t off, void* dst, size_t bytes) override { memcpy(dst, ctx_.heap[ctx_.rank] + off, bytes); }
  void reset_counters() override {
    uint64_t* c = counters_of(ctx_, ctx_.rank);
    for (int i = 0; i < C_NUM_COUNTERS; ++i) mem::st_relaxed(c + i, (uint64_t)0);
  }

 private:
  std::shared_ptr<Fabric> fabric_;
  Ctx ctx_;
  std::atomic<uint64_t> op_seq_[MAX_LOCAL_WORKERS];
};

}  // namespace

std::unique_ptr<Backend> make_cpu_backend(const Options& opt, const Layout& L, std::shared_ptr<Fabric> fabric) {
  if (opt.dtype == "float32") return std::unique_ptr<Backend>(new CpuBackend<float>(opt, L, fabric));
  if (opt.dtype == "float64") return std::unique_ptr<Backend>(new CpuBackend<double>(opt, L, fabric));
  if (opt.dtype == "int64") return std::unique_ptr<Backend>(new CpuBackend<int64_t>(opt, L, fabric));
  throw Error("unsupported dtype " + opt.dtype);
}

}  // namespace adapm
