// Server (one per rank / GPU), Worker (the ColoKVWorker-equivalent API), SyncEngine
// (the SyncManager-equivalent round driver) and Sampling.
//
// API parity (what a user of the reference finds here):
//   ColoKVServer   coloc_kv_server.h:59-477   -> Server
//   ColoKVWorker   coloc_kv_worker.h:76-921   -> Worker
//   SyncManager    sync_manager.h:162-824     -> SyncEngine
//   Sampling       sampling.h:32-534          -> Sampling (+ KeyDistribution)
//   Addressbook    addressbook.h:30-176       -> Server::owner_of / home_of (replicated dir)
#pragma once
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <queue>
#include <random>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include "action_timer.h"
#include "rpc.h"
#include "store.h"

namespace adapm {

class Server;
class Worker;

struct FutureIntent {
  Clock start, end;
  int worker;
  std::shared_ptr<std::vector<Key>> keys;
};
struct IntentLater {
  bool operator()(const FutureIntent& a, const FutureIntent& b) const { return a.start > b.start; }
};

// Key tracing events (reference coloc_kv_server_handle.h:86-104, PS_TRACE_KEYS build).
enum class TraceEvent : int { ALLOC = 0, DEALLOC, REPLICA_SETUP, REPLICA_DROP, INTENT_START, INTENT_STOP };

// -------------------------------------------------------------------------------------
class SyncEngine {
 public:
  SyncEngine(Server* server);
  ~SyncEngine();
  void start();
  void request_stop_and_join();   // collective: returns once all ranks agreed to stop

  void enqueue(FutureIntent&& fi);
  // key buffers of processed intents are recycled (a fresh 512 KB vector per batch would page-fault every step)
  std::shared_ptr<std::vector<Key>> acquire_key_buffer();
  uint64_t rounds_done() const;
  // Block until one complete round that started after this call has finished on this rank
  // (reference WaitSync, coloc_kv_worker.h:517-550). Requests guaranteed propagation.
  void wait_sync();
  std::string report() const;
  float avg_clocks_per_round() const { return timer_.avg_estimate(); }

 private:
  void loop();
  void loop_fused();
  void pace(bool device_round);
  void round(bool sweep);
  bool fused_ = false;
  void collect_intents(const std::vector<Clock>& clocks, const std::vector<Clock>& windows);

  Server* server_;
  std::thread thread_;
  ActionTimer timer_;
  std::mutex in_mu_;
  std::deque<FutureIntent> incoming_;
  std::mutex pool_mu_;
  std::vector<std::shared_ptr<std::vector<Key>>> key_pool_;
  std::vector<std::priority_queue<FutureIntent, std::vector<FutureIntent>, IntentLater>> heaps_;  // per worker
  std::vector<IntentRec> recs_, deferred_;
  std::vector<FutureIntent> due_;
  std::vector<uint32_t> seen_epoch_;   // intent dedupe table (one stamp per key)
  uint32_t epoch_ = 0;
  std::vector<uint64_t> rec_tab_;   // per-round merge of records per key (single worker per rank): stamp << 32 | index
  uint32_t rec_epoch_ = 0;
  std::atomic<uint64_t> deferred_pending_{0};  // intent records that could not be registered in the last round
  std::vector<uint8_t> status_;
  std::atomic<uint64_t> round_no_{0};   // written by the sync thread, read by report()
  std::chrono::steady_clock::time_point last_run_;
  Clock last_round_clock_ = 0;          // fastest worker clock when the previous round started (loop_fused pacing)
  Stopwatch sw_total_, sw_pausing_, sw_register_, sw_collect_, sw_phase_a_, sw_phase_b_, sw_grace_, sw_phase_c_, sw_barriers_;
  std::atomic<uint64_t> intents_seen_{0}, recs_registered_{0};
};

// -------------------------------------------------------------------------------------
// Key distributions for sampling (reference: app-provided `Key (*sample_key)()`,
// bindings.cc:60-76 uniform / log-uniform). `custom` wraps a C++ callback; `weights`
// builds an alias table (also used by the device sampler).
class KeyDistribution {
 public:
  virtual ~KeyDistribution() {}
  virtual Key draw(std::mt19937_64& rng) = 0;
  Key min_key = 0, max_key = 0;   // declared contiguous range [min,max) (0,0 = unknown)
  std::string name;
};
std::shared_ptr<KeyDistribution> make_uniform_distribution(Key min, Key max);
std::shared_ptr<KeyDistribution> make_log_uniform_distribution(Key min, Key max);
std::shared_ptr<KeyDistribution> make_alias_distribution(const double* weights, int64_t n, Key first_key, Key key_stride);
std::shared_ptr<KeyDistribution> make_callback_distribution(std::function<Key()> fn, Key min, Key max);
// Walker/Vose alias table: prob[i] in [0,1], alias[i] in [0,n)  (shared with the device sampler)
void build_alias_table(const double* weights, int64_t n, float* prob, int32_t* alias);

class Sampling {
 public:
  Sampling(Server* server, std::shared_ptr<KeyDistribution> dist, const std::string& scheme, bool with_replacement);
  SampleID prepare_sample(size_t K, int worker, Clock start, Clock end);
  // Fills `keys[0..n)` and pulls their values into vals; returns an op timestamp (LOCAL = -1).
  int pull_sample(SampleID id, Key* keys, size_t n, void* vals, Worker& w);
  void finish_sample(SampleID id, int worker);
  const std::string& scheme() const { return scheme_; }
  uint64_t local_checks() const { return checks_; }
  uint64_t local_pulls() const { return pulls_; }

 private:
  struct Sample { std::vector<Key> keys; size_t used = 0; size_t K = 0; };
  void draw(Key* out, size_t n, std::mt19937_64& rng);
  void draw_from_pool(Key* out, size_t n);
  Key next_local(int worker, Worker& w, void* vals, std::unordered_set<Key>* exclude);

  Server* server_;
  std::shared_ptr<KeyDistribution> dist_;
  std::string scheme_;
  bool with_replacement_;
  std::vector<std::mutex> mu_;                                        // per worker
  std::vector<std::unordered_map<SampleID, Sample>> samples_;        // per worker
  std::vector<SampleID> id_counter_;
  std::vector<std::mt19937_64> rng_;
  std::vector<std::deque<Key>> predrawn_;                             // local scheme
  std::vector<std::unordered_map<SampleID, std::unordered_set<Key>>> used_wor_;  // local WOR
  // pool scheme
  std::mutex pool_mu_;
  std::vector<Key> pool_;
  size_t pool_pos_ = 0, pool_uses_ = 0;
  std::mt19937_64 pool_rng_;
  std::atomic<uint64_t> checks_{0}, pulls_{0};
};

// -------------------------------------------------------------------------------------
class Server {
 public:
  Server(const Options& opt, const ValueSpec& spec);
  ~Server();

  // reference ColoKVServer::enable_sampling_support (coloc_kv_server.h:177-197)
  void enable_sampling_support(std::shared_ptr<KeyDistribution> dist, const std::string& scheme = "",
                               int with_replacement = -1);
  void barrier();    // among servers (one call per rank)
  // Sum-all-reduce of up to 64 doubles over the ranks through the control block (collective, one call per rank):
  // the host-side counterpart of ps_allreduce for values that must not live in a model key (losses, counters).
  void allreduce_sum(double* vals, int n);
  void shutdown();   // collective
  int my_rank() const { return opt_.rank; }
  int num_servers() const { return opt_.world; }
  int num_workers() const { return opt_.workers; }
  int64_t num_keys() const { return spec_.num_keys; }
  size_t get_len(Key k) const { return spec_.len_of(k); }
  const Options& options() const { return opt_; }
  Backend& backend() { return *backend_; }
  Fabric& fabric() { return *fabric_; }
  SyncEngine& sync() { return *sync_; }
  Sampling* sampling() { return sampling_.get(); }
  ControlBlock* control() { return fabric_->control(); }
  RankControl& my_control() { return fabric_->control()->ranks[opt_.rank]; }
  MailRouter& router();                        // host RPC (rpc.h); created on first use
  MailRouter* router_if_any() { return router_.get(); }

  // Addressbook-style queries (addressbook.h:50-112)
  int owner_of(Key k);                 // this rank's view of the current owner
  int home_of(Key k) const { return (int)(k % opt_.world); }
  bool is_local(Key k) { return backend_->key_is_local(k); }

  std::vector<Clock> worker_clocks();  // WORKER_FINISHED for finalized workers
  void register_worker(int id, Worker* w);
  void deregister_worker(int id);
  void worker_barrier();

  // statistics (reference coloc_kv_server.h:128-166 shutdown report)
  std::map<std::string, uint64_t> counters();
  std::string stats_string();
  void reset_stats();

  // key tracing (PS_TRACE_KEYS equivalent, run-time switch via sys.trace.keys)
  void trace(Key k, TraceEvent e);
  bool tracing() const { return trace_all_ || !traced_.empty(); }
  void write_traces();
  void write_locality_stats();   // locality_stats.rank.<r>.tsv (reference coloc_kv_server_handle.h:963-975)
  void observe_traced_keys();    // called by the sync thread after every round

 private:
  friend class Worker;
  friend class SyncEngine;
  friend class Sampling;
  void parse_trace_keys();

  Options opt_;
  ValueSpec spec_;
  std::shared_ptr<Fabric> fabric_;
  std::unique_ptr<Backend> backend_;
  std::unique_ptr<SyncEngine> sync_;
  std::unique_ptr<Sampling> sampling_;
  std::unique_ptr<MailRouter> router_;
  std::mutex mu_;
  std::vector<Worker*> workers_;
  bool shut_down_ = false;
  bool trace_all_ = false;
  std::unordered_set<Key> traced_;
  std::mutex trace_mu_;
  std::vector<std::tuple<int64_t, Key, int>> trace_log_;
  std::vector<Key> traced_list_;
  std::vector<uint8_t> traced_prev_;
};

// -------------------------------------------------------------------------------------
class Worker {
 public:
  Worker(int customer_id, Server& server);
  ~Worker();

  // Data ops. Return LOCAL (-1) when every key was served from local memory, else a timestamp
  // for Wait()/IsFinished() (reference coloc_kv_worker.h:120-186,253-318).
  int Push(const Key* keys, size_t n, const void* vals, bool set = false, const IoDesc& io = IoDesc());
  int Set(const Key* keys, size_t n, const void* vals, const IoDesc& io = IoDesc()) { return Push(keys, n, vals, true, io); }
  int Pull(const Key* keys, size_t n, void* vals, const IoDesc& io = IoDesc());
  bool PullIfLocal(Key key, void* vals);

  // Intent signalling (coloc_kv_worker.h:380-408); end == 0 means [start, start+1).
  int Intent(const Key* keys, size_t n, Clock start, Clock end = 0);
  int Intent(const std::vector<Key>& keys, Clock start, Clock end = 0) { return Intent(keys.data(), keys.size(), start, end); }
  int Intent(Key key, Clock start, Clock end = 0) { return Intent(&key, 1, start, end); }
  // Intent with the pre-pass done on the calling thread (CPU backend; the CUDA twin is cuda/ops_intent.cu): keys that
  // already have a usable local slot only get their end clock extended, the rest goes through Intent(). Returns the
  // number of keys that took the fast path.
  size_t IntentFast(const Key* keys, size_t n, Clock start, Clock end = 0);

  Clock advanceClock();
  Clock currentClock() const;

  SampleID PrepareSample(size_t K, Clock start, Clock end = 0);
  int PullSample(SampleID id, Key* keys, size_t n, void* vals);
  void FinishSample(SampleID id);

  void Wait(int ts);
  bool IsFinished(int ts);
  void WaitAll();
  void WaitSync();
  void Barrier();
  void BeginSetup();
  void EndSetup();
  void ResetStats();
  void Finalize();
  int StaggeredPush(const Key* keys, size_t n, const void* vals, size_t group_size = 100000);

  size_t GetLen(Key key) const { return server_.get_len(key); }
  int64_t GetNumKeys() const { return server_.num_keys(); }
  int id() const { return id_; }
  int worker_id() const { return server_.my_rank() * server_.num_workers() + id_; }
  Server& server() { return server_; }
  size_t total_len(const Key* keys, size_t n) const;

  // locality statistics (reference: num_pull_ops_local etc.)
  uint64_t num_pull_ops = 0, num_pull_ops_local = 0, num_push_ops = 0, num_push_ops_local = 0;
  uint64_t num_pull_params = 0, num_pull_params_local = 0, num_push_params = 0, num_push_params_local = 0;

 private:
  int new_ts(uint64_t ticket);
  Server& server_;
  int id_;
  bool finalized_ = false;
  std::mutex ts_mu_;
  std::vector<uint64_t> tickets_;   // ts -> backend ticket (0 = done)
};

}  // namespace adapm
