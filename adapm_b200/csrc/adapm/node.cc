#include "node.h"
#include "nvtx.h"

#include <fstream>
#include <sstream>

namespace adapm {

// ======================================================================== Server
Server::Server(const Options& opt, const ValueSpec& spec) : opt_(opt), spec_(spec) {
  ADAPM_CHECK(opt_.backend == "cpu" || opt_.backend == "cuda", "backend must be cpu or cuda");
  // (cuda: float32 rows for the fused application kernels; float64 / int64 rows - the reference's `double` apps and its
  //  exact `long` contract tests - run through the generic Pull/Push/Set and the sync round)
  fabric_ = Fabric::create(opt_);
  std::vector<uint8_t> key_class;
  Layout L = build_layout(spec_, opt_, &key_class);
  VLOG(1, "rank " << opt_.rank << ": heap " << (L.heap_bytes >> 20) << " MiB, " << L.total_slots << " slots, "
                  << L.num_classes << " length classes");
  if (opt_.backend == "cuda") backend_ = make_cuda_backend(opt_, L, fabric_);
  else backend_ = make_cpu_backend(opt_, L, fabric_);
  backend_->init_store(key_class, spec_.lens);
  workers_.assign(opt_.workers, nullptr);
  parse_trace_keys();
  sync_.reset(new SyncEngine(this));
  sync_->start();
  fabric_->node_barrier("server start");
  fabric_->start_failure_detector();
}

Server::~Server() {
  try {
    if (!shut_down_) shutdown();
  } catch (const std::exception& e) {
    ALOG("[adapm] error during shutdown: " << e.what());
  }
}

void Server::enable_sampling_support(std::shared_ptr<KeyDistribution> dist, const std::string& scheme, int with_replacement) {
  std::string sch = scheme.empty() ? opt_.sampling_scheme : scheme;
  bool wr = with_replacement < 0 ? opt_.sampling_with_replacement : (with_replacement != 0);
  sampling_.reset(new Sampling(this, dist, sch, wr));
}

void Server::barrier() { fabric_->node_barrier("Server::barrier"); }

void Server::allreduce_sum(double* vals, int n) {
  ADAPM_CHECK(n >= 0 && n <= 64, "allreduce_sum: at most 64 values");
  if (opt_.world == 1 || n == 0) return;
  ControlBlock* ctl = control();
  fabric_->node_barrier("allreduce: enter");
  if (opt_.rank == 0)
    for (int i = 0; i < n; ++i) ctl->allreduce_buf[i].store(0, std::memory_order_relaxed);   // bit pattern of +0.0
  fabric_->node_barrier("allreduce: cleared");
  for (int i = 0; i < n; ++i) {
    int64_t old = ctl->allreduce_buf[i].load(std::memory_order_relaxed);
    for (;;) {   // atomic add on the bit pattern of a double
      double cur;
      memcpy(&cur, &old, 8);
      cur += vals[i];
      int64_t nw;
      memcpy(&nw, &cur, 8);
      if (ctl->allreduce_buf[i].compare_exchange_weak(old, nw, std::memory_order_acq_rel)) break;
    }
  }
  fabric_->node_barrier("allreduce: added");
  for (int i = 0; i < n; ++i) {
    const int64_t bits = ctl->allreduce_buf[i].load(std::memory_order_acquire);
    memcpy(&vals[i], &bits, 8);
  }
  fabric_->node_barrier("allreduce: read");
}

MailRouter& Server::router() {
  std::lock_guard<std::mutex> lk(mu_);
  ADAPM_CHECK(!shut_down_, "rpc: the server is shut down");
  if (!router_) router_.reset(new MailRouter(this));
  return *router_;
}

void Server::worker_barrier() {
  control()->worker_barrier.wait(opt_.world * opt_.workers, opt_.wait_timeout_s, "Worker::Barrier");
}

void Server::shutdown() {
  if (shut_down_) return;
  shut_down_ = true;
  sync_->request_stop_and_join();
  fabric_->node_barrier("shutdown");
  if (verbosity() >= 1) ALOG(stats_string());
  if (tracing()) write_traces();
  if (opt_.locality_stats) write_locality_stats();
  fabric_->stop_failure_detector();   // from here on peers may exit at any time
  fabric_->node_barrier("shutdown done");
  if (router_) router_->stop();   // after the barrier: every peer's requests have been answered
  backend_.reset();
  fabric_.reset();
}

int Server::owner_of(Key k) {
  uint8_t st, ow;
  backend_->peek_states(&k, 1, &st, &ow);
  return ow;
}

std::vector<Clock> Server::worker_clocks() {
  std::vector<Clock> c(opt_.workers);
  RankControl& rc = my_control();
  for (int w = 0; w < opt_.workers; ++w) {
    int st = rc.worker_state[w].load(std::memory_order_acquire);
    c[w] = st == 2 ? WORKER_FINISHED : rc.worker_clock[w].load(std::memory_order_acquire);
  }
  return c;
}

void Server::register_worker(int id, Worker* w) {
  ADAPM_CHECK(id >= 0 && id < opt_.workers, "customer id " << id << " out of range [0," << opt_.workers << ")");
  std::lock_guard<std::mutex> lk(mu_);
  workers_[id] = w;
  my_control().worker_clock[id].store(0);
  my_control().worker_state[id].store(1, std::memory_order_release);
}
void Server::deregister_worker(int id) {
  std::lock_guard<std::mutex> lk(mu_);
  workers_[id] = nullptr;
  my_control().worker_state[id].store(2, std::memory_order_release);
}

std::map<std::string, uint64_t> Server::counters() {
  uint64_t c[C_NUM_COUNTERS];
  backend_->read_counters(c);
  std::map<std::string, uint64_t> m;
  m["pull_local"] = c[C_PULL_LOCAL]; m["pull_remote"] = c[C_PULL_REMOTE];
  m["push_local"] = c[C_PUSH_LOCAL]; m["push_remote"] = c[C_PUSH_REMOTE];
  m["relocations"] = c[C_RELOCATIONS]; m["replica_setups"] = c[C_REPLICA_SETUPS];
  m["replica_drops"] = c[C_REPLICA_DROPS]; m["refreshes"] = c[C_REFRESHES];
  m["deltas_shipped"] = c[C_DELTAS_SHIPPED]; m["intents_registered"] = c[C_INTENTS_REGISTERED];
  m["intents_deferred"] = c[C_INTENTS_DEFERRED]; m["alloc_fail"] = c[C_ALLOC_FAIL];
  m["protocol_errors"] = c[C_PROTOCOL_ERRORS];
  m["sync_rounds"] = sync_ ? sync_->rounds_done() : 0;
  return m;
}

std::string Server::stats_string() {
  auto m = counters();
  std::ostringstream os;
  auto pct = [](uint64_t a, uint64_t b) { return (a + b) ? 100.0 * (double)a / (double)(a + b) : 100.0; };
  os << "[rank " << opt_.rank << "] local pulls " << pct(m["pull_local"], m["pull_remote"]) << "% of "
     << (m["pull_local"] + m["pull_remote"]) << " params, local pushes " << pct(m["push_local"], m["push_remote"])
     << "% of " << (m["push_local"] + m["push_remote"]) << " params; relocations " << m["relocations"]
     << ", replica setups " << m["replica_setups"] << ", drops " << m["replica_drops"] << ", refreshes "
     << m["refreshes"] << ", deltas " << m["deltas_shipped"] << ", sync rounds " << m["sync_rounds"]
     << ", protocol errors " << m["protocol_errors"];
  if (sync_) os << "\n" << sync_->report();
  return os.str();
}

void Server::reset_stats() { backend_->reset_counters(); }

// sys.trace.keys "12,34" | all | random-N-seed-S-range-A-B   (coloc_kv_server_handle.h:213-255)
void Server::parse_trace_keys() {
  const std::string& s = opt_.trace_keys;
  if (s.empty()) return;
  if (s == "all") { trace_all_ = true; return; }
  if (s.rfind("random-", 0) == 0) {
    std::vector<std::string> tok;
    std::stringstream ss(s);
    std::string t;
    while (std::getline(ss, t, '-')) tok.push_back(t);
    ADAPM_CHECK(tok.size() >= 2, "bad sys.trace.keys spec " << s);
    int64_t n = std::stoll(tok[1]);
    uint64_t seed = 0;
    Key a = 0, b = spec_.num_keys;
    for (size_t i = 2; i + 1 < tok.size(); ++i) {
      if (tok[i] == "seed") seed = std::stoull(tok[i + 1]);
      if (tok[i] == "range" && i + 2 < tok.size()) { a = std::stoll(tok[i + 1]); b = std::stoll(tok[i + 2]); }
    }
    std::mt19937_64 rng(seed);
    std::uniform_int_distribution<Key> d(a, b - 1);
    while ((int64_t)traced_.size() < std::min<int64_t>(n, b - a)) traced_.insert(d(rng));
    return;
  }
  std::stringstream ss(s);
  std::string t;
  while (std::getline(ss, t, ',')) if (!t.empty()) traced_.insert(std::stoll(t));
}

void Server::trace(Key k, TraceEvent e) {
  auto now = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
  std::lock_guard<std::mutex> lk(trace_mu_);
  trace_log_.emplace_back((int64_t)now, k, (int)e);
}

void Server::write_traces() {
  static const char* names[] = {"ALLOC", "DEALLOC", "REPLICA_SETUP", "REPLICA_DROP", "INTENT_START", "INTENT_STOP"};
  std::string dir = opt_.stats_out.empty() ? std::string(".") : opt_.stats_out;
  std::string fn = dir + "/traces." + std::to_string(opt_.rank) + ".tsv";
  std::ofstream f(fn, std::ofstream::trunc);
  std::lock_guard<std::mutex> lk(trace_mu_);
  for (auto& t : trace_log_) f << std::get<0>(t) << "\t" << std::get<1>(t) << "\t" << opt_.rank << "\t" << names[std::get<2>(t)] << "\n";
  ALOG("Wrote " << trace_log_.size() << " key trace events to " << fn);
}

void Server::write_locality_stats() {
  const Layout& L = backend_->ctx().L;
  if (!L.off_access) return;
  std::vector<uint32_t> acc((size_t)L.num_keys * 2);
  backend_->read_heap(L.off_access, acc.data(), acc.size() * 4);
  std::string dir = opt_.stats_out.empty() ? std::string(".") : opt_.stats_out;
  std::string fn = dir + "/locality_stats.rank." + std::to_string(opt_.rank) + ".tsv";
  std::ofstream f(fn, std::ofstream::trunc);
  f << "Param\tAccesses\tLocalAccesses\n";
  uint64_t tot = 0, loc = 0;
  for (int64_t k = 0; k < L.num_keys; ++k) {
    if (acc[2 * k]) f << k << "\t" << acc[2 * k] << "\t" << acc[2 * k + 1] << "\n";
    tot += acc[2 * k]; loc += acc[2 * k + 1];
  }
  ALOG("Wrote locality stats for rank " << opt_.rank << " to " << fn << ": " << loc << " of " << tot << " accesses were local");
}

// Key tracing by observation: after every sync round the states of the traced keys are compared with the
// previous round and the transitions are logged as the reference's events
// (ALLOC, DEALLOC, REPLICA_SETUP, REPLICA_DROP; INTENT_START is logged when the intent is acted upon).
void Server::observe_traced_keys() {
  if (traced_list_.empty()) {
    if (trace_all_) {
      int64_t n = std::min<int64_t>(spec_.num_keys, 1 << 20);
      for (Key k = 0; k < n; ++k) traced_list_.push_back(k);
    } else {
      traced_list_.assign(traced_.begin(), traced_.end());
    }
    traced_prev_.assign(traced_list_.size(), 0xff);
  }
  std::vector<uint8_t> st(traced_list_.size()), ow(traced_list_.size());
  backend_->peek_states(traced_list_.data(), traced_list_.size(), st.data(), ow.data());
  auto resident = [](uint8_t s) { return s == S_OWNED || s == S_REPLICA || s == S_REPLICA_PENDING || state_is_incoming(s) || s == S_FINALIZING; };
  for (size_t i = 0; i < st.size(); ++i) {
    uint8_t p = traced_prev_[i], s = st[i];
    traced_prev_[i] = s;
    if (p == 0xff) { if (s == S_OWNED) trace(traced_list_[i], TraceEvent::ALLOC); continue; }
    if (p == s) continue;
    const Key k = traced_list_[i];
    if (!resident(p) && resident(s)) trace(k, TraceEvent::ALLOC);
    if (s == S_REPLICA && p != S_REPLICA) trace(k, TraceEvent::REPLICA_SETUP);
    if ((p == S_REPLICA || p == S_REPLICA_PENDING) && !resident(s)) { trace(k, TraceEvent::INTENT_STOP); trace(k, TraceEvent::REPLICA_DROP); }
    if (resident(p) && !resident(s)) trace(k, TraceEvent::DEALLOC);
  }
}

// ======================================================================== Worker
Worker::Worker(int customer_id, Server& server) : server_(server), id_(customer_id) {
  server_.register_worker(id_, this);
}
Worker::~Worker() {
  if (!finalized_) server_.deregister_worker(id_);
}

size_t Worker::total_len(const Key* keys, size_t n) const {
  if (server_.spec_.lens.empty()) return n * server_.spec_.uniform_len;
  size_t t = 0;
  for (size_t i = 0; i < n; ++i) t += server_.spec_.lens[keys[i]];
  return t;
}

int Worker::new_ts(uint64_t ticket) {
  std::lock_guard<std::mutex> lk(ts_mu_);
  tickets_.push_back(ticket);
  return (int)tickets_.size() - 1;
}

int Worker::Push(const Key* keys, size_t n, const void* vals, bool set, const IoDesc& io) {
  ADAPM_NVTX("adapm::Push");
  ++num_push_ops;
  num_push_params += n;
  OpResult res;
  uint64_t ticket = 0;
  if (set && server_.num_servers() > 1) {
    // Set cannot complete for a key whose relocation is in flight; such keys are repeated after the next sync round
    // (never wait for the round inside the op: the round's grace period waits for the op)
    std::vector<uint8_t> todo(n, 1);
    OpResult acc;
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
      const uint64_t round0 = server_.sync_->rounds_done();
      server_.backend_->push(id_, keys, n, vals, true, &res, io, todo.data());
      acc.n_local += res.n_local; acc.n_remote += res.n_remote; acc.n_failed += res.n_failed;
      if (res.n_retry == 0) break;
      while (server_.sync_->rounds_done() == round0) {
        std::this_thread::sleep_for(std::chrono::microseconds(50));
        const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        ADAPM_CHECK(el < server_.options().wait_timeout_s, "watchdog: Set is waiting for a sync round that never completes");
      }
    }
    res = acc;
  } else {
    ticket = server_.backend_->push(id_, keys, n, vals, set, &res, io);
  }
  if (ticket == 0) {
    ADAPM_CHECK(res.n_failed == 0, "push failed for " << res.n_failed << " keys (protocol error)");
    num_push_params_local += res.n_local;
    if (res.n_remote == 0) { ++num_push_ops_local; return LOCAL; }
  }
  return new_ts(ticket);
}

int Worker::Pull(const Key* keys, size_t n, void* vals, const IoDesc& io) {
  ADAPM_NVTX("adapm::Pull");
  ++num_pull_ops;
  num_pull_params += n;
  OpResult res;
  uint64_t ticket = server_.backend_->pull(id_, keys, n, vals, false, nullptr, &res, io);
  if (ticket == 0) {
    ADAPM_CHECK(res.n_failed == 0, "pull failed for " << res.n_failed << " keys (protocol error)");
    num_pull_params_local += res.n_local;
    if (res.n_remote == 0) { ++num_pull_ops_local; return LOCAL; }
  }
  return new_ts(ticket);
}

bool Worker::PullIfLocal(Key key, void* vals) {
  ADAPM_CHECK(key >= 0 && key < server_.num_keys(), "[ERROR] Pull key " << key << ", which is outside the configured key range [0," << server_.num_keys() << ")");
  uint8_t ok = 0;
  OpResult res;
  server_.backend_->pull(id_, &key, 1, vals, true, &ok, &res, IoDesc());
  return ok != 0;
}

int Worker::Intent(const Key* keys, size_t n, Clock start, Clock end) {
  ADAPM_NVTX("adapm::Intent");
  if (end == 0) end = start + 1;
  if (server_.num_servers() == 1 || n == 0) return LOCAL;  // single node: nothing to manage
  // Copy only: duplicates are removed by the sync thread (off the worker's critical path).
  auto uniq = server_.sync_->acquire_key_buffer();
  uniq->resize(n);
  Key* dst = uniq->data();
  const Key nk = server_.num_keys();
  Key bad = 0;
  for (size_t i = 0; i < n; ++i) {
    const Key k = keys[i];
    dst[i] = k;
    bad |= (Key)((uint64_t)k >= (uint64_t)nk);
  }
  if (bad) {
    for (size_t i = 0; i < n; ++i)
      ADAPM_CHECK(keys[i] >= 0 && keys[i] < nk, "[ERROR] Intent key " << keys[i] << " is outside the configured key range");
  }
  FutureIntent fi;
  fi.start = start; fi.end = end; fi.worker = id_; fi.keys = uniq;
  server_.sync_->enqueue(std::move(fi));
  return LOCAL;
}

size_t Worker::IntentFast(const Key* keys, size_t n, Clock start, Clock end) {
  if (end == 0) end = start + 1;
  if (server_.num_servers() == 1 || n == 0) return n;
  ADAPM_CHECK(!server_.backend_->is_cuda(), "IntentFast is the host pre-pass; on the cuda backend use ops.IntentPrepass");
  const Ctx& c = server_.backend_->ctx();
  std::vector<Key> rest;
  for (size_t i = 0; i < n; ++i) {
    const Key k = keys[i];
    ADAPM_CHECK(k >= 0 && k < server_.num_keys(), "[ERROR] Intent key " << k << " is outside the configured key range");
    if (!extend_intent_if_local(c, k, id_, end)) rest.push_back(k);
  }
  if (!rest.empty()) Intent(rest.data(), rest.size(), start, end);
  return n - rest.size();
}

Clock Worker::advanceClock() {
  return server_.my_control().worker_clock[id_].fetch_add(1, std::memory_order_acq_rel) + 1;
}
Clock Worker::currentClock() const {
  return server_.my_control().worker_clock[id_].load(std::memory_order_acquire);
}

SampleID Worker::PrepareSample(size_t K, Clock start, Clock end) {
  ADAPM_CHECK(server_.sampling_, "sampling support is not enabled (call enable_sampling_support)");
  if (end == 0) end = start + 1;
  return server_.sampling_->prepare_sample(K, id_, start, end);
}
int Worker::PullSample(SampleID id, Key* keys, size_t n, void* vals) {
  ADAPM_NVTX("adapm::PullSample");
  ADAPM_CHECK(server_.sampling_, "sampling support is not enabled (call enable_sampling_support)");
  return server_.sampling_->pull_sample(id, keys, n, vals, *this);
}
void Worker::FinishSample(SampleID id) {
  if (server_.sampling_) server_.sampling_->finish_sample(id, id_);
}

void Worker::Wait(int ts) {
  ADAPM_NVTX("adapm::Wait");
  if (ts == LOCAL) return;
  uint64_t t;
  {
    std::lock_guard<std::mutex> lk(ts_mu_);
    ADAPM_CHECK(ts >= 0 && (size_t)ts < tickets_.size(), "unknown timestamp " << ts);
    t = tickets_[ts];
  }
  if (t) {
    server_.backend_->wait_ticket(t);
    std::lock_guard<std::mutex> lk(ts_mu_);
    tickets_[ts] = 0;
  }
}
bool Worker::IsFinished(int ts) {
  if (ts == LOCAL) return true;
  std::lock_guard<std::mutex> lk(ts_mu_);
  ADAPM_CHECK(ts >= 0 && (size_t)ts < tickets_.size(), "unknown timestamp " << ts);
  if (tickets_[ts] == 0) return true;
  if (server_.backend_->ticket_done(tickets_[ts])) { tickets_[ts] = 0; return true; }
  return false;
}
void Worker::WaitAll() {
  ADAPM_NVTX("adapm::WaitAll");
  server_.backend_->wait_worker(id_);
  std::lock_guard<std::mutex> lk(ts_mu_);
  for (auto& t : tickets_) t = 0;
}
void Worker::WaitSync() {
  ADAPM_NVTX("adapm::WaitSync");
  if (server_.num_servers() == 1) return;
  server_.backend_->wait_worker(id_);
  server_.sync_->wait_sync();
}
void Worker::Barrier() {
  ADAPM_NVTX("adapm::Barrier");
  server_.worker_barrier();
}
void Worker::BeginSetup() { WaitSync(); Barrier(); }
void Worker::EndSetup() { WaitSync(); WaitSync(); Barrier(); ResetStats(); }
void Worker::ResetStats() {
  num_pull_ops = num_pull_ops_local = num_push_ops = num_push_ops_local = 0;
  num_pull_params = num_pull_params_local = num_push_params = num_push_params_local = 0;
  if (id_ == 0) server_.reset_stats();
}
void Worker::Finalize() {
  if (finalized_) return;
  WaitAll();
  server_.deregister_worker(id_);
  finalized_ = true;
  WaitSync();
  Barrier();
}
int Worker::StaggeredPush(const Key* keys, size_t n, const void* vals, size_t group_size) {
  const char* p = reinterpret_cast<const char*>(vals);
  const size_t vb = server_.backend_->ctx().L.val_bytes;
  for (size_t i = 0; i < n; i += group_size) {
    size_t m = std::min(group_size, n - i);
    Wait(Push(keys + i, m, p));
    p += total_len(keys + i, m) * vb;
  }
  return LOCAL;
}

}  // namespace adapm
