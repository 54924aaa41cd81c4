// Native C++ application on the core API (no Python): the reference's apps/simple.cc (:36-134) plus a
// stress mode that runs the reference's dynamic-allocation contract at full scale
// (tests/test_dynamic_allocation.cc: 100 000 fully asynchronous Push+Pull per worker with random intents).
//
//   adapm_simple [-s servers] [-t threads] [-k keys] [-i iterations] [-v values_per_key] [--stress runs]
//   adapm_simple --fuzz ops [--seed n] [--techniques all|replication_only|relocation_only] [-s .. -t .. -k ..]
//       random programs of Intent / Push / Pull / advanceClock / WaitSync on every worker; checks read-your-writes
//       during the run and the exact per-key sums at the end (the C++ twin of tests/test_protocol_property.py,
//       meant to be run under ThreadSanitizer / AddressSanitizer: scripts/sanitize.sh)
//
// Ranks are threads of this process (inproc fabric) so that the binary is self-contained; the same code
// runs one-process-per-GPU when fabric=shm and RANK/WORLD_SIZE are set.
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <random>
#include <thread>
#include <vector>

#include "adapm/node.h"

using namespace adapm;

struct Args {
  int servers = 2, threads = 2, keys = 10, iters = 4, vpk = 2;
  long stress = 0, fuzz = 0;
  unsigned seed = 1;
  std::string backend = "cpu", techniques = "all";
};

// ---- fuzz: shared bookkeeping of what was pushed, and the verdict
static std::vector<std::atomic<long>> g_pushed;
static std::atomic<int> g_fuzz_errors{0};
static std::atomic<long> g_fuzz_torn{0};   // rows whose elements were read at different points of concurrent pushes

static void run_fuzz(Server& server, Worker& kv, int wid, const Args& a) {
  std::mt19937 rng(a.seed * 7919u + (unsigned)wid * 104729u + 13u);
  std::vector<long> mine((size_t)a.keys, 0);
  std::vector<Key> ks;
  std::vector<double> vals;
  auto random_keys = [&] {
    ks.clear();
    const int n = 1 + (int)(rng() % 6);
    for (int i = 0; i < n; ++i) {
      Key k = (Key)(rng() % (unsigned)a.keys);
      bool dup = false;
      for (Key q : ks) dup = dup || q == k;
      if (!dup) ks.push_back(k);
    }
  };
  kv.Barrier();
  for (long op = 0; op < a.fuzz; ++op) {
    const unsigned what = rng() % 16;
    random_keys();
    if (what < 3) {
      const Clock c = kv.currentClock() + (Clock)(rng() % 4);
      if (rng() % 2) kv.Intent(ks.data(), ks.size(), c, c + 1 + (Clock)(rng() % 4));
      else kv.IntentFast(ks.data(), ks.size(), c, c + 1 + (Clock)(rng() % 4));   // pre-pass on this thread
    } else if (what < 9) {
      vals.assign(ks.size() * (size_t)a.vpk, 1.0);
      const int ts = kv.Push(ks.data(), ks.size(), vals.data());
      if (rng() % 2) kv.Wait(ts);
      for (Key k : ks) { ++mine[(size_t)k]; g_pushed[(size_t)k].fetch_add(1, std::memory_order_relaxed); }
    } else if (what < 13) {
      kv.WaitAll();     // read-your-writes is promised for completed pushes
      vals.assign(ks.size() * (size_t)a.vpk, -1.0);
      kv.Wait(kv.Pull(ks.data(), ks.size(), vals.data()));
      for (size_t i = 0; i < ks.size(); ++i) {
        // every element of the row must include this worker's completed pushes. Elements of one row may differ by the
        // pushes of OTHER workers that are in flight: rows are updated with per-element atomic adds, a Pull is not a
        // snapshot - counted, not an error.
        bool torn = false;
        for (int z = 0; z < a.vpk; ++z) {
          const double v = vals[i * (size_t)a.vpk + (size_t)z];
          torn = torn || v != vals[i * (size_t)a.vpk];
          if (v < (double)mine[(size_t)ks[i]]) {
            ALOG("fuzz: worker " << wid << " op " << op << ": key " << ks[i] << " element " << z << " read " << v
                                 << " but this worker alone pushed " << mine[(size_t)ks[i]]);
            g_fuzz_errors.fetch_add(1);
          }
        }
        if (torn) g_fuzz_torn.fetch_add(1, std::memory_order_relaxed);
      }
    } else if (what < 15) {
      kv.advanceClock();
    } else {
      kv.WaitSync();
    }
  }
  kv.WaitAll(); kv.WaitSync(); kv.Barrier(); kv.WaitSync(); kv.Barrier();
  // every worker checks the final state against the global sums
  ks.resize((size_t)a.keys);
  for (int k = 0; k < a.keys; ++k) ks[(size_t)k] = k;
  vals.assign((size_t)a.keys * (size_t)a.vpk, -1.0);
  kv.Wait(kv.Pull(ks.data(), ks.size(), vals.data()));
  for (int k = 0; k < a.keys; ++k) {
    const long want = g_pushed[(size_t)k].load();
    for (int z = 0; z < a.vpk; ++z) {
      if (vals[(size_t)k * (size_t)a.vpk + (size_t)z] != (double)want) {
        ALOG("fuzz: worker " << wid << ": final value of key " << k << " is " << vals[(size_t)k * (size_t)a.vpk + (size_t)z]
                             << ", expected " << want);
        g_fuzz_errors.fetch_add(1);
        break;
      }
    }
  }
  kv.Barrier();
  kv.Finalize();
  (void)server;
}

static void run_worker(Server& server, int cid, const Args& a, long* stress_result) {
  Worker kv(cid, server);
  const int wid = server.my_rank() * a.threads + cid;
  if (a.fuzz > 0) { run_fuzz(server, kv, wid, a); return; }
  if (a.stress == 0) {
    std::vector<double> vals(a.vpk), push(a.vpk);
    for (int x = 0; x < a.iters; ++x) {
      Key key = x % a.keys;
      kv.Intent(key, kv.currentClock());
      for (int z = 0; z < a.vpk; ++z) push[z] = z + 1;
      kv.Wait(kv.Push(&key, 1, push.data()));
      kv.Wait(kv.Pull(&key, 1, vals.data()));
      ALOG("Worker " << wid << " iteration " << x << ": key " << key << " = [" << vals[0] << (a.vpk > 1 ? ", ..." : "") << "]");
      kv.advanceClock();
    }
    kv.Barrier();
    kv.Finalize();
    return;
  }
  // ---- stress: no update may be lost or duplicated while key 9 relocates / replicates
  kv.Barrier();
  std::mt19937 rng(wid * 31 + 7);
  Key key = 9 % a.keys;
  std::vector<double> v1 = {1, 2}, v2(2);
  std::vector<int> ts;
  ts.reserve(2 * a.stress);
  for (long run = 0; run < a.stress; ++run) {
    if (rng() % 50 == 0) kv.Intent(key, kv.currentClock() + 10, kv.currentClock() + 40);
    ts.push_back(kv.Push(&key, 1, v1.data()));
    ts.push_back(kv.Pull(&key, 1, v2.data()));
    kv.advanceClock();
  }
  for (int t : ts) kv.Wait(t);
  kv.WaitSync(); kv.Barrier(); kv.WaitSync();
  if (wid == 0) {
    kv.Wait(kv.Pull(&key, 1, v2.data()));
    stress_result[0] = (long)v2[0];
    stress_result[1] = (long)v2[1];
  }
  kv.WaitSync();
  kv.Finalize();
}

int main(int argc, char** argv) {
  Args a;
  for (int i = 1; i < argc; ++i) {
    auto next = [&] { return i + 1 < argc ? argv[++i] : (char*)"0"; };
    if (!strcmp(argv[i], "-s")) a.servers = atoi(next());
    else if (!strcmp(argv[i], "-t")) a.threads = atoi(next());
    else if (!strcmp(argv[i], "-k")) a.keys = atoi(next());
    else if (!strcmp(argv[i], "-i")) a.iters = atoi(next());
    else if (!strcmp(argv[i], "-v")) a.vpk = atoi(next());
    else if (!strcmp(argv[i], "--stress")) { a.stress = atol(next()); a.vpk = 2; if (a.keys < 10) a.keys = 20; }
    else if (!strcmp(argv[i], "--fuzz")) { a.fuzz = atol(next()); if (a.keys < 16) a.keys = 24; }
    else if (!strcmp(argv[i], "--seed")) a.seed = (unsigned)atol(next());
    else if (!strcmp(argv[i], "--techniques")) a.techniques = next();
    else if (!strcmp(argv[i], "--backend")) a.backend = next();
  }
  long result[2] = {0, 0};
  g_pushed = std::vector<std::atomic<long>>((size_t)a.keys);
  for (auto& x : g_pushed) x.store(0);
  std::vector<std::thread> nodes;
  for (int r = 0; r < a.servers; ++r) {
    nodes.emplace_back([&, r] {
      Options opt;
      opt.backend = a.backend; opt.fabric = "inproc"; opt.job = "adapm_simple"; opt.rank = r; opt.world = a.servers;
      opt.workers = a.threads; opt.dtype = a.backend == "cuda" ? "float32" : "float64";
      opt.set("sys.techniques", a.techniques);
      opt.wait_timeout_s = 120;
      ValueSpec spec;
      spec.num_keys = a.keys; spec.uniform_len = a.vpk;
      Server server(opt, spec);
      std::vector<std::thread> ws;
      for (int c = 0; c < a.threads; ++c) ws.emplace_back(run_worker, std::ref(server), c, std::cref(a), result);
      for (auto& t : ws) t.join();
      if (r == 0) ALOG(server.stats_string());
      if (a.fuzz) {   // the protocol's own consistency checks count as fuzz errors too
        const uint64_t pe = server.counters()["protocol_errors"];
        if (pe) { ALOG("fuzz: rank " << r << " counted " << pe << " protocol errors"); g_fuzz_errors.fetch_add((int)pe); }
      }
      server.shutdown();
    });
  }
  for (auto& t : nodes) t.join();
  if (a.fuzz) {
    long total = 0;
    for (auto& x : g_pushed) total += x.load();
    std::cout << "Fuzz: " << a.servers * a.threads << " workers x " << a.fuzz << " ops, " << total << " key updates, "
              << g_fuzz_torn.load() << " non-snapshot row reads, " << g_fuzz_errors.load()
              << " errors: " << (g_fuzz_errors.load() == 0 ? "PASSED" : "FAILED") << std::endl;
    return g_fuzz_errors.load() == 0 ? 0 : 1;
  }
  if (a.stress) {
    const long expect = (long)a.servers * a.threads * a.stress;
    const bool ok = result[0] == expect && result[1] == 2 * expect;
    std::cout << "Result: [" << result[0] << ", " << result[1] << "]  Correct: [" << expect << ", " << 2 * expect << "]\n"
              << "Dynamic Allocation: " << (ok ? "PASSED" : "FAILED") << std::endl;
    return ok ? 0 : 1;
  }
  return 0;
}
