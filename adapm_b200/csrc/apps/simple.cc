// Native C++ application on the core API (no Python): the reference's apps/simple.cc (:36-134) plus a
// stress mode that runs the reference's dynamic-allocation contract at full scale
// (tests/test_dynamic_allocation.cc: 100 000 fully asynchronous Push+Pull per worker with random intents).
//
//   adapm_simple [-s servers] [-t threads] [-k keys] [-i iterations] [-v values_per_key] [--stress runs]
//
// Ranks are threads of this process (inproc fabric) so that the binary is self-contained; the same code
// runs one-process-per-GPU when fabric=shm and RANK/WORLD_SIZE are set.
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <random>
#include <thread>
#include <vector>

#include "adapm/node.h"

using namespace adapm;

struct Args { int servers = 2, threads = 2, keys = 10, iters = 4, vpk = 2; long stress = 0; std::string backend = "cpu"; };

static void run_worker(Server& server, int cid, const Args& a, long* stress_result) {
  Worker kv(cid, server);
  const int wid = server.my_rank() * a.threads + cid;
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
    else if (!strcmp(argv[i], "--backend")) a.backend = next();
  }
  long result[2] = {0, 0};
  std::vector<std::thread> nodes;
  for (int r = 0; r < a.servers; ++r) {
    nodes.emplace_back([&, r] {
      Options opt;
      opt.backend = a.backend; opt.fabric = "inproc"; opt.job = "adapm_simple"; opt.rank = r; opt.world = a.servers;
      opt.workers = a.threads; opt.dtype = a.backend == "cuda" ? "float32" : "float64";
      ValueSpec spec;
      spec.num_keys = a.keys; spec.uniform_len = a.vpk;
      Server server(opt, spec);
      std::vector<std::thread> ws;
      for (int c = 0; c < a.threads; ++c) ws.emplace_back(run_worker, std::ref(server), c, std::cref(a), result);
      for (auto& t : ws) t.join();
      if (r == 0) ALOG(server.stats_string());
      server.shutdown();
    });
  }
  for (auto& t : nodes) t.join();
  if (a.stress) {
    const long expect = (long)a.servers * a.threads * a.stress;
    const bool ok = result[0] == expect && result[1] == 2 * expect;
    std::cout << "Result: [" << result[0] << ", " << result[1] << "]  Correct: [" << expect << ", " << 2 * expect << "]\n"
              << "Dynamic Allocation: " << (ok ? "PASSED" : "FAILED") << std::endl;
    return ok ? 0 : 1;
  }
  return 0;
}
