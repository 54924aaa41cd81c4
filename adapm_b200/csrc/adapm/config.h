// System options. Same names and defaults as the reference's Boost.program_options
// tier (coloc_kv_server.h:205-222, sync_manager.h:120-123,805-814, sampling.h:166-170);
// options that only make sense for the ZeroMQ transport (sys.zmq_threads,
// sys.location_caches, sys.channels) are accepted and recorded for CLI parity but have
// no effect: there is no message transport and the directory is fully replicated.
#pragma once
#include <cmath>
#include <limits>
#include <map>
#include <string>
#include <vector>
#include "base.h"
#include "log.h"

namespace adapm {

struct Options {
  // ---- placement / fabric (new)
  std::string backend = "cpu";      // "cpu" | "cuda"
  std::string fabric = "inproc";    // "inproc" (ranks = threads) | "shm" (ranks = processes)
  std::string job = "default";      // rendezvous name (shm segment prefix / inproc registry key)
  std::string dtype = "float32";    // float32 | float64 | int64   (cuda: float32)
  int rank = 0;
  int world = 1;
  int device = -1;                  // cuda device index (default: LOCAL_RANK)
  int workers = 1;                  // logical workers per rank (reference: num_threads)
  double pool_factor = 0;           // slots per class per rank = max(min_pool, factor * keys/world); 0 = auto
  int64_t min_pool = 0;
  int64_t pool_bytes = (int64_t)16 << 30;  // auto pool sizing: memory budget per rank for rows + replica bases
  double wait_timeout_s = 300;      // watchdog on every blocking wait (failure detection)

  // ---- reference "sys.*" options
  int zmq_threads = 3;                       // sys.zmq_threads          (no effect)
  MgmtTechniques techniques = MgmtTechniques::ALL;  // sys.techniques
  bool time_intent_actions = true;           // sys.time_intent_actions
  bool location_caches = true;               // sys.location_caches      (no effect)
  int channels = 4;                          // sys.channels             (no effect)
  std::string trace_keys;                    // sys.trace.keys
  std::string stats_out;                     // sys.stats.out
  bool locality_stats = false;               // sys.stats.locality (reference: compile-time PS_LOCALITY_STATS)
  double sync_max_per_sec = 1000;            // sys.sync.max_per_sec
  int sync_pause_ms = 0;                     // sys.sync.pause
  int sync_min_clocks = -1;                  // sys.sync.min_clocks (new): a new round starts only after the workers advanced
                                             // this many clocks since the previous round started (or after
                                             // sync_min_clocks_wait_ms, or at once for WaitSync / shutdown); 0 = off,
                                             // -1 = auto: 8 for the device-resident round, off for host-sequenced rounds
  int sync_min_clocks_wait_ms = 10;          // sys.sync.min_clocks_wait_ms
  double sync_threshold = 0;                 // sys.sync.threshold (-1 all, 0 non-zero, >0 L2, inf off)
  int sweep_period = 64;                     // rolling sweep: every slot ignores dirty hints/versions once per n rounds (new)
  int idle_period = 4;                       // idle replicas check the owner's version every n-th round (new; 1 = every round)
  float timing_initial_estimate = 10;        // sys.timing.initial_estimate
  bool timing_autotune = true;               // sys.timing.autotune
  float timing_smoothing_factor = 0.1f;      // sys.timing.smoothing_factor
  float timing_buffer_quantile = 0.9999f;    // sys.timing.buffer_quantile

  // ---- reference "sampling.*" options
  std::string sampling_scheme = "local";     // naive | preloc | pool | local
  int64_t sampling_pool_size = 250;
  int64_t sampling_reuse = 1;
  int64_t sampling_batch_size = 10000;
  bool sampling_with_replacement = true;

  static MgmtTechniques parse_techniques(const std::string& s) {
    if (s == "all" || s.empty()) return MgmtTechniques::ALL;
    if (s == "replication_only") return MgmtTechniques::REPLICATION_ONLY;
    if (s == "relocation_only") return MgmtTechniques::RELOCATION_ONLY;
    throw Error("unknown management technique '" + s + "' (all|replication_only|relocation_only)");
  }
  static const char* techniques_name(MgmtTechniques t) {
    switch (t) {
      case MgmtTechniques::ALL: return "all";
      case MgmtTechniques::REPLICATION_ONLY: return "replication_only";
      default: return "relocation_only";
    }
  }

  // Set one option by its reference flag name ("sys.sync.max_per_sec", ...). Returns false if unknown.
  bool set(const std::string& name, const std::string& v) {
    auto b = [&](const std::string& x) { return x == "1" || x == "true" || x == "True" || x == "yes"; };
    if (name == "backend") backend = v;
    else if (name == "fabric") fabric = v;
    else if (name == "job") job = v;
    else if (name == "dtype") dtype = v;
    else if (name == "rank") rank = std::stoi(v);
    else if (name == "world") world = std::stoi(v);
    else if (name == "device") device = std::stoi(v);
    else if (name == "workers") workers = std::stoi(v);
    else if (name == "pool_factor") pool_factor = std::stod(v);
    else if (name == "min_pool") min_pool = std::stoll(v);
    else if (name == "pool_bytes") pool_bytes = std::stoll(v);
    else if (name == "wait_timeout_s") wait_timeout_s = std::stod(v);
    else if (name == "sys.zmq_threads") zmq_threads = std::stoi(v);
    else if (name == "sys.techniques") techniques = parse_techniques(v);
    else if (name == "sys.time_intent_actions") time_intent_actions = b(v);
    else if (name == "sys.location_caches") location_caches = b(v);
    else if (name == "sys.channels") {
      channels = std::stoi(v);
      ADAPM_CHECK(channels > 0 && (channels & (channels - 1)) == 0, "sys.channels must be a power of 2");
    }
    else if (name == "sys.trace.keys") trace_keys = v;
    else if (name == "sys.stats.out") stats_out = v;
    else if (name == "sys.stats.locality") locality_stats = b(v);
    else if (name == "sys.sync.max_per_sec") { sync_max_per_sec = std::stod(v); }
    else if (name == "sys.sync.pause") { sync_pause_ms = std::stoi(v); if (sync_pause_ms > 0) sync_max_per_sec = 0; }
    else if (name == "sys.sync.min_clocks") sync_min_clocks = std::stoi(v);
    else if (name == "sys.sync.min_clocks_wait_ms") sync_min_clocks_wait_ms = std::stoi(v);
    else if (name == "sys.sync.threshold") sync_threshold = (v == "inf") ? std::numeric_limits<double>::infinity() : std::stod(v);
    else if (name == "sys.sync.sweep_period") sweep_period = std::stoi(v);
    else if (name == "sys.sync.idle_period") idle_period = std::stoi(v);
    else if (name == "sys.timing.initial_estimate") timing_initial_estimate = std::stof(v);
    else if (name == "sys.timing.autotune") timing_autotune = b(v);
    else if (name == "sys.timing.smoothing_factor") timing_smoothing_factor = std::stof(v);
    else if (name == "sys.timing.buffer_quantile") timing_buffer_quantile = std::stof(v);
    else if (name == "sampling.scheme") sampling_scheme = v;
    else if (name == "sampling.pool_size") sampling_pool_size = std::stoll(v);
    else if (name == "sampling.reuse") sampling_reuse = std::stoll(v);
    else if (name == "sampling.batch_size") sampling_batch_size = std::stoll(v);
    else if (name == "sampling.with_replacement") sampling_with_replacement = b(v);
    else return false;
    return true;
  }
};

}  // namespace adapm
