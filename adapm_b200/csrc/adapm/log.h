// Logging, checks and stopwatch (parity: include/dmlc/logging.h ALOG/CHECK,
// include/utils.h:85-125 Stopwatch). Written fresh; no dmlc shim.
#pragma once
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <mutex>

namespace adapm {

inline int& verbosity() { static int v = [] { const char* e = getenv("ADAPM_VERBOSE"); if (!e) e = getenv("PS_VERBOSE"); return e ? atoi(e) : 0; }(); return v; }

inline void log_line(const std::string& s) {
  static std::mutex mu;
  auto now = std::chrono::system_clock::now();
  auto t = std::chrono::system_clock::to_time_t(now);
  auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(now.time_since_epoch()).count() % 1000;
  char buf[32];
  struct tm tmv;
  localtime_r(&t, &tmv);
  strftime(buf, sizeof(buf), "%H:%M:%S", &tmv);
  std::lock_guard<std::mutex> lk(mu);
  fprintf(stdout, "[%s.%03d] %s\n", buf, (int)ms, s.c_str());
  fflush(stdout);
}

#define ALOG(x)                              \
  do {                                       \
    std::ostringstream _adapm_os;            \
    _adapm_os << x;                          \
    ::adapm::log_line(_adapm_os.str());      \
  } while (0)
#define VLOG(level, x) \
  do { if (::adapm::verbosity() >= (level)) ALOG(x); } while (0)

struct Error : public std::runtime_error { using std::runtime_error::runtime_error; };

#define ADAPM_CHECK(cond, msg)                                                         \
  do {                                                                                 \
    if (!(cond)) {                                                                     \
      std::ostringstream _adapm_os;                                                    \
      _adapm_os << "[adapm] check failed: " #cond " at " << __FILE__ << ":" << __LINE__ \
                << ": " << msg;                                                        \
      throw ::adapm::Error(_adapm_os.str());                                           \
    }                                                                                  \
  } while (0)

// Accumulating stopwatch. One thread drives it (start / resume / stop); any thread may read elapsed_*() concurrently
// (statistics are printed while the sync thread is running), hence the atomics.
class Stopwatch {
 public:
  void start() { total_ns_.store(0, std::memory_order_relaxed); resume(); }
  void resume() { t0_ns_.store(now_ns(), std::memory_order_relaxed); running_.store(true, std::memory_order_release); }
  void stop() {
    if (running_.load(std::memory_order_relaxed)) {
      total_ns_.fetch_add(now_ns() - t0_ns_.load(std::memory_order_relaxed), std::memory_order_relaxed);
      running_.store(false, std::memory_order_release);
    }
  }
  double elapsed_s() const {
    int64_t t = total_ns_.load(std::memory_order_relaxed);
    if (running_.load(std::memory_order_acquire)) t += now_ns() - t0_ns_.load(std::memory_order_relaxed);
    return (double)t * 1e-9;
  }
  double elapsed_ms() const { return elapsed_s() * 1e3; }
 private:
  static int64_t now_ns() {
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
  }
  std::atomic<int64_t> t0_ns_{0}, total_ns_{0};
  std::atomic<bool> running_{false};
};
inline std::ostream& operator<<(std::ostream& os, const Stopwatch& sw) { return os << sw.elapsed_s() << "s"; }

}  // namespace adapm
