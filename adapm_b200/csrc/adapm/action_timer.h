// ActionTimer: decides how far ahead of a worker's clock the system acts on intents.
//
// Behavioural parity with the reference's ActionTimer (sync_manager.h:56-105): keep an
// exponentially smoothed estimate of "clocks per sync round" per worker, never estimate less
// than what the worker did in the last round, and open a window of
// quantile_q(Poisson(2 * estimate)) clocks. The reference uses Boost.Math; there is no Boost
// on the target image, so the Poisson quantile is computed here (exact summation for small
// lambda, Cornish-Fisher-corrected normal quantile for large lambda).
#pragma once
#include <cmath>
#include <vector>
#include "base.h"

namespace adapm {

// inverse standard-normal CDF (Acklam's rational approximation, |err| < 1.2e-9)
inline double normal_quantile(double p) {
  static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                             1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
  static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                             6.680131188771972e+01, -1.328068155288572e+01};
  static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                             -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
  static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                             3.754408661907416e+00};
  const double plow = 0.02425, phigh = 1 - plow;
  if (p <= 0) return -1e300;
  if (p >= 1) return 1e300;
  if (p < plow) {
    double q = std::sqrt(-2 * std::log(p));
    return (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
           ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
  }
  if (p > phigh) {
    double q = std::sqrt(-2 * std::log(1 - p));
    return -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
           ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
  }
  double q = p - 0.5, r = q * q;
  return (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
         (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
}

// smallest k with P(Poisson(lambda) <= k) >= q
inline int64_t poisson_quantile(double lambda, double q) {
  if (lambda <= 0) return 0;
  if (lambda < 400) {
    double p = std::exp(-lambda), cdf = p;
    int64_t k = 0;
    while (cdf < q && k < 100000) { ++k; p *= lambda / (double)k; cdf += p; }
    return k;
  }
  double z = normal_quantile(q);
  double k = lambda + z * std::sqrt(lambda) + (z * z - 1) / 6.0;  // Cornish-Fisher, skewness 1/sqrt(lambda)
  return (int64_t)std::ceil(k);
}

class ActionTimer {
 public:
  ActionTimer(int workers, float initial, bool autotune, float alpha, float quantile)
      : est_(workers, initial), last_(workers, 0), autotune_(autotune), alpha_(alpha), q_(quantile) {}

  std::vector<Clock> estimate_windows_and_tune(const std::vector<Clock>& clocks, uint64_t round) {
    std::vector<Clock> win(est_.size());
    for (size_t w = 0; w < est_.size(); ++w) {
      float e = est_[w];
      if (round > 0) {
        float ticks = (float)(clocks[w] - last_[w]);
        if (autotune_ && clocks[w] != WORKER_FINISHED && ticks != 0) {
          est_[w] = (1 - alpha_) * est_[w] + alpha_ * ticks;
          e = std::max(ticks, est_[w]);
        }
      }
      win[w] = (Clock)poisson_quantile(2.0 * e, q_);
    }
    last_ = clocks;
    return win;
  }
  float avg_estimate() const {
    float s = 0;
    for (float e : est_) s += e;
    return est_.empty() ? 0 : s / est_.size();
  }

 private:
  std::vector<float> est_;
  std::vector<Clock> last_;
  bool autotune_;
  float alpha_, q_;
};

}  // namespace adapm
