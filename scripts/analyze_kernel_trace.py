#!/usr/bin/env python
"""Reads a kernel timeline written with ADAPM_SYNC_TRACE=1 (bench.py --profile -> gpurun_out/kernel_trace.rank<r>.tsv)
and reports how the sync-round kernels and the training steps share the GPU:

    python scripts/analyze_kernel_trace.py gpurun_out/kernel_trace.rank0.tsv

* duration statistics per record type (register / phaseA / phaseB / phaseC on the high-priority sync stream, steps
  on the training stream);
* step duration split by how much sync-kernel time overlapped the step;
* idle gaps between consecutive steps.
"""
from __future__ import annotations

import statistics
import sys
from collections import defaultdict


def pct(v, p):
    v = sorted(v)
    return v[min(len(v) - 1, int(p * len(v)))]


def main(path):
    recs = []
    for i, line in enumerate(open(path)):
        if i == 0:
            continue
        n, a, b = line.rstrip("\n").split("\t")
        recs.append((n, float(a), float(b)))
    begins = [a for n, a, b in recs if n == "step_begin"]
    ends = [a for n, a, b in recs if n == "step_end"]
    steps = list(zip(begins, ends))
    if not steps:
        print("no step marks in the trace")
        return 1
    t0, t1 = steps[0][0], steps[-1][1]
    sync = [(n, a, b) for n, a, b in recs if not n.startswith("step_") and b > t0 and a < t1]
    by = defaultdict(list)
    for n, a, b in sync:
        by[n].append(b - a)
    print(f"window {t1 - t0:.1f} ms, {len(steps)} steps, {len(sync)} sync records")
    for n, v in sorted(by.items()):
        print(f"  {n:9s} n={len(v):5d} mean {statistics.mean(v):.3f} ms  p50 {pct(v, .5):.3f}  p90 {pct(v, .9):.3f}  "
              f"max {max(v):.3f}  total {sum(v):.1f} ms ({100 * sum(v) / (t1 - t0):.1f}% of the window)")
    dur = [e - b for b, e in steps]
    print(f"  step      n={len(dur):5d} mean {statistics.mean(dur):.3f} ms  p50 {pct(dur, .5):.3f}  p90 {pct(dur, .9):.3f}  "
          f"max {max(dur):.3f}")
    gaps = [steps[i + 1][0] - steps[i][1] for i in range(len(steps) - 1)]
    print(f"  gap between steps: mean {statistics.mean(gaps):.3f} ms  p90 {pct(gaps, .9):.3f}  max {max(gaps):.3f}  "
          f"total {sum(gaps):.1f} ms")
    # overlap of each step with sync-kernel intervals
    buckets = defaultdict(list)
    for b, e in steps:
        ov = 0.0
        kinds = set()
        for n, a, c in sync:
            lo, hi = max(a, b), min(c, e)
            if hi > lo:
                ov += hi - lo
                kinds.add(n)
        frac = ov / (e - b)
        key = "no sync overlap" if ov == 0 else ("<25% overlapped" if frac < 0.25 else ("25-75%" if frac < 0.75 else ">75%"))
        buckets[key].append((e - b, ov, tuple(sorted(kinds))))
    for k in ("no sync overlap", "<25% overlapped", "25-75%", ">75%"):
        v = buckets.get(k)
        if v:
            d = [x[0] for x in v]
            print(f"  steps with {k:16s}: n={len(v):4d} mean step {statistics.mean(d):.3f} ms (p90 {pct(d, .9):.3f}), "
                  f"mean overlapped sync time {statistics.mean(x[1] for x in v):.3f} ms")
    # slowest steps and what overlapped them
    worst = sorted(((e - b, b, e) for b, e in steps), reverse=True)[:8]
    print("  slowest steps:")
    for d, b, e in worst:
        ov = [(n, round(max(a, b) - b, 3), round(min(c, e) - b, 3)) for n, a, c in sync if min(c, e) > max(a, b)]
        print(f"    {d:.3f} ms at t={b - t0:.1f}: {ov[:8]}")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1]))
