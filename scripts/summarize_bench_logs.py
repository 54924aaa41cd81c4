#!/usr/bin/env python
"""One short line per bench log in a directory (value, ms/step, e2e, locality, churn per step)."""
import glob, json, os, sys
d = sys.argv[1]
for f in sorted(glob.glob(os.path.join(d, "*.log"))):
    name = os.path.basename(f)
    if name.startswith("pytest"):
        lines = open(f, errors="replace").read().strip().splitlines()
        print(name, "|", " / ".join(lines[-2:])[:300])
        continue
    js = [l for l in open(f, errors="replace") if l.startswith("{")]
    if not js:
        tail = open(f, errors="replace").read().strip().splitlines()[-3:]
        print(name, "| NO JSON:", " / ".join(tail)[:400])
        continue
    j = json.loads(js[-1])
    if "value" not in j:
        print(name, "|", js[-1].strip()[:600])
        continue
    k = j.get("steps", 1)
    pm = j.get("pm") or {}
    nst = 2 * k + 3
    loc = j.get("locality") or {}
    print(name, "| value %.3fG ms %.3f | e2e %.3fG | rem/slow per step %s/%s | reloc %s setup %s drop %s refr %s "
          "delta %s rounds %s err %s | host_ms %s" % (
        j["value"] / 1e9, j["ms_per_step"], (j.get("e2e") or {}).get("value", 0) / 1e9,
        loc.get("rows_remote", 0) // nst, loc.get("rows_slow_path", 0) // nst,
        pm.get("relocations", 0) // nst, pm.get("replica_setups", 0) // nst, pm.get("replica_drops", 0) // nst,
        pm.get("refreshes", 0) // nst, pm.get("deltas_shipped", 0) // nst, pm.get("sync_rounds"), pm.get("protocol_errors"),
        j.get("host_loop_ms_per_step")))
    if j.get("profile"):
        p = j["profile"]
        print("    profile: sgns_ms %.3f max %.3f | with intent %s | without %s" % (
            p.get("sgns_ms", 0), p.get("sgns_ms_max", 0), p.get("steps_with_intent"), p.get("steps_without_intent")))
    if j.get("sync_report"):
        print("    " + j["sync_report"][:400])
