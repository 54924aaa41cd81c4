"""Property-based test of the relocation / replication protocol (hypothesis): random programs of Intent / Push / Pull /
advanceClock on every rank, executed concurrently on the CPU backend, must never lose or duplicate an update, must
give each worker read-your-writes, and must converge to the exact sum after the WaitSync idiom - for every
management technique and for random replica-maintenance pacing."""
import os

import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from harness import run_cluster

NUM_KEYS, VPK = 24, 2

key_sets = st.lists(st.integers(0, NUM_KEYS - 1), min_size=1, max_size=6, unique=True)
op = st.one_of(
    st.tuples(st.just("intent"), key_sets, st.integers(0, 3), st.integers(1, 4)),   # keys, start offset, duration
    st.tuples(st.just("intent_fast"), key_sets, st.integers(0, 3), st.integers(1, 4)),   # same, pre-pass on this thread
    st.tuples(st.just("push"), key_sets),
    st.tuples(st.just("pull"), key_sets),
    st.tuples(st.just("clock")),
    st.tuples(st.just("sync")),
)
program = st.lists(st.lists(op, min_size=1, max_size=6), min_size=1, max_size=5)    # rounds of ops


def _run(kv, server, wid, programs=None):
    prog = programs[wid]
    rounds = max(len(p) for p in programs)
    mine = torch.zeros(NUM_KEYS, dtype=torch.int64)          # what this worker pushed so far, per key
    errs = []
    prepass = None
    for r in range(rounds):
        for o in (prog[r] if r < len(prog) else []):
            if o[0] == "intent":
                kv.intent(torch.tensor(o[1]), kv.current_clock() + o[2], kv.current_clock() + o[2] + o[3])
            elif o[0] == "intent_fast":
                c0 = kv.current_clock()
                if server.device.type == "cuda":     # the pre-pass runs on the device there (ops_intent.cu)
                    if prepass is None:
                        from adapm_b200.ops import IntentPrepass

                        prepass = IntentPrepass(server, kv, max_keys=8)
                    prepass.submit(torch.tensor(o[1]), c0 + o[2], c0 + o[2] + o[3])
                    prepass.harvest(block=True)
                else:
                    kv.intent_fast(torch.tensor(o[1]), c0 + o[2], c0 + o[2] + o[3])
            elif o[0] == "push":
                k = torch.tensor(o[1])
                kv.wait(kv.push(k, torch.ones(len(o[1]) * VPK, dtype=torch.int64)))
                mine[k] += 1
            elif o[0] == "pull":
                k = torch.tensor(o[1])
                v = torch.zeros(len(o[1]) * VPK, dtype=torch.int64)
                kv.wait(kv.pull(k, v))
                v = v.view(-1, VPK)
                # (a Pull is not a snapshot: the elements of a row may include different in-flight pushes of OTHER workers,
                #  DESIGN.md section 4 - but every element includes this worker's own completed pushes)
                if bool((v < mine[k].view(-1, 1)).any()):
                    errs.append(f"w{wid} round {r}: read-your-writes violated: {v.tolist()} < {mine[k].tolist()}")
            elif o[0] == "clock":
                kv.advance_clock()
            elif o[0] == "sync":
                kv.wait_sync()
        kv.barrier()
    kv.waitall(); kv.wait_sync(); kv.barrier(); kv.wait_sync(); kv.barrier()
    out = torch.zeros(NUM_KEYS * VPK, dtype=torch.int64)
    kv.wait(kv.pull(torch.arange(NUM_KEYS), out))
    kv.barrier()
    kv.finalize()
    return errs, out.view(-1, VPK)[:, 0].tolist(), mine.tolist()


@settings(max_examples=int(os.environ.get("ADAPM_HYP_EXAMPLES", "20")), deadline=None, suppress_health_check=list(HealthCheck))
@given(programs=st.lists(program, min_size=3, max_size=3),
       technique=st.sampled_from(["all", "replication_only", "relocation_only"]),
       idle_period=st.integers(1, 5), sweep_period=st.integers(0, 4),
       threshold=st.sampled_from(["-1", "0", "0.5", "3"]), min_clocks=st.sampled_from([0, 0, 2, 8]))
def test_random_programs_are_exact(programs, technique, idle_period, sweep_period, threshold, min_clocks):
    import functools

    res = run_cluster(functools.partial(_run, programs=programs), world=3, workers=1, mode="threads", value_lengths=VPK,
                      num_keys=NUM_KEYS, dtype="int64",
                      options={"sys.techniques": technique, "sys.sync.idle_period": idle_period,
                               "sys.sync.sweep_period": sweep_period, "sys.sync.threshold": threshold,
                               "sys.sync.min_clocks": min_clocks, "sys.sync.min_clocks_wait_ms": 3})
    total = [0] * NUM_KEYS
    for r in res.values():
        errs, final, mine = r[0]
        assert not errs, errs
        total = [a + b for a, b in zip(total, mine)]
    for r in res.values():
        assert r[0][1] == total, (r[0][1], total)           # every rank sees the exact sum of all pushes
        assert r["counters"]["protocol_errors"] == 0


@settings(max_examples=int(os.environ.get("ADAPM_HYP_EXAMPLES", "10")), deadline=None, suppress_health_check=list(HealthCheck))
@given(programs=st.lists(program, min_size=4, max_size=4), technique=st.sampled_from(["all", "replication_only"]))
def test_random_programs_two_workers_per_rank(programs, technique):
    """Two ranks x two worker threads: intents of co-located workers share slots (per-worker intent ends), pushes of
    co-located workers hit the same replica."""
    import functools

    res = run_cluster(functools.partial(_run, programs=programs), world=2, workers=2, mode="threads", value_lengths=VPK,
                      num_keys=NUM_KEYS, dtype="int64", options={"sys.techniques": technique})
    total = [0] * NUM_KEYS
    for r in res.values():
        for cid in (0, 1):
            errs, final, mine = r[cid]
            assert not errs, errs
            total = [a + b for a, b in zip(total, mine)]
    for r in res.values():
        for cid in (0, 1):
            assert r[cid][1] == total, (r[cid][1], total)
        assert r["counters"]["protocol_errors"] == 0


def _run_var(kv, server, wid, programs=None, lens=None):
    """Like _run, on a store whose keys have different value lengths (size classes + per-key length table)."""
    prog = programs[wid]
    rounds = max(len(p) for p in programs)
    lens_t = torch.tensor(lens, dtype=torch.int64)
    mine = torch.zeros(NUM_KEYS, dtype=torch.int64)
    errs = []
    for r in range(rounds):
        for o in (prog[r] if r < len(prog) else []):
            if o[0] in ("intent", "intent_fast"):
                c0 = kv.current_clock()
                fn = kv.intent_fast if (o[0] == "intent_fast" and server.device.type != "cuda") else kv.intent
                fn(torch.tensor(o[1]), c0 + o[2], c0 + o[2] + o[3])
            elif o[0] == "push":
                k = torch.tensor(sorted(o[1]))
                kv.wait(kv.push(k, torch.ones(int(lens_t[k].sum()), dtype=torch.int64)))
                mine[k] += 1
            elif o[0] == "pull":
                k = torch.tensor(sorted(o[1]))
                v = torch.zeros(int(lens_t[k].sum()), dtype=torch.int64)
                kv.wait(kv.pull(k, v))
                off = 0
                for key in k.tolist():
                    seg = v[off:off + lens[key]]
                    off += lens[key]
                    if bool((seg < mine[key]).any()):
                        errs.append(f"w{wid} round {r}: read-your-writes violated for key {key}: {seg.tolist()} < {int(mine[key])}")
            elif o[0] == "clock":
                kv.advance_clock()
            elif o[0] == "sync":
                kv.wait_sync()
        kv.barrier()
    kv.waitall(); kv.wait_sync(); kv.barrier(); kv.wait_sync(); kv.barrier()
    out = torch.zeros(int(lens_t.sum()), dtype=torch.int64)
    kv.wait(kv.pull(torch.arange(NUM_KEYS), out))
    kv.barrier()
    kv.finalize()
    final, off = [], 0
    for key in range(NUM_KEYS):
        seg = out[off:off + lens[key]]
        off += lens[key]
        if not bool((seg == seg[0]).all()):
            errs.append(f"key {key}: elements of one row differ: {seg.tolist()}")
        final.append(int(seg[0]))
    return errs, final, mine.tolist()


@settings(max_examples=int(os.environ.get("ADAPM_HYP_EXAMPLES", "12")), deadline=None, suppress_health_check=list(HealthCheck))
@given(programs=st.lists(program, min_size=3, max_size=3),
       lens=st.lists(st.sampled_from([1, 2, 3, 5, 8, 13, 40]), min_size=NUM_KEYS, max_size=NUM_KEYS),
       technique=st.sampled_from(["all", "replication_only", "relocation_only"]))
def test_random_programs_with_mixed_value_lengths(programs, lens, technique):
    """The same property on keys of different lengths: rows of up to 7 distinct lengths relocate / replicate between
    per-class pools (reference coloc_kv_server_handle.h:162-170: any value_lengths vector)."""
    import functools

    res = run_cluster(functools.partial(_run_var, programs=programs, lens=lens), world=3, workers=1, mode="threads",
                      value_lengths=torch.tensor(lens, dtype=torch.int64), num_keys=NUM_KEYS, dtype="int64",
                      options={"sys.techniques": technique})
    total = [0] * NUM_KEYS
    for r in res.values():
        errs, final, mine = r[0]
        assert not errs, errs
        total = [a + b for a, b in zip(total, mine)]
    for r in res.values():
        assert r[0][1] == total, (r[0][1], total)
        assert r["counters"]["protocol_errors"] == 0
