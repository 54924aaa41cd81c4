"""Many-key operations contract (assertions of tests/test_many_key_operations.cc:51-349, --quick):
 (1) exact reads under concurrent random intents,
 (2) monotonic reads + locality: after Intent + WaitSync every pull is local and >= expected,
 (3) eventual consistency: pushes followed by their negation leave every key unchanged after the
     propagation idiom  WaitAll; WaitSync; Barrier; WaitSync; Barrier.
Run for sys.techniques = all | replication_only | relocation_only (run_tests.sh:25-28).
"""
import random

import pytest
import torch

from harness import run_cluster

NUM_KEYS = 100
VPK = 10
RUNS = 60
MAX_KEYS = 20
AHEAD = 20
MAX_CONC = 10


def _random_keys(rng):
    k = rng.randrange(1, MAX_KEYS)
    return torch.tensor(rng.sample(range(NUM_KEYS), k), dtype=torch.int64)


def _worker(kv, server, wid):
    errs = []
    kv.barrier()
    dt = server.dtype
    g = torch.Generator().manual_seed(7)
    init_vals = torch.randint(-5000000, 5000000, (NUM_KEYS, VPK), generator=g, dtype=torch.int64).to(dt)
    all_keys = torch.arange(NUM_KEYS)
    if wid == 0:
        kv.wait(kv.push(all_keys, init_vals.clone().view(-1)))
    kv.wait_sync()
    kv.barrier()
    rng = random.Random(wid ^ 13)

    # ---------------- (1) pulls and localizes: exact reads
    loc = [_random_keys(rng) for _ in range(RUNS)]
    pk = [_random_keys(rng) for _ in range(RUNS)]
    pv = [None] * RUNS
    pts = [None] * RUNS
    fut = 0
    for i in range(RUNS):
        while fut <= i + AHEAD and fut < RUNS:
            kv.intent(loc[fut], kv.current_clock() + fut - i)
            fut += 1
        pv[i] = torch.full((pk[i].numel() * VPK,), 12, dtype=dt)
        pts[i] = kv.pull(pk[i], pv[i], True)
        if i > MAX_CONC:
            kv.wait(pts[i - MAX_CONC])
        kv.advance_clock()
    wrong = 0
    for i in range(RUNS):
        kv.wait(pts[i])
        if not torch.equal(pv[i].view(-1, VPK), init_vals[pk[i]]):
            wrong += 1
    if wrong:
        errs.append(f"w{wid}: pulls-and-localizes: {wrong} of {RUNS} pulls returned wrong values")
    kv.waitall()
    kv.barrier()

    # ---------------- (2) monotonic pushes + locality
    loc = [_random_keys(rng) for _ in range(RUNS)]
    pk = [_random_keys(rng) for _ in range(RUNS)]
    sk = [_random_keys(rng) for _ in range(RUNS)]
    sv = [None] * RUNS
    fut = 0
    nonlocal_pulls = 0
    for i in range(RUNS):
        while fut <= i + AHEAD and fut < RUNS:
            c = kv.current_clock() + fut - i
            kv.intent(pk[fut], c)
            kv.intent(sk[fut], c)
            kv.intent(loc[fut], c)
            fut += 1
        kv.wait_sync()
        pv[i] = torch.full((pk[i].numel() * VPK,), 12, dtype=dt)
        pts[i] = kv.pull(pk[i], pv[i], True)
        sv[i] = torch.randint(1, 1000, (sk[i].numel() * VPK,), generator=g, dtype=torch.int64).to(dt)
        kv.wait(kv.push(sk[i], sv[i], True))
        if i > MAX_CONC:
            kv.wait(pts[i - MAX_CONC])
        kv.advance_clock()
    kv.wait_sync()
    kv.barrier()
    expected = init_vals.clone()
    wrong = 0
    for i in range(RUNS):
        kv.wait(pts[i])
        if pts[i] != -1:
            nonlocal_pulls += 1
        if (pv[i].view(-1, VPK) < expected[pk[i]]).any():
            wrong += 1
        expected[sk[i]] += sv[i].view(-1, VPK)
    if wrong:
        errs.append(f"w{wid}: monotonic pushes: {wrong} of {RUNS} pulls went backwards")
    if nonlocal_pulls:
        errs.append(f"w{wid}: monotonic pushes: {nonlocal_pulls} pulls were not local despite intent + WaitSync")
    kv.waitall()
    kv.barrier()

    # ---------------- (3) eventual consistency
    before = torch.zeros(NUM_KEYS * VPK, dtype=dt)
    kv.wait(kv.pull(all_keys, before))
    kv.barrier()
    total = torch.zeros(NUM_KEYS, VPK, dtype=dt)
    loc = [_random_keys(rng) for _ in range(RUNS)]
    sk = [_random_keys(rng) for _ in range(RUNS)]
    sts = [None] * RUNS
    fut = 0
    for i in range(RUNS):
        while fut <= i + AHEAD and fut < RUNS:
            kv.intent(loc[fut], kv.current_clock() + fut - i)
            fut += 1
        sv[i] = torch.randint(1, 1000, (sk[i].numel() * VPK,), generator=g, dtype=torch.int64).to(dt)
        total[sk[i]] -= sv[i].view(-1, VPK)
        sts[i] = kv.push(sk[i], sv[i], True)
        if i > MAX_CONC:
            kv.wait(sts[i - MAX_CONC])
        kv.advance_clock()
    kv.push(all_keys, total.view(-1))
    kv.waitall(); kv.wait_sync(); kv.barrier()
    kv.wait_sync(); kv.barrier()
    after = torch.zeros(NUM_KEYS * VPK, dtype=dt)
    kv.wait(kv.pull(all_keys, after))
    nd = int((before != after).sum().item())
    if nd:
        errs.append(f"w{wid}: eventual consistency: {nd} of {NUM_KEYS * VPK} values differ")
    kv.waitall()
    kv.barrier()
    kv.finalize()
    return errs


@pytest.mark.parametrize("mode,technique", [("threads", "all"), ("procs", "all"), ("threads", "replication_only"),
                                            ("threads", "relocation_only")])
def test_many_key_operations(mode, technique):
    res = run_cluster(_worker, world=4, workers=2, mode=mode, value_lengths=VPK, num_keys=NUM_KEYS, dtype="int64",
                      options={"sys.techniques": technique, "sys.location_caches": 1})
    errs = [e for r in res.values() for k, v in r.items() if k != "counters" for e in v]
    if technique == "relocation_only":
        # with relocation only, two nodes with overlapping intent cannot both be local
        errs = [e for e in errs if "were not local" not in e]
    assert not errs, "\n".join(errs)
    assert all(r["counters"]["protocol_errors"] == 0 for r in res.values())


@pytest.mark.parametrize("opts", [{"sys.sync.idle_period": 1, "sys.sync.sweep_period": 0},
                                  {"sys.sync.idle_period": 7, "sys.sync.sweep_period": 3}])
def test_many_key_operations_sync_pacing_options(opts):
    """The replica-maintenance pacing knobs (rolling sweep, idle-replica check period) change when work is done,
    never what the store returns."""
    o = {"sys.techniques": "all"}
    o.update(opts)
    res = run_cluster(_worker, world=4, workers=2, mode="threads", value_lengths=VPK, num_keys=NUM_KEYS, dtype="int64",
                      options=o)
    errs = [e for r in res.values() for k, v in r.items() if k != "counters" for e in v]
    assert not errs, "\n".join(errs)
    assert all(r["counters"]["protocol_errors"] == 0 for r in res.values())
