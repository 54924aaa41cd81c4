"""Locality API contract (assertions of the reference's tests/test_locality_api.cc:31-140)."""
import pytest
import torch

from harness import run_cluster

NUM_LOCAL = 2


def _worker(kv, server, wid):
    num_workers = server.num_servers() * NUM_LOCAL
    errs = []
    kv.barrier()
    vals = torch.zeros(1, dtype=server.dtype)

    def expect(cond, msg):
        if not cond:
            errs.append(f"worker {wid}: {msg}")

    if wid == 0:  # rank 0 holds keys 0,3,6 (key % 3)
        expect(kv.pull_if_local(0, vals), "initial locality of key 0")
        expect(kv.pull_if_local(3, vals), "initial locality of key 3")
        expect(not kv.pull_if_local(4, vals), "initial locality of key 4")
        expect(not kv.pull_if_local(5, vals), "initial locality of key 5")
        expect(kv.pull_if_local(6, vals), "initial locality of key 6")
    if wid == num_workers - 1:  # rank 2
        expect(not kv.pull_if_local(1, vals), "initial locality of key 1")
        expect(not kv.pull_if_local(3, vals), "initial locality of key 3")
        expect(not kv.pull_if_local(4, vals), "initial locality of key 4")
        expect(kv.pull_if_local(5, vals), "initial locality of key 5")
        expect(not kv.pull_if_local(6, vals), "initial locality of key 6")
    kv.barrier()

    if wid == 0:
        kv.wait(kv.intent(torch.tensor([3, 4]), 1))
    if wid == num_workers - 1:
        kv.wait(kv.intent(torch.tensor([1, 3, 6]), 1))
    kv.advance_clock()
    kv.wait_sync()
    kv.barrier()

    if wid == 0:
        expect(kv.pull_if_local(0, vals), "changed locality of key 0")
        expect(kv.pull_if_local(3, vals), "changed locality of key 3 (replicated)")
        expect(kv.pull_if_local(4, vals), "changed locality of key 4 (relocated)")
        expect(not kv.pull_if_local(5, vals), "changed locality of key 5")
        expect(not kv.pull_if_local(6, vals), "changed locality of key 6 (relocated away)")
    if wid == num_workers - 1:
        expect(not kv.pull_if_local(0, vals), "changed locality of key 0")
        expect(kv.pull_if_local(1, vals), "changed locality of key 1 (relocated)")
        expect(kv.pull_if_local(3, vals), "changed locality of key 3 (replicated)")
        expect(not kv.pull_if_local(4, vals), "changed locality of key 4")
        expect(kv.pull_if_local(5, vals), "changed locality of key 5")
        expect(kv.pull_if_local(6, vals), "changed locality of key 6 (relocated)")

    # after the intents expire: relocated keys stay, replicas vanish
    kv.advance_clock()
    kv.wait_sync()
    if wid == 0:
        expect(kv.pull_if_local(0, vals), "no-intent locality of key 0")
        expect(kv.pull_if_local(4, vals), "no-intent locality of key 4")
        expect(not kv.pull_if_local(6, vals), "no-intent locality of key 6")
    if wid == num_workers - 1:
        expect(kv.pull_if_local(1, vals), "no-intent locality of key 1")
        expect(kv.pull_if_local(5, vals), "no-intent locality of key 5")
        expect(kv.pull_if_local(6, vals), "no-intent locality of key 6")

    # IsFinished / timestamps: -1 for local ops, a real timestamp for remote ones
    if wid == 0:
        v2 = torch.zeros(2, dtype=server.dtype)
        for _ in range(10):
            ts = kv.pull(torch.tensor([7]), vals, True)
            expect(ts != -1, "remote request returned -1")
            kv.wait(ts)
            expect(kv.is_finished(ts), "remote request isn't finished after wait")
            ts2 = kv.pull(torch.tensor([0, 4]), v2, True)
            expect(kv.is_finished(ts2), "local request isn't finished right away")
            expect(ts2 == -1, "local request returned a timestamp different from -1")
    kv.finalize()
    return errs


@pytest.mark.parametrize("mode", ["threads", "procs"])
def test_locality_api(mode):
    res = run_cluster(_worker, world=3, workers=NUM_LOCAL, mode=mode, value_lengths=1, num_keys=12, dtype="int64")
    errs = [e for r in res.values() for k, v in r.items() if k != "counters" for e in v]
    assert not errs, "\n".join(errs)


# ---- many distinct value lengths (reference: any value_lengths vector, coloc_kv_server_handle.h:162-170,996-999):
# more distinct lengths than size classes -> keys of different lengths share a slab class and carry their own length
def _many_lengths_worker(kv, server, wid):
    import torch

    nk = server.num_keys()
    lens = torch.tensor([(k * 7) % 97 + 1 for k in range(nk)], dtype=torch.int64)
    errors = []
    keys = torch.arange(nk, dtype=torch.int64)
    total = int(lens.sum())
    kv.barrier()
    if wid == 0:
        vals = torch.cat([torch.full((int(lens[k]),), float(k + 1), dtype=server.dtype) for k in range(nk)])
        kv.wait(kv.set(keys, vals))
    kv.barrier()
    # everybody intends a different third of the keys: relocations and replicas of rows of many lengths
    mine = keys[(keys % 3) == (wid % 3)]
    kv.intent(mine, kv.current_clock(), kv.current_clock() + 4)
    kv.wait_sync(); kv.barrier(); kv.wait_sync(); kv.barrier()
    for rep in range(3):
        out = torch.zeros(int(lens[mine].sum()), dtype=server.dtype)
        kv.wait(kv.pull(mine, out))
        exp = torch.cat([torch.full((int(lens[k]),), float(k + 1) + rep * server.num_servers() * 2, dtype=server.dtype)
                         for k in mine.tolist()])
        if not torch.equal(out, exp) and rep == 0:
            errors.append(f"worker {wid}: pull of mixed-length rows wrong: {out[:8].tolist()} vs {exp[:8].tolist()}")
        kv.barrier()          # nobody pushes before everybody has read
        kv.wait(kv.push(keys, torch.ones(total, dtype=server.dtype)))
        kv.advance_clock()
        kv.wait_sync(); kv.barrier(); kv.wait_sync(); kv.barrier()
    for _ in range(6):
        kv.advance_clock()
    kv.wait_sync(); kv.barrier(); kv.wait_sync(); kv.barrier()
    out = torch.zeros(total, dtype=server.dtype)
    kv.wait(kv.pull(keys, out))
    n_workers = server.num_servers() * 2
    exp = torch.cat([torch.full((int(lens[k]),), float(k + 1) + 3 * n_workers, dtype=server.dtype) for k in range(nk)])
    if not torch.equal(out, exp):
        bad = (out != exp).nonzero()[:4].view(-1).tolist()
        errors.append(f"worker {wid}: final values wrong at {bad}: {out[bad].tolist()} vs {exp[bad].tolist()}")
    kv.barrier()
    kv.finalize()
    return errors


def test_many_distinct_value_lengths():
    import torch

    nk = 150
    lens = torch.tensor([(k * 7) % 97 + 1 for k in range(nk)], dtype=torch.int64)
    assert lens.unique().numel() > 32     # more distinct lengths than size classes
    res = run_cluster(_many_lengths_worker, world=3, workers=2, mode="threads", value_lengths=lens, num_keys=nk,
                      dtype="float32")
    errs = [e for r in res.values() for k, v in r.items() if k != "counters" for e in v]
    assert not errs, "\n".join(errs)
    assert all(r["counters"]["protocol_errors"] == 0 for r in res.values())
