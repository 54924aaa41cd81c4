"""Sampling contract (assertions of the reference's tests/test_sampling.cc:40-289):
values of sampled keys are correct, with-replacement sampling reproduces the exact key
frequencies of a deterministic skewed sequence (for the local scheme: the sequence filtered to
locally resident keys), without-replacement samples are unique, partial pulls of 1..5 keys work,
and a third of the keys is never sampled. Schemes: naive, preloc, pool, local (run_tests.sh:34-40).
"""
import random
import threading

import pytest
import torch

from harness import run_cluster

NUM_KEYS = 60
VPK = 5
SAMPLES_PER_WORKER = 40
KEYS_PER_SAMPLE = 8
WORKERS = 2
WORLD = 3
AHEAD = 10


def _sequence(rank, total):
    rng = random.Random(17727 ^ rank)
    no = set()
    while len(no) != NUM_KEYS // 3:
        no.add(rng.randrange(NUM_KEYS))
    seq = []
    while len(seq) != total:
        k = rng.randrange(NUM_KEYS)
        if k in no:
            continue
        if rng.randrange(NUM_KEYS) < k:
            continue
        seq.append(k)
    return seq, no


class _SeqSampler:
    def __init__(self, seq):
        self.seq, self.pos, self.mu = seq, 0, threading.Lock()

    def __call__(self):
        with self.mu:
            k = self.seq[self.pos % len(self.seq)]
            self.pos += 1
            return k


def _make_setup(scheme, wr):
    def setup(server):
        total = SAMPLES_PER_WORKER * KEYS_PER_SAMPLE * WORKERS
        seq, no = _sequence(server.my_rank(), total)
        server._test_seq, server._test_no = seq, no
        server.enable_sampling_support(scheme=scheme, with_replacement=wr, sample_fn=_SeqSampler(seq))
    return setup


def _worker(kv, server, wid):
    errs = []
    dt = server.dtype
    if wid == 0:
        keys = torch.arange(NUM_KEYS)
        vals = (keys.view(-1, 1) * 100 + torch.arange(VPK).view(1, -1)).to(dt).contiguous().view(-1)
        kv.wait(kv.push(keys, vals))
    kv.wait_sync(); kv.barrier(); kv.wait_sync()
    rng = random.Random(173727 ^ wid)
    freq = [0] * NUM_KEYS
    ids = [None] * SAMPLES_PER_WORKER
    fut = 0
    wr = server._test_wr
    for i in range(SAMPLES_PER_WORKER):
        while fut <= i + AHEAD and fut < SAMPLES_PER_WORKER:
            ids[fut] = kv.prepare_sample(KEYS_PER_SAMPLE, kv.current_clock() + fut - i)
            fut += 1
        pulled = []
        remaining = KEYS_PER_SAMPLE
        while remaining > 0:
            n = min(rng.randint(1, VPK), remaining)
            k = torch.zeros(n, dtype=torch.int64)
            v = torch.zeros(n * VPK, dtype=dt)
            kv.wait(kv.pull_sample(ids[i], k, v, True))
            expect = (k.view(-1, 1) * 100 + torch.arange(VPK).view(1, -1)).to(dt)
            if not torch.equal(v.view(-1, VPK), expect):
                errs.append(f"w{wid}: wrong values for sampled keys {k.tolist()}")
            for kk in k.tolist():
                freq[kk] += 1
                if kk in server._test_no:
                    errs.append(f"w{wid}: key {kk} must never be sampled")
            pulled += k.tolist()
            remaining -= n
        if not wr and len(set(pulled)) != len(pulled):
            errs.append(f"w{wid}: without-replacement sample is not unique: {pulled}")
        kv.finish_sample(ids[i])
        kv.advance_clock()
    kv.barrier()
    kv.finalize()
    return {"errs": errs, "freq": freq}


@pytest.mark.parametrize("scheme,wr", [("naive", True), ("naive", False), ("preloc", True), ("preloc", False),
                                       ("pool", True), ("local", True), ("local", False)])
def test_sampling(scheme, wr):
    setup = _make_setup(scheme, wr)

    def setup2(server):
        server._test_wr = wr
        setup(server)

    res = run_cluster(_worker, world=WORLD, workers=WORKERS, mode="threads", setup_fn=setup2, value_lengths=VPK,
                      num_keys=NUM_KEYS, dtype="int64", options={"sampling.batch_size": 1, "sampling.pool_size": 16,
                                                                 "sampling.reuse": 3})
    errs = [e for r in res.values() for k, v in r.items() if k != "counters" for e in v["errs"]]
    assert not errs, "\n".join(errs[:10])
    if wr and scheme != "pool":
        total = SAMPLES_PER_WORKER * KEYS_PER_SAMPLE * WORKERS
        for rank, r in res.items():
            seq, _ = _sequence(rank, total)
            got = [sum(r[c]["freq"][k] for c in range(WORKERS)) for k in range(NUM_KEYS)]
            want = [0] * NUM_KEYS
            if scheme == "local":
                n, pos = 0, 0
                while n != total:
                    k = seq[pos % len(seq)]
                    pos += 1
                    if k % WORLD == rank:
                        want[k] += 1
                        n += 1
            else:
                for k in seq:
                    want[k] += 1
            assert got == want, f"rank {rank}: sampling frequencies differ from the distribution sequence"
