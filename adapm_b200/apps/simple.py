"""``simple`` - the reference's minimal application (apps/simple.cc:36-134), BASELINE config #1.

    python -m adapm_b200.launch -s 2 --backend cpu -m adapm_b200.apps.simple -- -k 10 -t 2 -i 5 -v 2

Every worker, for iteration x: ``Intent({x}, clock)``; ``Wait(Push({x}, [1..v]))``;
``Wait(Pull({x}))``; print; ``advanceClock()``. ValT = double like the reference.
"""
from __future__ import annotations

import argparse
import sys
import threading

import torch

import adapm_b200 as ad
from adapm_b200.apps._common import add_system_options, system_options


def run_worker(cid: int, server, args, out) -> None:
    kv = ad.Worker(cid, server)
    wid = server.my_rank() * args.num_threads + cid
    vals = torch.zeros(args.vpk, dtype=server.dtype)
    for x in range(args.num_iterations):
        key = torch.tensor([x % args.num_keys])
        kv.intent(key, kv.current_clock())
        push = torch.arange(1, args.vpk + 1, dtype=server.dtype)
        kv.wait(kv.push(key, push))
        kv.wait(kv.pull(key, vals))
        print(f"Worker {wid} iteration {x}: key {int(key)} = {vals.tolist()}", flush=True)
        kv.advance_clock()
    kv.barrier()
    out[cid] = vals.clone()
    kv.finalize()


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-k", "--num_keys", type=int, default=10)
    ap.add_argument("-t", "--num_threads", type=int, default=2)
    ap.add_argument("-i", "--num_iterations", type=int, default=4)
    ap.add_argument("-v", "--vpk", type=int, default=2, help="values per key")
    add_system_options(ap)      # sys.* / sampling.* flags by their reference names (+ --backend)
    args = ap.parse_args([a for a in (argv if argv is not None else sys.argv[1:]) if a != "--"])
    ad.setup(args.num_keys, args.num_threads)
    # ValT = double like the reference app: float64 rows live on the CPU backend (the CUDA backend stores float32)
    backend = args.backend or "cpu"
    server = ad.Server(args.vpk, dtype="float64" if backend == "cpu" else "float32", backend=backend,
                       options=system_options(args))
    out = {}
    ths = [threading.Thread(target=run_worker, args=(c, server, args, out)) for c in range(args.num_threads)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    server.shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
