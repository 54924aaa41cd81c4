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


def run_worker(cid: int, server, args, out) -> None:
    kv = ad.Worker(cid, server)
    wid = server.my_rank() * args.num_threads + cid
    vals = torch.zeros(args.vpk, dtype=torch.float64)
    for x in range(args.num_iterations):
        key = torch.tensor([x % args.num_keys])
        kv.intent(key, kv.current_clock())
        push = torch.arange(1, args.vpk + 1, dtype=torch.float64)
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
    ap.add_argument("--sys.techniques", dest="techniques", default="all")
    args = ap.parse_args([a for a in (argv if argv is not None else sys.argv[1:]) if a != "--"])
    ad.setup(args.num_keys, args.num_threads, use_techniques=args.techniques)
    server = ad.Server(args.vpk, dtype="float64", backend="cpu" if not ad._C.cuda_available() else None)
    out = {}
    ths = [threading.Thread(target=run_worker, args=(c, server, args, out)) for c in range(args.num_threads)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    server.shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
