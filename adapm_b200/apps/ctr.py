"""CTR DeepFM application (BASELINE config #5) on the bindings-style API.

    python -m adapm_b200.launch -s 8 --backend cuda -m adapm_b200.apps.ctr -- --num_features 100000000
    python -m adapm_b200.apps.ctr --backend cpu --num_features 20000 --steps 30 --batch_size 512
"""
from __future__ import annotations

import argparse
import sys
import time

import adapm_b200 as ad
from adapm_b200.apps._common import add_system_options, strip_dashes, system_options
from adapm_b200.models.deepfm import DeepFM, DeepFMConfig, synthetic_ctr_batch


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--num_features", type=int, default=1_000_000)
    ap.add_argument("--num_fields", type=int, default=26)
    ap.add_argument("--embed_dim", type=int, default=16)
    ap.add_argument("--batch_size", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--read_ahead", type=int, default=4)
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16", "fp8"])
    ap.add_argument("--table_precision", default="fp32", choices=["fp32", "fp8_block"],
                    help="fp8_block: train on the values a block-scaled e4m3 embedding table would deliver (fp32 master rows)")
    add_system_options(ap)
    args = ap.parse_args(strip_dashes(argv if argv is not None else sys.argv[1:]))
    cfg = DeepFMConfig(num_features=args.num_features, num_fields=args.num_fields, embed_dim=args.embed_dim,
                       batch_size=args.batch_size, read_ahead=args.read_ahead, precision=args.precision,
                       table_precision=args.table_precision)
    ad.setup(cfg.num_features, 1)
    server = ad.Server(cfg.row_len, backend=args.backend, options=system_options(args))
    kv = ad.Worker(0, server)
    model = DeepFM(server, kv, cfg)
    model.init_model()
    rank = server.my_rank()
    # look-ahead: while step s runs, the distinct feature ids of step s + read_ahead are signalled (parallel/schedules.py)
    from ..parallel.schedules import LookaheadIntents

    batches = {}

    def batch(b):
        if b not in batches:
            batches[b] = synthetic_ctr_batch(cfg, b, rank)
        return batches[b]

    look = LookaheadIntents(kv, args.steps + cfg.read_ahead, cfg.read_ahead, lambda b: batch(b)[0]) if server.num_servers() > 1 else None
    if look:
        look.prime()
    t0, losses = time.time(), []
    for s in range(args.steps):
        if look:
            look.signal(s)
        ids, y = batch(s)
        batches.pop(s - 1, None)
        losses.append(model.step(ids, y))
        kv.advance_clock()
        if rank == 0 and (s + 1) % max(1, args.steps // 5) == 0:
            print(f"[ctr] step {s + 1}: loss {sum(losses[-10:]) / len(losses[-10:]):.4f} "
                  f"({(s + 1) * cfg.batch_size / (time.time() - t0):.0f} examples/s/rank)", flush=True)
    kv.barrier()
    kv.finalize()
    if rank == 0:
        print(server.stats(), flush=True)
    server.shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
