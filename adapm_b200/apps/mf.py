"""Matrix-factorisation application (reference apps/matrix_factorization.cc), same flag names.

    python -m adapm_b200.launch -s 2 -m adapm_b200.apps.mf -- --dataset train.mmc --rank 10 --algorithm dsgd --epochs 10
    python -m adapm_b200.apps.mf --synthetic_nnz 2000000 --num_rows 100000 --num_cols 20000 --rank 128
"""
from __future__ import annotations

import argparse
import sys
import time

import numpy as np
import torch

import adapm_b200 as ad
from adapm_b200.apps._common import add_system_options, strip_dashes, system_options
from adapm_b200.models.mf import MatrixFactorization, MFConfig, SparseMatrix
from adapm_b200.utils.mmio import read_matrix_market_coo


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--dataset", default=None, help="MatrixMarket coordinate file")
    ap.add_argument("--synthetic_nnz", type=int, default=0)
    ap.add_argument("--num_rows", type=int, default=10000)
    ap.add_argument("--num_cols", type=int, default=2000)
    ap.add_argument("--rank", "-r", type=int, default=10)
    ap.add_argument("--epochs", type=int, default=10)
    ap.add_argument("--num_threads", "-t", type=int, default=1)
    ap.add_argument("--algorithm", default="dsgd", choices=["dsgd", "columnwise", "plain"])
    ap.add_argument("--eps", type=float, default=0.01)
    ap.add_argument("--lambda", dest="lam", type=float, default=0.05)
    ap.add_argument("--signal_intent_cols", type=int, default=1000)
    ap.add_argument("--batch_nnz", type=int, default=1 << 18)
    ap.add_argument("--bold_driver", type=int, default=1)
    ap.add_argument("--export_prefix", default="")
    ap.add_argument("--model_seed", type=int, default=134827)
    add_system_options(ap)
    args = ap.parse_args(strip_dashes(argv if argv is not None else sys.argv[1:]))
    world = int(__import__("os").environ.get("WORLD_SIZE", "1"))
    rank = int(__import__("os").environ.get("RANK", "0"))
    if args.dataset:
        i, j, x, m, n = read_matrix_market_coo(args.dataset)
        data = SparseMatrix(i, j, x, m, n, world, rank)
    else:
        m, n = args.num_rows, args.num_cols
        data = SparseMatrix.synthetic(m, n, args.synthetic_nnz or 200000, 8, world, rank, seed=args.model_seed)
    cfg = MFConfig(num_rows=m, num_cols=n, rank=args.rank, algorithm=args.algorithm, eps=args.eps, lam=args.lam,
                   batch_nnz=args.batch_nnz, signal_intent_cols=args.signal_intent_cols, bold_driver=bool(args.bold_driver),
                   model_seed=args.model_seed)
    ad.setup(cfg.num_keys(world), args.num_threads)
    server = ad.Server(cfg.row_len, backend=args.backend, options=system_options(args))
    kv = ad.Worker(0, server)
    model = MatrixFactorization(server, kv, cfg, data)
    model.init_model()
    kv.barrier()
    prev, t0 = None, time.time()
    for epoch in range(args.epochs):
        loss = model.run_epoch(epoch)
        kv.barrier()
        model.bold_driver(loss, prev)
        prev = loss
        if server.my_rank() == 0:
            print(f"[mf] epoch {epoch}: local squared error {loss:.6g}, eps {model.eps:.4g} ({time.time() - t0:.1f}s)", flush=True)
    if args.export_prefix:
        model.write_factors(args.export_prefix)
    kv.barrier()
    kv.finalize()
    if server.my_rank() == 0:
        print(server.stats(), flush=True)
    server.shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
