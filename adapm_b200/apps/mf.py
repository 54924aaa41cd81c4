"""Matrix-factorisation application (reference apps/matrix_factorization.cc), same flag names.

    python -m adapm_b200.launch -s 2 -m adapm_b200.apps.mf -- --dataset train.mmc --rank 10 --algorithm dsgd --epochs 10
    python -m adapm_b200.apps.mf --synthetic_nnz 2000000 --num_rows 100000 --num_cols 20000 --rank 128
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch

import adapm_b200 as ad
from adapm_b200.apps._common import (add_ablation_options, add_system_options, enforce_full_replication, id_permutation,
                                      strip_dashes, system_options)
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
    ap.add_argument("--increase_step_factor", type=float, default=1.05, help="bold driver: factor after a successful epoch")
    ap.add_argument("--decrease_step_factor", type=float, default=0.5, help="bold driver: factor after an unsuccessful epoch")
    ap.add_argument("--signal_intent_rows", type=int, default=1, help="localise this rank's row parameters (default on)")
    ap.add_argument("--wor_blocks", type=int, default=1, help="WOR schedule for the DSGD blocks")
    ap.add_argument("--wor_points", type=int, default=1, help="WOR schedule for the data points")
    ap.add_argument("--compute_loss", type=int, default=1, help="report the training (and test, when <dataset> has a test file) loss")
    ap.add_argument("--init_parameters", type=int, default=2,
                    help="0: no init, 1: read W.mma / H.mma next to the dataset, 2: draw random factors")
    ap.add_argument("--early_stop", type=int, default=0, help="stop an epoch after N data points (debugging)")
    ap.add_argument("--max_runtime", type=float, default=float("inf"), help="stop after this many seconds")
    ap.add_argument("--prevent_full_model_pull", type=int, default=0,
                    help="accepted for parity: losses are accumulated inside the step, the model is never pulled in full")
    ap.add_argument("--write_generated_factors", default="", help="path prefix to write the initial factors to")
    add_ablation_options(ap)
    add_system_options(ap)
    args = ap.parse_args(strip_dashes(argv if argv is not None else sys.argv[1:]))
    world = int(__import__("os").environ.get("WORLD_SIZE", "1"))
    rank = int(__import__("os").environ.get("RANK", "0"))
    test = None
    if args.dataset:
        i, j, x, m, n = read_matrix_market_coo(args.dataset)
        if args.enforce_random_keys:      # row / column ids are labels: relabel them randomly (same on every rank)
            fr, fc = id_permutation(m, args.model_seed, True), id_permutation(n, args.model_seed + 1, True)
            i, j = fr[i], fc[j]
        data = SparseMatrix(i, j, x, m, n, world, rank)
        tpath = args.dataset.replace("train", "test")
        if args.compute_loss and tpath != args.dataset and os.path.exists(tpath):
            ti, tj, tx, _, _ = read_matrix_market_coo(tpath)
            if args.enforce_random_keys:
                ti, tj = fr[ti], fc[tj]
            test = (ti[rank::world], tj[rank::world], tx[rank::world])
    else:
        m, n = args.num_rows, args.num_cols
        data = SparseMatrix.synthetic(m, n, args.synthetic_nnz or 200000, 8, world, rank, seed=args.model_seed)
    cfg = MFConfig(num_rows=m, num_cols=n, rank=args.rank, algorithm=args.algorithm, eps=args.eps, lam=args.lam,
                   batch_nnz=args.batch_nnz, signal_intent_cols=args.signal_intent_cols, bold_driver=bool(args.bold_driver),
                   model_seed=args.model_seed, eps_inc=args.increase_step_factor, eps_dec=args.decrease_step_factor,
                   signal_intent_rows=bool(args.signal_intent_rows), wor_blocks=bool(args.wor_blocks),
                   wor_points=bool(args.wor_points), early_stop=args.early_stop)
    ad.setup(cfg.num_keys(world), 1)   # one worker per rank: the per-thread loops of the reference are batched kernels here
    server = ad.Server(cfg.row_len, backend=args.backend, options=system_options(args),
                       dtype={"float": "float32", "double": "float64"}[args.value_type])
    kv = ad.Worker(0, server)
    model = MatrixFactorization(server, kv, cfg, data)
    if args.init_parameters == 2:
        model.init_model()
    elif args.init_parameters == 1:
        d = os.path.dirname(args.dataset or ".")
        model.init_model()                 # row intents etc.; the factors are overwritten from the files
        model.load_factors(os.path.join(d, "W.mma"), os.path.join(d, "H.mma"))
    if args.write_generated_factors:
        model.write_factors(args.write_generated_factors)
    if args.enforce_full_replication:
        enforce_full_replication(kv, cfg.num_keys(world))
    kv.barrier()
    prev, t0 = None, time.time()
    for epoch in range(args.epochs):
        loss = model.run_epoch(epoch)
        kv.barrier()
        model.bold_driver(loss, prev)
        prev = loss
        if args.compute_loss:
            msg = f"[mf] epoch {epoch}: local squared error {loss:.6g}, eps {model.eps:.4g}"
            if test is not None:
                kv.wait_sync()
                tl = server.allreduce_sum([model.evaluate(*test), float(len(test[2]))])
                msg += f", test rmse {float((tl[0] / max(1.0, float(tl[1]))) ** 0.5):.4f}"
            if server.my_rank() == 0:
                print(msg + f" ({time.time() - t0:.1f}s)", flush=True)
        if time.time() - t0 > args.max_runtime:
            break
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
