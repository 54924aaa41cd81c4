"""Device-timed updates/s of the KGE (ComplEx, FB15k scale, d=512) and MF (10M x 1M, d=128) training loops on N
GPUs (BASELINE.json configs 3 and 4). Launch with torchrun; one JSON line per app from rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/app_bench_multi.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adapm_b200 as ad  # noqa: E402
from adapm_b200.models.kge import KGE, KGEConfig, synthetic_triples  # noqa: E402
from adapm_b200.models.mf import MatrixFactorization, MFConfig, SparseMatrix  # noqa: E402
from adapm_b200.ops import mf_step  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[lr])
        torch.cuda.synchronize()

    def max_ms(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    fab = "shm" if world > 1 else "inproc"
    K, W, RA = 100, 10, 8
    # ------------------------------------------------------------------ KGE
    cfg = KGEConfig(embed_dim=512, batch_triples=8192, read_ahead=RA)
    server = ad.Server(cfg.value_lengths(), num_keys=cfg.num_keys, num_threads=1, rank=rank, world=world, backend="cuda",
                       fabric=fab, device=lr, job=ad.default_job() + "k")
    kv = ad.Worker(0, server)
    model = KGE(server, kv, cfg)
    model.init_model()
    tr = synthetic_triples(cfg, cfg.batch_triples * 32, seed=100 + rank)
    batches = [tr[i * cfg.batch_triples:(i + 1) * cfg.batch_triples].pin_memory() for i in range(32)]

    def kge_step(s):
        model.signal_intent(batches[(s + RA) % 32], kv.current_clock() + RA)
        model.step(batches[s % 32])
        kv.advance_clock()
        if s % 3 == 2:
            torch.cuda.current_stream().synchronize() if False else None

    for s in range(RA):
        model.signal_intent(batches[s], kv.current_clock() + s)
    kv.wait_sync() if world > 1 else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    evq = []
    for s in range(W):
        kge_step(s)
    barrier()
    ev[0].record()
    for s in range(W, W + K):
        kge_step(s)
        e = torch.cuda.Event(); e.record(); evq.append(e)
        if len(evq) > 3:
            evq.pop(0).synchronize()      # bounded run-ahead
    ev[1].record()
    barrier()
    ms = max_ms(ev[0].elapsed_time(ev[1])) / K
    st = model.stats.tolist()
    if rank == 0:
        upd = world * cfg.batch_triples * cfg.updates_per_triple
        print(json.dumps({"bench": "kge_complex", "n_gpus": world, "config": "FB15k scale, d=512, neg_ratio=6, 8192 triples/GPU/step",
                          "ms_per_step": ms, "updates_per_s": upd / ms * 1e3,
                          "rows_local_remote_slow": st[:3],
                          "pm": {k: v for k, v in server.counters().items()
                                 if k in ("relocations", "replica_setups", "refreshes", "protocol_errors")}}), flush=True)
    kv.finalize(); server.shutdown()
    barrier()

    # ------------------------------------------------------------------ MF
    cfg = MFConfig(num_rows=10_000_000, num_cols=1_000_000, rank=128, algorithm="dsgd", batch_nnz=1 << 18)
    nnz_per_rank = (1 << 18) * 24
    # this rank's row block x all column blocks, uniform non-zeros
    rng = np.random.default_rng(5 + rank)
    rpb = (cfg.num_rows + world - 1) // world
    cpb = (cfg.num_cols + world - 1) // world
    server = ad.Server(cfg.row_len, num_keys=cfg.num_keys(world), num_threads=1, rank=rank, world=world, backend="cuda",
                       fabric=fab, device=lr, job=ad.default_job() + "m")
    kv = ad.Worker(0, server)
    fck = cfg.first_col_key(world)
    # init: every rank sets the keys it is home for (device path)
    kv.begin_setup()
    keys = torch.arange(rank, cfg.num_keys(world), world, dtype=torch.int64)
    keys = keys[(keys < cfg.num_rows) | (keys >= fck)]
    gen = torch.Generator(device=dev).manual_seed(rank)
    for s in range(0, keys.numel(), 1 << 18):
        k = keys[s:s + (1 << 18)].to(dev)
        rows = torch.zeros(k.numel(), cfg.row_len, device=dev)
        rows[:, :cfg.rank] = torch.rand(k.numel(), cfg.rank, generator=gen, device=dev) / cfg.rank ** 0.5
        kv.set(k, rows.view(-1))
    kv.waitall(); kv.end_setup()
    if world > 1:   # row intents for the whole run + the DSGD column block of every sub-epoch
        kv.intent(torch.arange(rank * rpb, min(cfg.num_rows, (rank + 1) * rpb)), 0, ad.CLOCK_MAX)
    loss = torch.zeros(1, device=dev); stats = torch.zeros(4, dtype=torch.int64, device=dev)
    n = cfg.batch_nnz
    steps_per_block = 32
    packs = {}
    for b in range(world):
        packs[b] = []
        for s in range(steps_per_block):
            i = torch.from_numpy(rng.integers(rank * rpb, min(cfg.num_rows, (rank + 1) * rpb), n)).to(dev)
            j = torch.from_numpy(rng.integers(b * cpb, min(cfg.num_cols, (b + 1) * cpb), n)).to(dev) + fck
            packs[b].append((i, j, torch.randn(n, device=dev), torch.full((n,), 20, dtype=torch.int32, device=dev),
                             torch.full((n,), 200, dtype=torch.int32, device=dev)))

    def subepoch(se):
        b = (rank + se) % world                     # DSGD stratum: no two ranks share a column block
        if world > 1:
            kv.intent(torch.arange(b * cpb, min(cfg.num_cols, (b + 1) * cpb)) + fck, kv.current_clock())
            kv.wait_sync()
        for p in packs[b]:
            mf_step(server, *p, cfg.rank, 0.01, 0.05, loss, stats)
        kv.advance_clock()
        torch.cuda.synchronize()
        kv.barrier()

    subepoch(0)
    barrier()
    ev[0].record()
    n_se = max(2, world)
    for se in range(1, 1 + n_se):
        subepoch(se)
    ev[1].record()
    barrier()
    ms = max_ms(ev[0].elapsed_time(ev[1]))
    st = stats.tolist()
    if rank == 0:
        upd = world * n_se * steps_per_block * 2 * n
        print(json.dumps({"bench": "mf_dsgd", "n_gpus": world,
                          "config": "10M x 1M, rank 128, DSGD sub-epochs (intent + WaitSync + barrier per sub-epoch included)",
                          "ms_total": ms, "updates_per_s": upd / ms * 1e3, "rows_local_remote_slow": st[:3]}), flush=True)
    kv.finalize(); server.shutdown()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
