"""NCCL-only comparison arm (``bench.py --impl nccl``): the same word2vec SGNS metric and config
implemented with stock PyTorch ops and NCCL collectives only. No kernel, engine or model code of
adapm_b200 is on the timed path (the synthetic corpus generator is shared so that both arms see the
same batches). BASELINE.md section 2 item 3: "a path that only calls NCCL for the named ops is the
baseline, not the product".

The arm is built to be a *fair* anchor, not a strawman:

* static hash partitioning by word (``(key // 2) % world``: the syn0 and syn1 row of a word live on the same GPU, every
  GPU owns 1/world of BOTH tables), one dense ``[keys/world, 2d]`` fp32 shard per GPU;
* everything a data loader can prepare is prepared outside the timed loop, exactly like the native arm's loader
  prepares the distinct keys of a batch: the distinct keys of every batch sorted by owner, the inverse map, the
  per-destination split sizes - and the receive sizes of every step are exchanged once at start-up, so the timed
  loop has **no host synchronisation at all** (no ``.tolist()`` / ``.item()``);
* negatives are drawn from the *local* shard (unigram^0.75 restricted to the keys this rank owns) - the same
  locality trick the native arm's ``local`` sampling scheme uses - so only centers and contexts travel;
* per step: H2D of the prepared keys (pinned) -> ``all_to_all`` keys -> ``index_select`` at the owners ->
  ``all_to_all`` rows -> SGNS/AdaGrad math with PyTorch ops -> duplicate updates pre-aggregated with
  ``index_add_`` -> ``all_to_all`` updates -> ``index_add_`` at the owners.

What it cannot do (by construction): keep hot rows local between steps, overlap the exchanges with the math, or
avoid materialising the gathered ``[B, 1+neg, 2d]`` operands - that is the work the fused kernels remove.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch
import torch.distributed as dist


def run_nccl_word2vec(args, rank: int, world: int, local_rank: int) -> int:
    from ..models.word2vec import SyntheticPairs, Word2VecConfig, zipf_counts

    dev = torch.device("cuda", local_rank)
    V, d, neg, B = args.vocab, args.dim, args.negative, args.batch_pairs
    cfg = Word2VecConfig(vocab_size=V, embed_dim=d, negative=neg, batch_pairs=B)
    nkeys = 2 * V
    n_local = 2 * ((V + world - 1) // world)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    table = torch.empty(n_local, 2 * d, device=dev)
    table[:, :d] = (torch.rand(n_local, d, generator=g, device=dev) - 0.5) / d
    table[:, d:] = 1e-6
    counts = zipf_counts(V, cfg.zipf_exponent)
    data = SyntheticPairs(cfg, counts, rank, seed=1)
    # key -> (owner, shard row): words are dealt round-robin, both rows of a word on its owner
    def owner_of(k):
        return (k >> 1) % world

    def shard_row(k):
        return ((k >> 1) // world) * 2 + (k & 1)

    # local negative sampling: unigram^0.75 over the syn1 keys (2w + 1) this rank owns
    w_all = np.arange(V, dtype=np.int64)
    mine = (w_all % world) == rank
    loc_rows = torch.from_numpy(shard_row(2 * w_all[mine] + 1)).to(dev)            # shard row of every local syn1 key
    pw = counts[mine] ** 0.75
    ncdf = torch.from_numpy(np.cumsum(pw / pw.sum())).to(dev)
    alpha = 0.025
    K, W = args.steps, args.warmup
    n_steps = W + K

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ loader (outside the timed loop)
    prepared = []
    send_mat = torch.zeros(n_steps, world, dtype=torch.int64)
    for s in range(n_steps):
        kb = data.batch(s)                                       # [2, B]: syn0 keys of the centers, syn1 keys of the contexts
        uk, inv = torch.unique(kb.reshape(-1), return_inverse=True)
        owner = owner_of(uk)
        order = torch.argsort(owner, stable=True)
        rank_of = torch.empty_like(order)
        rank_of[order] = torch.arange(order.numel())
        send_mat[s] = torch.bincount(owner, minlength=world)
        # shard rows to request (owner order), position of every pair's row in that order
        prepared.append((shard_row(uk[order]).pin_memory(), rank_of[inv].pin_memory()))
    if world > 1:
        # recv_mat[s][p] = number of rows rank p requests from me at step s (everybody's send matrix, column `rank`)
        sm = send_mat.to(dev)
        allm = [torch.empty_like(sm) for _ in range(world)]
        dist.all_gather(allm, sm)
        recv_mat = torch.stack([m[:, rank] for m in allm], 1).cpu()
    else:
        recv_mat = send_mat.clone()
    send_l, recv_l = send_mat.tolist(), recv_mat.tolist()
    label = torch.zeros(B, 1 + neg, device=dev)
    label[:, 0] = 1
    gen = torch.Generator(device=dev).manual_seed(99 + rank)

    def step(s):
        req_h, pos_h = prepared[s]
        req = req_h.to(dev, non_blocking=True)                  # H2D of this step's prepared inputs
        pos = pos_h.to(dev, non_blocking=True)
        sc, rc = send_l[s], recv_l[s]
        if world > 1:
            got_req = req.new_empty(sum(rc))
            dist.all_to_all_single(got_req, req, rc, sc)                                  # row requests -> owners
        else:
            got_req = req
        rows = table.index_select(0, got_req)                                             # owner-side gather
        if world > 1:
            urows = rows.new_empty(sum(sc), 2 * d)
            dist.all_to_all_single(urows, rows, sc, rc)                                   # rows -> requesters
        else:
            urows = rows
        nrow = loc_rows[torch.searchsorted(ncdf, torch.rand(B * neg, generator=gen, device=dev, dtype=torch.float64)
                                           ).clamp_(max=loc_rows.numel() - 1)]           # local negatives
        cr = urows.index_select(0, pos[:B])
        pr = urows.index_select(0, pos[B:])
        nr = table.index_select(0, nrow).view(B, neg, 2 * d)
        e0, a0 = cr[:, :d], cr[:, d:]
        e1 = torch.cat([pr[:, :d].unsqueeze(1), nr[:, :, :d]], 1)                          # [B, 1+neg, d]
        a1 = torch.cat([pr[:, d:].unsqueeze(1), nr[:, :, d:]], 1)
        f = torch.einsum("bd,btd->bt", e0, e1)
        gsc = label - torch.sigmoid(f)
        gsc = torch.where(f > 6, label - 1, gsc)
        gsc = torch.where(f < -6, label, gsc)
        grad0 = torch.einsum("bt,btd->bd", gsc, e1)
        grad1 = gsc.unsqueeze(-1) * e0.unsqueeze(1)
        g1s = grad1 * grad1
        u1 = torch.cat([alpha * grad1 * torch.rsqrt(a1 + g1s), g1s], -1)                  # [B, 1+neg, 2d]
        g0s = grad0 * grad0
        u0 = torch.cat([alpha * grad0 * torch.rsqrt(a0 + g0s), g0s], -1)
        upd = torch.zeros_like(urows)
        upd.index_add_(0, pos[:B], u0)                                                    # duplicates aggregated before they travel
        upd.index_add_(0, pos[B:], u1[:, 0])
        if world > 1:
            back = upd.new_empty(sum(rc), 2 * d)
            dist.all_to_all_single(back, upd, rc, sc)                                     # updates -> owners
        else:
            back = upd
        table.index_add_(0, got_req, back)                                                # owner-side apply
        table.index_add_(0, nrow, u1[:, 1:].reshape(-1, 2 * d))                           # negatives are local
        return torch.log1p(torch.exp(-torch.where(label > 0.5, f, -f).clamp(-6, 6))).sum()

    for s in range(W):
        step(s)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s in range(W, W + K):
        loss = step(s)
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    upd = world * B * (neg + 2)
    if rank == 0:
        print(json.dumps({"metric": "word2vec SGNS updates/sec (device-timed, max over ranks)", "impl": "nccl",
                          "value": upd * K / (ms * 1e-3), "unit": "updates/s", "n_gpus": world, "steps": K, "warmup": W,
                          "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "dtype": "fp32",
                          "data": "synthetic",
                          "config": {"model": "word2vec SGNS 1M-vocab d=300", "vocab": V, "embed_dim": d, "negative": neg,
                                     "global_batch": world * B,
                                     "parallelism": f"hash-partitioned table, NCCL all_to_all x{world}, stock PyTorch ops",
                                     "loader": "distinct keys / owner order / split sizes prepared outside the timed loop; "
                                               "no host sync inside it; negatives drawn from the local shard"},
                          "loss_last": float(loss) / (B * (neg + 1))}))
    if world > 1:
        dist.destroy_process_group()
    return 0
