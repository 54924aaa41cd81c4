"""NCCL-only baseline arm (``bench.py --impl nccl``): the same word2vec SGNS metric and config
implemented with stock PyTorch ops and NCCL collectives only - static hash partitioning of the
key space, per step an all_to_all of keys, index_select at the owners, all_to_all of rows, the
SGNS/AdaGrad math in PyTorch, and an all_to_all of the updates applied with index_add_.
No kernel, engine or model code of adapm_b200 is on this path. BASELINE.md section 2 item 3:
"a path that only calls NCCL for the named ops is the baseline, not the product".
"""
from __future__ import annotations

import json
import os
import statistics

import numpy as np
import torch
import torch.distributed as dist


def _exchange(rows_or_keys: torch.Tensor, send_counts: torch.Tensor, world: int):
    """all_to_all_single with variable splits along dim 0."""
    recv_counts = torch.empty_like(send_counts)
    if world > 1:
        dist.all_to_all_single(recv_counts, send_counts)
    else:
        recv_counts.copy_(send_counts)
    sc, rc = send_counts.tolist(), recv_counts.tolist()
    out = rows_or_keys.new_empty((sum(rc),) + tuple(rows_or_keys.shape[1:]))
    if world > 1:
        dist.all_to_all_single(out, rows_or_keys, rc, sc)
    else:
        out.copy_(rows_or_keys)
    return out, sc, rc


def run_nccl_word2vec(args, rank: int, world: int, local_rank: int) -> int:
    dev = torch.device("cuda", local_rank)
    V, d, neg, B = args.vocab, args.dim, args.negative, args.batch_pairs
    nkeys = 2 * V
    n_local = (nkeys + world - 1) // world
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    table = torch.empty(n_local, 2 * d, device=dev)
    table[:, :d] = (torch.rand(n_local, d, generator=g, device=dev) - 0.5) / d
    table[:, d:] = 1e-6
    # Zipf corpus + unigram^0.75 negatives (same shapes as the native arm)
    r = np.arange(1, V + 1, dtype=np.float64)
    p = r ** -1.0
    p /= p.sum()
    cdf = torch.from_numpy(np.cumsum(p)).to(dev)
    pn = p ** 0.75
    pn /= pn.sum()
    ncdf = torch.from_numpy(np.cumsum(pn)).to(dev)
    alpha = 0.025

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def step(seed):
        gg = torch.Generator(device=dev).manual_seed(seed * 131 + rank)
        w = torch.searchsorted(cdf, torch.rand(2 * B, generator=gg, device=dev, dtype=torch.float64)).clamp_(max=V - 1)
        nw = torch.searchsorted(ncdf, torch.rand(B * neg, generator=gg, device=dev, dtype=torch.float64)).clamp_(max=V - 1)
        centers = 2 * w[:B]
        tk = torch.cat([(2 * w[B:] + 1).view(B, 1), (2 * nw + 1).view(B, neg)], 1)      # [B, 1+neg]
        allk = torch.cat([centers, tk.reshape(-1)])
        uk, inv = torch.unique(allk, return_inverse=True)
        owner = uk % world
        order = torch.argsort(owner)
        uk_sorted = uk[order]
        send_counts = torch.bincount(owner, minlength=world)
        req_keys, sc, rc = _exchange(uk_sorted, send_counts, world)                      # keys -> owners
        rows = table.index_select(0, req_keys // world)                                   # owner-side gather
        got, _, _ = _exchange(rows, torch.tensor(rc, device=dev), world)                  # rows -> requesters
        urows = torch.empty_like(got)
        urows[order] = got
        cr = urows[inv[:B]]
        tr = urows[inv[B:]].view(B, 1 + neg, 2 * d)
        e0, a0 = cr[:, :d], cr[:, d:]
        e1, a1 = tr[:, :, :d], tr[:, :, d:]
        label = torch.zeros(B, 1 + neg, device=dev)
        label[:, 0] = 1
        f = torch.einsum("bd,btd->bt", e0, e1)
        gsc = label - torch.sigmoid(f)
        gsc = torch.where(f > 6, label - 1, gsc)
        gsc = torch.where(f < -6, label, gsc)
        valid = torch.ones_like(gsc, dtype=torch.bool)
        valid[:, 1:] = tk[:, 1:] != tk[:, :1]
        gsc = gsc * valid
        grad0 = torch.einsum("bt,btd->bd", gsc, e1)
        grad1 = gsc.unsqueeze(-1) * e0.unsqueeze(1)
        u1 = torch.cat([alpha * grad1 / torch.sqrt(a1 + grad1 * grad1), grad1 * grad1], -1)
        u0 = torch.cat([alpha * grad0 / torch.sqrt(a0 + grad0 * grad0), grad0 * grad0], -1)
        upd = torch.zeros_like(urows)
        upd.index_add_(0, inv[:B], u0)
        upd.index_add_(0, inv[B:], u1.view(-1, 2 * d))
        back, _, _ = _exchange(upd[order], send_counts, world)                            # updates -> owners
        table.index_add_(0, req_keys // world, back)                                      # owner-side apply
        return (torch.log1p(torch.exp(-torch.where(label > 0.5, f, -f).clamp(-6, 6))) * valid).sum()

    K, W = args.steps, args.warmup
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
        print(json.dumps({"metric": "word2vec SGNS updates/sec (device-timed, max over ranks)", "impl": "nccl-baseline",
                          "value": upd * K / (ms * 1e-3), "unit": "updates/s", "n_gpus": world, "steps": K, "warmup": W,
                          "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "dtype": "fp32",
                          "data": "synthetic",
                          "config": {"model": "word2vec SGNS 1M-vocab d=300", "vocab": V, "embed_dim": d, "negative": neg,
                                     "global_batch": world * B, "parallelism": f"hash-partitioned table, NCCL all_to_all x{world}"},
                          "loss_last": float(loss) / (B * (neg + 1))}))
    if world > 1:
        dist.destroy_process_group()
    return 0
