"""All-reduce through a parameter-server key (reference include/utils.h:163-197 ``ps_allreduce``):
pull the key, push its negation (reset to zero), barrier, push the local contribution, barrier, pull."""
from __future__ import annotations

import torch


def ps_allreduce(kv, key: int, local: torch.Tensor) -> torch.Tensor:
    n = kv.get_key_size(key)
    assert local.numel() <= n, "all-reduce payload does not fit the key"
    dt = kv.server.dtype
    k = torch.tensor([key])
    buf = torch.zeros(n, dtype=dt)
    if kv.server.my_rank() == 0 and kv._impl.id() == 0:
        kv.wait(kv.pull(k, buf))
        kv.wait(kv.push(k, -buf))
    kv.waitall()
    kv.wait_sync()
    kv.barrier()
    contrib = torch.zeros(n, dtype=dt)
    contrib[: local.numel()] = local.to(dt)
    kv.wait(kv.push(k, contrib))
    kv.waitall()
    kv.wait_sync()
    kv.barrier()
    kv.wait_sync()
    out = torch.zeros(n, dtype=dt)
    kv.wait(kv.pull(k, out))
    return out[: local.numel()]
