"""Parallelism strategies of the framework and where they live.

* asynchronous data parallelism  - every rank trains its share of the data and applies additive (Hogwild-style) updates
                                   through the parameter manager: ``adapm_b200.Worker`` / the fused ops.
* key-space sharding             - a key's home is ``key % world``; :func:`home_rank`, :func:`home_keys`.
* relocation / replication       - decided per key and per sync round by the owner (``csrc/adapm/protocol.h``,
                                   ``sync_engine.cc``); steered by ``Worker.intent`` and ``sys.techniques``.
* schedules -> intents            - :mod:`adapm_b200.parallel.schedules`: :func:`wor_block_schedule` (DSGD: ranks own row
                                   blocks, column blocks rotate), :func:`column_intent_plan` (column-wise: one ranged
                                   intent per column), :class:`LookaheadIntents` (batched apps: signal batch s + k).
* data partitioning              - :func:`partition_rows` (rank r gets rows r, r + world, ...: the apps' rule).
* comparison arm                 - :mod:`adapm_b200.parallel.nccl_baseline` (stock PyTorch + NCCL all_to_all only).
"""
from __future__ import annotations

import torch

from .schedules import LookaheadIntents, column_intent_plan, wor_block_schedule  # noqa: F401


def home_rank(keys: torch.Tensor, world: int) -> torch.Tensor:
    """Rank that holds a key initially (and to which nothing ever has to be forwarded: the directory is replicated)."""
    return keys % world


def home_keys(num_keys: int, rank: int, world: int) -> torch.Tensor:
    """The keys whose home is ``rank`` (the ones a rank initialises in ``init_model``)."""
    return torch.arange(rank, num_keys, world, dtype=torch.int64)


def partition_rows(n: int, rank: int, world: int) -> torch.Tensor:
    """Indices of the data items of ``rank`` under the apps' striping rule (item i belongs to rank i % world)."""
    return torch.arange(rank, n, world, dtype=torch.int64)
