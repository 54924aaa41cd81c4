"""Schedules that decide WHICH rank touches WHICH keys WHEN - and therefore what to tell the parameter manager ahead
of time. The applications (models/mf.py, models/word2vec.py, models/kge.py, bench.py) use them to turn their access
pattern into `Worker.intent` calls; nothing here talks to a GPU.

* :func:`wor_block_schedule`   DSGD stratified block schedule (reference apps/mf/data.h:182-210): in sub-epoch `se` rank `w`
                               works on column block `schedule[se, w]`; a random Latin square, so no two ranks share a
                               block in a sub-epoch and every rank sees every block once per epoch.
* :func:`column_intent_plan`   column-wise schedule (reference mf.cc:477-482): the data points are sorted by column, a
                               column occupies a contiguous run of batches and is signalled ONCE for the clocks it spans.
* :class:`LookaheadIntents`    the generic pattern of the batched applications: while batch `s` runs, signal the distinct
                               keys of batch `s + read_ahead` for the clock at which that batch will run.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np
import torch


def wor_block_schedule(world: int, epoch: int, seed: int = 0) -> np.ndarray:
    """schedule[subepoch, worker] = column block (a random Latin square: WOR block schedule)."""
    rng = np.random.default_rng(seed * 7919 + epoch)
    perm, shift = rng.permutation(world), rng.permutation(world)
    return np.array([[perm[(w + shift[se]) % world] for w in range(world)] for se in range(world)])


def column_intent_plan(sorted_cols: np.ndarray, batch: int):
    """``sorted_cols``: the column id of every data point in processing order (sorted by column). Returns
    ``(cols, duration, ptr)``: the distinct columns in order, the number of batches each spans, and ``ptr`` such that
    the columns whose FIRST batch is ``b`` are ``cols[ptr[b]:ptr[b + 1]]``."""
    n = int(sorted_cols.shape[0])
    n_batches = (n + batch - 1) // batch
    cols, first_idx = np.unique(sorted_cols, return_index=True)
    last_idx = np.r_[first_idx[1:], n] - 1
    first_b, last_b = first_idx // batch, last_idx // batch
    ptr = np.searchsorted(first_b, np.arange(n_batches + 1))
    return cols, last_b - first_b + 1, ptr


class LookaheadIntents:
    """``signal(s)`` calls ``worker.intent(distinct keys of batch s + read_ahead, clock + read_ahead)`` - once per batch,
    nothing for batches beyond the end. ``keys_of(b)`` returns the (not necessarily distinct) keys of batch ``b`` as an
    int64 tensor; a batch object may carry its distinct keys as ``unique_keys`` (a data loader that de-duplicates on
    its own thread, like the native corpus loader does)."""

    def __init__(self, worker, num_batches: int, read_ahead: int, keys_of: Callable[[int], torch.Tensor],
                 duration: int = 1):
        self.worker, self.n, self.ra, self.keys_of, self.duration = worker, int(num_batches), int(read_ahead), keys_of, int(duration)
        self.keys_signalled = 0

    def prime(self) -> None:
        """The first ``read_ahead`` batches have no earlier batch to signal them."""
        c = self.worker.current_clock()
        for b in range(min(self.ra, self.n)):
            self._emit(b, c + b)

    def signal(self, s: int) -> None:
        b = s + self.ra
        if b < self.n:
            self._emit(b, self.worker.current_clock() + self.ra)

    def _emit(self, b: int, start: int) -> None:
        k = self.keys_of(b)
        u: Optional[torch.Tensor] = getattr(k, "unique_keys", None)
        if u is None:
            u = torch.unique(k.reshape(-1))
        self.worker.intent(u, start, start + self.duration)
        self.keys_signalled += int(u.numel())
