"""Shared command-line plumbing of the applications: the reference's system options
(``sys.*``, ``sampling.*``; coloc_kv_server.h:205-222, sync_manager.h:805-814, sampling.h:166-170)
are accepted with their original names and forwarded to the server."""
from __future__ import annotations

import argparse
import threading

SYS_FLAGS = ["sys.zmq_threads", "sys.techniques", "sys.time_intent_actions", "sys.location_caches", "sys.channels",
             "sys.trace.keys", "sys.stats.out", "sys.sync.max_per_sec", "sys.sync.pause", "sys.sync.threshold",
             "sys.timing.initial_estimate", "sys.timing.autotune", "sys.timing.smoothing_factor",
             "sys.timing.buffer_quantile", "sampling.scheme", "sampling.pool_size", "sampling.reuse",
             "sampling.batch_size", "sampling.with_replacement"]


def add_system_options(ap: argparse.ArgumentParser) -> None:
    g = ap.add_argument_group("system options (same names as the reference)")
    for f in SYS_FLAGS:
        g.add_argument("--" + f, dest=f, default=None)
    g.add_argument("--backend", default=None, choices=["cpu", "cuda"])


def system_options(args) -> dict:
    return {f: getattr(args, f) for f in SYS_FLAGS if getattr(args, f, None) is not None}


def strip_dashes(argv):
    return [a for a in argv if a != "--"]


def run_workers(num_threads: int, fn) -> list:
    """fn(customer_id) in num_threads threads (the reference spawns worker threads per node)."""
    out, errs = [None] * num_threads, []

    def w(c):
        try:
            out[c] = fn(c)
        except BaseException as e:  # noqa
            errs.append(e)

    ths = [threading.Thread(target=w, args=(c,)) for c in range(num_threads)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    if errs:
        raise errs[0]
    return out
