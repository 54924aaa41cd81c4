"""Prometheus exporter for a running parameter-manager node (the reference prints its counters at shutdown;
a production deployment wants them scraped while the job runs).

    from adapm_b200.utils.metrics import start_metrics_server
    start_metrics_server(server, workers=[kv], port=9400 + server.my_rank())

Every scrape reads the node's counters (`Server.counters()`: relocations, replica set-ups / drops, refreshes, deltas,
sync rounds, local / remote pulls and pushes, deferred intents, protocol errors) and the workers' locality counters
(`Worker.locality()`); nothing is sampled between scrapes and nothing runs on the training path. Needs the
`prometheus_client` package (soft dependency: importing this module without it raises ImportError)."""
from __future__ import annotations

from typing import Iterable, Optional

from prometheus_client import CollectorRegistry, start_http_server
from prometheus_client.core import CounterMetricFamily, GaugeMetricFamily

_HELP = {
    "relocations": "keys whose ownership moved to this rank",
    "replica_setups": "replicas created on this rank",
    "replica_drops": "replicas dropped on this rank",
    "refreshes": "replica rows refreshed from their owners",
    "deltas_shipped": "replica deltas shipped to the owners",
    "sync_rounds": "completed synchronisation rounds",
    "intents_registered": "intent records registered by the sync round",
    "intents_deferred": "intent records deferred to a later round (pool full / slot being recycled)",
    "pull_local": "rows pulled from local memory",
    "pull_remote": "rows pulled from a peer",
    "push_local": "rows pushed into local memory",
    "push_remote": "rows pushed to a peer",
    "alloc_fail": "slot allocations that found the pool empty",
    "protocol_errors": "protocol invariants violated (must stay 0)",
}


class NodeCollector:
    """prometheus_client collector over one `adapm_b200.Server` (and optionally its workers)."""

    def __init__(self, server, workers: Optional[Iterable] = None, prefix: str = "adapm"):
        self.server, self.workers, self.prefix = server, list(workers or []), prefix

    def collect(self):
        rank = str(self.server.my_rank())
        for name, value in sorted(self.server.counters().items()):
            m = CounterMetricFamily(f"{self.prefix}_{name}", _HELP.get(name, name), labels=["rank"])
            m.add_metric([rank], float(value))
            yield m
        g = GaugeMetricFamily(f"{self.prefix}_world_size", "ranks of the job", labels=["rank"])
        g.add_metric([rank], float(self.server.num_servers()))
        yield g
        for w in self.workers:
            loc = w.locality()
            for name, value in sorted(loc.items()):
                m = CounterMetricFamily(f"{self.prefix}_worker_{name}", f"worker-side {name.replace('_', ' ')}",
                                        labels=["rank", "worker"])
                m.add_metric([rank, str(getattr(w, "customer_id", 0))], float(value))
                yield m


def start_metrics_server(server, workers: Optional[Iterable] = None, port: int = 0, addr: str = "127.0.0.1"):
    """Starts the HTTP endpoint (`/metrics`) in a daemon thread. Returns ``(registry, port)``; ``port=0`` picks a free one."""
    import socket

    if port == 0:
        with socket.socket() as sk:
            sk.bind((addr, 0))
            port = sk.getsockname()[1]
    registry = CollectorRegistry()
    registry.register(NodeCollector(server, workers))
    start_http_server(port, addr=addr, registry=registry)
    return registry, port
