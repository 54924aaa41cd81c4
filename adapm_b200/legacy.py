"""Legacy PS-Lite application API: ``SimpleApp`` (int + byte-string RPC) and the range-sliced
``KVWorker`` / ``KVServer`` pair.

Capability parity with ``include/ps/simple_app.h:32-184`` and ``include/ps/kv_app.h:33-697`` of the reference
(``KVPairs``, ``KVMeta``, ``DefaultSlicer`` (:477-534), ``KVServerDefaultHandle`` (:388-410), per-request
callbacks (:250-264)). The reference keeps these classes only as the base of its ColoKV classes (its
``KVWorker::Send`` never sends); here they are a small, working host-side service on top of the shared-memory
mailboxes of ``csrc/adapm/rpc.{h,cc}``: requests are fragmented into the receiver's ring in the control block,
the receiver's router thread reassembles them and runs the handle, the response travels back the same way.
Use :class:`adapm_b200.Worker` for anything performance-relevant: its traffic never touches the host.
"""
from __future__ import annotations

import struct
import threading
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from . import _C

kServerGroup = _C.kServerGroup
kWorkerGroup = _C.kWorkerGroup
kAllNodes = _C.kAllNodes
SimpleData = _C.SimpleData


class SimpleApp:
    """``Request(head, body, recv_id) -> ts``, ``Wait(ts)``, ``Response(req, body)``, request/response handles.
    Handles run on the rank's router thread and receive ``(SimpleData, app)``."""

    def __init__(self, app_id: int, customer_id: int, server, serves_requests: bool = True) -> None:
        self.server = server
        self._impl = _C.SimpleApp(int(app_id), int(customer_id), server._impl, bool(serves_requests))

    def request(self, head: int, body: bytes | str, recv_id: int) -> int:
        if isinstance(body, str):
            body = body.encode()
        return self._impl.request(int(head), bytes(body), int(recv_id))

    def wait(self, ts: int) -> None:
        self._impl.wait(int(ts))

    def num_response(self, ts: int) -> int:
        return self._impl.num_response(int(ts))

    def response(self, req, body: bytes | str = b"") -> None:
        if isinstance(body, str):
            body = body.encode()
        self._impl.response(req, bytes(body))

    def set_request_handle(self, fn: Callable) -> None:
        self._impl.set_request_handle(lambda d, _app: fn(d, self))

    def set_response_handle(self, fn: Callable) -> None:
        self._impl.set_response_handle(lambda d, _app: fn(d, self))


# ------------------------------------------------------------------------------------------ KV
@dataclass
class KVPairs:
    """keys (int64, ascending), vals (float32), lens (int32, optional: per-key value lengths)."""
    keys: np.ndarray = field(default_factory=lambda: np.empty(0, np.int64))
    vals: np.ndarray = field(default_factory=lambda: np.empty(0, np.float32))
    lens: np.ndarray = field(default_factory=lambda: np.empty(0, np.int32))


@dataclass
class KVMeta:
    cmd: int
    push: bool
    sender: int
    timestamp: int
    customer_id: int
    _raw: object = None
    _kv_ts: int = 0


_HDR = struct.Struct("<iiqqqq")  # cmd, push, kv_ts, n_keys, n_vals, n_lens


def _pack(cmd: int, push: bool, kv_ts: int, kv: KVPairs) -> bytes:
    k = np.ascontiguousarray(kv.keys, dtype=np.int64)
    v = np.ascontiguousarray(kv.vals, dtype=np.float32)
    ln = np.ascontiguousarray(kv.lens, dtype=np.int32)
    return _HDR.pack(cmd, 1 if push else 0, kv_ts, k.size, v.size, ln.size) + k.tobytes() + v.tobytes() + ln.tobytes()


def _unpack(body: bytes) -> Tuple[int, bool, int, KVPairs]:
    cmd, push, kv_ts, nk, nv, nl = _HDR.unpack_from(body, 0)
    o = _HDR.size
    k = np.frombuffer(body, np.int64, nk, o).copy(); o += 8 * nk
    v = np.frombuffer(body, np.float32, nv, o).copy(); o += 4 * nv
    ln = np.frombuffer(body, np.int32, nl, o).copy()
    return cmd, bool(push), kv_ts, KVPairs(k, v, ln)


def server_key_ranges(num_keys: int, num_servers: int) -> List[Tuple[int, int]]:
    """Equal-width static key ranges, one per server (reference Postoffice::GetServerKeyRanges)."""
    return [(num_keys * i // num_servers, num_keys * (i + 1) // num_servers) for i in range(num_servers)]


def default_slicer(send: KVPairs, ranges: List[Tuple[int, int]]) -> List[Optional[KVPairs]]:
    """Splits an ascending key list at the server range boundaries (reference DefaultSlicer, kv_app.h:477-534).
    Returns one KVPairs (or None when the slice is empty) per range."""
    keys = np.asarray(send.keys, dtype=np.int64)
    n = keys.size
    if n > 1 and not np.all(keys[1:] >= keys[:-1]):
        raise ValueError("KVWorker: keys must be sorted ascending")
    pos = [int(np.searchsorted(keys, ranges[0][0], side="left"))]
    pos += [int(np.searchsorted(keys, hi, side="left")) for _, hi in ranges]
    has_lens = send.lens.size > 0
    if has_lens:
        if send.lens.size != n:
            raise ValueError("KVWorker: lens must have one entry per key")
        voff = np.concatenate([[0], np.cumsum(send.lens, dtype=np.int64)])
    else:
        k = send.vals.size // n if n else 0
        if n and send.vals.size and send.vals.size % n:
            raise ValueError("KVWorker: vals size must be a multiple of the number of keys")
        voff = np.arange(n + 1, dtype=np.int64) * k
    out: List[Optional[KVPairs]] = []
    for i in range(len(ranges)):
        a, b = pos[i], pos[i + 1]
        if a == b:
            out.append(None)
            continue
        vals = send.vals[voff[a]:voff[b]] if send.vals.size else np.empty(0, np.float32)
        out.append(KVPairs(keys[a:b], vals, send.lens[a:b] if has_lens else np.empty(0, np.int32)))
    return out


class KVServer:
    """Holds one static key range; ``set_request_handle(fn(meta, kvpairs, server))``; ``response(meta, kvpairs)``."""

    def __init__(self, app_id: int, server) -> None:
        self.app = SimpleApp(app_id, 0, server)
        self._handle: Optional[Callable] = None
        self.app.set_request_handle(self._on_request)

    def set_request_handle(self, fn: Callable) -> None:
        self._handle = fn

    def _on_request(self, d, _app) -> None:
        cmd, push, kv_ts, kv = _unpack(d.body)
        meta = KVMeta(cmd, push, d.sender, d.timestamp, d.customer_id, _raw=d, _kv_ts=kv_ts)
        if self._handle is None:
            raise RuntimeError("KVServer: no request handle set")
        self._handle(meta, kv, self)

    def response(self, meta: KVMeta, res: Optional[KVPairs] = None) -> None:
        self.app.response(meta._raw, _pack(meta.cmd, meta.push, meta._kv_ts, res or KVPairs()))


class KVServerDefaultHandle:
    """``store[key] += val`` on push, ``store[key]`` on pull (reference kv_app.h:388-410); values of any length."""

    def __init__(self) -> None:
        self.store: Dict[int, np.ndarray] = {}

    def __call__(self, meta: KVMeta, req: KVPairs, server: KVServer) -> None:
        n = req.keys.size
        res = KVPairs()
        if meta.push:
            if req.lens.size:
                off = np.concatenate([[0], np.cumsum(req.lens, dtype=np.int64)])
            else:
                off = np.arange(n + 1, dtype=np.int64) * (req.vals.size // max(n, 1))
            for i, k in enumerate(req.keys.tolist()):
                v = req.vals[off[i]:off[i + 1]]
                cur = self.store.get(k)
                self.store[k] = v.copy() if cur is None else cur + v
        else:
            vals = [self.store.get(k, np.zeros(0, np.float32)) for k in req.keys.tolist()]
            res.keys = req.keys
            res.lens = np.array([v.size for v in vals], np.int32)
            res.vals = np.concatenate(vals) if vals else np.empty(0, np.float32)
        server.response(meta, res)


class KVWorker:
    """``push(keys, vals, lens=None, cmd=0, callback=None) -> ts``, ``pull(keys, vals_out, lens_out=None, ...)
    -> ts``, ``wait(ts)``. Keys must be ascending; each server receives the slice of its static key range."""

    def __init__(self, app_id: int, customer_id: int, server, num_keys: Optional[int] = None) -> None:
        self.app = SimpleApp(app_id, customer_id, server, serves_requests=False)
        self.app.set_response_handle(self._on_response)
        self.ranges = server_key_ranges(int(num_keys if num_keys is not None else server.num_keys()),
                                        server.num_servers())
        self._slicer = default_slicer
        self._mu = threading.Lock()
        self._pending: Dict[int, dict] = {}
        self._next = 0

    def set_slicer(self, fn: Callable) -> None:
        self._slicer = fn

    def _send(self, push: bool, kv: KVPairs, cmd: int, callback, out) -> int:
        sliced = self._slicer(kv, self.ranges)
        with self._mu:
            ts = self._next
            self._next += 1
            st = {"expected": sum(1 for s in sliced if s is not None), "recv": [], "app_ts": [], "cb": callback,
                  "out": out, "done": threading.Event()}
            self._pending[ts] = st
        if st["expected"] == 0:
            self._finish(ts, st)
            return ts
        for rank, s in enumerate(sliced):
            if s is not None:
                st["app_ts"].append(self.app.request(cmd, _pack(cmd, push, ts, s), rank))
        return ts

    def push(self, keys, vals, lens=None, cmd: int = 0, callback: Optional[Callable] = None) -> int:
        kv = KVPairs(np.asarray(keys, np.int64), np.asarray(vals, np.float32).reshape(-1),
                     np.asarray(lens, np.int32) if lens is not None else np.empty(0, np.int32))
        return self._send(True, kv, cmd, callback, None)

    def pull(self, keys, vals_out: np.ndarray, lens_out: Optional[np.ndarray] = None, cmd: int = 0,
             callback: Optional[Callable] = None) -> int:
        kv = KVPairs(np.asarray(keys, np.int64))
        return self._send(False, kv, cmd, callback, (vals_out, lens_out))

    # zero-copy variants of the reference (kv_app.h:182-210: SArray arguments instead of std::vector): numpy arrays are
    # passed by reference here anyway, so they are the same calls
    zpush = push
    zpull = pull

    def add_callback(self, ts: int, callback: Callable) -> None:
        """Runs ``callback`` when request ``ts`` completes (reference ``AddCallback``, kv_app.h:250-258); at once if it
        already has."""
        with self._mu:
            st = self._pending.get(ts)
            if st is not None and not st.get("finished"):
                prev = st["cb"]
                st["cb"] = callback if prev is None else (lambda: (prev(), callback()))
                return
        callback()

    def wait(self, ts: int) -> None:
        with self._mu:
            st = self._pending.get(ts)
        if st is None:
            return
        for a in list(st["app_ts"]):
            self.app.wait(a)
        st["done"].wait(300)
        with self._mu:
            self._pending.pop(ts, None)

    def _on_response(self, d, _app) -> None:
        cmd, push, kv_ts, kv = _unpack(d.body)
        with self._mu:
            st = self._pending.get(kv_ts)
            if st is None:
                return
            st["recv"].append(kv)
            complete = len(st["recv"]) == st["expected"]
        if complete:
            self._finish(kv_ts, st)

    def _finish(self, ts: int, st: dict) -> None:
        if st["out"] is not None and st["recv"]:
            vals_out, lens_out = st["out"]
            parts = sorted(st["recv"], key=lambda kv: int(kv.keys[0]) if kv.keys.size else -1)
            vals = np.concatenate([p.vals for p in parts])
            if vals_out.size < vals.size:
                raise ValueError("KVWorker.pull: output buffer too small")
            vals_out.reshape(-1)[:vals.size] = vals
            if lens_out is not None:
                lens = np.concatenate([p.lens for p in parts])
                lens_out.reshape(-1)[:lens.size] = lens
        with self._mu:                      # (add_callback either lands before this point or runs its callback itself)
            cb, st["finished"] = st["cb"], True
        if cb is not None:
            cb()
        st["done"].set()
