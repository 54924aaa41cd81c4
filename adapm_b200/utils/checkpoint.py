"""Generic checkpoint / restore of a whole parameter store (any application, both backends).

The reference only has application-level checkpoints (word2vec binary vectors, KGE ``export/checkpoint``, MF
``W.mma/H.mma`` - all provided by the models here); this adds the system-level one: every rank writes the rows it
currently *owns* (after a ``WaitSync``, so replicas are folded in), and a restore ``Set``s them back, whatever the
world size or key placement of the restoring job is.

File ``<prefix>.rank<r>.adapm`` (little endian): magic ``ADAPMCK1``, ``int64 num_keys``, ``int64 n``, ``int32 val_bytes``,
``n`` x ``int64`` key, ``n`` x ``int32`` len, values (dtype of the store) concatenated in key order.
"""
from __future__ import annotations

import glob
import struct

import numpy as np
import torch

MAGIC = b"ADAPMCK1"


def owned_keys(server, chunk: int = 1 << 20) -> torch.Tensor:
    """Keys whose current owner is this rank (ascending)."""
    me, nk = server.my_rank(), server.num_keys()
    out = []
    for a in range(0, nk, chunk):
        k = torch.arange(a, min(nk, a + chunk), dtype=torch.int64)
        st = torch.empty(k.numel(), dtype=torch.uint8)
        ow = torch.empty(k.numel(), dtype=torch.uint8)
        server._impl.peek_into(k.data_ptr(), k.numel(), st.data_ptr(), ow.data_ptr())
        out.append(k[ow == me])
    return torch.cat(out) if out else torch.empty(0, dtype=torch.int64)


def save_store(worker, prefix: str, chunk: int = 1 << 16, attempts: int = 3) -> int:
    """Collective over the RANKS: exactly one worker per rank calls it (worker 0 by convention; with
    ``num_threads > 1`` the other workers of the rank must be idle - e.g. parked at their own barrier - while the
    store is saved). Uses the node barrier, so it works for any number of workers per rank.

    The sync thread keeps running during the save. After the quiescence idiom no relocation is pending, but to be
    safe against a late one (an intent registered long ago that only now falls into the action window) the ownership
    of every chunk is re-checked after its rows were pulled, and the ranks agree that exactly ``num_keys`` keys
    were written; otherwise the save is repeated. Returns the number of keys this rank wrote."""
    server = worker.server
    np_dtype = {torch.float32: np.float32, torch.float64: np.float64, torch.int64: np.int64}[server.dtype]
    fn = f"{prefix}.rank{server.my_rank()}.adapm"
    for attempt in range(attempts):
        worker.waitall()
        worker.wait_sync()
        server.barrier()
        worker.wait_sync()
        server.barrier()                       # quiescent: no relocation in flight, replicas synchronised
        keys = owned_keys(server)
        lens = (torch.full((keys.numel(),), server._uniform_len, dtype=torch.int32) if server._uniform_len is not None
                else server._lens_t[keys].to(torch.int32))
        moved = 0
        with open(fn, "wb") as f:
            f.write(MAGIC + struct.pack("<qqi", server.num_keys(), keys.numel(), np.dtype(np_dtype).itemsize))
            f.write(keys.numpy().tobytes())
            f.write(lens.numpy().tobytes())
            me = server.my_rank()
            for a in range(0, keys.numel(), chunk):
                k = keys[a:a + chunk]
                vals = torch.empty(int(lens[a:a + chunk].sum()), dtype=server.dtype)
                worker.wait(worker.pull(k, vals))
                f.write(vals.numpy().tobytes())
                st = torch.empty(k.numel(), dtype=torch.uint8)
                ow = torch.empty(k.numel(), dtype=torch.uint8)
                server._impl.peek_into(k.data_ptr(), k.numel(), st.data_ptr(), ow.data_ptr())
                moved += int((ow != me).sum())     # the key left this rank while it was being written
        total, moved_all = server.allreduce_sum([float(keys.numel()), float(moved)])
        if int(total) == server.num_keys() and int(moved_all) == 0:
            server.barrier()
            return keys.numel()
    raise RuntimeError(f"save_store: the placement kept changing during {attempts} attempts "
                       f"(keys written {int(total)} of {server.num_keys()}, moved {int(moved_all)}); quiesce the workers first")


def load_store(worker, prefix: str, chunk: int = 1 << 16) -> int:
    """Collective. The files of the saving job (any world size) are distributed round-robin over the ranks of this
    job; rows are written with ``Set``. Returns the number of keys this rank restored."""
    server = worker.server
    files = sorted(glob.glob(f"{prefix}.rank*.adapm"))
    if not files:
        raise FileNotFoundError(f"no checkpoint files match {prefix}.rank*.adapm")
    np_dtype = {torch.float32: np.float32, torch.float64: np.float64, torch.int64: np.int64}[server.dtype]
    restored = 0
    worker.begin_setup()
    for i, fn in enumerate(files):
        if i % server.num_servers() != server.my_rank():
            continue
        with open(fn, "rb") as f:
            if f.read(8) != MAGIC:
                raise ValueError(f"{fn}: not an adapm_b200 store checkpoint")
            nk, n, vb = struct.unpack("<qqi", f.read(20))
            if nk != server.num_keys() or vb != np.dtype(np_dtype).itemsize:
                raise ValueError(f"{fn}: checkpoint of a different store ({nk} keys, {vb}-byte values)")
            keys = torch.from_numpy(np.frombuffer(f.read(8 * n), dtype=np.int64).copy())
            lens = torch.from_numpy(np.frombuffer(f.read(4 * n), dtype=np.int32).copy()).to(torch.int64)
            for a in range(0, n, chunk):
                cnt = int(lens[a:a + chunk].sum())
                vals = torch.from_numpy(np.frombuffer(f.read(cnt * vb), dtype=np_dtype).copy())
                worker.wait(worker.set(keys[a:a + chunk], vals))
            restored += n
    worker.waitall()
    worker.end_setup()
    return restored
