"""Read-only serving of a trained store: embedding lookup and nearest neighbours over HTTP.

    python -m adapm_b200.serve --checkpoint /path/prefix --embed_dim 300 [--port 8000] [--backend cpu|cuda]

loads a store checkpoint (`utils/checkpoint.py`: the files of a training job of ANY world size) into a single-rank
server and answers

    GET /health                           -> {"status": "ok", "keys": N, "value_length": L}
    GET /pull?keys=3,17,42                -> {"keys": [...], "values": [[...], ...]}          rows through Worker.pull
    GET /topk?key=42&k=10[&stride=2&offset=0]
                                          -> {"key": 42, "neighbours": [[key, cosine], ...]}   over the first embed_dim
                                             values of the rows key = offset + i * stride (word2vec: stride 2 = syn0)
    GET /metrics                          -> Prometheus text (counters of the node)

The candidate matrix for /topk is pulled once at start-up (normalised, kept on the server's device: HBM on the cuda
backend, where the scoring GEMM runs on the GPU). The reference has no serving path; this is the deployment side of
"train with the parameter manager, serve from its checkpoint"."""
from __future__ import annotations

import argparse
import struct
import sys
from typing import Optional

import torch


def checkpoint_header(prefix: str):
    """(num_keys, value_bytes, uniform_length or None) of a store checkpoint."""
    import glob

    import numpy as np

    from .utils.checkpoint import MAGIC

    files = sorted(glob.glob(f"{prefix}.rank*.adapm"))
    if not files:
        raise FileNotFoundError(f"no checkpoint files match {prefix}.rank*.adapm")
    lens_all = []
    nk = vb = None
    for fn in files:
        with open(fn, "rb") as f:
            if f.read(8) != MAGIC:
                raise ValueError(f"{fn}: not an adapm_b200 store checkpoint")
            nk, n, vb = struct.unpack("<qqi", f.read(20))
            f.seek(8 * n, 1)
            lens_all.append(np.frombuffer(f.read(4 * n), dtype=np.int32))
    lens = np.concatenate(lens_all) if lens_all else np.zeros(0, np.int32)
    uniform = int(lens[0]) if lens.size and bool((lens == lens[0]).all()) else None
    return int(nk), int(vb), uniform


class EmbeddingService:
    def __init__(self, server, worker, embed_dim: int, stride: int = 1, offset: int = 0, max_candidates: int = 5_000_000):
        self.server, self.worker, self.d = server, worker, int(embed_dim)
        self.stride, self.offset = int(stride), int(offset)
        nk = server.num_keys()
        cand = torch.arange(self.offset, nk, self.stride, dtype=torch.int64)[:max_candidates]
        L = server.get_len(int(cand[0]))
        rows = torch.empty(cand.numel() * L, dtype=server.dtype)
        chunk = 1 << 16
        pos = 0
        for a in range(0, cand.numel(), chunk):                  # chunked Pull of the candidate rows
            k = cand[a:a + chunk]
            v = rows[pos:pos + k.numel() * L]
            worker.wait(worker.pull(k, v))
            pos += k.numel() * L
        E = rows.view(-1, L)[:, :self.d].to(torch.float32)
        E = E / E.norm(dim=1, keepdim=True).clamp(min=1e-12)
        self.cand, self.E = cand, E.to(server.device)

    def pull(self, keys):
        k = torch.tensor(list(keys), dtype=torch.int64)
        if k.numel() == 0 or int(k.min()) < 0 or int(k.max()) >= self.server.num_keys():
            raise KeyError("key out of range")
        lens = [self.server.get_len(int(x)) for x in k.tolist()]
        v = torch.empty(sum(lens), dtype=self.server.dtype)
        self.worker.wait(self.worker.pull(k, v))
        out, p = [], 0
        for n in lens:
            out.append(v[p:p + n].tolist())
            p += n
        return out

    def topk(self, key: int, k: int = 10):
        idx = (int(key) - self.offset) // self.stride
        if (int(key) - self.offset) % self.stride or not (0 <= idx < self.cand.numel()):
            raise KeyError("key is not among the candidates (offset / stride)")
        scores = self.E @ self.E[idx]                            # cosine against every candidate (GPU GEMV on cuda)
        scores[idx] = -2.0
        val, pos = torch.topk(scores, min(int(k), self.cand.numel() - 1))
        return [[int(self.cand[p]), float(s)] for p, s in zip(pos.tolist(), val.tolist())]


def make_app(svc: EmbeddingService):
    from fastapi import FastAPI, HTTPException
    from fastapi.responses import PlainTextResponse

    app = FastAPI(title="adapm_b200 embedding service")

    @app.get("/health")
    def health():
        return {"status": "ok", "keys": svc.server.num_keys(), "candidates": int(svc.cand.numel()), "embed_dim": svc.d}

    @app.get("/pull")
    def pull(keys: str):
        try:
            ks = [int(t) for t in keys.split(",") if t.strip()]
            return {"keys": ks, "values": svc.pull(ks)}
        except (KeyError, ValueError) as e:
            raise HTTPException(status_code=400, detail=str(e))

    @app.get("/topk")
    def topk(key: int, k: int = 10):
        try:
            return {"key": key, "neighbours": svc.topk(key, k)}
        except KeyError as e:
            raise HTTPException(status_code=400, detail=str(e))

    @app.get("/metrics", response_class=PlainTextResponse)
    def metrics():
        from prometheus_client import CollectorRegistry, generate_latest

        from .utils.metrics import NodeCollector

        reg = CollectorRegistry()
        reg.register(NodeCollector(svc.server, [svc.worker]))
        return generate_latest(reg).decode()

    return app


def load_service(prefix: str, embed_dim: int, backend: Optional[str] = None, stride: int = 1, offset: int = 0,
                 dtype: Optional[str] = None) -> EmbeddingService:
    from . import Server, Worker, setup
    from .utils.checkpoint import load_store

    nk, vb, uniform = checkpoint_header(prefix)
    if uniform is None:
        raise ValueError("serving needs a store with one value length (per-key lengths are not supported here)")
    dtype = dtype or {4: "float32", 8: "float64"}[vb]
    setup(nk, 1)
    server = Server(uniform, num_keys=nk, num_threads=1, rank=0, world=1, backend=backend, fabric="inproc", dtype=dtype,
                    job="serve")
    kv = Worker(0, server)
    load_store(kv, prefix)
    return EmbeddingService(server, kv, embed_dim, stride, offset)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--checkpoint", required=True, help="prefix of the store checkpoint (<prefix>.rank<r>.adapm)")
    ap.add_argument("--embed_dim", type=int, required=True)
    ap.add_argument("--stride", type=int, default=1)
    ap.add_argument("--offset", type=int, default=0)
    ap.add_argument("--backend", default=None, choices=["cpu", "cuda"])
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=8000)
    args = ap.parse_args(argv)
    import uvicorn

    svc = load_service(args.checkpoint, args.embed_dim, args.backend, args.stride, args.offset)
    uvicorn.run(make_app(svc), host=args.host, port=args.port, log_level="warning")
    return 0


if __name__ == "__main__":
    sys.exit(main())
