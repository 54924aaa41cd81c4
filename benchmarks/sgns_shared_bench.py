"""Device-timed step of the shared-negative SGNS variant on the tensor cores (adapm_b200.ops.SgnsSharedStep,
csrc/cuda/ops_sgns_shared.cu) on one GPU at the headline shape (1M vocab, d = 300, 32768 pairs per step), next to the
reference-faithful fused step (25 private negatives per pair). The two do different work per step - the line reports
pairs/s, sample pairs/s (B * (1 + Nn) vs B * 26) and row updates/s for both - so this is NOT the headline metric."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adapm_b200 as ad  # noqa: E402
from adapm_b200.models.word2vec import Word2Vec, Word2VecConfig, zipf_counts  # noqa: E402
from adapm_b200.ops import SgnsSharedStep, sgns_step  # noqa: E402


def main():
    V, d, B = int(os.environ.get("VOCAB", 1_000_000)), 300, 32768
    cfg = Word2VecConfig(vocab_size=V, embed_dim=d, negative=25, batch_pairs=B)
    server = ad.Server(2 * d, num_keys=2 * V, num_threads=1, rank=0, world=1, backend="cuda", fabric="inproc",
                       job="benchshared", device=0)
    kv = ad.Worker(0, server)
    model = Word2Vec(server, kv, cfg, zipf_counts(V, cfg.zipf_exponent))
    model.init_model()
    dev = server.device
    g = torch.Generator(device=dev).manual_seed(1)
    w = torch.from_numpy(zipf_counts(V, cfg.zipf_exponent) ** 0.75).to(dev)
    K = 20
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    loss = torch.zeros(1, device=dev)
    out = {}
    for Nn in (256, 1024):
        step = SgnsSharedStep(server, kv, B, Nn, d)
        batches = []
        for _ in range(8):
            cw = torch.multinomial(w, B, replacement=True, generator=g)
            xw = torch.multinomial(w, B, replacement=True, generator=g)
            nw = torch.multinomial(w, Nn, replacement=False, generator=g)
            batches.append((2 * cw, 2 * xw + 1, 2 * nw + 1))
        for i in range(3):
            step(*batches[i], 0.025, loss)
        torch.cuda.synchronize()
        a.record()
        for i in range(K):
            step(*batches[i % 8], 0.025, loss)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / K
        out[f"shared_Nn{Nn}"] = {"ms_per_step": ms, "pairs_per_s": B / ms * 1e3, "sample_pairs_per_s": B * (1 + Nn) / ms * 1e3,
                                 "row_updates_per_s": (2 * B + Nn) / ms * 1e3,
                                 "gemm_tflops": 3 * 2.0 * B * Nn * 304 / ms * 1e-9}
    # the fused reference-faithful step on the same store
    cw = torch.multinomial(w, B, replacement=True, generator=g)
    xw = torch.multinomial(w, B, replacement=True, generator=g)
    nw = torch.multinomial(w, B * 25, replacement=True, generator=g).view(B, 25)
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    for i in range(3):
        sgns_step(server, 2 * cw, 2 * xw + 1, (2 * nw + 1).contiguous(), d, 0.025, loss, stats)
    torch.cuda.synchronize()
    a.record()
    for i in range(K):
        sgns_step(server, 2 * cw, 2 * xw + 1, (2 * nw + 1).contiguous(), d, 0.025, loss, stats)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / K
    out["fused_neg25"] = {"ms_per_step": ms, "pairs_per_s": B / ms * 1e3, "sample_pairs_per_s": B * 26 / ms * 1e3,
                          "row_updates_per_s": B * 27 / ms * 1e3}
    print(json.dumps({"bench": "sgns_shared_negatives", "config": f"vocab {V}, d {d}, {B} pairs per step, 1 GPU, fp32 rows",
                      **out}), flush=True)
    kv.finalize(); server.shutdown()


if __name__ == "__main__":
    main()
