"""Application models on the CPU backend (reference semantics through the public Pull/Push API):
word2vec SGNS, KGE ComplEx/RESCAL (+ filtered ranking eval, checkpoints), MF (DSGD/columnwise/plain)."""
import os

import numpy as np
import pytest
import torch

from harness import run_cluster


def _w2v_worker(kv, server, wid):
    from adapm_b200.models.word2vec import SyntheticPairs, Word2Vec, Word2VecConfig, zipf_counts

    cfg = Word2VecConfig(vocab_size=300, embed_dim=16, negative=3, batch_pairs=256, read_ahead=2, sampling_scheme="naive")
    counts = zipf_counts(cfg.vocab_size)
    model = Word2Vec(server, kv, cfg, counts)
    model.init_model()
    data = SyntheticPairs(cfg, counts, server.my_rank())
    losses = []
    for s in range(12):
        model.signal_intent(data.batch((s + cfg.read_ahead) % 3), kv.current_clock() + cfg.read_ahead)
        losses.append(float(model.step(data.batch(s % 3))))
        kv.advance_clock()
    kv.barrier()
    if wid == 0:
        model.write_checkpoint(os.path.join(server._tmp, "vectors.bin"))
    kv.finalize()
    return losses


@pytest.mark.parametrize("world", [1, 2])
def test_word2vec_cpu(tmp_path, world):
    def setup(server):
        server._tmp = str(tmp_path)

    res = run_cluster(_w2v_worker, world=world, workers=1, mode="threads", setup_fn=setup, value_lengths=32, num_keys=600)
    for r in res.values():
        losses = r[0]
        assert losses[-1] < losses[0]
    hdr = open(tmp_path / "vectors.bin", "rb").readline()
    assert hdr == b"300 16\n"
    # header + per word: "w<i> " + 16 float32 + "\n"
    assert os.path.getsize(tmp_path / "vectors.bin") > 300 * (16 * 4 + 3)


KGE_EVAL_KEYS = ("mrr_s", "mrr_r", "mrr_o", "mrr_s_raw", "mrr_o_raw", "mr_s", "mr_r", "mr_o", "hits01_s", "hits03_r", "hits10_o")


def _kge_worker(kv, server, wid):
    from adapm_b200.models.kge import KGE, KGEConfig, synthetic_triples

    cfg = server._cfg
    model = KGE(server, kv, cfg)
    model.init_model()
    tr = synthetic_triples(cfg, 600, seed=3)
    mine = tr[server.my_rank()::server.num_servers()]
    losses = []
    for ep in range(25):
        tot = 0.0
        for s in range(0, mine.shape[0], cfg.batch_triples):
            b = mine[s:s + cfg.batch_triples]
            model.signal_intent(b, kv.current_clock())
            tot += float(model.step(b))
            kv.advance_clock()
        losses.append(tot)
    # propagation idiom: every replica delta has reached its owner before anybody evaluates (otherwise the local and the
    # distributed evaluation below may see models that differ by the deltas still in flight)
    kv.barrier(); kv.wait_sync(); kv.barrier(); kv.wait_sync(); kv.barrier()
    out = {"losses": losses}
    if wid == 0:
        out["eval"] = model.evaluate(tr[:100], tr)
        model.save(os.path.join(server._tmp, "m."), 6, write_checkpoint=True)
    kv.barrier()
    # the reference's distributed evaluation: every rank ranks its share, the 19 metric sums meet in the eval_key row
    out["eval_dist"] = model.evaluate_distributed(tr[:100], tr)
    kv.finalize()
    return out


@pytest.mark.parametrize("algo,world", [("ComplEx", 1), ("ComplEx", 2), ("RESCAL", 1)])
def test_kge_cpu(tmp_path, algo, world):
    from adapm_b200.models.kge import KGEConfig

    cfg = KGEConfig(num_entities=60, num_relations=7, embed_dim=8, algorithm=algo, neg_ratio=2, batch_triples=100,
                    sampling_scheme="naive", eta=0.2)

    def setup(server):
        server._cfg, server._tmp = cfg, str(tmp_path)

    res = run_cluster(_kge_worker, world=world, workers=1, mode="threads", setup_fn=setup,
                      value_lengths=cfg.value_lengths(), num_keys=cfg.num_keys)
    for r in res.values():
        assert r[0]["losses"][-1] < r[0]["losses"][0]
    ev = res[0][0]["eval"]
    assert 0 < ev["mrr"] <= 1 and ev["mrr"] >= ev["mrr_raw"] - 1e-9 and ev["n"] == 100
    # chance level is ~0.08 for 60 entities; two asynchronous ranks (Hogwild) land between 0.18 and 0.36 after 6 epochs
    assert ev["mrr"] > (0.2 if world == 1 else 0.13), ev
    # subject, relation and object are ranked (kge.cc:742-757); the distributed aggregation reproduces the local numbers
    for k in ("mrr_s", "mrr_r", "mrr_o", "mr_r", "hits10_r", "mrr_s_raw", "mrr_o_raw"):
        assert k in ev, k
    assert 0 < ev["mrr_r"] <= 1 and ev["mr_r"] >= 1
    evd = res[0][0]["eval_dist"]
    for k in KGE_EVAL_KEYS:
        assert abs(evd[k] - ev[k]) <= 1e-4 * max(1.0, abs(ev[k])), (k, evd[k], ev[k])
    for r in range(1, world):
        assert res[r][0]["eval_dist"] == {}
    e = np.fromfile(tmp_path / "m.export.epoch.6.entities.bin", dtype=np.float32)
    assert e.size == cfg.num_entities * cfg.embed_dim
    a = np.fromfile(tmp_path / "m.checkpoint.epoch.6.relations.adagrad.bin", dtype=np.float64)
    assert a.size == cfg.num_relations * cfg.relation_len // 2 and (a > 0).all()


def _mf_worker(kv, server, wid):
    from adapm_b200.models.mf import MatrixFactorization, SparseMatrix

    cfg = server._cfg
    data = SparseMatrix.synthetic(cfg.num_rows, cfg.num_cols, 6000, 4, server.num_servers(), server.my_rank(), seed=5)
    model = MatrixFactorization(server, kv, cfg, data)
    model.init_model()
    kv.barrier()
    losses, prev = [], None
    for ep in range(8):
        l = model.run_epoch(ep)
        model.bold_driver(l, prev)
        prev = l
        losses.append(l)
        kv.barrier()
    if wid == 0:
        model.write_factors(os.path.join(server._tmp, ""))
    kv.barrier()
    kv.finalize()
    return losses


@pytest.mark.parametrize("algo,world", [("dsgd", 2), ("columnwise", 2), ("plain", 1)])
def test_mf_cpu(tmp_path, algo, world):
    from adapm_b200.models.mf import MFConfig

    cfg = MFConfig(num_rows=80, num_cols=40, rank=8, algorithm=algo, eps=0.02, lam=0.01, batch_nnz=250, read_ahead=1)

    def setup(server):
        server._cfg, server._tmp = cfg, str(tmp_path)

    res = run_cluster(_mf_worker, world=world, workers=1, mode="threads", setup_fn=setup, value_lengths=2 * cfg.rank,
                      num_keys=cfg.num_keys(world))
    for r in res.values():
        assert min(r[0][-3:]) < 0.9 * r[0][0], r[0]
    lines = open(tmp_path / "W.mma").read().splitlines()
    assert lines[0].startswith("%%MatrixMarket matrix array") and lines[1] == "80 8" and len(lines) == 2 + 80 * 8


def test_sgns_shared_reference_matches_autograd():
    """The fp32 reference of the shared-negative SGNS step (the yardstick of the tensor-core kernels' GPU test) against
    autograd of the same objective: accumulated g^2 of every row that is touched exactly once."""
    import collections

    import torch

    from adapm_b200.ops import sgns_shared_reference_step

    torch.manual_seed(0)
    d, B, Nn, n = 8, 6, 4, 40
    t = torch.randn(n, 2 * d) * 0.5
    t[:, d:] = torch.rand(n, d) + 0.1
    c = torch.randint(0, n, (B,)); x = torch.randint(0, n, (B,)); ng = torch.randperm(n)[:Nn]
    x[0] = ng[1]
    out, loss = sgns_shared_reference_step(t, c, x, ng, d, 0.1)
    E = t[:, :d].clone().requires_grad_(True)
    fp = (E[c] * E[x]).sum(-1)
    S = E[c] @ E[ng].t()
    mask = x.view(-1, 1) != ng.view(1, -1)
    L = torch.nn.functional.softplus(-fp).sum() + (torch.nn.functional.softplus(S) * mask).sum()
    L.backward()
    assert abs(loss.item() - L.item()) < 1e-4
    cnt = collections.Counter(c.tolist() + x.tolist() + ng.tolist())
    once = [k for k, v in cnt.items() if v == 1]
    assert once
    for k in once:
        g = -E.grad[k]
        torch.testing.assert_close(out[k, d:] - t[k, d:], g * g, atol=1e-5, rtol=1e-4)
        torch.testing.assert_close(out[k, :d] - t[k, :d], 0.1 * g * torch.rsqrt(t[k, d:] + g * g), atol=1e-5, rtol=1e-4)


def _w2v_shared_worker(kv, server, wid):
    from adapm_b200.models.word2vec import SyntheticPairs, Word2Vec, Word2VecConfig, zipf_counts

    V = 2000
    cfg = Word2VecConfig(vocab_size=V, embed_dim=16, negative=5, batch_pairs=256, shared_negatives=32, read_ahead=2)
    cnt = zipf_counts(V, 1.0)
    model = Word2Vec(server, kv, cfg, cnt)
    model.init_model()
    data = SyntheticPairs(cfg, cnt, server.my_rank(), seed=1)
    losses = []
    for it in range(80):
        b = data.batch(it % 4)
        model.signal_intent(data.batch((it + 2) % 4), kv.current_clock() + 2)
        losses.append(float(model.step(b)) / (cfg.batch_pairs * (1 + cfg.shared_negatives)))
        kv.advance_clock()
    kv.barrier()
    kv.finalize()
    return losses


@pytest.mark.parametrize("world", [1, 2])
def test_word2vec_shared_negatives_through_pull_push(world):
    """--shared_negatives on the CPU backend: one set of negatives per batch, the same rule as the tensor-core variant
    (ops.SgnsSharedStep) computed on pulled rows and pushed back as additive updates; the loss goes down, also with
    relocation / replication between two ranks."""
    res = run_cluster(_w2v_shared_worker, world=world, workers=1, mode="threads", value_lengths=32, num_keys=4000)
    for r in res.values():
        ls = r[0]
        assert ls[-1] < 0.97 * ls[0], (ls[0], ls[-1])
        assert r["counters"]["protocol_errors"] == 0


def _mf_ranged_worker(kv, server, wid):
    from adapm_b200.models.mf import MatrixFactorization, SparseMatrix

    cfg = server._cfg
    data = SparseMatrix.synthetic(cfg.num_rows, cfg.num_cols, 4000, 4, server.num_servers(), server.my_rank(), seed=7)
    model = MatrixFactorization(server, kv, cfg, data)
    model.init_model()
    kv.barrier()
    calls = []
    orig = kv.intent
    kv.intent = lambda keys, start, end=0: (calls.append((keys.numel(), int(end) - int(start) if end else 1)), orig(keys, start, end))[1]
    l0 = model.run_epoch(0)
    n_batches = (data.i.shape[0] + cfg.batch_nnz - 1) // cfg.batch_nnz
    kv.intent = orig
    for ep in range(1, 6):
        l = model.run_epoch(ep)
    kv.barrier()
    loc = kv.locality()
    kv.finalize()
    return {"l0": l0, "l": l, "calls": calls, "n_batches": n_batches, "loc": loc}


def test_mf_columnwise_ranged_intents(tmp_path):
    """Column-wise MF signals ONE intent per column for the clocks its data points span (reference mf.cc:477-482): with 6
    columns spread over ~40 batches every column is signalled once per epoch, with a multi-clock window."""
    from adapm_b200.models.mf import MFConfig

    cfg = MFConfig(num_rows=200, num_cols=6, rank=4, algorithm="columnwise", eps=0.02, lam=0.01, batch_nnz=50, read_ahead=2)

    def setup(server):
        server._cfg, server._tmp = cfg, str(tmp_path)

    res = run_cluster(_mf_ranged_worker, world=2, workers=1, mode="threads", setup_fn=setup, value_lengths=2 * cfg.rank,
                      num_keys=cfg.num_keys(2))
    for r in res.values():
        out = r[0]
        keys_signalled = sum(n for n, _ in out["calls"])
        assert keys_signalled <= cfg.num_cols, out["calls"]              # once per column, not once per batch
        assert out["n_batches"] > 3 * cfg.num_cols
        assert max(d for _, d in out["calls"]) >= 3, out["calls"]           # windows cover the batches a column spans
        assert out["l"] < out["l0"]
        assert r["counters"]["protocol_errors"] == 0
