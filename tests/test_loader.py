"""Native corpus loader (csrc/adapm/corpus.cc) against the Python reference implementation (utils/text.py)."""
import os
import time

import numpy as np
import torch

from adapm_b200.utils.text import NativeCorpus, Vocabulary, read_sentences

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TEXT = os.path.join(ROOT, "data", "lm", "small.txt")


def test_vocabulary_matches_python_builder():
    py = Vocabulary.build(TEXT, min_count=3)
    nat = NativeCorpus.build(TEXT, min_count=3).vocabulary()
    assert nat.words == py.words
    assert np.array_equal(nat.counts, py.counts)
    assert nat.words[0] == "</s>" and np.all(np.diff(nat.counts[1:]) <= 0)


def test_encoding_matches_python_reader_and_partitions_lines():
    vocab = Vocabulary.build(TEXT, min_count=3)
    for rank, world in ((0, 1), (1, 3)):
        nat = NativeCorpus.from_vocabulary(vocab).encode(TEXT, rank, world)
        want = [s for s in read_sentences(TEXT, vocab, rank, world, 0.0, np.random.default_rng(0))]
        got = [s for s in nat.sentences() if s.size > 1]        # the python reader drops 1-word sentences
        assert len(got) == len(want)
        assert all(np.array_equal(a, b) for a, b in zip(got, want))


def test_pair_stream_pairs_are_window_neighbours_and_cover_the_corpus():
    window, B = 3, 4096
    nat = NativeCorpus.build(TEXT, min_count=1).encode(TEXT)
    sents = nat.sentences()
    # every (context, centre) pair that can occur at distance <= window
    allowed = set()
    for s in sents:
        for i in range(s.size):
            for j in range(max(0, i - window), min(s.size, i + window + 1)):
                if i != j:
                    allowed.add((int(s[j]), int(s[i])))
    n_valid, batches = 0, 0
    seen_centres = set()
    for t in nat.pair_batches(window, 0.0, B, seed=5):
        assert t.shape == (2, B) and t.dtype == torch.int64
        v = t.valid_pairs
        c, x = t[0, :v].numpy(), t[1, :v].numpy()
        assert np.all(c % 2 == 0) and np.all(x % 2 == 1)         # syn0 keys of context words, syn1 keys of centres
        pairs = set(zip((c // 2).tolist(), ((x - 1) // 2).tolist()))
        assert pairs <= allowed
        seen_centres.update(((x - 1) // 2).tolist())
        if v < B:                                                # padded tail repeats real pairs
            assert torch.equal(t[:, v:], t[:, :v].repeat(1, (B + v - 1) // v)[:, :B - v])
        n_valid += v
        batches += 1
    tokens = sum(s.size for s in sents if s.size > 1)
    # window shrink b ~ U[0, window): expected pairs per position = 2 * mean(window - b) minus boundary effects
    assert 0.8 * tokens * (window + 1) * 0.85 <= n_valid <= tokens * (window + 1) * 1.0 + B
    assert len(seen_centres) >= 0.99 * len({int(w) for s in sents if s.size > 1 for w in s})
    # deterministic per (seed, epoch); different epochs differ
    a = next(iter(nat.pair_batches(window, 1e-3, B, seed=5, epoch=1)))
    b = next(iter(nat.pair_batches(window, 1e-3, B, seed=5, epoch=1)))
    c = next(iter(nat.pair_batches(window, 1e-3, B, seed=5, epoch=2)))
    assert torch.equal(a, b) and not torch.equal(a, c)


def test_subsampling_thins_frequent_words():
    nat = NativeCorpus.build(TEXT, min_count=1).encode(TEXT)
    full = sum(t.valid_pairs for t in nat.pair_batches(5, 0.0, 8192, seed=1))
    thin = sum(t.valid_pairs for t in nat.pair_batches(5, 1e-4, 8192, seed=1))
    assert thin < 0.7 * full


def test_loader_throughput_is_gpu_class():
    nat = NativeCorpus.build(TEXT, min_count=1).encode(TEXT)
    t0 = time.time()
    n = 0
    for ep in range(6):
        n += sum(t.valid_pairs for t in nat.pair_batches(5, 0.0, 32768, seed=1, epoch=ep))
    rate = n / (time.time() - t0)
    assert rate > 5e6, f"{rate:.3g} pairs/s"      # the python generator manages ~0.3 M pairs/s


def test_native_text_parsers_match_numpy(tmp_path):
    from adapm_b200.models.kge import load_triples
    from adapm_b200.utils import mmio

    kge = os.path.join(ROOT, "data", "kge", "train.del")
    assert np.array_equal(load_triples(kge).numpy(), np.loadtxt(kge, dtype=np.int64))
    mm = os.path.join(ROOT, "data", "mf", "train.mmc")
    a, b = mmio.read_matrix_market_coo(mm), mmio.read_matrix_market_coo_py(mm)
    assert all(np.array_equal(x, y) for x, y in zip(a[:3], b[:3])) and a[3:] == b[3:]
    # a larger random matrix, with comment lines
    rng = np.random.default_rng(0)
    n = 20000
    i, j, x = rng.integers(1, 501, n), rng.integers(1, 301, n), rng.normal(size=n)
    p = tmp_path / "big.mmc"
    with open(p, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n% comment\n%\n" + f"500 300 {n}\n")
        for r in zip(i, j, x):
            f.write("%d %d %.8g\n" % r)
    gi, gj, gx, m, k = mmio.read_matrix_market_coo(str(p))
    assert (m, k) == (500, 300) and np.array_equal(gi, i - 1) and np.array_equal(gj, j - 1)
    assert np.allclose(gx, x.astype(np.float32), rtol=1e-6)
    bad = tmp_path / "bad.del"
    bad.write_text("1 2 3\n4 5\n")
    import pytest

    with pytest.raises(RuntimeError):
        load_triples(str(bad))
