"""Corpus handling for word2vec (reference apps/word2vec.cc:147-364: vocabulary building, sorting by
count, min_count pruning, frequent-word subsampling, per-worker file partitioning)."""
from __future__ import annotations

from collections import Counter
from typing import Iterator, List

import numpy as np
import torch


class Vocabulary:
    def __init__(self, words: List[str], counts: np.ndarray):
        self.words, self.counts = words, counts
        self.index = {w: i for i, w in enumerate(words)}

    @staticmethod
    def build(path: str, min_count: int = 5) -> "Vocabulary":
        c = Counter()
        with open(path, "r", errors="ignore") as f:
            for line in f:
                c.update(line.split())
        items = sorted(((w, n) for w, n in c.items() if n >= min_count), key=lambda t: -t[1])
        # like word2vec.c, index 0 is the sentence delimiter token
        words = ["</s>"] + [w for w, _ in items]
        counts = np.array([max(1, sum(1 for _ in open(path, errors="ignore")))] + [n for _, n in items], dtype=np.int64)
        return Vocabulary(words, counts)

    def save(self, path: str) -> None:
        with open(path, "w") as f:
            for w, n in zip(self.words, self.counts):
                f.write(f"{w} {int(n)}\n")

    @staticmethod
    def load(path: str) -> "Vocabulary":
        words, counts = [], []
        with open(path) as f:
            for line in f:
                w, n = line.rsplit(" ", 1)
                words.append(w); counts.append(int(n))
        return Vocabulary(words, np.array(counts, dtype=np.int64))


def read_sentences(path: str, vocab: Vocabulary, rank: int, world: int, subsample: float, rng) -> Iterator[np.ndarray]:
    """This rank's share of the corpus (line i belongs to rank i % world), as arrays of word ids with
    frequent-word subsampling (word2vec.c rule)."""
    total = float(vocab.counts.sum())
    with open(path, "r", errors="ignore") as f:
        for li, line in enumerate(f):
            if li % world != rank:
                continue
            ids = np.array([vocab.index[w] for w in line.split() if w in vocab.index], dtype=np.int64)
            if subsample > 0 and ids.size:
                fr = vocab.counts[ids] / total
                keep = (np.sqrt(fr / subsample) + 1) * subsample / fr
                ids = ids[rng.random(ids.size) < keep]
            if ids.size > 1:
                yield ids


def pairs_from_sentences(sentences: Iterator[np.ndarray], window: int, batch_pairs: int, rng) -> Iterator[torch.Tensor]:
    """(center, context) key batches [2, B]: for every position a random window shrink b in [0, window)
    like the reference (word2vec.cc:672-690); centers are syn0 keys of the context word, targets syn1 keys."""
    buf_c, buf_t, n = [], [], 0
    for s in sentences:
        L = s.size
        for pos in range(L):
            b = int(rng.integers(0, window))
            lo, hi = max(0, pos - (window - b)), min(L, pos + (window - b) + 1)
            ctx = np.concatenate([s[lo:pos], s[pos + 1:hi]])
            if ctx.size == 0:
                continue
            buf_c.append(2 * ctx)                                   # syn0 key of the context word
            buf_t.append(np.full(ctx.size, 2 * s[pos] + 1))         # syn1 key of the centre word
            n += ctx.size
            if n >= batch_pairs:
                c, t = np.concatenate(buf_c), np.concatenate(buf_t)
                yield torch.from_numpy(np.stack([c[:batch_pairs], t[:batch_pairs]]))
                buf_c, buf_t = [c[batch_pairs:]], [t[batch_pairs:]]
                n = buf_c[0].size
    if n:
        yield torch.from_numpy(np.stack([np.concatenate(buf_c), np.concatenate(buf_t)]))


# ------------------------------------------------------------------------------------------ native loader
class NativeCorpus:
    """C++ corpus (csrc/adapm/corpus.{h,cc}): same vocabulary rules as :class:`Vocabulary` (count-descending, ties by
    first occurrence, ``</s>`` = word 0), plus the encoded sentences of this rank and a background pair generator."""

    def __init__(self, impl):
        self._impl = impl

    @staticmethod
    def build(path: str, min_count: int = 5) -> "NativeCorpus":
        from .. import _C

        return NativeCorpus(_C.Corpus.build(path, int(min_count)))

    @staticmethod
    def from_vocabulary(vocab: Vocabulary) -> "NativeCorpus":
        from .. import _C

        return NativeCorpus(_C.Corpus.from_vocab(list(vocab.words), [int(c) for c in vocab.counts]))

    def vocabulary(self) -> Vocabulary:
        return Vocabulary(list(self._impl.words()), np.array(self._impl.counts(), dtype=np.int64))

    def encode(self, path: str, rank: int = 0, world: int = 1) -> "NativeCorpus":
        self._impl.encode(path, int(rank), int(world))
        return self

    def num_tokens(self) -> int:
        return self._impl.num_tokens()

    def num_sentences(self) -> int:
        return self._impl.num_sentences()

    def sentences(self) -> List[np.ndarray]:
        """The encoded sentences (copies; for tests and small corpora)."""
        import ctypes

        n, ns = self.num_tokens(), self.num_sentences()
        tok = np.ctypeslib.as_array((ctypes.c_int32 * max(n, 1)).from_address(self._impl.tokens_ptr()))[:n].copy()
        off = np.ctypeslib.as_array((ctypes.c_int64 * (ns + 1)).from_address(self._impl.sentence_offsets_ptr())).copy()
        return [tok[off[i]:off[i + 1]] for i in range(ns)]

    def pair_batches(self, window: int, subsample: float, batch_pairs: int, seed: int, epoch: int = 0,
                     pin: bool = False, queue_depth: int = 8) -> Iterator[torch.Tensor]:
        """[2, batch_pairs] key batches of one epoch, generated by a background C++ thread. A short final batch is
        padded by repeating its pairs; ``valid_pairs`` of the yielded tensor tells how many are real."""
        from .. import _C

        ps = _C.PairStream(self._impl, int(window), float(subsample), int(batch_pairs), int(seed), int(queue_depth))
        ps.start_epoch(int(epoch))
        while True:
            t = torch.empty(2, batch_pairs, dtype=torch.int64)
            if pin:
                t = t.pin_memory()
            u = torch.empty(2 * batch_pairs, dtype=torch.int64)
            valid, n_u = ps.next_with_unique(t.data_ptr(), u.data_ptr())
            if valid == 0:
                return
            t.valid_pairs = int(valid)
            t.unique_keys = u[:n_u]          # what to call Intent() with (deduplicated on the loader thread)
            yield t
