"""Synthetic C4-shaped token streams (there is no network for real C4).

Shape contract of the reference's converted dataset: fully packed 2048-token
int32 samples, documents separated by the tokenizer's ``<|endoftext|>`` id (0
for gpt-neox-20b), ids < 50277 inside a model vocab padded to 50368
(ref: scripts/convert_c4_dataset.sh:46-50).  Token ids are
Zipf-distributed and document lengths log-normal (C4-like median ≈ 400
tokens) so the unigram metrics and the loss curve behave like text, and every
sample is a pure function of ``(seed, stream_id, index)`` → resumable and
identical across processes without any files.
"""
from __future__ import annotations

import numpy as np

EOS_ID = 0
TOKENIZER_VOCAB = 50277


class SyntheticC4:
    def __init__(self, seq_len: int = 2048, vocab_size: int = TOKENIZER_VOCAB, seed: int = 17,
                 stream_id: int = 0, num_samples: int | None = None, zipf_a: float = 1.15) -> None:
        self.seq_len, self.vocab_size, self.seed = int(seq_len), int(vocab_size), int(seed)
        self.stream_id, self.num_samples = int(stream_id), num_samples
        ranks = np.arange(1, self.vocab_size, dtype=np.float64)
        p = ranks ** (-zipf_a)
        self._cdf = np.cumsum(p / p.sum())
        # a fixed permutation so frequent ids are not simply the small integers
        self._perm = np.random.default_rng(1234567).permutation(self.vocab_size - 1) + 1

    def __len__(self) -> int:
        if self.num_samples is None:
            raise TypeError("infinite synthetic stream has no len()")
        return self.num_samples

    def unigram_probabilities(self) -> np.ndarray:
        """Exact generating distribution over ids (EOS mass approximated by 1/median_doc_len)."""
        p = np.zeros(self.vocab_size, dtype=np.float64)
        pdf = np.diff(np.concatenate([[0.0], self._cdf]))
        p[self._perm] = pdf
        eos = 1.0 / 400.0
        p *= (1.0 - eos)
        p[EOS_ID] = eos
        return p

    def __getitem__(self, idx: int) -> np.ndarray:
        if self.num_samples is not None and not 0 <= idx < self.num_samples:
            raise IndexError(idx)
        rng = np.random.default_rng([self.seed, self.stream_id, int(idx)])
        u = rng.random(self.seq_len)
        toks = self._perm[np.searchsorted(self._cdf, u, side="left").clip(max=self.vocab_size - 2)]
        toks = toks.astype(np.int32)
        pos = int(rng.integers(0, 400))
        while pos < self.seq_len:  # sprinkle document boundaries
            toks[pos] = EOS_ID
            pos += max(8, int(rng.lognormal(mean=6.0, sigma=0.9)))
        return toks


def synthetic_batch(batch: int, seq_len: int, vocab_size: int = TOKENIZER_VOCAB, seed: int = 0,
                    stream_id: int = 0, start: int = 0) -> np.ndarray:
    ds = SyntheticC4(seq_len, vocab_size, seed, stream_id)
    return np.stack([ds[start + i] for i in range(batch)]).astype(np.int64)
