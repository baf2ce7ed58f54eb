"""Sample generators (ref: photon/dataset/samples_generators.py:63-221): produce fixed-length token
samples from raw text, or RE-tokenise / re-chunk samples out of existing shard directories
(e.g. change ``seq_len`` or vocabulary without touching the raw corpus)."""
from __future__ import annotations

from pathlib import Path
from typing import Any, Iterable, Iterator

import numpy as np

from photon_b200.data.shards import ShardReader
from photon_b200.dataset.utils import concat_tokens


def generate_samples_from_text(docs: Iterable[str], tokenizer: Any, seq_len: int, eos_text: str = "<|endoftext|>",
                               bos_text: str = "") -> Iterator[dict[str, np.ndarray]]:
    for s in concat_tokens(docs, tokenizer, seq_len, bos_text=bos_text, eos_text=eos_text):
        yield {"tokens": s}


def generate_samples_from_shards(directory: str | Path, new_seq_len: int | None = None) -> Iterator[dict[str, np.ndarray]]:
    """Stream samples back out of a shard directory, optionally re-chunked to ``new_seq_len``."""
    r = ShardReader(directory)
    if new_seq_len is None or new_seq_len == r.seq_len:
        for i in range(len(r)):
            yield {"tokens": r[i]}
        return
    buf = np.empty(0, dtype=np.int32)
    for i in range(len(r)):
        buf = np.concatenate([buf, r[i]])
        while buf.size >= new_seq_len:
            yield {"tokens": buf[:new_seq_len].copy()}
            buf = buf[new_seq_len:]


def retokenize_samples(directory: str | Path, old_tokenizer: Any, new_tokenizer: Any, seq_len: int) -> Iterator[dict[str, np.ndarray]]:
    """Decode with the old tokenizer, re-encode + re-pack with the new one."""
    def docs() -> Iterator[str]:
        for s in generate_samples_from_shards(directory):
            yield old_tokenizer.decode(s["tokens"].tolist())

    yield from generate_samples_from_text(docs(), new_tokenizer, seq_len)
