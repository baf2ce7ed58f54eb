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


# ----------------------------------------------------------------------------- reference-named, loader-based forms
def generate_samples_from_dataloader(loader: Iterable[dict[str, Any]], truncate_num_samples: int | None = None) -> Iterator[dict[str, Any]]:
    """Un-batch a loader: one dict per sample, at most ``truncate_num_samples`` of them (ref: samples_generators.py:23-60)."""
    n = 0
    for batch in loader:
        keys = list(batch)
        for i in range(len(batch[keys[0]])):
            if truncate_num_samples is not None and 0 <= truncate_num_samples <= n:
                return
            yield {k: batch[k][i] for k in keys}
            n += 1


def stream_and_untokenize(loader: Iterable[dict[str, Any]], tokenizer: Any, truncate_num_batches: int | None = None) -> Iterator[str]:
    """Token batches back to text, one string per sample (ref: samples_generators.py:63-126). Batches carry ``tokens`` (our shards)
    or ``input_ids`` (a text loader)."""
    for b, batch in enumerate(loader):
        if truncate_num_batches is not None and b >= truncate_num_batches:
            return
        rows = batch["tokens"] if "tokens" in batch else batch["input_ids"]
        for row in rows:
            yield tokenizer.decode(np.asarray(row).tolist())


def generate_samples_retokenized_streaming_text_dataset(loader: Iterable[dict[str, Any]], tokenizer_couple: Any, max_length: int,
                                                        truncate_num_samples: int | None = None, *, no_wrap: bool = False,
                                                        bos_text: str = "", eos_text: str = "") -> Iterator[dict[str, np.ndarray]]:
    """Change the vocabulary of an already tokenised stream (ref: samples_generators.py:129-221): every sample is decoded with
    ``tokenizer_couple.decode_tokenizer``, re-encoded with ``.encode_tokenizer`` and re-packed into ``max_length`` windows."""
    docs = stream_and_untokenize(loader, tokenizer_couple.decode_tokenizer)
    n = 0
    for s in concat_tokens(docs, tokenizer_couple.encode_tokenizer, max_length, bos_text=bos_text, eos_text=eos_text, no_wrap=no_wrap):
        if truncate_num_samples is not None and 0 <= truncate_num_samples <= n:
            return
        yield {"tokens": s}
        n += 1
