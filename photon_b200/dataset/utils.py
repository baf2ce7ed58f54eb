"""Tokenisation + packing helpers of the converter (role of ref photon/dataset/utils.py and the
llm-foundry ``ConcatTokensDataset``): documents → token ids → fixed-length int32 samples with
bos/eos separators, plus the unigram frequency counter behind ``1_gram.json``."""
from __future__ import annotations

import json
from collections import Counter
from pathlib import Path
from typing import Any, Iterable, Iterator

import numpy as np


class ByteTokenizer:
    """Offline fallback when no HF tokenizer files are available: bytes + 1, id 0 = <|endoftext|>."""

    vocab_size = 257
    eos_token_id = 0
    name = "byte-fallback"

    def encode(self, text: str) -> list[int]:
        return [b + 1 for b in text.encode("utf-8")]

    def decode(self, ids: Iterable[int]) -> str:
        return bytes(int(i) - 1 for i in ids if 0 < int(i) <= 256).decode("utf-8", errors="replace")

    def save_pretrained(self, path: str | Path) -> None:
        Path(path).mkdir(parents=True, exist_ok=True)
        (Path(path) / "tokenizer_config.json").write_text(json.dumps({"tokenizer_class": "ByteTokenizer", "vocab_size": 257}))


def build_tokenizer(name: str = "EleutherAI/gpt-neox-20b", **kwargs: Any) -> Any:
    """HF tokenizer when its files are locally cached; otherwise the byte-level fallback
    (there is no network here)."""
    try:
        from transformers import AutoTokenizer

        tok = AutoTokenizer.from_pretrained(name, local_files_only=True, **kwargs)
        tok.model_max_length = int(1e30)
        return tok
    except Exception:  # noqa: BLE001
        return ByteTokenizer()


def _encode(tok: Any, text: str) -> list[int]:
    if isinstance(tok, ByteTokenizer):
        return tok.encode(text)
    return tok(text, truncation=False, padding=False, add_special_tokens=False)["input_ids"]


def concat_tokens(docs: Iterable[str], tokenizer: Any, max_length: int, bos_text: str = "", eos_text: str = "<|endoftext|>",
                  no_wrap: bool = False) -> Iterator[np.ndarray]:
    """Pack documents into ``max_length`` windows (``--concat_tokens 2048 --eos_text '<|endoftext|>'``;
    ref: scripts/convert_c4_dataset.sh:46-50). The trailing partial window is dropped."""
    bos = _encode(tokenizer, bos_text) if bos_text else []
    eos = _encode(tokenizer, eos_text) if eos_text and not isinstance(tokenizer, ByteTokenizer) else ([tokenizer.eos_token_id] if eos_text else [])
    buf: list[int] = []
    for doc in docs:
        buf.extend(bos + _encode(tokenizer, doc) + eos)
        while len(buf) >= max_length:
            yield np.asarray(buf[:max_length], dtype=np.int32)
            buf = [] if no_wrap else buf[max_length:]


class UnigramCounter:
    def __init__(self) -> None:
        self.counts: Counter[int] = Counter()

    def update(self, sample: np.ndarray) -> None:
        ids, c = np.unique(sample, return_counts=True)
        self.counts.update(dict(zip(ids.tolist(), c.tolist())))

    def dump(self, path: str | Path) -> None:
        Path(path).write_text(json.dumps({str(k): int(v) for k, v in sorted(self.counts.items())}))


def iter_text_source(source: str, split: str | None = None, text_key: str = "text", limit: int | None = None) -> Iterator[str]:
    """Documents from: a ``.txt`` (one doc per line) / ``.jsonl`` file, a directory of those, an HF dataset
    name (only if cached locally), or ``synthetic://N`` (N lorem-ipsum-like docs for smoke tests)."""
    n = 0
    if source.startswith("synthetic://"):
        rng = np.random.default_rng(0)
        words = ["lorem", "ipsum", "dolor", "sit", "amet", "federated", "photon", "blackwell", "tensor", "memory", "kernel", "round"]
        for _ in range(int(source.split("://")[1] or 1000)):
            yield " ".join(rng.choice(words, size=int(rng.integers(20, 400))).tolist())
        return
    p = Path(source)
    files = sorted(p.rglob("*")) if p.is_dir() else [p] if p.exists() else []
    if files:
        for f in files:
            if f.suffix == ".jsonl":
                for line in f.open():
                    if line.strip():
                        yield json.loads(line)[text_key]
                        n += 1
                        if limit and n >= limit:
                            return
            elif f.suffix == ".txt":
                for line in f.open():
                    if line.strip():
                        yield line.rstrip("\n")
                        n += 1
                        if limit and n >= limit:
                            return
        return
    import datasets  # type: ignore[import-not-found]  (works only with a local HF cache)

    ds = datasets.load_dataset(source, split=split, streaming=True)
    for row in ds:
        yield row[text_key]
        n += 1
        if limit and n >= limit:
            return
