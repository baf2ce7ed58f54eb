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
    """HF tokenizer when its files are locally cached (there is no network here). ``name="byte"`` asks for the byte-level tokenizer
    explicitly; any OTHER name that cannot be loaded falls back to it LOUDLY — a corpus tokenised with the wrong vocabulary trains and
    evaluates without complaint, so ``PHOTON_STRICT_DATA=1`` (exported by the launch scripts) makes the fall-back an error."""
    import os
    import sys

    if str(name).lower() in ("byte", "bytes", "byte-fallback"):
        return ByteTokenizer()
    try:
        from transformers import AutoTokenizer

        tok = AutoTokenizer.from_pretrained(name, local_files_only=True, **kwargs)
        tok.model_max_length = int(1e30)
        return tok
    except Exception as e:  # noqa: BLE001
        if os.environ.get("PHOTON_STRICT_DATA", "0") not in ("", "0", "false", "False"):
            raise RuntimeError(f"tokenizer {name!r} is not available locally ({type(e).__name__}: {str(e)[:160]}); with PHOTON_STRICT_DATA "
                               "set the byte-level fall-back is not used — cache the tokenizer files or pass --tokenizer byte") from e
        print(f"[tokenizer] WARNING: {name!r} could not be loaded from the local cache ({type(e).__name__}) -> BYTE-LEVEL fall-back "
              "(vocabulary 257). Token ids will NOT match a model trained with the real tokenizer. Set PHOTON_STRICT_DATA=1 to make "
              "this an error, or pass --tokenizer byte to silence it.", file=sys.stderr, flush=True)
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


# ----------------------------------------------------------------------------- reference-named entry points
def check_tokenizer_config(tokenizer: Any, bos_text: str = "", eos_text: str = "") -> None:
    """Catch the double-special-token trap before converting a corpus (ref: photon/dataset/utils.py:27-102): the converter adds
    ``bos_text`` / ``eos_text`` itself, so a tokenizer that ALSO inserts BOS / EOS on its own would put two of them at every document
    boundary. Documents are encoded with ``add_special_tokens=False`` here, which makes that impossible; this check verifies it on
    a probe string and that the boundary texts encode to something."""
    if isinstance(tokenizer, ByteTokenizer):
        return
    probe = _encode(tokenizer, "probe text")
    for name, tid in (("bos", getattr(tokenizer, "bos_token_id", None)), ("eos", getattr(tokenizer, "eos_token_id", None))):
        if tid is not None and probe and (probe[0] == tid or probe[-1] == tid):
            raise ValueError(f"the tokenizer inserts its {name} token by itself even with add_special_tokens=False; "
                             f"pass an empty {name}_text or fix the tokenizer configuration")
    for label, text in (("bos_text", bos_text), ("eos_text", eos_text)):
        if text and not _encode(tokenizer, text):
            raise ValueError(f"{label}={text!r} encodes to no tokens with this tokenizer")


def build_hf_dataset(path: str, split: str, mode: Any = "CONCAT_TOKENS", temp_dir: Any = None, max_length: int | None = None,
                     bos_text: str = "", eos_text: str = "", tokenizer: Any = None, name: str | None = None, *,
                     no_wrap: bool = False, limit: int | None = None) -> Iterator[dict[str, Any]]:
    """An iterable of converter samples from a text source (ref: photon/dataset/utils.py:105-210): ``NO_CONCAT`` yields
    ``{"text": …}`` per document, ``CONCAT_TOKENS`` yields ``{"tokens": int32[max_length]}`` windows packed with BOS / EOS.
    ``path`` is anything :func:`iter_text_source` reads (text / jsonl files, a directory, ``synthetic://N``, or a Hugging Face
    dataset that is in the local cache); ``name`` selects the language folder of a C4-style local tree; ``temp_dir`` is accepted
    for signature compatibility (nothing is staged)."""
    del temp_dir
    mode_s = getattr(mode, "value", mode)
    src = str(Path(path) / name) if name and (Path(path) / name).exists() else path
    docs = iter_text_source(src, split=split, limit=limit)
    if mode_s == "NO_CONCAT":
        return ({"text": d} for d in docs)
    if mode_s != "CONCAT_TOKENS":
        raise ValueError(f"unknown concatenation mode {mode!r}")
    if tokenizer is None or not max_length:
        raise ValueError("CONCAT_TOKENS needs a tokenizer and max_length")
    check_tokenizer_config(tokenizer, bos_text, eos_text)
    return ({"tokens": s} for s in concat_tokens(docs, tokenizer, int(max_length), bos_text=bos_text, eos_text=eos_text, no_wrap=no_wrap))


def build_dataloader(dataset: Any, batch_size: int, num_workers: int | None = None) -> Any:
    """Batches of converter samples (ref: photon/dataset/utils.py:213-260: a torch DataLoader whose worker count defaults to the
    CPU count). The samples here come out of a generator that is already the bottleneck-free part of the conversion (tokenisation
    dominates), so batching happens in-process: lists of ``batch_size`` samples, stacked when they are token windows."""
    del num_workers

    def batches() -> Iterator[dict[str, Any]]:
        buf: list[dict[str, Any]] = []
        for s in dataset:
            buf.append(s)
            if len(buf) == batch_size:
                yield _collate(buf)
                buf = []
        if buf:
            yield _collate(buf)

    return batches()


def _collate(samples: list[dict[str, Any]]) -> dict[str, Any]:
    out: dict[str, Any] = {}
    for k in samples[0]:
        vals = [s[k] for s in samples]
        out[k] = np.stack(vals) if isinstance(vals[0], np.ndarray) else vals
    return out
