"""Corpus → federated token shards (ref: photon/dataset/convert_dataset_hf.py:234-363).

Pipeline: stream documents → tokenize (gpt-neox-20b when its files are cached, byte fallback
otherwise) → pack into ``seq_len`` int32 samples → split SEQUENTIALLY into ``num_clients`` contiguous
partitions → one shard directory per client/split::

    {out_root}/{name}/client_{i}/{folder_split}/index.json + shard.*.tok[.z]
    {out_root}/{name}/client_{i}/{folder_split}/1_gram.json     (token-frequency map → unigram metrics)
    {out_root}/{name}/tokenizer/                                (tokenizer dump)

    python -m photon_b200.dataset.convert_dataset_hf --dataset c4_en --splits train_small val_xxsmall \
        --source /data/c4-texts --out_root ./fed-c4 --num_clients 8 --concat_tokens 2048
"""
from __future__ import annotations

import argparse
from pathlib import Path
from typing import Any

from photon_b200.data.shards import ShardWriter
from photon_b200.dataset.constants import DATASETS_CONSTANTS
from photon_b200.dataset.utils import UnigramCounter, build_tokenizer, concat_tokens, iter_text_source


def convert_split(docs: Any, tokenizer: Any, out_dirs: list[Path], seq_len: int, total_samples: int | None,
                  compression: str | None = "zlib", shard_samples: int = 8192, eos_text: str = "<|endoftext|>") -> list[int]:
    """Write samples into ``len(out_dirs)`` contiguous partitions. When ``total_samples`` is unknown the
    samples are first counted into memory-light temporary order: we buffer and cut at the end."""
    n_clients = len(out_dirs)
    samples = list(concat_tokens(docs, tokenizer, seq_len, eos_text=eos_text)) if total_samples is None else None
    it = iter(samples) if samples is not None else concat_tokens(docs, tokenizer, seq_len, eos_text=eos_text)
    total = len(samples) if samples is not None else int(total_samples)
    per = total // n_clients
    written = []
    for c, d in enumerate(out_dirs):
        counter = UnigramCounter()
        n = 0
        with ShardWriter(d, seq_len=seq_len, shard_samples=shard_samples, compression=compression) as w:
            for _ in range(per if c < n_clients - 1 else total - per * (n_clients - 1)):
                try:
                    s = next(it)
                except StopIteration:
                    break
                w.write(s)
                counter.update(s)
                n += 1
        counter.dump(d / "1_gram.json")
        written.append(n)
    return written


def main(argv: list[str] | None = None) -> dict[str, list[int]]:
    ap = argparse.ArgumentParser(description="Convert a text corpus into per-client token shards")
    ap.add_argument("--dataset", default="c4_en", choices=sorted(DATASETS_CONSTANTS))
    ap.add_argument("--splits", nargs="+", default=["train_small", "val_xxsmall"])
    ap.add_argument("--source", default=None, help="text/jsonl file or dir, HF dataset name, or synthetic://N (default: the HF path of --dataset)")
    ap.add_argument("--out_root", required=True)
    ap.add_argument("--num_clients", type=int, default=8)
    ap.add_argument("--concat_tokens", type=int, default=2048)
    ap.add_argument("--tokenizer", default="EleutherAI/gpt-neox-20b")
    ap.add_argument("--eos_text", default="<|endoftext|>")
    ap.add_argument("--compression", default="zlib", choices=["zlib", "none"])
    args = ap.parse_args(argv)
    consts = DATASETS_CONSTANTS[args.dataset]
    tok = build_tokenizer(args.tokenizer)
    lang = args.dataset.split("_", 1)[1]
    root = Path(args.out_root) / f"c{args.num_clients}" / lang
    out: dict[str, list[int]] = {}
    for fs in args.splits:
        sc = consts.splits[fs]
        src = args.source or sc.path
        docs = iter_text_source(src, split=sc.split, limit=sc.truncated_samples)
        dirs = [root / f"client_{i}" / fs for i in range(args.num_clients)]
        out[fs] = convert_split(docs, tok, dirs, args.concat_tokens, None, None if args.compression == "none" else "zlib", eos_text=args.eos_text)
        print(f"[convert] {args.dataset}/{fs}: {out[fs]} samples per client -> {root}")
    tok.save_pretrained(str(Path(args.out_root) / "tokenizer"))
    return out


if __name__ == "__main__":
    main()
