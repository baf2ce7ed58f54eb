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

from photon_b200.data.shards import MDSWriter, ShardWriter
from photon_b200.dataset.constants import DATASETS_CONSTANTS, resolve_split
from photon_b200.dataset.utils import UnigramCounter, build_tokenizer, concat_tokens, iter_text_source


def convert_split(docs: Any, tokenizer: Any, out_dirs: list[Path], seq_len: int, total_samples: int | None,
                  compression: str | None = "zlib", shard_samples: int = 8192, eos_text: str = "<|endoftext|>",
                  bos_text: str = "", no_wrap: bool = False, only_client: int | None = None, fmt: str = "photon") -> list[int]:
    """Write samples into ``len(out_dirs)`` contiguous partitions. When ``total_samples`` is unknown the
    samples are first counted into memory-light temporary order: we buffer and cut at the end. ``only_client`` writes
    just that partition (its samples are the same ones a full conversion would give it)."""
    n_clients = len(out_dirs)
    kw = dict(eos_text=eos_text, bos_text=bos_text, no_wrap=no_wrap)
    samples = list(concat_tokens(docs, tokenizer, seq_len, **kw)) if total_samples is None else None
    it = iter(samples) if samples is not None else concat_tokens(docs, tokenizer, seq_len, **kw)
    total = len(samples) if samples is not None else int(total_samples)
    per = total // n_clients
    written = []
    for c, d in enumerate(out_dirs):
        quota = per if c < n_clients - 1 else total - per * (n_clients - 1)
        if only_client is not None and c != only_client:
            for _ in range(quota):     # consume this partition's samples so the next one starts where it should
                next(it, None)
            written.append(0)
            continue
        counter = UnigramCounter()
        n = 0
        writer = (MDSWriter(d, seq_len=seq_len, compression="zstd" if compression == "zstd" else None) if fmt == "mds"
                  else ShardWriter(d, seq_len=seq_len, shard_samples=shard_samples, compression=compression))
        with writer as w:
            for _ in range(quota):
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


def parse_args(argv: list[str] | None = None) -> argparse.Namespace:
    """The reference's flag surface (ref: photon/dataset/convert_dataset_hf.py:30-172). ``--path`` / ``--name`` select an HF dataset
    when it is cached locally; ``--source`` takes a local text/jsonl file or directory (or ``synthetic://N``)."""
    ap = argparse.ArgumentParser(description="Convert a text corpus into per-client token shards")
    ap.add_argument("--dataset", default="c4_en", choices=sorted(DATASETS_CONSTANTS))
    ap.add_argument("--path", default=None, help="HF dataset path (reference flag); default: the table entry of --dataset")
    ap.add_argument("--name", default=None, help='HF dataset config name, e.g. "en" (reference flag; selects c4_<name> when --dataset is not given)')
    ap.add_argument("--splits", nargs="+", default=["train_small", "val_xxsmall"])
    ap.add_argument("--source", default=None, help="text/jsonl file or dir, HF dataset name, or synthetic://N (default: --path / the HF path of --dataset)")
    ap.add_argument("--out_root", required=True)
    ap.add_argument("--remote_path", default=None, help="mirror the converted tree here as well (the object-store upload of the reference)")
    ap.add_argument("--num_clients", type=int, default=8)
    ap.add_argument("--client", type=int, default=None, help="only write this client's partition (the others are skipped, not re-numbered)")
    ap.add_argument("--concat_tokens", type=int, default=2048)
    ap.add_argument("--tokenizer", default="EleutherAI/gpt-neox-20b")
    ap.add_argument("--tokenizer_kwargs", default=None, help="JSON dict forwarded to the tokenizer constructor")
    ap.add_argument("--bos_text", default=None)
    ap.add_argument("--eos_text", default="<|endoftext|>")
    ap.add_argument("--pad_text", default=None, help="accepted for flag parity; packed samples are never padded")
    ap.add_argument("--no_wrap", action="store_true", help="drop the tail of a document that does not fill the current sample instead of carrying it over")
    ap.add_argument("--num_workers", type=int, default=None, help="accepted for flag parity; conversion is a single streaming pass")
    ap.add_argument("--compression", default="zlib", choices=["zlib", "zstd", "none"])
    ap.add_argument("--format", default="photon", choices=["photon", "mds"],
                    help="photon = this package's shard format; mds = mosaicml-streaming's layout (what the reference writes and reads; "
                         "raw or zstd)")
    args = ap.parse_args(argv)
    if args.name and f"c4_{args.name}" in DATASETS_CONSTANTS and args.dataset == "c4_en":
        args.dataset = f"c4_{args.name}"
    return args


def main(argv: list[str] | None = None) -> dict[str, list[int]]:
    """Convert the selected splits into per-client token shards (ref: photon/dataset/convert_dataset_hf.py:175-330)."""
    import json
    import shutil

    args = parse_args(argv)
    if args.compression == "zstd":
        from photon_b200.data.shards import _zstd

        if _zstd() is None:
            print("[convert] no zstd codec importable here; writing zlib shards (the reader picks the codec from index.json)")
            args.compression = "zlib"
    if args.dataset not in DATASETS_CONSTANTS:
        raise SystemExit(f"unknown dataset {args.dataset!r} (have {sorted(DATASETS_CONSTANTS)})")
    tok = build_tokenizer(args.tokenizer, **(json.loads(args.tokenizer_kwargs) if args.tokenizer_kwargs else {}))
    lang = args.dataset.split("_", 1)[1]
    root = Path(args.out_root) / f"c{args.num_clients}" / lang
    out: dict[str, list[int]] = {}
    for key in args.splits:
        sc = resolve_split(args.dataset, key)      # table key (reference names; "val" = "validation") -> HF split, folder, truncation
        fs = sc.folder_split
        src = args.source or args.path or sc.path
        docs = iter_text_source(src, split=sc.split, limit=sc.truncated_samples)
        dirs = [root / f"client_{i}" / fs for i in range(args.num_clients)]
        out[fs] = convert_split(docs, tok, dirs, args.concat_tokens, None, None if args.compression == "none" else args.compression,
                                eos_text=args.eos_text, bos_text=args.bos_text or "", no_wrap=args.no_wrap, only_client=args.client,
                                fmt=args.format)
        print(f"[convert] {args.dataset}/{fs}: {out[fs]} samples per client -> {root}")
    tok.save_pretrained(str(Path(args.out_root) / "tokenizer"))
    if args.remote_path:
        shutil.copytree(args.out_root, args.remote_path, dirs_exist_ok=True)
        print(f"[convert] mirrored {args.out_root} -> {args.remote_path}")
    return out


if __name__ == "__main__":
    main()
