"""Regroup the streams of a streams-YAML into K clients, IID by contiguous blocks
(ref: photon/dataset/stream_partitioner.py:11-41).

    python -m photon_b200.dataset.stream_partitioner <in.yaml> <out.yaml> <num_clients>
"""
from __future__ import annotations

import sys
from pathlib import Path

import yaml


def partition_stream_list(entries: list[dict], num_clients: int) -> list[dict]:
    flat = [st for entry in entries for st in (entry.get("client_streams") or {}).values()]
    if len(flat) % num_clients:
        raise ValueError(f"number of streams ({len(flat)}) must be divisible by num_clients ({num_clients})")
    step = len(flat) // num_clients
    return [{"client_streams": {f"stream_{i}": s for i, s in enumerate(flat[c * step:(c + 1) * step])}} for c in range(num_clients)]


def partition_streams(input_file: Path, output_file: Path, num_clients: int) -> None:
    out = partition_stream_list(yaml.safe_load(Path(input_file).read_text()), num_clients)
    Path(output_file).write_text(yaml.safe_dump(out, sort_keys=False))


if __name__ == "__main__":
    if len(sys.argv) != 4:
        raise SystemExit("usage: python -m photon_b200.dataset.stream_partitioner <input_file> <output_file> <num_clients>")
    partition_streams(Path(sys.argv[1]), Path(sys.argv[2]), int(sys.argv[3]))
