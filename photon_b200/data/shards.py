"""Token shard format (reader + writer).

Plays the role of mosaicml-streaming's MDS shards in the reference
(ref: photon/dataset/convert_dataset_hf.py:323-327 writes
``columns={"tokens": "ndarray:int32"}`` zstd shards + ``index.json``).
The on-disk layout is our own, sized for fast sequential reads into pinned
host memory: a directory holding

* ``index.json`` — ``{"version", "format": "pb200-tokens", "seq_len",
  "dtype": "int32", "compression": null|"zlib"|"zstd", "shards": [{"basename",
  "samples", "raw_bytes", "zip_bytes", "sha1"}...]}``
* ``shard.00000.tok[.z]`` — a C-contiguous ``[samples, seq_len]`` int32 matrix.

(zstd — the reference's MDS default — comes from pyarrow's bundled codec; zlib needs nothing.)
"""
from __future__ import annotations

import hashlib
import json
import os
import zlib
from pathlib import Path
from typing import Any, Iterable

import numpy as np

INDEX_NAME = "index.json"
FORMAT = "pb200-tokens"


def _zstd() -> Any:
    """zstd through pyarrow's bundled codec (the ``zstandard`` module is not installed here); None if unavailable."""
    try:
        import pyarrow as pa

        return pa.Codec("zstd", compression_level=3) if pa.Codec.is_available("zstd") else None
    except ImportError:
        return None


def _zstd_decompress(blob: bytes, raw_bytes: int) -> bytes:
    codec = _zstd()
    if codec is None:
        raise RuntimeError("this shard is zstd-compressed and no zstd codec is importable")
    return codec.decompress(blob, decompressed_size=raw_bytes, asbytes=True)


class ShardWriter:
    """Accumulates fixed-length int32 samples and cuts a shard every ``shard_samples``."""

    def __init__(self, out_dir: str | os.PathLike, seq_len: int, shard_samples: int = 8192,
                 compression: str | None = None) -> None:
        if compression not in (None, "zlib", "zstd"):
            raise ValueError("compression must be null, 'zlib' or 'zstd'")
        if compression == "zstd" and _zstd() is None:
            raise RuntimeError("zstd shards need pyarrow's codec (no zstandard module in this image)")
        self.out = Path(out_dir)
        self.out.mkdir(parents=True, exist_ok=True)
        self.seq_len, self.shard_samples, self.compression = int(seq_len), int(shard_samples), compression
        self._buf: list[np.ndarray] = []
        self._shards: list[dict[str, Any]] = []

    def write(self, tokens: np.ndarray) -> None:
        a = np.asarray(tokens, dtype=np.int32).reshape(-1)
        if a.size != self.seq_len:
            raise ValueError(f"sample has {a.size} tokens, expected {self.seq_len}")
        self._buf.append(a)
        if len(self._buf) >= self.shard_samples:
            self._flush()

    def write_many(self, rows: Iterable[np.ndarray]) -> None:
        for r in rows:
            self.write(r)

    def _flush(self) -> None:
        if not self._buf:
            return
        mat = np.stack(self._buf).astype(np.int32, copy=False)
        raw = mat.tobytes()
        base = f"shard.{len(self._shards):05d}.tok"
        payload = raw
        if self.compression == "zlib":
            payload, base = zlib.compress(raw, 3), base + ".z"
        elif self.compression == "zstd":
            payload, base = _zstd().compress(raw, asbytes=True), base + ".zstd"
        (self.out / base).write_bytes(payload)
        self._shards.append({"basename": base, "samples": int(mat.shape[0]), "raw_bytes": len(raw),
                             "zip_bytes": len(payload), "sha1": hashlib.sha1(raw).hexdigest()})  # noqa: S324
        self._buf = []

    def finish(self) -> dict[str, Any]:
        self._flush()
        index = {"version": 1, "format": FORMAT, "seq_len": self.seq_len, "dtype": "int32",
                 "compression": self.compression, "shards": self._shards}
        (self.out / INDEX_NAME).write_text(json.dumps(index, indent=1))
        return index

    def __enter__(self) -> "ShardWriter":
        return self

    def __exit__(self, *exc: Any) -> None:
        self.finish()


class ShardReader:
    """Random access over one shard directory (memory-maps raw shards)."""

    def __init__(self, directory: str | os.PathLike, validate_hash: bool = False) -> None:
        self.dir = Path(directory)
        idx = self.dir / INDEX_NAME
        if not idx.exists():
            raise FileNotFoundError(f"no {INDEX_NAME} under {self.dir}")
        self.index = json.loads(idx.read_text())
        if self.index.get("format") != FORMAT:
            raise ValueError(f"{idx}: unknown shard format {self.index.get('format')!r}")
        self.seq_len = int(self.index["seq_len"])
        self.validate_hash = validate_hash
        counts = [int(s["samples"]) for s in self.index["shards"]]
        self._starts = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        self._cache: dict[int, np.ndarray] = {}

    def __len__(self) -> int:
        return int(self._starts[-1])

    def _shard(self, i: int) -> np.ndarray:
        if i not in self._cache:
            meta = self.index["shards"][i]
            path = self.dir / meta["basename"]
            codec = self.index.get("compression")
            if codec in ("zlib", "zstd"):
                blob = path.read_bytes()
                raw = zlib.decompress(blob) if codec == "zlib" else _zstd_decompress(blob, int(meta["raw_bytes"]))
                mat = np.frombuffer(raw, dtype=np.int32).reshape(meta["samples"], self.seq_len)
            else:
                mat = np.memmap(path, dtype=np.int32, mode="r", shape=(meta["samples"], self.seq_len))
            if self.validate_hash and hashlib.sha1(np.ascontiguousarray(mat).tobytes()).hexdigest() != meta["sha1"]:  # noqa: S324
                raise OSError(f"hash mismatch in {path}")
            if len(self._cache) > 8:
                self._cache.pop(next(iter(self._cache)))
            self._cache[i] = mat
        return self._cache[i]

    def __getitem__(self, idx: int) -> np.ndarray:
        if not 0 <= idx < len(self):
            raise IndexError(idx)
        s = int(np.searchsorted(self._starts, idx, side="right") - 1)
        return np.asarray(self._shard(s)[idx - int(self._starts[s])])


class MDSReader:
    """Read-only access to token datasets written by mosaicml-streaming's ``MDSWriter`` (the reference's on-disk format:
    ``MDSWriter(columns={"tokens": "ndarray:int32"}, compression="zstd")``, ref: photon/dataset/convert_dataset_hf.py:234-363),
    so datasets converted for the reference can be trained on without re-tokenising.

    Layout handled (streaming ``format: mds``, index ``version: 2``): ``index.json`` lists shards with ``raw_data`` /
    ``zip_data`` basenames, ``samples``, ``column_names`` / ``column_encodings`` / ``column_sizes`` and ``compression``.
    A raw shard is ``uint32 n | uint32 offsets[n+1] | json config | samples``; offsets are absolute byte positions; a sample
    is ``uint32 size`` per variable-size column followed by the column payloads; an ``ndarray:<dtype>`` payload is a short
    shape header followed by the C-order data, so the tokens are the trailing ``4·n`` bytes and the header is cross-checked
    against ``n``. Only the token column is decoded. Compressed shards (``.zstd``) go through pyarrow's codec.

    NB: implemented from the format description — there is no mosaicml-streaming install or MDS sample in this environment;
    the unit test builds shards following the same description.
    """

    def __init__(self, directory: str | os.PathLike, column: str = "tokens", seq_len: int | None = None, validate_hash: bool = False) -> None:
        self.dir = Path(directory)
        self.index = json.loads((self.dir / INDEX_NAME).read_text())
        shards = self.index.get("shards") or []
        if not shards or any(sh.get("format") != "mds" for sh in shards):
            raise ValueError(f"{self.dir}: not an MDS index")
        self.column = column
        self._meta = shards
        counts = [int(sh["samples"]) for sh in shards]
        self._starts = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        self._cache: dict[int, tuple[bytes, np.ndarray]] = {}
        self.validate_hash = validate_hash
        self.seq_len = int(seq_len) if seq_len else int(self[0].shape[0])

    def __len__(self) -> int:
        return int(self._starts[-1])

    def _load(self, i: int) -> tuple[bytes, np.ndarray]:
        if i not in self._cache:
            sh = self._meta[i]
            raw_name, zip_info = sh["raw_data"]["basename"], sh.get("zip_data")
            if (self.dir / raw_name).exists():
                blob = (self.dir / raw_name).read_bytes()
            elif zip_info and (self.dir / zip_info["basename"]).exists():
                codec = str(sh.get("compression") or "").split(":")[0]
                z = (self.dir / zip_info["basename"]).read_bytes()
                if codec == "zstd":
                    blob = _zstd_decompress(z, int(sh["raw_data"]["bytes"]))
                elif codec in ("zlib", "gz"):
                    blob = zlib.decompress(z, 15 + 32)
                else:
                    raise ValueError(f"{self.dir}: unsupported MDS compression {sh.get('compression')!r}")
            else:
                raise FileNotFoundError(f"{self.dir}: neither {raw_name} nor its compressed form is present")
            n = int(np.frombuffer(blob, np.uint32, 1)[0])
            if n != int(sh["samples"]):
                raise OSError(f"{self.dir / raw_name}: header says {n} samples, index says {sh['samples']}")
            offsets = np.frombuffer(blob, np.uint32, n + 1, offset=4)
            if len(self._cache) > 4:
                self._cache.pop(next(iter(self._cache)))
            self._cache[i] = (blob, offsets)
        return self._cache[i]

    def _decode(self, sample: bytes, sh: dict[str, Any]) -> np.ndarray:
        names, encs, sizes = sh["column_names"], sh["column_encodings"], sh["column_sizes"]
        pos, lens = 0, []
        for sz in sizes:
            if sz:
                lens.append(int(sz))
            else:
                lens.append(int(np.frombuffer(sample, np.uint32, 1, offset=pos)[0]))
                pos += 4
        for name, enc, ln in zip(names, encs, lens):
            if name != self.column:
                pos += ln
                continue
            payload = sample[pos:pos + ln]
            kind, _, rest = enc.partition(":")
            if kind == "bytes":
                # older llm-foundry converters store ``np.int64`` token ids as raw bytes (read back with
                # ``np.frombuffer(sample["tokens"], dtype=np.int64)``); accept int32 too when the length only fits that
                want = getattr(self, "seq_len", None)
                dt = np.dtype(np.int64) if (ln % 8 == 0 and (want is None or ln == 8 * want)) else np.dtype(np.int32)
                return np.frombuffer(payload, dt).astype(np.int32)
            if kind != "ndarray":
                raise ValueError(f"column {name!r} has encoding {enc!r}; expected ndarray:<dtype> or bytes")
            dtype = np.dtype(rest.split(":")[0] or "int32")
            # dynamic-shape header of mosaicml-streaming's NDArray encoding: ``uint8 ndim | uint8 code of the dims' integer type |
            # ndim dims`` (dims in the narrowest unsigned type that holds them: 2048 tokens -> uint16), then the C-order data. Token rows
            # have ndim 1. A one-byte prefix (rank and width packed, or no type code) is accepted as well; either way the dims must
            # equal the element count the remaining bytes hold, so a mis-parse cannot slip through.
            for prefix in (2, 1):
                for width in (1, 2, 4, 8):
                    h = prefix + width
                    body = ln - h
                    if body <= 0 or body % dtype.itemsize:
                        continue
                    n = body // dtype.itemsize
                    if int.from_bytes(payload[prefix:h], "little") != n:
                        continue
                    if prefix == 2 and (payload[0] != 1 or payload[1] != {1: 0, 2: 2, 4: 4, 8: 6}[width]):   # ndim 1; uint8/16/32/64 codes
                        continue
                    return np.frombuffer(payload, dtype, n, offset=h).astype(np.int32, copy=False)
            if ":" in rest and ln % dtype.itemsize == 0:      # static shape in the encoding string: no header at all
                return np.frombuffer(payload, dtype, ln // dtype.itemsize).astype(np.int32, copy=False)
            raise ValueError(f"cannot locate the token payload in a {ln}-byte {enc} column")
        raise KeyError(f"column {self.column!r} not in {names}")

    def __getitem__(self, idx: int) -> np.ndarray:
        if not 0 <= idx < len(self):
            raise IndexError(idx)
        s = int(np.searchsorted(self._starts, idx, side="right") - 1)
        blob, offsets = self._load(s)
        j = idx - int(self._starts[s])
        return self._decode(blob[int(offsets[j]):int(offsets[j + 1])], self._meta[s])


class MDSWriter:
    """Token shards in mosaicml-streaming's MDS layout (``columns={"tokens": "ndarray:int32"}``, optional zstd) — what the reference's
    converter writes (ref: photon/dataset/convert_dataset_hf.py:323-327) — so a corpus converted HERE can be read by the reference's
    loaders too. Same context-manager interface as :class:`ShardWriter`.

    A raw shard is ``uint32 n | uint32 offsets[n+1] | json config | samples`` (absolute offsets); a sample is ``uint32 size`` of every
    variable-size column followed by the payloads; the ``ndarray:int32`` payload is ``uint8 ndim | uint8 code of the dims' integer
    type | dims | C-order data`` (codes: uint8 0, uint16 2, uint32 4, uint64 6). ``index.json`` (version 2) lists the shards with
    their raw / compressed basenames and byte counts. Written from the format description (no mosaicml-streaming in this
    environment); :class:`MDSReader` reads it back."""

    def __init__(self, out_dir: str | os.PathLike, seq_len: int, shard_samples: int | None = None, compression: str | None = "zstd",
                 size_limit: int = 1 << 26) -> None:
        if compression not in (None, "zstd"):
            raise ValueError("MDS shards are written raw or zstd-compressed")
        if compression == "zstd" and _zstd() is None:
            raise RuntimeError("zstd shards need pyarrow's codec (no zstandard module in this image)")
        self.out = Path(out_dir)
        self.out.mkdir(parents=True, exist_ok=True)
        self.seq_len, self.compression, self.size_limit = int(seq_len), compression, int(size_limit)
        self.shard_samples = int(shard_samples) if shard_samples else None
        self._samples: list[bytes] = []
        self._bytes = 0
        self._shards: list[dict[str, Any]] = []
        self._config = {"version": 2, "format": "mds", "compression": compression, "hashes": [], "size_limit": self.size_limit,
                        "column_names": ["tokens"], "column_encodings": ["ndarray:int32"], "column_sizes": [None]}

    @staticmethod
    def _payload(a: np.ndarray) -> bytes:
        n = int(a.shape[0])
        code, dt = (0, np.uint8) if n < 1 << 8 else (2, np.uint16) if n < 1 << 16 else (4, np.uint32) if n < 1 << 32 else (6, np.uint64)
        return bytes([1, code]) + np.array([n], dt).tobytes() + a.tobytes()

    def write(self, tokens: np.ndarray) -> None:
        a = np.ascontiguousarray(np.asarray(tokens, dtype=np.int32).reshape(-1))
        if a.size != self.seq_len:
            raise ValueError(f"sample has {a.size} tokens, expected {self.seq_len}")
        body = self._payload(a)
        sample = np.uint32(len(body)).tobytes() + body
        full = (self.shard_samples is not None and len(self._samples) >= self.shard_samples) or \
               (self._samples and self._bytes + len(sample) + 4 * (len(self._samples) + 3) > self.size_limit)
        if full:
            self._flush()
        self._samples.append(sample)
        self._bytes += len(sample)

    def write_many(self, rows: Iterable[np.ndarray]) -> None:
        for r in rows:
            self.write(r)

    def _flush(self) -> None:
        if not self._samples:
            return
        cfg = json.dumps(self._config, sort_keys=True).encode()
        n = len(self._samples)
        offsets = np.concatenate([[0], np.cumsum([len(x) for x in self._samples])]).astype(np.uint32)
        offsets += np.uint32(4 + 4 * (n + 1) + len(cfg))
        raw = np.uint32(n).tobytes() + offsets.tobytes() + cfg + b"".join(self._samples)
        base = f"shard.{len(self._shards):05d}.mds"
        entry: dict[str, Any] = {**self._config, "samples": n, "raw_data": {"basename": base, "bytes": len(raw), "hashes": {}}, "zip_data": None}
        if self.compression == "zstd":
            z = _zstd().compress(raw, asbytes=True)
            (self.out / (base + ".zstd")).write_bytes(z)
            entry["zip_data"] = {"basename": base + ".zstd", "bytes": len(z), "hashes": {}}
        else:
            (self.out / base).write_bytes(raw)
        self._shards.append(entry)
        self._samples, self._bytes = [], 0

    def finish(self) -> dict[str, Any]:
        self._flush()
        index = {"version": 2, "shards": self._shards}
        (self.out / INDEX_NAME).write_text(json.dumps(index, sort_keys=True))
        return index

    def __enter__(self) -> "MDSWriter":
        return self

    def __exit__(self, *exc: Any) -> None:
        self.finish()


def open_shard_dir(directory: str | os.PathLike, validate_hash: bool = False) -> Any:
    """:class:`ShardReader` for our own format, :class:`MDSReader` for a mosaicml-streaming directory."""
    index = json.loads((Path(directory) / INDEX_NAME).read_text())
    if index.get("format") == FORMAT:
        return ShardReader(directory, validate_hash=validate_hash)
    return MDSReader(directory, validate_hash=validate_hash)
