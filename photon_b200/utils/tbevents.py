"""TensorBoard event files without the tensorboard package.

Composer's TensorBoard logger (the reference's ``loggers.tensorboard``) needs ``tensorboard`` importable; when it is not, scalars are
written here in the same on-disk format, so ``tensorboard --logdir`` on any machine that has it reads the run: a TFRecord stream
(``uint64 length · masked CRC32C(length) · payload · masked CRC32C(payload)``) of ``Event`` protocol-buffer messages, hand-encoded —
``Event{wall_time=1:double, step=2:int64, file_version=3:string | summary=5:Summary}``, ``Summary{value=1:repeated Value}``,
``Value{tag=1:string, simple_value=2:float}``. The first record carries ``file_version = "brain.Event:2"``.
"""
from __future__ import annotations

import os
import socket
import struct
import time
from pathlib import Path
from typing import Iterator

_POLY = 0x82F63B78      # CRC-32C (Castagnoli), reflected
_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ _POLY if _c & 1 else _c >> 1
    _TABLE.append(_c)


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c = _TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _masked(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _varint(n: int) -> bytes:
    n &= (1 << 64) - 1          # int64 two's complement, like protobuf
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _len_field(field: int, payload: bytes) -> bytes:
    return _varint(field << 3 | 2) + _varint(len(payload)) + payload


def encode_event(wall_time: float, step: int, scalars: dict[str, float] | None = None, file_version: str | None = None) -> bytes:
    ev = _varint(1 << 3 | 1) + struct.pack("<d", wall_time) + _varint(2 << 3 | 0) + _varint(step)
    if file_version is not None:
        ev += _len_field(3, file_version.encode())
    if scalars:
        summary = b"".join(_len_field(1, _len_field(1, tag.encode()) + _varint(2 << 3 | 5) + struct.pack("<f", float(v)))
                           for tag, v in scalars.items())
        ev += _len_field(5, summary)
    return ev


def record(payload: bytes) -> bytes:
    head = struct.pack("<Q", len(payload))
    return head + struct.pack("<I", _masked(head)) + payload + struct.pack("<I", _masked(payload))


class EventFileWriter:
    def __init__(self, log_dir: str | os.PathLike, flush_secs: float = 10.0) -> None:
        d = Path(log_dir)
        d.mkdir(parents=True, exist_ok=True)
        self.path = d / f"events.out.tfevents.{int(time.time())}.{socket.gethostname()}.{os.getpid()}"
        self._f = open(self.path, "ab")
        self._flush_secs, self._last = float(flush_secs), time.time()
        self._f.write(record(encode_event(time.time(), 0, file_version="brain.Event:2")))
        self._f.flush()

    def add_scalars(self, scalars: dict[str, float], step: int) -> None:
        self._f.write(record(encode_event(time.time(), int(step), scalars)))
        if time.time() - self._last >= self._flush_secs:
            self.flush()

    def add_scalar(self, tag: str, value: float, step: int) -> None:
        self.add_scalars({tag: value}, step)

    def flush(self) -> None:
        self._f.flush()
        self._last = time.time()

    def close(self) -> None:
        if not self._f.closed:
            self._f.flush()
            self._f.close()


# ----------------------------------------------------------------------------------------------- reader (tests, offline inspection)
def _read_varint(buf: bytes, i: int) -> tuple[int, int]:
    n = shift = 0
    while True:
        b = buf[i]
        i += 1
        n |= (b & 0x7F) << shift
        shift += 7
        if not b & 0x80:
            return n, i


def _fields(buf: bytes) -> Iterator[tuple[int, int, bytes | int]]:
    i = 0
    while i < len(buf):
        key, i = _read_varint(buf, i)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _read_varint(buf, i)
            yield field, wt, v
        elif wt == 1:
            yield field, wt, buf[i: i + 8]
            i += 8
        elif wt == 5:
            yield field, wt, buf[i: i + 4]
            i += 4
        elif wt == 2:
            n, i = _read_varint(buf, i)
            yield field, wt, buf[i: i + n]
            i += n
        else:
            raise ValueError(f"unsupported wire type {wt}")


def read_events(path: str | os.PathLike) -> list[dict]:
    """Decode an event file written by anybody: ``[{wall_time, step, file_version?, scalars: {tag: value}}]``; checks both CRCs."""
    data = Path(path).read_bytes()
    out, i = [], 0
    while i < len(data):
        head = data[i: i + 8]
        (n,) = struct.unpack("<Q", head)
        if struct.unpack("<I", data[i + 8: i + 12])[0] != _masked(head):
            raise ValueError(f"corrupt length CRC at byte {i}")
        payload = data[i + 12: i + 12 + n]
        if struct.unpack("<I", data[i + 12 + n: i + 16 + n])[0] != _masked(payload):
            raise ValueError(f"corrupt payload CRC at byte {i}")
        i += 16 + n
        ev: dict = {"scalars": {}}
        for field, _wt, v in _fields(payload):
            if field == 1:
                ev["wall_time"] = struct.unpack("<d", v)[0]      # type: ignore[arg-type]
            elif field == 2:
                ev["step"] = v if v < 1 << 63 else v - (1 << 64)  # type: ignore[operator]
            elif field == 3:
                ev["file_version"] = bytes(v).decode()           # type: ignore[arg-type]
            elif field == 5:
                for f2, _w2, val in _fields(v):                  # type: ignore[arg-type]
                    if f2 == 1:
                        tag, x = "", 0.0
                        for f3, _w3, vv in _fields(val):         # type: ignore[arg-type]
                            if f3 == 1:
                                tag = bytes(vv).decode()         # type: ignore[arg-type]
                            elif f3 == 2:
                                x = struct.unpack("<f", vv)[0]   # type: ignore[arg-type]
                        ev["scalars"][tag] = x
        out.append(ev)
    return out
