"""Phase tracing: Chrome-trace (``chrome://tracing`` / Perfetto) spans + NVTX ranges + optional CUDA-event timing.

The reference wires Composer's ``Profiler`` + ``JSONTraceHandler`` (off by default; ref:
photon/clients/trainer_utils.py:1456-1482) and otherwise reports wall-clock spans as metrics
(SURVEY §5.1).  Here one tiny tracer serves both: every ``with tracer.span("fit_round")`` becomes
a complete event in a Chrome trace, an NVTX range for nsys/ncu, and — when ``device=True`` — a pair
of CUDA events whose elapsed time (device time on the launching stream) is attached as ``args.device_ms``.
Enable globally with ``PHOTON_TRACE=/path/trace.json`` (written at exit / ``flush()``).
"""
from __future__ import annotations

import atexit
import contextlib
import json
import os
import threading
import time
from typing import Any, Iterator

import torch


class PhaseTracer:
    def __init__(self, path: str | None = None, rank: int = 0, enabled: bool | None = None) -> None:
        self.path = path or os.environ.get("PHOTON_TRACE")
        self.enabled = bool(self.path) if enabled is None else enabled
        self.rank = rank
        self.events: list[dict[str, Any]] = []
        self._pending: list[tuple[dict[str, Any], Any, Any]] = []
        self._lock = threading.Lock()
        if self.enabled and self.path:
            atexit.register(self.flush)

    @contextlib.contextmanager
    def span(self, name: str, cat: str = "phase", device: bool = False, **args: Any) -> Iterator[dict[str, Any]]:
        if not self.enabled:
            yield {}
            return
        cuda = device and torch.cuda.is_available()
        if cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.nvtx.range_push(name)
            e0.record()
        t0 = time.perf_counter_ns()
        ev: dict[str, Any] = {"name": name, "cat": cat, "ph": "X", "pid": self.rank, "tid": threading.get_ident() % 100000,
                              "ts": t0 / 1e3, "args": dict(args)}
        try:
            yield ev["args"]
        finally:
            ev["dur"] = (time.perf_counter_ns() - t0) / 1e3
            if cuda:
                e1.record()
                torch.cuda.nvtx.range_pop()
                self._pending.append((ev, e0, e1))
            with self._lock:
                self.events.append(ev)

    def instant(self, name: str, **args: Any) -> None:
        if self.enabled:
            self.events.append({"name": name, "ph": "i", "s": "p", "pid": self.rank, "tid": 0, "ts": time.perf_counter_ns() / 1e3, "args": args})

    def resolve_device_times(self) -> None:
        """Attach device-side durations (needs the events to have completed → synchronises)."""
        if self._pending:
            torch.cuda.synchronize()
            for ev, e0, e1 in self._pending:
                ev["args"]["device_ms"] = e0.elapsed_time(e1)
            self._pending = []

    def flush(self) -> None:
        if not (self.enabled and self.path):
            return
        self.resolve_device_times()
        path = self.path if self.rank == 0 else f"{self.path}.rank{self.rank}"
        with open(path, "w") as f:
            json.dump({"traceEvents": self.events, "displayTimeUnit": "ms"}, f)


_GLOBAL: PhaseTracer | None = None


def tracer(rank: int | None = None) -> PhaseTracer:
    global _GLOBAL
    if _GLOBAL is None:
        _GLOBAL = PhaseTracer(rank=int(os.environ.get("RANK", "0")) if rank is None else rank)
    return _GLOBAL
