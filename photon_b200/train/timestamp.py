"""Training clock: counters + ``"500ba"``-style time strings.

Re-implements the slice of Composer's ``Timestamp``/``Time`` the reference
depends on (ref: photon/clients/llm_client_functions.py:164-174 sets
``timestamp.batch`` from ``server_steps_cumulative``; photon/utils.py:41-53
ships ``local_timestamp`` as a literal-evaluable dict inside ``ClientState``).
Units: ``ep`` epoch, ``ba`` batch, ``sp`` sample, ``tok`` token, ``dur``
fraction of ``max_duration``.
"""
from __future__ import annotations

import re
from dataclasses import asdict, dataclass, fields
from typing import Any

_UNITS = ("ep", "ba", "sp", "tok", "dur")
_TIME_RE = re.compile(r"^\s*([-+]?[0-9]*\.?[0-9]+(?:[eE][-+]?[0-9]+)?)\s*(ep|ba|sp|tok|dur)\s*$")


@dataclass(frozen=True)
class Time:
    value: float
    unit: str

    @classmethod
    def parse(cls, spec: "str | int | Time", default_unit: str = "ba") -> "Time":
        if isinstance(spec, Time):
            return spec
        if isinstance(spec, (int, float)):
            return cls(float(spec), default_unit)
        m = _TIME_RE.match(str(spec))
        if not m:
            raise ValueError(f"invalid time string {spec!r} (expected e.g. '500ba', '1ep', '1e9tok', '0.5dur')")
        v, u = float(m.group(1)), m.group(2)
        return cls(v, u)

    def __str__(self) -> str:
        v = int(self.value) if float(self.value).is_integer() else self.value
        return f"{v}{self.unit}"

    def to_batches(self, *, max_duration: "Time | None" = None, samples_per_batch: int | None = None,
                   tokens_per_batch: int | None = None, batches_per_epoch: int | None = None) -> int:
        """Convert to a whole number of batches given the run geometry."""
        if self.unit == "ba":
            return int(self.value)
        if self.unit == "dur":
            if max_duration is None:
                raise ValueError("'dur' needs max_duration")
            return int(self.value * max_duration.to_batches(samples_per_batch=samples_per_batch,
                                                            tokens_per_batch=tokens_per_batch,
                                                            batches_per_epoch=batches_per_epoch))
        if self.unit == "sp":
            if not samples_per_batch:
                raise ValueError("'sp' needs samples_per_batch")
            return int(self.value // samples_per_batch)
        if self.unit == "tok":
            if not tokens_per_batch:
                raise ValueError("'tok' needs tokens_per_batch")
            return int(self.value // tokens_per_batch)
        if self.unit == "ep":
            if not batches_per_epoch:
                raise ValueError("'ep' needs a sized dataloader")
            return int(self.value * batches_per_epoch)
        raise ValueError(self.unit)


@dataclass
class Timestamp:
    epoch: int = 0
    batch: int = 0
    sample: int = 0
    token: int = 0
    batch_in_epoch: int = 0
    sample_in_epoch: int = 0
    token_in_epoch: int = 0
    iteration: int = 0
    total_wct_s: float = 0.0

    def advance_batch(self, samples: int, tokens: int, wct_s: float = 0.0) -> None:
        self.batch += 1
        self.batch_in_epoch += 1
        self.sample += samples
        self.sample_in_epoch += samples
        self.token += tokens
        self.token_in_epoch += tokens
        self.total_wct_s += wct_s

    def advance_epoch(self) -> None:
        self.epoch += 1
        self.batch_in_epoch = self.sample_in_epoch = self.token_in_epoch = 0

    def get(self, unit: str) -> float:
        return {"ep": self.epoch, "ba": self.batch, "sp": self.sample, "tok": self.token}[unit]

    def state_dict(self) -> dict[str, Any]:
        return asdict(self)

    def load_state_dict(self, sd: dict[str, Any]) -> None:
        names = {f.name for f in fields(self)}
        for k, v in sd.items():
            if k in names:
                setattr(self, k, type(getattr(self, k))(v))

    def reset(self) -> None:
        self.load_state_dict(asdict(Timestamp()))

    def copy(self) -> "Timestamp":
        return Timestamp(**asdict(self))
