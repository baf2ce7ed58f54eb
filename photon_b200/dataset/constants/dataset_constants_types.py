"""Split names, the concatenation mode switch and the two record types of the converter's dataset tables
(ref: photon/dataset/constants/dataset_constants_types.py:7-50)."""
from __future__ import annotations

from dataclasses import dataclass, field
from enum import Enum
from typing import Iterator

TRAIN_CONSTANT, TRAIN_SMALL_CONSTANT = "train", "train_small"
VALIDATION_CONSTANT = "validation"                      # the Hugging Face split AND the table key of the full validation set ...
VAL_CONSTANT = "val"                                    # ... which is written to the folder "val"
VAL_SMALL_CONSTANT, VAL_XSMALL_CONSTANT, VAL_XXSMALL_CONSTANT = "val_small", "val_xsmall", "val_xxsmall"


class ConcatMode(Enum):
    NO_CONCAT = "NO_CONCAT"
    CONCAT_TOKENS = "CONCAT_TOKENS"


@dataclass(frozen=True)
class DataSplitConstants:
    path: str                       # Hugging Face dataset path
    name: str                       # dataset configuration (the language)
    split: str                      # Hugging Face split the samples come from
    folder_split: str               # folder (and stream split) the shards are written to
    truncated_samples: int | None   # take only the first n samples


@dataclass(frozen=True)
class DatasetConstants:
    splits: dict[str, DataSplitConstants] = field(default_factory=dict)

    def __iter__(self) -> Iterator[DataSplitConstants]:
        return iter(self.splits.values())
