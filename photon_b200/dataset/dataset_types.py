"""Small shared types of the dataset tooling (ref: photon/dataset/dataset_types.py)."""
from __future__ import annotations

from dataclasses import dataclass, field
from enum import Enum
from typing import Any, Iterator


class ConcatMode(str, Enum):
    NO_CONCAT = "NO_CONCAT"
    CONCAT_TOKENS = "CONCAT_TOKENS"


@dataclass
class TokenizersCouple:
    """The two tokenizers of a re-tokenisation: samples are DEcoded with the one they were written with and ENcoded with the new one
    (ref: photon/dataset/dataset_types.py:8-13)."""

    encode_tokenizer: Any
    decode_tokenizer: Any


# ---- split names and the record types of the converter's dataset tables (the reference keeps them in
# ``photon/dataset/constants/dataset_constants_types.py``; the ``photon`` alias maps that module path here)
SPLIT_KEYS = {"TRAIN": "train", "TRAIN_SMALL": "train_small", "VALIDATION": "validation", "VAL": "val", "VAL_SMALL": "val_small",
              "VAL_XSMALL": "val_xsmall", "VAL_XXSMALL": "val_xxsmall"}
globals().update({f"{k}_CONSTANT": v for k, v in SPLIT_KEYS.items()})      # TRAIN_CONSTANT = "train", ... ("validation" is the HF split and
TRAIN_CONSTANT: str                                                          # the key of the full validation set, written to the folder "val")
TRAIN_SMALL_CONSTANT: str
VALIDATION_CONSTANT: str
VAL_CONSTANT: str
VAL_SMALL_CONSTANT: str
VAL_XSMALL_CONSTANT: str
VAL_XXSMALL_CONSTANT: str


@dataclass(frozen=True)
class DataSplitConstants:
    """One row of a dataset table: where the samples come from and where they go."""

    path: str                       # Hugging Face dataset path
    name: str                       # dataset configuration (the language)
    split: str                      # Hugging Face split the samples are read from
    folder_split: str               # folder (and stream split) the shards are written to
    truncated_samples: int | None   # only the first n samples


@dataclass(frozen=True)
class DatasetConstants:
    """All rows of one dataset, keyed by the name used on the converter's command line."""

    splits: dict[str, DataSplitConstants] = field(default_factory=dict)

    def __iter__(self) -> Iterator[DataSplitConstants]:
        return iter(self.splits.values())
