"""Small shared types of the dataset tooling (ref: photon/dataset/dataset_types.py)."""
from __future__ import annotations

from enum import Enum


class ConcatMode(str, Enum):
    NO_CONCAT = "NO_CONCAT"
    CONCAT_TOKENS = "CONCAT_TOKENS"
