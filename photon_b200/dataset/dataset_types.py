"""Small shared types of the dataset tooling (ref: photon/dataset/dataset_types.py)."""
from __future__ import annotations

from enum import Enum


class ConcatMode(str, Enum):
    NO_CONCAT = "NO_CONCAT"
    CONCAT_TOKENS = "CONCAT_TOKENS"


from dataclasses import dataclass  # noqa: E402
from typing import Any  # noqa: E402


@dataclass
class TokenizersCouple:
    """The two tokenizers of a re-tokenisation: samples are DEcoded with the one they were written with and ENcoded with the new one
    (ref: photon/dataset/dataset_types.py:8-13)."""

    encode_tokenizer: Any
    decode_tokenizer: Any
