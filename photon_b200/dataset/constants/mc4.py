"""C4 English and the twelve mC4 languages the reference ships (ref: photon/dataset/constants/mc4.py:14-339).

English has six entries — the two full splits plus four truncated ones (``train_small`` = first 100 k training samples,
``val_small`` / ``val_xsmall`` / ``val_xxsmall`` = first 10 k / 3 k / 100 validation samples); every other language has the two
full splits only. Table keys are the reference's (``validation`` is the key of the full validation set, its folder is ``val``).
The tables are generated from two small specs instead of thirteen literal blocks; ``c4_<lang>_constants`` names are kept."""
from __future__ import annotations

from photon_b200.dataset.dataset_types import (TRAIN_CONSTANT, TRAIN_SMALL_CONSTANT, VAL_CONSTANT, VAL_SMALL_CONSTANT,
                                                                   VAL_XSMALL_CONSTANT, VAL_XXSMALL_CONSTANT, VALIDATION_CONSTANT,
                                                                   DatasetConstants, DataSplitConstants)

C4_PATH = "allenai/c4"
LANGUAGE_CODES = {"ENGLISH": "en", "SERBIAN": "sr", "LATIN": "la", "SWAHILI": "sw", "URDU": "ur", "MALAY": "ms", "CHINESE": "zh",
                  "ITALIAN": "it", "SPANISH": "es", "GERMAN": "de", "GREEK": "el", "RUSSIAN": "ru", "HINDI": "hi"}
globals().update({f"{k}_CONSTANT": v for k, v in LANGUAGE_CODES.items()})      # ENGLISH_CONSTANT = "en", ...

# table key -> (hf split, folder, truncation)
_FULL = {TRAIN_CONSTANT: (TRAIN_CONSTANT, TRAIN_CONSTANT, None), VALIDATION_CONSTANT: (VALIDATION_CONSTANT, VAL_CONSTANT, None)}
_ENGLISH_EXTRA = {TRAIN_SMALL_CONSTANT: (TRAIN_CONSTANT, TRAIN_SMALL_CONSTANT, 100_000),
                  VAL_SMALL_CONSTANT: (VALIDATION_CONSTANT, VAL_SMALL_CONSTANT, 10_000),
                  VAL_XSMALL_CONSTANT: (VALIDATION_CONSTANT, VAL_XSMALL_CONSTANT, 3_000),
                  VAL_XXSMALL_CONSTANT: (VALIDATION_CONSTANT, VAL_XXSMALL_CONSTANT, 100)}


def _table(lang: str) -> DatasetConstants:
    spec = {**_FULL, **(_ENGLISH_EXTRA if lang == "en" else {})}
    order = [TRAIN_CONSTANT, TRAIN_SMALL_CONSTANT, VALIDATION_CONSTANT, VAL_SMALL_CONSTANT, VAL_XSMALL_CONSTANT, VAL_XXSMALL_CONSTANT]
    return DatasetConstants({k: DataSplitConstants(C4_PATH, lang, *spec[k]) for k in order if k in spec})


ALL_CONSTANTS: dict[str, DatasetConstants] = {lang: _table(lang) for lang in LANGUAGE_CODES.values()}
globals().update({f"c4_{lang}_constants": t for lang, t in ALL_CONSTANTS.items()})     # c4_en_constants, c4_sr_constants, ...
