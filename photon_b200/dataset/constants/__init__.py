"""Dataset split tables for the converter: C4 English + the mC4 languages the reference ships
(ref: photon/dataset/constants/__init__.py:20-34, constants/mc4.py:30-77). Built programmatically:
every language has the same six folder splits with the same truncation counts."""
from __future__ import annotations

from dataclasses import dataclass, field

C4_PATH, MC4_PATH = "allenai/c4", "allenai/c4"
LANGUAGES = ("en", "it", "zh", "ms", "ur", "sw", "la", "sr", "es", "de", "el", "ru", "hi")
# folder_split -> (hf_split, truncated_samples)
SPLIT_TABLE = {"train": ("train", None), "train_small": ("train", 100_000), "val": ("validation", None),
               "val_small": ("validation", 10_000), "val_xsmall": ("validation", 3_000), "val_xxsmall": ("validation", 100)}


@dataclass(frozen=True)
class DataSplitConstants:
    path: str
    name: str
    split: str
    folder_split: str
    truncated_samples: int | None


@dataclass(frozen=True)
class DatasetConstants:
    splits: dict[str, DataSplitConstants] = field(default_factory=dict)

    def __iter__(self):  # noqa: ANN204
        return iter(self.splits.values())


def _lang(lang: str) -> DatasetConstants:
    return DatasetConstants({fs: DataSplitConstants(C4_PATH if lang == "en" else MC4_PATH, lang, hf, fs, trunc)
                             for fs, (hf, trunc) in SPLIT_TABLE.items()})


DATASETS_CONSTANTS: dict[str, DatasetConstants] = {f"c4_{lang}": _lang(lang) for lang in LANGUAGES}
