"""Dataset split tables for the converter: C4 English + the mC4 languages the reference ships
(ref: photon/dataset/constants/__init__.py:20-34). ``DATASETS_CONSTANTS["c4_en"].splits["val_xxsmall"]`` etc.;
``val`` is accepted as an alias of the reference's ``validation`` key (the folder it is written to)."""
from __future__ import annotations

from photon_b200.dataset.dataset_types import ConcatMode, DatasetConstants, DataSplitConstants  # noqa: F401
from photon_b200.dataset.constants.mc4 import ALL_CONSTANTS, C4_PATH  # noqa: F401

# the reference's order
DATASETS_CONSTANTS: dict[str, DatasetConstants] = {f"c4_{lang}": ALL_CONSTANTS[lang]
                                                   for lang in ("en", "it", "zh", "ms", "ur", "sw", "la", "sr", "es", "de", "el", "ru", "hi")}


def resolve_split(dataset: str, split: str) -> DataSplitConstants:
    """Table lookup with the ``val`` → ``validation`` alias and a message that lists what exists."""
    table = DATASETS_CONSTANTS[dataset].splits
    key = "validation" if split == "val" and "val" not in table else split
    if key not in table:
        raise KeyError(f"dataset {dataset} has no split '{split}' (have {sorted(table)})")
    return table[key]
