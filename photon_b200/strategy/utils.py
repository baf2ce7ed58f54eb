"""Strategy (re)initialisation helper (ref: photon/strategy/utils.py:13-72)."""
from __future__ import annotations

from typing import Any, Sequence

import numpy as np
import torch

from photon_b200.strategy.strategies import ServerStrategy
from photon_b200.utils.flat import FlatLayout


def _as_flat(x: Any, layout: FlatLayout | None, like: torch.Tensor | None) -> torch.Tensor | None:
    """Accept a flat tensor or the reference's per-tensor ndarray list."""
    if x is None or torch.is_tensor(x):
        return x
    if layout is None:
        raise ValueError("a FlatLayout is required to install per-tensor ndarrays")
    arrays: Sequence[np.ndarray] = list(x)
    flat = torch.zeros(layout.total, dtype=torch.float32, device=like.device if like is not None else "cpu")
    layout.from_ndarrays(flat, arrays)
    return flat


def initialize_strategy(strategy: ServerStrategy, parameters: Any, momentum_vector: Any = None,
                        second_momentum_vector: Any = None, layout: FlatLayout | None = None) -> None:
    """Hand the strategy its global model and optimizer state after construction.

    Momenta a strategy does not use are ignored; momenta it needs but did not receive raise, like the
    reference's asserts do (``ServerStrategy.initialize`` itself would zero-fill them, which is what
    ``initialize_round`` wants at round 0 but not what a resume wants)."""
    params = _as_flat(parameters, layout, None)
    if params is None:
        raise ValueError("parameters must be given")
    m = _as_flat(momentum_vector, layout, params) if strategy.n_moments >= 1 else None
    v = _as_flat(second_momentum_vector, layout, params) if strategy.n_moments >= 2 else None
    if strategy.n_moments >= 1 and m is None:
        raise ValueError("Momentum vector must be initialized")
    if strategy.n_moments >= 2 and v is None:
        raise ValueError("Second momentum vector must be initialized")
    strategy.initialize(params, m, v, layout=layout if layout is not None else strategy.layout)
