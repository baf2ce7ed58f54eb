"""LR schedules (multiplicative factor on the base LR, stepped EVERY batch and
indexed by the *global* batch counter so a federated client resumes the
schedule where the federation left off; ref: photon/clients/
llm_client_functions.py:172-174, photon/conf/llm_config/mpt-125m.yaml:50-56,
scripts/cen_125m_example.sh:51-54; closed forms in SURVEY App. C)."""
from __future__ import annotations

import math
from typing import Any, Callable

from photon_b200.train.timestamp import Time

Scheduler = Callable[[int], float]


def _warm(t: int, t_warmup: int) -> float | None:
    if t_warmup > 0 and t < t_warmup:
        return t / t_warmup
    return None


def cosine_with_warmup(t_warmup: int, t_max: int, alpha_f: float = 0.0) -> Scheduler:
    def f(t: int) -> float:
        w = _warm(t, t_warmup)
        if w is not None:
            return w
        frac = min(1.0, max(0.0, (t - t_warmup) / max(1, t_max - t_warmup)))
        return alpha_f + (1.0 - alpha_f) * 0.5 * (1.0 + math.cos(math.pi * frac))
    return f


def linear_decay_with_warmup(t_warmup: int, t_max: int, alpha_f: float = 0.0) -> Scheduler:
    def f(t: int) -> float:
        w = _warm(t, t_warmup)
        if w is not None:
            return w
        frac = min(1.0, max(0.0, (t - t_warmup) / max(1, t_max - t_warmup)))
        return 1.0 + (alpha_f - 1.0) * frac
    return f


def constant_with_warmup(t_warmup: int, t_max: int | None = None, alpha: float = 1.0) -> Scheduler:
    def f(t: int) -> float:
        w = _warm(t, t_warmup)
        return alpha * (w if w is not None else 1.0)
    return f


def constant_with_sqrt_cooldown_with_warmup(t_warmup: int, t_max: int, t_cooldown: int,
                                            alpha_f: float = 0.0) -> Scheduler:
    """Warmup–stable–decay with a 1-sqrt cooldown over the last ``t_cooldown`` batches."""
    def f(t: int) -> float:
        w = _warm(t, t_warmup)
        if w is not None:
            return w
        start = t_max - t_cooldown
        if t <= start or t_cooldown <= 0:
            return 1.0
        frac = min(1.0, (t - start) / t_cooldown)
        return alpha_f + (1.0 - alpha_f) * (1.0 - math.sqrt(frac))
    return f


def build_scheduler(cfg: dict[str, Any] | None, *, max_duration: Time | None = None,
                    samples_per_batch: int | None = None, tokens_per_batch: int | None = None,
                    batches_per_epoch: int | None = None) -> Scheduler:
    """``cfg`` is ``llm_config.scheduler.schedulers.lr`` (or llm-foundry's flat ``scheduler``)."""
    if not cfg:
        return lambda t: 1.0
    if "schedulers" in cfg:
        cfg = cfg["schedulers"].get("lr", next(iter(cfg["schedulers"].values())))
    cfg = dict(cfg)
    name = cfg.pop("name")

    def nb(key: str, default: Any = None) -> int:
        v = cfg.get(key, default)
        if v is None:
            raise ValueError(f"scheduler '{name}' needs '{key}'")
        return Time.parse(v).to_batches(max_duration=max_duration, samples_per_batch=samples_per_batch,
                                        tokens_per_batch=tokens_per_batch, batches_per_epoch=batches_per_epoch)

    t_max_default = str(max_duration) if max_duration is not None else None
    if name == "cosine_with_warmup":
        return cosine_with_warmup(nb("t_warmup", "0ba"), nb("t_max", t_max_default), float(cfg.get("alpha_f", 0.0)))
    if name == "linear_decay_with_warmup":
        return linear_decay_with_warmup(nb("t_warmup", "0ba"), nb("t_max", t_max_default), float(cfg.get("alpha_f", 0.0)))
    if name == "constant_with_warmup":
        return constant_with_warmup(nb("t_warmup", "0ba"), None, float(cfg.get("alpha", 1.0)))
    if name == "constant_with_sqrt_cooldown_with_warmup":
        return constant_with_sqrt_cooldown_with_warmup(nb("t_warmup", "0ba"), nb("t_max", t_max_default),
                                                       nb("t_cooldown", "0ba"), float(cfg.get("alpha_f", 0.0)))
    raise ValueError(f"unknown scheduler '{name}'")
