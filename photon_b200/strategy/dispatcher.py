"""``cfg.fl.strategy_name`` → strategy instance (ref: photon/strategy/dispatcher.py:44-165)."""
from __future__ import annotations

from typing import Any

from photon_b200.config.schema import StrategyName
from photon_b200.strategy.metrics import FedSimpleNoiseScale
from photon_b200.strategy.strategies import FedAdam, FedAvgEfficient, FedMom, FedNesterov, FedYogi, ServerStrategy


def dispatch_strategy(cfg: Any) -> ServerStrategy:
    fl = cfg["fl"] if isinstance(cfg, dict) else cfg.fl
    name = StrategyName.parse(fl["strategy_name"])
    kw = dict(fl.get("strategy_kwargs") or {})
    common: dict[str, Any] = dict(n_clients_per_round=int(fl["n_clients_per_round"]),
                                  reference_sign_compat=bool(fl.get("reference_sign_compat", False)))
    # the reference only wires the noise-scale callback into nestorov / fedavg (dispatcher.py:69-71,116-118); the estimate
    # only needs per-client and aggregate pseudo-gradient norms, so every strategy gets it here
    if fl.get("use_noise_scale_metric"):
        common["metrics_callback"] = FedSimpleNoiseScale(float(fl.get("noise_scale_beta", 0.99)))
    for extra in ("track_norms", "track_inplace_aggregation", "scaling_fn"):
        if extra in kw:
            common[extra] = kw.pop(extra)
    if name in (StrategyName.NESTOROV, StrategyName.FEDMOM):
        for req in ("server_learning_rate", "server_momentum"):
            if kw.get(req) is None:
                raise ValueError(f"fl.strategy_kwargs.{req} is required for {name.value}")
    if name == StrategyName.NESTOROV:
        return FedNesterov(kw["server_learning_rate"], kw["server_momentum"], **common)
    if name == StrategyName.FEDMOM:
        return FedMom(kw["server_learning_rate"], kw["server_momentum"], **common)
    if name == StrategyName.FEDAVG:  # the reference pins η=1.0 here (dispatcher.py:99-120)
        return FedAvgEfficient(1.0, **common)
    adam_kw = {k: kw[k] for k in ("eta", "beta_1", "beta_2", "tau") if kw.get(k) is not None}
    if name == StrategyName.FEDADAM:
        return FedAdam(**adam_kw, **common)
    if name == StrategyName.FEDYOGI:
        return FedYogi(**adam_kw, **common)
    raise ValueError("Unknown strategy")
