"""The reference's ``FedAvgWithConfig`` mixin (Flower ``FedAvg`` + config-function plumbing, ref:
photon/strategy/strategy_with_cfg.py:88-162) is the base class here: ``ServerStrategy`` carries the fit / evaluate
metric-aggregation functions, the failure policy lives in ``server/fit_utils.py`` and the per-round config functions in
``clients/configs.py`` (``get_photon_fit_config_fn`` / ``get_photon_evaluate_config_fn``)."""
from photon_b200.clients.configs import get_photon_evaluate_config_fn, get_photon_fit_config_fn  # noqa: F401
from photon_b200.strategy.strategies import ServerStrategy

FedAvgWithConfig = ServerStrategy

__all__ = ["FedAvgWithConfig", "ServerStrategy"]
