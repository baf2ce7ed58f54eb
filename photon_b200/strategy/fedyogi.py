"""``FedYogi`` under the reference's module path (ref: photon/strategy/fedyogi.py:295-322). All five server optimizers share one
implementation — ``photon_b200.strategy.strategies`` (host / oracle path) and ``csrc/comm.cu`` (fused NVLink round kernel)."""
from photon_b200.strategy.strategies import FedYogi, server_opt_step  # noqa: F401

__all__ = ["FedYogi"]
