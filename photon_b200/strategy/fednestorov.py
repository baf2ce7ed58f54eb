"""``FedNesterov`` under the reference's module path (ref: photon/strategy/fednestorov.py:313-334). All five server optimizers share one
implementation — ``photon_b200.strategy.strategies`` (host / oracle path) and ``csrc/comm.cu`` (fused NVLink round kernel)."""
from photon_b200.strategy.strategies import FedNesterov, server_opt_step  # noqa: F401

__all__ = ["FedNesterov"]
