from photon_b200.strategy.aggregation import StreamingMean, aggregate_inplace, weighted_average, weighted_loss_avg
from photon_b200.strategy.dispatcher import dispatch_strategy
from photon_b200.strategy.strategies import (FedAdam, FedAvgEfficient, FedMom, FedNesterov, FedYogi, ServerStrategy,
                                            server_opt_step)
from photon_b200.strategy.aggregation import (aggregate_cumulative_average, aggregate_parameters,  # noqa: E402
                                             parameters_to_ndarrays_gen)
from photon_b200.strategy.utils import initialize_strategy  # noqa: E402
