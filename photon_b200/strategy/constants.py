"""Checkpoint file stems == strategy state keys (ref: photon/strategy/constants.py:4-6)."""
SERVER_PARAMETERS_KEY = "current_server_parameters"
MOMENTUM_KEY = "current_momentum_vector"
SECOND_MOMENTUM_KEY = "current_second_momentum_vector"
