"""photon_b200 — a B200-native federated LLM pre-training engine.

Capabilities follow relogu/photon (federated LocalSGD/FedOpt pre-training of
MPT decoder-only LLMs, centralised DDP training, Hydra-style config surface,
shm hand-off, server/client checkpoints) with the hot paths re-designed for
Blackwell: hand-written sm_100a kernels (tcgen05/TMEM/TMA GEMMs, fused
norm/loss/optimizer kernels) and in-kernel NVLink collectives for the round
reduce/broadcast and the DDP gradient all-reduce.
"""
__version__ = "0.1.0"
