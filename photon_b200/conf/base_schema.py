"""``photon.conf.base_schema`` of the reference (ref: photon/conf/base_schema.py): the same class names, defined in
:mod:`photon_b200.config.schema` (pydantic models instead of OmegaConf structured configs)."""
from photon_b200.config.schema import (FL, BackendKwargs, BaseConfig, Centralized, ClientConfig, CommStack, Dataset, Fleet, Kernels,  # noqa: F401
                                       LLMConfig, Photon, S3CommConfig, StrategyKWArgs, StrategyName, Wandb, WandbSetup, register_config,
                                       validate_config)
