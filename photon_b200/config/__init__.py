"""Config subsystem: composer (mini-Hydra), schema, time-string helpers."""
from photon_b200.config.composer import (ConfigError, ConfigNode, apply_overrides, compose, load_config,
                                         parse_value, resolve, save_yaml, to_container)
from photon_b200.config.schema import BaseConfig, StrategyName, validate_config

__all__ = ["ConfigError", "ConfigNode", "BaseConfig", "StrategyName", "apply_overrides", "compose",
           "load_config", "parse_value", "resolve", "save_yaml", "to_container", "validate_config"]
