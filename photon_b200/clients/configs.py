"""Per-round runtime configs: ``FitConfig``, ``EvaluateConfig``, ``CentralizedConfig``.

Same field surface as the reference (ref: photon/clients/configs.py:55-214,
289-425, 488-573) so round logs and semantics line up.  In the reference these
cross a Flower ``ConfigsRecord`` that only carries scalars, hence everything
non-scalar is stringified and ``ast.literal_eval``-ed on the other side; the
in-memory control plane here keeps real objects but ``to_wire``/``from_wire``
reproduce the stringly form for the shm / file transports and for tests.
"""
from __future__ import annotations

import ast
from typing import Any, Callable

from pydantic import BaseModel, ConfigDict, field_validator

from photon_b200.messages import ClientState, decode_client_states, encode_client_states


def typed_field_validator(field: str, /, *field_names: str, mode: str = "after") -> Callable[[Callable[..., Any]], Callable[..., Any]]:
    """``pydantic.field_validator`` behind a typed signature, so decorated validators keep their type for the checker
    (ref: photon/clients/configs.py:18-44)."""
    return field_validator(field, *field_names, mode=mode)  # type: ignore[call-overload,no-any-return]


def _maybe_literal(v: Any) -> Any:
    if isinstance(v, str):
        try:
            return ast.literal_eval(v)
        except (ValueError, SyntaxError):
            return v
    return v


class _Wire(BaseModel):
    model_config = ConfigDict(arbitrary_types_allowed=True, extra="forbid")

    def to_wire(self) -> dict[str, Any]:
        """Scalars stay; everything else becomes ``str(obj)`` (ConfigsRecord-compatible)."""
        out: dict[str, Any] = {}
        for k, v in self.model_dump().items():
            if k == "client_state" and v is not None:
                out[k] = encode_client_states({int(c): ClientState.from_literal(s) for c, s in v.items()})
            elif isinstance(v, (bool, int, float, str)):
                out[k] = v
            else:
                out[k] = str(v)
        return out

    @classmethod
    def from_wire(cls, d: dict[str, Any]) -> Any:
        return cls(**{k: _maybe_literal(v) for k, v in d.items()})


class FitConfig(_Wire):
    cid: int | str
    server_round: int
    batch_size: int
    n_local_steps: int
    n_local_epochs: int
    reset_checkpoint: bool
    reset_optimizer: bool
    reset_dataset_state: bool
    reset_timestamp: bool
    use_unigram_metrics: bool
    allow_unigram_metrics_failures: bool
    aggregate_momenta: bool
    resize_vocab: int | None = None
    s3_comm_config: dict[str, Any] = {}
    random_layers: list[str] | None = None
    random_init_freq: int = 0
    personalized_layers: list[str] | None = None
    truly_random_init: bool = True
    frozen_layers: list[str] | None = None
    unfrozen_layers: list[str] | None = None
    split_eval: bool = False
    set_trainer_params_filter_keys: bool = True
    set_trainer_key_to_filter: str = "transformer"
    client_state: dict[int, dict[str, Any]] | None = None
    server_steps_cumulative: int | None = None

    @field_validator("client_state", mode="before")
    @classmethod
    def _cs(cls, v: Any) -> Any:
        if v is None:
            return None
        states = decode_client_states(v) if isinstance(v, str) else {int(c): (s if isinstance(s, ClientState) else ClientState.from_literal(s)) for c, s in v.items()}
        return {c: s.to_literal() for c, s in states.items()}

    def state_of(self, cid: int) -> ClientState:
        if self.client_state and int(cid) in self.client_state:
            return ClientState.from_literal(self.client_state[int(cid)])
        return ClientState()


class EvaluateConfig(_Wire):
    cid: int | str | None = None
    server_round: int
    batch_size: int
    use_unigram_metrics: bool = False
    allow_unigram_metrics_failures: bool = False
    resize_vocab: int | None = None
    s3_comm_config: dict[str, Any] = {}
    frozen_layers: list[str] | None = None
    unfrozen_layers: list[str] | None = None
    split_eval: bool = False
    set_trainer_params_filter_keys: bool = True
    set_trainer_key_to_filter: str = "transformer"
    client_state: dict[int, dict[str, Any]] | None = None
    server_steps_cumulative: int | None = None

    _cs = FitConfig.__dict__["_cs"]


class CentralizedConfig(_Wire):
    """Knobs of the non-federated entry point (ref: configs.py:488-573)."""

    store_init_model: bool = False
    store_final_model: bool = False
    stream_id: int | str | None = None
    eval_only: bool = False
    split_eval: bool = False
    reset_timestamp: bool = False
    use_unigram_metrics: bool = False
    allow_unigram_metrics_failures: bool = False
    resize_vocab: int | None = None
    frozen_layers: list[str] | None = None
    unfrozen_layers: list[str] | None = None
    pretrained_model_path: str | None = None
    wte_parameters_path: str | None = None


def get_photon_fit_config_fn(cfg: Any) -> Callable[[int, int], FitConfig]:
    """``(server_round, cid) -> FitConfig`` built from ``cfg.fl`` (ref: configs.py:217-286)."""
    fl, llm = cfg["fl"], cfg["llm_config"]

    def fn(server_round: int, cid: int, client_states: dict[int, ClientState] | None = None,
           server_steps_cumulative: int | None = None) -> FitConfig:
        return FitConfig(
            cid=cid, server_round=server_round, batch_size=int(llm["global_train_batch_size"]),
            n_local_steps=int(fl["n_local_steps"]), n_local_epochs=int(fl["n_local_epochs"]),
            reset_checkpoint=bool(fl["reset_checkpoint"]), reset_optimizer=bool(fl["reset_optimizer"]),
            reset_dataset_state=bool(fl["reset_dataset_state"]), reset_timestamp=bool(fl["reset_timestamp"]),
            use_unigram_metrics=bool(fl["use_unigram_metrics"]),
            allow_unigram_metrics_failures=bool(fl["allow_unigram_metrics_failures"]),
            aggregate_momenta=bool(fl["aggregate_momenta"]), resize_vocab=fl.get("resize_vocab"),
            s3_comm_config=dict(cfg.get("s3_comm_config") or {}), random_layers=list(fl.get("random_layers") or []),
            random_init_freq=int(fl.get("random_init_freq", 0)), personalized_layers=list(fl.get("personalized_layers") or []),
            truly_random_init=bool(fl.get("truly_random_init", True)), frozen_layers=fl.get("frozen_layers"),
            unfrozen_layers=fl.get("unfrozen_layers"), split_eval=bool(fl.get("split_eval", False)),
            set_trainer_params_filter_keys=bool(fl.get("set_trainer_params_filter_keys", True)),
            set_trainer_key_to_filter=str(fl.get("set_trainer_key_to_filter", "transformer")),
            client_state={c: s.to_literal() for c, s in (client_states or {}).items()} or None,
            server_steps_cumulative=server_steps_cumulative)

    return fn


def get_photon_evaluate_config_fn(cfg: Any) -> Callable[..., EvaluateConfig]:
    fl, llm = cfg["fl"], cfg["llm_config"]

    def fn(server_round: int, cid: int | None = None, client_states: dict[int, ClientState] | None = None,
           server_steps_cumulative: int | None = None) -> EvaluateConfig:
        return EvaluateConfig(
            cid=cid, server_round=server_round, batch_size=int(llm.get("device_eval_batch_size", 1)),
            use_unigram_metrics=bool(fl["use_unigram_metrics"]),
            allow_unigram_metrics_failures=bool(fl["allow_unigram_metrics_failures"]), resize_vocab=fl.get("resize_vocab"),
            s3_comm_config=dict(cfg.get("s3_comm_config") or {}), frozen_layers=fl.get("frozen_layers"),
            unfrozen_layers=fl.get("unfrozen_layers"), split_eval=bool(fl.get("split_eval", False)),
            set_trainer_params_filter_keys=bool(fl.get("set_trainer_params_filter_keys", True)),
            set_trainer_key_to_filter=str(fl.get("set_trainer_key_to_filter", "transformer")),
            client_state={c: s.to_literal() for c, s in (client_states or {}).items()} or None,
            server_steps_cumulative=server_steps_cumulative)

    return fn
