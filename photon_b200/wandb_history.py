"""Round-level metric history (+ optional wandb mirroring).

``History`` keeps what Flower's did (``losses_distributed/centralized``,
``metrics_distributed_fit / distributed / centralized`` as ``{key: [(round, value)…]}``);
``WandbHistory`` mirrors every ``add_*`` to ``wandb.log(step=server_round)`` when a run is
active (ref: photon/wandb_history.py:27-70). It is pickled into ``state.bin`` by the
server checkpoint (ref: photon/server/s3_utils.py:374-383).
"""
from __future__ import annotations

from typing import Any


class History:
    def __init__(self) -> None:
        self.losses_distributed: list[tuple[int, float]] = []
        self.losses_centralized: list[tuple[int, float]] = []
        self.metrics_distributed_fit: dict[str, list[tuple[int, Any]]] = {}
        self.metrics_distributed: dict[str, list[tuple[int, Any]]] = {}
        self.metrics_centralized: dict[str, list[tuple[int, Any]]] = {}

    def add_loss_distributed(self, server_round: int, loss: float) -> None:
        self.losses_distributed.append((server_round, loss))

    def add_loss_centralized(self, server_round: int, loss: float) -> None:
        self.losses_centralized.append((server_round, loss))

    @staticmethod
    def _add(store: dict[str, list[tuple[int, Any]]], server_round: int, metrics: dict[str, Any]) -> None:
        for k, v in metrics.items():
            store.setdefault(k, []).append((server_round, v))

    def add_metrics_distributed_fit(self, server_round: int, metrics: dict[str, Any]) -> None:
        self._add(self.metrics_distributed_fit, server_round, metrics)

    def add_metrics_distributed(self, server_round: int, metrics: dict[str, Any]) -> None:
        self._add(self.metrics_distributed, server_round, metrics)

    def add_metrics_centralized(self, server_round: int, metrics: dict[str, Any]) -> None:
        self._add(self.metrics_centralized, server_round, metrics)

    def latest(self, key: str) -> Any:
        for store in (self.metrics_distributed_fit, self.metrics_distributed, self.metrics_centralized):
            if key in store:
                return store[key][-1][1]
        return None


class WandbHistory(History):
    def __init__(self, use_wandb: bool = False) -> None:
        super().__init__()
        self.use_wandb = use_wandb

    def _log(self, server_round: int, metrics: dict[str, Any]) -> None:
        if not self.use_wandb:
            return
        try:
            import wandb  # type: ignore[import-not-found]

            if wandb.run is not None:
                wandb.log({k: v for k, v in metrics.items() if isinstance(v, (int, float))}, step=server_round)
        except Exception:  # noqa: BLE001
            self.use_wandb = False

    def add_loss_distributed(self, server_round: int, loss: float) -> None:
        super().add_loss_distributed(server_round, loss)
        self._log(server_round, {"distributed_loss": loss})

    def add_loss_centralized(self, server_round: int, loss: float) -> None:
        super().add_loss_centralized(server_round, loss)
        self._log(server_round, {"centralized_loss": loss})

    def add_metrics_distributed_fit(self, server_round: int, metrics: dict[str, Any]) -> None:
        super().add_metrics_distributed_fit(server_round, metrics)
        self._log(server_round, metrics)

    def add_metrics_distributed(self, server_round: int, metrics: dict[str, Any]) -> None:
        super().add_metrics_distributed(server_round, metrics)
        self._log(server_round, metrics)

    def add_metrics_centralized(self, server_round: int, metrics: dict[str, Any]) -> None:
        super().add_metrics_centralized(server_round, metrics)
        self._log(server_round, metrics)

    def __getstate__(self) -> dict[str, Any]:
        return dict(self.__dict__)
