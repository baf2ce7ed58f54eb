"""Server-side metric callbacks (ref: photon/strategy/metrics.py:53-267).

``FedSimpleNoiseScale`` = McCandlish et al. "simple noise scale" with
B_small = 1 client and B_big = K clients, EMA-smoothed with de-biasing.
Unlike the reference (which re-instantiates the callback every round, so its
EMA never accumulates), the instance here lives as long as the strategy.
"""
from __future__ import annotations

from typing import Any

import torch


def ema_with_debias(prev: float, beta: float, value: float, step: int) -> tuple[float, float]:
    """Returns (new running value, de-biased estimate)."""
    new = beta * prev + (1.0 - beta) * value
    return new, new / (1.0 - beta ** step)


class ServerMetricCallback:
    def round_start(self, global_params: torch.Tensor, server_round: int) -> None: ...
    def per_client(self, client_params: torch.Tensor, num_samples: float) -> None: ...
    def round_end(self, fedavg_result: torch.Tensor | None, metrics: dict[str, Any]) -> None: ...


class FedSimpleNoiseScale(ServerMetricCallback):
    def __init__(self, beta: float = 0.99) -> None:
        self.beta = float(beta)
        self.run_trace = self.run_sq = self.run_ns = 0.0
        self.counter = 0
        self._sum_sq = 0.0
        self._n = 0
        self._old: torch.Tensor | None = None

    def round_start(self, global_params: torch.Tensor, server_round: int) -> None:
        self._old, self._sum_sq, self._n = global_params, 0.0, 0

    def per_client(self, client_params: torch.Tensor, num_samples: float) -> None:
        assert self._old is not None
        d = self._old.to(client_params.device) - client_params
        self._sum_sq += float(torch.dot(d, d))
        self._n += 1

    def round_end(self, fedavg_result: torch.Tensor | None, metrics: dict[str, Any]) -> None:
        if fedavg_result is None or self._old is None:
            return
        pg = self._old.to(fedavg_result.device) - fedavg_result
        self.round_end_from_stats(self._sum_sq, self._n, float(torch.dot(pg, pg)), metrics)

    def round_end_from_stats(self, sum_sq: float, n: int, g_big: float, metrics: dict[str, Any]) -> None:
        """The same estimate from additive statistics — Σ_k ‖x − x_k‖², K and ‖x − x̄‖² — which is what the
        distributed round transports have (each rank sees only its own clients; two scalars are all-reduced)."""
        if sum_sq == 0.0 or n < 2:
            return
        g_small = sum_sq / n
        b_small, b_big = 1, int(n)
        trace = (g_small - g_big) / (1.0 / b_small - 1.0 / b_big)
        sq = (b_big * g_big - b_small * g_small) / (b_big - b_small)
        self.counter += 1
        self.run_trace, scale = ema_with_debias(self.run_trace, self.beta, trace, self.counter)
        self.run_sq, noise = ema_with_debias(self.run_sq, self.beta, sq, self.counter)
        raw = trace / sq if sq else float("nan")
        self.run_ns, ns_debiased = ema_with_debias(self.run_ns, self.beta, raw, self.counter)
        metrics.update({
            "noise_scale/b_small": b_small, "noise_scale/b_big": b_big,
            "noise_scale/g_small_l2norm_squared": g_small, "noise_scale/g_big_l2norm_squared": g_big,
            "noise_scale/trace_estimate": trace, "noise_scale/squared_gradients_estimate": sq,
            "noise_scale/noise_scale_with_emas": scale / noise if noise else float("nan"),
            "noise_scale/noise_scale_ema": self.run_ns, "noise_scale/noise_scale_ema_bias": ns_debiased,
            "noise_scale/noise_scale_raw": raw,
        })
        self._sum_sq, self._n = 0.0, 0
