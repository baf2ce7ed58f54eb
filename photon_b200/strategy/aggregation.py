"""Streaming (in-place) weighted aggregation of client results.

Same arithmetic as the reference's NumPy path — running mean
``acc *= N_prev/N_new; acc += cur * n/N_new`` consumed lazily as replies arrive
(ref: photon/strategy/aggregation.py:57-118) — but on flat torch buffers, so
one fused op per client instead of 148 per-layer NumPy calls, and the same
code runs on a CUDA shard.  ``weighted_average`` merges metric dicts and the
stringly ``client_state_acc`` field (ref: aggregation.py:172-211).
"""
from __future__ import annotations

import ast
from typing import Any, Callable, Iterable

import torch


class StreamingMean:
    """Running sample-weighted mean over equally-shaped flat tensors."""

    def __init__(self) -> None:
        self.acc: torch.Tensor | None = None
        self.total = 0.0
        self.count = 0

    def add(self, cur: torch.Tensor, weight: float) -> None:
        if weight <= 0:
            raise ValueError("client weight (num_examples) must be positive")
        if self.acc is None:
            self.acc = cur.detach().to(torch.float32).clone()
            self.total = float(weight)
        else:
            new_total = self.total + float(weight)
            self.acc.mul_(self.total / new_total).add_(cur.to(self.acc.device, torch.float32), alpha=float(weight) / new_total)
            self.total = new_total
        self.count += 1

    def result(self) -> torch.Tensor | None:
        return self.acc


def aggregate_inplace(results: Iterable[tuple[torch.Tensor, float]],
                      per_client_hook: Callable[[torch.Tensor, float], None] | None = None) -> tuple[torch.Tensor | None, float, int]:
    """Reduce a lazy iterable of ``(flat_params, num_examples)``; returns (mean, Σn, K)."""
    sm = StreamingMean()
    for flat, n in results:
        if per_client_hook is not None:
            per_client_hook(flat, n)
        sm.add(flat, n)
    return sm.result(), sm.total, sm.count


def aggregate_parameters(acc: torch.Tensor | None, cur: torch.Tensor, n_prev: float, n_cur: float) -> tuple[torch.Tensor, float]:
    """One step of the running mean, functional form: ``acc ← acc·N_prev/N_new + cur·n/N_new``
    (ref: photon/strategy/aggregation.py:19-87). ``acc`` is updated in place when given."""
    if n_cur <= 0:
        raise ValueError("client weight (num_examples) must be positive")
    if acc is None:
        return cur.detach().to(torch.float32).clone(), float(n_cur)
    n_new = float(n_prev) + float(n_cur)
    acc.mul_(float(n_prev) / n_new).add_(cur.to(acc.device, torch.float32), alpha=float(n_cur) / n_new)
    return acc, n_new


def aggregate_cumulative_average(fit_results: Iterable[Any], metrics_callback: Any = None) -> torch.Tensor | None:
    """Lazy in-place weighted average over ``FitRes``-like objects (``.parameters`` flat tensor, ``.num_examples``);
    only one client payload is alive at a time (ref: aggregation.py:121-150)."""
    hook = None
    if metrics_callback is not None:
        hook = lambda flat, n: metrics_callback.per_client(flat, n)  # noqa: E731
    mean, _, _ = aggregate_inplace(((r.parameters, r.num_examples) for r in fit_results), per_client_hook=hook)
    return mean


def parameters_to_ndarrays_gen(flat: torch.Tensor, layout: Any) -> Any:
    """Yield one per-tensor ndarray VIEW at a time from a flat payload — the lazy per-layer decode of
    ref aggregation.py:153-169, without the ``np.load`` of serialized bytes (payloads are never serialized here)."""
    host = flat.detach().cpu() if flat.is_cuda else flat.detach()
    for off, ne, shp in zip(layout.offsets, layout.numels, layout.shapes):
        yield host[off:off + ne].view(shp).numpy()


def naive_weighted_mean(results: list[tuple[torch.Tensor, float]]) -> torch.Tensor:
    """Textbook Σ n_k x_k / Σ n_k in float64 — the oracle for ``track_inplace_aggregation``
    (ref: photon/strategy/fedavg_eff.py:366-391 ``server/l2_norm_fedavg_gap``)."""
    tot = sum(n for _, n in results)
    acc = torch.zeros_like(results[0][0], dtype=torch.float64)
    for x, n in results:
        acc += x.to(torch.float64) * (n / tot)
    return acc.to(torch.float32)


def weighted_average(metrics: list[tuple[int, dict[str, Any]]]) -> dict[str, Any]:
    """Sample-weighted mean of scalar metrics; ``client_state_acc`` dict-strings are merged."""
    if not metrics:
        return {}
    total = float(sum(n for n, _ in metrics)) or 1.0
    out: dict[str, Any] = {}
    merged_state: dict[Any, Any] = {}
    keys: list[str] = []
    for _, m in metrics:
        for k in m:
            if k not in keys:
                keys.append(k)
    for k in keys:
        if k == "client_state_acc":
            for _, m in metrics:
                if k in m:
                    v = m[k]
                    merged_state.update(ast.literal_eval(v) if isinstance(v, str) else v)
            out[k] = str(merged_state)
            continue
        num, den = 0.0, 0.0
        for n, m in metrics:
            v = m.get(k)
            if isinstance(v, bool) or not isinstance(v, (int, float)):
                continue
            num += n * float(v)
            den += n
        if den:
            out[k] = num / den if den != total else num / total
    return out


def weighted_loss_avg(results: list[tuple[int, float]]) -> float:
    tot = sum(n for n, _ in results)
    return sum(n * l for n, l in results) / tot if tot else 0.0
