"""Server optimizers: FedAvgEfficient, FedNesterov, FedMom, FedAdam, FedYogi.

Math per SURVEY §2.4 (ref: photon/strategy/fedavg_eff.py:312-330,
fednestorov.py:313-334, fedmom.py:259-279, fedadam.py:291-318,
fedyogi.py:295-322) on FLAT fp32 tensors: ``x`` global model, ``a`` the
sample-weighted mean of the returned client models (optionally scaled by
``scaling_fn(K)``), ``pg = x − a`` the pseudo-gradient.

Two execution paths share these classes:

* **host/oracle** — the torch expressions in :func:`server_opt_step` (CPU or a
  single CUDA device); also the numerics oracle for the kernel tests;
* **nvl** — ``photon_b200.parallel.fed_round`` runs the same update inside the
  fused reduce+optimizer kernel (``csrc/fed_round.cu``) on each GPU's shard and
  only hands the per-tensor norm partials back to this class for metrics.

Sign convention: FedAdam/FedYogi in the reference step ``x + η·m̂/(√v̂+τ)`` with
``pg = x − a`` — an *ascent* direction (SURVEY §2.4 "sign quirk").  Default here
is the descent form; ``reference_sign_compat=True`` reproduces the reference
bit-for-bit semantics.
"""
from __future__ import annotations

import math
from typing import Any, Callable, Iterable

import torch

from photon_b200.messages import Code, EvaluateRes, FitRes
from photon_b200.strategy.aggregation import (aggregate_inplace, naive_weighted_mean, weighted_average,
                                             weighted_loss_avg)
from photon_b200.strategy.metrics import ServerMetricCallback
from photon_b200.utils.flat import FlatLayout

KINDS = ("fedavg", "nesterov", "fedmom", "fedadam", "fedyogi")

# strategy state keys == the stems of the server checkpoint files (part of the on-disk compatibility surface; the reference keeps
# them in photon/strategy/constants.py as MODEL_PARAMETERS / FIRST_MOMENTUM / SECOND_MOMENTUM — aliases below)
SERVER_PARAMETERS_KEY, MOMENTUM_KEY, SECOND_MOMENTUM_KEY = (f"current_{stem}" for stem in ("server_parameters", "momentum_vector", "second_momentum_vector"))
MODEL_PARAMETERS, FIRST_MOMENTUM, SECOND_MOMENTUM = SERVER_PARAMETERS_KEY, MOMENTUM_KEY, SECOND_MOMENTUM_KEY


def server_opt_step(kind: str, x: torch.Tensor, a: torch.Tensor, m: torch.Tensor | None, v: torch.Tensor | None,
                    hp: dict[str, float], t: int, sign_compat: bool = False) -> torch.Tensor:
    """In-place server update of ``x`` (and ``m``/``v``); returns the pseudo-gradient.

    ``hp``: ``lr``, ``mu`` for fedavg/nesterov/fedmom; ``eta, beta1, beta2, tau`` for
    fedadam/fedyogi. ``t`` is the 1-based server round (bias correction)."""
    pg = x - a
    if kind == "fedavg":
        x.sub_(pg, alpha=hp["lr"])
    elif kind == "nesterov":  # torch.optim.SGD(nesterov=True) form
        assert m is not None
        m.mul_(hp["mu"]).add_(pg)
        x.sub_(pg + hp["mu"] * m, alpha=hp["lr"])
    elif kind == "fedmom":
        assert m is not None
        v_new = x - hp["lr"] * pg
        x.copy_((1.0 + hp["mu"]) * v_new - hp["mu"] * m)
        m.copy_(v_new)
    elif kind in ("fedadam", "fedyogi"):
        assert m is not None and v is not None
        b1, b2 = hp["beta1"], hp["beta2"]
        m.mul_(b1).add_(pg, alpha=1.0 - b1)
        g2 = pg * pg
        if kind == "fedadam":
            v.mul_(b2).add_(g2, alpha=1.0 - b2)
        else:  # yogi: additive, sign-controlled second moment
            v.add_((1.0 - b2) * g2 * torch.sign(g2 - v))
        m_hat = m / (1.0 - b1 ** t)
        v_hat = v / (1.0 - b2 ** t)
        step = hp["eta"] * m_hat / (v_hat.sqrt() + hp["tau"])
        if sign_compat:
            x.add_(step)   # reference behaviour: ascends along pg = x − a
        else:
            x.sub_(step)   # descent: equivalent to Δ = a − x with x + η·…
    else:
        raise ValueError(f"unknown server optimizer '{kind}'")
    return pg


class ServerStrategy:
    """Common machinery: streaming aggregation → optimizer → norms/metrics."""

    kind = "fedavg"
    n_moments = 0

    def __init__(self, *, n_clients_per_round: int = 1, track_norms: bool = True,
                 track_inplace_aggregation: bool = False, scaling_fn: str | None = None,
                 metrics_callback: ServerMetricCallback | None = None, reference_sign_compat: bool = False,
                 fit_metrics_aggregation_fn: Callable[..., dict[str, Any]] = weighted_average,
                 evaluate_metrics_aggregation_fn: Callable[..., dict[str, Any]] = weighted_average,
                 **hyper: float) -> None:
        if scaling_fn not in (None, "linear", "sqrt"):
            raise ValueError("Scaling function must be either 'linear' or 'sqrt'.")
        self.min_fit_clients = int(n_clients_per_round)
        self._scaling = scaling_fn
        self.track_norms, self.track_inplace = track_norms, track_inplace_aggregation
        self.metrics_callback = metrics_callback
        self.sign_compat = reference_sign_compat
        self.fit_metrics_aggregation_fn = fit_metrics_aggregation_fn
        self.evaluate_metrics_aggregation_fn = evaluate_metrics_aggregation_fn
        self.hp: dict[str, float] = {k: float(v) for k, v in hyper.items()}
        self.parameters: torch.Tensor | None = None
        self.momentum_vector: torch.Tensor | None = None
        self.second_momentum_vector: torch.Tensor | None = None
        self.layout: FlatLayout | None = None
        self.last_pg_sq: float | None = None   # ‖x − x̄‖² of the last server update (kept when a metrics callback is set)

    # -- state ----------------------------------------------------------------------
    @property
    def state_keys(self) -> list[str]:
        return [SERVER_PARAMETERS_KEY, MOMENTUM_KEY, SECOND_MOMENTUM_KEY][: 1 + self.n_moments]

    def state_tensors(self) -> dict[str, torch.Tensor]:
        vals = [self.parameters, self.momentum_vector, self.second_momentum_vector]
        return {k: t for k, t in zip(self.state_keys, vals) if t is not None}

    def initialize(self, parameters: torch.Tensor, momentum: torch.Tensor | None = None,
                   second_momentum: torch.Tensor | None = None, layout: FlatLayout | None = None) -> None:
        """Inject params + momenta (ref: photon/strategy/utils.py:31-72). Missing momenta → zeros
        (the reference's ``initialize_round`` overrides FedMom's copy-of-x ctor value with zeros,
        init_utils.py:190-193; we keep the zeros)."""
        self.parameters = parameters
        self.layout = layout
        if self.n_moments >= 1:
            self.momentum_vector = momentum if momentum is not None else torch.zeros_like(parameters)
        if self.n_moments >= 2:
            self.second_momentum_vector = second_momentum if second_momentum is not None else torch.zeros_like(parameters)

    def scaling_factor(self, k: int | None = None) -> float:
        k = self.min_fit_clients if k is None else k
        return 1.0 if self._scaling is None else (float(k) if self._scaling == "linear" else math.sqrt(k))

    # -- fit ------------------------------------------------------------------------
    def aggregate_fit(self, server_round: int, results: Iterable[FitRes], failures: list[Any] | None = None
                      ) -> tuple[torch.Tensor | None, dict[str, Any]]:
        """Consume a (lazy) iterable of successful ``FitRes`` whose ``parameters.data`` is a
        flat tensor in this strategy's layout. Returns (new global params | None, metrics)."""
        assert self.parameters is not None, "strategy not initialised"
        metrics: dict[str, Any] = {}
        cb = self.metrics_callback
        if cb is not None:
            cb.round_start(self.parameters.clone(), server_round)
        kept: list[tuple[torch.Tensor, float]] = []
        fit_metrics: list[tuple[int, dict[str, Any]]] = []

        def stream() -> Iterable[tuple[torch.Tensor, float]]:
            for r in results:
                flat = r.parameters.data if r.parameters is not None else None
                if flat is None:
                    continue
                fit_metrics.append((r.num_examples, r.metrics))
                if self.track_inplace:
                    kept.append((flat.clone(), float(r.num_examples)))
                yield flat, float(r.num_examples)

        avg, _total, k = aggregate_inplace(stream(), cb.per_client if cb is not None else None)
        if avg is None:
            return None, {}
        avg = avg.to(self.parameters.device)
        s = self.scaling_factor()
        if s != 1.0:
            avg.mul_(s)
        if cb is not None:
            cb.round_end(avg, metrics)
        metrics.update(self.apply_server_update(avg, server_round))
        if self.track_inplace and kept:
            gap = (naive_weighted_mean(kept).to(avg.device) * s - avg).norm()
            metrics["server/l2_norm_fedavg_gap"] = float(gap)
        if self.fit_metrics_aggregation_fn is not None and fit_metrics:
            metrics.update(self.fit_metrics_aggregation_fn(fit_metrics))
        metrics["server/n_aggregated_clients"] = k
        return self.parameters, metrics

    def apply_server_update(self, avg: torch.Tensor, server_round: int) -> dict[str, Any]:
        """Host/oracle path of the server optimizer + the reference's norm metrics."""
        assert self.parameters is not None
        pg = server_opt_step(self.kind, self.parameters, avg, self.momentum_vector, self.second_momentum_vector,
                             self.hp, max(1, server_round), self.sign_compat)
        if self.metrics_callback is not None:   # the round transports feed the noise-scale estimate from this
            flat = pg.reshape(-1)
            self.last_pg_sq = float(torch.dot(flat, flat))
        return self.norm_metrics(pg, avg) if self.track_norms else {}

    def norm_metrics(self, pg: torch.Tensor, avg: torch.Tensor) -> dict[str, Any]:
        planes = {"pseudo_gradient": pg, "fedavg_result": avg, "model": self.parameters}
        if self.momentum_vector is not None:
            planes["momentum_vector"] = self.momentum_vector
        if self.second_momentum_vector is not None:
            planes["second_momentum_vector"] = self.second_momentum_vector
        out: dict[str, Any] = {}
        for name, t in planes.items():
            if self.layout is not None:
                sq = [float(torch.dot(w := self.layout.view(t, i).reshape(-1), w)) for i in range(len(self.layout.names))]
                for i, q in enumerate(sq):
                    out[f"server/layer/{i}/l2_norm_{name}"] = math.sqrt(q)
                out[f"server/l2_norm_{name}"] = math.sqrt(sum(sq))
            else:
                out[f"server/l2_norm_{name}"] = float(t.norm())
        return out

    def metrics_from_partials(self, sq_partials: dict[str, torch.Tensor]) -> dict[str, Any]:
        """Norm metrics from the fused kernel's per-tensor Σx² by-products (nvl path)."""
        out: dict[str, Any] = {}
        for name, sq in sq_partials.items():
            vals = sq.double().cpu().tolist()
            for i, q in enumerate(vals):
                out[f"server/layer/{i}/l2_norm_{name}"] = math.sqrt(max(q, 0.0))
            out[f"server/l2_norm_{name}"] = math.sqrt(max(sum(vals), 0.0))
        return out

    # -- evaluate ---------------------------------------------------------------------
    def aggregate_evaluate(self, server_round: int, results: Iterable[EvaluateRes], failures: list[Any] | None = None
                           ) -> tuple[float | None, dict[str, Any]]:
        res = [r for r in results if r.status.code == Code.OK]
        if not res:
            return None, {}
        loss = weighted_loss_avg([(r.num_examples, r.loss) for r in res])
        metrics = self.evaluate_metrics_aggregation_fn([(r.num_examples, r.metrics) for r in res]) \
            if self.evaluate_metrics_aggregation_fn else {}
        return loss, metrics


class FedAvgEfficient(ServerStrategy):
    kind, n_moments = "fedavg", 0

    def __init__(self, server_learning_rate: float = 0.7, **kw: Any) -> None:
        super().__init__(lr=server_learning_rate, **kw)


class FedNesterov(ServerStrategy):
    kind, n_moments = "nesterov", 1

    def __init__(self, server_learning_rate: float = 0.7, server_momentum: float = 0.9, **kw: Any) -> None:
        super().__init__(lr=server_learning_rate, mu=server_momentum, **kw)


class FedMom(ServerStrategy):
    kind, n_moments = "fedmom", 1

    def __init__(self, server_learning_rate: float = 0.7, server_momentum: float = 0.9, **kw: Any) -> None:
        super().__init__(lr=server_learning_rate, mu=server_momentum, **kw)


class FedAdam(ServerStrategy):
    kind, n_moments = "fedadam", 2

    def __init__(self, eta: float = 0.1, beta_1: float = 0.9, beta_2: float = 0.95, tau: float = 1e-9, **kw: Any) -> None:
        super().__init__(eta=eta, beta1=beta_1, beta2=beta_2, tau=tau, **kw)


class FedYogi(ServerStrategy):
    kind, n_moments = "fedyogi", 2

    def __init__(self, eta: float = 1e-2, beta_1: float = 0.9, beta_2: float = 0.99, tau: float = 1e-3, **kw: Any) -> None:
        super().__init__(eta=eta, beta1=beta_1, beta2=beta_2, tau=tau, **kw)


# the reference's name for the common base (Flower ``FedAvg`` + the config-function plumbing, ref: photon/strategy/strategy_with_cfg.py:88-162):
# here ``ServerStrategy`` carries the fit / evaluate metric aggregation, the failure policy lives in ``server/fit_utils.py`` and the
# per-round config functions in ``clients/configs.py`` (``get_photon_fit_config_fn`` / ``get_photon_evaluate_config_fn``)
FedAvgWithConfig = ServerStrategy
