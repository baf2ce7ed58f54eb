"""Language-model metrics as plain sum/count accumulators.

``LanguageCrossEntropy`` / ``LanguagePerplexity`` / ``TokenAccuracy`` are what
Composer/llm-foundry attach to ``mpt_causal_lm``; the four unigram-normalised
metrics reproduce the reference's own additions (ref: photon/metrics/
unigram_normalized_metrics.py:12-264): ``PureUnigramCrossEntropy`` is the CE of
the *unigram* distribution on the targets, ``UnigramNormalizedLanguageCrossEntropy``
is ``CE(model) − CE(unigram)``, and the perplexities are ``exp`` of those.
States are (sum, count) pairs → distributed sync is ONE tiny all-reduce of the
stacked states per evaluation instead of one per metric.
"""
from __future__ import annotations

import math
from typing import Any

import torch


class Metric:
    """Minimal metric protocol: ``update(**batch_stats)``, ``compute()``, ``state``."""

    name = "metric"

    def __init__(self) -> None:
        self.reset()

    def reset(self) -> None:
        self.sum = 0.0
        self.count = 0.0

    def state(self) -> list[float]:
        return [float(self.sum), float(self.count)]

    def load_state(self, st: list[float]) -> None:
        self.sum, self.count = float(st[0]), float(st[1])

    def update(self, stats: dict[str, Any]) -> None:
        raise NotImplementedError

    def compute(self) -> float:
        return self.sum / self.count if self.count else float("nan")


class LanguageCrossEntropy(Metric):
    name = "LanguageCrossEntropy"

    def update(self, stats: dict[str, Any]) -> None:
        self.sum += float(stats["loss_sum"])
        self.count += float(stats["n_tokens"])


class LanguagePerplexity(LanguageCrossEntropy):
    name = "LanguagePerplexity"

    def compute(self) -> float:
        ce = super().compute()
        return math.exp(ce) if ce == ce and ce < 700 else float("inf") if ce == ce else ce


class TokenAccuracy(Metric):
    name = "TokenAccuracy"

    def update(self, stats: dict[str, Any]) -> None:
        if "n_correct" in stats:
            self.sum += float(stats["n_correct"])
            self.count += float(stats["n_tokens"])


class PureUnigramCrossEntropy(Metric):
    name = "PureUnigramCrossEntropy"

    def update(self, stats: dict[str, Any]) -> None:
        if "unigram_loss_sum" in stats:
            self.sum += float(stats["unigram_loss_sum"])
            self.count += float(stats["n_tokens"])


class PureUnigramPerplexity(PureUnigramCrossEntropy):
    name = "PureUnigramPerplexity"

    def compute(self) -> float:
        ce = super().compute()
        return math.exp(ce) if ce == ce else ce


class UnigramNormalizedLanguageCrossEntropy(Metric):
    name = "UnigramNormalizedLanguageCrossEntropy"

    def update(self, stats: dict[str, Any]) -> None:
        if "unigram_loss_sum" in stats:
            self.sum += float(stats["loss_sum"]) - float(stats["unigram_loss_sum"])
            self.count += float(stats["n_tokens"])


class UnigramNormalizedLanguagePerplexity(UnigramNormalizedLanguageCrossEntropy):
    name = "UnigramNormalizedLanguagePerplexity"

    def compute(self) -> float:
        ce = super().compute()
        return math.exp(ce) if ce == ce else ce


UNIGRAM_METRIC_NAMES_AND_CLASSES = {
    c.name: c for c in (PureUnigramCrossEntropy, PureUnigramPerplexity,
                        UnigramNormalizedLanguageCrossEntropy, UnigramNormalizedLanguagePerplexity)
}
BASE_METRICS = {c.name: c for c in (LanguageCrossEntropy, LanguagePerplexity, TokenAccuracy)}


def create_wrapped_subclass(base_class: type, **kwargs: Any) -> type:
    """A same-named subclass of ``base_class`` whose constructor has ``kwargs`` pre-bound, so a metric that needs
    arguments (e.g. a custom unigram table) can sit in a name → class registry that instantiates with no arguments
    (ref: photon/metrics/unigram_normalized_metrics.py:233-256). Call-time kwargs override the bound ones."""
    def __init__(self: Any, **init_kwargs: Any) -> None:
        base_class.__init__(self, **{**kwargs, **init_kwargs})

    return type(base_class.__name__, (base_class,), {"__init__": __init__, "__module__": base_class.__module__,
                                                    "__doc__": base_class.__doc__})


def register_metric(cls: type, *, unigram: bool = False, **bound_kwargs: Any) -> type:
    """Add a ``Metric`` subclass to the registry ``build_metrics`` instantiates from."""
    wrapped = create_wrapped_subclass(cls, **bound_kwargs) if bound_kwargs else cls
    (UNIGRAM_METRIC_NAMES_AND_CLASSES if unigram else BASE_METRICS)[getattr(cls, "name", cls.__name__)] = wrapped
    return wrapped


def build_metrics(use_unigram: bool = False) -> dict[str, Metric]:
    out = {n: c() for n, c in BASE_METRICS.items()}
    if use_unigram:
        out.update({n: c() for n, c in UNIGRAM_METRIC_NAMES_AND_CLASSES.items()})
    return out


def unigram_log_probs(freq: dict[int, int] | dict[str, int], vocab_size: int, device: Any = None) -> torch.Tensor:
    """``log p_unigram[token]`` from a ``1_gram.json`` frequency map; unseen ids get the
    mass of a single pseudo-count so the CE stays finite
    (ref: photon/utils.py:1039-1063 ``get_unigram_probabilities_tensor``)."""
    counts = torch.ones(vocab_size, dtype=torch.float64)
    for k, v in freq.items():
        i = int(k)
        if 0 <= i < vocab_size:
            counts[i] += float(v)
    return (counts / counts.sum()).log().to(torch.float32).to(device or "cpu")


def unigram_loss_sum(targets: torch.Tensor, log_probs: torch.Tensor) -> torch.Tensor:
    """Σ −log p_unigram(target) over non-ignored targets."""
    valid = targets != -100
    return -(log_probs[targets.clamp(min=0)] * valid).sum()


def sync_metrics(metrics: dict[str, Metric], group: Any = None) -> None:
    """One all-reduce(SUM) over the stacked (sum,count) states of every metric."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    names = sorted(metrics)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([x for n in names for x in metrics[n].state()], dtype=torch.float64, device=dev)
    dist.all_reduce(t, group=group)
    vals = t.cpu().tolist()
    for i, n in enumerate(names):
        metrics[n].load_state(vals[2 * i: 2 * i + 2])
