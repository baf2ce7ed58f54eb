"""Reference module path for the unigram-normalised metrics (ref: photon/metrics/unigram_normalized_metrics.py).
The implementations live in ``photon_b200.metrics.language`` next to the base language metrics because all of them
are fed from the same fused cross-entropy by-products (loss sum, token count, unigram loss sum)."""
from photon_b200.metrics.language import (UNIGRAM_METRIC_NAMES_AND_CLASSES, PureUnigramCrossEntropy,  # noqa: F401
                                          PureUnigramPerplexity, UnigramNormalizedLanguageCrossEntropy,
                                          UnigramNormalizedLanguagePerplexity, create_wrapped_subclass, register_metric)
