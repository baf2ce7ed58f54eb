from photon_b200.metrics.language import (BASE_METRICS, UNIGRAM_METRIC_NAMES_AND_CLASSES, LanguageCrossEntropy,
                                          LanguagePerplexity, Metric, TokenAccuracy, build_metrics, sync_metrics,
                                          unigram_log_probs, unigram_loss_sum)
