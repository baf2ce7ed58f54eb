import json
import math

import torch

from photon_b200.dataset.utils import ByteTokenizer
from photon_b200.eval import EvalGauntlet, ICLEvaluator, expand_task


def _fake_logits(ids):
    """A 'model' that always predicts next byte = current byte + 1 (mod 257)."""
    B, S = ids.shape
    out = torch.full((B, S, 257), -10.0)
    out[torch.arange(B)[:, None], torch.arange(S)[None], (ids + 1) % 257] = 10.0
    return out


def test_icl_language_modeling_and_multiple_choice(tmp_path):
    tok = ByteTokenizer()
    (tmp_path / "lm.jsonl").write_text("\n".join(json.dumps(r) for r in [{"context": "\x1f", "continuation": "!"}, {"context": "xy", "continuation": "zz"}]))
    (tmp_path / "mc.jsonl").write_text(json.dumps({"query": "\x1f", "choices": ["!", "q"], "gold": 0}) + "\n" + json.dumps({"query": "\x1f", "choices": ["q", "!"], "gold": 1}))
    ev = ICLEvaluator(_fake_logits, tok, 64, str(tmp_path))
    lm = ev.evaluate_task({"label": "lm", "dataset_uri": "lm.jsonl", "icl_task_type": "language_modeling", "num_fewshot": [0]})
    assert lm["lm/0-shot/accuracy"] == 0.5            # 0x1f -> ' ' -> '!' is the +1 chain; 'xy zz' is not
    mc = ev.evaluate_task({"label": "mc", "dataset_uri": "mc.jsonl", "icl_task_type": "multiple_choice", "num_fewshot": 0})
    assert mc["mc/0-shot/accuracy"] == 1.0
    assert "skipped" in ev.evaluate_task({"label": "x", "dataset_uri": "nope.jsonl", "icl_task_type": "multiple_choice"})["status"]
    assert expand_task({"label": "a", "dataset_uri": "b", "icl_task_type": "schema", "num_fewshot": 3})["num_fewshot"] == [3]


def test_gauntlet_aggregation():
    g = EvalGauntlet({"weighting": "EQUAL", "subtract_random_baseline": True, "rescale_accuracy": True,
                      "averages": {"core": ["a", "b"]},
                      "categories": [{"name": "a", "benchmarks": [{"name": "t1", "num_fewshot": 0, "random_baseline": 0.25}]},
                                     {"name": "b", "benchmarks": [{"name": "t2", "num_fewshot": 5, "random_baseline": 0.0}]}]})
    out = g.aggregate({"t1/0-shot/accuracy": 0.625, "t2/5-shot/accuracy": 0.3})
    assert math.isclose(out["icl/metrics/eval_gauntlet/a"], 0.5) and math.isclose(out["icl/metrics/eval_gauntlet/core"], 0.4)
